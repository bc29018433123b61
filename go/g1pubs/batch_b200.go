// +build b200

// batch_b200.go -- package g1pubs (pubkeys in G1, signatures in G2) under `-tags b200`.  The exported API of
// g1pubs/bls.go is unchanged: Verify / VerifyWithDomain / VerifyAggregateCommon* keep calling bls.CompareTwoPairings,
// which the tag routes to the GPU engine (go/bls/pairing_b200.go).  Three functions of bls.go are REPLACED here under the
// tag -- AggregatePublicKeys, AggregateSignatures, (*Signature).VerifyAggregate: the maintainer moves those three, unchanged,
// from bls.go into a file `aggregate_ref.go` that starts with `// +build !b200` (INTEGRATION.md section 2) -- and the batch
// entry points the reference lacks are added.
package g1pubs

import (
	"github.com/phoreproject/bls"
)

// AggregatePublicKeys replaces g1pubs/bls.go:192-198 (serial fold of G1Projective.Add) with one device reduction; the
// result is normalised and Equal()s the fold.
func AggregatePublicKeys(p []*PublicKey) *PublicKey {
	aff := make([]bls.G1Affine, len(p))
	for i, pk := range p {
		aff[i] = *pk.p.ToAffine()
	}
	return &PublicKey{p: bls.SumG1(aff)}
}

// AggregateSignatures replaces g1pubs/bls.go:177-183.
func AggregateSignatures(s []*Signature) *Signature {
	aff := make([]bls.G2Affine, len(s))
	for i, sig := range s {
		aff[i] = *sig.s.ToAffine()
	}
	return &Signature{s: bls.SumG2(aff)}
}

// VerifyAggregate replaces g1pubs/bls.go:252-282: same duplicate-message rule (incl. quirk Q7: an
// empty message is rejected), but ONE product of n+1 Miller loops and one final exponentiation
// instead of n+1 full pairings.  e(G1, sig) == prod e(pk_i, H(m_i))  <=>
// FE(ML(-G1, sig) * prod ML(pk_i, H(m_i))) == 1.
// An infinite signature or key makes the check false (the reference panics in MillerLoop, pairing.go:17-26).
func (s *Signature) VerifyAggregate(pubKeys []*PublicKey, msgs [][]byte) bool {
	if len(pubKeys) != len(msgs) {
		return false
	}
	if s.s.IsZero() {
		return false
	}
	for _, pk := range pubKeys {
		if pk.p.IsZero() {
			return false
		}
	}
	if hasDuplicates(msgs) { // the sort + dedupe of bls.go:257-273, unchanged
		return false
	}
	n := len(pubKeys)
	p := make([]bls.G1Affine, n+1)
	q := make([]bls.G2Affine, n+1)
	g := bls.G1AffineOne.Copy()
	g.NegAssign()
	p[0], q[0] = *g, *s.s.ToAffine()
	hs := bls.HashG2Batch(msgs) // HashG2 of every message on the device (go/bls/codec_b200.go)
	for i := range pubKeys {
		p[i+1], q[i+1] = *pubKeys[i].p.ToAffine(), hs[i]
	}
	return bls.PairingProductsAreOne(p, q, []uint32{0, uint32(n + 1)})[0]
}

// VerifyBatchCommonWithDomain verifies many (aggregate signature, committee, message) triples in one
// launch: the Ethereum-beacon shape of BASELINE config 5.  ok[i] ==
// sigs[i].VerifyAggregateCommonWithDomain(committees[i], msgs[i], domain)  (g1pubs/bls.go:294-297).
// An empty committee, keys that cancel to infinity or an infinite signature give false.
func VerifyBatchCommonWithDomain(sigs []*Signature, committees [][]*PublicKey, msgs [][32]byte, domain [8]byte) []bool {
	n := len(sigs)
	if n == 0 {
		return []bool{}
	}
	valid := make([]bool, n)
	p := make([]bls.G1Affine, 0, 2*n)
	q := make([]bls.G2Affine, 0, 2*n)
	off := make([]uint32, 1, n+1)
	hs := bls.HashG2WithDomainBatch(msgs, domain) // one launch for all message points
	for i := range sigs {
		aggp := AggregatePublicKeys(committees[i]).p
		valid[i] = len(committees[i]) > 0 && !aggp.IsZero() && !sigs[i].s.IsZero()
		agg := aggp.ToAffine()
		agg.NegAssign()
		p = append(p, *bls.G1AffineOne, *agg)
		q = append(q, *sigs[i].s.ToAffine(), hs[i])
		off = append(off, uint32(len(p)))
	}
	ok := bls.PairingProductsAreOne(p, q, off)
	for i := range ok {
		ok[i] = ok[i] && valid[i]
	}
	return ok
}

// VerifyWithDomainBatch verifies n independent wire-format (public key, message hash, signature) triples with one
// call; deserialisation, subgroup checks, hashing to G2 and the pairing checks all run on the device.
// ok[i] == VerifyWithDomain(msgs[i], DeserializePublicKey(pubs[i]), DeserializeSignature(sigs[i]), domain), false
// where either Deserialize* would have returned an error (g1pubs/bls.go:38-58,91-111,171-174).
func VerifyWithDomainBatch(pubs [][48]byte, msgs [][32]byte, domain [8]byte, sigs [][96]byte) []bool {
	return bls.VerifyWithDomainWire(pubs, msgs, domain, sigs)
}

// DeserializePublicKeys is DeserializePublicKey (g1pubs/bls.go:91-98) over a batch.
func DeserializePublicKeys(b [][48]byte) ([]*PublicKey, []error) {
	aff, errs := bls.DecompressG1Batch(b, true)
	out := make([]*PublicKey, len(b))
	for i := range aff {
		if errs[i] == nil {
			out[i] = &PublicKey{p: aff[i].ToProjective()}
		}
	}
	return out, errs
}

func hasDuplicates(msgs [][]byte) bool {
	cp := make([][]byte, len(msgs))
	for i, m := range msgs {
		cp[i] = append([]byte(nil), m...)
	}
	sorted := sortByteArrays(cp)
	last := []byte(nil)
	for _, m := range sorted {
		if bytesEqual(m, last) {
			return true
		}
		last = m
	}
	return false
}

func bytesEqual(a, b []byte) bool {
	if len(a) != len(b) {
		return false
	}
	for i := range a {
		if a[i] != b[i] {
			return false
		}
	}
	return true
}

// VerifyBatch verifies n independent wire-format (public key, message, signature) triples: ok[i] ==
// Verify(msgs[i], DeserializePublicKey(pubs[i]), DeserializeSignature(sigs[i])) (g1pubs/bls.go:38-58,91-111,165-168).
func VerifyBatch(pubs [][48]byte, msgs [][]byte, sigs [][96]byte) []bool {
	p := make([]byte, 0, 48*len(pubs))
	s := make([]byte, 0, 96*len(sigs))
	for i := range pubs {
		p = append(p, pubs[i][:]...)
		s = append(s, sigs[i][:]...)
	}
	return bls.VerifyWire(false, p, msgs, s)
}
