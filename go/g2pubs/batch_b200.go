// +build b200

// batch_b200.go -- package g2pubs (pubkeys in G2, signatures in G1) under `-tags b200`.  Verify (g2pubs/bls.go:159-162)
// keeps calling bls.CompareTwoPairings(sig, G2One, HashG1(m), pub), which the tag routes to the GPU engine.
// AggregateSignatures and AggregatePublicKeys of bls.go are REPLACED here under the tag (the maintainer moves the two
// originals, unchanged, into `aggregate_ref.go` starting with `// +build !b200`, INTEGRATION.md section 2); the batch entry
// points are additions.
package g2pubs

import (
	"github.com/phoreproject/bls"
)

// AggregateSignatures replaces g2pubs/bls.go:165-171 (signatures are G1 points here).
func AggregateSignatures(s []*Signature) *Signature {
	aff := make([]bls.G1Affine, len(s))
	for i, sig := range s {
		aff[i] = *sig.s.ToAffine()
	}
	return &Signature{s: bls.SumG1(aff)}
}

// AggregatePublicKeys replaces g2pubs/bls.go:180-186 (public keys are G2 points here).
func AggregatePublicKeys(p []*PublicKey) *PublicKey {
	aff := make([]bls.G2Affine, len(p))
	for i, pk := range p {
		aff[i] = *pk.p.ToAffine()
	}
	return &PublicKey{p: bls.SumG2(aff)}
}

// VerifyBatchObjects verifies many independent (message, public key, signature) triples held as objects in one launch:
// ok[i] == Verify(msgs[i], pubs[i], sigs[i])  (g2pubs/bls.go:159-162), i.e.
// e(sig, G2One) == e(HashG1(m), pub)  <=>  FE(ML(sig, G2One) * ML(-HashG1(m), pub)) == 1; false for infinite keys or signatures.
func VerifyBatchObjects(msgs [][]byte, pubs []*PublicKey, sigs []*Signature) []bool {
	n := len(msgs)
	if n == 0 {
		return []bool{}
	}
	p := make([]bls.G1Affine, 0, 2*n)
	q := make([]bls.G2Affine, 0, 2*n)
	off := make([]uint32, 1, n+1)
	for i := range msgs {
		h := bls.HashG1(msgs[i]).Copy() // *G1Affine (hash.go:326); hashing stays on the host (SURVEY N1)
		h.NegAssign()
		p = append(p, *sigs[i].s.ToAffine(), *h)
		q = append(q, *bls.G2AffineOne, *pubs[i].p.ToAffine())
		off = append(off, uint32(len(p)))
	}
	ok := bls.PairingProductsAreOne(p, q, off)
	for i := range ok {
		ok[i] = ok[i] && !pubs[i].p.IsZero() && !sigs[i].s.IsZero()
	}
	return ok
}

// VerifyBatch verifies n independent wire-format (public key, message, signature) triples: ok[i] ==
// Verify(msgs[i], DeserializePublicKey(pubs[i]), DeserializeSignature(sigs[i])) (g2pubs/bls.go:33-53,91-111,159-162).
func VerifyBatch(pubs [][96]byte, msgs [][]byte, sigs [][48]byte) []bool {
	p := make([]byte, 0, 96*len(pubs))
	s := make([]byte, 0, 48*len(sigs))
	for i := range pubs {
		p = append(p, pubs[i][:]...)
		s = append(s, sigs[i][:]...)
	}
	return bls.VerifyWire(true, p, msgs, s)
}
