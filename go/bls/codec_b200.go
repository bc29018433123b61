// +build b200

// codec_b200.go -- additions to package bls: the wire-format, hashing and scalar-multiplication entry points of
// libb381.so (include/b381.h), i.e. the work the callers of the verification path do on either side of it.
// Batch functions are additions; the single-item functions of the reference (DecompressG1/G2, CompressG1/G2,
// HashG2WithDomain, G1Affine.MulFR, G2Affine.MulFR) keep their signatures and can forward here with n = 1.
// NOT COMPILED in this repository's build image (no Go toolchain); see INTEGRATION.md.
package bls

/*
#include "b381.h"
*/
import "C"

import (
	"encoding/binary"
	"errors"
	"io"
	"unsafe"
)

// decodeErrors are the reference's own messages (g1.go:120,192,204,213; g2.go:158,226,234,242), indexed by status.
var decodeErrors = []error{
	nil,
	errors.New("unexpected compression mode"),
	errors.New("unexpected information in compressed infinity"),
	errors.New("point not on curve"),
	errors.New("not in correct subgroup"),
}

// DecompressG1Batch is DecompressG1 (g1.go:185-195) over many 48-byte encodings; errs[i] is nil or the
// reference's error.  checked = false gives DecompressG1Unchecked (g1.go:199-227).
func DecompressG1Batch(in [][48]byte, checked bool) ([]G1Affine, []error) {
	c, leave := enter()
	defer leave()
	n := len(in)
	out := make([]G1Affine, n)
	st := make([]uint8, n)
	errs := make([]error, n)
	if n == 0 {
		return out, errs
	}
	chk := C.int(0)
	if checked {
		chk = 1
	}
	rc := C.b381_g1_decompress_batch(c, (*C.uint8_t)(unsafe.Pointer(&in[0])), C.size_t(n), chk,
		(*C.b381_g1_affine)(unsafe.Pointer(&out[0])), (*C.uint8_t)(unsafe.Pointer(&st[0])))
	must(rc)
	for i, s := range st {
		errs[i] = decodeErrors[s]
	}
	return out, errs
}

// DecompressG2Batch is DecompressG2 (g2.go:219-229) / DecompressG2Unchecked (g2.go:232-265) over 96-byte encodings.
func DecompressG2Batch(in [][96]byte, checked bool) ([]G2Affine, []error) {
	c, leave := enter()
	defer leave()
	n := len(in)
	out := make([]G2Affine, n)
	st := make([]uint8, n)
	errs := make([]error, n)
	if n == 0 {
		return out, errs
	}
	chk := C.int(0)
	if checked {
		chk = 1
	}
	rc := C.b381_g2_decompress_batch(c, (*C.uint8_t)(unsafe.Pointer(&in[0])), C.size_t(n), chk,
		(*C.b381_g2_affine)(unsafe.Pointer(&out[0])), (*C.uint8_t)(unsafe.Pointer(&st[0])))
	must(rc)
	for i, s := range st {
		errs[i] = decodeErrors[s]
	}
	return out, errs
}

// CompressG1Batch / CompressG2Batch: CompressG1 (g1.go:230-249) / CompressG2 (g2.go:268-289).
func CompressG1Batch(p []G1Affine) [][48]byte {
	c, leave := enter()
	defer leave()
	out := make([][48]byte, len(p))
	if len(p) > 0 {
		must(C.b381_g1_compress_batch(c, (*C.b381_g1_affine)(unsafe.Pointer(&p[0])), C.size_t(len(p)),
			(*C.uint8_t)(unsafe.Pointer(&out[0]))))
	}
	return out
}

func CompressG2Batch(p []G2Affine) [][96]byte {
	c, leave := enter()
	defer leave()
	out := make([][96]byte, len(p))
	if len(p) > 0 {
		must(C.b381_g2_compress_batch(c, (*C.b381_g2_affine)(unsafe.Pointer(&p[0])), C.size_t(len(p)),
			(*C.uint8_t)(unsafe.Pointer(&out[0]))))
	}
	return out
}

// scalars flattens []*FR (a pointer type, fr.go:11-13) into canonical 4 x u64 integers: FR.ToRepr (fr.go:316-329).
func scalars(k []*FR) []FRRepr {
	out := make([]FRRepr, len(k))
	for i, s := range k {
		out[i] = *s.ToRepr()
	}
	return out
}

// MulG1Batch: out[i] = p[i].MulFR(k[i]).ToAffine() (g1.go:80-90,322-340).  len(p) == 1 broadcasts the base
// (PrivToPub, g1pubs/bls.go:144-146); len(k) == 1 broadcasts the scalar.
func MulG1Batch(p []G1Affine, k []*FR) []G1Affine {
	if len(p) == 0 || len(k) == 0 {
		return []G1Affine{}
	}
	c, leave := enter()
	defer leave()
	n := len(p)
	if len(k) > n {
		n = len(k)
	}
	out := make([]G1Affine, n)
	ks := scalars(k)
	ps, kst := C.size_t(1), C.size_t(1)
	if len(p) == 1 && n > 1 {
		ps = 0
	}
	if len(k) == 1 && n > 1 {
		kst = 0
	}
	must(C.b381_g1_mul_batch(c, (*C.b381_g1_affine)(unsafe.Pointer(&p[0])), ps,
		(*C.b381_scalar)(unsafe.Pointer(&ks[0])), kst, C.size_t(n), (*C.b381_g1_affine)(unsafe.Pointer(&out[0]))))
	return out
}

// MulG2Batch: out[i] = p[i].MulFR(k[i]).ToAffine() (g2.go:92-102,365-386): Sign of many hashed messages.
func MulG2Batch(p []G2Affine, k []*FR) []G2Affine {
	if len(p) == 0 || len(k) == 0 {
		return []G2Affine{}
	}
	c, leave := enter()
	defer leave()
	n := len(p)
	if len(k) > n {
		n = len(k)
	}
	out := make([]G2Affine, n)
	ks := scalars(k)
	ps, kst := C.size_t(1), C.size_t(1)
	if len(p) == 1 && n > 1 {
		ps = 0
	}
	if len(k) == 1 && n > 1 {
		kst = 0
	}
	must(C.b381_g2_mul_batch(c, (*C.b381_g2_affine)(unsafe.Pointer(&p[0])), ps,
		(*C.b381_scalar)(unsafe.Pointer(&ks[0])), kst, C.size_t(n), (*C.b381_g2_affine)(unsafe.Pointer(&out[0]))))
	return out
}

// mulStrides: a single point or a single scalar is broadcast over the batch.
func mulStrides(np, nk int) (n int, ps, ks C.size_t) {
	n = np
	if nk > n {
		n = nk
	}
	ps, ks = 1, 1
	if np == 1 && n > 1 {
		ps = 0
	}
	if nk == 1 && n > 1 {
		ks = 0
	}
	return
}

// MulG1SubgroupBatch is MulG1Batch for points known to lie in G1 (the generator, a hash to the curve, a checked key):
// the engine takes the endomorphism ladder (1.5x / 1.9x faster), same result.  PrivToPub (g2pubs keys live in G2,
// g1pubs keys in G1) and Sign use these.
func MulG1SubgroupBatch(p []G1Affine, k []*FR) []G1Affine {
	if len(p) == 0 || len(k) == 0 {
		return []G1Affine{}
	}
	c, leave := enter()
	defer leave()
	n, ps, kst := mulStrides(len(p), len(k))
	out := make([]G1Affine, n)
	ks := scalars(k)
	must(C.b381_g1_mul_subgroup_batch(c, (*C.b381_g1_affine)(unsafe.Pointer(&p[0])), ps,
		(*C.b381_scalar)(unsafe.Pointer(&ks[0])), kst, C.size_t(n), (*C.b381_g1_affine)(unsafe.Pointer(&out[0]))))
	return out
}

// MulG2SubgroupBatch: see MulG1SubgroupBatch.
func MulG2SubgroupBatch(p []G2Affine, k []*FR) []G2Affine {
	if len(p) == 0 || len(k) == 0 {
		return []G2Affine{}
	}
	c, leave := enter()
	defer leave()
	n, ps, kst := mulStrides(len(p), len(k))
	out := make([]G2Affine, n)
	ks := scalars(k)
	must(C.b381_g2_mul_subgroup_batch(c, (*C.b381_g2_affine)(unsafe.Pointer(&p[0])), ps,
		(*C.b381_scalar)(unsafe.Pointer(&ks[0])), kst, C.size_t(n), (*C.b381_g2_affine)(unsafe.Pointer(&out[0]))))
	return out
}

// HashG2WithDomainBatch: out[i] = HashG2WithDomain(msgs[i], domain).ToAffine() (g2.go:1041-1085).
func HashG2WithDomainBatch(msgs [][32]byte, domain [8]byte) []G2Affine {
	c, leave := enter()
	defer leave()
	out := make([]G2Affine, len(msgs))
	if len(msgs) > 0 {
		must(C.b381_hash_g2_with_domain_batch(c, (*C.uint8_t)(unsafe.Pointer(&msgs[0])),
			(*C.uint8_t)(unsafe.Pointer(&domain[0])), 0, C.size_t(len(msgs)), (*C.b381_g2_affine)(unsafe.Pointer(&out[0]))))
	}
	return out
}

// VerifyWithDomainWire verifies n wire-format triples entirely on the device: ok[i] ==
// g1pubs.VerifyWithDomain(msgs[i], DeserializePublicKey(pubs[i]), DeserializeSignature(sigs[i]), domain)
// (g1pubs/bls.go:38-58,91-111,171-174), false when either deserialisation fails.
func VerifyWithDomainWire(pubs [][48]byte, msgs [][32]byte, domain [8]byte, sigs [][96]byte) []bool {
	c, leave := enter()
	defer leave()
	n := len(pubs)
	ok8 := make([]uint8, n)
	out := make([]bool, n)
	if n == 0 {
		return out
	}
	must(C.b381_verify_with_domain_batch(c, (*C.uint8_t)(unsafe.Pointer(&pubs[0])), (*C.uint8_t)(unsafe.Pointer(&msgs[0])),
		(*C.uint8_t)(unsafe.Pointer(&domain[0])), 0, (*C.uint8_t)(unsafe.Pointer(&sigs[0])), C.size_t(n),
		(*C.uint8_t)(unsafe.Pointer(&ok8[0]))))
	for i, v := range ok8 {
		out[i] = v != 0
	}
	return out
}

// packMessages lays variable-length messages back to back with an offsets array (n + 1 entries).
func packMessages(msgs [][]byte) ([]byte, []uint64) {
	off := make([]uint64, len(msgs)+1)
	total := 0
	for i, m := range msgs {
		total += len(m)
		off[i+1] = uint64(total)
	}
	buf := make([]byte, 0, total+1)
	for _, m := range msgs {
		buf = append(buf, m...)
	}
	return append(buf, 0), off
}

// HashG1Batch / HashG2Batch: HashG1 (hash.go:320-331) / HashG2 (hash.go:404-411) of every message.
func HashG1Batch(msgs [][]byte) []G1Affine {
	c, leave := enter()
	defer leave()
	out := make([]G1Affine, len(msgs))
	if len(msgs) > 0 {
		buf, off := packMessages(msgs)
		must(C.b381_hash_g1_batch(c, (*C.uint8_t)(unsafe.Pointer(&buf[0])), (*C.uint64_t)(unsafe.Pointer(&off[0])),
			C.size_t(len(msgs)), (*C.b381_g1_affine)(unsafe.Pointer(&out[0]))))
	}
	return out
}

func HashG2Batch(msgs [][]byte) []G2Affine {
	c, leave := enter()
	defer leave()
	out := make([]G2Affine, len(msgs))
	if len(msgs) > 0 {
		buf, off := packMessages(msgs)
		must(C.b381_hash_g2_batch(c, (*C.uint8_t)(unsafe.Pointer(&buf[0])), (*C.uint64_t)(unsafe.Pointer(&off[0])),
			C.size_t(len(msgs)), (*C.b381_g2_affine)(unsafe.Pointer(&out[0]))))
	}
	return out
}

// VerifyWire: g1pubs.Verify (g2pubs = false: 48-byte keys, 96-byte signatures, g1pubs/bls.go:165-168) or g2pubs.Verify
// (g2pubs = true: 96-byte keys, 48-byte signatures, g2pubs/bls.go:159-162) for n wire-format triples on the device.
func VerifyWire(g2pubs bool, pubs []byte, msgs [][]byte, sigs []byte) []bool {
	c, leave := enter()
	defer leave()
	n := len(msgs)
	ok8 := make([]uint8, n)
	out := make([]bool, n)
	if n == 0 {
		return out
	}
	buf, off := packMessages(msgs)
	if g2pubs {
		must(C.b381_g2pubs_verify_batch(c, (*C.uint8_t)(unsafe.Pointer(&pubs[0])), (*C.uint8_t)(unsafe.Pointer(&buf[0])),
			(*C.uint64_t)(unsafe.Pointer(&off[0])), (*C.uint8_t)(unsafe.Pointer(&sigs[0])), C.size_t(n), (*C.uint8_t)(unsafe.Pointer(&ok8[0]))))
	} else {
		must(C.b381_g1pubs_verify_batch(c, (*C.uint8_t)(unsafe.Pointer(&pubs[0])), (*C.uint8_t)(unsafe.Pointer(&buf[0])),
			(*C.uint64_t)(unsafe.Pointer(&off[0])), (*C.uint8_t)(unsafe.Pointer(&sigs[0])), C.size_t(n), (*C.uint8_t)(unsafe.Pointer(&ok8[0]))))
	}
	for i, v := range ok8 {
		out[i] = v != 0
	}
	return out
}

// must is called with the engine locked (enter): the error text belongs to the call that just failed.
func must(rc C.int) {
	if rc != C.B381_OK {
		panic(C.GoString(C.b381_last_error(engine.ctx)))
	}
}

// enter locks the engine for one cgo call and returns the context with the unlock function.  A b381_ctx serves one OS
// thread at a time (include/b381.h): every entry point shares its stream and its grow-only device scratch, so concurrent
// goroutines -- the normal case for a Go verifier -- are serialised here, exactly as the functions of pairing_b200.go do.
func enter() (*C.b381_ctx, func()) {
	engine.mu.Lock()
	return ctx(), engine.mu.Unlock
}

// VerifyWithDomainRLC checks n wire-format triples with ONE final exponentiation (random linear combination):
// true iff every VerifyWithDomain equation holds, up to a false-accept probability of 2^-64.  The weights must be
// drawn after the signatures are fixed, from a cryptographic source.
func VerifyWithDomainRLC(pubs [][48]byte, msgs [][32]byte, domain [8]byte, sigs [][96]byte, rnd io.Reader) bool {
	c, leave := enter()
	defer leave()
	n := len(pubs)
	if n == 0 {
		return true
	}
	w := make([]FRRepr, n)
	var b [8]byte
	for i := range w {
		if _, err := io.ReadFull(rnd, b[:]); err != nil {
			panic(err)
		}
		w[i][0] = binary.LittleEndian.Uint64(b[:]) | 1
	}
	var ok C.uint8_t
	must(C.b381_set_rlc_weight_bits(c, 64))
	must(C.b381_verify_with_domain_rlc_batch(c, (*C.uint8_t)(unsafe.Pointer(&pubs[0])), (*C.uint8_t)(unsafe.Pointer(&msgs[0])),
		(*C.uint8_t)(unsafe.Pointer(&domain[0])), 0, (*C.uint8_t)(unsafe.Pointer(&sigs[0])),
		(*C.b381_scalar)(unsafe.Pointer(&w[0])), C.size_t(n), &ok))
	return ok != 0
}

// KeyRegistry is a table of validated public keys resident on the device (BASELINE config 5: the validator registry).
// Attestations refer to its entries by index, so a batch moves only indices, signatures and message points.
type KeyRegistry struct {
	dev unsafe.Pointer
	n   int
}

// NewKeyRegistry uploads the keys once (b381_dev_alloc + b381_h2d).  Keys must be group elements
// (DeserializePublicKey / DecompressG1Batch with the subgroup check).
func NewKeyRegistry(keys []G1Affine) *KeyRegistry {
	c, leave := enter()
	defer leave()
	r := &KeyRegistry{n: len(keys)}
	if len(keys) == 0 {
		return r
	}
	bytes := C.size_t(len(keys)) * C.size_t(unsafe.Sizeof(keys[0]))
	must(C.b381_dev_alloc(c, bytes, &r.dev))
	must(C.b381_h2d(c, r.dev, unsafe.Pointer(&keys[0]), bytes))
	must(C.b381_sync(c))
	return r
}

// Free releases the device table.
func (r *KeyRegistry) Free() {
	c, leave := enter()
	defer leave()
	if r.dev != nil {
		C.b381_dev_free(c, r.dev)
		r.dev = nil
	}
}

// attestationBatch stages the per-batch arrays on the device: CSR committees, signatures, message points, message indices
// (and the weights of the one-boolean form).  The caller holds the engine lock.
type attestationBatch struct {
	ptr [6]unsafe.Pointer
}

func (b *attestationBatch) free(c *C.b381_ctx) {
	for _, p := range b.ptr {
		if p != nil {
			C.b381_dev_free(c, p)
		}
	}
}

// upload allocates bytes on the device (at least 8) and, when host is not nil, copies them from host.
func upload(c *C.b381_ctx, host unsafe.Pointer, bytes int) unsafe.Pointer {
	var d unsafe.Pointer
	alloc := bytes
	if alloc < 8 {
		alloc = 8
	}
	must(C.b381_dev_alloc(c, C.size_t(alloc), &d))
	if host != nil && bytes > 0 {
		must(C.b381_h2d(c, d, host, C.size_t(bytes)))
	}
	return d
}

func stageAttestations(c *C.b381_ctx, committees [][]uint32, sigs []G2Affine, msgPoints []G2Affine, msgIdx []uint32) *attestationBatch {
	keyOff := make([]uint32, 1, len(committees)+1)
	keyIdx := make([]uint32, 0, 128*len(committees))
	for _, cm := range committees {
		keyIdx = append(keyIdx, cm...)
		keyOff = append(keyOff, uint32(len(keyIdx)))
	}
	b := &attestationBatch{}
	var ki unsafe.Pointer
	if len(keyIdx) > 0 {
		ki = unsafe.Pointer(&keyIdx[0])
	}
	b.ptr[0] = upload(c, ki, 4*len(keyIdx))
	b.ptr[1] = upload(c, unsafe.Pointer(&keyOff[0]), 4*len(keyOff))
	b.ptr[2] = upload(c, unsafe.Pointer(&sigs[0]), len(sigs)*int(unsafe.Sizeof(sigs[0])))
	b.ptr[3] = upload(c, unsafe.Pointer(&msgPoints[0]), len(msgPoints)*int(unsafe.Sizeof(msgPoints[0])))
	b.ptr[4] = upload(c, unsafe.Pointer(&msgIdx[0]), 4*len(msgIdx))
	return b
}

// VerifyAttestations: ok[a] == sigs[a].VerifyAggregateCommon(keys of committees[a], message msgIdx[a])
// (g1pubs/bls.go:287-290) for a batch of attestations over the registry; msgPoints are the distinct hashed messages
// (HashG2Batch / HashG2WithDomainBatch).  An empty committee, an index outside the tables, an infinite signature or
// aggregate key give false.
func (r *KeyRegistry) VerifyAttestations(committees [][]uint32, sigs []G2Affine, msgPoints []G2Affine, msgIdx []uint32) []bool {
	n := len(sigs)
	if n == 0 {
		return []bool{}
	}
	if len(committees) != n || len(msgIdx) != n || len(msgPoints) == 0 || r.n == 0 {
		panic("bls: attestation batch arrays do not match")
	}
	c, leave := enter()
	defer leave()
	b := stageAttestations(c, committees, sigs, msgPoints, msgIdx)
	defer b.free(c)
	b.ptr[5] = upload(c, nil, n)
	must(C.b381_verify_aggregate_common_batch_dev(c, (*C.b381_g1_affine)(r.dev), (*C.uint32_t)(b.ptr[0]), (*C.uint32_t)(b.ptr[1]),
		(*C.b381_g2_affine)(b.ptr[2]), (*C.b381_g2_affine)(b.ptr[3]), (*C.uint32_t)(b.ptr[4]), C.size_t(n), C.size_t(r.n),
		C.size_t(len(msgPoints)), (*C.uint8_t)(b.ptr[5])))
	st := make([]uint8, n)
	must(C.b381_d2h(c, unsafe.Pointer(&st[0]), b.ptr[5], C.size_t(n)))
	ok := make([]bool, n)
	for i, s := range st {
		ok[i] = s != 0
	}
	return ok
}

// VerifyAttestationsOneBoolean is the batch as a single check (random linear combination grouped by message,
// b381_verify_aggregate_common_rlc_dev): true iff every attestation of VerifyAttestations is true, up to a false-accept
// probability of 2^-64; len(msgPoints) + 1 Miller loops and one final exponentiation for the whole batch.  The weights are
// drawn here, after the batch is fixed, from rnd (a cryptographic source).  On false, VerifyAttestations locates the offenders.
func (r *KeyRegistry) VerifyAttestationsOneBoolean(committees [][]uint32, sigs []G2Affine, msgPoints []G2Affine, msgIdx []uint32, rnd io.Reader) bool {
	n := len(sigs)
	if n == 0 {
		return true
	}
	if len(committees) != n || len(msgIdx) != n || len(msgPoints) == 0 || r.n == 0 {
		return false
	}
	w := make([]FRRepr, n)
	var buf [8]byte
	for i := range w {
		if _, err := io.ReadFull(rnd, buf[:]); err != nil {
			panic(err)
		}
		w[i][0] = binary.LittleEndian.Uint64(buf[:]) | 1
	}
	c, leave := enter()
	defer leave()
	b := stageAttestations(c, committees, sigs, msgPoints, msgIdx)
	defer b.free(c)
	b.ptr[5] = upload(c, unsafe.Pointer(&w[0]), n*int(unsafe.Sizeof(w[0])))
	var dok unsafe.Pointer
	must(C.b381_dev_alloc(c, 8, &dok))
	defer C.b381_dev_free(c, dok)
	must(C.b381_set_rlc_weight_bits(c, 64))
	must(C.b381_verify_aggregate_common_rlc_dev(c, (*C.b381_g1_affine)(r.dev), (*C.uint32_t)(b.ptr[0]), (*C.uint32_t)(b.ptr[1]),
		(*C.b381_g2_affine)(b.ptr[2]), (*C.b381_g2_affine)(b.ptr[3]), (*C.uint32_t)(b.ptr[4]), (*C.b381_scalar)(b.ptr[5]),
		C.size_t(n), C.size_t(r.n), C.size_t(len(msgPoints)), (*C.uint8_t)(dok)))
	var ok uint8
	must(C.b381_d2h(c, unsafe.Pointer(&ok), dok, 1))
	return ok != 0
}
