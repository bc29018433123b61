// +build b200

// pairing_b200.go -- drop-in for pairing.go of github.com/phoreproject/bls when built with
// `-tags b200`: same exported functions, work done by libb381.so (hand-written sm_100a CUDA)
// through cgo.  Add `// +build !b200` to the top of the original pairing.go; nothing else in the
// package changes -- the same mechanism the package already uses to pick stub.go (amd64 asm) over
// stub_fallback.go (pure Go).
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain.  The C ABI below is
// exercised by tests/ through ctypes with the identical memory layouts.
//
// Layout contract (checked by static_asserts in bls_b200/csrc/b381.cu and tests/test_abi.py):
//   G1Affine{x, y FQ; infinity bool}   == b381_g1_affine (104 B)      g1.go:10-14
//   G2Affine{x, y FQ2; infinity bool}  == b381_g2_affine (200 B)      g2.go:12-16
//   G1Projective / G2Projective        == b381_g1_jac / b381_g2_jac   g1.go:252-256, g2.go:298-302
//   FQ12{c0, c1 *FQ6} is two POINTERS (fq12.go:9-12): flattened to b381_fp12 (576 B) here.
package bls

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../bls_b200 -lb381 -Wl,-rpath,${SRCDIR}/../../bls_b200
#include "b381.h"
*/
import "C"

import (
	"runtime"
	"sync"
	"unsafe"
)

// engine is one b381_ctx per process (a ctx may be used by one OS thread at a time).
var engine struct {
	once sync.Once
	mu   sync.Mutex
	ctx  *C.b381_ctx
	err  C.int
}

func ctx() *C.b381_ctx {
	engine.once.Do(func() { engine.err = C.b381_init(0, &engine.ctx) })
	if engine.err != C.B381_OK {
		panic("bls/b200: no usable CUDA device (libb381 has no CPU fallback)")
	}
	return engine.ctx
}

func flatFQ12(f *C.b381_fp12) *FQ12 {
	// b381_fp12 = c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2, each FQ2 = c0 || c1 (Montgomery limbs)
	v := (*[6]FQ2)(unsafe.Pointer(f))
	return &FQ12{
		c0: &FQ6{c0: v[0], c1: v[1], c2: v[2]},
		c1: &FQ6{c0: v[3], c1: v[4], c2: v[5]},
	}
}

func packFQ12(f *FQ12) (out C.b381_fp12) {
	v := (*[6]FQ2)(unsafe.Pointer(&out))
	v[0], v[1], v[2], v[3], v[4], v[5] = f.c0.c0, f.c0.c1, f.c0.c2, f.c1.c0, f.c1.c1, f.c1.c2
	return
}

// MillerLoopItem is pairing.go:4-7 (the file this one replaces under -tags b200): a G1 point and a prepared G2 point.
// G2Prepared and G2AffineToPrepared stay the reference's (g2.go:639-801, not replaced); G2AffineToPreparedBatch below
// prepares many points in one launch.
type MillerLoopItem struct {
	P *G1Affine
	Q *G2Prepared
}

// packPrepared copies a G2Prepared (coeffs [][3]FQ2, 68 entries, g2.go:639-642) into the engine's flat form.  An infinite
// point has no coefficients in the reference; the engine reads none of them.
func packPrepared(dst *C.b381_g2_prepared, q *G2Prepared) {
	if q.infinity {
		dst.infinity = 1
		return
	}
	if len(q.coeffs) != 68 {
		panic("bls/b200: G2Prepared with an unexpected number of coefficient triples")
	}
	v := (*[68][3]FQ2)(unsafe.Pointer(&dst.coeffs[0]))
	copy(v[:], q.coeffs)
}

// MillerLoop replaces pairing.go:16-75: the product over the items of f_{|x|,Q}(P), conjugated.  Each item's Miller value
// is computed from the prepared coefficients on the device (b381_miller_loop_prepared_batch); the reference's shared
// accumulator (pairing.go:40-69) computes exactly the product of the per-item values, formed here with FQ12.MulAssign.
// An item with P or Q at infinity contributes the factor 1 (the reference panics there, indexing an empty slice).
func MillerLoop(items []MillerLoopItem) *FQ12 {
	f := FQ12One.Copy()
	n := len(items)
	if n == 0 {
		return f
	}
	ps := make([]G1Affine, n)
	prep := make([]C.b381_g2_prepared, n)
	for i, it := range items {
		ps[i] = *it.P
		packPrepared(&prep[i], it.Q)
	}
	ml := make([]C.b381_fp12, n)
	engine.mu.Lock()
	rc := C.b381_miller_loop_prepared_batch(ctx(), (*C.b381_g1_affine)(unsafe.Pointer(&ps[0])), &prep[0], C.size_t(n), nil,
		C.size_t(n), &ml[0])
	engine.mu.Unlock()
	if rc != C.B381_OK {
		panic(C.GoString(C.b381_last_error(ctx())))
	}
	runtime.KeepAlive(ps)
	for i := range ml {
		f.MulAssign(flatFQ12(&ml[i]))
	}
	return f
}

// G2AffineToPreparedBatch is G2AffineToPrepared (g2.go:650-801) for many points in one launch.
func G2AffineToPreparedBatch(qs []G2Affine) []*G2Prepared {
	n := len(qs)
	out := make([]*G2Prepared, n)
	if n == 0 {
		return out
	}
	prep := make([]C.b381_g2_prepared, n)
	engine.mu.Lock()
	rc := C.b381_g2_prepare_batch(ctx(), (*C.b381_g2_affine)(unsafe.Pointer(&qs[0])), C.size_t(n), &prep[0])
	engine.mu.Unlock()
	if rc != C.B381_OK {
		panic(C.GoString(C.b381_last_error(ctx())))
	}
	for i := range prep {
		if prep[i].infinity != 0 {
			out[i] = &G2Prepared{infinity: true}
			continue
		}
		v := (*[68][3]FQ2)(unsafe.Pointer(&prep[i].coeffs[0]))
		c := make([][3]FQ2, 68)
		copy(c, v[:])
		out[i] = &G2Prepared{coeffs: c}
	}
	return out
}

// MillerLoopAffine is the batched form on unprepared points: the engine fuses G2AffineToPrepared into the loop
// (no 19.6 KB coefficient table per point); the product of the Miller values is returned.
func MillerLoopAffine(ps []G1Affine, qs []G2Affine) *FQ12 {
	f := FQ12One.Copy()
	n := len(ps)
	if n == 0 {
		return f
	}
	if len(qs) != n {
		panic("bls/b200: MillerLoopAffine needs as many G2 points as G1 points")
	}
	ml := make([]C.b381_fp12, n)
	engine.mu.Lock()
	rc := C.b381_miller_loop_batch(ctx(), (*C.b381_g1_affine)(unsafe.Pointer(&ps[0])),
		(*C.b381_g2_affine)(unsafe.Pointer(&qs[0])), C.size_t(n), &ml[0])
	engine.mu.Unlock()
	if rc != C.B381_OK {
		panic(C.GoString(C.b381_last_error(ctx())))
	}
	runtime.KeepAlive(ps)
	runtime.KeepAlive(qs)
	for i := range ml {
		f.MulAssign(flatFQ12(&ml[i])) // shared accumulator of pairing.go:40-69
	}
	return f
}

// FinalExponentiation replaces pairing.go:79-129 (nil for r == 0, as the reference).
func FinalExponentiation(r *FQ12) *FQ12 {
	in := packFQ12(r)
	var out C.b381_fp12
	var ok C.uint8_t
	engine.mu.Lock()
	defer engine.mu.Unlock()
	if C.b381_final_exp_batch(ctx(), &in, 1, &out, &ok) != C.B381_OK {
		panic(C.GoString(C.b381_last_error(ctx())))
	}
	if ok == 0 {
		return nil
	}
	return flatFQ12(&out)
}

// Pairing replaces pairing.go:132-136.
func Pairing(p *G1Projective, q *G2Projective) *FQ12 {
	return PairingBatch([]*G1Affine{p.ToAffine()}, []*G2Affine{q.ToAffine()})[0]
}

// PairingBatch is the batched form the reference lacks: out[i] = Pairing(ps[i], qs[i]).
func PairingBatch(ps []*G1Affine, qs []*G2Affine) []*FQ12 {
	n := len(ps)
	if n == 0 {
		return []*FQ12{}
	}
	p := make([]G1Affine, n)
	q := make([]G2Affine, n)
	for i := range ps {
		p[i], q[i] = *ps[i], *qs[i]
	}
	out := make([]C.b381_fp12, n)
	engine.mu.Lock()
	defer engine.mu.Unlock()
	// batches of 2^16 pairs per launch; with more than one batch the copies overlap the kernels (b381_pairing_batch_stream)
	rc := C.b381_pairing_batch_stream(ctx(), (*C.b381_g1_affine)(unsafe.Pointer(&p[0])),
		(*C.b381_g2_affine)(unsafe.Pointer(&q[0])), C.size_t(n), C.size_t(1<<16), &out[0])
	if rc != C.B381_OK {
		panic(C.GoString(C.b381_last_error(ctx())))
	}
	res := make([]*FQ12, n)
	for i := range out {
		res[i] = flatFQ12(&out[i])
	}
	return res
}

// CompareTwoPairings replaces pairing.go:140-147: e(P1,Q1) == e(P2,Q2) as one 2-pair group.
// A point at infinity on either side makes the comparison false: the engine's Miller loop treats such a pair as the factor 1
// (the reference panics there, pairing.go:17-26), and Verify(m, infinity, infinity) must not hold for every m.
func CompareTwoPairings(P1 *G1Projective, Q1 *G2Projective, P2 *G1Projective, Q2 *G2Projective) bool {
	if P1.IsZero() || Q1.IsZero() || P2.IsZero() || Q2.IsZero() {
		return false
	}
	negP2 := P2.ToAffine()
	negP2.NegAssign()
	return PairingProductsAreOne(
		[]G1Affine{*P1.ToAffine(), *negP2},
		[]G2Affine{*Q1.ToAffine(), *Q2.ToAffine()},
		[]uint32{0, 2})[0]
}

// PairingProductsAreOne: for every group g, FinalExponentiation(prod MillerLoop(p[i], q[i])) == 1
// over i in [off[g], off[g+1]).  One launch verifies any number of signatures.
func PairingProductsAreOne(p []G1Affine, q []G2Affine, off []uint32) []bool {
	ng := len(off) - 1
	if ng <= 0 {
		return []bool{}
	}
	if len(p) == 0 { // every group is empty: the empty product is 1
		res := make([]bool, ng)
		for i := range res {
			res[i] = true
		}
		return res
	}
	ok := make([]C.uint8_t, ng)
	engine.mu.Lock()
	defer engine.mu.Unlock()
	rc := C.b381_pairing_product_is_one(ctx(), (*C.b381_g1_affine)(unsafe.Pointer(&p[0])),
		(*C.b381_g2_affine)(unsafe.Pointer(&q[0])), C.size_t(len(p)),
		(*C.uint32_t)(unsafe.Pointer(&off[0])), C.size_t(ng), &ok[0])
	if rc != C.B381_OK {
		panic(C.GoString(C.b381_last_error(ctx())))
	}
	res := make([]bool, ng)
	for i := range ok {
		res[i] = ok[i] != 0
	}
	return res
}

// SumG1 == the G1Projective.Add fold of g1pubs.AggregatePublicKeys (g1pubs/bls.go:192-198);
// the result is normalised (z = 1 or the canonical zero) and Equal()s the fold.
func SumG1(ps []G1Affine) *G1Projective {
	var out G1Projective
	engine.mu.Lock()
	defer engine.mu.Unlock()
	var p0 *C.b381_g1_affine
	if len(ps) > 0 {
		p0 = (*C.b381_g1_affine)(unsafe.Pointer(&ps[0]))
	}
	if C.b381_g1_sum(ctx(), p0, C.size_t(len(ps)), (*C.b381_g1_jac)(unsafe.Pointer(&out))) != C.B381_OK {
		panic(C.GoString(C.b381_last_error(ctx())))
	}
	return &out
}

// SumG2 == the G2Projective.Add fold of g1pubs.AggregateSignatures (g1pubs/bls.go:177-183).
func SumG2(qs []G2Affine) *G2Projective {
	var out G2Projective
	engine.mu.Lock()
	defer engine.mu.Unlock()
	var q0 *C.b381_g2_affine
	if len(qs) > 0 {
		q0 = (*C.b381_g2_affine)(unsafe.Pointer(&qs[0]))
	}
	if C.b381_g2_sum(ctx(), q0, C.size_t(len(qs)), (*C.b381_g2_jac)(unsafe.Pointer(&out))) != C.B381_OK {
		panic(C.GoString(C.b381_last_error(ctx())))
	}
	return &out
}

// MSMG1 = sum_i k[i] * ps[i] with k[i] = FR.ToRepr() limbs (fr.go:316-329): the weighted aggregate
// the reference would compute as a fold of G1Affine.MulFR (g1.go:80-90).
func MSMG1(ps []G1Affine, ks []FRRepr) *G1Projective {
	var out G1Projective
	if len(ps) == 0 {
		return G1ProjectiveZero.Copy()
	}
	if len(ks) != len(ps) {
		panic("bls/b200: MSMG1 needs one scalar per point")
	}
	engine.mu.Lock()
	defer engine.mu.Unlock()
	rc := C.b381_g1_msm(ctx(), (*C.b381_g1_affine)(unsafe.Pointer(&ps[0])),
		(*C.b381_scalar)(unsafe.Pointer(&ks[0])), C.size_t(len(ps)), (*C.b381_g1_jac)(unsafe.Pointer(&out)))
	if rc != C.B381_OK {
		panic(C.GoString(C.b381_last_error(ctx())))
	}
	return &out
}
