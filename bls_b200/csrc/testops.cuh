// testops.cuh -- TEST HOOKS: the device build of the field / tower / group-law routines exposed one operation at a time, so
// that the reference's known-answer tests (fq_test.go:189-207, fq2_test.go:71-246, g1_test.go:62-104), edge operands and random
// vectors reach the PTX multiplier and the lane-cooperative routines directly on the GPU (tests/test_gpu_ops.py), not only
// through whole pairings.  Nothing on the product path calls these kernels.
#pragma once
#include "pairing.cuh"
#include "curve.cuh"
#include "quad.cuh"
#include "duo.cuh"

namespace b381 {
#if defined(__CUDACC__)

// family 0: Fq (6 x u64 per element).  op: 0 mul 1 add 2 sub 3 sqr 4 neg 5 dbl 6 inv 8 inv by the Fermat chain 9 a*b + a*b as one dot product
// 10 the multiplication body inlined 11 the dedicated squaring body (FQ.SquareAssign, fq.go:151-198)
__global__ void k_test_fp(int op, const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp x, y, r;
    fp_load_u64(x, a + 6 * i); fp_load_u64(y, b + 6 * i);
    switch (op) {
        case 0: fp_mul(r, x, y); break;
        case 1: fp_add(r, x, y); break;
        case 2: fp_sub(r, x, y); break;
        case 3: fp_sqr(r, x); break;
        case 4: fp_neg(r, x); break;
        case 5: fp_dbl(r, x); break;
        case 6: fp_inv(&r, &x); break;
        case 8: fp_inv_fermat(&r, &x); break;
        case 9: r = fp_dot2_v(x, y, y, x); break;
        case 10: fp_mul_inl(r, x, y); break;
        case 11: r = fp_sqr_v(x); break;
        default: r = x;
    }
    fp_store_u64(o + 6 * i, r);
}
// family 1: Fq2 (12 x u64).  op: 0 mul 1 add 2 sub 3 sqr 4 neg 5 dbl 6 inv 10 times (1 + u) 11 a * b.c0 (Fq scalar) 12 conjugate
__global__ void k_test_fp2(int op, const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp2 x, y, r;
    fp2_load_u64(x, a + 12 * i); fp2_load_u64(y, b + 12 * i);
    r = x;
    switch (op) {
        case 0: fp2_mul(&r, &x, &y); break;
        case 1: fp2_add(r, x, y); break;
        case 2: fp2_sub(r, x, y); break;
        case 3: fp2_sqr(&r, &x); break;
        case 4: fp2_neg(r, x); break;
        case 5: fp2_dbl(r, x); break;
        case 6: fp2_inv(&r, &x); break;
        case 10: fp2_mul_nr(r, x); break;
        case 11: fp2_mul_fp(&r, &x, &y.c0); break;
        case 12: fp2_conj(r, x); break;
    }
    fp2_store_u64(o + 12 * i, r);
}
// family 2: Fq6 (36 x u64).  op: 0 mul 1 add 2 sub 4 neg 6 inv 7 frobenius(arg) 10 times v 13 mul_by_01(b.c0, b.c1) 14 mul_by_1(b.c1)
__global__ void k_test_fp6(int op, uint64_t arg, const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp6 x, y, r;
    fp *px = fp_array(&x), *py = fp_array(&y), *pr = fp_array(&r);
    for (int k = 0; k < 6; k++) { fp_load_u64(px[k], a + 36 * i + 6 * k); fp_load_u64(py[k], b + 36 * i + 6 * k); }
    r = x;
    switch (op) {
        case 0: fp6_mul(&r, &x, &y); break;
        case 1: fp6_add(&r, &x, &y); break;
        case 2: fp6_sub(&r, &x, &y); break;
        case 4: fp6_neg(&r, &x); break;
        case 6: fp6_inv(&r, &x); break;
        case 7: fp6_frobenius(&r, &x, (int)arg); break;
        case 10: fp6_mul_nr(&r, &x); break;
        case 13: fp6_mul_by_01(&r, &x, &y.c0, &y.c1); break;
        case 14: fp6_mul_by_1(&r, &x, &y.c1); break;
    }
    for (int k = 0; k < 6; k++) fp_store_u64(o + 36 * i + 6 * k, pr[k]);
}
// family 3: Fq12, one element per thread (72 x u64).  op: 0 mul 3 sqr 6 inv 7 frobenius(arg) 12 conj 13 mul_by_014(b.c0.c0, b.c0.c1,
// b.c1.c1) 15 cyclotomic sqr 16 exp_by_x(arg) 17 exp_by_x_gs(arg).  ok[i] = 0 where an inverse does not exist.
__global__ void k_test_fp12(int op, uint64_t arg, const uint64_t *a, const uint64_t *b, uint64_t *o, uint8_t *ok, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp12 x, y, r;
    fp12_load_u64(&x, a + 72 * i); fp12_load_u64(&y, b + 72 * i);
    fp12_copy(&r, &x);
    bool good = true;
    switch (op) {
        case 0: fp12_mul(&r, &x, &y); break;
        case 3: fp12_sqr(&r, &x); break;
        case 6: good = fp12_inv(&r, &x); break;
        case 7: fp12_frobenius(&r, &x, (int)arg); break;
        case 12: fp12_conj(&r, &x); break;
        case 13: fp12_mul_by_014(&r, &y.c0.c0, &y.c0.c1, &y.c1.c1); break;
        case 15: fp12_cyclotomic_sqr(&r, &x); break;
        case 16: exp_by_x(&r, &x, arg); break;
        case 17: exp_by_x_gs(&r, &x, arg); break;
    }
    fp12_store_u64(o + 72 * i, &r);
    ok[i] = good ? 1 : 0;
}
// family 4: Fq12 on four lanes (csrc/quad.cuh); ops as family 3, 13 also returns the extra product b.c1.c0 * b.c1.c2 in o2 (12 x u64)
__global__ void __launch_bounds__(64) k_test_quad12(int op, uint64_t arg, const uint64_t *a, const uint64_t *b, uint64_t *o, uint64_t *o2,
                                                    uint8_t *ok, size_t n) {
    using namespace quad;
    const size_t first = ((size_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u)) >> 2;
    if (first >= n) return;
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const bool active = i < n;
    if (!active) i = n - 1;
    q6 x, y, r;
    q12_load(&x, a + 72 * i); q12_load(&y, b + 72 * i);
    r = x;
    bool good[1] = {true};
    switch (op) {
        case 0: q12_mul(&r, &x, &y); break;
        case 3: q12_sqr(&r); break;
        case 6: q12_inv(&r, &x, good); break;
        case 7: q12_frobenius(&r, &x, (int)arg); break;
        case 12: q12_conj(&r, &x); break;
        case 13: {
            qfp d0, d1, d4, ea, eb, eo;
            const uint64_t *bb = b + 72 * i;
            q_load(d0, bb, 0, 1, 0); q_load(d1, bb, 2, 1, 0); q_load(d4, bb, 8, 1, 0);
            q_load(ea, bb, 6, 1, 0); q_load(eb, bb, 10, 1, 0);
            q12_mul_by_014(&r, &d0, &d1, &d4, &ea, &eb, &eo);
            if (active && o2 && (threadIdx.x & 2u) == 0) q_store(o2 + 12 * i, eo, 0, 1, 0);
            break;
        }
        case 15: q12_cyc_sqr(&r); break;
        case 16: q_exp_by_x(&r, &x, arg); break;
        case 17: q_exp_by_x_gs(&r, &x, arg); break;
        case 18: {   // arg compressed squarings, decompressed with their own inversion: equals arg cyclotomic squarings
            qfp G[2], C[4], d, g0, g1;
            q_cyc_compress(G, &x);
            for (uint64_t k = 0; k < arg; k++) q_cyc_sqr_compressed(G);
            q_cyc_gather(C, G);
            qv_dbl(&d, &C[0], 1); qv_dbl(&d, &d, 1);
            q2_inv(&d, &d);
            q_cyc_decompress(&g0, &g1, C, &d);
            q_cyc_place(&r, &g0, &g1, C);
            break;
        }
        case 20: { bool bad[QL]; q_exp_by_x_main(&r, bad, &x, 0xd201000000010000ULL, (int)arg); break; }
        case 19: {   // compress / gather / place only: the compressed coordinates in their slots, g0 = g1 = 0
            qfp G[2], C[4], z;
            q_cyc_compress(G, &x);
            for (uint64_t k = 0; k < arg; k++) q_cyc_sqr_compressed(G);
            q_cyc_gather(C, G);
            q_set_zero(z);
            q_cyc_place(&r, &z, &z, C);
            break;
        }
    }
    if (active) {
        q12_store(o + 72 * i, &r);
        if ((threadIdx.x & 3u) == 0) ok[i] = good[0] ? 1 : 0;
    }
}

// family 7: Fq12 on two lanes (csrc/duo.cuh); ops as family 3
__global__ void __launch_bounds__(64) k_test_duo12(int op, uint64_t arg, const uint64_t *a, const uint64_t *b, uint64_t *o, uint8_t *ok, size_t n) {
    using namespace duo;
    const size_t first = ((size_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u)) >> 1;
    if (first >= n) return;
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    const bool active = i < n;
    if (!active) i = n - 1;
    d12 x, y, r;
    d12_load(&x, a + 72 * i); d12_load(&y, b + 72 * i);
    r = x;
    dflag good;
    good.on[0] = true;
    switch (op) {
        case 0: d12_mul(&r, &x, &y); break;
        case 3: d12_sqr(&r, &x); break;
        case 6: d12_inv(&r, &x, good); break;
        case 7: d12_frobenius(&r, &x, (int)arg); break;
        case 12: d12_conj(&r, &x); break;
        case 13: d12_mul_by_014(&r, &y.c0.c0, &y.c0.c1, &y.c1.c1); break;
        case 15: d12_cyc_sqr(&r, &x); break;
        case 16: d_exp_by_x(&r, &x, arg); break;
        case 17: d_exp_by_x_gs(&r, &x, arg); break;
    }
    if (active) {
        d12_store(o + 72 * i, &r);
        if ((threadIdx.x & 1u) == 0) ok[i] = good.on[0] ? 1 : 0;
    }
}
// families 5 / 6: the group law on G1 / G2 as the sum and MSM kernels use it (XYZZ accumulators, csrc/curve.cuh), on affine inputs
// (g1.go:400-559, g2.go:446-606).  op: 0 a + b (mixed addition into an accumulator holding a, incl. a = b, a = -b, infinities)
// 1 2a 2 (a + b) + (a + b) through the full XYZZ addition 3 a + b through the Jacobian mixed addition 4 2a through the
// Jacobian doubling.  Output: normalised Jacobian.
template <class F, class APOD, class JPOD>
__global__ void k_test_group(int op, const APOD *a, const APOD *b, JPOD *o, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    typename F::T ax, ay, bx, by, ox, oy, oz;
    bool ainf, binf;
    load_affine(ax, ay, ainf, a + i); load_affine(bx, by, binf, b + i);
    xyzz<F> p, q;
    xyzz_set_inf(p);
    if (!ainf) xyzz_madd(p, ax, ay);
    switch (op) {
        case 0: if (!binf) xyzz_madd(p, bx, by); break;
        case 1: xyzz_dbl(p); break;
        case 2: if (!binf) xyzz_madd(p, bx, by); q = p; xyzz_add(p, q); break;
        case 3: {
            jac_pt<F> j;
            xyzz_to_jac_pt(j, p);
            if (!binf) jac_madd(j, bx, by);
            jac_to_xyzz(p, j);
            break;
        }
        case 4: { jac_pt<F> j; xyzz_to_jac_pt(j, p); jac_dbl(j); jac_to_xyzz(p, j); break; }
    }
    xyzz_to_jac_normalised(ox, oy, oz, p);
    store_jac(o + i, ox, oy, oz);
}
#endif
}  // namespace b381
