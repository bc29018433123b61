// tower.cuh -- Fq2 = Fq[u]/(u^2+1), Fq6 = Fq2[v]/(v^3-(1+u)), Fq12 = Fq6[w]/(w^2-v) on the device.
// Replaces the reference's L2 layer on the hot path (fq2.go, fq6.go, fq12.go); every function
// cites the reference operation whose value it reproduces.  Values are exact field elements in
// canonical Montgomery form, so any correct formula yields the reference's bits.
//
// Layout: elements live in per-thread local memory (L1/L2 resident) as the reference's flattened
// structs (fq12 = c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2; fq2 = c0 || c1; 48 B per Fq).  Fq2
// multiplication/squaring are the call granularity: operands are pulled into registers with
// 128-bit loads, ~900 IMAD.WIDE of work are done on them, and the result is written back.
#pragma once
#include "fp.cuh"
#include "constants.inc"

namespace b381 {

struct fp2 { fp c0, c1; };
struct fp6 { fp2 c0, c1, c2; };
struct fp12 { fp6 c0, c1; };
// The Fq coefficients of a tower element as an array, taken from the WHOLE object.  `&x->c0` followed by indexing past that
// member is undefined behaviour, and cicc 12.9 exploits it: a callee that receives `&x->c0` is assumed to touch the first
// coefficient only, so the caller keeps stale copies of the others in registers (device-only wrong values as soon as such a
// call is out of line in a function that also reads the members; profiles/r02_experiments.md).
template <class T> HD fp *fp_array(T *x) { return reinterpret_cast<fp *>(x); }
template <class T> HD const fp *fp_array(const T *x) { return reinterpret_cast<const fp *>(x); }

// C-ABI PODs (include/b381.h): Go's G1Affine / G2Affine structs incl. padding (g1.go:10-14, g2.go:12-16)
struct g1_affine_pod { uint64_t x[6], y[6]; uint8_t inf; uint8_t pad[7]; };
struct g2_affine_pod { uint64_t x[12], y[12]; uint8_t inf; uint8_t pad[7]; };

HD void fp_load_u64(fp &r, const uint64_t *p) {
#pragma unroll
    for (int i = 0; i < 6; i++) { uint64_t v = p[i]; r.l[2 * i] = (uint32_t)v; r.l[2 * i + 1] = (uint32_t)(v >> 32); }
}
HD void fp_store_u64(uint64_t *p, const fp &a) {
#pragma unroll
    for (int i = 0; i < 6; i++) p[i] = (uint64_t)a.l[2 * i] | ((uint64_t)a.l[2 * i + 1] << 32);
}

struct g1_jac_pod { uint64_t x[6], y[6], z[6]; };
struct g2_jac_pod { uint64_t x[12], y[12], z[12]; };

// ---- constant tables -------------------------------------------------------------------------
#if defined(__CUDACC__)
__device__ __constant__ uint32_t d_frob6_c1[6 * 24] = B381_FROB6_C1_INIT;
__device__ __constant__ uint32_t d_frob6_c2[6 * 24] = B381_FROB6_C2_INIT;
__device__ __constant__ uint32_t d_frob12_c1[12 * 24] = B381_FROB12_C1_INIT;
__device__ __constant__ uint32_t d_q_minus_2[12] = {B381_Q_MINUS_2_LIMBS};
#endif
#if !defined(__CUDA_ARCH__)
static const uint32_t h_frob6_c1[6 * 24] = B381_FROB6_C1_INIT;
static const uint32_t h_frob6_c2[6 * 24] = B381_FROB6_C2_INIT;
static const uint32_t h_frob12_c1[12 * 24] = B381_FROB12_C1_INIT;
static const uint32_t h_q_minus_2[12] = {B381_Q_MINUS_2_LIMBS};
#endif
#if defined(__CUDA_ARCH__)
#define B381_TAB(name) d_##name
#else
#define B381_TAB(name) h_##name
#endif

HD void fp_load_tab(fp &r, const uint32_t *t) {
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = t[i];
}

// ---- pointer forms (kept for callers that hold operands in memory) -------------------------------
HD void fp_mul_n(fp *r, const fp *a, const fp *b) { *r = fp_mul_v(*a, *b); }
HD void fp_sqr_n(fp *r, const fp *a) { fp x = *a; *r = fp_mul_v(x, x); }

// ---- vector forms of the cheap operations: one out-of-line loop serves Fq2 (n=2), Fq6 (n=6) and
// Fq12 (n=12) callers, so an addition costs a call instead of 37 inlined instructions per Fq ------
// n is even everywhere (Fq2 = 2, Fq6 = 6, Fq12 = 12): two elements per iteration, so that the four loads are in
// flight together and the two carry chains interleave (the one-element loop was latency-bound on one load pair
// and one chain).  One body for add / sub / double keeps the code small: the hot set of a kernel has to stay inside
// the 32 KB L1.5 instruction cache (profiles/r01_v4_experiments.md).
HDN void fpv_addsub(fp *r, const fp *a, const fp *b, int n, int sub) {
#pragma unroll 1
    for (int i = 0; i < n; i += 2) {
        fp x0 = a[i], y0 = b[i], x1 = a[i + 1], y1 = b[i + 1];
        if (sub) { fp_sub(x0, x0, y0); fp_sub(x1, x1, y1); }
        else { fp_add(x0, x0, y0); fp_add(x1, x1, y1); }
        r[i] = x0; r[i + 1] = x1;
    }
}
HD void fpv_add(fp *r, const fp *a, const fp *b, int n) { fpv_addsub(r, a, b, n, 0); }
HD void fpv_sub(fp *r, const fp *a, const fp *b, int n) { fpv_addsub(r, a, b, n, 1); }
HD void fpv_dbl(fp *r, const fp *a, int n) { fpv_addsub(r, a, a, n, 0); }
HDN void fpv_neg(fp *r, const fp *a, int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) { fp x = a[i]; fp_neg(x, x); r[i] = x; }
}

// a^(Q-2): Fermat inversion with a fixed chain (570 multiplications).  Kept as the cross-check of fp_inv in the tests and
// behind B381_INV_FERMAT for A/B timing.
HDN void fp_inv_fermat(fp *r, const fp *a) {
    fp x = *a, acc;
    fp_set_one(acc);
    const uint32_t *e = B381_TAB(q_minus_2);
    bool started = false;
#pragma unroll 1
    for (int i = 380; i >= 0; i--) {
        if (started) fp_sqr(acc, acc);
        if ((e[i >> 5] >> (i & 31)) & 1) {
            if (started) fp_mul(acc, acc, x);
            else { acc = x; started = true; }
        }
    }
    *r = acc;
}
// FQ.Inverse (fq.go:224-266 is a binary Euclid as well; the value is canonical, so any correct method is bit-identical);
// 0 -> 0 here (the reference returns "no inverse").  Kaliski's almost-inverse on the twelve limbs of the Montgomery
// representative A, with the factors of two stripped in batches:
//   invariants  A r = -+ u 2^k,  A s = +- v 2^k  (mod Q),  u s + v r = Q  (so r, s <= Q: twelve limbs suffice)
//   v even: v >>= z, r <<= z, k += z        v < u: exchange (u, r) with (v, s), which flips both signs        v -= u, s += r
// until v = 0, u = 1: A^-1 = -+ r 2^-k with 381 <= k <= 762.  The Montgomery form of the inverse is A^-1 R^2 = -+ r 2^(768-k):
// two Montgomery products (by R^2 and by the one-bit operand 2^(768-k)).  ~30 k integer instructions against ~190 k
// for the exponentiation; the trip count depends on the value (lanes of a warp differ by a few per cent).
HDN void fp_inv(fp *out, const fp *a) {
#if defined(B381_INV_FERMAT)
    fp_inv_fermat(out, a);
#else
    const uint32_t q[12] = {B381_Q_LIMBS};
    uint32_t u[12], v[12], r[12], s[12], k = 0, neg = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) { u[i] = q[i]; v[i] = a->l[i]; r[i] = 0; s[i] = 0; }
    s[0] = 1;
    for (;;) {
        uint32_t low = v[0];
        if (low == 0) {
            uint32_t any = 0;
#pragma unroll
            for (int i = 1; i < 12; i++) any |= v[i];
            if (any == 0) break;
#pragma unroll
            for (int i = 0; i < 11; i++) { v[i] = v[i + 1]; r[11 - i] = r[10 - i]; }
            v[11] = 0; r[0] = 0;
            k += 32;
            continue;
        }
#if defined(__CUDA_ARCH__)
        uint32_t z = __ffs(low) - 1;
#else
        uint32_t z = (uint32_t)__builtin_ctz(low);
#endif
        if (z) {
#pragma unroll
            for (int i = 0; i < 11; i++) v[i] = (v[i] >> z) | (v[i + 1] << (32 - z));
            v[11] >>= z;
#pragma unroll
            for (int i = 11; i > 0; i--) r[i] = (r[i] << z) | (r[i - 1] >> (32 - z));
            r[0] <<= z;
            k += z;
        }
        uint32_t d[12];
        uint64_t br = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            uint64_t x = (uint64_t)v[i] - u[i] - br;
            d[i] = (uint32_t)x;
            br = (x >> 32) & 1u;
        }
        if (br) {
            uint64_t c = 1;
#pragma unroll
            for (int i = 0; i < 12; i++) {
                u[i] = v[i];
                uint64_t x = (uint64_t)(~d[i]) + c;
                d[i] = (uint32_t)x;
                c = x >> 32;
                uint32_t t = r[i]; r[i] = s[i]; s[i] = t;
            }
            neg ^= 1u;
        }
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            v[i] = d[i];
            uint64_t x = (uint64_t)s[i] + r[i] + c;
            s[i] = (uint32_t)x;
            c = x >> 32;
        }
    }
    if (k == 0) { fp_set_zero(*out); return; }         // A = 0 (for 0 < A < Q Kaliski's bound gives 381 <= k <= 762)
    const uint32_t r2[12] = {B381_R2_RAW_LIMBS};
    fp t, w;
#pragma unroll
    for (int i = 0; i < 12; i++) { t.l[i] = r[i]; w.l[i] = r2[i]; }
    fp_mul(t, t, w);                                   // r R
    uint32_t e = 768u - k;                             // 6 <= e <= 387
    if (e > 380u) {                                    // 2^e would not be a reduced operand: go through r R^2 and two factors
        fp_mul(t, t, w);
        fp_set_zero(w); w.l[190 >> 5] = 1u << (190 & 31);
        fp_mul(t, t, w);
        e -= 190u;
    }
    fp_set_zero(w);
    w.l[e >> 5] = 1u << (e & 31);
    fp_mul(t, t, w);                                   // r 2^e
    if (!neg) fp_neg(t, t);
    *out = t;
#endif
}

// ---- Fq2 (fq2.go) ----------------------------------------------------------------------------
// three-pointer entry points: ~170 call sites in the Miller loop each save the two constant arguments (length, mode) --
// the kernel sits at the instruction-cache limit and call-site code is most of what is not a multiplier
HDN void fp2_add_p(fp2 *r, const fp2 *a, const fp2 *b) { fpv_addsub(fp_array(r), fp_array(a), fp_array(b), 2, 0); }
HDN void fp2_sub_p(fp2 *r, const fp2 *a, const fp2 *b) { fpv_addsub(fp_array(r), fp_array(a), fp_array(b), 2, 1); }
HD void fp2_add(fp2 &r, const fp2 &a, const fp2 &b) { fp2_add_p(&r, &a, &b); }            // fq2.go:104-107
HD void fp2_sub(fp2 &r, const fp2 &a, const fp2 &b) { fp2_sub_p(&r, &a, &b); }            // fq2.go:110-113
HD void fp2_dbl(fp2 &r, const fp2 &a) { fp2_add_p(&r, &a, &a); }                          // fq2.go:92-95
HD void fp2_neg(fp2 &r, const fp2 &a) { fpv_neg(fp_array(&r), fp_array(&a), 2); }                       // fq2.go:98-101
HD void fp2_conj(fp2 &r, const fp2 &a) { r.c0 = a.c0; fp_neg(r.c1, a.c1); }
HD bool fp2_is_zero(const fp2 &a) { return fp_is_zero(a.c0) && fp_is_zero(a.c1); }
HD bool fp2_eq(const fp2 &a, const fp2 &b) { return fp_eq(a.c0, b.c0) && fp_eq(a.c1, b.c1); }
HD void fp2_set_zero(fp2 &r) { fp_set_zero(r.c0); fp_set_zero(r.c1); }
HD void fp2_set_one(fp2 &r) { fp_set_one(r.c0); fp_set_zero(r.c1); }
// r = a * (1 + u)   (fq2.go:41-45)
HDN void fp2_mul_nr_p(fp2 *r, const fp2 *a) {
    fp a0 = a->c0, a1 = a->c1, t0, t1;
    fp_sub(t0, a0, a1);
    fp_add(t1, a0, a1);
    r->c0 = t0; r->c1 = t1;
}
HD void fp2_mul_nr(fp2 &r, const fp2 &a) { fp2_mul_nr_p(&r, &a); }

// r = a * b   (fq2.go:116-130: same value; the two rows are two-product dot products with one reduction each,
// 888 wide MACs and no Karatsuba fix-up additions instead of 900 + five additions)
HDN void fp2_mul(fp2 *r, const fp2 *a, const fp2 *b) {
#ifdef B381_FP2_MUL_KARATSUBA
    {   // A/B: three plain products (the two-product body leaves the hot instruction set); measured slower, profiles/r02_experiments.md
        fp a0 = a->c0, a1 = a->c1, b0 = b->c0, b1v = b->c1, v0, v1, s, t;
        fp_add_nr(s, a0, a1); fp_add_nr(t, b0, b1v);
        fp_mul(v0, a0, b0); fp_mul(v1, a1, b1v); fp_mul(s, s, t);
        fp_sub(s, s, v0); fp_sub(s, s, v1);
        fp_sub(v0, v0, v1);
        r->c0 = v0; r->c1 = s;
        return;
    }
#endif
    fp nb1, b1 = b->c1, c0;
    fp_qminus(nb1, b1);
    fp_dot2_p(&c0, &a->c0, &b->c0, &a->c1, &nb1);      // c0 is a temporary: r may alias a or b
    fp_dot2_p(&r->c1, &a->c0, &b->c1, &a->c1, &b->c0);
    r->c0 = c0;
}
// r = a^2   (fq2.go:75-89; complex squaring, 2 Fq mul; the operand sums stay unreduced, below 2Q)
HDN void fp2_sqr(fp2 *r, const fp2 *a) {
#ifdef B381_FP2_SQR_DOT2
    { fp2_mul(r, a, a); return; }                      // A/B: the plain multiplier leaves the hot instruction set (+288 wide MACs per squaring)
#endif
    fp a0 = a->c0, a1 = a->c1, s, d, na1, t;
    fp_add_nr(s, a0, a1);
    fp_qminus(na1, a1);
    fp_add_nr(d, a0, na1);
    fp_add_nr(t, a1, a1);
    fp_mul(s, s, d);
    fp_mul(t, a0, t);
    r->c0 = s; r->c1 = t;
}
// r = a * s with s in Fq   (the c0.c0/c0.c1 *= p.y scaling of pairing.go:33-36)
HDN void fp2_mul_fp(fp2 *r, const fp2 *a, const fp *s) {
    fp a0 = a->c0, a1 = a->c1, k = *s;
    fp_mul(a0, a0, k);
    fp_mul(a1, a1, k);
    r->c0 = a0; r->c1 = a1;
}
// r = a^-1, 0 -> 0   (fq2.go:133-147)
HDN void fp2_inv(fp2 *r, const fp2 *a) {
    fp t0, t1;
    fp_sqr_n(&t0, &a->c0);
    fp_sqr_n(&t1, &a->c1);
    fp_add(t0, t0, t1);
    fp_inv(&t0, &t0);
    fp_mul_n(&r->c0, &a->c0, &t0);
    fp_mul_n(&t1, &a->c1, &t0);
    fp_neg(r->c1, t1);
}

// ---- Fq6 (fq6.go) ------------------------------------------------------------------------------
HD void fp6_add(fp6 *r, const fp6 *a, const fp6 *b) { fpv_add(fp_array(r), fp_array(a), fp_array(b), 6); }   // fq6.go:123-127
HD void fp6_sub(fp6 *r, const fp6 *a, const fp6 *b) { fpv_sub(fp_array(r), fp_array(a), fp_array(b), 6); }   // fq6.go:130-134
HD void fp6_neg(fp6 *r, const fp6 *a) { fpv_neg(fp_array(r), fp_array(a), 6); }                  // fq6.go:116-120
// r = a * v   (fq6.go:34-37)
HD void fp6_mul_nr(fp6 *r, const fp6 *a) {
    fp2 t = a->c2, c0 = a->c0, c1 = a->c1;
    fp2_mul_nr(t, t);
    r->c0 = t; r->c1 = c0; r->c2 = c1;
}
HD bool fp6_is_zero(const fp6 *a) { return fp2_is_zero(a->c0) && fp2_is_zero(a->c1) && fp2_is_zero(a->c2); }

// r = a * b   (fq6.go:255-292; 6 Fq2 mul)
HDN void fp6_mul(fp6 *r, const fp6 *a, const fp6 *b) {
    fp2 v0, v1, v2, s, t, x;
    fp2_mul(&v0, &a->c0, &b->c0);
    fp2_mul(&v1, &a->c1, &b->c1);
    fp2_mul(&v2, &a->c2, &b->c2);
    // c0 = v0 + xi*((a1+a2)(b1+b2) - v1 - v2)
    fp2_add(s, a->c1, a->c2);
    fp2_add(t, b->c1, b->c2);
    fp2_mul(&x, &s, &t);
    fp2_sub(x, x, v1);
    fp2_sub(x, x, v2);
    fp2_mul_nr(x, x);
    fp2_add(x, x, v0);
    // c1 = (a0+a1)(b0+b1) - v0 - v1 + xi*v2
    fp2 y;
    fp2_add(s, a->c0, a->c1);
    fp2_add(t, b->c0, b->c1);
    fp2_mul(&y, &s, &t);
    fp2_sub(y, y, v0);
    fp2_sub(y, y, v1);
    fp2_mul_nr(s, v2);
    fp2_add(y, y, s);
    // c2 = (a0+a2)(b0+b2) - v0 - v2 + v1: the last use of a and b, so it can go straight to r->c2 even when r aliases them
    fp2_add(s, a->c0, a->c2);
    fp2_add(t, b->c0, b->c2);
    fp2_mul(&r->c2, &s, &t);
    fp2_sub(r->c2, r->c2, v0);
    fp2_sub(r->c2, r->c2, v2);
    fp2_add(r->c2, r->c2, v1);
    r->c0 = x; r->c1 = y;
}
// r = a * (b0 + b1 v)   (fq6.go:60-90; 5 Fq2 mul)
HDN void fp6_mul_by_01(fp6 *r, const fp6 *a, const fp2 *b0, const fp2 *b1) {
    fp2 v0, v1, s, t, x, y;
    fp2_mul(&v0, &a->c0, b0);
    fp2_mul(&v1, &a->c1, b1);
    // c0 = v0 + xi*((a1+a2)*b1 - v1)
    fp2_add(s, a->c1, a->c2);
    fp2_mul(&x, &s, b1);
    fp2_sub(x, x, v1);
    fp2_mul_nr(x, x);
    fp2_add(x, x, v0);
    // c1 = (a0+a1)(b0+b1) - v0 - v1
    fp2_add(s, a->c0, a->c1);
    fp2_add(t, *b0, *b1);
    fp2_mul(&y, &s, &t);
    fp2_sub(y, y, v0);
    fp2_sub(y, y, v1);
    // c2 = (a0+a2)*b0 - v0 + v1 (last use of a: straight to r->c2)
    fp2_add(s, a->c0, a->c2);
    fp2_mul(&r->c2, &s, b0);
    fp2_sub(r->c2, r->c2, v0);
    fp2_add(r->c2, r->c2, v1);
    r->c0 = x; r->c1 = y;
}
// r = a * (b1 v)   (fq6.go:40-57; 3 Fq2 mul)
HDN void fp6_mul_by_1(fp6 *r, const fp6 *a, const fp2 *b1) {
    fp2 x, y, z;
    fp2_mul(&x, &a->c2, b1);
    fp2_mul_nr(x, x);
    fp2_mul(&y, &a->c0, b1);
    fp2_mul(&z, &a->c1, b1);
    r->c0 = x; r->c1 = y; r->c2 = z;
}
// r = a^-1, 0 -> 0   (fq6.go:295-336)
HDN void fp6_inv(fp6 *r, const fp6 *a) {
    fp2 k0, k1, k2, t, u;
    // k0 = a0^2 - xi*a1*a2
    fp2_sqr(&k0, &a->c0);
    fp2_mul(&t, &a->c1, &a->c2);
    fp2_mul_nr(t, t);
    fp2_sub(k0, k0, t);
    // k1 = xi*a2^2 - a0*a1
    fp2_sqr(&k1, &a->c2);
    fp2_mul_nr(k1, k1);
    fp2_mul(&t, &a->c0, &a->c1);
    fp2_sub(k1, k1, t);
    // k2 = a1^2 - a0*a2
    fp2_sqr(&k2, &a->c1);
    fp2_mul(&t, &a->c0, &a->c2);
    fp2_sub(k2, k2, t);
    // t = a0*k0 + xi*(a2*k1 + a1*k2)
    fp2_mul(&t, &a->c2, &k1);
    fp2_mul(&u, &a->c1, &k2);
    fp2_add(t, t, u);
    fp2_mul_nr(t, t);
    fp2_mul(&u, &a->c0, &k0);
    fp2_add(t, t, u);
    fp2_inv(&t, &t);
    fp2_mul(&r->c0, &k0, &t);
    fp2_mul(&r->c1, &k1, &t);
    fp2_mul(&r->c2, &k2, &t);
}
// r = a^(q^power), power in {1,2,3}   (fq6.go:211-218)
HDN void fp6_frobenius(fp6 *r, const fp6 *a, int power) {
    fp2 c0 = a->c0, c1 = a->c1, c2 = a->c2, k;
    if (power & 1) { fp2_conj(c0, c0); fp2_conj(c1, c1); fp2_conj(c2, c2); }   // fq2.go:156-158
    r->c0 = c0;
    fp_load_tab(k.c0, B381_TAB(frob6_c1) + power * 24); fp_load_tab(k.c1, B381_TAB(frob6_c1) + power * 24 + 12);
    fp2_mul(&r->c1, &c1, &k);
    fp_load_tab(k.c0, B381_TAB(frob6_c2) + power * 24); fp_load_tab(k.c1, B381_TAB(frob6_c2) + power * 24 + 12);
    fp2_mul(&r->c2, &c2, &k);
}

// ---- Fq12 (fq12.go) ----------------------------------------------------------------------------
HD void fp12_set_one(fp12 *r) {
    fp *p = fp_array(r);
    fp z; fp_set_zero(z);
#pragma unroll 1
    for (int i = 1; i < 12; i++) p[i] = z;
    fp_set_one(z); p[0] = z;
}
HD void fp12_copy(fp12 *r, const fp12 *a) {
    const fp *pa = fp_array(a); fp *pr = fp_array(r);
#pragma unroll 1
    for (int i = 0; i < 12; i++) { fp x = pa[i]; pr[i] = x; }
}
HD void fp12_conj(fp12 *r, const fp12 *a) {   // fq12.go:27-29
    if (r != a) r->c0 = a->c0;
    fp6_neg(&r->c1, &a->c1);
}
HD bool fp12_is_one(const fp12 *a) {           // fq12.go:56-58 against FQ12One
    const fp *p = fp_array(a);
    fp one; fp_set_one(one);
    bool ok = fp_eq(p[0], one);
#pragma unroll 1
    for (int i = 1; i < 12; i++) ok = ok && fp_is_zero(p[i]);
    return ok;
}
HD bool fp12_is_zero(const fp12 *a) {
    const fp *p = fp_array(a);
    bool z = true;
#pragma unroll 1
    for (int i = 0; i < 12; i++) z = z && fp_is_zero(p[i]);
    return z;
}
// r = a + v*b and r = a - v*b with v*b = (xi b2, b0, b1) (fq6.go:34-37) formed on the fly: no copy of the rotated element
HD void fp6_add_vmul(fp6 *r, const fp6 *a, const fp6 *b) {
    fp2 x;
    fp2_mul_nr(x, b->c2);
    fp2_add(r->c0, a->c0, x);
    fp2_add(r->c1, a->c1, b->c0);
    fp2_add(r->c2, a->c2, b->c1);
}
HD void fp6_sub_vmul(fp6 *r, const fp6 *a, const fp6 *b) {
    fp2 x;
    fp2_mul_nr(x, b->c2);
    fp2_sub(r->c0, a->c0, x);
    fp2_sub(r->c1, a->c1, b->c0);
    fp2_sub(r->c2, a->c2, b->c1);
}
// r = a * b   (fq12.go:198-213; 3 Fq6 mul)
HDN void fp12_mul(fp12 *r, const fp12 *a, const fp12 *b) {
    fp6 aa, bb, s, t;
    fp6_mul(&aa, &a->c0, &b->c0);
    fp6_mul(&bb, &a->c1, &b->c1);
    fp6_add(&s, &a->c0, &a->c1);
    fp6_add(&t, &b->c0, &b->c1);
    fp6_mul(&s, &s, &t);
    fp6_sub(&s, &s, &aa);
    fp6_sub(&r->c1, &s, &bb);
    fp6_mul_nr(&bb, &bb);                  // (the on-the-fly form of fp12_sqr / fp12_mul_by_014 measured slower here: this
    fp6_add(&r->c0, &bb, &aa);             // function runs in the final-exponentiation kernel, those in the Miller loop)
}
// r = a^2   (fq12.go:180-195; complex squaring, 2 Fq6 mul)
HDN void fp12_sqr(fp12 *r, const fp12 *a) {
    fp6 ab, s, t;
    fp6_mul(&ab, &a->c0, &a->c1);
    fp6_add(&s, &a->c0, &a->c1);
    fp6_add_vmul(&t, &a->c0, &a->c1);
    fp6_mul(&s, &s, &t);
    fp6_sub(&s, &s, &ab);
    fp6_add(&r->c1, &ab, &ab);
    fp6_sub_vmul(&r->c0, &s, &ab);
}
// f *= (d0 + d1 v) + (d4 v) w   (fq12.go:32-47; 13 Fq2 mul)
HDN void fp12_mul_by_014(fp12 *f, const fp2 *d0, const fp2 *d1, const fp2 *d4) {
    fp6 aa, bb, s;
    fp2 o;
    fp6_mul_by_01(&aa, &f->c0, d0, d1);
    fp6_mul_by_1(&bb, &f->c1, d4);
    fp2_add(o, *d1, *d4);
    fp6_add(&s, &f->c1, &f->c0);
    fp6_mul_by_01(&s, &s, d0, &o);
    fp6_sub(&s, &s, &aa);
    fp6_sub(&f->c1, &s, &bb);
    fp6_add_vmul(&f->c0, &aa, &bb);
}
// r = a^-1; returns false (and leaves r untouched) for a == 0   (fq12.go:216-237)
HDN bool fp12_inv(fp12 *r, const fp12 *a) {
    fp6 t0, t1;
    fp6_mul(&t0, &a->c0, &a->c0);
    fp6_mul(&t1, &a->c1, &a->c1);
    fp6_mul_nr(&t1, &t1);
    fp6_sub(&t0, &t0, &t1);
    if (fp6_is_zero(&t0)) return false;
    fp6_inv(&t0, &t0);
    fp6_mul(&t1, &a->c1, &t0);
    fp6_mul(&r->c0, &a->c0, &t0);
    fp6_neg(&r->c1, &t1);
    return true;
}
// r = a^(q^power), power in {1,2,3}   (fq12.go:171-177)
HDN void fp12_frobenius(fp12 *r, const fp12 *a, int power) {
    fp6_frobenius(&r->c0, &a->c0, power);
    fp6_frobenius(&r->c1, &a->c1, power);
    fp2 k;
    fp_load_tab(k.c0, B381_TAB(frob12_c1) + power * 24); fp_load_tab(k.c1, B381_TAB(frob12_c1) + power * 24 + 12);
    fp2_mul(&r->c1.c0, &r->c1.c0, &k);
    fp2_mul(&r->c1.c1, &r->c1.c1, &k);
    fp2_mul(&r->c1.c2, &r->c1.c2, &k);
}

// Squaring in the cyclotomic subgroup (Granger-Scott): valid after the easy part of the final
// exponentiation (pairing.go:80-90).  Same value as fq12.go:180-195 on such inputs at 18 Fq mul
// instead of 36; the reference squares generically inside FQ12.Exp (fq12.go:108-120).
HD void fp4_sqr(fp2 &o0, fp2 &o1, const fp2 &a, const fp2 &b) {
    fp2 t1;                       // the outputs double as the other two temporaries: 96 B of hot stack instead of 288 B
    fp2_sqr(&o0, &a);
    fp2_sqr(&t1, &b);
    // (direct calls of the add/sub loop: this runs in the final-exponentiation kernel, where the extra hop through the
    // three-pointer entry points costs more than their smaller call sites save)
    fpv_add(fp_array(&o1), fp_array(&a), fp_array(&b), 2);
    fp2_sqr(&o1, &o1);
    fpv_sub(fp_array(&o1), fp_array(&o1), fp_array(&o0), 2);
    fpv_sub(fp_array(&o1), fp_array(&o1), fp_array(&t1), 2);          // 2ab
    fp2_mul_nr(t1, t1);
    fpv_add(fp_array(&o0), fp_array(&o0), fp_array(&t1), 2);          // a^2 + xi b^2
}
// r = 3t - 2z (plus = 0) or 3t + 2z (plus = 1): the six output rows of the cyclotomic squaring as ONE pass each (three
// separate add/sub/double passes cost three local-memory round trips per row)
HDN void fp2_tri(fp2 *r, const fp2 *t, const fp2 *z, int plus) {
    const fp *tp = fp_array(t), *zp = fp_array(z);
    fp *rp = fp_array(r);
#pragma unroll 1
    for (int i = 0; i < 2; i++) {
        fp x = tp[i], y = zp[i], u;
        if (plus) fp_add(u, x, y); else fp_sub(u, x, y);
        fp_add(u, u, u);
        fp_add(u, u, x);
        rp[i] = u;
    }
}
HDN void fp12_cyclotomic_sqr(fp12 *r, const fp12 *a) {
    // z0 = a.c0.c0, z4 = a.c0.c1, z3 = a.c0.c2, z2 = a.c1.c0, z1 = a.c1.c1, z5 = a.c1.c2.  In place (r == a) without copies:
    // every output row is written to the slot its own z operand is read from (fp2_tri reads before it writes), and the two
    // Fp4 squarings whose outputs cross (z2,z3 -> z4',z5' and z4,z5 -> z2',z3') are both finished before either is written.
    fp2 t0, t1, t2, t3;
    // (t0,t1) = fp4sq(z0,z1);  z0' = 3t0 - 2z0;  z1' = 3t1 + 2z1
    fp4_sqr(t0, t1, a->c0.c0, a->c1.c1);
    fp2_tri(&r->c0.c0, &t0, &a->c0.c0, 0);
    fp2_tri(&r->c1.c1, &t1, &a->c1.c1, 1);
    // (t0,t1) = fp4sq(z2,z3); (t2,t3) = fp4sq(z4,z5)
    fp4_sqr(t0, t1, a->c1.c0, a->c0.c2);
    fp4_sqr(t2, t3, a->c0.c1, a->c1.c2);
    // z4' = 3t0 - 2z4;  z5' = 3t1 + 2z5
    fp2_tri(&r->c0.c1, &t0, &a->c0.c1, 0);
    fp2_tri(&r->c1.c2, &t1, &a->c1.c2, 1);
    // z2' = 3 xi t3 + 2z2;  z3' = 3t2 - 2z3
    fp2_mul_nr(t3, t3);
    fp2_tri(&r->c1.c0, &t3, &a->c1.c0, 1);
    fp2_tri(&r->c0.c2, &t2, &a->c0.c2, 0);
}

}  // namespace b381
