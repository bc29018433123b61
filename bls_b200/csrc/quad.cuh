// quad.cuh -- the pairing on FOUR LANES per pairing: the throughput form of G2AffineToPrepared + MillerLoop +
// FinalExponentiation (g2.go:650-801, pairing.go:16-129) for large batches.
//
// Why: with one pairing per thread (pairing.cuh) the Fq12 working set (2.8 / 5.7 KB of stack per thread, 2^16 threads)
// lives in L2 / DRAM -- ncu showed 30-46 GB of local-memory traffic per launch and `long_scoreboard` as the second stall
// (profiles/r01_v6_ncu_summary.md).  A B200 SM has 256 KB of registers + 228 KB of L1/shared memory for the 443
// pairings of a 2^16 batch that land on it: 1.1 KB each, less than the state of ONE Miller loop.  So the batch has to
// run in rounds with fewer pairings resident, and to keep 14 warps per SM busy each pairing has to spread over lanes.
//
// Layout: lane (h, j) of a quad, h = lane bit 1, j = lane bit 0.  Every Fq2 value x = x0 + x1 u lives as ONE Fq per
// lane: lane j holds x_j.  The two halves h = 0 / 1 of a quad run the SAME instruction stream on DIFFERENT Fq2
// operations (two independent products of the tower's Karatsuba levels per round), and an Fq12 value f = c0 + c1 w is
// split by halves: half h holds the Fq6 coefficient c_h (three Fq per lane).  Per lane the state is a quarter of the
// thread form's: ~0.7 KB of stack of which ~0.35 KB is hot -- 448 lanes x 0.35 KB fits the L1 of an SM.
//   * an Fq2 product needs the partner lane's coefficients: 24 SHFL per 444 wide MACs (lane j = 0 computes
//     a0 b0 + a1 (Q - b1), lane j = 1 computes a1 b0 + a0 b1: one two-product dot product with one reduction per lane);
//   * additions, subtractions, doublings, multiplications by an Fq scalar are lane-local;
//   * multiplication by xi = 1 + u and conjugation touch the partner (12 SHFL) / one lane only;
//   * the halves exchange Fq6 values at the Fq12 level (SHFL xor 2) -- 36 SHFL per exchange.
// The schedule below fills both halves in every multiplier round of the doubling iteration of the Miller loop
// (4 squaring rounds, 1 scaling round, 14 product rounds = exactly the thread form's 25 + 4 + 39 + 36 Fq
// multiplications / 4) and of the compressed cyclotomic squarings that dominate the final exponentiation.
//
// Values are canonical residues throughout, and the step formulas are the reference's (g2.go:655-772), so both the
// Miller value and the final result are bit-identical to pairing.cuh's and to the reference's.
//
// Host build: QL = 4 and every primitive loops over the four lanes of ONE quad, so tests/emu runs this exact logic
// against the oracle on the CPU (tests/test_emu_quad.py).  Composite functions use lane identity only through the
// primitives (qv_selh, q_bcast, mode arguments), never directly.
#pragma once
#include "pairing.cuh"

#ifndef LANE_SQR_DOT2_MIN_BLOCKS
#define LANE_SQR_DOT2_MIN_BLOCKS 800u   // 64-thread blocks in the grid (148 SMs x 5.4) from which the lane kernels square through the two-product body
#endif
namespace b381 {
namespace quad {

#if defined(__CUDA_ARCH__)
#define QL 1
#define QLANE ((int)(threadIdx.x & 3u))
#else
#define QL 4
#define QLANE (l_)
#endif
#define QFOR for (int l_ = 0; l_ < QL; l_++)
#define QJ (QLANE & 1)
#define QH ((QLANE >> 1) & 1)

struct qfp { fp v[QL]; };                 // one Fq per lane
struct q6 { qfp c0, c1, c2; };            // an Fq6 value (coefficient j of each Fq2 coefficient), or half of an Fq12
// Coefficient k of a q6 (or the first of a run of coefficients) as a pointer into the WHOLE object.  `qa(&x, 0)` followed by
// indexing past that member is undefined behaviour which cicc 12.9 exploits (a callee given `qa(&x, 0)` is assumed to touch that
// member only; tower.cuh:fp_array, profiles/r02_experiments.md); every multi-coefficient call below goes through this.
template <class T> HD qfp *qa(T *x, int k = 0) { return reinterpret_cast<qfp *>(x) + k; }
template <class T> HD const qfp *qa(const T *x, int k = 0) { return reinterpret_cast<const qfp *>(x) + k; }

// ---- lane exchange ---------------------------------------------------------------------------------------------
HD void q_shfl(qfp &r, const qfp &a, int mask) {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int i = 0; i < 12; i++) r.v[0].l[i] = __shfl_xor_sync(0xffffffffu, a.v[0].l[i], mask);
#else
    qfp t = a;
    QFOR r.v[l_] = t.v[l_ ^ mask];
#endif
}
// true if the predicate holds on any lane of the warp (host: of the quad)
HD bool q_any(const bool p[QL]) {
#if defined(__CUDA_ARCH__)
    return __any_sync(0xffffffffu, p[0]);
#else
    bool o = false;
    QFOR o = o || p[l_];
    return o;
#endif
}
// per lane: does the predicate hold on all four lanes of the own quad
HD void q_all4(bool r[QL], const bool p[QL]) {
#if defined(__CUDA_ARCH__)
    unsigned b = __ballot_sync(0xffffffffu, p[0]);
    unsigned sh = (threadIdx.x & 31u) & ~3u;
    r[0] = ((b >> sh) & 0xFu) == 0xFu;
#else
    bool o = true;
    QFOR o = o && p[l_];
    QFOR r[l_] = o;
#endif
}

// ---- lane-local vector operations (n Fq per lane) ------------------------------------------------------------------
// mode: 0 add, 1 sub, 2 add on half 0 / sub on half 1, 3 sub on half 0 / add on half 1
HDN void qv_addsub(qfp *r, const qfp *a, const qfp *b, int n, int mode) {
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        QFOR {
            fp x = a[i].v[l_], y = b[i].v[l_];
            const int sub = (mode ^ ((mode >> 1) & QH)) & 1;
            if (sub) fp_sub(x, x, y); else fp_add(x, x, y);
            r[i].v[l_] = x;
        }
    }
}
HD void qv_add(qfp *r, const qfp *a, const qfp *b, int n) { qv_addsub(r, a, b, n, 0); }
HD void qv_sub(qfp *r, const qfp *a, const qfp *b, int n) { qv_addsub(r, a, b, n, 1); }
HD void qv_dbl(qfp *r, const qfp *a, int n) { qv_addsub(r, a, a, n, 0); }
// which: 0 all lanes, 1 lanes with j = 1 (conjugation, fq2.go:156-158), 2 lanes of half 0, 3 lanes of half 1
HDN void qv_neg(qfp *r, const qfp *a, int n, int which) {
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        QFOR {
            fp x = a[i].v[l_];
            const bool on = which == 0 || (which == 1 && QJ) || (which == 2 && !QH) || (which == 3 && QH);
            if (on) fp_neg(x, x);
            r[i].v[l_] = x;
        }
    }
}
// r = half 0 ? a0 : a1
HDN void qv_selh(qfp *r, const qfp *a0, const qfp *a1, int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        QFOR {
            fp x = a0[i].v[l_], y = a1[i].v[l_];
#pragma unroll
            for (int k = 0; k < 12; k++) x.l[k] = QH ? y.l[k] : x.l[k];
            r[i].v[l_] = x;
        }
    }
}
// r = the values held by the other half
HDN void qv_xh(qfp *r, const qfp *a, int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) { qfp x = a[i], y; q_shfl(y, x, 2); r[i] = y; }
}
// both halves computed a value in `own`: r0 = half 0's, r1 = half 1's, on all four lanes
HDN void q_bcast(qfp *r0, qfp *r1, const qfp *own) {
    qfp x = *own, o, a, b;
    q_shfl(o, x, 2);
    QFOR {
#pragma unroll
        for (int k = 0; k < 12; k++) {
            a.v[l_].l[k] = QH ? o.v[l_].l[k] : x.v[l_].l[k];
            b.v[l_].l[k] = QH ? x.v[l_].l[k] : o.v[l_].l[k];
        }
    }
    *r0 = a; *r1 = b;
}
// lane-local Fq products: r[i] = a[i] * s (the Fq2 x Fq scalings of pairing.go:33-36 and the norm inversions)
// (routing these through the two-product body like q2_sqr was measured and is slower: k_duo_miller_loop 23.8 -> 24.3 ms)
HDN void qv_mul(qfp *r, const qfp *a, const qfp *s, int n) {
    qfp k = *s;
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        QFOR { fp x = a[i].v[l_]; x = fp_mul_v(x, k.v[l_]); r[i].v[l_] = x; }
    }
}

// ---- Fq2 operations on the lane pair (fq2.go) ------------------------------------------------------------------
// r = a * b   (fq2.go:116-130: same value; one dot product with one reduction per lane)
HDN void q2_mul(qfp *r, const qfp *a, const qfp *b) {
    qfp A = *a, B = *b, AO, BO, R;
    q_shfl(AO, A, 1); q_shfl(BO, B, 1);
    QFOR {
        fp nbo, x, y;
        fp_qminus(nbo, BO.v[l_]);
#pragma unroll
        for (int k = 0; k < 12; k++) {
            x.l[k] = QJ ? BO.v[l_].l[k] : B.v[l_].l[k];       // j = 0: a0 b0 + a1 (Q - b1)    j = 1: a1 b0 + a0 b1
            y.l[k] = QJ ? B.v[l_].l[k] : nbo.l[k];
        }
#if defined(__CUDA_ARCH__)
        fp_dot2_inl(R.v[l_], A.v[l_], x, AO.v[l_], y);
#else
        R.v[l_] = fp_dot2_v(A.v[l_], x, AO.v[l_], y);
#endif
    }
    *r = R;
}
#if !defined(__CUDA_ARCH__)
inline bool &host_sqr_dot2() { static bool on = true; return on; }     // host emulation: which squaring form q2_sqr runs (tests flip it)
#endif
// r = a^2   (fq2.go:75-89: (a0 + a1)(a0 - a1) on lane 0, a1 (2 a0) on lane 1; the operand sums stay unreduced, below 2Q)
HDN void q2_sqr(qfp *r, const qfp *a) {
#ifndef B381_LANE_SQR_FPMUL
    // On a loaded machine: a0 a0 + a1 (Q - a1) | a1 a0 + a0 a1 through the two-product body -- 144 wide MACs more per squaring than
    // the form below, but the plain multiplier (6 KB) leaves the hot set of the lane kernels, which sit at the instruction-cache
    // limit: at 2^16 pairings k_duo_final_exp 27.1 -> 25.8 ms, four lanes 33.5 / 39.8 -> 31.9 / 36.0 ms.  A batch that leaves
    // the machine mostly empty is latency-bound and fetches do not contend, so it keeps the shorter form (64 pairings on four
    // lanes: 7.29 ms against 7.49; profiles/r02_experiments.md).  The host emulation runs the form the test selects.
#if defined(__CUDA_ARCH__)
    if (gridDim.x >= LANE_SQR_DOT2_MIN_BLOCKS) { q2_mul(r, a, a); return; }
#else
    if (host_sqr_dot2()) { q2_mul(r, a, a); return; }
#endif
#endif
    qfp A = *a, AO, R;
    q_shfl(AO, A, 1);
    QFOR {
        fp s, d, t, n, x, y;
        fp_add_nr(s, A.v[l_], AO.v[l_]);
        fp_qminus(n, AO.v[l_]);
        fp_add_nr(d, A.v[l_], n);
        fp_add_nr(t, AO.v[l_], AO.v[l_]);
#pragma unroll
        for (int k = 0; k < 12; k++) {
            x.l[k] = QJ ? A.v[l_].l[k] : s.l[k];
            y.l[k] = QJ ? t.l[k] : d.l[k];
        }
        R.v[l_] = fp_mul_v(x, y);
    }
    *r = R;
}
// r = a * (1 + u) = (a0 - a1) + (a0 + a1) u   (fq2.go:41-45)
HDN void q2_mul_nr(qfp *r, const qfp *a) {
    qfp A = *a, AO, R;
    q_shfl(AO, A, 1);
    QFOR {
        fp x = A.v[l_], y = AO.v[l_];
        if (QJ) fp_add(x, x, y); else fp_sub(x, x, y);
        R.v[l_] = x;
    }
    *r = R;
}
// r = a^-1, 0 -> 0   (fq2.go:133-147); both lanes invert the norm
HDN void q2_inv(qfp *r, const qfp *a) {
    qfp A = *a, S, SO, R;
    QFOR S.v[l_] = fp_mul_v(A.v[l_], A.v[l_]);
    q_shfl(SO, S, 1);
    QFOR {
        fp n, t;
        fp_add(n, S.v[l_], SO.v[l_]);
        fp_inv(&n, &n);
#if defined(__CUDA_ARCH__)
        __syncwarp();                              // the trip count of fp_inv depends on the value: reconverge before the next shuffle
#endif
        t = fp_mul_v(A.v[l_], n);
        if (QJ) fp_neg(t, t);
        R.v[l_] = t;
    }
    *r = R;
}
HD void q_set_zero(qfp &r) { QFOR fp_set_zero(r.v[l_]); }
// the Fq2 value 1: lane j = 0 holds R mod Q, lane j = 1 holds 0
HD void q2_set_one(qfp &r) {
    QFOR { if (QJ) fp_set_zero(r.v[l_]); else fp_set_one(r.v[l_]); }
}
// an Fq2 constant from a 24-word table entry (c0 || c1)
HD void q2_load_tab(qfp &r, const uint32_t *t) { QFOR fp_load_tab(r.v[l_], t + 12 * QJ); }
// per lane: is the Fq2 value (own coefficient AND the partner's) zero
HD void q2_is_zero(bool z[QL], const qfp &a) {
    qfp o;
    q_shfl(o, a, 1);
    QFOR z[l_] = fp_is_zero(a.v[l_]) && fp_is_zero(o.v[l_]);
}

// ---- Fq6 on a half (fq6.go), both halves running the same code on their own operands ---------------------------------
// r = a * b   (fq6.go:255-292; six Fq2 products)
HDN void q6_mul(q6 *r, const q6 *a, const q6 *b) {
    qfp v0, v1, v2, s, t, x, y, z;
    q2_mul(&v0, qa(a, 0), qa(b, 0));
    q2_mul(&v1, qa(a, 1), qa(b, 1));
    q2_mul(&v2, qa(a, 2), qa(b, 2));
    // c0 = v0 + xi ((a1 + a2)(b1 + b2) - v1 - v2)
    qv_add(&s, qa(a, 1), qa(a, 2), 1);
    qv_add(&t, qa(b, 1), qa(b, 2), 1);
    q2_mul(&x, &s, &t);
    qv_sub(&x, &x, &v1, 1);
    qv_sub(&x, &x, &v2, 1);
    q2_mul_nr(&x, &x);
    qv_add(&x, &x, &v0, 1);
    // c1 = (a0 + a1)(b0 + b1) - v0 - v1 + xi v2
    qv_add(&s, qa(a, 0), qa(a, 1), 1);
    qv_add(&t, qa(b, 0), qa(b, 1), 1);
    q2_mul(&y, &s, &t);
    qv_sub(&y, &y, &v0, 1);
    qv_sub(&y, &y, &v1, 1);
    q2_mul_nr(&s, &v2);
    qv_add(&y, &y, &s, 1);
    // c2 = (a0 + a2)(b0 + b2) - v0 - v2 + v1
    qv_add(&s, qa(a, 0), qa(a, 2), 1);
    qv_add(&t, qa(b, 0), qa(b, 2), 1);
    q2_mul(&z, &s, &t);
    qv_sub(&z, &z, &v0, 1);
    qv_sub(&z, &z, &v2, 1);
    qv_add(&z, &z, &v1, 1);
    r->c0 = x; r->c1 = y; r->c2 = z;
}
// r = a * (b0 + b1 v)   (fq6.go:60-90; five Fq2 products)
HDN void q6_mul_by_01(q6 *r, const q6 *a, const qfp *b0, const qfp *b1) {
    qfp v0, v1, s, t, x, y, z;
    q2_mul(&v0, qa(a, 0), b0);
    q2_mul(&v1, qa(a, 1), b1);
    qv_add(&s, qa(a, 1), qa(a, 2), 1);
    q2_mul(&x, &s, b1);
    qv_sub(&x, &x, &v1, 1);
    q2_mul_nr(&x, &x);
    qv_add(&x, &x, &v0, 1);
    qv_add(&s, qa(a, 0), qa(a, 1), 1);
    qv_add(&t, b0, b1, 1);
    q2_mul(&y, &s, &t);
    qv_sub(&y, &y, &v0, 1);
    qv_sub(&y, &y, &v1, 1);
    qv_add(&s, qa(a, 0), qa(a, 2), 1);
    q2_mul(&z, &s, b0);
    qv_sub(&z, &z, &v0, 1);
    qv_add(&z, &z, &v1, 1);
    r->c0 = x; r->c1 = y; r->c2 = z;
}
// r = a * v = (xi a2, a0, a1)   (fq6.go:34-37)
HD void q6_mul_v(q6 *r, const q6 *a) {
    qfp t, c0 = a->c0, c1 = a->c1;
    q2_mul_nr(&t, qa(a, 2));
    r->c0 = t; r->c1 = c0; r->c2 = c1;
}
// r = a^-1, 0 -> 0   (fq6.go:295-336); run by both halves on the same value (one per final exponentiation)
HDN void q6_inv(q6 *r, const q6 *a) {
    qfp k0, k1, k2, t, u;
    q2_sqr(&k0, qa(a, 0));
    q2_mul(&t, qa(a, 1), qa(a, 2));
    q2_mul_nr(&t, &t);
    qv_sub(&k0, &k0, &t, 1);
    q2_sqr(&k1, qa(a, 2));
    q2_mul_nr(&k1, &k1);
    q2_mul(&t, qa(a, 0), qa(a, 1));
    qv_sub(&k1, &k1, &t, 1);
    q2_sqr(&k2, qa(a, 1));
    q2_mul(&t, qa(a, 0), qa(a, 2));
    qv_sub(&k2, &k2, &t, 1);
    q2_mul(&t, qa(a, 2), &k1);
    q2_mul(&u, qa(a, 1), &k2);
    qv_add(&t, &t, &u, 1);
    q2_mul_nr(&t, &t);
    q2_mul(&u, qa(a, 0), &k0);
    qv_add(&t, &t, &u, 1);
    q2_inv(&t, &t);
    q2_mul(qa(r, 0), &k0, &t);
    q2_mul(qa(r, 1), &k1, &t);
    q2_mul(qa(r, 2), &k2, &t);
}

// ---- Fq12 split over the halves (fq12.go): F = the own half's Fq6 coefficient ----------------------------------------
// f <- f^2   (fq12.go:180-195, complex squaring): half 0 forms ab = c0 c1, half 1 forms (c0 + c1)(c0 + v c1)
HDN void q12_sqr(q6 *F) {
    q6 O, S, T, Pr;
    qfp x;
    qv_xh(qa(&O, 0), qa(F, 0), 3);
    qv_add(qa(&S, 0), qa(F, 0), qa(&O, 0), 3);
    // on half 1 (F = c1, O = c0): c0 + v c1 = (O0 + xi F2, O1 + F0, O2 + F1)
    q2_mul_nr(&x, qa(F, 2));
    qv_add(qa(&T, 0), qa(&O, 0), &x, 1);
    qv_add(qa(&T, 1), qa(&O, 1), qa(F, 0), 2);
    qv_selh(qa(&S, 0), qa(F, 0), qa(&S, 0), 3);              // X = c0      | c0 + c1
    qv_selh(qa(&T, 0), qa(&O, 0), qa(&T, 0), 3);               // Y = c1      | c0 + v c1
    q6_mul(&Pr, &S, &T);
    qv_xh(qa(&O, 0), qa(&Pr, 0), 3);
    qv_selh(qa(&S, 0), qa(&Pr, 0), qa(&O, 0), 3);              // ab on all lanes
    qv_selh(qa(&T, 0), qa(&O, 0), qa(&Pr, 0), 3);              // (c0 + c1)(c0 + v c1) on all lanes
    // c0' = T - ab - v ab   (half 0)        c1' = 2 ab   (half 1)
    qv_sub(qa(&T, 0), qa(&T, 0), qa(&S, 0), 3);
    q2_mul_nr(&x, qa(&S, 2));
    qv_sub(qa(&T, 0), qa(&T, 0), &x, 1);
    qv_sub(qa(&T, 1), qa(&T, 1), qa(&S, 0), 2);
    qv_dbl(qa(&S, 0), qa(&S, 0), 3);
    qv_selh(qa(F, 0), qa(&T, 0), qa(&S, 0), 3);
}
// f <- f * ((d0 + d1 v) + (d4 v) w)   (fq12.go:32-47; 13 Fq2 products) and, in the free slot of the seventh round,
// *eout = ea * eb (the y coordinate of the line step that produced d: its last product has no partner there)
HDN void q12_mul_by_014(q6 *F, const qfp *d0, const qfp *d1, const qfp *d4, const qfp *ea, const qfp *eb, qfp *eout) {
    q6 O, X, Pr, U;
    qfp D1, ia, ib, own, b0, b1, b2;
    qv_xh(qa(&O, 0), qa(F, 0), 3);
    qv_add(qa(&X, 0), qa(F, 0), qa(&O, 0), 3);
    qv_selh(qa(&X, 0), qa(F, 0), qa(&X, 0), 3);              // c0 | c0 + c1
    qv_add(&D1, d1, d4, 1);
    qv_selh(&D1, d1, &D1, 1);                      // d1 | d1 + d4
    q6_mul_by_01(&Pr, &X, d0, &D1);                // aa = c0 (d0, d1) | (c0 + c1)(d0, d1 + d4)
    // bb = c1 * (d4 v) = (xi c1_2 d4, c1_0 d4, c1_1 d4): three products + the extra one
    qv_selh(qa(&X, 0), qa(&O, 0), qa(F, 0), 3);              // c1 on all lanes
    qv_selh(&ia, qa(&X, 0), qa(&X, 1), 1);
    q2_mul(&own, &ia, d4);
    q_bcast(&b0, &b1, &own);
    qv_selh(&ia, qa(&X, 2), ea, 1);
    qv_selh(&ib, d4, eb, 1);
    q2_mul(&own, &ia, &ib);
    q_bcast(&b2, eout, &own);
    q2_mul_nr(&b2, &b2);                           // bb = (b2, b0, b1) from here on
    // c0' = aa + v bb = (aa0 + xi b1, aa1 + b2, aa2 + b0)   (half 0)
    // c1' = cc - aa - bb                                      (half 1)
    q2_mul_nr(&own, &b1);
    qv_xh(qa(&O, 0), qa(&Pr, 0), 3);                       // half 1 receives aa
    U.c0 = own; U.c1 = b2; U.c2 = b0;
    X.c0 = b2; X.c1 = b0; X.c2 = b1;
    qv_selh(qa(&U, 0), qa(&U, 0), qa(&X, 0), 3);
    qv_sub(qa(&O, 0), qa(&Pr, 0), qa(&O, 0), 3);               // cc - aa (meaningful on half 1)
    qv_selh(qa(&Pr, 0), qa(&Pr, 0), qa(&O, 0), 3);
    qv_addsub(qa(F, 0), qa(&Pr, 0), qa(&U, 0), 3, 2);
}
// r = a * b   (fq12.go:198-213: aa = a0 b0 on half 0, bb = a1 b1 on half 1, and the six products of
// (a0 + a1)(b0 + b1) split three and three: 9 rounds for 18 Fq2 products).  r may alias a or b.
HDN void q12_mul(q6 *r, const q6 *a, const q6 *b) {
    q6 Pr, SA, SB, O;
    qfp ia, ib, own, v0, v1, v2, m12, m01, m02, s;
    q6_mul(&Pr, a, b);
    qv_xh(qa(&O, 0), qa(a, 0), 3);
    qv_add(qa(&SA, 0), qa(a, 0), qa(&O, 0), 3);
    qv_xh(qa(&O, 0), qa(b, 0), 3);
    qv_add(qa(&SB, 0), qa(b, 0), qa(&O, 0), 3);
    // round 7: v0 = SA0 SB0 | (SA1 + SA2)(SB1 + SB2)
    qv_add(&ia, qa(&SA, 1), qa(&SA, 2), 1); qv_add(&ib, qa(&SB, 1), qa(&SB, 2), 1);
    qv_selh(&ia, qa(&SA, 0), &ia, 1); qv_selh(&ib, qa(&SB, 0), &ib, 1);
    q2_mul(&own, &ia, &ib);
    q_bcast(&v0, &m12, &own);
    // round 8: v1 = SA1 SB1 | (SA0 + SA1)(SB0 + SB1)
    qv_add(&ia, qa(&SA, 0), qa(&SA, 1), 1); qv_add(&ib, qa(&SB, 0), qa(&SB, 1), 1);
    qv_selh(&ia, qa(&SA, 1), &ia, 1); qv_selh(&ib, qa(&SB, 1), &ib, 1);
    q2_mul(&own, &ia, &ib);
    q_bcast(&v1, &m01, &own);
    // round 9: v2 = SA2 SB2 | (SA0 + SA2)(SB0 + SB2)
    qv_add(&ia, qa(&SA, 0), qa(&SA, 2), 1); qv_add(&ib, qa(&SB, 0), qa(&SB, 2), 1);
    qv_selh(&ia, qa(&SA, 2), &ia, 1); qv_selh(&ib, qa(&SB, 2), &ib, 1);
    q2_mul(&own, &ia, &ib);
    q_bcast(&v2, &m02, &own);
    // cc = (v0 + xi (m12 - v1 - v2), m01 - v0 - v1 + xi v2, m02 - v0 - v2 + v1)   (fq6.go:255-292)
    qv_sub(&m12, &m12, &v1, 1); qv_sub(&m12, &m12, &v2, 1);
    q2_mul_nr(&m12, &m12);
    qv_add(qa(&SA, 0), &m12, &v0, 1);
    qv_sub(&m01, &m01, &v0, 1); qv_sub(&m01, &m01, &v1, 1);
    q2_mul_nr(&s, &v2);
    qv_add(qa(&SA, 1), &m01, &s, 1);
    qv_sub(&m02, &m02, &v0, 1); qv_sub(&m02, &m02, &v2, 1);
    qv_add(qa(&SA, 2), &m02, &v1, 1);
    // c0 = aa + v bb   (half 0: Pr = aa, O = bb)        c1 = cc - aa - bb   (half 1: Pr = bb, O = aa)
    qv_xh(qa(&O, 0), qa(&Pr, 0), 3);
    q6_mul_v(&SB, &O);
    qv_add(qa(&SB, 0), qa(&Pr, 0), qa(&SB, 0), 3);
    qv_sub(qa(&SA, 0), qa(&SA, 0), qa(&Pr, 0), 3);
    qv_sub(qa(&SA, 0), qa(&SA, 0), qa(&O, 0), 3);
    qv_selh(qa(r, 0), qa(&SB, 0), qa(&SA, 0), 3);
}
HD void q12_conj(q6 *r, const q6 *a) { qv_neg(qa(r, 0), qa(a, 0), 3, 3); }     // fq12.go:27-29: c1 <- -c1
HD void q12_set_one(q6 *r) {
    qfp one, z;
    q2_set_one(one); q_set_zero(z);
    QFOR { if (QH) one.v[l_] = z.v[l_]; }
    r->c0 = one; r->c1 = z; r->c2 = z;
}
// per lane: is the own quad's Fq12 value equal to 1 / to 0
HD void q12_is_one(bool r[QL], const q6 *a) {
    q6 one;
    q12_set_one(&one);
    bool p[QL];
    QFOR p[l_] = fp_eq(a->c0.v[l_], one.c0.v[l_]) && fp_is_zero(a->c1.v[l_]) && fp_is_zero(a->c2.v[l_]);
    q_all4(r, p);
}
HD void q12_is_zero(bool r[QL], const q6 *a) {
    bool p[QL];
    QFOR p[l_] = fp_is_zero(a->c0.v[l_]) && fp_is_zero(a->c1.v[l_]) && fp_is_zero(a->c2.v[l_]);
    q_all4(r, p);
}
// r = a^(q^power), power in {1, 2, 3}   (fq12.go:171-177, fq6.go:211-218).  Both halves multiply their c1, c2 by the Fq6
// table entries; half 1 then multiplies its three coefficients by the Fq12 entry (half 0 multiplies by one).
HDN void q12_frobenius(q6 *r, const q6 *a, int power) {
    q6 t = *a;
    qfp k, one;
    if (power & 1) qv_neg(qa(&t, 0), qa(&t, 0), 3, 1);                               // fq2.go:156-158
    q2_load_tab(k, B381_TAB(frob6_c1) + power * 24);
    q2_mul(qa(&t, 1), qa(&t, 1), &k);
    q2_load_tab(k, B381_TAB(frob6_c2) + power * 24);
    q2_mul(qa(&t, 2), qa(&t, 2), &k);
    q2_load_tab(k, B381_TAB(frob12_c1) + power * 24);
    q2_set_one(one);
    qv_selh(&k, &one, &k, 1);
    q2_mul(qa(&t, 0), qa(&t, 0), &k);
    q2_mul(qa(&t, 1), qa(&t, 1), &k);
    q2_mul(qa(&t, 2), qa(&t, 2), &k);
    *r = t;
}
// r = a^-1; ok = false (r untouched) for a == 0   (fq12.go:216-237)
HDN void q12_inv(q6 *r, const q6 *a, bool ok[QL]) {
    q6 Pr, O, t0;
    q6_mul(&Pr, a, a);                             // c0^2 | c1^2
    qv_xh(qa(&O, 0), qa(&Pr, 0), 3);
    qv_selh(qa(&t0, 0), qa(&Pr, 0), qa(&O, 0), 3);             // c0^2 on all lanes
    qv_selh(qa(&O, 0), qa(&O, 0), qa(&Pr, 0), 3);              // c1^2 on all lanes
    q6_mul_v(&O, &O);
    qv_sub(qa(&t0, 0), qa(&t0, 0), qa(&O, 0), 3);
    bool z[QL];
    q12_is_zero(z, &t0);                           // (both halves hold the same Fq6 value)
    QFOR ok[l_] = !z[l_];
    q6_inv(&t0, &t0);
    q6_mul(&Pr, a, &t0);                           // c0 t | c1 t
    qv_neg(qa(r, 0), qa(&Pr, 0), 3, 3);                  // (c0 t, -c1 t)
}

// Squaring in the cyclotomic subgroup (Granger-Scott; fp12_cyclotomic_sqr of tower.cuh): nine Fq2 squarings in five rounds.
// z0 = c0.c0, z4 = c0.c1, z3 = c0.c2 (half 0), z2 = c1.c0, z1 = c1.c1, z5 = c1.c2 (half 1).
HDN void q12_cyc_sqr(q6 *F) {
    q6 O, A, B;        // A = half 0's coefficients (z0, z4, z3), B = half 1's (z2, z1, z5), on all lanes
    qfp in, own, s, a2[3], b2[3], ab2[3], t;
    qv_xh(qa(&O, 0), qa(F, 0), 3);
    qv_selh(qa(&A, 0), qa(F, 0), qa(&O, 0), 3);
    qv_selh(qa(&B, 0), qa(&O, 0), qa(F, 0), 3);
    // Fp4 pairs (a, b): (z0, z1) = (A0, B1), (z2, z3) = (B0, A2), (z4, z5) = (A1, B2)
    qv_selh(&in, qa(&A, 0), qa(&B, 1), 1); q2_sqr(&own, &in); q_bcast(&a2[0], &b2[0], &own);
    qv_add(&s, qa(&A, 0), qa(&B, 1), 1);
    qv_selh(&in, &s, qa(&B, 0), 1); q2_sqr(&own, &in); q_bcast(&ab2[0], &a2[1], &own);
    qv_add(&s, qa(&B, 0), qa(&A, 2), 1);
    qv_selh(&in, qa(&A, 2), &s, 1); q2_sqr(&own, &in); q_bcast(&b2[1], &ab2[1], &own);
    qv_selh(&in, qa(&A, 1), qa(&B, 2), 1); q2_sqr(&own, &in); q_bcast(&a2[2], &b2[2], &own);
    qv_add(&s, qa(&A, 1), qa(&B, 2), 1);
    q2_sqr(&ab2[2], &s);                           // (both halves: the ninth squaring has no partner)
    // fp4_sqr: t0 = a^2 + xi b^2, t1 = (a + b)^2 - a^2 - b^2
    //   z0' = 3 t0(01) - 2 z0    z4' = 3 t0(23) - 2 z4    z3' = 3 t0(45) - 2 z3             (half 0: 3 T - 2 z)
    //   z2' = 3 xi t1(45) + 2 z2    z1' = 3 t1(01) + 2 z1    z5' = 3 t1(23) + 2 z5           (half 1: 3 T + 2 z)
    q6 T0, T1;
    qfp *t0p[3] = {qa(&T0, 0), qa(&T0, 1), qa(&T0, 2)}, *t1p[3] = {qa(&T1, 0), qa(&T1, 1), qa(&T1, 2)};
#pragma unroll 1
    for (int i = 0; i < 3; i++) {
        q2_mul_nr(&t, &b2[i]);
        qv_add(t0p[i], &a2[i], &t, 1);
        qv_sub(&t, &ab2[i], &a2[i], 1);
        qv_sub(t1p[i], &t, &b2[i], 1);
    }
    q2_mul_nr(&t, qa(&T1, 2));
    // half 0 rows (z0, z4, z3) take (t0(01), t0(23), t0(45)); half 1 rows (z2, z1, z5) take (xi t1(45), t1(01), t1(23))
    O.c0 = t; O.c1 = T1.c0; O.c2 = T1.c1;
    qv_selh(qa(&T0, 0), qa(&T0, 0), qa(&O, 0), 3);
    // 3 T -+ 2 z = T + 2 (T -+ z)
    qv_addsub(qa(&O, 0), qa(&T0, 0), qa(F, 0), 3, 3);
    qv_dbl(qa(&O, 0), qa(&O, 0), 3);
    qv_add(qa(F, 0), qa(&O, 0), qa(&T0, 0), 3);
}

// ---- global memory <-> lanes ---------------------------------------------------------------------------------------
// lane (h, j) reads / writes the Fq at base[off + j sj + h sh] (offsets in Fq = 6 x u64)
HD void q_load(qfp &r, const uint64_t *base, int off, int sj, int sh) {
    QFOR fp_load_u64(r.v[l_], base + 6 * (size_t)(off + QJ * sj + QH * sh));
}
HD void q_store(uint64_t *base, const qfp &a, int off, int sj, int sh) {
    QFOR fp_store_u64(base + 6 * (size_t)(off + QJ * sj + QH * sh), a.v[l_]);
}
// the own half of an Fq12 in the reference's flattened order (c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2; each c0 || c1)
HD void q12_load(q6 *F, const uint64_t *src) {
    q_load(F->c0, src, 0, 1, 6); q_load(F->c1, src, 2, 1, 6); q_load(F->c2, src, 4, 1, 6);
}
HD void q12_store(uint64_t *dst, const q6 *F) {
    q_store(dst, F->c0, 0, 1, 6); q_store(dst, F->c1, 2, 1, 6); q_store(dst, F->c2, 4, 1, 6);
}

// ---- Miller loop ---------------------------------------------------------------------------------------------------
struct qpair {
    qfp rx, ry, rz;          // the running point R (Jacobian, own coefficient j; both halves hold it)
    qfp px, py;              // P (Fq, the same on all four lanes)
    qfp qx, qy, ysq;         // Q and qy^2 (own coefficient j)
};
HD void qpair_load(qpair *S, const g1_affine_pod *P, const g2_affine_pod *Q) {
    q_load(S->px, P->x, 0, 0, 0); q_load(S->py, P->y, 0, 0, 0);
    q_load(S->qx, Q->x, 0, 1, 0); q_load(S->qy, Q->y, 0, 1, 0);
    S->rx = S->qx; S->ry = S->qy; q2_set_one(S->rz);
    q2_sqr(&S->ysq, &S->qy);
}
// scale the raw line coefficients of a step by P (the `ell` closure, pairing.go:28-39) -- own = m1 | m2 with
// o1 = -2 m1 -> d1 = o1 px (half 0), o0 = 2 m2 -> d4 = o0 py (half 1) -- and hand both to all lanes
HD void qml_scale(qfp *d1, qfp *d4, qfp *own, const qpair *S) {
    qfp k;
    qv_dbl(own, own, 1);
    qv_neg(own, own, 1, 2);
    qv_selh(&k, &S->px, &S->py, 1);
    qv_mul(own, own, &k, 1);
    q_bcast(d1, d4, own);
}
// doubling step (g2.go:655-708) + f <- f * line(P): 4 squaring rounds, 1 product round, 1 scaling round, then the 7 rounds of
// the sparse multiplication, whose free slot takes the step's last product (the new y)
HDN void qml_double(q6 *F, qpair *S) {
    qfp in, own, t0, t1, zsq, zy, t2, t3, t4, t5, t6, tmp, d1, d4, ya, yp;
    qv_selh(&in, &S->rx, &S->ry, 1);
    q2_sqr(&own, &in); q_bcast(&t0, &t1, &own);                    // x^2 | y^2
    qv_add(&tmp, &S->rz, &S->ry, 1);
    qv_selh(&in, &S->rz, &tmp, 1);
    q2_sqr(&own, &in); q_bcast(&zsq, &zy, &own);                   // z^2 | (z + y)^2
    qv_sub(&zy, &zy, &t1, 1);
    qv_sub(&S->rz, &zy, &zsq, 1);                                  // z' = (z + y)^2 - y^2 - z^2
    qv_add(&tmp, &t1, &S->rx, 1);
    qv_selh(&in, &t1, &tmp, 1);
    q2_sqr(&own, &in); q_bcast(&t2, &t3, &own);                    // y^4 | (y^2 + x)^2
    qv_dbl(&t4, &t0, 1); qv_add(&t4, &t4, &t0, 1);                 // 3 x^2
    qv_add(&tmp, &S->rx, &t4, 1);
    qv_selh(&in, &t4, &tmp, 1);
    q2_sqr(&own, &in); q_bcast(&t5, &t6, &own);                    // t4^2 | (x + t4)^2
    qv_selh(&in, &t4, &S->rz, 1);
    q2_mul(&own, &in, &zsq);                                       // t4 z^2 | z' z^2
    qml_scale(&d1, &d4, &own, S);
    qv_sub(&t3, &t3, &t0, 1); qv_sub(&t3, &t3, &t2, 1); qv_dbl(&t3, &t3, 1);
    qv_sub(&tmp, &t5, &t3, 1); qv_sub(&S->rx, &tmp, &t3, 1);       // x' = t4^2 - 2 t3
    qv_sub(&t6, &t6, &t0, 1); qv_sub(&t6, &t6, &t5, 1);
    qv_dbl(&t1, &t1, 1); qv_dbl(&t1, &t1, 1);
    qv_sub(&t6, &t6, &t1, 1);                                      // d0 = (x + t4)^2 - x^2 - t4^2 - 4 y^2
    qv_sub(&ya, &t3, &S->rx, 1);
    q12_mul_by_014(F, &t6, &d1, &d4, &ya, &t4, &yp);
    qv_dbl(&t2, &t2, 1); qv_dbl(&t2, &t2, 1); qv_dbl(&t2, &t2, 1);
    qv_sub(&S->ry, &yp, &t2, 1);                                   // y' = (t3 - x') t4 - 8 y^4
}
// addition step (g2.go:710-772) + f <- f * line(P)
HDN void qml_add(q6 *F, qpair *S) {
    qfp in, ib, own, zsq, t0, t1, t2, t3, t4, t5, t6, t7, t9, t10, zn, tmp, d1, d4, d0, ya, yp, z2n, x6;
    qv_add(&tmp, &S->qy, &S->rz, 1);
    qv_selh(&in, &S->rz, &tmp, 1);
    q2_sqr(&own, &in); q_bcast(&zsq, &t1, &own);                   // z^2 | (qy + z)^2
    qv_sub(&t1, &t1, &S->ysq, 1); qv_sub(&t1, &t1, &zsq, 1);
    qv_selh(&in, &S->qx, &t1, 1);
    q2_mul(&own, &in, &zsq); q_bcast(&t0, &t1, &own);              // qx z^2 | t1 z^2
    qv_sub(&t2, &t0, &S->rx, 1);
    qv_add(&tmp, &S->rz, &t2, 1);
    qv_selh(&in, &t2, &tmp, 1);
    q2_sqr(&own, &in); q_bcast(&t3, &zn, &own);                    // t2^2 | (z + t2)^2
    qv_sub(&zn, &zn, &zsq, 1); qv_sub(&zn, &zn, &t3, 1);           // z'
    qv_sub(&t6, &t1, &S->ry, 1); qv_sub(&t6, &t6, &S->ry, 1);
    qv_add(&t10, &S->qy, &zn, 1);
    qv_selh(&in, &t6, &t10, 1);
    q2_sqr(&own, &in); q_bcast(&x6, &t10, &own);                   // t6^2 | (qy + z')^2
    q2_sqr(&z2n, &zn);                                             // z'^2 (both halves)
    qv_dbl(&t4, &t3, 1); qv_dbl(&t4, &t4, 1);
    qv_selh(&ib, &t2, &S->rx, 1);
    q2_mul(&own, &t4, &ib); q_bcast(&t5, &t7, &own);               // t4 t2 | t4 x
    qv_sub(&x6, &x6, &t5, 1); qv_sub(&x6, &x6, &t7, 1); qv_sub(&x6, &x6, &t7, 1);   // x'
    qv_selh(&in, &t6, &S->ry, 1);
    qv_selh(&ib, &S->qx, &t5, 1);
    q2_mul(&own, &in, &ib); q_bcast(&t9, &t0, &own);               // t6 qx | y t5
    qv_sub(&t10, &t10, &S->ysq, 1); qv_sub(&t10, &t10, &z2n, 1);
    qv_dbl(&t9, &t9, 1);
    qv_sub(&d0, &t9, &t10, 1);                                     // coefficient 2
    // raw coefficients 0 and 1: 2 z' and -2 t6 -> qml_scale wants own = (m1 | m2) with o1 = -2 m1, o0 = 2 m2
    qv_selh(&own, &t6, &zn, 1);
    qml_scale(&d1, &d4, &own, S);
    qv_sub(&ya, &t7, &x6, 1);
    q12_mul_by_014(F, &d0, &d1, &d4, &ya, &t6, &yp);
    qv_dbl(&t0, &t0, 1);
    qv_sub(&S->ry, &yp, &t0, 1);                                   // y' = (t7 - x') t6 - 2 y t5
    S->rx = x6; S->rz = zn;
}

// Miller loop of NP pairs sharing the accumulator (pairing.go:16-75), conjugated; a pair with P or Q at infinity
// contributes the factor 1 (pairing.cuh): live[k] per pair, the same on the four lanes of a quad.
// Control flow is uniform over the warp: dead pairs compute on whatever their coordinates are and are masked out.
struct qlive { bool on[QL]; };
HD void q6_keep_if(q6 *F, const q6 *G, const qlive &keep) {      // F <- keep ? F : G, lane by lane
    QFOR { if (!keep.on[l_]) { F->c0.v[l_] = G->c0.v[l_]; F->c1.v[l_] = G->c1.v[l_]; F->c2.v[l_] = G->c2.v[l_]; } }
}
template <int NP>
HD void q_miller_loop(q6 *F, qpair *S, const qlive *live) {
    q12_set_one(F);
    const uint64_t xr = 0xd201000000010000ULL >> 1;
#pragma unroll 1
    for (int bit = 61; bit >= -1; bit--) {
#pragma unroll 1
        for (int k = 0; k < NP; k++) {
            q6 G;
            if (NP > 1) G = *F;
            qml_double(F, &S[k]);
            if (bit >= 0 && ((xr >> bit) & 1)) qml_add(F, &S[k]);
            if (NP > 1) q6_keep_if(F, &G, live[k]);            // a dead pair leaves the accumulator as it was
        }
        if (bit >= 0) q12_sqr(F);
    }
    q12_conj(F, F);
    if (NP == 1) { q6 one; q12_set_one(&one); q6_keep_if(F, &one, live[0]); }
}
HD void qpair_live(qlive &lv, const g1_affine_pod *P, const g2_affine_pod *Q) {
    QFOR lv.on[l_] = !(P->inf || Q->inf);
}

// ---- final exponentiation ------------------------------------------------------------------------------------------
// Compressed cyclotomic squaring (Karabina; cyc_sqr_compressed of pairing.cuh) with half 0 holding (g2, g3) and half 1
// holding (g4, g5) as G[0], G[1]:  A = (ga + gb)(ga + xi gb), B = ga gb on the own pair, then
//   half 0 needs  h2 = 3 (2 xi B45) + 2 g2,  h3 = 3 (A45 - (xi + 1) B45) - 2 g3     from half 1's pair
//   half 1 needs  h4 = 3 (A23 - (xi + 1) B23) - 2 g4,  h5 = 3 (2 B23) + 2 g5        from half 0's pair
// Two product rounds per squaring, both halves busy.
HDN void q_cyc_sqr_compressed(qfp *G) {
    qfp t0, t1, A, B, xB, U[2], E[2];
    q2_mul_nr(&t0, &G[1]); qv_add(&t0, &t0, &G[0], 1);
    qv_add(&t1, &G[0], &G[1], 1);
    q2_mul(&A, &t0, &t1);
    q2_mul(&B, &G[0], &G[1]);
    q2_mul_nr(&xB, &B);
    qv_sub(&A, &A, &xB, 1); qv_sub(&A, &A, &B, 1);
    // what the other half needs: (tripled-and-added term of its G[0], of its G[1])
    //   half 0 sends (A23', 2 B23) for (g4, g5);   half 1 sends (2 xi B45, A45') for (g2, g3)
    qv_selh(&B, &B, &xB, 1);
    qv_dbl(&B, &B, 1);
    qv_selh(&U[0], &A, &B, 1);
    qv_selh(&U[1], &B, &A, 1);
    qv_xh(E, U, 2);
    // half 0: G0 = 3 E0 + 2 G0, G1 = 3 E1 - 2 G1;   half 1: G0 = 3 E0 - 2 G0, G1 = 3 E1 + 2 G1   (3 E +- 2 G = E + 2 (E +- G))
    qv_addsub(&t0, &E[0], &G[0], 1, 2);
    qv_addsub(&t1, &E[1], &G[1], 1, 3);
    qv_dbl(&t0, &t0, 1); qv_dbl(&t1, &t1, 1);
    qv_add(&G[0], &t0, &E[0], 1);
    qv_add(&G[1], &t1, &E[1], 1);
}
HD void q_cyc_compress(qfp *G, const q6 *F) {
    // g2 = c1.c0 (half 1, F.c0), g3 = c0.c2 (half 0, F.c2), g4 = c0.c1 (half 0, F.c1), g5 = c1.c2 (half 1, F.c2)
    // half 0 holds (g2, g3) = (the other half's c0, own c2);  half 1 holds (g4, g5) = (the other half's c1, own c2)
    q6 O;
    qv_xh(qa(&O, 0), qa(F, 0), 3);
    qv_selh(&G[0], qa(&O, 0), qa(&O, 1), 1);
    G[1] = F->c2;
}
// all four compressed coefficients (g2, g3, g4, g5) of the own quad on every lane
HD void q_cyc_gather(qfp *g, const qfp *G) {
    qfp E[2];
    qv_xh(E, G, 2);
    qv_selh(&g[0], &G[0], &E[0], 1); qv_selh(&g[1], &G[1], &E[1], 1);
    qv_selh(&g[2], &E[0], &G[0], 1); qv_selh(&g[3], &E[1], &G[1], 1);
}
// decompression (cyc_decompress of pairing.cuh) of the value g = (g2, g3, g4, g5) held by the own half (the halves may hold
// different values): out = (g0, g1):  g1 = (xi g5^2 + 3 g4^2 - 2 g3) inv4g2,  g0 = (2 g1^2 + g2 g5 - 3 g3 g4) xi + 1
HDN void q_cyc_decompress(qfp *g0, qfp *g1, const qfp *g, const qfp *inv4g2) {
    qfp t0, t1, one;
    q2_sqr(&t0, &g[3]); q2_mul_nr(&t0, &t0);
    q2_sqr(&t1, &g[2]);
    qv_sub(&one, &t1, &g[1], 1); qv_dbl(&one, &one, 1); qv_add(&t1, &one, &t1, 1);     // 3 g4^2 - 2 g3
    qv_add(&t0, &t0, &t1, 1);
    q2_mul(g1, &t0, inv4g2);
    q2_sqr(&t0, g1); qv_dbl(&t0, &t0, 1);
    q2_mul(&t1, &g[0], &g[3]); qv_add(&t0, &t0, &t1, 1);
    q2_mul(&t1, &g[1], &g[2]);
    qv_sub(&t0, &t0, &t1, 1); qv_dbl(&t1, &t1, 1); qv_sub(&t0, &t0, &t1, 1);
    q2_mul_nr(&t0, &t0);
    q2_set_one(one);
    qv_add(g0, &t0, &one, 1);
}
// the Fq12 value (g0 .. g5), all on every lane of the quad, in the split layout: half 0 = (g0, g4, g3), half 1 = (g2, g1, g5)
HD void q_cyc_place(q6 *F, const qfp *g0, const qfp *g1, const qfp *g) {
    qfp a[3], b[3];
    a[0] = *g0; a[1] = g[2]; a[2] = g[1];
    b[0] = g[0]; b[1] = *g1; b[2] = g[3];
    qv_selh(qa(F, 0), a, b, 3);
}
// conj(f^x), square and multiply with Granger-Scott squarings (exp_by_x_gs of pairing.cuh); r must not alias f
HDN void q_exp_by_x_gs(q6 *r, const q6 *f, uint64_t x) {
    q6 acc = *f;
    int top = 63;
    while (!((x >> top) & 1)) top--;
#pragma unroll 1
    for (int bit = top - 1; bit >= 0; bit--) {
        q12_cyc_sqr(&acc);
        if ((x >> bit) & 1) q12_mul(&acc, &acc, f);
    }
    q12_conj(r, &acc);
}
// conj(f^x) for x = |x_BLS| and |x_BLS| / 2 (exp_by_x of pairing.cuh: e + 41 compressed squarings, three decompressions with
// one shared inversion, six Granger-Scott squarings).  A quad whose compressed form degenerates (some g2 = 0: f = 1, ...)
// makes its warp run the square-and-multiply form as well; every quad keeps the result of the form that is valid for it.
HDN void q_exp_by_x_main(q6 *r, bool bad[QL], const q6 *f, uint64_t x, int stage = 0) {
    int e = 0;
    while (!((x >> e) & 1)) e++;
    qfp G[2], C[3][4];
    q_cyc_compress(G, f);
    if (stage == 11) { qfp z; q_set_zero(z); q_cyc_gather(C[2], G); q_cyc_place(r, &z, &z, C[2]); return; }
    if (stage == 12) { qfp z; q_set_zero(z); q_cyc_sqr_compressed(G); q_cyc_gather(C[2], G); q_cyc_place(r, &z, &z, C[2]); return; }
#pragma unroll 1
    for (int i = 1; i <= e + 41; i++) {
        q_cyc_sqr_compressed(G);
        if (i == e) q_cyc_gather(C[0], G);
        if (i == e + 32) q_cyc_gather(C[1], G);
    }
    q_cyc_gather(C[2], G);
    if (stage >= 13 && stage <= 15) { qfp z; q_set_zero(z); q_cyc_place(r, &z, &z, C[stage - 13]); return; }
    // 1 / (4 g2) of the three values with one inversion (all lanes)
    qfp d[3], p01, inv;
    bool z[QL];
    QFOR bad[l_] = false;
    for (int i = 0; i < 3; i++) {
        qv_dbl(&d[i], &C[i][0], 1); qv_dbl(&d[i], &d[i], 1);
        q2_is_zero(z, d[i]);
        QFOR bad[l_] = bad[l_] || z[l_];
    }
    if (stage == 6) { q_cyc_place(r, &d[0], &d[1], C[0]); return; }
    q2_mul(&p01, &d[0], &d[1]);
    q2_mul(&inv, &p01, &d[2]);
    if (stage == 7) { q_cyc_place(r, &p01, &inv, C[0]); return; }
    q2_inv(&inv, &inv);
    if (stage == 8) { q_cyc_place(r, &inv, &d[2], C[0]); return; }
    q2_mul(&p01, &p01, &inv);                      // 1 / d2
    q2_mul(&inv, &inv, &d[2]);                     // 1 / (d0 d1)
    if (stage == 9) { q_cyc_place(r, &inv, &d[0], C[0]); return; }
    q2_mul(&d[2], &inv, &d[0]);                    // 1 / d1
    if (stage == 10) { q_cyc_place(r, &d[2], &d[1], C[0]); return; }
    q2_mul(&d[0], &inv, &d[1]);                    // 1 / d0
    // decompress value 0 on half 0 and value 1 on half 1 in one pass, value 2 on both halves in a second one
    qfp gs[4], iv, g0, g1, h0, h1;
    for (int k = 0; k < 4; k++) qv_selh(&gs[k], &C[0][k], &C[1][k], 1);
    qv_selh(&iv, &d[0], &d[2], 1);
    q_cyc_decompress(&g0, &g1, gs, &iv);
    q_cyc_decompress(&h0, &h1, C[2], &p01);
    q6 D0, D1, D2, own, oth;
    // own = the full value decompressed by the own half: (g0, g4, g3 | g2, g1, g5) -> D0 = half 0's, D1 = half 1's
    {
        q6 lo, hi;
        lo.c0 = g0; lo.c1 = gs[2]; lo.c2 = gs[1];          // the coefficients that live on half 0
        hi.c0 = gs[0]; hi.c1 = g1; hi.c2 = gs[3];          // the coefficients that live on half 1
        // half 0 keeps lo of value 0 and needs lo of value 1 (held by half 1); half 1 keeps hi of value 1, needs hi of value 0
        qv_selh(qa(&own, 0), qa(&hi, 0), qa(&lo, 0), 3);               // what the OTHER half is missing of my value
        qv_xh(qa(&oth, 0), qa(&own, 0), 3);
        qv_selh(qa(&D0, 0), qa(&lo, 0), qa(&oth, 0), 3);               // value 0: half 0 own lo | half 1 receives hi of value 0
        qv_selh(qa(&D1, 0), qa(&oth, 0), qa(&hi, 0), 3);               // value 1: half 0 receives lo of value 1 | half 1 own hi
    }
    q_cyc_place(&D2, &h0, &h1, C[2]);
    if (stage == 1) { *r = D0; return; }
    if (stage == 2) { *r = D1; return; }
    if (stage == 3) { *r = D2; return; }
    if (stage == 4) { q_cyc_place(r, &d[0], &d[2], C[0]); return; }
    if (stage == 5) { q_cyc_place(r, &p01, &p01, C[1]); return; }
    q6 R, T;
    q12_mul(&R, &D0, &D1);
    q12_mul(&R, &R, &D2);
    T = D2;
    q12_cyc_sqr(&T); q12_cyc_sqr(&T); q12_cyc_sqr(&T);
    q12_mul(&R, &R, &T);                           // 2^(e+44)
    q12_cyc_sqr(&T); q12_cyc_sqr(&T);
    q12_mul(&R, &R, &T);                           // 2^(e+46)
    q12_cyc_sqr(&T);
    q12_mul(&R, &R, &T);                           // 2^(e+47)
    q12_conj(&R, &R);
    bool b4[QL], nb[QL];
    QFOR nb[l_] = !bad[l_];
    q_all4(b4, nb);                                // the own quad is fine
    QFOR bad[l_] = !b4[l_];
    *r = R;
}
HDN void q_exp_by_x(q6 *r, const q6 *f, uint64_t x) {
    q6 R, T;
    bool bad[QL];
    q_exp_by_x_main(&R, bad, f, x);
    if (q_any(bad)) {
        q_exp_by_x_gs(&T, f, x);
        QFOR { if (bad[l_]) { R.c0.v[l_] = T.c0.v[l_]; R.c1.v[l_] = T.c1.v[l_]; R.c2.v[l_] = T.c2.v[l_]; } }
    }
    *r = R;
}

// FinalExponentiation (pairing.go:79-129; final_exp_one of pairing.cuh).  ok = false for f == 0 (the reference returns nil);
// the value is then unspecified (the kernel writes 1).  F is updated in place.
HD void q_final_exp(q6 *F, bool ok[QL]) {
    const uint64_t X = 0xd201000000010000ULL;
    q6 r, y0, y1, y2, y3;
    q12_conj(&y0, F);
    q12_inv(&y1, F, ok);
    q12_mul(&r, &y0, &y1);
    y1 = r;
    q12_frobenius(&r, &r, 2);
    q12_mul(&r, &r, &y1);                          // f^((q^6 - 1)(q^2 + 1)): cyclotomic from here on
    y0 = r;
    q12_cyc_sqr(&y0);
    q_exp_by_x(&y1, &y0, X);
    q_exp_by_x(&y2, &y1, X >> 1);
    q12_conj(&y3, &r);
    q12_mul(&y1, &y1, &y3);
    q12_conj(&y1, &y1);
    q12_mul(&y1, &y1, &y2);
    q_exp_by_x(&y2, &y1, X);
    q_exp_by_x(&y3, &y2, X);
    q12_conj(&y1, &y1);
    q12_mul(&y3, &y3, &y1);
    q12_conj(&y1, &y1);
    q12_frobenius(&y1, &y1, 3);
    q12_frobenius(&y2, &y2, 2);
    q12_mul(&y1, &y1, &y2);
    q_exp_by_x(&y2, &y3, X);
    q12_mul(&y2, &y2, &y0);
    q12_mul(&y2, &y2, &r);
    q12_mul(&y1, &y1, &y2);
    q12_frobenius(&y3, &y3, 1);
    q12_mul(F, &y1, &y3);
}

}  // namespace quad
}  // namespace b381
