// duo.cuh -- the pairing on TWO LANES per pairing: G2AffineToPrepared + MillerLoop + FinalExponentiation
// (g2.go:650-801, pairing.go:16-129) with every Fq2 value x = x0 + x1 u split over a lane pair (lane j holds x_j).
//
// Why two lanes: the one-pairing-per-thread kernels (pairing.cuh) keep 2.8 / 5.7 KB of Fq12 state per thread in local
// memory; with 2^16 threads resident that is 1.2 MB per SM against 228 KB of L1 and 180-370 MB against 126 MB of L2, so
// 30-46 GB per launch reach DRAM and `long_scoreboard` is the second stall (profiles/r01_v6_ncu_summary.md).  Splitting
// every Fq2 over two lanes halves the state per lane and lets the batch run as TWO rounds of 2^15 resident pairings:
// the stacks of one round fit the L2, and the hot part of a lane's stack has twice the chance to be in L1.
//   * Fq2 product: the partner's coefficients come by SHFL (24 per 444 wide MACs); lane 0 forms a0 b0 + a1 (Q - b1),
//     lane 1 forms a1 b0 + a0 b1 -- ONE two-product dot product with one Montgomery reduction per lane;
//   * Fq2 squaring: (a0 + a1)(a0 - a1) on lane 0, a1 (2 a0) on lane 1 -- one multiplication per lane;
//   * additions, subtractions, doublings, negations, scalings by an Fq value: lane-local, half the work per lane;
//   * multiplication by xi = 1 + u: one exchange; conjugation: lane 1 negates.
// Everything above the Fq2 level is the code of tower.cuh / pairing.cuh with an Fq2 replaced by its lane view: same
// formulas, same operation order, so the Miller value and the result are bit-identical to that path and the reference.
// The four-lane form (quad.cuh) splits the Fq12 once more; measured on B200 it pays more in exchanges and selections
// than it gains (profiles/r02_experiments.md), the two-lane form is the throughput path.
//
// Control flow is uniform over a warp (the shuffles name all 32 lanes): pairs at infinity and degenerate values are
// computed like any other and masked at the end.
// Host build: the lane primitives of quad.cuh loop over four emulated lanes = two independent copies of the pair.
#pragma once
#include "quad.cuh"

namespace b381 {
namespace duo {

using quad::qfp;
using quad::q6;
using quad::qa;
typedef qfp d2;                            // an Fq2 value: this lane's coefficient
struct d12 { q6 c0, c1; };                 // Fq12 = c0 + c1 w, six Fq per lane
struct dg2 { d2 x, y, z; };                // the running point of the Miller loop (Jacobian)
struct dflag { bool on[QL]; };

// n Fq2 values: r = a + b (sub = 0) or a - b (sub = 1), two elements per iteration while they last (two carry chains
// in flight, like fpv_addsub of tower.cuh)
HDN void dv_addsub(d2 *r, const d2 *a, const d2 *b, int n, int sub) {
    int i = 0;
#pragma unroll 1
    for (; i + 1 < n; i += 2) {
        QFOR {
            fp x0 = a[i].v[l_], y0 = b[i].v[l_], x1 = a[i + 1].v[l_], y1 = b[i + 1].v[l_];
            if (sub) { fp_sub(x0, x0, y0); fp_sub(x1, x1, y1); }
            else { fp_add(x0, x0, y0); fp_add(x1, x1, y1); }
            r[i].v[l_] = x0; r[i + 1].v[l_] = x1;
        }
    }
    if (i < n) {
        QFOR {
            fp x0 = a[i].v[l_], y0 = b[i].v[l_];
            if (sub) fp_sub(x0, x0, y0); else fp_add(x0, x0, y0);
            r[i].v[l_] = x0;
        }
    }
}
HDN void d2_add_p(d2 *r, const d2 *a, const d2 *b) {
    QFOR { fp x = a->v[l_], y = b->v[l_]; fp_add(x, x, y); r->v[l_] = x; }
}
HDN void d2_sub_p(d2 *r, const d2 *a, const d2 *b) {
    QFOR { fp x = a->v[l_], y = b->v[l_]; fp_sub(x, x, y); r->v[l_] = x; }
}
HD void d2_add(d2 &r, const d2 &a, const d2 &b) { d2_add_p(&r, &a, &b); }
HD void d2_sub(d2 &r, const d2 &a, const d2 &b) { d2_sub_p(&r, &a, &b); }
HD void d2_dbl(d2 &r, const d2 &a) { d2_add_p(&r, &a, &a); }
HD void d2_neg(d2 &r, const d2 &a) { quad::qv_neg(&r, &a, 1, 0); }
HD void d2_conj(d2 &r, const d2 &a) { quad::qv_neg(&r, &a, 1, 1); }
HD void d2_mul(d2 *r, const d2 *a, const d2 *b) { quad::q2_mul(r, a, b); }
HD void d2_sqr(d2 *r, const d2 *a) { quad::q2_sqr(r, a); }
HD void d2_mul_nr(d2 &r, const d2 &a) { quad::q2_mul_nr(&r, &a); }
HD void d2_mul_fp(d2 *r, const d2 *a, const d2 *s) { quad::qv_mul(r, a, s, 1); }      // s: an Fq value, the same on both lanes
HD void d2_inv(d2 *r, const d2 *a) { quad::q2_inv(r, a); }
HD void d2_set_one(d2 &r) { quad::q2_set_one(r); }
// r = 3t - 2z (plus = 0) or 3t + 2z (plus = 1) in one pass (fp2_tri of tower.cuh)
HDN void d2_tri(d2 *r, const d2 *t, const d2 *z, int plus) {
    QFOR {
        fp x = t->v[l_], y = z->v[l_], u;
        if (plus) fp_add(u, x, y); else fp_sub(u, x, y);
        fp_add(u, u, u);
        fp_add(u, u, x);
        r->v[l_] = u;
    }
}
// both lanes of the pair: is the Fq2 value zero
HD void d2_is_zero(dflag &z, const d2 &a) { quad::q2_is_zero(z.on, a); }
HD bool d_any(const dflag &f) { return quad::q_any(f.on); }
// per lane: the predicate holds on both lanes of the own pair
HD void d_both(dflag &r, const dflag &p) {
#if defined(__CUDA_ARCH__)
    unsigned b = __ballot_sync(0xffffffffu, p.on[0]);
    unsigned sh = (threadIdx.x & 31u) & ~1u;
    r.on[0] = ((b >> sh) & 3u) == 3u;
#else
    QFOR r.on[l_] = p.on[l_] && p.on[l_ ^ 1];
#endif
}

// ---- Fq6 (fq6.go): q6_mul, q6_mul_by_01, q6_inv, q6_mul_v of quad.cuh are lane-pair code already -------------------------
HD void d6_add(q6 *r, const q6 *a, const q6 *b) { dv_addsub(qa(r), qa(a), qa(b), 3, 0); }
HD void d6_sub(q6 *r, const q6 *a, const q6 *b) { dv_addsub(qa(r), qa(a), qa(b), 3, 1); }
HD void d6_neg(q6 *r, const q6 *a) { quad::qv_neg(qa(r), qa(a), 3, 0); }
// r = a * b   (fq6.go:255-292; fp6_mul of tower.cuh: outputs double as temporaries)
HDN void d6_mul(q6 *r, const q6 *a, const q6 *b) {
    d2 v0, v1, v2, s, t, x, y;
    d2_mul(&v0, &a->c0, &b->c0);
    d2_mul(&v1, &a->c1, &b->c1);
    d2_mul(&v2, &a->c2, &b->c2);
    d2_add(s, a->c1, a->c2);
    d2_add(t, b->c1, b->c2);
    d2_mul(&x, &s, &t);
    d2_sub(x, x, v1);
    d2_sub(x, x, v2);
    d2_mul_nr(x, x);
    d2_add(x, x, v0);
    d2_add(s, a->c0, a->c1);
    d2_add(t, b->c0, b->c1);
    d2_mul(&y, &s, &t);
    d2_sub(y, y, v0);
    d2_sub(y, y, v1);
    d2_mul_nr(s, v2);
    d2_add(y, y, s);
    d2_add(s, a->c0, a->c2);
    d2_add(t, b->c0, b->c2);
    d2_mul(&r->c2, &s, &t);                        // last use of a and b: r may alias them
    d2_sub(r->c2, r->c2, v0);
    d2_sub(r->c2, r->c2, v2);
    d2_add(r->c2, r->c2, v1);
    r->c0 = x; r->c1 = y;
}
// r = a * (b0 + b1 v)   (fq6.go:60-90)
HDN void d6_mul_by_01(q6 *r, const q6 *a, const d2 *b0, const d2 *b1) {
    d2 v0, v1, s, t, x, y;
    d2_mul(&v0, &a->c0, b0);
    d2_mul(&v1, &a->c1, b1);
    d2_add(s, a->c1, a->c2);
    d2_mul(&x, &s, b1);
    d2_sub(x, x, v1);
    d2_mul_nr(x, x);
    d2_add(x, x, v0);
    d2_add(s, a->c0, a->c1);
    d2_add(t, *b0, *b1);
    d2_mul(&y, &s, &t);
    d2_sub(y, y, v0);
    d2_sub(y, y, v1);
    d2_add(s, a->c0, a->c2);
    d2_mul(&r->c2, &s, b0);
    d2_sub(r->c2, r->c2, v0);
    d2_add(r->c2, r->c2, v1);
    r->c0 = x; r->c1 = y;
}
// r = a * (b1 v)   (fq6.go:40-57)
HDN void d6_mul_by_1(q6 *r, const q6 *a, const d2 *b1) {
    d2 x, y, z;
    d2_mul(&x, &a->c2, b1);
    d2_mul_nr(x, x);
    d2_mul(&y, &a->c0, b1);
    d2_mul(&z, &a->c1, b1);
    r->c0 = x; r->c1 = y; r->c2 = z;
}
// r = a^(q^power)   (fq6.go:211-218)
HDN void d6_frobenius(q6 *r, const q6 *a, int power) {
    q6 t = *a;
    d2 k;
    if (power & 1) quad::qv_neg(qa(&t), qa(&t), 3, 1);
    quad::q2_load_tab(k, B381_TAB(frob6_c1) + power * 24);
    d2_mul(&t.c1, &t.c1, &k);
    quad::q2_load_tab(k, B381_TAB(frob6_c2) + power * 24);
    d2_mul(&t.c2, &t.c2, &k);
    *r = t;
}
// r = a + v b, r = a - v b with v b = (xi b2, b0, b1) formed on the fly (tower.cuh)
HD void d6_add_vmul(q6 *r, const q6 *a, const q6 *b) {
    d2 x;
    d2_mul_nr(x, b->c2);
    d2_add(r->c0, a->c0, x);
    dv_addsub(qa(r, 1), qa(a, 1), qa(b, 0), 2, 0);
}
HD void d6_sub_vmul(q6 *r, const q6 *a, const q6 *b) {
    d2 x;
    d2_mul_nr(x, b->c2);
    d2_sub(r->c0, a->c0, x);
    dv_addsub(qa(r, 1), qa(a, 1), qa(b, 0), 2, 1);
}

// ---- Fq12 (fq12.go) ------------------------------------------------------------------------------------------------
HD void d12_set_one(d12 *r) {
    d2 one, z;
    d2_set_one(one); quad::q_set_zero(z);
    r->c0.c0 = one; r->c0.c1 = z; r->c0.c2 = z; r->c1.c0 = z; r->c1.c1 = z; r->c1.c2 = z;
}
HD void d12_conj(d12 *r, const d12 *a) {           // fq12.go:27-29
    if (r != a) r->c0 = a->c0;
    d6_neg(&r->c1, &a->c1);
}
// r = a * b   (fq12.go:198-213)
HDN void d12_mul(d12 *r, const d12 *a, const d12 *b) {
    q6 aa, bb, s, t;
    d6_mul(&aa, &a->c0, &b->c0);
    d6_mul(&bb, &a->c1, &b->c1);
    d6_add(&s, &a->c0, &a->c1);
    d6_add(&t, &b->c0, &b->c1);
    d6_mul(&s, &s, &t);
    d6_sub(&s, &s, &aa);
    d6_sub(&r->c1, &s, &bb);
    d6_add_vmul(&r->c0, &aa, &bb);
}
// r = a^2   (fq12.go:180-195)
HDN void d12_sqr(d12 *r, const d12 *a) {
    q6 ab, s, t;
    d6_mul(&ab, &a->c0, &a->c1);
    d6_add(&s, &a->c0, &a->c1);
    d6_add_vmul(&t, &a->c0, &a->c1);
    d6_mul(&s, &s, &t);
    d6_sub(&s, &s, &ab);
    d6_add(&r->c1, &ab, &ab);
    d6_sub_vmul(&r->c0, &s, &ab);
}
// f *= (d0 + d1 v) + (d4 v) w   (fq12.go:32-47)
HDN void d12_mul_by_014(d12 *f, const d2 *d0, const d2 *d1, const d2 *d4) {
    q6 aa, bb, s;
    d2 o;
    d6_mul_by_01(&aa, &f->c0, d0, d1);
    d6_mul_by_1(&bb, &f->c1, d4);
    d2_add(o, *d1, *d4);
    d6_add(&s, &f->c1, &f->c0);
    d6_mul_by_01(&s, &s, d0, &o);
    d6_sub(&s, &s, &aa);
    d6_sub(&f->c1, &s, &bb);
    d6_add_vmul(&f->c0, &aa, &bb);
}
// per lane: the own pair's Fq12 value is zero / one
HD void d12_is_zero(dflag &r, const d12 *a) {
    dflag p;
    const d2 *c = qa(a);
    QFOR { bool z = true; for (int i = 0; i < 6; i++) z = z && fp_is_zero(c[i].v[l_]); p.on[l_] = z; }
    d_both(r, p);
}
HD void d12_is_one(dflag &r, const d12 *a) {
    dflag p;
    d2 one;
    d2_set_one(one);
    const d2 *c = qa(a);
    QFOR { bool z = fp_eq(c[0].v[l_], one.v[l_]); for (int i = 1; i < 6; i++) z = z && fp_is_zero(c[i].v[l_]); p.on[l_] = z; }
    d_both(r, p);
}
// r = a^-1; ok = false where a == 0 (r is then 0)   (fq12.go:216-237)
HDN void d12_inv(d12 *r, const d12 *a, dflag &ok) {
    q6 t0, t1;
    d6_mul(&t0, &a->c0, &a->c0);
    d6_mul(&t1, &a->c1, &a->c1);
    quad::q6_mul_v(&t1, &t1);
    d6_sub(&t0, &t0, &t1);
    dflag z, p;
    QFOR p.on[l_] = fp_is_zero(t0.c0.v[l_]) && fp_is_zero(t0.c1.v[l_]) && fp_is_zero(t0.c2.v[l_]);
    d_both(z, p);
    QFOR ok.on[l_] = !z.on[l_];
    quad::q6_inv(&t0, &t0);
    d6_mul(&t1, &a->c1, &t0);
    d6_mul(&r->c0, &a->c0, &t0);
    d6_neg(&r->c1, &t1);
}
// r = a^(q^power)   (fq12.go:171-177)
HDN void d12_frobenius(d12 *r, const d12 *a, int power) {
    d6_frobenius(&r->c0, &a->c0, power);
    d6_frobenius(&r->c1, &a->c1, power);
    d2 k;
    quad::q2_load_tab(k, B381_TAB(frob12_c1) + power * 24);
    d2_mul(&r->c1.c0, &r->c1.c0, &k);
    d2_mul(&r->c1.c1, &r->c1.c1, &k);
    d2_mul(&r->c1.c2, &r->c1.c2, &k);
}
// Granger-Scott squaring in the cyclotomic subgroup (fp12_cyclotomic_sqr / fp4_sqr of tower.cuh), in place
HD void d4_sqr(d2 &o0, d2 &o1, const d2 &a, const d2 &b) {
    d2 t1;
    d2_sqr(&o0, &a);
    d2_sqr(&t1, &b);
    d2_add(o1, a, b);
    d2_sqr(&o1, &o1);
    d2_sub(o1, o1, o0);
    d2_sub(o1, o1, t1);
    d2_mul_nr(t1, t1);
    d2_add(o0, o0, t1);
}
HDN void d12_cyc_sqr(d12 *r, const d12 *a) {
    d2 t0, t1, t2, t3;
    d4_sqr(t0, t1, a->c0.c0, a->c1.c1);
    d2_tri(&r->c0.c0, &t0, &a->c0.c0, 0);
    d2_tri(&r->c1.c1, &t1, &a->c1.c1, 1);
    d4_sqr(t0, t1, a->c1.c0, a->c0.c2);
    d4_sqr(t2, t3, a->c0.c1, a->c1.c2);
    d2_tri(&r->c0.c1, &t0, &a->c0.c1, 0);
    d2_tri(&r->c1.c2, &t1, &a->c1.c2, 1);
    d2_mul_nr(t3, t3);
    d2_tri(&r->c1.c0, &t3, &a->c1.c0, 1);
    d2_tri(&r->c0.c2, &t2, &a->c0.c2, 0);
}

// ---- global memory <-> lanes ---------------------------------------------------------------------------------------
HD void d12_load(d12 *F, const uint64_t *src) {
    d2 *c = qa(F);
    for (int i = 0; i < 6; i++) quad::q_load(c[i], src, 2 * i, 1, 0);
}
HD void d12_store(uint64_t *dst, const d12 *F) {
    const d2 *c = qa(F);
    for (int i = 0; i < 6; i++) quad::q_store(dst, c[i], 2 * i, 1, 0);
}

// ---- Miller loop (g2.go:655-772, pairing.go:16-75; line_double / line_add / ell of pairing.cuh) ---------------------
HDN void dline_double(dg2 *r, d2 *o0, d2 *o1, d2 *o2) {
    d2 t0, t1, t2, t3, t4, t5, t6, zsq;
    d2_sqr(&t0, &r->x);
    d2_sqr(&t1, &r->y);
    d2_sqr(&t2, &t1);
    d2_add(t3, t1, r->x);
    d2_sqr(&t3, &t3);
    d2_sub(t3, t3, t0);
    d2_sub(t3, t3, t2);
    d2_dbl(t3, t3);
    d2_dbl(t4, t0);
    d2_add(t4, t4, t0);
    d2_add(t6, r->x, t4);
    d2_sqr(&t5, &t4);
    d2_sqr(&zsq, &r->z);
    d2_sub(r->x, t5, t3);
    d2_sub(r->x, r->x, t3);
    d2_add(r->z, r->z, r->y);
    d2_sqr(&r->z, &r->z);
    d2_sub(r->z, r->z, t1);
    d2_sub(r->z, r->z, zsq);
    d2_sub(r->y, t3, r->x);
    d2_mul(&r->y, &r->y, &t4);
    d2_dbl(t2, t2); d2_dbl(t2, t2); d2_dbl(t2, t2);
    d2_sub(r->y, r->y, t2);
    d2_mul(&t3, &t4, &zsq);
    d2_dbl(t3, t3);
    d2_neg(*o1, t3);
    d2_sqr(&t6, &t6);
    d2_sub(t6, t6, t0);
    d2_sub(t6, t6, t5);
    d2_dbl(t1, t1); d2_dbl(t1, t1);
    d2_sub(*o2, t6, t1);
    d2_mul(&t0, &r->z, &zsq);
    d2_dbl(*o0, t0);
}
HDN void dline_add(dg2 *r, const d2 *qx, const d2 *qy, d2 *o0, d2 *o1, d2 *o2) {
    d2 zsq, ysq, t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10;
    d2_sqr(&zsq, &r->z);
    d2_sqr(&ysq, qy);
    d2_mul(&t0, &zsq, qx);
    d2_add(t1, *qy, r->z);
    d2_sqr(&t1, &t1);
    d2_sub(t1, t1, ysq);
    d2_sub(t1, t1, zsq);
    d2_mul(&t1, &t1, &zsq);
    d2_sub(t2, t0, r->x);
    d2_sqr(&t3, &t2);
    d2_dbl(t4, t3); d2_dbl(t4, t4);
    d2_mul(&t5, &t4, &t2);
    d2_sub(t6, t1, r->y);
    d2_sub(t6, t6, r->y);
    d2_mul(&t9, &t6, qx);
    d2_mul(&t7, &t4, &r->x);
    d2_sqr(&r->x, &t6);
    d2_sub(r->x, r->x, t5);
    d2_sub(r->x, r->x, t7);
    d2_sub(r->x, r->x, t7);
    d2_add(r->z, r->z, t2);
    d2_sqr(&r->z, &r->z);
    d2_sub(r->z, r->z, zsq);
    d2_sub(r->z, r->z, t3);
    d2_add(t10, *qy, r->z);
    d2_sub(t8, t7, r->x);
    d2_mul(&t8, &t8, &t6);
    d2_mul(&t0, &r->y, &t5);
    d2_dbl(t0, t0);
    d2_sub(r->y, t8, t0);
    d2_sqr(&t10, &t10);
    d2_sub(t10, t10, ysq);
    d2_sqr(&zsq, &r->z);
    d2_sub(t10, t10, zsq);
    d2_dbl(t9, t9);
    d2_sub(*o2, t9, t10);
    d2_dbl(*o0, r->z);
    d2_neg(t6, t6);
    d2_dbl(*o1, t6);
}
HD void dell(d12 *f, d2 *c0, d2 *c1, const d2 *c2, const d2 *px, const d2 *py) {
    d2_mul_fp(c0, c0, py);
    d2_mul_fp(c1, c1, px);
    d12_mul_by_014(f, c2, c1, c0);
}
struct dpair { d2 px, py, qx, qy; dg2 r; dflag live; };
HD void dpair_load(dpair *S, const g1_affine_pod *P, const g2_affine_pod *Q) {
    quad::q_load(S->px, P->x, 0, 0, 0); quad::q_load(S->py, P->y, 0, 0, 0);
    quad::q_load(S->qx, Q->x, 0, 1, 0); quad::q_load(S->qy, Q->y, 0, 1, 0);
    S->r.x = S->qx; S->r.y = S->qy; d2_set_one(S->r.z);
    QFOR S->live.on[l_] = !(P->inf || Q->inf);
}
HD void d12_keep_if(d12 *F, const d12 *G, const dflag &keep) {       // F <- keep ? F : G
    d2 *f = qa(F); const d2 *g = qa(G);
    QFOR { if (!keep.on[l_]) for (int i = 0; i < 6; i++) f[i].v[l_] = g[i].v[l_]; }
}
// Miller loop of NP pairs sharing the accumulator, conjugated; a pair with P or Q at infinity contributes the factor 1
template <int NP>
HD void d_miller_loop(d12 *f, dpair *S) {
    d12_set_one(f);
    d2 c0, c1, c2;
    const uint64_t xr = 0xd201000000010000ULL >> 1;
#pragma unroll 1
    for (int bit = 61; bit >= -1; bit--) {
#pragma unroll 1
        for (int k = 0; k < NP; k++) {
            d12 g;
            if (NP > 1) g = *f;
            dline_double(&S[k].r, &c0, &c1, &c2);
            dell(f, &c0, &c1, &c2, &S[k].px, &S[k].py);
            if (bit >= 0 && ((xr >> bit) & 1)) {
                dline_add(&S[k].r, &S[k].qx, &S[k].qy, &c0, &c1, &c2);
                dell(f, &c0, &c1, &c2, &S[k].px, &S[k].py);
            }
            if (NP > 1) d12_keep_if(f, &g, S[k].live);
        }
        if (bit >= 0) d12_sqr(f, f);
    }
    d12_conj(f, f);
    if (NP == 1) { d12 one; d12_set_one(&one); d12_keep_if(f, &one, S[0].live); }
}

// ---- final exponentiation (pairing.go:79-129; pairing.cuh) -----------------------------------------------------------
struct dcyc4 { d2 g2, g3, g4, g5; };
HD void dcyc_compress(dcyc4 *c, const d12 *f) { c->g2 = f->c1.c0; c->g3 = f->c0.c2; c->g4 = f->c0.c1; c->g5 = f->c1.c2; }
HDN void dcyc_sqr_compressed(dcyc4 *c) {
    d2 A, B, t0, t1, n2, n3;
    d2_mul_nr(t0, c->g5); d2_add(t0, t0, c->g4);
    d2_add(t1, c->g4, c->g5);
    d2_mul(&A, &t0, &t1);
    d2_mul(&B, &c->g4, &c->g5);
    d2_mul_nr(t0, B);
    d2_sub(A, A, t0); d2_sub(A, A, B);
    d2_dbl(t0, t0);
    d2_tri(&n2, &t0, &c->g2, 1);
    d2_tri(&n3, &A, &c->g3, 0);
    d2_mul_nr(t0, c->g3); d2_add(t0, t0, c->g2);
    d2_add(t1, c->g2, c->g3);
    d2_mul(&A, &t0, &t1);
    d2_mul(&B, &c->g2, &c->g3);
    d2_mul_nr(t0, B);
    d2_sub(A, A, t0); d2_sub(A, A, B);
    d2_tri(&c->g4, &A, &c->g4, 0);
    d2_dbl(t0, B);
    d2_tri(&c->g5, &t0, &c->g5, 1);
    c->g2 = n2; c->g3 = n3;
}
HDN void dcyc_decompress(d12 *f, const dcyc4 *c, const d2 *inv4g2) {
    d2 t0, t1, g1;
    d2_sqr(&t0, &c->g5); d2_mul_nr(t0, t0);
    d2_sqr(&t1, &c->g4);
    d2_tri(&t1, &t1, &c->g3, 0);
    d2_add(t0, t0, t1);
    d2_mul(&g1, &t0, inv4g2);
    d2_sqr(&t0, &g1); d2_dbl(t0, t0);
    d2_mul(&t1, &c->g2, &c->g5); d2_add(t0, t0, t1);
    d2_mul(&t1, &c->g3, &c->g4);
    d2_sub(t0, t0, t1); d2_dbl(t1, t1); d2_sub(t0, t0, t1);
    d2_mul_nr(t0, t0);
    d2_set_one(t1);
    d2_add(f->c0.c0, t0, t1);
    f->c1.c1 = g1; f->c1.c0 = c->g2; f->c0.c2 = c->g3; f->c0.c1 = c->g4; f->c1.c2 = c->g5;
}
HDN void d_exp_by_x_gs(d12 *r, const d12 *f, uint64_t x) {
    d12 acc = *f;
    int top = 63;
    while (!((x >> top) & 1)) top--;
#pragma unroll 1
    for (int bit = top - 1; bit >= 0; bit--) {
        d12_cyc_sqr(&acc, &acc);
        if ((x >> bit) & 1) d12_mul(&acc, &acc, f);
    }
    d12_conj(r, &acc);
}
// conj(f^x) for |x_BLS| and |x_BLS| / 2 (exp_by_x of pairing.cuh).  bad = the own pair's compressed form degenerated
// (some g2 = 0 on the way: f = 1, ...): the result is then meaningless and the caller takes d_exp_by_x_gs.
HDN void d_exp_by_x_main(d12 *r, dflag &bad, const d12 *f, uint64_t x) {
    int e = 0;
    while (!((x >> e) & 1)) e++;
    dcyc4 c[3];
    dcyc_compress(&c[2], f);
#pragma unroll 1
    for (int i = 1; i <= e + 41; i++) {
        dcyc_sqr_compressed(&c[2]);
        if (i == e) c[0] = c[2];
        if (i == e + 32) c[1] = c[2];
    }
    d2 d[3], p01, inv;
    dflag z, nz;
    QFOR nz.on[l_] = true;
    for (int i = 0; i < 3; i++) {
        d2_dbl(d[i], c[i].g2); d2_dbl(d[i], d[i]);
        d2_is_zero(z, d[i]);
        QFOR nz.on[l_] = nz.on[l_] && !z.on[l_];
    }
    d_both(z, nz);
    QFOR bad.on[l_] = !z.on[l_];
    d2_mul(&p01, &d[0], &d[1]);
    d2_mul(&inv, &p01, &d[2]);
    d2_inv(&inv, &inv);
    d2_mul(&p01, &p01, &inv);                      // 1 / d2
    d2_mul(&inv, &inv, &d[2]);                     // 1 / (d0 d1)
    d2_mul(&d[2], &inv, &d[0]);                    // 1 / d1
    d2_mul(&d[0], &inv, &d[1]);                    // 1 / d0
    d12 t;
    dcyc_decompress(r, &c[0], &d[0]);
    dcyc_decompress(&t, &c[1], &d[2]);
    d12_mul(r, r, &t);
    dcyc_decompress(&t, &c[2], &p01);
    d12_mul(r, r, &t);
    d12_cyc_sqr(&t, &t); d12_cyc_sqr(&t, &t); d12_cyc_sqr(&t, &t);
    d12_mul(r, r, &t);
    d12_cyc_sqr(&t, &t); d12_cyc_sqr(&t, &t);
    d12_mul(r, r, &t);
    d12_cyc_sqr(&t, &t);
    d12_mul(r, r, &t);
    d12_conj(r, r);
}
// r must not alias f
HDN void d_exp_by_x(d12 *r, const d12 *f, uint64_t x) {
    dflag bad;
    d_exp_by_x_main(r, bad, f, x);
    if (d_any(bad)) {                              // some pair of the warp: everybody runs the square-and-multiply form too
        d12 t;
        d_exp_by_x_gs(&t, f, x);
        dflag good;
        QFOR good.on[l_] = !bad.on[l_];
        d12_keep_if(r, &t, good);
    }
}
// FinalExponentiation in place (final_exp_one of pairing.cuh); ok = false where f == 0 (the value is then unspecified)
HD void d_final_exp(d12 *out, dflag &ok) {
    const uint64_t X = 0xd201000000010000ULL;
    d12 r, y0, y2, y3;
    d12 *y1 = out;
    d12_conj(&y0, out);
    d12_inv(y1, out, ok);
    d12_mul(&r, &y0, y1);
    *y1 = r;
    d12_frobenius(&r, &r, 2);
    d12_mul(&r, &r, y1);
    d12_cyc_sqr(&y0, &r);
    d_exp_by_x(y1, &y0, X);
    d_exp_by_x(&y2, y1, X >> 1);
    d12_conj(&y3, &r);
    d12_mul(y1, y1, &y3);
    d12_conj(y1, y1);
    d12_mul(y1, y1, &y2);
    d_exp_by_x(&y2, y1, X);
    d_exp_by_x(&y3, &y2, X);
    d12_conj(y1, y1);
    d12_mul(&y3, &y3, y1);
    d12_conj(y1, y1);
    d12_frobenius(y1, y1, 3);
    d12_frobenius(&y2, &y2, 2);
    d12_mul(y1, y1, &y2);
    d_exp_by_x(&y2, &y3, X);
    d12_mul(&y2, &y2, &y0);
    d12_mul(&y2, &y2, &r);
    d12_mul(y1, y1, &y2);
    d12_frobenius(&y3, &y3, 1);
    d12_mul(out, y1, &y3);
}

}  // namespace duo
}  // namespace b381
