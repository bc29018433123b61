// curve.cuh -- G1 (over Fq) and G2 (over Fq2) point arithmetic for aggregation and MSM.
// Replaces the group-law calls behind AggregatePublicKeys / AggregateSignatures
// (g1pubs/bls.go:177-204 -> G1Projective.Add g1.go:400-482, G2Projective.Add g2.go:446-529) and
// provides the bucket arithmetic of the Pippenger MSM the north star adds.
//
// Accumulators use extended Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2):
// a mixed addition costs 8M+2S against 7M+4S for the reference's madd-2007-bl (g1.go:485-559) and
// needs no Z.  Only affine results cross the ABI, and those are canonical field elements, so the
// coordinate system is invisible to callers; the exceptional cases the reference branches on
// (g1.go:401-437: either input zero, equal points) are handled the same way.
#pragma once
#include "tower.cuh"

namespace b381 {

// ---- field policies: same point code over Fq (inline or out-of-line multiplies) and Fq2 -------
struct FpInl {
    typedef fp T;
    static HD void mul(T &r, const T &a, const T &b) { fp_mul(r, a, b); }
    static HD void sqr(T &r, const T &a) { fp_sqr(r, a); }
    static HD void add(T &r, const T &a, const T &b) { fp_add(r, a, b); }
    static HD void sub(T &r, const T &a, const T &b) { fp_sub(r, a, b); }
    static HD void dbl(T &r, const T &a) { fp_dbl(r, a); }
    static HD void neg(T &r, const T &a) { fp_neg(r, a); }
    static HD bool is_zero(const T &a) { return fp_is_zero(a); }
    static HD void set_zero(T &r) { fp_set_zero(r); }
    static HD void set_one(T &r) { fp_set_one(r); }
    static HD void inv(T &r, const T &a) { fp_inv(&r, &a); }
};
// multiplications inlined at every use: for latency-bound single-thread chains (the MSM window shifts), where ptxas can
// overlap the independent products of a doubling; far too much code for the throughput kernels
struct FpIlp : FpInl {
    static HD void mul(T &r, const T &a, const T &b) { fp_mul_inl(r, a, b); }
    static HD void sqr(T &r, const T &a) { fp_mul_inl(r, a, a); }
};
template <class F> struct shift_policy { typedef F type; };
template <> struct shift_policy<FpInl> { typedef FpIlp type; };
// the dedicated squaring body (fp_sqr_v) for long squaring chains: square-root exponentiations
struct FpSqr : FpInl {
    static HD void sqr(T &r, const T &a) {
#if defined(B381_NO_FP_SQR)
        fp_sqr(r, a);
#else
        r = fp_sqr_v(a);
#endif
    }
};
// policy of a long accumulation loop over F (MSM chunk sums): same element type, the dedicated squaring where there is one
template <class F> struct acc_policy { typedef F type; };
template <> struct acc_policy<FpInl> { typedef FpSqr type; };
struct FpOut : FpInl {
    static HD void mul(T &r, const T &a, const T &b) { fp_mul_n(&r, &a, &b); }
    static HD void sqr(T &r, const T &a) { fp_sqr_n(&r, &a); }
};
struct Fp2Out {
    typedef fp2 T;
    static HD void mul(T &r, const T &a, const T &b) { fp2_mul(&r, &a, &b); }
    static HD void sqr(T &r, const T &a) { fp2_sqr(&r, &a); }
    static HD void add(T &r, const T &a, const T &b) { fp2_add(r, a, b); }
    static HD void sub(T &r, const T &a, const T &b) { fp2_sub(r, a, b); }
    static HD void dbl(T &r, const T &a) { fp2_dbl(r, a); }
    static HD void neg(T &r, const T &a) { fp2_neg(r, a); }
    static HD bool is_zero(const T &a) { return fp2_is_zero(a); }
    static HD void set_zero(T &r) { fp2_set_zero(r); }
    static HD void set_one(T &r) { fp2_set_one(r); }
    static HD void inv(T &r, const T &a) { fp2_inv(&r, &a); }
};

template <class F> struct xyzz { typename F::T x, y, zz, zzz; };   // infinity <=> zz == 0

template <class F> HD void xyzz_set_inf(xyzz<F> &p) {
    F::set_one(p.x); F::set_one(p.y); F::set_zero(p.zz); F::set_zero(p.zzz);
}
template <class F> HD bool xyzz_is_inf(const xyzz<F> &p) { return F::is_zero(p.zz); }

// EXPERIMENT, off by default (-DB381_GROUP_LAW_OOL): the group law out of line for the policies whose multiplications are calls
// already (FpOut, Fp2Out: codec, hashing, scalar multiplication, G2 sum / MSM).  Inlined at every use those kernels carry
// 0.1-0.6 MB of SASS and stall on instruction fetch (profiles/r01_v6_ncu_summary_hash.md).  Measured in round 2
// (profiles/r02_experiments.md): doubling / mixed addition out of line are correct and gain 1-2 % (HashG2WithDomain 27.5 ->
// 26.9 ms); the full addition out of line still gives WRONG G2 MSM results on the device (host build of the same source
// correct; cicc infers the address space of the callee's first pointer from its callers) -- so the inlined form ships.
#ifndef B381_OOL_DBLA
#define B381_OOL_DBLA 1
#endif
#ifndef B381_OOL_DBL
#define B381_OOL_DBL 1
#endif
#ifndef B381_OOL_MADD
#define B381_OOL_MADD 1
#endif
#ifndef B381_OOL_ADD
#define B381_OOL_ADD 0
#endif
template <class F> struct group_law_ool { static const bool value = false; };
struct FpOut; struct Fp2Out;
#if defined(B381_GROUP_LAW_OOL)
template <> struct group_law_ool<FpOut> { static const bool value = true; };
template <> struct group_law_ool<Fp2Out> { static const bool value = true; };
#endif
template <class F> HD void xyzz_dbl_affine_body(xyzz<F> &p, const typename F::T &x, const typename F::T &y);
template <class F> HD void xyzz_dbl_body(xyzz<F> &p);
template <class F> HD void xyzz_madd_body(xyzz<F> &p, const typename F::T &x2, const typename F::T &y2);
template <class F> HD void xyzz_add_body(xyzz<F> &p, const xyzz<F> &q);
template <class F> HDN void xyzz_dbl_affine_ool(xyzz<F> *p, const typename F::T *x, const typename F::T *y) { xyzz_dbl_affine_body<F>(*p, *x, *y); }
template <class F> HDN void xyzz_dbl_ool(xyzz<F> *p) { xyzz_dbl_body<F>(*p); }
template <class F> HDN void xyzz_madd_ool(xyzz<F> *p, const typename F::T *x2, const typename F::T *y2) { xyzz_madd_body<F>(*p, *x2, *y2); }
template <class F> HDN void xyzz_add_ool(xyzz<F> *p, const xyzz<F> *q) { xyzz_add_body<F>(*p, *q); }
template <class F> HD void xyzz_dbl_affine(xyzz<F> &p, const typename F::T &x, const typename F::T &y) {
    if constexpr (group_law_ool<F>::value && B381_OOL_DBLA) xyzz_dbl_affine_ool<F>(&p, &x, &y); else xyzz_dbl_affine_body<F>(p, x, y);
}
template <class F> HD void xyzz_dbl(xyzz<F> &p) {
    if constexpr (group_law_ool<F>::value && B381_OOL_DBL) xyzz_dbl_ool<F>(&p); else xyzz_dbl_body<F>(p);
}
template <class F> HD void xyzz_madd(xyzz<F> &p, const typename F::T &x2, const typename F::T &y2) {
    if constexpr (group_law_ool<F>::value && B381_OOL_MADD) xyzz_madd_ool<F>(&p, &x2, &y2); else xyzz_madd_body<F>(p, x2, y2);
}
template <class F> HD void xyzz_add(xyzz<F> &p, const xyzz<F> &q) {
    if constexpr (group_law_ool<F>::value && B381_OOL_ADD) xyzz_add_ool<F>(&p, &q); else xyzz_add_body<F>(p, q);
}

// p = 2 * (x, y) for a finite affine point (mdbl-2008-s-1, a = 0)
template <class F> HD void xyzz_dbl_affine_body(xyzz<F> &p, const typename F::T &x, const typename F::T &y) {
    typename F::T u, v, w, s, m, t;
    F::dbl(u, y);
    if (F::is_zero(u)) { xyzz_set_inf(p); return; }   // 2-torsion: not on these curves, kept for totality
    F::sqr(v, u);
    F::mul(w, u, v);
    F::mul(s, x, v);
    F::sqr(m, x);
    F::dbl(t, m); F::add(m, t, m);
    F::sqr(p.x, m);
    F::sub(p.x, p.x, s); F::sub(p.x, p.x, s);
    F::sub(t, s, p.x);
    F::mul(t, m, t);
    F::mul(u, w, y);
    F::sub(p.y, t, u);
    p.zz = v; p.zzz = w;
}
// p = 2p (dbl-2008-s-1, a = 0); same case split as G1Projective.Double (g1.go:343-397)
template <class F> HD void xyzz_dbl_body(xyzz<F> &p) {
    if (xyzz_is_inf(p)) return;
    typename F::T u, v, w, s, m, t;
    F::dbl(u, p.y);
    F::sqr(v, u);
    F::mul(w, u, v);
    F::mul(s, p.x, v);
    F::sqr(m, p.x);
    F::dbl(t, m); F::add(m, t, m);
    F::mul(u, w, p.y);                 // W*Y1 (uses the old y)
    F::sqr(p.x, m);
    F::sub(p.x, p.x, s); F::sub(p.x, p.x, s);
    F::sub(t, s, p.x);
    F::mul(t, m, t);
    F::sub(p.y, t, u);
    F::mul(p.zz, v, p.zz);
    F::mul(p.zzz, w, p.zzz);
}
// p += (x2, y2), a finite affine point (madd-2008-s); case split of G1Projective.AddAffine (g1.go:485-559)
template <class F> HD void xyzz_madd_body(xyzz<F> &p, const typename F::T &x2, const typename F::T &y2) {
    if (xyzz_is_inf(p)) { p.x = x2; p.y = y2; F::set_one(p.zz); F::set_one(p.zzz); return; }
    typename F::T u2, s2, pp, ppp, q, t;
    F::mul(u2, x2, p.zz);
    F::mul(s2, y2, p.zzz);
    F::sub(u2, u2, p.x);               // P
    F::sub(s2, s2, p.y);               // R
    if (F::is_zero(u2)) {
        if (F::is_zero(s2)) xyzz_dbl_affine(p, x2, y2);   // same point (g1.go:506-509)
        else xyzz_set_inf(p);                             // P + (-P)
        return;
    }
    F::sqr(pp, u2);
    F::mul(ppp, u2, pp);
    F::mul(q, p.x, pp);
    F::sqr(t, s2);
    F::sub(t, t, ppp);
    F::sub(t, t, q);
    F::sub(p.x, t, q);                 // X3 = R^2 - PPP - 2Q
    F::sub(q, q, p.x);
    F::mul(q, s2, q);
    F::mul(t, p.y, ppp);
    F::sub(p.y, q, t);                 // Y3 = R(Q - X3) - Y1*PPP
    F::mul(p.zz, p.zz, pp);
    F::mul(p.zzz, p.zzz, ppp);
}
// p += q (add-2008-s); case split of G1Projective.Add (g1.go:400-482)
template <class F> HD void xyzz_add_body(xyzz<F> &p, const xyzz<F> &q) {
    if (xyzz_is_inf(q)) return;
    if (xyzz_is_inf(p)) { p = q; return; }
    typename F::T u1, u2, s1, s2, pp, ppp, qq, t;
    F::mul(u1, p.x, q.zz);
    F::mul(u2, q.x, p.zz);
    F::mul(s1, p.y, q.zzz);
    F::mul(s2, q.y, p.zzz);
    F::sub(u2, u2, u1);                // P
    F::sub(s2, s2, s1);                // R
    if (F::is_zero(u2)) {
        if (F::is_zero(s2)) xyzz_dbl(p);
        else xyzz_set_inf(p);
        return;
    }
    F::sqr(pp, u2);
    F::mul(ppp, u2, pp);
    F::mul(qq, u1, pp);
    F::sqr(t, s2);
    F::sub(t, t, ppp);
    F::sub(t, t, qq);
    F::sub(p.x, t, qq);
    F::sub(qq, qq, p.x);
    F::mul(qq, s2, qq);
    F::mul(t, s1, ppp);
    F::sub(p.y, qq, t);
    F::mul(p.zz, p.zz, q.zz);
    F::mul(p.zz, p.zz, pp);
    F::mul(p.zzz, p.zzz, q.zzz);
    F::mul(p.zzz, p.zzz, ppp);
}
// normalised Jacobian out: (x, y, 1) or the canonical zero (0, 1, 0) of g1.go:269 / g2.go:310
template <class F> HD void xyzz_to_jac_normalised(typename F::T &ox, typename F::T &oy, typename F::T &oz, const xyzz<F> &p) {
    if (xyzz_is_inf(p)) { F::set_zero(ox); F::set_one(oy); F::set_zero(oz); return; }
    typename F::T i3, i2;
    F::inv(i3, p.zzz);                 // 1/ZZZ
    F::mul(i2, i3, p.zz);              // ZZ/ZZZ = 1/Z
    F::sqr(i2, i2);                    // 1/ZZ
    F::mul(ox, p.x, i2);
    F::mul(oy, p.y, i3);
    F::set_one(oz);
}
// Jacobian (X, Y, Z) -> XYZZ: ZZ = Z^2, ZZZ = Z^3
template <class F> HD void xyzz_from_jac(xyzz<F> &p, const typename F::T &X, const typename F::T &Y, const typename F::T &Z) {
    if (F::is_zero(Z)) { xyzz_set_inf(p); return; }
    p.x = X; p.y = Y;
    F::sqr(p.zz, Z);
    F::mul(p.zzz, p.zz, Z);
}
// p = k * p for a small non-negative integer k (double-and-add, MSB first; G1Affine.MulFR's schedule, g1.go:80-90)
template <class F> HD void xyzz_mul_small(xyzz<F> &p, uint32_t k) {
    xyzz<F> acc;
    xyzz_set_inf(acc);
    for (int b = 31; b >= 0; b--) {
        xyzz_dbl(acc);
        if ((k >> b) & 1) xyzz_add(acc, p);
    }
    p = acc;
}

// ---- Jacobian points for doubling-heavy work (ladders, window shifts): the reference's own formulas -------------------
// doubling 2M + 5S (G1Projective.Double, g1.go:343-397: dbl-2009-l), mixed addition 7M + 4S (AddAffine, g1.go:485-559:
// madd-2007-bl); the XYZZ doubling above costs 6M + 3S.
template <class F> struct jac_pt { typename F::T x, y, z; };          // infinity <=> z == 0
template <class F> HDN void jac_dbl(jac_pt<F> &p) {
    if (F::is_zero(p.z)) return;
    typename F::T a, b, c, d, e, f;
    F::sqr(a, p.x);
    F::sqr(b, p.y);
    F::sqr(c, b);
    F::add(d, p.x, b);
    F::sqr(d, d);
    F::sub(d, d, a);
    F::sub(d, d, c);
    F::dbl(d, d);                      // D = 2((X + B)^2 - A - C)
    F::dbl(e, a); F::add(e, e, a);     // E = 3A
    F::sqr(f, e);
    F::mul(p.z, p.y, p.z);
    F::dbl(p.z, p.z);                  // Z3 = 2 Y Z
    F::sub(p.x, f, d);
    F::sub(p.x, p.x, d);               // X3 = F - 2D
    F::sub(d, d, p.x);
    F::mul(d, e, d);
    F::dbl(c, c); F::dbl(c, c); F::dbl(c, c);
    F::sub(p.y, d, c);                 // Y3 = E(D - X3) - 8C
}
template <class F> HDN void jac_madd(jac_pt<F> &p, const typename F::T &x2, const typename F::T &y2) {
    if (F::is_zero(p.z)) { p.x = x2; p.y = y2; F::set_one(p.z); return; }
    typename F::T z1z1, u2, s2, h, hh, i, j, r, v, t;
    F::sqr(z1z1, p.z);
    F::mul(u2, x2, z1z1);
    F::mul(s2, y2, p.z);
    F::mul(s2, s2, z1z1);
    F::sub(h, u2, p.x);
    F::sub(r, s2, p.y);
    if (F::is_zero(h)) {
        if (F::is_zero(r)) { p.x = x2; p.y = y2; F::set_one(p.z); jac_dbl(p); }   // same point (g1.go:506-509)
        else { F::set_one(p.x); F::set_one(p.y); F::set_zero(p.z); }              // P + (-P)
        return;
    }
    F::dbl(r, r);
    F::sqr(hh, h);
    F::dbl(i, hh); F::dbl(i, i);       // I = 4 HH
    F::mul(j, h, i);
    F::mul(v, p.x, i);
    F::add(t, p.z, h);                 // Z3 = (Z1 + H)^2 - Z1Z1 - HH
    F::sqr(t, t);
    F::sub(t, t, z1z1);
    F::sub(p.z, t, hh);
    F::sqr(p.x, r);
    F::sub(p.x, p.x, j);
    F::sub(p.x, p.x, v);
    F::sub(p.x, p.x, v);               // X3 = r^2 - J - 2V
    F::sub(v, v, p.x);
    F::mul(v, r, v);
    F::mul(t, p.y, j);
    F::dbl(t, t);
    F::sub(p.y, v, t);                 // Y3 = r(V - X3) - 2 Y1 J
}
template <class F> HD void jac_to_xyzz(xyzz<F> &o, const jac_pt<F> &p) {
    if (F::is_zero(p.z)) { xyzz_set_inf(o); return; }
    o.x = p.x; o.y = p.y;
    F::sqr(o.zz, p.z);
    F::mul(o.zzz, o.zz, p.z);
}
// XYZZ -> Jacobian without an inversion: (X ZZ, Y ZZZ, ZZ) represents the same point (ZZZ^2 = ZZ^3)
template <class F> HD void xyzz_to_jac_pt(jac_pt<F> &o, const xyzz<F> &p) {
    if (xyzz_is_inf(p)) { F::set_one(o.x); F::set_one(o.y); F::set_zero(o.z); return; }
    F::mul(o.x, p.x, p.zz);
    F::mul(o.y, p.y, p.zzz);
    o.z = p.zz;
}

// ---- loads / stores of the ABI PODs --------------------------------------------------------------
HD void fp2_load_u64(fp2 &r, const uint64_t *p) { fp_load_u64(r.c0, p); fp_load_u64(r.c1, p + 6); }
HD void fp2_store_u64(uint64_t *p, const fp2 &a) { fp_store_u64(p, a.c0); fp_store_u64(p + 6, a.c1); }

}  // namespace b381

// ---- MSM building blocks (Pippenger bucket method; the reference has no MSM) -----
namespace b381 {

HD void load_affine(fp &x, fp &y, bool &inf, const g1_affine_pod *p) {
    fp_load_u64(x, p->x); fp_load_u64(y, p->y); inf = p->inf != 0;
}
HD void load_affine(fp2 &x, fp2 &y, bool &inf, const g2_affine_pod *p) {
    fp2_load_u64(x, p->x); fp2_load_u64(y, p->y); inf = p->inf != 0;
}
HD void store_jac(g1_jac_pod *o, const fp &x, const fp &y, const fp &z) {
    fp_store_u64(o->x, x); fp_store_u64(o->y, y); fp_store_u64(o->z, z);
}
HD void store_jac(g2_jac_pod *o, const fp2 &x, const fp2 &y, const fp2 &z) {
    fp2_store_u64(o->x, x); fp2_store_u64(o->y, y); fp2_store_u64(o->z, z);
}
HD void load_jac(fp &x, fp &y, fp &z, const g1_jac_pod *p) {
    fp_load_u64(x, p->x); fp_load_u64(y, p->y); fp_load_u64(z, p->z);
}

// c-bit digit number w of a 256-bit scalar (4 x u64, LS limb first): FRRepr bits [w*c, w*c + c)
HD uint32_t msm_digit(const uint64_t *k, int w, int c) {
    int bit = w * c;
    if (bit >= 256) return 0;
    int limb = bit >> 6, off = bit & 63;
    uint64_t v = k[limb] >> off;
    if (off + c > 64 && limb + 1 < 4) v |= k[limb + 1] << (64 - off);
    return (uint32_t)(v & ((1ull << c) - 1));
}
HD int msm_window_bits(size_t n) {
    int lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) lg++;
    int c = lg - 5;
    return c < 4 ? 4 : (c > 16 ? 16 : c);
}
HD int msm_num_windows(int c) { return (255 + c - 1) / c; }
// Signed c-bit digits of a scalar of at most nbits bits: k = sum_w d_w 2^(c w) with -2^(c-1) < d_w <= 2^(c-1), so a window needs
// 2^(c-1) buckets instead of 2^c - 1 (a negative digit adds -P: y -> Q - y).  W = msm_signed_windows(nbits, c) digits hold the
// final carry.  Returns digit w; the carry chain from window 0 is recomputed (a few integer operations per window).
HD int msm_signed_windows(int nbits, int c) { return (nbits + c) / c; }       // ceil((nbits + 1) / c)
HD int32_t msm_signed_digit(const uint64_t *k, int w, int c) {
    uint32_t carry = 0, half = 1u << (c - 1);
    int32_t d = 0;
    for (int v = 0; v <= w; v++) {
        uint32_t raw = msm_digit(k, v, c) + carry;
        if (raw > half) { d = (int32_t)raw - (int32_t)(1u << c); carry = 1; }
        else { d = (int32_t)raw; carry = 0; }
    }
    return d;
}

// sum of the points idx[lo..hi) into acc (one bucket); bit 31 of an index entry = add the NEGATED point (signed digits)
#define MSM_NEG_BIT 0x80000000u
template <class F, class APOD> HD void msm_bucket_sum(xyzz<F> &acc, const APOD *pts, const uint32_t *idx, uint32_t lo, uint32_t hi) {
    xyzz_set_inf(acc);
    for (uint32_t j = lo; j < hi; j++) {
        typename F::T x, y; bool inf;
        const uint32_t e = idx[j];
        load_affine(x, y, inf, pts + (e & ~MSM_NEG_BIT));
        if (e & MSM_NEG_BIT) F::neg(y, y);
        if (!inf) xyzz_madd(acc, x, y);
    }
}
// out = sum_{k in [lo, hi)} k * B[k]   (lo >= 1): running-sum trick inside the segment, then the
// segment offset (lo - 1) * sum(B) by a short double-and-add
template <class F> HD void msm_segment_reduce(xyzz<F> &out, const xyzz<F> *B, uint32_t lo, uint32_t hi) {
    xyzz<F> running, acc;
    xyzz_set_inf(running); xyzz_set_inf(acc);
    for (uint32_t k = hi; k-- > lo;) {
        xyzz<F> b = B[k];
        xyzz_add(running, b);
        xyzz_add(acc, running);
    }
    if (lo > 1) { xyzz_mul_small(running, lo - 1); xyzz_add(acc, running); }
    out = acc;
}
// acc = sum_j 2^(c * (w0 + j * wstep)) * S[j], j < nw: Horner from the top window down
template <class F> HD void msm_combine_windows(xyzz<F> &acc, const xyzz<F> *S, int nw, int c, int w0, int wstep) {
    xyzz_set_inf(acc);
    for (int j = nw - 1; j >= 0; j--) {
        xyzz<F> s = S[j];
        xyzz_add(acc, s);
        int shifts = (j > 0) ? c * wstep : c * w0;
        for (int i = 0; i < shifts; i++) xyzz_dbl(acc);
    }
}

}  // namespace b381
