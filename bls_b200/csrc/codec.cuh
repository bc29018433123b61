// codec.cuh -- wire formats and scalar multiplication on the device: what the callers of
// the verification path do on either side of it.
//   DecompressG1 / DecompressG1Unchecked / GetG1PointFromX / IsInCorrectSubgroupAssumingOnCurve   g1.go:111-141,185-227
//   DecompressG2 / DecompressG2Unchecked / GetG2PointFromX                                        g2.go:149-169,219-265,293-295
//   CompressG1 / CompressG2                                                                       g1.go:230-249, g2.go:268-289
//   FQ.Sqrt (fq.go:203-217), FQ2.Sqrt (fq2.go:198-232), FQ.Cmp / FQ2.Cmp (fq.go:134-137, fq2.go:31-37)
//   G1Affine.MulFR / G2Affine.MulFR + ToAffine (g1.go:80-90,322-340; g2.go:92-102,365-386): PrivToPub and Sign
// One thread per point.  Results are canonical field elements / bytes / status codes, so they are the reference's
// bits whatever formulas run underneath (here: Jacobian / XYZZ double-and-add from curve.cuh and the integer almost-inverse).
#pragma once
#include "curve.cuh"

#ifndef CODEC_MIN_BLOCKS
#define CODEC_MIN_BLOCKS 8   // 128 registers: measured on B200 G2 scalar multiplication 48.9 -> 40.1 ms, hash 85.5 -> 77.3 ms vs the unconstrained build (140-200 registers)
#endif

namespace b381 {

// status codes of the decompression entry points (0 = ok); the reference returns errors with these messages
enum {
    CODEC_OK = 0,
    CODEC_ERR_MODE = 1,       // "unexpected compression mode"                     g1.go:203-205, g2.go:233-235
    CODEC_ERR_INFINITY = 2,   // "unexpected information in compressed infinity"  g1.go:211-214, g2.go:241-244
    CODEC_ERR_NOT_ON_CURVE = 3,   // "point not on curve"                          g1.go:119-121, g2.go:157-159
    CODEC_ERR_SUBGROUP = 4    // "not in correct subgroup"                         g1.go:191-193, g2.go:225-227
};

#if defined(__CUDACC__)
__device__ __constant__ uint32_t d_r2_raw[12] = {B381_R2_RAW_LIMBS};
__device__ __constant__ uint32_t d_neg_one[12] = {B381_NEG_ONE_LIMBS};
__device__ __constant__ uint32_t d_b_coeff[12] = {B381_B_COEFF_LIMBS};
__device__ __constant__ uint32_t d_qm3o4[12] = {B381_Q_MINUS_3_OVER_4_LIMBS};
__device__ __constant__ uint32_t d_qm1o2[12] = {B381_Q_MINUS_1_OVER_2_LIMBS};
__device__ __constant__ uint32_t d_r_order[8] = {B381_R_ORDER_LIMBS};
__device__ __constant__ uint32_t d_beta[12] = {B381_BETA_LIMBS};
__device__ __constant__ uint32_t d_psi_cx_c1[12] = {B381_PSI_CX_C1_LIMBS};
__device__ __constant__ uint32_t d_psi_cy[24] = {B381_PSI_CY_LIMBS};
__device__ __constant__ uint32_t d_bls_x[2] = {B381_BLS_X_LIMBS};
__device__ __constant__ uint32_t d_bls_x2[4] = {B381_BLS_X2_LIMBS};
#endif
#if !defined(__CUDA_ARCH__)
static const uint32_t h_r2_raw[12] = {B381_R2_RAW_LIMBS};
static const uint32_t h_neg_one[12] = {B381_NEG_ONE_LIMBS};
static const uint32_t h_b_coeff[12] = {B381_B_COEFF_LIMBS};
static const uint32_t h_qm3o4[12] = {B381_Q_MINUS_3_OVER_4_LIMBS};
static const uint32_t h_qm1o2[12] = {B381_Q_MINUS_1_OVER_2_LIMBS};
static const uint32_t h_r_order[8] = {B381_R_ORDER_LIMBS};
static const uint32_t h_beta[12] = {B381_BETA_LIMBS};
static const uint32_t h_psi_cx_c1[12] = {B381_PSI_CX_C1_LIMBS};
static const uint32_t h_psi_cy[24] = {B381_PSI_CY_LIMBS};
static const uint32_t h_bls_x[2] = {B381_BLS_X_LIMBS};
static const uint32_t h_bls_x2[4] = {B381_BLS_X2_LIMBS};
#endif

// ---- integers <-> field elements -----------------------------------------------------------------------------------
// big-endian 48 bytes -> plain limbs (FQReprFromBytes, fqrepr.go:194-202)
HD void fp_raw_from_be48(fp &r, const uint8_t *b) {
#pragma unroll 1
    for (int i = 0; i < 12; i++) {
        const uint8_t *p = b + 44 - 4 * i;
        r.l[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
    }
}
HD void fp_raw_to_be48(uint8_t *b, const fp &a) {   // FQRepr.Bytes, fqrepr.go:181-191
#pragma unroll 1
    for (int i = 0; i < 12; i++) {
        uint8_t *p = b + 44 - 4 * i;
        uint32_t v = a.l[i];
        p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v;
    }
}
// -1, 0, +1 on plain limbs (FQRepr.Cmp, fqrepr.go:125-136)
HD int fp_raw_cmp(const fp &a, const fp &b) {
    int c = 0;
#pragma unroll 1
    for (int i = 11; i >= 0; i--)
        if (c == 0 && a.l[i] != b.l[i]) c = a.l[i] < b.l[i] ? -1 : 1;
    return c;
}
// plain integer -> Montgomery element; values >= Q become 0 exactly like FQReprToFQ (fq.go:49-56)
HD void fp_from_raw(fp &r, const fp &raw) {
    const uint32_t q[12] = {B381_Q_LIMBS};
    fp qq, k;
#pragma unroll
    for (int i = 0; i < 12; i++) qq.l[i] = q[i];
    if (fp_raw_cmp(raw, qq) >= 0) { fp_set_zero(r); return; }
    fp_load_tab(k, B381_TAB(r2_raw));
    fp_mul(r, raw, k);
}
// Montgomery element -> plain integer (FQ.ToRepr, fq.go:334-338)
HD void fp_to_raw(fp &r, const fp &a) {
    fp one;
    fp_set_zero(one); one.l[0] = 1;
    fp_mul(r, a, one);
}
// FQ.Cmp (fq.go:134-137): compares the plain integers
HD int fp_cmp(const fp &a, const fp &b) {
    fp x, y;
    fp_to_raw(x, a); fp_to_raw(y, b);
    return fp_raw_cmp(x, y);
}
// FQ2.Cmp (fq2.go:31-37): c1 first
HD int fp2_cmp(const fp2 &a, const fp2 &b) {
    int c = fp_cmp(a.c1, b.c1);
    return c ? c : fp_cmp(a.c0, b.c0);
}

// ---- exponentiation by a fixed 384-bit exponent and square roots -----------------------------------------------------
// a^e, e as 12 plain u32 limbs (FQ.Exp fq.go:268-284 / FQ2.Exp fq2.go:170-187 are MSB-first square-and-multiply; the value
// is the same): sliding window of four bits over the odd powers a, a^3 .. a^15 -- for the 381-bit square-root exponents
// ~380 squarings + ~84 multiplications instead of ~190.  The exponent is a constant, so a warp never diverges here.
template <class F> HDN void field_pow(typename F::T *r, const typename F::T *a, const uint32_t *e) {
    typename F::T tab[8], acc, x2;
    tab[0] = *a;
    F::sqr(x2, tab[0]);
#pragma unroll 1
    for (int k = 1; k < 8; k++) F::mul(tab[k], tab[k - 1], x2);
    F::set_one(acc);
    bool started = false;
    int i = 383;
#pragma unroll 1
    while (i >= 0) {
        if (!((e[i >> 5] >> (i & 31)) & 1)) {
            if (started) F::sqr(acc, acc);
            i--;
            continue;
        }
        int j = i >= 3 ? i - 3 : 0;
        while (!((e[j >> 5] >> (j & 31)) & 1)) j++;               // the window ends on a set bit
        uint32_t val = 0;
        for (int b = i; b >= j; b--) val = (val << 1) | ((e[b >> 5] >> (b & 31)) & 1);
        if (started) {
#pragma unroll 1
            for (int b = i; b >= j; b--) F::sqr(acc, acc);
            F::mul(acc, acc, tab[val >> 1]);
        } else { acc = tab[val >> 1]; started = true; }
        i = j - 1;
    }
    *r = acc;
}
// FQ.Sqrt (fq.go:203-217): q = 3 mod 4
HDN bool fp_sqrt(fp *out, const fp *a) {
    fp a1, a0, m1;
    field_pow<FpSqr>(&a1, a, B381_TAB(qm3o4));
    fp_sqr(a0, a1);
    fp_mul(a0, a0, *a);
    fp_load_tab(m1, B381_TAB(neg_one));
    if (fp_eq(a0, m1)) return false;
    fp_mul(*out, a1, *a);
    return true;
}
// Is a a square in Fq?  Jacobi symbol (a / Q) by the binary algorithm on the twelve limbs (subtract, strip the factors of two,
// quadratic reciprocity on swaps): ~20 k integer instructions instead of the 570-multiplication Euler exponentiation
// (~190 k), used where a failed FQ.Sqrt / FQ2.Sqrt only steers a search (HashG2WithDomain's x + 1 loop g2.go:1056-1064, the
// first SWU candidate g1.go:679-700).  The symbol of the Montgomery representative a 2^384 equals that of a (2^384 is a
// square), so no conversion is needed.  0 counts as a square, as FQ.Sqrt(0) = 0 succeeds.
HDN bool fp_is_square(const fp &v) {
    const uint32_t q[12] = {B381_Q_LIMBS};
    uint32_t a[12], n[12], t = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) { a[i] = v.l[i]; n[i] = q[i]; }
    for (;;) {
        // strip the factors of two of a: (2 / n) = -1 iff n = 3, 5 mod 8
        uint32_t low = a[0];
        if (low == 0) {
            uint32_t any = 0;
#pragma unroll
            for (int i = 1; i < 12; i++) any |= a[i];
            if (any == 0) break;                                   // a == 0: n holds gcd(v, Q)
#pragma unroll
            for (int i = 0; i < 11; i++) a[i] = a[i + 1];          // 32 factors of two: an even count, no sign change
            a[11] = 0;
            continue;
        }
#if defined(__CUDA_ARCH__)
        uint32_t z = __ffs(low) - 1;
#else
        uint32_t z = (uint32_t)__builtin_ctz(low);
#endif
        if (z) {
#pragma unroll
            for (int i = 0; i < 11; i++) a[i] = (a[i] >> z) | (a[i + 1] << (32 - z));
            a[11] >>= z;
            uint32_t n8 = n[0] & 7u;
            if ((z & 1u) && (n8 == 3u || n8 == 5u)) t ^= 1u;
        }
        // both odd: d = a - n; on a borrow swap roles (reciprocity: sign flips iff a = n = 3 mod 4) and negate
        uint32_t d[12];
        uint64_t br = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            uint64_t x = (uint64_t)a[i] - n[i] - br;
            d[i] = (uint32_t)x;
            br = (x >> 32) & 1u;
        }
        if (br) {
            if ((a[0] & n[0] & 3u) == 3u) t ^= 1u;
            uint64_t c = 1;
#pragma unroll
            for (int i = 0; i < 12; i++) {
                n[i] = a[i];
                uint64_t x = (uint64_t)(~d[i]) + c;
                d[i] = (uint32_t)x;
                c = x >> 32;
            }
        }
#pragma unroll
        for (int i = 0; i < 12; i++) a[i] = d[i];
    }
    // gcd is 1 unless v == 0 (Q is prime)
    uint32_t rest = n[0] ^ 1u;
#pragma unroll
    for (int i = 1; i < 12; i++) rest |= n[i];
    return rest != 0 || t == 0;
}
// r = a / 2: (a + (a odd ? Q : 0)) >> 1, valid on the Montgomery representative as well
HD void fp_half(fp &r, const fp &a) {
    const uint32_t q[12] = {B381_Q_LIMBS};
    uint32_t m = 0u - (a.l[0] & 1u), t[13];
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        uint64_t v = (uint64_t)a.l[i] + (q[i] & m) + c;
        t[i] = (uint32_t)v;
        c = v >> 32;
    }
    t[12] = (uint32_t)c;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = (t[i] >> 1) | (t[i + 1] << 31);
}
// FQ2.Sqrt (fq2.go:198-232; Algorithm 9 of eprint 2012/685), the form the reference runs: two Fq2 exponentiations
HDN bool fp2_sqrt_alg9(fp2 *out, const fp2 *a) {
    if (fp2_is_zero(*a)) { fp2_set_zero(*out); return true; }
    fp2 a1, alpha, a0, neg1;
    field_pow<Fp2Out>(&a1, a, B381_TAB(qm3o4));
    fp2_sqr(&alpha, &a1);
    fp2_mul(&alpha, &alpha, a);
    fp2_conj(a0, alpha);                       // Frobenius(1)
    fp2_mul(&a0, &a0, &alpha);
    fp_load_tab(neg1.c0, B381_TAB(neg_one)); fp_set_zero(neg1.c1);
    if (fp2_eq(a0, neg1)) return false;
    fp2_mul(&a1, &a1, a);
    if (fp2_eq(alpha, neg1)) {                 // multiply by u
        fp t = a1.c0;
        fp_neg(a1.c0, a1.c1);
        a1.c1 = t;
        *out = a1;
        return true;
    }
    fp2 one;
    fp2_set_one(one);
    fp2_add(alpha, alpha, one);
    field_pow<Fp2Out>(&alpha, &alpha, B381_TAB(qm1o2));
    fp2_mul(out, &alpha, &a1);
    return true;
}
// A square root of a in Fq2 = Fq[u]/(u^2 + 1) through the norm, two Fq exponentiations instead of two Fq2 ones:
//   n = a0^2 + a1^2 must be a square s^2 in Fq (else a is a non-square);  t = (a0 + s) / 2;  w = t^((Q-3)/4);  x = w t
//   w^2 t = +1:  (x, a1 w / 2)^2 = a          w^2 t = -1 (t a non-residue, -1 being one):  (a1 w / 2, -x)^2 = a
// The root may be the negative of the one FQ2.Sqrt returns; every caller in the reference normalises the sign afterwards
// (GetG2PointFromX g2.go:247-262, HashG2WithDomain g2.go:1066-1070, the SWU sign rule g2.go:1024-1027), which is why the
// results stay bit-identical (tests compare with fp2_sqrt_alg9 up to sign and with the oracle end to end).
// second half, given s with s^2 = a0^2 + a1^2 and a1 != 0
HDN void fp2_sqrt_from_norm_root(fp2 *out, const fp2 *a, const fp *sp) {
    fp t, w, x, c, one;
    fp_add(t, a->c0, *sp);
    fp_half(t, t);
    field_pow<FpSqr>(&w, &t, B381_TAB(qm3o4));
    fp_mul(x, w, t);
    fp_mul(c, w, x);                           // t^((Q-1)/2)
    fp_mul(w, w, a->c1);
    fp_half(w, w);
    fp_set_one(one);
    if (fp_eq(c, one)) { out->c0 = x; out->c1 = w; }
    else { out->c0 = w; fp_neg(out->c1, x); }
}
HDN bool fp2_sqrt(fp2 *out, const fp2 *a) {
    if (fp_is_zero(a->c1)) return fp2_sqrt_alg9(out, a);
    fp n, s;
    fp_sqr(n, a->c0); fp_sqr(s, a->c1);
    fp_add(n, n, s);
    if (!fp_sqrt(&s, &n)) return false;
    fp2_sqrt_from_norm_root(out, a, &s);
    return true;
}

// ---- field-generic helpers for the two curves ----------------------------------------------------------------------
struct G1Codec {
    typedef FpOut F;
    typedef fp T;
    typedef g1_affine_pod APOD;
    enum { BYTES = 48 };
    static HD void b_coeff(T &b) { fp_load_tab(b, B381_TAB(b_coeff)); }
    static HD bool sqrt(T *o, const T *a) { return fp_sqrt(o, a); }
    static HD bool is_square(const T &a) { return fp_is_square(a); }
    static HD int cmp(const T &a, const T &b) { return fp_cmp(a, b); }
    static HD bool in_subgroup(const T &x, const T &y);
    static HD void x_from_bytes(T &x, const uint8_t *c) { fp raw; fp_raw_from_be48(raw, c); fp_from_raw(x, raw); }
    static HD void x_to_bytes(uint8_t *c, const T &x) { fp raw; fp_to_raw(raw, x); fp_raw_to_be48(c, raw); }
    static HD void load(T &x, T &y, const APOD *p) { fp_load_u64(x, p->x); fp_load_u64(y, p->y); }
    static HD void store(APOD *p, const T &x, const T &y, bool inf) {
        fp_store_u64(p->x, x); fp_store_u64(p->y, y); p->inf = inf ? 1 : 0;
        for (int i = 0; i < 7; i++) p->pad[i] = 0;
    }
};
struct G2Codec {
    typedef Fp2Out F;
    typedef fp2 T;
    typedef g2_affine_pod APOD;
    enum { BYTES = 96 };
    static HD void b_coeff(T &b) { fp_load_tab(b.c0, B381_TAB(b_coeff)); b.c1 = b.c0; }   // 4(1 + u), g2.go:32
    static HD bool sqrt(T *o, const T *a) { return fp2_sqrt(o, a); }
    static HD bool is_square(const T &a) { fp n, m; fp_sqr(n, a.c0); fp_sqr(m, a.c1); fp_add(n, n, m); return fp_is_square(n); }   // through the norm
    static HD int cmp(const T &a, const T &b) { return fp2_cmp(a, b); }
    static HD bool in_subgroup(const T &x, const T &y);
    // x.c1 comes first on the wire (g2.go:250-257, 275-278)
    static HD void x_from_bytes(T &x, const uint8_t *c) {
        fp raw;
        fp_raw_from_be48(raw, c); fp_from_raw(x.c1, raw);
        fp_raw_from_be48(raw, c + 48); fp_from_raw(x.c0, raw);
    }
    static HD void x_to_bytes(uint8_t *c, const T &x) {
        fp raw;
        fp_to_raw(raw, x.c1); fp_raw_to_be48(c, raw);
        fp_to_raw(raw, x.c0); fp_raw_to_be48(c + 48, raw);
    }
    static HD void load(T &x, T &y, const APOD *p) {
        fp_load_u64(x.c0, p->x); fp_load_u64(x.c1, p->x + 6); fp_load_u64(y.c0, p->y); fp_load_u64(y.c1, p->y + 6);
    }
    static HD void store(APOD *p, const T &x, const T &y, bool inf) {
        fp_store_u64(p->x, x.c0); fp_store_u64(p->x + 6, x.c1); fp_store_u64(p->y, y.c0); fp_store_u64(p->y + 6, y.c1);
        p->inf = inf ? 1 : 0;
        for (int i = 0; i < 7; i++) p->pad[i] = 0;
    }
};

// ---- double-and-add ladders -------------------------------------------------------------------------------------------
// The ladders run in Jacobian coordinates with the formulas the reference itself uses -- doubling 2M + 5S
// (G1Projective.Double, g1.go:343-397: dbl-2009-l) and mixed addition 7M + 4S (AddAffine, g1.go:485-559: madd-2007-bl):
// a ladder is mostly doublings, and the XYZZ doubling of curve.cuh costs 6M + 3S (24 instead of 16 Fq multiplications
// over Fq2).  The result is handed back in XYZZ form (ZZ = Z^2, ZZZ = Z^3) for the callers that go on accumulating.
// (jac_pt, jac_dbl, jac_madd, jac_to_xyzz live in curve.cuh)
// acc = k * (x, y) for a finite affine point, k as `nlimbs` plain u32 limbs: MSB-first double-and-add, the loop of
// G1Affine.Mul (g1.go:59-77)
template <class F> HDN void point_mul(xyzz<F> *acc, const typename F::T *x, const typename F::T *y, const uint32_t *k, int nlimbs) {
    jac_pt<F> a;
    F::set_one(a.x); F::set_one(a.y); F::set_zero(a.z);
#pragma unroll 1
    for (int i = nlimbs * 32 - 1; i >= 0; i--) {
        jac_dbl(a);
        if ((k[i >> 5] >> (i & 31)) & 1) jac_madd(a, *x, *y);
    }
    jac_to_xyzz(*acc, a);
}
// acc = (pos - neg) * (x, y): MSB-first ladder over a non-adjacent form given as two bit masks (constants only)
template <class F> HDN void point_mul_naf(xyzz<F> *acc, const typename F::T *x, const typename F::T *y, const uint32_t *pos,
                                          const uint32_t *neg, int nlimbs) {
    jac_pt<F> a;
    F::set_one(a.x); F::set_one(a.y); F::set_zero(a.z);
    typename F::T ny;
    F::neg(ny, *y);
#pragma unroll 1
    for (int i = nlimbs * 32 - 1; i >= 0; i--) {
        jac_dbl(a);
        if ((pos[i >> 5] >> (i & 31)) & 1) jac_madd(a, *x, *y);
        else if ((neg[i >> 5] >> (i & 31)) & 1) jac_madd(a, *x, ny);
    }
    jac_to_xyzz(*acc, a);
}
// affine coordinates of a finite XYZZ point (ToAffine, g1.go:322-340: same canonical values)
template <class F> HD void xyzz_to_affine(typename F::T &x, typename F::T &y, const xyzz<F> &p) {
    typename F::T zi, t;
    F::inv(zi, p.zzz);                 // 1/ZZZ
    F::mul(y, p.y, zi);                // y = Y/ZZZ
    F::mul(t, zi, p.zz);               // 1/Z = ZZ/ZZZ
    F::sqr(t, t);                      // 1/ZZ
    F::mul(x, p.x, t);
}

// ---- membership in the r-torsion ---------------------------------------------------------------------------------------
// The reference tests [r]P == O with a 255-bit double-and-add (IsInCorrectSubgroupAssumingOnCurve, g1.go:137-141,
// g2.go:293-295).  For points ON THE CURVE the endomorphism criteria of M. Scott, "A note on group membership tests
// for G1, G2 and GT on BLS pairing-friendly curves" (eprint 2021/1130) decide the same predicate with a 128-bit
// (G1) / 64-bit (G2) ladder:   G1:  (beta x, y) == -[x^2] P        G2:  psi(P) == [x] P      (x = -0xd201000000010000)
// B381_SUBGROUP_LADDER selects the reference's full ladder instead (tests compare both with the oracle, including
// points of the cofactor subgroups).
// is the finite XYZZ point `a` equal to the affine point (x, y)?
template <class F> HD bool xyzz_eq_affine(const xyzz<F> &a, const typename F::T &x, const typename F::T &y) {
    if (xyzz_is_inf(a)) return false;
    typename F::T t;
    F::mul(t, x, a.zz);
    F::sub(t, t, a.x);
    if (!F::is_zero(t)) return false;
    F::mul(t, y, a.zzz);
    F::sub(t, t, a.y);
    return F::is_zero(t);
}
HD bool g1_in_subgroup(const fp &x, const fp &y) {
#ifdef B381_SUBGROUP_LADDER
    xyzz<FpOut> acc;
    point_mul<FpOut>(&acc, &x, &y, B381_TAB(r_order), 8);
    return xyzz_is_inf(acc);
#else
    xyzz<FpOut> acc;
    point_mul<FpOut>(&acc, &x, &y, B381_TAB(bls_x2), 4);        // [x^2] P
    fp bx, ny, beta;
    fp_load_tab(beta, B381_TAB(beta));
    fp_mul(bx, x, beta);
    fp_neg(ny, y);
    return xyzz_eq_affine<FpOut>(acc, bx, ny);                    // == -(beta x, y)
#endif
}
// psi(x, y) = (PSI_CX conj(x), PSI_CY conj(y)), PSI_CX = (0, k): the reference's psi (hash.go:341-366)
HD void g2_psi(fp2 &ox, fp2 &oy, const fp2 &x, const fp2 &y) {
    fp k;
    fp_load_tab(k, B381_TAB(psi_cx_c1));
    fp2 cy, t;
    // (0 + k u)(x0 - x1 u) = k x1 + k x0 u
    fp_mul(t.c0, x.c1, k);
    fp_mul(t.c1, x.c0, k);
    ox = t;
    fp_load_tab(cy.c0, B381_TAB(psi_cy)); fp_load_tab(cy.c1, B381_TAB(psi_cy) + 12);
    fp2_conj(t, y);
    fp2_mul(&oy, &t, &cy);
}
HD bool g2_in_subgroup(const fp2 &x, const fp2 &y) {
#ifdef B381_SUBGROUP_LADDER
    xyzz<Fp2Out> acc;
    point_mul<Fp2Out>(&acc, &x, &y, B381_TAB(r_order), 8);
    return xyzz_is_inf(acc);
#else
    xyzz<Fp2Out> acc;
    point_mul<Fp2Out>(&acc, &x, &y, B381_TAB(bls_x), 2);         // [|x|] P = -[x] P
    fp2 px, py;
    g2_psi(px, py, x, y);
    fp2_neg(py, py);
    return xyzz_eq_affine<Fp2Out>(acc, px, py);                   // [|x|] P == -psi(P)
#endif
}

HD bool G1Codec::in_subgroup(const fp &x, const fp &y) { return g1_in_subgroup(x, y); }
HD bool G2Codec::in_subgroup(const fp2 &x, const fp2 &y) { return g2_in_subgroup(x, y); }

// ---- decompression ---------------------------------------------------------------------------------------------------
// DecompressG1[Unchecked] / DecompressG2[Unchecked]; the affine point is written for status 0 and 4 (the reference
// returns no point with an error; callers must look at the status)
template <class C> HD int decompress_one(typename C::APOD *out, const uint8_t *in, bool check_subgroup) {
    typedef typename C::T T;
    uint8_t c[C::BYTES];
    for (int i = 0; i < C::BYTES; i++) c[i] = in[i];
    T x, y, ny, t;
    C::F::set_zero(x); C::F::set_one(y);
    if ((c[0] & 0x80) == 0) { C::store(out, x, y, true); return CODEC_ERR_MODE; }
    if (c[0] & 0x40) {
        c[0] &= 0x3f;
        uint8_t o = 0;
        for (int i = 0; i < C::BYTES; i++) o |= c[i];
        C::store(out, x, y, true);             // G1AffineZero = (0, 1, infinity), g1.go:269
        return o ? CODEC_ERR_INFINITY : CODEC_OK;
    }
    bool greatest = (c[0] & 0x20) != 0;
    c[0] &= 0x1f;
    C::x_from_bytes(x, c);
    // y^2 = x^3 + b
    C::F::sqr(t, x);
    C::F::mul(t, t, x);
    C::b_coeff(ny);
    C::F::add(t, t, ny);
    if (!C::sqrt(&y, &t)) { C::F::set_zero(x); C::F::set_one(y); C::store(out, x, y, true); return CODEC_ERR_NOT_ON_CURVE; }
    C::F::neg(ny, y);
    if (!((C::cmp(y, ny) < 0) != greatest)) y = ny;     // g1.go:126-129
    C::store(out, x, y, false);
    if (check_subgroup && !C::in_subgroup(x, y)) return CODEC_ERR_SUBGROUP;
    return CODEC_OK;
}
// CompressG1 / CompressG2
template <class C> HD void compress_one(uint8_t *out, const typename C::APOD *in) {
    typedef typename C::T T;
    for (int i = 0; i < C::BYTES; i++) out[i] = 0;
    if (in->inf) { out[0] = 0xc0; return; }
    T x, y, ny;
    C::load(x, y, in);
    C::x_to_bytes(out, x);
    C::F::neg(ny, y);
    if (C::cmp(y, ny) > 0) out[0] |= 0x20;
    out[0] |= 0x80;
}
// ---- scalar multiplication of r-torsion points through the endomorphisms ----------------------------------------------------
// For P in G1 / G2 (NOT for other curve points) the 255-bit ladder of MulFR (g1.go:80-90, g2.go:92-102) is replaced by
//   G2:  k = d0 + d1 X + d2 X^2 + d3 X^3 (X = |x|),  [k]P = [d0]P - [d1]psi(P) + [d2]psi^2(P) - [d3]psi^3(P)   (psi = [x] = [-X])
//   G1:  k = k0 + k1 X^2,  [k]P = [k0]P + [k1](beta x, -y)                         ((beta x, y) = -[X^2]P, the subgroup test's identity)
// with joint fixed windows over the bases: 64 doublings + 64 table additions (G2), 128 + 64 (G1), instead of 255 + 255 (a
// warp pays for an addition at every bit some lane has set, i.e. at every bit).
// The affine result is the same canonical pair of field elements.  Used for Sign / PrivToPub, whose bases are the generator
// or a hash to the curve; b381_g{1,2}_mul_batch keeps the plain ladder because it must also serve points outside the subgroup.
// base-X digits of a 256-bit scalar by restoring division (X has its top bit set); d[4] = 0 whenever k < X^4 (r < X^4)
HD void scalar_digits_base_x(uint64_t d[5], const uint64_t *k) {
    const uint64_t X = 0xd201000000010000ull;
    uint64_t n[4] = {k[0], k[1], k[2], k[3]};
#pragma unroll 1
    for (int j = 0; j < 4; j++) {
        uint64_t rem = 0;
#pragma unroll 1
        for (int b = 255; b >= 0; b--) {
            uint64_t carry = rem >> 63;
            rem = (rem << 1) | ((n[b >> 6] >> (b & 63)) & 1u);
            uint64_t ge = (carry | (uint64_t)(rem >= X)) & 1u;
            if (ge) rem -= X;
            n[b >> 6] = (n[b >> 6] & ~(1ull << (b & 63))) | (ge << (b & 63));
        }
        d[j] = rem;
    }
    d[4] = n[0] | n[1] | n[2] | n[3];
}
// affine coordinates of n finite XYZZ points with ONE inversion (Montgomery's trick on the ZZZ); n <= 16
template <class F> HDN void xyzz_batch_to_affine(typename F::T *ox, typename F::T *oy, const xyzz<F> *p, int n) {
    typename F::T pre[16], acc, inv, t;
    acc = p[0].zzz;
    pre[0] = acc;
#pragma unroll 1
    for (int i = 1; i < n; i++) { F::mul(acc, acc, p[i].zzz); pre[i] = acc; }
    F::inv(inv, acc);
#pragma unroll 1
    for (int i = n - 1; i >= 0; i--) {
        typename F::T zi;
        if (i) { F::mul(zi, inv, pre[i - 1]); F::mul(inv, inv, p[i].zzz); }
        else zi = inv;
        F::mul(oy[i], p[i].y, zi);         // y = Y / ZZZ
        F::mul(t, zi, p[i].zz);            // 1 / Z
        F::sqr(t, t);
        F::mul(ox[i], p[i].x, t);          // x = X / ZZ
    }
}
// acc = k * (x, y) for ANY finite curve point and a per-lane scalar (8 limbs): fixed four-bit windows over the affine table
// P .. 15 P (one shared inversion) -- 256 doublings + 64 table additions on a schedule common to the warp, where the bitwise
// ladder pays an addition at every bit some lane has set.  A point of order <= 15 makes a table entry infinite; such inputs
// take the bitwise ladder.  Same affine result as G1Affine.Mul / G2Affine.Mul (g1.go:59-90).
template <class F> HDN void point_mul_w4(xyzz<F> *acc, const typename F::T *x, const typename F::T *y, const uint32_t *k) {
    typename F::T tx[16], ty[16];
    {
        xyzz<F> s[15];
        s[0].x = *x; s[0].y = *y; F::set_one(s[0].zz); F::set_one(s[0].zzz);
        xyzz_dbl_affine(s[1], *x, *y);
        bool finite = !xyzz_is_inf(s[1]);
#pragma unroll 1
        for (int m = 2; m < 15; m++) { s[m] = s[m - 1]; xyzz_madd(s[m], *x, *y); finite = finite && !xyzz_is_inf(s[m]); }
        if (!finite) { point_mul<F>(acc, x, y, k, 8); return; }
        xyzz_batch_to_affine<F>(tx + 1, ty + 1, s, 15);
    }
    jac_pt<F> a;
    F::set_one(a.x); F::set_one(a.y); F::set_zero(a.z);
#pragma unroll 1
    for (int j = 252; j >= 0; j -= 4) {
        jac_dbl(a); jac_dbl(a); jac_dbl(a); jac_dbl(a);
        int m = (int)((k[j >> 5] >> (j & 31)) & 15u);
        if (m) jac_madd(a, tx[m], ty[m]);
    }
    jac_to_xyzz(*acc, a);
}
// A warp executes an addition whenever ANY of its lanes needs one, so per-lane sparse digits (NAF) buy nothing under SIMT;
// what pays is a fixed schedule: every lane adds a table entry at the same steps and only the INDEX depends on its scalar.
// G2: joint one-bit window over the four bases, T[m] = sum of the B_i with bit i of m set (15 affine entries, built with
// 11 mixed additions and one shared inversion), then 64 x (doubling + one mixed addition).
HDN void torsion_mul(xyzz<Fp2Out> *acc, const fp2 *x, const fp2 *y, const uint64_t *k) {
    uint64_t d[5];
    scalar_digits_base_x(d, k);
    if (d[4]) {                                            // k >= X^4 > r: not a canonical scalar, take the plain ladder
        uint32_t kk[8];
        for (int i = 0; i < 4; i++) { kk[2 * i] = (uint32_t)k[i]; kk[2 * i + 1] = (uint32_t)(k[i] >> 32); }
        point_mul<Fp2Out>(acc, x, y, kk, 8);
        return;
    }
    fp2 tx[16], ty[16];                                    // entry m at index m; index 0 unused
    tx[1] = *x; ty[1] = *y;
    for (int i = 1; i < 4; i++) {                          // B_i = -psi(B_{i-1}) = (-1)^i psi^i(P) at index 2^i
        g2_psi(tx[1 << i], ty[1 << i], tx[1 << (i - 1)], ty[1 << (i - 1)]);
        fp2_neg(ty[1 << i], ty[1 << i]);
    }
    {
        xyzz<Fp2Out> s[11];
        const uint8_t comp[11] = {3, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15};
        fp2 sx[11], sy[11];
#pragma unroll 1
        for (int j = 0; j < 11; j++) {                     // T[m] = T[m without its top bit] + B_top: the smaller entry is already
            int m = comp[j], top = m >= 8 ? 8 : (m >= 4 ? 4 : 2), lowm = m - top;   // affine (a base) or an earlier s[]
            if ((lowm & (lowm - 1)) == 0) { s[j].x = tx[lowm]; s[j].y = ty[lowm]; fp2_set_one(s[j].zz); fp2_set_one(s[j].zzz); }
            else {
                int q = 0;
                while (comp[q] != lowm) q++;
                s[j] = s[q];
            }
            xyzz_madd(s[j], tx[top], ty[top]);
        }
        xyzz_batch_to_affine<Fp2Out>(sx, sy, s, 11);
        for (int j = 0; j < 11; j++) { tx[comp[j]] = sx[j]; ty[comp[j]] = sy[j]; }
    }
    jac_pt<Fp2Out> a;
    fp2_set_one(a.x); fp2_set_one(a.y); fp2_set_zero(a.z);
#pragma unroll 1
    for (int j = 63; j >= 0; j--) {
        jac_dbl(a);
        int m = (int)(((d[0] >> j) & 1) | (((d[1] >> j) & 1) << 1) | (((d[2] >> j) & 1) << 2) | (((d[3] >> j) & 1) << 3));
        if (m) jac_madd(a, tx[m], ty[m]);
    }
    jac_to_xyzz(*acc, a);
}
// G1: k = k0 + k1 X^2 over P and Q = (beta x, -y) = [X^2]P, joint two-bit windows: T[a + 4 b] = a P + b Q (15 affine
// entries, 14 mixed additions / doublings and one shared inversion), then 64 x (two doublings + one mixed addition).
HDN void torsion_mul(xyzz<FpOut> *acc, const fp *x, const fp *y, const uint64_t *k) {
    uint64_t d[5];
    scalar_digits_base_x(d, k);
    if (d[4]) {
        uint32_t kk[8];
        for (int i = 0; i < 4; i++) { kk[2 * i] = (uint32_t)k[i]; kk[2 * i + 1] = (uint32_t)(k[i] >> 32); }
        point_mul<FpOut>(acc, x, y, kk, 8);
        return;
    }
    const uint64_t X = 0xd201000000010000ull;
    uint32_t kw[2][4];
    for (int i = 0; i < 2; i++) {                          // k_i = d[2i] + d[2i+1] X < X^2 < 2^128, by 32-bit partial products
        uint32_t *n = kw[i], m[2] = {(uint32_t)d[2 * i + 1], (uint32_t)(d[2 * i + 1] >> 32)}, xs[2] = {(uint32_t)X, (uint32_t)(X >> 32)};
        n[0] = n[1] = n[2] = n[3] = 0;
        for (int u = 0; u < 2; u++) {
            uint64_t c = 0;
            for (int v = 0; v < 2; v++) {
                uint64_t t = (uint64_t)m[u] * xs[v] + n[u + v] + c;
                n[u + v] = (uint32_t)t;
                c = t >> 32;
            }
            n[u + 2] = (uint32_t)c;
        }
        uint64_t c = 0, add[4] = {(uint32_t)d[2 * i], d[2 * i] >> 32, 0, 0};
        for (int u = 0; u < 4; u++) {
            uint64_t t = (uint64_t)n[u] + add[u] + c;
            n[u] = (uint32_t)t;
            c = t >> 32;
        }
    }
    fp tx[16], ty[16], qx, qy, beta;
    fp_load_tab(beta, B381_TAB(beta));
    fp_mul(qx, *x, beta);
    fp_neg(qy, *y);
    {
        xyzz<FpOut> s[15];                                 // s[m - 1] = T[m]
        s[0].x = *x; s[0].y = *y; fp_set_one(s[0].zz); fp_set_one(s[0].zzz);
        xyzz_dbl_affine(s[1], *x, *y);
        s[2] = s[1]; xyzz_madd(s[2], *x, *y);
        s[3].x = qx; s[3].y = qy; fp_set_one(s[3].zz); fp_set_one(s[3].zzz);
#pragma unroll 1
        for (int m = 5; m <= 15; m++) { s[m - 1] = s[m - 5]; xyzz_madd(s[m - 1], qx, qy); }   // a P + b Q = (a P + (b - 1) Q) + Q
        xyzz_batch_to_affine<FpOut>(tx + 1, ty + 1, s, 15);
    }
    jac_pt<FpOut> a;
    fp_set_one(a.x); fp_set_one(a.y); fp_set_zero(a.z);
#pragma unroll 1
    for (int j = 126; j >= 0; j -= 2) {
        jac_dbl(a); jac_dbl(a);
        int m = (int)(((kw[0][j >> 5] >> (j & 31)) & 3u) | (((kw[1][j >> 5] >> (j & 31)) & 3u) << 2));
        if (m) jac_madd(a, tx[m], ty[m]);
    }
    jac_to_xyzz(*acc, a);
}
// out = k * p as an affine point (MulFR + ToAffine); p at infinity or k = 0 give the canonical zero (0, 1, inf)
template <class C, bool TORSION> HD void mul_one(typename C::APOD *out, const typename C::APOD *p, const uint64_t *k) {
    typedef typename C::T T;
    T x, y;
    uint32_t kk[8];
    for (int i = 0; i < 4; i++) { kk[2 * i] = (uint32_t)k[i]; kk[2 * i + 1] = (uint32_t)(k[i] >> 32); }
    xyzz<typename C::F> acc;
    xyzz_set_inf(acc);
    if (!p->inf) {
        C::load(x, y, p);
        if (TORSION) torsion_mul(&acc, &x, &y, k);
        else point_mul_w4<typename C::F>(&acc, &x, &y, kk);
    }
    if (xyzz_is_inf(acc)) { C::F::set_zero(x); C::F::set_one(y); C::store(out, x, y, true); return; }
    xyzz_to_affine<typename C::F>(x, y, acc);
    C::store(out, x, y, false);
}

#if defined(__CUDACC__)
template <class C> __global__ void __launch_bounds__(64, CODEC_MIN_BLOCKS) k_decompress(const uint8_t *__restrict__ in, size_t n, int check_subgroup,
                                                                      typename C::APOD *__restrict__ out, uint8_t *__restrict__ status) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    status[i] = (uint8_t)decompress_one<C>(out + i, in + (size_t)C::BYTES * i, check_subgroup != 0);
}
template <class C> __global__ void __launch_bounds__(64) k_compress(const typename C::APOD *__restrict__ in, size_t n,
                                                                    uint8_t *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    compress_one<C>(out + (size_t)C::BYTES * i, in + i);
}
// out[i] = k[i] * p[i * p_stride]: p_stride 0 multiplies one base by every scalar (PrivToPub: the generator)
template <class C, bool TORSION> __global__ void __launch_bounds__(64, CODEC_MIN_BLOCKS) k_point_mul(const typename C::APOD *__restrict__ p, size_t p_stride,
                                                                     const uint64_t *__restrict__ k, size_t k_stride, size_t n,
                                                                     typename C::APOD *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    mul_one<C, TORSION>(out + i, p + i * p_stride, k + 4 * i * k_stride);
}
#endif

}  // namespace b381
