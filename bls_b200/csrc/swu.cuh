// swu.cuh -- HashG1 / HashG2 on the device (second half): the message points of g2pubs.Verify /
// Sign (HashG1, hash.go:320-331) and of g1pubs.Verify / Sign / VerifyAggregate (HashG2, hash.go:404-411):
//   hash_to_field  hp / hp2 (hash.go:41-113): SHA-256(0x01 || msg), then two (per coefficient) single-block SHA-256
//                  digests folded into a residue mod Q
//   simplified SWU onto the isogenous curve  optimizedSWUMapHelper (g1.go:614-714), OptimizedSWU2MapHelper (g2.go:933-1031)
//   the sum of the two mapped points (ordinary AddAffine, hash.go:311-318,391-399)
//   the 11- / 3-isogeny  iso11 (hash.go:185-203), iso3 (hash.go:282-303)
//   cofactor clearing  ClearH = [x + 1]P with x = 0xd201000000010000 (hash.go:305-309), clearH2 with psi (hash.go:341-389)
// One thread per message.  Every intermediate the reference normalises is a canonical affine point, so the formulas
// underneath (XYZZ accumulators, psi applied in XYZZ form, shared inversions) are free; roots are fixed by the
// reference's sign rule, not by which root a square-root algorithm happens to return.
#pragma once
#include "hash.cuh"

namespace b381 {

#if defined(__CUDACC__)
__device__ __constant__ uint32_t d_two_256[12] = {B381_TWO_256_LIMBS};
__device__ __constant__ uint32_t d_ellpa[12] = {B381_ELLPA_LIMBS};
__device__ __constant__ uint32_t d_ellpb[12] = {B381_ELLPB_LIMBS};
__device__ __constant__ uint32_t d_ell2pa[24] = {B381_ELL2PA_LIMBS};
__device__ __constant__ uint32_t d_ell2pb[24] = {B381_ELL2PB_LIMBS};
__device__ __constant__ uint32_t d_iso11[55 * 12] = B381_ISO11_INIT;
__device__ __constant__ uint32_t d_iso3[15 * 24] = B381_ISO3_INIT;
#endif
#if !defined(__CUDA_ARCH__)
static const uint32_t h_two_256[12] = {B381_TWO_256_LIMBS};
static const uint32_t h_ellpa[12] = {B381_ELLPA_LIMBS};
static const uint32_t h_ellpb[12] = {B381_ELLPB_LIMBS};
static const uint32_t h_ell2pa[24] = {B381_ELL2PA_LIMBS};
static const uint32_t h_ell2pb[24] = {B381_ELL2PB_LIMBS};
static const uint32_t h_iso11[55 * 12] = B381_ISO11_INIT;
static const uint32_t h_iso3[15 * 24] = B381_ISO3_INIT;
#endif

// SHA-256 of (prefix byte || msg[0..len)) for any length: the msgHash of hp / hp2 with the cipher-suite byte 0x01
// prepended by HashG1 / HashG2 (hash.go:318-331,404-411)
HD void sha256_prefixed(uint32_t dg[8], uint8_t prefix, const uint8_t *msg, size_t len) {
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    const uint32_t *K = B381_TAB(sha_k);
    size_t total = len + 1, nblk = (total + 9 + 63) / 64;
    uint64_t bits = (uint64_t)total * 8;
#define B381_ROR(x, n) (((x) >> (n)) | ((x) << (32 - (n))))
#pragma unroll 1
    for (size_t blk = 0; blk < nblk; blk++) {
        uint32_t w[64];
        for (int i = 0; i < 16; i++) {
            uint32_t v = 0;
            for (int b = 0; b < 4; b++) {
                size_t pos = blk * 64 + 4 * i + b;
                uint8_t byte;
                if (pos == 0) byte = prefix;
                else if (pos < total) byte = msg[pos - 1];
                else if (pos == total) byte = 0x80;
                else if (pos >= nblk * 64 - 8) byte = (uint8_t)(bits >> (8 * (nblk * 64 - 1 - pos)));
                else byte = 0;
                v = (v << 8) | byte;
            }
            w[i] = v;
        }
        for (int i = 16; i < 64; i++) {
            uint32_t s0 = B381_ROR(w[i - 15], 7) ^ B381_ROR(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = B381_ROR(w[i - 2], 17) ^ B381_ROR(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll 1
        for (int i = 0; i < 64; i++) {
            uint32_t S1 = B381_ROR(e, 6) ^ B381_ROR(e, 11) ^ B381_ROR(e, 25);
            uint32_t t1 = hh + S1 + ((e & f) ^ (~e & g)) + K[i] + w[i];
            uint32_t t2 = (B381_ROR(a, 2) ^ B381_ROR(a, 13) ^ B381_ROR(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
#undef B381_ROR
    for (int i = 0; i < 8; i++) dg[i] = h[i];
}
// one coefficient of hash_to_field: (SHA-256(prime || tag || 1) || SHA-256(prime || tag || 2)) as a 512-bit big-endian
// integer mod Q, prime = msgHash || ctr (hash.go:49-72: tag 0x01; hash.go:86-110: tag = i)
HD void hash_to_fp(fp &r, const uint32_t msg_hash[8], uint8_t ctr, uint8_t tag) {
    uint8_t buf[35];
    for (int i = 0; i < 8; i++) {
        buf[4 * i] = (uint8_t)(msg_hash[i] >> 24); buf[4 * i + 1] = (uint8_t)(msg_hash[i] >> 16);
        buf[4 * i + 2] = (uint8_t)(msg_hash[i] >> 8); buf[4 * i + 3] = (uint8_t)msg_hash[i];
    }
    buf[32] = ctr; buf[33] = tag;
    uint32_t dg[8];
    fp hi, lo, c;
    buf[34] = 1; sha256_short(dg, buf, 35); fp_from_digest(hi, dg);
    buf[34] = 2; sha256_short(dg, buf, 35); fp_from_digest(lo, dg);
    fp_load_tab(c, B381_TAB(two_256));
    fp_mul(hi, hi, c);
    fp_add(r, hi, lo);
}

// ---- the two suites ------------------------------------------------------------------------------------------------------
struct G1Swu : G1Codec {
    static HD void ell_a(T &a) { fp_load_tab(a, B381_TAB(ellpa)); }
    static HD void ell_b(T &b) { fp_load_tab(b, B381_TAB(ellpb)); }
    static HD void mul_nqr(T &r, const T &a) { fp_neg(r, a); }                      // xi = -1 (g1.go:630)
    static HD bool eq(const T &a, const T &b) { return fp_eq(a, b); }
    // signFQ (g1.go:621-626): -1 iff f > (Q-1)/2
    static HD int sign(const T &f) {
        fp raw, th;
        fp_to_raw(raw, f);
        fp_load_tab(th, B381_TAB(qm1o2));
        return fp_raw_cmp(raw, th) > 0 ? -1 : 1;
    }
    static HD void field_of_hash(T &t, const uint32_t mh[8], uint8_t ctr) { hash_to_fp(t, mh, ctr, 1); }
    static HD void iso_coeff(T &c, int idx) { fp_load_tab(c, B381_TAB(iso11) + 12 * idx); }
    static HD int iso_len(int m) { const int l[4] = {B381_ISO11_LENS}; return l[m]; }
    // ClearH (hash.go:305-309): [x]P + P
    static HD void clear_h(xyzz<F> &acc, const T &x, const T &y) {
        point_mul<F>(&acc, &x, &y, B381_TAB(bls_x), 2);
        xyzz_madd(acc, x, y);
    }
};
struct G2Swu : G2Codec {
    static HD void ell_a(T &a) { fp_load_tab(a.c0, B381_TAB(ell2pa)); fp_load_tab(a.c1, B381_TAB(ell2pa) + 12); }
    static HD void ell_b(T &b) { fp_load_tab(b.c0, B381_TAB(ell2pb)); fp_load_tab(b.c1, B381_TAB(ell2pb) + 12); }
    static HD void mul_nqr(T &r, const T &a) { fp2_mul_nr(r, a); }                  // fq2nqr = 1 + u (fq6.go:139-142)
    static HD bool eq(const T &a, const T &b) { return fp2_eq(a, b); }
    // signFQ2 (g2.go:920-934)
    static HD int sign(const T &f) {
        fp raw, th;
        fp_load_tab(th, B381_TAB(qm1o2));
        fp_to_raw(raw, f.c1);
        if (fp_raw_cmp(raw, th) > 0) return -1;
        if (!fp_is_zero(raw)) return 1;
        fp_to_raw(raw, f.c0);
        return fp_raw_cmp(raw, th) > 0 ? -1 : 1;
    }
    static HD void field_of_hash(T &t, const uint32_t mh[8], uint8_t ctr) {
        hash_to_fp(t.c0, mh, ctr, 1);
        hash_to_fp(t.c1, mh, ctr, 2);
    }
    static HD void iso_coeff(T &c, int idx) { fp_load_tab(c.c0, B381_TAB(iso3) + 24 * idx); fp_load_tab(c.c1, B381_TAB(iso3) + 24 * idx + 12); }
    static HD int iso_len(int m) { const int l[4] = {B381_ISO3_LENS}; return l[m]; }
    // clearH2 (hash.go:368-389)
    static HD void clear_h(xyzz<F> &acc, const T &x, const T &y) { g2_clear_h2(acc, x, y); }
};

// simplified SWU onto y^2 = x^3 + A x + B (optimizedSWUMapHelper g1.go:628-714, OptimizedSWU2MapHelper g2.go:933-1031), split
// around its one inversion so that the two maps of a hash share a single inversion (Montgomery's trick):
//   swu_prepare: u = xi t^2, common = u^2 + u, den = A common (xi A when common = 0)    -- den is never zero
//   swu_finish : x0 = -B (common + 1) / den (B / den when common = 0), the candidate test and the root
template <class S> struct swu_state { typename S::T t, u, common, den; };
template <class S> HDN void swu_prepare(swu_state<S> *st, const typename S::T *tp) {
    typedef typename S::T T;
    typedef typename S::F F;
    T A, t2;
    S::ell_a(A);
    st->t = *tp;
    F::sqr(t2, st->t);
    S::mul_nqr(st->u, t2);                 // xi t^2
    F::sqr(st->common, st->u);             // xi^2 t^4
    F::add(st->common, st->common, st->u);
    if (F::is_zero(st->common)) S::mul_nqr(st->den, A);
    else F::mul(st->den, A, st->common);
}
template <class S> HDN void swu_finish(typename S::T *ox, typename S::T *oy, const swu_state<S> *st, const typename S::T *inv_den) {
    typedef typename S::T T;
    typedef typename S::F F;
    T A, B, x0, gx, y, v, one;
    S::ell_a(A); S::ell_b(B);
    if (F::is_zero(st->common)) {          // x0 = B / (xi A)
        F::mul(x0, B, *inv_den);
    } else {                               // x0 = -B (common + 1) / (A common)
        F::set_one(one);
        F::add(v, st->common, one);
        F::neg(gx, B);
        F::mul(x0, gx, v);
        F::mul(x0, x0, *inv_den);
    }
    F::sqr(gx, x0);
    F::mul(gx, gx, x0);
    F::mul(v, A, x0);
    F::add(gx, gx, v);
    F::add(gx, gx, B);                     // g(x0)
    // the reference takes sqrt(g(x0)) and falls back to x1 when it fails; which of the two happens is decided here by a
    // Jacobi symbol, so that every lane of the warp runs ONE square root, on its own candidate
    if (!S::is_square(gx)) {               // x1 = xi t^2 x0, g(x1) = xi^3 t^6 g(x0)
        F::mul(x0, st->u, x0);
        F::sqr(gx, x0);
        F::mul(gx, gx, x0);
        F::mul(v, A, x0);
        F::add(gx, gx, v);
        F::add(gx, gx, B);
    }
    S::sqrt(&y, &gx);
    if (S::sign(st->t) != S::sign(y)) F::neg(y, y);
    *ox = x0; *oy = y;
}
// The rational map (xNum/xDen, y yNum/yDen) (iso11 hash.go:185-203, iso3 hash.go:282-303) applied to a finite XYZZ point
// (x = X/ZZ, y = Y/ZZZ) without normalising it first: each polynomial p of degree d is evaluated in homogeneous form,
//   P = ZZ^d p(X/ZZ) = sum c_i X^i ZZ^(d-i)   (Horner: acc = acc X + c_(d-k) ZZ^k, the powers of ZZ shared by the four),
// so that x' = P0 / (P1 ZZ^(d0-d1)), y' = Y P2 / (ZZZ P3 ZZ^(d2-d3)) and ONE inversion yields the affine image -- the same
// field elements as ToAffine followed by the affine map.
template <class S> HDN void iso_map_xyzz(typename S::T *ox, typename S::T *oy, const xyzz<typename S::F> *p) {
    typedef typename S::T T;
    typedef typename S::F F;
    T acc[4], pw, c;
    int deg[4], base[4], maxd = 0, b0 = 0;
    for (int m = 0; m < 4; m++) {
        deg[m] = S::iso_len(m) - 1;
        base[m] = b0;
        b0 += deg[m] + 1;
        if (deg[m] > maxd) maxd = deg[m];
        S::iso_coeff(acc[m], base[m] + deg[m]);
    }
    pw = p->zz;
#pragma unroll 1
    for (int k = 1; k <= maxd; k++) {
#pragma unroll 1
        for (int m = 0; m < 4; m++) {
            if (k > deg[m]) continue;
            F::mul(acc[m], acc[m], p->x);
            S::iso_coeff(c, base[m] + deg[m] - k);
            F::mul(c, c, pw);
            F::add(acc[m], acc[m], c);
        }
        if (k < maxd) F::mul(pw, pw, p->zz);
    }
    // balance the powers of ZZ between numerators and denominators (d0 - d1 = 1, d2 = d3 for both curves)
    for (int k = deg[1]; k < deg[0]; k++) F::mul(acc[1], acc[1], p->zz);
    for (int k = deg[0]; k < deg[1]; k++) F::mul(acc[0], acc[0], p->zz);
    for (int k = deg[3]; k < deg[2]; k++) F::mul(acc[3], acc[3], p->zz);
    for (int k = deg[2]; k < deg[3]; k++) F::mul(acc[2], acc[2], p->zz);
    F::mul(acc[2], acc[2], p->y);
    F::mul(acc[3], acc[3], p->zzz);
    // one inversion for both denominators
    T d, di;
    F::mul(d, acc[1], acc[3]);
    F::inv(di, d);
    F::mul(c, di, acc[3]);                 // 1 / xDen
    F::mul(*ox, acc[0], c);
    F::mul(c, di, acc[1]);                 // 1 / yDen
    F::mul(*oy, acc[2], c);
}
// HashG1(msg) / HashG2(msg) as an affine point: three inversions (the two SWU maps together, the isogeny, the result)
template <class S> HD void hash_to_curve_one(typename S::APOD *out, const uint8_t *msg, size_t len) {
    typedef typename S::T T;
    typedef typename S::F F;
    uint32_t mh[8];
    sha256_prefixed(mh, 0x01, msg, len);
    T t, x1, y1, x2, y2, inv, i1, i2;
    swu_state<S> s1, s2;
    S::field_of_hash(t, mh, 0);
    swu_prepare<S>(&s1, &t);
    S::field_of_hash(t, mh, 1);
    swu_prepare<S>(&s2, &t);
    F::mul(inv, s1.den, s2.den);
    F::inv(inv, inv);
    F::mul(i1, inv, s2.den);
    F::mul(i2, inv, s1.den);
    swu_finish<S>(&x1, &y1, &s1, &i1);
    swu_finish<S>(&x2, &y2, &s2, &i2);
    xyzz<F> acc;
    acc.x = x1; acc.y = y1; F::set_one(acc.zz); F::set_one(acc.zzz);
    xyzz_madd(acc, x2, y2);                // Pp.ToProjective().AddAffine(Pp2)
    if (xyzz_is_inf(acc)) { F::set_zero(x1); F::set_one(y1); S::store(out, x1, y1, true); return; }
    iso_map_xyzz<S>(&x1, &y1, &acc);
    S::clear_h(acc, x1, y1);
    if (xyzz_is_inf(acc)) { F::set_zero(x1); F::set_one(y1); S::store(out, x1, y1, true); return; }
    xyzz_to_affine<F>(x1, y1, acc);
    S::store(out, x1, y1, false);
}

#if defined(__CUDACC__)
// out[i] = Hash(msgs[off[i] .. off[i+1])): variable-length messages packed back to back
template <class S> __global__ void __launch_bounds__(64, CODEC_MIN_BLOCKS) k_hash_to_curve(const uint8_t *__restrict__ msgs, const uint64_t *__restrict__ off,
                                                                                          size_t n, typename S::APOD *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    hash_to_curve_one<S>(out + i, msgs + off[i], (size_t)(off[i + 1] - off[i]));
}
// The two Miller pairs of g2pubs.Verify (g2pubs/bls.go:159-162 -> CompareTwoPairings(sig, G2One, HashG1(m), pub)):
// (P, Q)[2i] = (sig[i], G2One), (P, Q)[2i+1] = (-H[i], pub[i]); same validity rule as k_verify_pairs
__global__ void __launch_bounds__(128) k_verify_pairs_g2pubs(const g2_affine_pod *__restrict__ pub, const uint8_t *__restrict__ pub_status,
                                                             const g1_affine_pod *__restrict__ sig, const uint8_t *__restrict__ sig_status,
                                                             const g1_affine_pod *__restrict__ H, size_t n, g1_affine_pod *__restrict__ P,
                                                             g2_affine_pod *__restrict__ Q, uint32_t *__restrict__ group_off,
                                                             uint8_t *__restrict__ valid) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) group_off[0] = 0;
    if (i >= n) return;
    group_off[i + 1] = (uint32_t)(2 * i + 2);
    valid[i] = (pub_status[i] == 0 && sig_status[i] == 0 && !pub[i].inf && !sig[i].inf) ? 1 : 0;
    g1_affine_pod h = H[i];
    fp y;
    fp_load_u64(y, h.y);
    fp_neg(y, y);
    fp_store_u64(h.y, y);
    P[2 * i] = sig[i];
    P[2 * i + 1] = h;
    g2_affine_pod one;
    const uint32_t gx[24] = {B381_G2_GEN_X_LIMBS}, gy[24] = {B381_G2_GEN_Y_LIMBS};
    fp t;
    fp_load_tab(t, gx); fp_store_u64(one.x, t); fp_load_tab(t, gx + 12); fp_store_u64(one.x + 6, t);
    fp_load_tab(t, gy); fp_store_u64(one.y, t); fp_load_tab(t, gy + 12); fp_store_u64(one.y + 6, t);
    one.inf = 0;
    for (int k = 0; k < 7; k++) one.pad[k] = 0;
    Q[2 * i] = one;
    Q[2 * i + 1] = pub[i];
}
// Random-linear-combination batch check: closes the product  prod_i e(r_i pk_i, H_i) * e(-G1One, sum_i r_i sig_i)  with its
// last pair (P, Q)[n] = (-G1One, S), S = the normalised Jacobian sum; *all_valid = 0 when any key or signature is the point
// at infinity or failed to deserialise (status arrays may be null for already-decoded inputs).
__global__ void __launch_bounds__(128) k_rlc_close(const g2_jac_pod *__restrict__ S, g1_affine_pod *__restrict__ P_last,
                                                   g2_affine_pod *__restrict__ Q_last, uint32_t *__restrict__ group_off, uint32_t n) {
    if (blockIdx.x || threadIdx.x) return;
    group_off[0] = 0; group_off[1] = n + 1;
    g1_affine_pod one;
    const uint32_t gx[12] = {B381_G1_GEN_X_LIMBS}, gy[12] = {B381_G1_GEN_Y_LIMBS};
    fp t;
    fp_load_tab(t, gx); fp_store_u64(one.x, t);
    fp_load_tab(t, gy); fp_neg(t, t); fp_store_u64(one.y, t);
    one.inf = 0;
    for (int k = 0; k < 7; k++) one.pad[k] = 0;
    *P_last = one;
    g2_affine_pod q;
    uint64_t zor = 0;
    for (int k = 0; k < 12; k++) { q.x[k] = S->x[k]; q.y[k] = S->y[k]; zor |= S->z[k]; }   // the sum kernels return z = 1 or 0
    q.inf = zor ? 0 : 1;
    for (int k = 0; k < 7; k++) q.pad[k] = 0;
    *Q_last = q;
}
// a weight of zero would drop its triple from the combination (r pk = infinity contributes the factor 1), and a weight of more
// than `bits` bits would be truncated by the short ladder / the window count: both make the whole check false
__global__ void k_rlc_valid(const g1_affine_pod *__restrict__ pub, const g2_affine_pod *__restrict__ sig, const uint8_t *__restrict__ pub_status,
                            const uint8_t *__restrict__ sig_status, const uint64_t *__restrict__ r, int bits, size_t n,
                            uint32_t *__restrict__ any_bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool bad = pub[i].inf || sig[i].inf || (pub_status && pub_status[i]) || (sig_status && sig_status[i]);
    uint64_t any = 0, over = 0;
    for (int l = 0; l < 4; l++) {
        uint64_t w = r[4 * i + l];
        any |= w;
        int lo = 64 * l;
        if (bits <= lo) over |= w; else if (bits < lo + 64) over |= w >> (bits - lo);
    }
    bad = bad || any == 0 || over != 0;
    if (bad) atomicOr(any_bad, 1u);
}
__global__ void k_rlc_finish(uint8_t *__restrict__ ok, const uint32_t *__restrict__ any_bad) {
    if (blockIdx.x == 0 && threadIdx.x == 0) ok[0] = (ok[0] && !*any_bad) ? 1 : 0;
}
// ok[i] &= (status[i] == 0 && !pt[i].inf): a deserialisation failure (or an infinite signature) makes the check false
__global__ void k_and_status_g2(uint8_t *__restrict__ ok, const uint8_t *__restrict__ status, const g2_affine_pod *__restrict__ pt, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ok[i] = (ok[i] && status[i] == 0 && !pt[i].inf) ? 1 : 0;
}
#endif

}  // namespace b381
