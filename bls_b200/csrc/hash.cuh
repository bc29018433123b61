// hash.cuh -- hashing to G2 on the device: HashG2WithDomain (g2.go:1041-1085), the message-point
// computation of VerifyWithDomain / SignWithDomain / VerifyAggregate*WithDomain (g1pubs/bls.go:138-141,171-174,294-311).
// It is ~40 % of the reference's VerifyWithDomain (22.5 k of ~54 k Fq multiplications) and the dominant host cost once
// the pairings run on the GPU.  One thread per message: two single-block SHA-256 (41-byte inputs), try-and-increment
// on x, FQ2.Sqrt, the "larger than its negative" sign rule (FQ2.Parity, fq2.go:256-260), multiplication by the
// 507-bit G2 cofactor (g2.go:133-138).  The reference returns the unnormalised projective point; the ABI returns its
// affine form (canonical coordinates, so any correct ladder gives the same bits).
#pragma once
#include "codec.cuh"

namespace b381 {

#define B381_SHA_K                                                                                                     \
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98,        \
        0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786,    \
        0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8,    \
        0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,    \
        0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819,    \
        0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a,    \
        0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7,    \
        0xc67178f2
#if defined(__CUDACC__)
__device__ __constant__ uint32_t d_g2_cofactor[16] = {B381_G2_COFACTOR_LIMBS};
__device__ __constant__ uint32_t d_g2_cof_pos[16] = {B381_G2_COFACTOR_NAF_POS_LIMBS};
__device__ __constant__ uint32_t d_g2_cof_neg[16] = {B381_G2_COFACTOR_NAF_NEG_LIMBS};
__device__ __constant__ uint32_t d_kh2_naf[24] = B381_KH2_DIGITS_NAF_INIT;
__device__ __constant__ uint32_t d_sha_k[64] = {B381_SHA_K};
#endif
#if !defined(__CUDA_ARCH__)
static const uint32_t h_g2_cofactor[16] = {B381_G2_COFACTOR_LIMBS};
static const uint32_t h_g2_cof_pos[16] = {B381_G2_COFACTOR_NAF_POS_LIMBS};
static const uint32_t h_g2_cof_neg[16] = {B381_G2_COFACTOR_NAF_NEG_LIMBS};
static const uint32_t h_kh2_naf[24] = B381_KH2_DIGITS_NAF_INIT;
static const uint32_t h_sha_k[64] = {B381_SHA_K};
#endif

// SHA-256 of a message shorter than 56 bytes (one block); digest as 8 big-endian words (crypto/sha256 as used by
// hashFunc, g2.go:1034-1038)
HD void sha256_short(uint32_t dg[8], const uint8_t *msg, int len) {
    uint32_t w[64];
    uint8_t blk[64];
    for (int i = 0; i < 64; i++) blk[i] = i < len ? msg[i] : 0;
    blk[len] = 0x80;
    blk[62] = (uint8_t)((len * 8) >> 8); blk[63] = (uint8_t)(len * 8);
    for (int i = 0; i < 16; i++)
        w[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) | blk[4 * i + 3];
#define B381_ROR(x, n) (((x) >> (n)) | ((x) << (32 - (n))))
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = B381_ROR(w[i - 15], 7) ^ B381_ROR(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = B381_ROR(w[i - 2], 17) ^ B381_ROR(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    const uint32_t *K = B381_TAB(sha_k);
#pragma unroll 1
    for (int i = 0; i < 64; i++) {
        uint32_t S1 = B381_ROR(e, 6) ^ B381_ROR(e, 11) ^ B381_ROR(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + K[i] + w[i];
        uint32_t S0 = B381_ROR(a, 2) ^ B381_ROR(a, 13) ^ B381_ROR(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
#undef B381_ROR
    dg[0] = h[0] + a; dg[1] = h[1] + b; dg[2] = h[2] + c; dg[3] = h[3] + d;
    dg[4] = h[4] + e; dg[5] = h[5] + f; dg[6] = h[6] + g; dg[7] = h[7] + hh;
}
// the 256-bit digest as a field element (big.Int.SetBytes + FQReprFromBigInt + FQReprToFQ, g2.go:1053-1062; always < Q)
HD void fp_from_digest(fp &r, const uint32_t dg[8]) {
    fp raw;
    fp_set_zero(raw);
    for (int i = 0; i < 8; i++) raw.l[i] = dg[7 - i];
    fp_from_raw(r, raw);
}

// HashG2WithDomain(messageHash, domain) as an affine point
// psi on an XYZZ point: conjugate every coordinate, scale X and Y (the affine map of hash.go:341-366)
HD void g2_psi_xyzz(xyzz<Fp2Out> &p) {
    fp2 t;
    g2_psi(t, p.y, p.x, p.y);
    p.x = t;
    fp2_conj(p.zz, p.zz);
    fp2_conj(p.zzz, p.zzz);
}
// clearH2 (hash.go:368-389)
HDN void g2_clear_h2(xyzz<Fp2Out> &acc, const fp2 &x, const fp2 &y) {
    xyzz<Fp2Out> a, d;
    fp2 px, py, ny;
    point_mul<Fp2Out>(&a, &x, &y, B381_TAB(bls_x), 2);       // work = [x]P
    xyzz_madd(a, x, y);                                 //      + P
    g2_psi(px, py, x, y);
    fp2_neg(py, py);                                    // minusPsiP
    xyzz_madd(a, px, py);                               //      - psi(P)
    // work = [x]work: the inversion that makes `work` affine (~100 multiplications' worth, fp_inv) buys the Jacobian
    // doubling and mixed additions of point_mul (16 / 29 multiplications against 26 / 40 for an XYZZ base)
    xyzz<Fp2Out> b;
    if (xyzz_is_inf(a)) xyzz_set_inf(b);
    else {
        fp2 ax, ay;
        xyzz_to_affine<Fp2Out>(ax, ay, a);
        point_mul<Fp2Out>(&b, &ax, &ay, B381_TAB(bls_x), 2);
    }
    xyzz_madd(b, px, py);                               //      - psi(P)
    fp2_neg(ny, y);
    xyzz_madd(b, x, ny);                                //      - P
    xyzz_dbl_affine(d, x, y);                           // psi(psi(2P))
    g2_psi_xyzz(d); g2_psi_xyzz(d);
    xyzz_add(b, d);
    acc = b;
}
// ScaleByCofactor (g2.go:1041-1085 -> [h2] P, a 507-bit scalar) through the 64-bit endomorphism ladder:
//   [h2] P = [k] Q,  Q = clearH2(P) = [3 (x^2 - 1) h2] P,  k = (3 (x^2 - 1))^-1 mod r          ([h2] P has order r)
//   [k] Q  = [d0] Q - [d1] psi(Q) + [d2] psi^2(Q) - [d3] psi^3(Q),  k = sum d_i |x|^i           (psi = [x] = [-|x|] on G2)
// Q is normalised once so that the four bases are affine (psi maps affine coordinates to affine coordinates) and the
// ladder is 65 Jacobian doublings + 85 mixed additions (the d_i in non-adjacent form) instead of 508 + 176.
// B381_COFACTOR_LADDER selects the plain [h2] ladder (tests compare both with the oracle).
HDN void g2_scale_by_cofactor(xyzz<Fp2Out> *acc, const fp2 *px, const fp2 *py) {
#if defined(B381_COFACTOR_LADDER)
    point_mul_naf<Fp2Out>(acc, px, py, B381_TAB(g2_cof_pos), B381_TAB(g2_cof_neg), 16);
#else
    g2_clear_h2(*acc, *px, *py);
    if (xyzz_is_inf(*acc)) return;
    fp2 bx[4], by[4];
    xyzz_to_affine<Fp2Out>(bx[0], by[0], *acc);
    for (int i = 1; i < 4; i++) {
        g2_psi(bx[i], by[i], bx[i - 1], by[i - 1]);    // B_i = -psi(B_{i-1}) = (-1)^i psi^i(Q)
        fp2_neg(by[i], by[i]);
    }
    jac_pt<Fp2Out> a;
    fp2_set_one(a.x); fp2_set_one(a.y); fp2_set_zero(a.z);
    const uint32_t *tab = B381_TAB(kh2_naf);
#pragma unroll 1
    for (int j = 64; j >= 0; j--) {
        jac_dbl(a);
#pragma unroll 1
        for (int i = 0; i < 4; i++) {
            uint32_t pos = (tab[6 * i + (j >> 5)] >> (j & 31)) & 1, neg = (tab[6 * i + 3 + (j >> 5)] >> (j & 31)) & 1;
            if (pos) jac_madd(a, bx[i], by[i]);
            else if (neg) { fp2 ny; fp2_neg(ny, by[i]); jac_madd(a, bx[i], ny); }
        }
    }
    jac_to_xyzz(*acc, a);
#endif
}

HD void hash_g2_with_domain_one(g2_affine_pod *out, const uint8_t *msg32, const uint8_t *domain8) {
    uint8_t buf[41];
    uint32_t dg[8];
    for (int i = 0; i < 32; i++) buf[i] = msg32[i];
    for (int i = 0; i < 8; i++) buf[32 + i] = domain8[i];
    fp2 x, y, t, b, one;
    buf[40] = 1; sha256_short(dg, buf, 41); fp_from_digest(x.c0, dg);
    buf[40] = 2; sha256_short(dg, buf, 41); fp_from_digest(x.c1, dg);
    G2Codec::b_coeff(b);
    fp2_set_one(one);
    // FQ2.Sqrt fails exactly when x^3 + b is a non-square, i.e. when its norm is a non-residue of Fq: one Jacobi
    // symbol (fp_is_square) per rejected candidate instead of an Fq2 exponentiation (1 330 multiplications), the lanes
    // of a warp leave the divergent search before the rest of the root, and the accepted candidate's norm root is
    // the first half of that root (fp2_sqrt_from_norm_root)
    fp n0, n1;
    for (;;) {
        fp2_sqr(&t, &x);
        fp2_mul(&t, &t, &x);
        fp2_add(t, t, b);
        fp_sqr(n0, t.c0); fp_sqr(n1, t.c1);
        fp_add(n0, n0, n1);
        if (fp_is_square(n0)) { fp_sqrt(&n1, &n0); break; }
        fp2_add(x, x, one);
    }
    if (fp_is_zero(t.c1)) fp2_sqrt_alg9(&y, &t);
    else fp2_sqrt_from_norm_root(&y, &t, &n1);
    fp2_neg(t, y);
    if (!(fp2_cmp(y, t) > 0)) y = t;                   // "favor the lower y value": keep the one with Parity() true
    xyzz<Fp2Out> acc;
    g2_scale_by_cofactor(&acc, &x, &y);
    if (xyzz_is_inf(acc)) { fp2_set_zero(x); fp2_set_one(y); G2Codec::store(out, x, y, true); return; }
    xyzz_to_affine<Fp2Out>(x, y, acc);
    G2Codec::store(out, x, y, false);
}

#if defined(__CUDACC__)
// out[i] = HashG2WithDomain(msg[i], domain[i * domain_stride]); domain_stride 0 = one domain for the whole batch
__global__ void __launch_bounds__(64, CODEC_MIN_BLOCKS) k_hash_g2_with_domain(const uint8_t *__restrict__ msg, const uint8_t *__restrict__ domain,
                                                            size_t domain_stride, size_t n, g2_affine_pod *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    hash_g2_with_domain_one(out + i, msg + 32 * i, domain + 8 * i * domain_stride);
}
// The two Miller pairs of g1pubs.VerifyWithDomain (g1pubs/bls.go:171-174 -> CompareTwoPairings(G1One, sig, pub, H),
// pairing.go:140-147): (P, Q)[2i] = (G1One, sig[i]), (P, Q)[2i+1] = (-pub[i], H[i]).  valid[i] = 0 when the key or the
// signature failed to deserialise or is the point at infinity (the reference returns an error from Deserialize* in
// the first case and panics in MillerLoop in the second): such a check is reported false.
__global__ void __launch_bounds__(128) k_verify_pairs(const g1_affine_pod *__restrict__ pub, const uint8_t *__restrict__ pub_status,
                                                      const g2_affine_pod *__restrict__ sig, const uint8_t *__restrict__ sig_status,
                                                      const g2_affine_pod *__restrict__ H, size_t n, g1_affine_pod *__restrict__ P,
                                                      g2_affine_pod *__restrict__ Q, uint32_t *__restrict__ group_off,
                                                      uint8_t *__restrict__ valid) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) group_off[0] = 0;
    if (i >= n) return;
    group_off[i + 1] = (uint32_t)(2 * i + 2);
    g1_affine_pod a = pub[i];
    valid[i] = (pub_status[i] == 0 && sig_status[i] == 0 && !a.inf && !sig[i].inf) ? 1 : 0;
    fp y;
    fp_load_u64(y, a.y);
    fp_neg(y, y);
    fp_store_u64(a.y, y);
    P[2 * i + 1] = a;
    g1_affine_pod one;
    const uint32_t gx[12] = {B381_G1_GEN_X_LIMBS}, gy[12] = {B381_G1_GEN_Y_LIMBS};
    fp t;
    fp_load_tab(t, gx); fp_store_u64(one.x, t);
    fp_load_tab(t, gy); fp_store_u64(one.y, t);
    one.inf = 0;
    for (int k = 0; k < 7; k++) one.pad[k] = 0;
    P[2 * i] = one;
    Q[2 * i] = sig[i];
    Q[2 * i + 1] = H[i];
}
__global__ void k_and_bytes(uint8_t *__restrict__ ok, const uint8_t *__restrict__ valid, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ok[i] = (ok[i] && valid[i]) ? 1 : 0;
}
#endif

}  // namespace b381
