// emu.cc -- TEST-ONLY host build of the device headers (bls_b200/csrc/*.cuh compile as plain
// C++ with the portable limb bodies).  It lets the CPU test suite exercise the exact tower /
// pairing / group-law logic the kernels run, against the oracle, without a GPU.  The product
// library (libb381.so) never contains or calls this path.
#define B381_HOST_COUNT 1
#include "pairing.cuh"
#include "curve.cuh"
#include <cstring>
#include <cstddef>
using namespace b381;

static void ld12(fp12 *a, const uint64_t *s) { fp12_load_u64(a, s); }

extern "C" {
// op: 0 mul 1 add 2 sub 3 sqr 4 neg 5 dbl 6 inv 8 inv by the Fermat chain
void emu_fp_op(int op, const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) {
    for (size_t i = 0; i < n; i++) {
        fp x, y, r; fp_load_u64(x, a + 6 * i); fp_load_u64(y, b + 6 * i);
        switch (op) {
            case 0: fp_mul(r, x, y); break;
            case 1: fp_add(r, x, y); break;
            case 2: fp_sub(r, x, y); break;
            case 3: fp_sqr(r, x); break;
            case 4: fp_neg(r, x); break;
            case 5: fp_dbl(r, x); break;
            case 6: fp_inv(&r, &x); break;
            case 8: fp_inv_fermat(&r, &x); break;
            default: r = x;
        }
        fp_store_u64(o + 6 * i, r);
    }
}
// op: 0 mul 3 sqr 6 inv 7 frobenius(arg) 12 conj 13 mul_by_014(b.c0.c0,b.c0.c1,b.c1.c1) 15 cyclotomic sqr 16 exp_by_x(arg)
int emu_fp12_op(int op, uint64_t arg, const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) {
    int ok = 1;
    for (size_t i = 0; i < n; i++) {
        fp12 x, y, r; ld12(&x, a + 72 * i); ld12(&y, b + 72 * i);
        fp12_copy(&r, &x);
        switch (op) {
            case 0: fp12_mul(&r, &x, &y); break;
            case 3: fp12_sqr(&r, &x); break;
            case 6: ok &= fp12_inv(&r, &x) ? 1 : 0; break;
            case 7: fp12_frobenius(&r, &x, (int)arg); break;
            case 12: fp12_conj(&r, &x); break;
            case 13: fp12_mul_by_014(&r, &y.c0.c0, &y.c0.c1, &y.c1.c1); break;
            case 15: fp12_cyclotomic_sqr(&r, &x); break;
            case 16: exp_by_x(&r, &x, arg); break;
        }
        fp12_store_u64(o + 72 * i, &r);
    }
    return ok;
}
void emu_miller_loop(const g1_affine_pod *p, const g2_affine_pod *q, size_t n, uint64_t *out) {
    for (size_t i = 0; i < n; i++) { fp12 f; miller_loop_one(&f, p + i, q + i); fp12_store_u64(out + 72 * i, &f); }
}
void emu_final_exp(const uint64_t *in, size_t n, uint64_t *out, uint8_t *ok) {
    for (size_t i = 0; i < n; i++) {
        // odd units run the in-place form the kernels use, even units the two-buffer form: both must agree with the oracle
        fp12 f, r; ld12(&f, in + 72 * i);
        fp12_set_one(&r);
        if (i & 1) { ok[i] = final_exp_one(&f, &f) ? 1 : 0; if (ok[i]) fp12_copy(&r, &f); }
        else ok[i] = final_exp_one(&r, &f) ? 1 : 0;
        fp12_store_u64(out + 72 * i, &r);
    }
}
}
#include "emu_curve.inc"
extern "C" void emu_miller_loop2(const g1_affine_pod *p, const g2_affine_pod *q, size_t ngroups, uint64_t *out) {
    for (size_t i = 0; i < ngroups; i++) { fp12 f; miller_loop_two(&f, p + 2 * i, q + 2 * i); fp12_store_u64(out + 72 * i, &f); }
}
#include "emu_quad.inc"
#include "emu_duo.inc"

// Fq multiplications (hostimpl::mul calls: products and squarings, one each; a two-product dot product counts two)
// executed on this thread since the last reset
extern "C" unsigned long long emu_mul_count(int reset) {
    unsigned long long v = hostimpl::g_mul_count;
    if (reset) hostimpl::g_mul_count = 0;
    return v;
}
// two-product dot products (fp_dot2_v: one 444-MAC body on the device) among them; each also added 2 to emu_mul_count
extern "C" unsigned long long emu_dot2_count(int reset) {
    unsigned long long v = hostimpl::g_dot2_count;
    if (reset) hostimpl::g_dot2_count = 0;
    return v;
}

// prepared G2 points (pairing.cuh): coefficients, Miller loop from them, the fused + prepared two-pair loop
extern "C" void emu_g2_prepare(const g2_affine_pod *q, size_t n, g2_prepared_pod *out) {
    for (size_t i = 0; i < n; i++) g2_prepare_one(out + i, q + i);
}
extern "C" void emu_miller_loop_prepared(const g1_affine_pod *p, const g2_prepared_pod *prep, const uint32_t *idx, size_t n, uint64_t *out) {
    for (size_t i = 0; i < n; i++) { fp12 f; miller_loop_prepared_one(&f, p + i, prep + (idx ? idx[i] : i)); fp12_store_u64(out + 72 * i, &f); }
}
extern "C" void emu_miller_loop_fused_prepared(const g1_affine_pod *p, const g2_affine_pod *q0, const g2_prepared_pod *prep, const uint32_t *idx,
                                               size_t ngroups, uint64_t *out) {
    for (size_t i = 0; i < ngroups; i++) { fp12 f; miller_loop_fused_prepared(&f, p + 2 * i, q0 + i, prep + idx[i]); fp12_store_u64(out + 72 * i, &f); }
}
