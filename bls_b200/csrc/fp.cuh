// fp.cuh -- BLS12-381 base field Fq on the device: 12 x u32 limbs in registers, Montgomery form
// (R = 2^384), every value canonical in [0, Q) exactly like the reference keeps it
// (fq.go:41-45).  Replaces the reference's L0/L1 layers on the hot path:
//   MultiplyFQRepr + MontReduce (stub_fallback.go:11-116, primitivefuncs_amd64.s) -> fp_mul
//   FQ.AddAssign/SubAssign/NegAssign/DoubleAssign (fq.go:65-143)                 -> fp_add/...
//   FQ.Inverse (fq.go:224-266, binary Euclid)        -> fp_inv (almost-inverse on the limbs, tower.cuh; same value)
// Memory layout of an element is the reference's FQRepr: 6 x u64 little-endian limbs = 12 x u32.
//
// The arithmetic bodies are __host__ __device__: the device path is inline PTX (carry chains that
// ptxas turns into IMAD.WIDE.U32 + carry); the host path is portable C used ONLY by the CPU unit
// tests of this logic (tests/emu), never by the product library.
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define HDN __host__ __device__ __noinline__
#else
#define HD inline
#define HDN __attribute__((noinline))
#define __align__(n) __attribute__((aligned(n)))
#endif

#include "fp_mul_asm.inc"

namespace b381 {

struct __align__(16) fp { uint32_t l[12]; };

#define B381_Q_LIMBS                                                                                  \
    0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u, 0xf38512bfu,        \
        0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau
// R mod Q = FQOne (fq.go:23)
#define B381_ONE_LIMBS                                                                                \
    0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u, 0x70525745u,        \
        0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u

HD void fp_set_zero(fp &r) {
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = 0;
}
HD void fp_set_one(fp &r) {
    const uint32_t one[12] = {B381_ONE_LIMBS};
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = one[i];
}
HD bool fp_is_zero(const fp &a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) o |= a.l[i];
    return o == 0;
}
HD bool fp_eq(const fp &a, const fp &b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) o |= a.l[i] ^ b.l[i];
    return o == 0;
}

// ---------------------------------------------------------------------------------------------
// host reference bodies (CPU unit tests of the device logic only)
// ---------------------------------------------------------------------------------------------
#if !defined(__CUDA_ARCH__)
namespace hostimpl {
// Fq multiplications executed by the host build (tests/emu pins the per-pairing counts bench.py's roofline uses)
#if defined(B381_HOST_COUNT)
static thread_local unsigned long long g_mul_count = 0, g_dot2_count = 0;
#define B381_COUNT_MUL(k) (hostimpl::g_mul_count += (k))
#define B381_COUNT_DOT2(k) (hostimpl::g_dot2_count += (k))
#else
#define B381_COUNT_MUL(k) ((void)0)
#define B381_COUNT_DOT2(k) ((void)0)
#endif
static inline void cond_sub_q(uint32_t x[12], uint32_t top) {
    const uint32_t q[12] = {B381_Q_LIMBS};
    uint32_t d[12];
    int64_t br = 0;
    for (int i = 0; i < 12; i++) {
        int64_t t = (int64_t)x[i] - q[i] + br;
        d[i] = (uint32_t)t;
        br = t >> 32;
    }
    if (top || br == 0)
        for (int i = 0; i < 12; i++) x[i] = d[i];
}
static inline void mul(uint32_t r[12], const uint32_t a[12], const uint32_t b[12]) {
    B381_COUNT_MUL(1);
    const uint32_t q[12] = {B381_Q_LIMBS};
    uint32_t t[14] = {0};
    for (int i = 0; i < 12; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 12; j++) {
            uint64_t v = (uint64_t)a[j] * b[i] + t[j] + c;
            t[j] = (uint32_t)v; c = v >> 32;
        }
        uint64_t v = (uint64_t)t[12] + c;
        t[12] = (uint32_t)v; t[13] = (uint32_t)(v >> 32);
        uint32_t m = t[0] * 0xfffcfffdu;
        c = ((uint64_t)m * q[0] + t[0]) >> 32;
        for (int j = 1; j < 12; j++) {
            uint64_t w = (uint64_t)m * q[j] + t[j] + c;
            t[j - 1] = (uint32_t)w; c = w >> 32;
        }
        v = (uint64_t)t[12] + c;
        t[11] = (uint32_t)v;
        t[12] = t[13] + (uint32_t)(v >> 32);
    }
    for (int i = 0; i < 12; i++) r[i] = t[i];
    cond_sub_q(r, t[12]);
}
}  // namespace hostimpl
#endif

// r = a + b mod Q   (fq.go:65-68)
HD void fp_add(fp &r, const fp &a, const fp &b) {
#if defined(__CUDA_ARCH__)
    uint32_t s[12], d[12], bw;
    asm("add.cc.u32 %0, %12, %24;\n\t"
        "addc.cc.u32 %1, %13, %25;\n\t"
        "addc.cc.u32 %2, %14, %26;\n\t"
        "addc.cc.u32 %3, %15, %27;\n\t"
        "addc.cc.u32 %4, %16, %28;\n\t"
        "addc.cc.u32 %5, %17, %29;\n\t"
        "addc.cc.u32 %6, %18, %30;\n\t"
        "addc.cc.u32 %7, %19, %31;\n\t"
        "addc.cc.u32 %8, %20, %32;\n\t"
        "addc.cc.u32 %9, %21, %33;\n\t"
        "addc.cc.u32 %10, %22, %34;\n\t"
        "addc.u32 %11, %23, %35;\n\t"
        : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]),
          "=r"(s[8]), "=r"(s[9]), "=r"(s[10]), "=r"(s[11])
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
          "r"(a.l[8]), "r"(a.l[9]), "r"(a.l[10]), "r"(a.l[11]),
          "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]),
          "r"(b.l[8]), "r"(b.l[9]), "r"(b.l[10]), "r"(b.l[11]));
    asm("sub.cc.u32 %0, %13, 0xffffaaab;\n\t"
        "subc.cc.u32 %1, %14, 0xb9feffff;\n\t"
        "subc.cc.u32 %2, %15, 0xb153ffff;\n\t"
        "subc.cc.u32 %3, %16, 0x1eabfffe;\n\t"
        "subc.cc.u32 %4, %17, 0xf6b0f624;\n\t"
        "subc.cc.u32 %5, %18, 0x6730d2a0;\n\t"
        "subc.cc.u32 %6, %19, 0xf38512bf;\n\t"
        "subc.cc.u32 %7, %20, 0x64774b84;\n\t"
        "subc.cc.u32 %8, %21, 0x434bacd7;\n\t"
        "subc.cc.u32 %9, %22, 0x4b1ba7b6;\n\t"
        "subc.cc.u32 %10, %23, 0x397fe69a;\n\t"
        "subc.cc.u32 %11, %24, 0x1a0111ea;\n\t"
        "subc.u32 %12, 0, 0;\n\t"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]),
          "=r"(d[8]), "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(bw)
        : "r"(s[0]), "r"(s[1]), "r"(s[2]), "r"(s[3]), "r"(s[4]), "r"(s[5]), "r"(s[6]), "r"(s[7]),
          "r"(s[8]), "r"(s[9]), "r"(s[10]), "r"(s[11]));
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = bw ? s[i] : d[i];
#else
    uint32_t s[12];
    uint64_t c = 0;
    for (int i = 0; i < 12; i++) { uint64_t v = (uint64_t)a.l[i] + b.l[i] + c; s[i] = (uint32_t)v; c = v >> 32; }
    hostimpl::cond_sub_q(s, (uint32_t)c);
    for (int i = 0; i < 12; i++) r.l[i] = s[i];
#endif
}

// r = a - b mod Q   (fq.go:82-87)
HD void fp_sub(fp &r, const fp &a, const fp &b) {
#if defined(__CUDA_ARCH__)
    uint32_t d[12], bw;
    asm("sub.cc.u32 %0, %13, %25;\n\t"
        "subc.cc.u32 %1, %14, %26;\n\t"
        "subc.cc.u32 %2, %15, %27;\n\t"
        "subc.cc.u32 %3, %16, %28;\n\t"
        "subc.cc.u32 %4, %17, %29;\n\t"
        "subc.cc.u32 %5, %18, %30;\n\t"
        "subc.cc.u32 %6, %19, %31;\n\t"
        "subc.cc.u32 %7, %20, %32;\n\t"
        "subc.cc.u32 %8, %21, %33;\n\t"
        "subc.cc.u32 %9, %22, %34;\n\t"
        "subc.cc.u32 %10, %23, %35;\n\t"
        "subc.cc.u32 %11, %24, %36;\n\t"
        "subc.u32 %12, 0, 0;\n\t"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]),
          "=r"(d[8]), "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(bw)
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
          "r"(a.l[8]), "r"(a.l[9]), "r"(a.l[10]), "r"(a.l[11]),
          "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]),
          "r"(b.l[8]), "r"(b.l[9]), "r"(b.l[10]), "r"(b.l[11]));
    // add Q back under the borrow mask
    asm("add.cc.u32 %0, %0, %12;\n\t"
        "addc.cc.u32 %1, %1, %13;\n\t"
        "addc.cc.u32 %2, %2, %14;\n\t"
        "addc.cc.u32 %3, %3, %15;\n\t"
        "addc.cc.u32 %4, %4, %16;\n\t"
        "addc.cc.u32 %5, %5, %17;\n\t"
        "addc.cc.u32 %6, %6, %18;\n\t"
        "addc.cc.u32 %7, %7, %19;\n\t"
        "addc.cc.u32 %8, %8, %20;\n\t"
        "addc.cc.u32 %9, %9, %21;\n\t"
        "addc.cc.u32 %10, %10, %22;\n\t"
        "addc.u32 %11, %11, %23;\n\t"
        : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]), "+r"(d[4]), "+r"(d[5]), "+r"(d[6]), "+r"(d[7]),
          "+r"(d[8]), "+r"(d[9]), "+r"(d[10]), "+r"(d[11])
        : "r"(bw & 0xffffaaabu), "r"(bw & 0xb9feffffu), "r"(bw & 0xb153ffffu), "r"(bw & 0x1eabfffeu),
          "r"(bw & 0xf6b0f624u), "r"(bw & 0x6730d2a0u), "r"(bw & 0xf38512bfu), "r"(bw & 0x64774b84u),
          "r"(bw & 0x434bacd7u), "r"(bw & 0x4b1ba7b6u), "r"(bw & 0x397fe69au), "r"(bw & 0x1a0111eau));
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = d[i];
#else
    const uint32_t q[12] = {B381_Q_LIMBS};
    uint32_t d[12];
    int64_t br = 0;
    for (int i = 0; i < 12; i++) { int64_t t = (int64_t)a.l[i] - b.l[i] + br; d[i] = (uint32_t)t; br = t >> 32; }
    if (br) {
        uint64_t c = 0;
        for (int i = 0; i < 12; i++) { uint64_t v = (uint64_t)d[i] + q[i] + c; d[i] = (uint32_t)v; c = v >> 32; }
    }
    for (int i = 0; i < 12; i++) r.l[i] = d[i];
#endif
}

HD void fp_dbl(fp &r, const fp &a) { fp_add(r, a, a); }   // fq.go:140-143

// r = -a mod Q, with -0 = 0   (fq.go:121-127)
HD void fp_neg(fp &r, const fp &a) {
    fp z;
    fp_set_zero(z);
    fp_sub(r, z, a);   // 0 - a borrows iff a != 0, and then adds Q
}

// r = a * b * 2^-384 mod Q   (fq.go:76-79) -- fully inlined body (~300 IMAD.WIDE.U32[.X])
HD void fp_mul_inl(fp &r, const fp &a, const fp &b) {
#if defined(__CUDA_ARCH__)
    asm(FP_MUL_PTX
        : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]),
          "=r"(r.l[7]), "=r"(r.l[8]), "=r"(r.l[9]), "=r"(r.l[10]), "=r"(r.l[11])
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
          "r"(a.l[8]), "r"(a.l[9]), "r"(a.l[10]), "r"(a.l[11]),
          "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]),
          "r"(b.l[8]), "r"(b.l[9]), "r"(b.l[10]), "r"(b.l[11]));
#else
    uint32_t t[12];
    hostimpl::mul(t, a.l, b.l);
    for (int i = 0; i < 12; i++) r.l[i] = t[i];
#endif
}
// r = (a*b + c*d) * 2^-384 mod Q, canonical; operands may be as large as 2Q (device only: the VM's MUL2 step)
#if defined(__CUDACC__)
__device__ __forceinline__ void fp_dot2_inl(fp &r, const fp &a, const fp &b, const fp &c, const fp &d) {
    asm(FP_DOT2_PTX
        : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]),
          "=r"(r.l[7]), "=r"(r.l[8]), "=r"(r.l[9]), "=r"(r.l[10]), "=r"(r.l[11])
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
          "r"(a.l[8]), "r"(a.l[9]), "r"(a.l[10]), "r"(a.l[11]),
          "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]),
          "r"(b.l[8]), "r"(b.l[9]), "r"(b.l[10]), "r"(b.l[11]),
          "r"(c.l[0]), "r"(c.l[1]), "r"(c.l[2]), "r"(c.l[3]), "r"(c.l[4]), "r"(c.l[5]), "r"(c.l[6]), "r"(c.l[7]),
          "r"(c.l[8]), "r"(c.l[9]), "r"(c.l[10]), "r"(c.l[11]),
          "r"(d.l[0]), "r"(d.l[1]), "r"(d.l[2]), "r"(d.l[3]), "r"(d.l[4]), "r"(d.l[5]), "r"(d.l[6]), "r"(d.l[7]),
          "r"(d.l[8]), "r"(d.l[9]), "r"(d.l[10]), "r"(d.l[11]));
}
#endif
// r = a + b without reduction (operands of a multiplication may be as large as 2Q)
HD void fp_add_nr(fp &r, const fp &a, const fp &b) {
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %0, %12, %24;\n\taddc.cc.u32 %1, %13, %25;\n\taddc.cc.u32 %2, %14, %26;\n\taddc.cc.u32 %3, %15, %27;\n\t"
        "addc.cc.u32 %4, %16, %28;\n\taddc.cc.u32 %5, %17, %29;\n\taddc.cc.u32 %6, %18, %30;\n\taddc.cc.u32 %7, %19, %31;\n\t"
        "addc.cc.u32 %8, %20, %32;\n\taddc.cc.u32 %9, %21, %33;\n\taddc.cc.u32 %10, %22, %34;\n\taddc.u32 %11, %23, %35;"
        : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7]),
          "=r"(r.l[8]), "=r"(r.l[9]), "=r"(r.l[10]), "=r"(r.l[11])
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
          "r"(a.l[8]), "r"(a.l[9]), "r"(a.l[10]), "r"(a.l[11]),
          "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]),
          "r"(b.l[8]), "r"(b.l[9]), "r"(b.l[10]), "r"(b.l[11]));
#else
    uint64_t c = 0;
    for (int i = 0; i < 12; i++) { uint64_t v = (uint64_t)a.l[i] + b.l[i] + c; r.l[i] = (uint32_t)v; c = v >> 32; }
#endif
}
// r = Q - a   (in (0, Q]; used as a negated multiplication operand)
HD void fp_qminus(fp &r, const fp &a) {
    const uint32_t q[12] = {B381_Q_LIMBS};
#if defined(__CUDA_ARCH__)
    asm("sub.cc.u32 %0, %12, %24;\n\tsubc.cc.u32 %1, %13, %25;\n\tsubc.cc.u32 %2, %14, %26;\n\tsubc.cc.u32 %3, %15, %27;\n\t"
        "subc.cc.u32 %4, %16, %28;\n\tsubc.cc.u32 %5, %17, %29;\n\tsubc.cc.u32 %6, %18, %30;\n\tsubc.cc.u32 %7, %19, %31;\n\t"
        "subc.cc.u32 %8, %20, %32;\n\tsubc.cc.u32 %9, %21, %33;\n\tsubc.cc.u32 %10, %22, %34;\n\tsubc.u32 %11, %23, %35;"
        : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7]),
          "=r"(r.l[8]), "=r"(r.l[9]), "=r"(r.l[10]), "=r"(r.l[11])
        : "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]), "r"(q[4]), "r"(q[5]), "r"(q[6]), "r"(q[7]), "r"(q[8]), "r"(q[9]),
          "r"(q[10]), "r"(q[11]),
          "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
          "r"(a.l[8]), "r"(a.l[9]), "r"(a.l[10]), "r"(a.l[11]));
#else
    int64_t br = 0;
    for (int i = 0; i < 12; i++) { int64_t t = (int64_t)q[i] - a.l[i] + br; r.l[i] = (uint32_t)t; br = t >> 32; }
#endif
}
// Out-of-line multiply with BY-VALUE operands: nvcc's device ABI passes the 2 x 12 limbs and the
// result in registers (no local-memory traffic), so every caller shares one 5 KB copy of the
// multiplication -- the instruction footprint of the tower stays inside the 32 KB L1.5 I-cache
// (profiles/r01_v1_ncu_summary.md: `no_instruction` was the top stall with the body inlined).
HDN fp fp_mul_v(fp a, fp b) { fp r; fp_mul_inl(r, a, b); return r; }
// r = (a*b + c*d) * 2^-384 mod Q with ONE reduction (444 instead of 600 wide MACs): the rows of an Fq2 product
// (fq2.go:116-130 computes the same values with Karatsuba: three reduced multiplications and five additions)
HDN fp fp_dot2_v(fp a, fp b, fp c, fp d) {
    fp r;
#if defined(__CUDA_ARCH__)
    fp_dot2_inl(r, a, b, c, d);
#else
    uint32_t t[12], u[12];
    B381_COUNT_DOT2(1);                                // (the device runs ONE 444-MAC body; the two host products below are counted too)
    hostimpl::mul(t, a.l, b.l);
    hostimpl::mul(u, c.l, d.l);
    fp x, y;
    for (int i = 0; i < 12; i++) { x.l[i] = t[i]; y.l[i] = u[i]; }
    fp_add(r, x, y);
#endif
    return r;
}
// the same with operands and result in memory: the caller passes five pointers instead of marshalling 48 registers in and 12
// out around every call (fp2_mul shrinks from 158 to ~40 instructions; the kernels are instruction-cache bound)
HDN void fp_dot2_p(fp *r, const fp *a, const fp *b, const fp *c, const fp *d) {
#if defined(__CUDA_ARCH__)
    fp x = *a, y = *b, z = *c, w = *d, o;
    fp_dot2_inl(o, x, y, z, w);
    *r = o;
#else
    *r = fp_dot2_v(*a, *b, *c, *d);
#endif
}
HD void fp_mul(fp &r, const fp &a, const fp &b) { r = fp_mul_v(a, b); }
HD void fp_sqr(fp &r, const fp &a) { r = fp_mul_v(a, a); }   // fq.go:151-198 through the multiplier (the pairing kernels: no second body)
// FQ.SquareAssign (fq.go:151-198) with its own body: the 66 cross products once, doubled, and the 12 diagonal products -- 228 wide
// MACs instead of 300 (78 products + 144 reduction + 12 quotient words; FP_SQR_PTX, tools/gen_fp_asm.py).  Used by the long
// squaring chains: square roots (380 squarings each) in decompression and hashing.  Out of line, by value.
HDN fp fp_sqr_v(fp a) {
#if defined(__CUDA_ARCH__)
    fp r;
    asm(FP_SQR_PTX
        : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]),
          "=r"(r.l[7]), "=r"(r.l[8]), "=r"(r.l[9]), "=r"(r.l[10]), "=r"(r.l[11])
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
          "r"(a.l[8]), "r"(a.l[9]), "r"(a.l[10]), "r"(a.l[11]));
    return r;
#else
    return fp_mul_v(a, a);
#endif
}

}  // namespace b381
