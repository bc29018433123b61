// b381.cu -- kernels and C ABI (include/b381.h) of the B200 BLS12-381 engine.
// One TU on purpose: the __constant__ tables and the out-of-line Fq2 routines are shared by
// every kernel.  There is no CPU fallback in this library: without a CUDA device b381_init
// returns B381_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <new>

#include "../../include/b381.h"
#include "pairing.cuh"
#include "curve.cuh"
#include "agg.cuh"
#include "codec.cuh"
#include "hash.cuh"
#include "swu.cuh"
#include "vm.cuh"
#include "quad.cuh"
#include "duo.cuh"
#include "testops.cuh"
#include "vm_programs.inc"
#include <stdlib.h>

using namespace b381;

static_assert(sizeof(b381_g1_affine) == sizeof(g1_affine_pod) && sizeof(b381_g1_affine) == 104, "layout");
static_assert(sizeof(b381_g2_affine) == sizeof(g2_affine_pod) && sizeof(b381_g2_affine) == 200, "layout");
static_assert(sizeof(b381_fp12) == 576 && sizeof(fp12) == 576, "layout");
static_assert(sizeof(b381_g2_prepared) == sizeof(g2_prepared_pod) && sizeof(b381_g2_prepared) == 68 * 288 + 8, "layout");
static_assert(sizeof(b381_g1_jac) == 144 && sizeof(b381_g2_jac) == 288, "layout");

// ---------------------------------------------------------------------------------------------
// kernels: one thread per independent unit
// ---------------------------------------------------------------------------------------------
#ifdef PAIRING_BLOCK_OVERRIDE
#define PAIRING_BLOCK PAIRING_BLOCK_OVERRIDE
#else
#define PAIRING_BLOCK 64    // 1024 blocks for 2^16 pairings spread evenly over 148 SMs x 8 resident blocks (measured +2.5 % over 128)
#endif
#ifndef PAIRING_MIN_BLOCKS
#define PAIRING_MIN_BLOCKS 8   // 128 registers -> 16 warps/SM: measured 1.18 M pairings/s vs 0.90 M at 12 warps (162 regs), 1.10 M at 20
#endif

// out[i] = MillerLoop(p[i], q[i])   (pairing.go:16-75 fused with g2.go:650-801)
__global__ void __launch_bounds__(PAIRING_BLOCK, PAIRING_MIN_BLOCKS) k_miller_loop(const g1_affine_pod *__restrict__ p,
                                                                 const g2_affine_pod *__restrict__ q, size_t n,
                                                                 uint64_t *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp12 f;
    miller_loop_one(&f, p + i, q + i);
    fp12_store_u64(out + 72 * i, &f);
}

// prod[g] = MillerLoop of the two pairs (p, q)[2g], (p, q)[2g+1] with a shared accumulator (pairing.go:16-75, two items)
__global__ void __launch_bounds__(PAIRING_BLOCK, PAIRING_MIN_BLOCKS) k_miller_loop2(const g1_affine_pod *__restrict__ p,
                                                                  const g2_affine_pod *__restrict__ q, size_t ngroups,
                                                                  uint64_t *__restrict__ out) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    fp12 f;
    miller_loop_two(&f, p + 2 * g, q + 2 * g);
    fp12_store_u64(out + 72 * g, &f);
}

// prep[i] = G2AffineToPrepared(q[i])   (g2.go:650-801)
__global__ void __launch_bounds__(PAIRING_BLOCK, PAIRING_MIN_BLOCKS) k_g2_prepare(const g2_affine_pod *__restrict__ q, size_t n,
                                                                                  g2_prepared_pod *__restrict__ prep) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    g2_prepare_one(prep + i, q + i);
}
// out[i] = MillerLoop({p[i], prep[idx ? idx[i] : i]})   (pairing.go:16-75 on a MillerLoopItem whose Q was prepared before)
__global__ void __launch_bounds__(PAIRING_BLOCK, PAIRING_MIN_BLOCKS) k_miller_loop_prepared(const g1_affine_pod *__restrict__ p,
                                                                                            const g2_prepared_pod *__restrict__ prep,
                                                                                            const uint32_t *__restrict__ idx, size_t n,
                                                                                            uint64_t *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp12 f;
    miller_loop_prepared_one(&f, p + i, prep + (idx ? idx[i] : i));
    fp12_store_u64(out + 72 * i, &f);
}
// prod[g] = MillerLoop of (p[2g], q[2g]) computed and (p[2g+1], prep[idx[g]]) read from its prepared coefficients, shared accumulator
__global__ void __launch_bounds__(PAIRING_BLOCK, PAIRING_MIN_BLOCKS) k_miller_loop_fused_prepared(const g1_affine_pod *__restrict__ p,
                                                                                                  const g2_affine_pod *__restrict__ q,
                                                                                                  const g2_prepared_pod *__restrict__ prep,
                                                                                                  const uint32_t *__restrict__ idx, size_t ngroups,
                                                                                                  uint64_t *__restrict__ out) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    fp12 f;
    miller_loop_fused_prepared(&f, p + 2 * g, q + 2 * g, prep + idx[g]);
    fp12_store_u64(out + 72 * g, &f);
}

// out[i] = FinalExponentiation(in[i])   (pairing.go:79-129); in-place allowed
__global__ void __launch_bounds__(PAIRING_BLOCK, PAIRING_MIN_BLOCKS) k_final_exp(const uint64_t *in, size_t n, uint64_t *out,
                                                               uint8_t *ok) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp12 f;
    fp12_load_u64(&f, in + 72 * i);
    bool good = final_exp_one(&f, &f);                 // in place; f == 0 leaves f untouched: report 1 like the VM path
    if (!good) fp12_set_one(&f);
    fp12_store_u64(out + 72 * i, &f);
    if (ok) ok[i] = good ? 1 : 0;
}

// prod[g] = product of ml[group_off[g] .. group_off[g+1])   (the shared accumulator f of pairing.go:40-69)
__global__ void __launch_bounds__(PAIRING_BLOCK, PAIRING_MIN_BLOCKS) k_group_product(const uint64_t *__restrict__ ml,
                                                                   const uint32_t *__restrict__ group_off,
                                                                   size_t ngroups, uint64_t *__restrict__ prod, int first_only) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    fp12 acc, t;
    fp12_set_one(&acc);
    uint32_t lo = group_off[g], hi = group_off[g + 1];
    if (first_only && hi > lo) hi = lo + 1;        // k_group_tree has already folded the group into its first element
    for (uint32_t i = lo; i < hi; i++) {
        fp12_load_u64(&t, ml + 72 * (size_t)i);
        if (i == lo) fp12_copy(&acc, &t);
        else fp12_mul(&acc, &acc, &t);
    }
    fp12_store_u64(prod + 72 * g, &acc);
}

// Large groups (VerifyAggregate with many messages, g1pubs/bls.go:252-282; random-linear-combination batches): the product
// of a group is formed as a tree over its Miller values instead of one serial chain.  Level l: element at position pos of
// its group absorbs the element at pos + 2^l when pos is a multiple of 2^(l+1); after ceil(log2(size)) levels the first
// element of the group holds the product.  *maxsize (largest group) lets surplus launches exit at once.
__global__ void k_group_maxsize(const uint32_t *__restrict__ group_off, size_t ngroups, uint32_t *__restrict__ maxsize) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < ngroups) atomicMax(maxsize, group_off[g + 1] - group_off[g]);
}
__global__ void __launch_bounds__(PAIRING_BLOCK, PAIRING_MIN_BLOCKS) k_group_tree(uint64_t *__restrict__ ml, const uint32_t *__restrict__ group_off,
                                                                                  size_t ngroups, size_t npairs, int level,
                                                                                  const uint32_t *__restrict__ maxsize) {
    if ((1u << level) >= *maxsize) return;
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= npairs) return;
    size_t lo = 0, hi = ngroups;                   // largest g with group_off[g] <= e
    while (hi - lo > 1) { size_t mid = (lo + hi) >> 1; if (group_off[mid] <= e) lo = mid; else hi = mid; }
    uint32_t start = group_off[lo], end = group_off[lo + 1];
    if (e >= end) return;                          // e lies in an empty tail
    uint32_t pos = (uint32_t)e - start, step = 1u << level;
    if ((pos & (2 * step - 1)) || e + step >= end) return;
    fp12 a, b;
    fp12_load_u64(&a, ml + 72 * e);
    fp12_load_u64(&b, ml + 72 * (e + step));
    fp12_mul(&a, &a, &b);
    fp12_store_u64(ml + 72 * e, &a);
}

// ok[g] = FinalExponentiation(prod[g]) == 1   (pairing.go:143-146)
__global__ void __launch_bounds__(PAIRING_BLOCK, PAIRING_MIN_BLOCKS) k_final_exp_is_one(const uint64_t *__restrict__ prod, size_t n,
                                                                      uint8_t *__restrict__ ok) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp12 f;
    fp12_load_u64(&f, prod + 72 * i);
    bool good = final_exp_one(&f, &f);
    ok[i] = (good && fp12_is_one(&f)) ? 1 : 0;
}

// group offsets of ONE group of n values, written on the device (no host staging, nothing to wait for)
__global__ void k_one_group(uint32_t *off, uint32_t n) { off[0] = 0; off[1] = n; }

// ok[i] &= (fe[i] == 1)   (the Equals(FQ12One) of pairing.go:146 after a separate final-exponentiation pass)
__global__ void k_fp12_is_one(const uint64_t *__restrict__ fe, size_t n, uint8_t *__restrict__ ok) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t one[12] = {B381_ONE_LIMBS};
    const uint64_t *p = fe + 72 * i;
    uint64_t d = 0;
    for (int k = 0; k < 6; k++) d |= p[k] ^ ((uint64_t)one[2 * k] | ((uint64_t)one[2 * k + 1] << 32));
    for (int k = 6; k < 72; k++) d |= p[k];
    ok[i] = (ok[i] && d == 0) ? 1 : 0;
}


// ---------------------------------------------------------------------------------------------
// kernels: four lanes per pairing (csrc/quad.cuh), persistent -- every warp walks the batch in rounds of 8 pairings
// ---------------------------------------------------------------------------------------------
#ifndef QUAD_BLOCK
#define QUAD_BLOCK 64          // two warps = 16 pairings per block and round
#endif
#ifndef QUAD_MIN_BLOCKS
#define QUAD_MIN_BLOCKS 7      // 14 warps per SM at <= 144 registers: 2^16 pairings = 3.95 rounds of 148 x 7 x 16
#endif
// unit of the calling lane in round `it`, warp-uniform activity; returns false when the whole warp is past the end
__device__ __forceinline__ bool quad_unit(size_t it, size_t n, size_t &unit, bool &active) {
    const size_t per_block = QUAD_BLOCK / 4;
    const size_t warp_first = (it * gridDim.x + blockIdx.x) * per_block + (threadIdx.x >> 5) * 8;
    if (warp_first >= n) return false;
    unit = warp_first + ((threadIdx.x & 31u) >> 2);
    active = unit < n;
    if (!active) unit = n - 1;                      // idle quads of the last warp recompute the last unit (shuffles need them)
    return true;
}
// out[i] = MillerLoop of the NP pairs (p, q)[NP i .. NP i + NP)   (pairing.go:16-75 fused with g2.go:650-801)
template <int NP>
__global__ void __launch_bounds__(QUAD_BLOCK, QUAD_MIN_BLOCKS) k_quad_miller_loop(const g1_affine_pod *__restrict__ p,
                                                                                   const g2_affine_pod *__restrict__ q, size_t n,
                                                                                   uint64_t *__restrict__ out) {
    for (size_t it = 0;; it++) {
        size_t u; bool active;
        if (!quad_unit(it, n, u, active)) break;
        quad::q6 F; quad::qpair S[NP]; quad::qlive lv[NP];
#pragma unroll
        for (int k = 0; k < NP; k++) { quad::qpair_load(&S[k], p + NP * u + k, q + NP * u + k); quad::qpair_live(lv[k], p + NP * u + k, q + NP * u + k); }
        quad::q_miller_loop<NP>(&F, S, lv);
        if (active) quad::q12_store(out + 72 * u, &F);
    }
}
// out[i] = FinalExponentiation(in[i])   (pairing.go:79-129); in place allowed; MODE 1: ok[i] = (result == 1) and no value is written
template <int MODE>
__global__ void __launch_bounds__(QUAD_BLOCK, QUAD_MIN_BLOCKS) k_quad_final_exp(const uint64_t *in, size_t n, uint64_t *out, uint8_t *ok) {
    for (size_t it = 0;; it++) {
        size_t u; bool active;
        if (!quad_unit(it, n, u, active)) break;
        quad::q6 F; bool good[1], isone[1];
        quad::q12_load(&F, in + 72 * u);
        quad::q_final_exp(&F, good);
        if (MODE == 1) {
            quad::q12_is_one(isone, &F);
            if (active && (threadIdx.x & 3u) == 0) ok[u] = (good[0] && isone[0]) ? 1 : 0;
        } else {
            if (!good[0]) quad::q12_set_one(&F);           // f == 0: report 1 with ok = 0 like the other paths
            if (active) {
                quad::q12_store(out + 72 * u, &F);
                if (ok && (threadIdx.x & 3u) == 0) ok[u] = good[0] ? 1 : 0;
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// kernels: two lanes per pairing (csrc/duo.cuh), persistent -- every warp walks the batch in rounds of 16 pairings
// ---------------------------------------------------------------------------------------------
#ifndef DUO_BLOCK
#define DUO_BLOCK 64           // two warps = 32 pairings per block and round
#endif
#ifndef DUO_MIN_BLOCKS
#define DUO_MIN_BLOCKS 7       // 14 warps per SM: 2^16 pairings = 2048 block-rounds = 1.98 rounds of 148 x 7 blocks
#endif
__device__ __forceinline__ bool duo_unit(size_t it, size_t n, size_t &unit, bool &active) {
    const size_t per_block = DUO_BLOCK / 2;
    const size_t warp_first = (it * gridDim.x + blockIdx.x) * per_block + (threadIdx.x >> 5) * 16;
    if (warp_first >= n) return false;
    unit = warp_first + ((threadIdx.x & 31u) >> 1);
    active = unit < n;
    if (!active) unit = n - 1;                      // idle pairs of the last warp recompute the last unit (shuffles name all lanes)
    return true;
}
template <int NP>
__global__ void __launch_bounds__(DUO_BLOCK, DUO_MIN_BLOCKS) k_duo_miller_loop(const g1_affine_pod *__restrict__ p,
                                                                                const g2_affine_pod *__restrict__ q, size_t n,
                                                                                uint64_t *__restrict__ out) {
    for (size_t it = 0;; it++) {
        size_t u; bool active;
        if (!duo_unit(it, n, u, active)) break;
        duo::d12 F; duo::dpair S[NP];
#pragma unroll
        for (int k = 0; k < NP; k++) duo::dpair_load(&S[k], p + NP * u + k, q + NP * u + k);
        duo::d_miller_loop<NP>(&F, S);
        if (active) duo::d12_store(out + 72 * u, &F);
    }
}
// MODE 0: out[i] = FinalExponentiation(in[i]), ok[i] = (in[i] != 0) if ok;  MODE 1: ok[i] = (FinalExponentiation(in[i]) == 1), no value
template <int MODE>
__global__ void __launch_bounds__(DUO_BLOCK, DUO_MIN_BLOCKS) k_duo_final_exp(const uint64_t *in, size_t n, uint64_t *out, uint8_t *ok) {
    for (size_t it = 0;; it++) {
        size_t u; bool active;
        if (!duo_unit(it, n, u, active)) break;
        duo::d12 F; duo::dflag good, isone;
        duo::d12_load(&F, in + 72 * u);
        duo::d_final_exp(&F, good);
        if (MODE == 1) {
            duo::d12_is_one(isone, &F);
            if (active && (threadIdx.x & 1u) == 0) ok[u] = (good.on[0] && isone.on[0]) ? 1 : 0;
        } else {
            if (!good.on[0]) duo::d12_set_one(&F);
            if (active) {
                duo::d12_store(out + 72 * u, &F);
                if (ok && (threadIdx.x & 1u) == 0) ok[u] = good.on[0] ? 1 : 0;
            }
        }
    }
}

// Roofline denominator: sustained issue rate of IMAD.WIDE.U32 (the 32x32->64 multiply-accumulate
// every Fq multiplication is made of).  8 independent chains per thread; each thread executes
// iters * 8 wide MACs.
__global__ void k_imad_probe(uint32_t *out, int iters) {
    uint32_t b = blockIdx.x * 40503u + threadIdx.x * 2654435761u + 7u;
    uint64_t acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = (uint64_t)(b + j) * 0x9E3779B97F4A7C15ull;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            uint32_t x = (uint32_t)acc[(j + 1) & 7];
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(x), "r"(b));
        }
    }
    uint64_t r = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) r ^= acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)r ^ (uint32_t)(r >> 32);
}

// The same roofline measured through the engine's own multiplier: a register-only dependent chain of
// fp_mul (300 IMAD.WIDE.U32[.X] each).  This is the number the pairing kernels are compared with:
// the carry-chained form reaches 32 lanes/clk/SM, the plain 64-bit accumulate of k_imad_probe less.
__global__ void __launch_bounds__(256) k_fpmul_probe(uint32_t *out, int iters) {
    fp x, y;
#pragma unroll
    for (int j = 0; j < 12; j++) { x.l[j] = (blockIdx.x * 977u + threadIdx.x * 131u + j) & 0x0fffffffu; y.l[j] = (threadIdx.x * 7919u + j * 13u) & 0x0fffffffu; }
#pragma unroll 1
    for (int i = 0; i < iters; i++) fp_mul_inl(x, x, y);
    uint32_t r = 0;
#pragma unroll
    for (int j = 0; j < 12; j++) r ^= x.l[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
struct b381_ctx {
    int device;
    int sms;
    cudaStream_t own_stream;
    cudaStream_t stream;
    cudaStream_t side_stream;        // a second stream of the ctx's own for work that is independent of the main chain (fork / join by events)
    cudaEvent_t side_fork, side_join;
    cudaEvent_t pipe_in[2], pipe_done[2], pipe_out[2];   // double-buffered host streaming (b381_pairing_batch_stream)
    uint64_t launches;
    char err[256];
    // grow-only device scratch
    void *scratch[40];
    size_t scratch_bytes[40];
    // warp-cooperative VM programs resident on the device (csrc/vm.cuh)
    struct { uint4 *code, *consts; int lanes, nsteps, nslots, spill_fq; } vm[3];
    int path;                // -1 by batch size, 0 one pairing per thread, 1 warp-cooperative VM, 2 four lanes per pairing
    int vm_split;
    int rlc_bits;            // bit length the caller promises for the weights of the random-linear-combination checks
    int prepared_attest;     // attestation batches prepare repeated message points (B381_PREPARED=0 disables: A/B measurements)
};
enum { VM_ML1 = 0, VM_FE_A = 1, VM_FE_C = 2 };

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            snprintf(ctx->err, sizeof ctx->err, "%s failed: %s", #call, cudaGetErrorString(e_));      \
            return e_ == cudaErrorMemoryAllocation ? B381_ERR_NOMEM : B381_ERR_CUDA;                  \
        }                                                                                             \
    } while (0)

// Grow-only device scratch, one slot per role so that nested entry points never alias each other's buffers:
//    0  Miller values (n x Fq12)            1  group products (ngroups x Fq12)      2, 3  staged / assembled pairs (G1, G2)
//    4  partial sums of the point-sum kernels   5  MSM work area (counts, offsets, indices, chunk / segment / window sums)
//    6  group offsets of assembled checks   7  staged sum / MSM results             8-11  VM: norms, spill, input copy, ok flags
//   12-14  codec / hash host staging (encoded bytes or messages, decoded points, status or scalars or offsets)
//   15-18  wire-level verify: decoded keys, decoded signatures, message points, status + validity bytes
//   19, 20  wire-level verify host staging (inputs, verdicts)                       21  largest group size (tree product)
//   22  group offsets of the random-linear-combination check                        23  its G2 sum + "any invalid" flag
//   24  validity bytes of an attestation batch        25, 26  group offsets / verdicts of b381_pairing_product_is_one        27, 28  prepared G2 points
//   32-35  attestation-level random-linear-combination check: weighted keys, group keys as scalars, per-message pairs (G1, G2)
//   36  work area of the G2 MSM that runs on the side stream        37-39  double buffers of b381_pairing_batch_stream (G1, G2, Fq12)
static int scratch_get(b381_ctx *ctx, int slot, size_t bytes, void **out) {
    if (ctx->scratch_bytes[slot] < bytes) {
        if (ctx->scratch[slot]) {
            CK(cudaStreamSynchronize(ctx->stream));
            CK(cudaFree(ctx->scratch[slot]));
            ctx->scratch[slot] = nullptr; ctx->scratch_bytes[slot] = 0;
        }
        CK(cudaMalloc(&ctx->scratch[slot], bytes));
        ctx->scratch_bytes[slot] = bytes;
    }
    *out = ctx->scratch[slot];
    return B381_OK;
}

static inline unsigned grid_for(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }
// Which schedule runs a batch of n units when the caller did not force one (measured on B200, tools/path_sweep.py,
// profiles/r02_path_sweep.json: ms per b381_pairing_batch_dev call)
//      n      VM    thread    duo    quad
//     64     7.3    18.6    10.2     7.3        few units: only latency counts -- four lanes per pairing (and the VM's
//   4096     9.4    18.7    10.3     7.7        eight-lane split form up to 592 units) finish first
//   8192    11.3    18.7    10.4    10.2
//  16384    22.7    20.2    13.8    17.8        half a wave of threads: two lanes per pairing fill the SMs
//  32768    45.3    27.2    25.1    36.1
//  49152      -     37.2    39.8    54.5        a full wave of threads: one pairing per thread executes the fewest instructions
//  65536      -     46.6    50.8    72.8
#define QUAD_MAX_UNITS 10240
#define DUO_MAX_UNITS 40960
static inline int path_for(const b381_ctx *ctx, size_t n) {
    if (ctx->path >= 0) return ctx->path;
    return n <= QUAD_MAX_UNITS ? B381_PATH_QUAD : n <= DUO_MAX_UNITS ? B381_PATH_DUO : B381_PATH_THREAD;
}
static inline bool vm_for(const b381_ctx *ctx, size_t n) { return path_for(ctx, n) == B381_PATH_VM; }
static inline bool quad_for(const b381_ctx *ctx, size_t n) { return path_for(ctx, n) == B381_PATH_QUAD; }
static inline bool duo_for(const b381_ctx *ctx, size_t n) { return path_for(ctx, n) == B381_PATH_DUO; }
// persistent grid: the block-rounds of the batch spread evenly over the fewest rounds of sms x min_blocks resident blocks
static inline unsigned lane_grid(const b381_ctx *ctx, size_t n, size_t per_block, size_t min_blocks) {
    size_t want = (n + per_block - 1) / per_block, cap = (size_t)ctx->sms * min_blocks;
    size_t rounds = (want + cap - 1) / cap;
    return (unsigned)(rounds ? (want + rounds - 1) / rounds : 1);
}
static inline unsigned quad_grid(const b381_ctx *ctx, size_t n) { return lane_grid(ctx, n, QUAD_BLOCK / 4, QUAD_MIN_BLOCKS); }
static inline unsigned duo_grid(const b381_ctx *ctx, size_t n) { return lane_grid(ctx, n, DUO_BLOCK / 2, DUO_MIN_BLOCKS); }
static inline size_t vm_smem_bytes(int lanes, int nslots) { return (size_t)VM2_WARPS * (32 / lanes) * nslots * 96; }
#define VM_SPLIT_MAX_UNITS 592     // up to one warp of 4 units per SM: below this only latency matters -> two lanes per Fq2 operation

// run VM program `which` over n units; seg[i] = (base, stride) of the per-unit global areas
static int vm_run(b381_ctx *ctx, int which, const vm_seg seg[4], size_t n, const unsigned char *flag_a, size_t fsa,
                  const unsigned char *flag_b, size_t fsb, unsigned char *ok) {
    vm2_args A;
    for (int i = 0; i < 4; i++) A.seg[i] = seg[i];
    A.seg[4].base = (unsigned char *)ctx->vm[which].consts; A.seg[4].stride = 0;
    A.code = ctx->vm[which].code;
    A.nsteps = ctx->vm[which].nsteps; A.nslots = ctx->vm[which].nslots; A.n = n;
    A.flag_a = flag_a; A.flag_b = flag_b; A.flag_stride_a = fsa; A.flag_stride_b = fsb; A.ok = ok;
    int lanes = ctx->vm[which].lanes, upb = VM2_WARPS * (32 / lanes);
    if ((n <= VM_SPLIT_MAX_UNITS && ctx->vm_split) || ctx->vm_split == 2)
        k_vm2<4, 2><<<grid_for(n, upb / 2), VM2_WARPS * 32, vm_smem_bytes(2 * lanes, A.nslots), ctx->stream>>>(A);
    else
        k_vm2<4, 1><<<grid_for(n, upb), VM2_WARPS * 32, vm_smem_bytes(lanes, A.nslots), ctx->stream>>>(A);
    ctx->launches++;
    CK(cudaGetLastError());
    return B381_OK;
}

extern "C" {

int b381_init(int device, b381_ctx **out) {
    if (!out) return B381_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return B381_ERR_NO_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return B381_ERR_NO_DEVICE;
    b381_ctx *ctx = new (std::nothrow) b381_ctx();
    if (!ctx) return B381_ERR_NOMEM;
    memset(ctx, 0, sizeof *ctx);
    ctx->device = device;
    cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, device);
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return B381_ERR_CUDA; }
    ctx->stream = ctx->own_stream;
    if (cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->side_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->side_join, cudaEventDisableTiming) != cudaSuccess) { b381_free(ctx); return B381_ERR_CUDA; }
    for (int i = 0; i < 2; i++)
        if (cudaEventCreateWithFlags(&ctx->pipe_in[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->pipe_done[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->pipe_out[i], cudaEventDisableTiming) != cudaSuccess) { b381_free(ctx); return B381_ERR_CUDA; }
    // the tower state of a pairing lives in local memory: prefer L1 over shared memory
    if (cudaFuncSetAttribute(k_miller_loop, cudaFuncAttributePreferredSharedMemoryCarveout, 0) != cudaSuccess ||
        cudaFuncSetAttribute(k_miller_loop2, cudaFuncAttributePreferredSharedMemoryCarveout, 0) != cudaSuccess ||
        cudaFuncSetAttribute(k_final_exp, cudaFuncAttributePreferredSharedMemoryCarveout, 0) != cudaSuccess ||
        cudaFuncSetAttribute(k_final_exp_is_one, cudaFuncAttributePreferredSharedMemoryCarveout, 0) != cudaSuccess ||
        cudaFuncSetAttribute(k_group_product, cudaFuncAttributePreferredSharedMemoryCarveout, 0) != cudaSuccess) {
        b381_free(ctx);
        return B381_ERR_CUDA;
    }
    // upload the VM programs and opt in to the shared memory they need
    static_assert(sizeof(vm_program_images) / sizeof(vm_program_images[0]) == 3, "ml1, fe_a, fe_c");
    size_t max_smem = 0;
    for (int i = 0; i < 3; i++) {
        const vm_program_image &im = vm_program_images[i];
        size_t cb = (size_t)im.nsteps * im.lanes * 64, kb = (size_t)im.nconsts * 96;
        if (cudaMalloc(&ctx->vm[i].code, cb) != cudaSuccess || cudaMalloc(&ctx->vm[i].consts, kb) != cudaSuccess ||
            cudaMemcpy(ctx->vm[i].code, im.code, cb, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(ctx->vm[i].consts, im.consts, kb, cudaMemcpyHostToDevice) != cudaSuccess) { b381_free(ctx); return B381_ERR_CUDA; }
        ctx->vm[i].lanes = im.lanes; ctx->vm[i].nsteps = im.nsteps; ctx->vm[i].nslots = im.nslots; ctx->vm[i].spill_fq = im.spill_fq;
        size_t sm = vm_smem_bytes(im.lanes, im.nslots);
        if (sm > max_smem) max_smem = sm;
    }
    if (max_smem < 200 * 1024) max_smem = 200 * 1024;
    if (cudaFuncSetAttribute(k_vm2<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem) != cudaSuccess) { b381_free(ctx); return B381_ERR_CUDA; }
    // Two schedules of the same arithmetic: the warp-cooperative VM (8 pairings per warp, state in shared memory:
    // 2-3x lower latency, no DRAM traffic, fills the GPU from ~8 k pairings) and one pairing per thread (higher
    // throughput once ~65 k threads are resident).  Default: by batch size; B381_VM=0/1 forces one of them for A/B measurements.
    const char *ev = getenv("B381_VM");
    ctx->path = ev ? (ev[0] == '0' ? 0 : 1) : -1;
    ctx->rlc_bits = 255;
    const char *epr = getenv("B381_PREPARED");
    ctx->prepared_attest = epr ? (epr[0] != '0') : 1;
    const char *ep = getenv("B381_PATH");          // thread | vm | quad | auto (same as b381_set_kernel_path)
    if (ep) ctx->path = ep[0] == 't' ? 0 : ep[0] == 'v' ? 1 : ep[0] == 'q' ? 2 : ep[0] == 'd' ? 3 : -1;
    const char *es = getenv("B381_VM_SPLIT");       // 0 disables the two-lanes-per-operation latency form (A/B measurements)
    ctx->vm_split = es ? (es[0] == '2' ? 2 : es[0] != '0') : 1;      // 2: the two-lane form at every batch size (measurements)
    *out = ctx;
    return B381_OK;
}

void b381_free(b381_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 40; i++) if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    for (int i = 0; i < 3; i++) { if (ctx->vm[i].code) cudaFree(ctx->vm[i].code); if (ctx->vm[i].consts) cudaFree(ctx->vm[i].consts); }
    if (ctx->side_stream) { cudaStreamSynchronize(ctx->side_stream); cudaStreamDestroy(ctx->side_stream); }
    if (ctx->side_fork) cudaEventDestroy(ctx->side_fork);
    if (ctx->side_join) cudaEventDestroy(ctx->side_join);
    for (int i = 0; i < 2; i++) {
        if (ctx->pipe_in[i]) cudaEventDestroy(ctx->pipe_in[i]);
        if (ctx->pipe_done[i]) cudaEventDestroy(ctx->pipe_done[i]);
        if (ctx->pipe_out[i]) cudaEventDestroy(ctx->pipe_out[i]);
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char *b381_last_error(const b381_ctx *ctx) { return ctx ? ctx->err : "null ctx"; }

int b381_set_stream(b381_ctx *ctx, void *cuda_stream) {
    if (!ctx) return B381_ERR_ARG;
    // work queued on the stream being left may still use the grow-only scratch (scratch_get frees after synchronising the
    // CURRENT stream only): drain it before switching
    if ((cudaStream_t)cuda_stream != ctx->stream) CK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = (cudaStream_t)cuda_stream;   // NULL is the legacy default stream, as everywhere in CUDA
    return B381_OK;
}
int b381_use_own_stream(b381_ctx *ctx) {
    if (!ctx) return B381_ERR_ARG;
    if (ctx->stream != ctx->own_stream) CK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = ctx->own_stream;
    return B381_OK;
}
int b381_set_kernel_path(b381_ctx *ctx, int path) {
    if (!ctx || path < B381_PATH_AUTO || path > B381_PATH_DUO) return B381_ERR_ARG;
    ctx->path = path;
    return B381_OK;
}
int b381_set_rlc_weight_bits(b381_ctx *ctx, int bits) {
    if (!ctx || bits < 8 || bits > 255) return B381_ERR_ARG;
    ctx->rlc_bits = bits;
    return B381_OK;
}
int b381_sync(b381_ctx *ctx) {
    if (!ctx) return B381_ERR_ARG;
    CK(cudaStreamSynchronize(ctx->stream));
    return B381_OK;
}
uint64_t b381_launch_count(const b381_ctx *ctx) { return ctx ? ctx->launches : 0; }

int b381_dev_alloc(b381_ctx *ctx, size_t bytes, void **dptr) {
    if (!ctx || !dptr) return B381_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMalloc(dptr, bytes ? bytes : 1));
    return B381_OK;
}
int b381_dev_free(b381_ctx *ctx, void *dptr) {
    if (!ctx) return B381_ERR_ARG;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaFree(dptr));
    return B381_OK;
}
int b381_h2d(b381_ctx *ctx, void *dptr, const void *host, size_t bytes) {
    if (!ctx || (bytes && (!dptr || !host))) return B381_ERR_ARG;
    CK(cudaMemcpyAsync(dptr, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return B381_OK;
}
int b381_d2h(b381_ctx *ctx, void *host, const void *dptr, size_t bytes) {
    if (!ctx || (bytes && (!dptr || !host))) return B381_ERR_ARG;
    CK(cudaMemcpyAsync(host, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B381_OK;
}

// ---- measurement ---------------------------------------------------------------------------------
int b381_imad_probe_dev(b381_ctx *ctx, uint32_t *d_out, int blocks, int threads, int iters) {
    if (!ctx || !d_out || blocks <= 0 || threads <= 0 || threads > 1024 || iters <= 0) return B381_ERR_ARG;
    k_imad_probe<<<blocks, threads, 0, ctx->stream>>>(d_out, iters);
    ctx->launches++;
    CK(cudaGetLastError());
    return B381_OK;
}

int b381_fpmul_probe_dev(b381_ctx *ctx, uint32_t *d_out, int blocks, int threads, int iters) {
    if (!ctx || !d_out || blocks <= 0 || threads <= 0 || threads > 256 || iters <= 0) return B381_ERR_ARG;
    k_fpmul_probe<<<blocks, threads, 0, ctx->stream>>>(d_out, iters);
    ctx->launches++;
    CK(cudaGetLastError());
    return B381_OK;
}

// test hook: run an arbitrary VM program (host-side encoded image) over n units
int b381_vm_exec_dev(b381_ctx *ctx, const void *code, int lanes, int nsteps, int nslots, const void *consts, int nconsts,
                     void *const d_seg[4], const size_t stride[4], size_t n) {
    if (!ctx || !code || lanes != 4 || nsteps <= 0 || nslots <= 0 || !d_seg || !stride) return B381_ERR_ARG;
    uint4 *dcode = nullptr, *dconst = nullptr;
    size_t cb = (size_t)nsteps * lanes * 64, kb = (size_t)(nconsts > 0 ? nconsts : 1) * 96;
    CK(cudaMalloc(&dcode, cb));
    if (cudaMalloc(&dconst, kb) != cudaSuccess) { cudaFree(dcode); return B381_ERR_NOMEM; }
    if (cudaMemcpy(dcode, code, cb, cudaMemcpyHostToDevice) != cudaSuccess ||
        (nconsts > 0 && cudaMemcpy(dconst, consts, kb, cudaMemcpyHostToDevice) != cudaSuccess)) {
        cudaFree(dcode); cudaFree(dconst);
        snprintf(ctx->err, sizeof ctx->err, "vm_exec: upload failed");
        return B381_ERR_CUDA;
    }
    vm2_args A;
    for (int i = 0; i < 4; i++) { A.seg[i].base = (unsigned char *)d_seg[i]; A.seg[i].stride = stride[i]; }
    A.seg[4].base = (unsigned char *)dconst; A.seg[4].stride = 0;
    A.code = dcode; A.nsteps = nsteps; A.nslots = nslots; A.n = n;
    A.flag_a = A.flag_b = nullptr; A.flag_stride_a = A.flag_stride_b = 0; A.ok = nullptr;
    size_t sm = vm_smem_bytes(lanes, nslots);
    if (sm > 200 * 1024) { cudaFree(dcode); cudaFree(dconst); return B381_ERR_ARG; }
    if (cudaFuncSetAttribute(k_vm2<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
        cudaFree(dcode); cudaFree(dconst);
        return B381_ERR_CUDA;
    }
    k_vm2<4, 1><<<grid_for(n, VM2_WARPS * (32 / lanes)), VM2_WARPS * 32, sm, ctx->stream>>>(A);
    ctx->launches++;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dcode); cudaFree(dconst);
    if (e != cudaSuccess) { snprintf(ctx->err, sizeof ctx->err, "vm_exec: %s", cudaGetErrorString(e)); return B381_ERR_CUDA; }
    return B381_OK;
}

// test hook (csrc/testops.cuh): one field / tower / group-law operation of the DEVICE build over n host-resident operand pairs
int b381_test_op(b381_ctx *ctx, int family, int op, uint64_t arg, const void *a, const void *b, void *out, void *out2, uint8_t *ok, size_t n) {
    static const size_t in_sz[8] = {48, 96, 288, 576, 576, sizeof(b381_g1_affine), sizeof(b381_g2_affine), 576};
    static const size_t out_sz[8] = {48, 96, 288, 576, 576, sizeof(b381_g1_jac), sizeof(b381_g2_jac), 576};
    if (!ctx || family < 0 || family > 7 || (n && (!a || !b || !out))) return B381_ERR_ARG;
    if (!n) return B381_OK;
    CK(cudaSetDevice(ctx->device));
    void *da, *db, *dout, *dok;
    int rc = scratch_get(ctx, 12, n * in_sz[family], &da); if (rc) return rc;
    rc = scratch_get(ctx, 13, n * in_sz[family], &db); if (rc) return rc;
    rc = scratch_get(ctx, 14, n * (out_sz[family] + 96), &dout); if (rc) return rc;
    rc = scratch_get(ctx, 11, n, &dok); if (rc) return rc;
    uint64_t *dout2 = (uint64_t *)((char *)dout + n * out_sz[family]);
    CK(cudaMemcpyAsync(da, a, n * in_sz[family], cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(db, b, n * in_sz[family], cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(dok, 1, n, ctx->stream));
    CK(cudaMemsetAsync(dout2, 0, n * 96, ctx->stream));
    const uint64_t *ua = (const uint64_t *)da, *ub = (const uint64_t *)db;
    uint64_t *uo = (uint64_t *)dout;
    switch (family) {
        case 0: k_test_fp<<<grid_for(n, 64), 64, 0, ctx->stream>>>(op, ua, ub, uo, n); break;
        case 1: k_test_fp2<<<grid_for(n, 64), 64, 0, ctx->stream>>>(op, ua, ub, uo, n); break;
        case 2: k_test_fp6<<<grid_for(n, 64), 64, 0, ctx->stream>>>(op, arg, ua, ub, uo, n); break;
        case 3: k_test_fp12<<<grid_for(n, 64), 64, 0, ctx->stream>>>(op, arg, ua, ub, uo, (uint8_t *)dok, n); break;
        case 4: k_test_quad12<<<grid_for(4 * n, 64), 64, 0, ctx->stream>>>(op, arg, ua, ub, uo, dout2, (uint8_t *)dok, n); break;
        case 7: k_test_duo12<<<grid_for(2 * n, 64), 64, 0, ctx->stream>>>(op, arg, ua, ub, uo, (uint8_t *)dok, n); break;
        case 5: k_test_group<FpInl, g1_affine_pod, g1_jac_pod><<<grid_for(n, 64), 64, 0, ctx->stream>>>(op, (const g1_affine_pod *)da, (const g1_affine_pod *)db, (g1_jac_pod *)dout, n); break;
        case 6: k_test_group<Fp2Out, g2_affine_pod, g2_jac_pod><<<grid_for(n, 64), 64, 0, ctx->stream>>>(op, (const g2_affine_pod *)da, (const g2_affine_pod *)db, (g2_jac_pod *)dout, n); break;
    }
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, dout, n * out_sz[family], cudaMemcpyDeviceToHost, ctx->stream));
    if (out2) CK(cudaMemcpyAsync(out2, dout2, n * 96, cudaMemcpyDeviceToHost, ctx->stream));
    if (ok) CK(cudaMemcpyAsync(ok, dok, n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B381_OK;
}

}   // extern "C" (templates below)
// ---- wire formats and scalar multiplication (csrc/codec.cuh) -------------------------
template <class C, class APOD>
static int decompress_dev(b381_ctx *ctx, const uint8_t *d_in, size_t n, int check, APOD *d_out, uint8_t *d_status) {
    if (!ctx || (n && (!d_in || !d_out || !d_status))) return B381_ERR_ARG;
    if (!n) return B381_OK;
    k_decompress<C><<<grid_for(n, 64), 64, 0, ctx->stream>>>(d_in, n, check, (typename C::APOD *)d_out, d_status);
    ctx->launches++;
    CK(cudaGetLastError());
    return B381_OK;
}
template <class C, class APOD>
static int decompress_host(b381_ctx *ctx, const uint8_t *in, size_t n, int check, APOD *out, uint8_t *status) {
    if (!ctx || (n && (!in || !out || !status))) return B381_ERR_ARG;
    if (!n) return B381_OK;
    CK(cudaSetDevice(ctx->device));
    void *di, *dout, *ds;
    int rc = scratch_get(ctx, 12, n * C::BYTES, &di); if (rc) return rc;
    rc = scratch_get(ctx, 13, n * sizeof(APOD), &dout); if (rc) return rc;
    rc = scratch_get(ctx, 14, n, &ds); if (rc) return rc;
    CK(cudaMemcpyAsync(di, in, n * C::BYTES, cudaMemcpyHostToDevice, ctx->stream));
    rc = decompress_dev<C, APOD>(ctx, (const uint8_t *)di, n, check, (APOD *)dout, (uint8_t *)ds); if (rc) return rc;
    CK(cudaMemcpyAsync(out, dout, n * sizeof(APOD), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(status, ds, n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B381_OK;
}
template <class C, class APOD>
static int compress_dev(b381_ctx *ctx, const APOD *d_in, size_t n, uint8_t *d_out) {
    if (!ctx || (n && (!d_in || !d_out))) return B381_ERR_ARG;
    if (!n) return B381_OK;
    k_compress<C><<<grid_for(n, 64), 64, 0, ctx->stream>>>((const typename C::APOD *)d_in, n, d_out);
    ctx->launches++;
    CK(cudaGetLastError());
    return B381_OK;
}
template <class C, class APOD>
static int compress_host(b381_ctx *ctx, const APOD *in, size_t n, uint8_t *out) {
    if (!ctx || (n && (!in || !out))) return B381_ERR_ARG;
    if (!n) return B381_OK;
    CK(cudaSetDevice(ctx->device));
    void *di, *dout;
    int rc = scratch_get(ctx, 12, n * C::BYTES, &dout); if (rc) return rc;
    rc = scratch_get(ctx, 13, n * sizeof(APOD), &di); if (rc) return rc;
    CK(cudaMemcpyAsync(di, in, n * sizeof(APOD), cudaMemcpyHostToDevice, ctx->stream));
    rc = compress_dev<C, APOD>(ctx, (const APOD *)di, n, (uint8_t *)dout); if (rc) return rc;
    CK(cudaMemcpyAsync(out, dout, n * C::BYTES, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B381_OK;
}
template <class C, bool TORSION, class APOD>
static int mul_dev(b381_ctx *ctx, const APOD *d_p, size_t p_stride, const b381_scalar *d_k, size_t k_stride, size_t n, APOD *d_out) {
    if (!ctx || p_stride > 1 || k_stride > 1 || (n && (!d_p || !d_k || !d_out))) return B381_ERR_ARG;
    if (!n) return B381_OK;
    k_point_mul<C, TORSION><<<grid_for(n, 64), 64, 0, ctx->stream>>>((const typename C::APOD *)d_p, p_stride, (const uint64_t *)d_k, k_stride, n,
                                                             (typename C::APOD *)d_out);
    ctx->launches++;
    CK(cudaGetLastError());
    return B381_OK;
}
template <class C, bool TORSION, class APOD>
static int mul_host(b381_ctx *ctx, const APOD *p, size_t p_stride, const b381_scalar *k, size_t k_stride, size_t n, APOD *out) {
    if (!ctx || p_stride > 1 || k_stride > 1 || (n && (!p || !k || !out))) return B381_ERR_ARG;
    if (!n) return B381_OK;
    CK(cudaSetDevice(ctx->device));
    size_t np = p_stride ? n : 1, nk = k_stride ? n : 1;
    void *dp, *dk, *dout;
    int rc = scratch_get(ctx, 12, np * sizeof(APOD), &dp); if (rc) return rc;
    rc = scratch_get(ctx, 13, n * sizeof(APOD), &dout); if (rc) return rc;
    rc = scratch_get(ctx, 14, nk * sizeof(b381_scalar), &dk); if (rc) return rc;
    CK(cudaMemcpyAsync(dp, p, np * sizeof(APOD), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dk, k, nk * sizeof(b381_scalar), cudaMemcpyHostToDevice, ctx->stream));
    rc = mul_dev<C, TORSION, APOD>(ctx, (const APOD *)dp, p_stride, (const b381_scalar *)dk, k_stride, n, (APOD *)dout); if (rc) return rc;
    CK(cudaMemcpyAsync(out, dout, n * sizeof(APOD), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B381_OK;
}
extern "C" {
int b381_g1_decompress_batch(b381_ctx *ctx, const uint8_t *in, size_t n, int check_subgroup, b381_g1_affine *out, uint8_t *status) {
    return decompress_host<G1Codec>(ctx, in, n, check_subgroup, out, status);
}
int b381_g1_decompress_batch_dev(b381_ctx *ctx, const uint8_t *d_in, size_t n, int check_subgroup, b381_g1_affine *d_out, uint8_t *d_status) {
    return decompress_dev<G1Codec>(ctx, d_in, n, check_subgroup, d_out, d_status);
}
int b381_g2_decompress_batch(b381_ctx *ctx, const uint8_t *in, size_t n, int check_subgroup, b381_g2_affine *out, uint8_t *status) {
    return decompress_host<G2Codec>(ctx, in, n, check_subgroup, out, status);
}
int b381_g2_decompress_batch_dev(b381_ctx *ctx, const uint8_t *d_in, size_t n, int check_subgroup, b381_g2_affine *d_out, uint8_t *d_status) {
    return decompress_dev<G2Codec>(ctx, d_in, n, check_subgroup, d_out, d_status);
}
int b381_g1_compress_batch(b381_ctx *ctx, const b381_g1_affine *in, size_t n, uint8_t *out) { return compress_host<G1Codec>(ctx, in, n, out); }
int b381_g1_compress_batch_dev(b381_ctx *ctx, const b381_g1_affine *d_in, size_t n, uint8_t *d_out) { return compress_dev<G1Codec>(ctx, d_in, n, d_out); }
int b381_g2_compress_batch(b381_ctx *ctx, const b381_g2_affine *in, size_t n, uint8_t *out) { return compress_host<G2Codec>(ctx, in, n, out); }
int b381_g2_compress_batch_dev(b381_ctx *ctx, const b381_g2_affine *d_in, size_t n, uint8_t *d_out) { return compress_dev<G2Codec>(ctx, d_in, n, d_out); }
int b381_g1_mul_batch(b381_ctx *ctx, const b381_g1_affine *p, size_t p_stride, const b381_scalar *k, size_t k_stride, size_t n, b381_g1_affine *out) {
    return mul_host<G1Codec, false>(ctx, p, p_stride, k, k_stride, n, out);
}
int b381_g1_mul_batch_dev(b381_ctx *ctx, const b381_g1_affine *d_p, size_t p_stride, const b381_scalar *d_k, size_t k_stride, size_t n, b381_g1_affine *d_out) {
    return mul_dev<G1Codec, false>(ctx, d_p, p_stride, d_k, k_stride, n, d_out);
}
int b381_g2_mul_batch(b381_ctx *ctx, const b381_g2_affine *p, size_t p_stride, const b381_scalar *k, size_t k_stride, size_t n, b381_g2_affine *out) {
    return mul_host<G2Codec, false>(ctx, p, p_stride, k, k_stride, n, out);
}
int b381_g2_mul_batch_dev(b381_ctx *ctx, const b381_g2_affine *d_p, size_t p_stride, const b381_scalar *d_k, size_t k_stride, size_t n, b381_g2_affine *d_out) {
    return mul_dev<G2Codec, false>(ctx, d_p, p_stride, d_k, k_stride, n, d_out);
}
int b381_g1_mul_subgroup_batch(b381_ctx *ctx, const b381_g1_affine *p, size_t p_stride, const b381_scalar *k, size_t k_stride, size_t n, b381_g1_affine *out) {
    return mul_host<G1Codec, true>(ctx, p, p_stride, k, k_stride, n, out);
}
int b381_g1_mul_subgroup_batch_dev(b381_ctx *ctx, const b381_g1_affine *d_p, size_t p_stride, const b381_scalar *d_k, size_t k_stride, size_t n, b381_g1_affine *d_out) {
    return mul_dev<G1Codec, true>(ctx, d_p, p_stride, d_k, k_stride, n, d_out);
}
int b381_g2_mul_subgroup_batch(b381_ctx *ctx, const b381_g2_affine *p, size_t p_stride, const b381_scalar *k, size_t k_stride, size_t n, b381_g2_affine *out) {
    return mul_host<G2Codec, true>(ctx, p, p_stride, k, k_stride, n, out);
}
int b381_g2_mul_subgroup_batch_dev(b381_ctx *ctx, const b381_g2_affine *d_p, size_t p_stride, const b381_scalar *d_k, size_t k_stride, size_t n, b381_g2_affine *d_out) {
    return mul_dev<G2Codec, true>(ctx, d_p, p_stride, d_k, k_stride, n, d_out);
}
int b381_hash_g2_with_domain_batch_dev(b381_ctx *ctx, const uint8_t *d_msg32, const uint8_t *d_domain8, size_t domain_stride, size_t n,
                                       b381_g2_affine *d_out) {
    if (!ctx || domain_stride > 1 || (n && (!d_msg32 || !d_domain8 || !d_out))) return B381_ERR_ARG;
    if (!n) return B381_OK;
    k_hash_g2_with_domain<<<grid_for(n, 64), 64, 0, ctx->stream>>>(d_msg32, d_domain8, domain_stride, n, (g2_affine_pod *)d_out);
    ctx->launches++;
    CK(cudaGetLastError());
    return B381_OK;
}
int b381_hash_g2_with_domain_batch(b381_ctx *ctx, const uint8_t *msg32, const uint8_t *domain8, size_t domain_stride, size_t n,
                                   b381_g2_affine *out) {
    if (!ctx || domain_stride > 1 || (n && (!msg32 || !domain8 || !out))) return B381_ERR_ARG;
    if (!n) return B381_OK;
    CK(cudaSetDevice(ctx->device));
    void *dm, *dd, *dout;
    size_t nd = domain_stride ? n : 1;
    int rc = scratch_get(ctx, 12, n * 32, &dm); if (rc) return rc;
    rc = scratch_get(ctx, 13, n * sizeof(b381_g2_affine), &dout); if (rc) return rc;
    rc = scratch_get(ctx, 14, nd * 8, &dd); if (rc) return rc;
    CK(cudaMemcpyAsync(dm, msg32, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dd, domain8, nd * 8, cudaMemcpyHostToDevice, ctx->stream));
    rc = b381_hash_g2_with_domain_batch_dev(ctx, (const uint8_t *)dm, (const uint8_t *)dd, domain_stride, n, (b381_g2_affine *)dout);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, dout, n * sizeof(b381_g2_affine), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B381_OK;
}

// ---- pairing, device-resident ------------------------------------------------------------------
int b381_miller_loop_batch_dev(b381_ctx *ctx, const b381_g1_affine *d_p, const b381_g2_affine *d_q, size_t n,
                               b381_fp12 *d_out) {
    if (!ctx || (n && (!d_p || !d_q || !d_out))) return B381_ERR_ARG;
    if (!n) return B381_OK;
    if (vm_for(ctx, n)) {
        vm_seg seg[4] = {{(unsigned char *)d_p, sizeof(b381_g1_affine)}, {(unsigned char *)d_q, sizeof(b381_g2_affine)},
                         {(unsigned char *)d_out, sizeof(b381_fp12)}, {nullptr, 0}};
        return vm_run(ctx, VM_ML1, seg, n, (const unsigned char *)d_p + 96, sizeof(b381_g1_affine),
                      (const unsigned char *)d_q + 192, sizeof(b381_g2_affine), nullptr);
    }
    if (duo_for(ctx, n))
        k_duo_miller_loop<1><<<duo_grid(ctx, n), DUO_BLOCK, 0, ctx->stream>>>((const g1_affine_pod *)d_p, (const g2_affine_pod *)d_q, n,
                                                                              (uint64_t *)d_out);
    else if (quad_for(ctx, n))
        k_quad_miller_loop<1><<<quad_grid(ctx, n), QUAD_BLOCK, 0, ctx->stream>>>((const g1_affine_pod *)d_p, (const g2_affine_pod *)d_q, n,
                                                                                (uint64_t *)d_out);
    else
        k_miller_loop<<<grid_for(n, PAIRING_BLOCK), PAIRING_BLOCK, 0, ctx->stream>>>(
            (const g1_affine_pod *)d_p, (const g2_affine_pod *)d_q, n, (uint64_t *)d_out);
    ctx->launches++;
    CK(cudaGetLastError());
    return B381_OK;
}
int b381_final_exp_batch_dev(b381_ctx *ctx, const b381_fp12 *d_in, size_t n, b381_fp12 *d_out, uint8_t *d_ok) {
    if (!ctx || (n && (!d_in || !d_out))) return B381_ERR_ARG;
    if (!n) return B381_OK;
    if (vm_for(ctx, n)) {
        // fe_a: f -> the Fq norm the inversion needs; one thread per element inverts it; fe_c: the rest
        void *nrm, *spill, *fcopy, *dok = d_ok;
        int rc = scratch_get(ctx, 8, n * 2 * sizeof(b381_fp), &nrm);
        if (rc) return rc;
        rc = scratch_get(ctx, 9, n * (size_t)ctx->vm[VM_FE_C].spill_fq * sizeof(b381_fp), &spill);
        if (rc) return rc;
        const b381_fp12 *src = d_in;
        if (d_in == d_out) {                      // in place: the zero test and fe_c read the input after outputs exist
            rc = scratch_get(ctx, 10, n * sizeof(b381_fp12), &fcopy);
            if (rc) return rc;
            CK(cudaMemcpyAsync(fcopy, d_in, n * sizeof(b381_fp12), cudaMemcpyDeviceToDevice, ctx->stream));
            src = (const b381_fp12 *)fcopy;
        }
        if (!dok) { rc = scratch_get(ctx, 11, n, &dok); if (rc) return rc; }
        b381_fp *nbuf = (b381_fp *)nrm, *ninv = nbuf + n;
        vm_seg sa[4] = {{(unsigned char *)src, sizeof(b381_fp12)}, {nullptr, 0}, {(unsigned char *)nbuf, sizeof(b381_fp)}, {nullptr, 0}};
        rc = vm_run(ctx, VM_FE_A, sa, n, nullptr, 0, nullptr, 0, nullptr);
        if (rc) return rc;
        k_fp_inv_batch<<<grid_for(n, 128), 128, 0, ctx->stream>>>((const uint64_t *)nbuf, n, (uint64_t *)ninv);
        ctx->launches++;
        vm_seg sc[4] = {{(unsigned char *)src, sizeof(b381_fp12)}, {(unsigned char *)ninv, sizeof(b381_fp)},
                        {(unsigned char *)d_out, sizeof(b381_fp12)}, {(unsigned char *)spill, (size_t)ctx->vm[VM_FE_C].spill_fq * sizeof(b381_fp)}};
        return vm_run(ctx, VM_FE_C, sc, n, nullptr, 0, nullptr, 0, (unsigned char *)dok);
    }
    if (duo_for(ctx, n))
        k_duo_final_exp<0><<<duo_grid(ctx, n), DUO_BLOCK, 0, ctx->stream>>>((const uint64_t *)d_in, n, (uint64_t *)d_out, d_ok);
    else if (quad_for(ctx, n))
        k_quad_final_exp<0><<<quad_grid(ctx, n), QUAD_BLOCK, 0, ctx->stream>>>((const uint64_t *)d_in, n, (uint64_t *)d_out, d_ok);
    else
        k_final_exp<<<grid_for(n, PAIRING_BLOCK), PAIRING_BLOCK, 0, ctx->stream>>>((const uint64_t *)d_in, n,
                                                                                  (uint64_t *)d_out, d_ok);
    ctx->launches++;
    CK(cudaGetLastError());
    return B381_OK;
}
int b381_pairing_batch_dev(b381_ctx *ctx, const b381_g1_affine *d_p, const b381_g2_affine *d_q, size_t n,
                           b381_fp12 *d_out) {
    if (ctx && n && vm_for(ctx, n)) {
        void *ml;
        int rc = scratch_get(ctx, 10, n * sizeof(b381_fp12), &ml);
        if (rc) return rc;
        rc = b381_miller_loop_batch_dev(ctx, d_p, d_q, n, (b381_fp12 *)ml);
        if (rc) return rc;
        return b381_final_exp_batch_dev(ctx, (const b381_fp12 *)ml, n, d_out, nullptr);
    }
    int rc = b381_miller_loop_batch_dev(ctx, d_p, d_q, n, d_out);
    if (rc) return rc;
    return b381_final_exp_batch_dev(ctx, d_out, n, d_out, nullptr);
}
// ---- prepared G2 points: G2AffineToPrepared once, MillerLoop from the coefficients (g2.go:639-801, pairing.go:4-75) ---------------
int b381_g2_prepare_batch_dev(b381_ctx *ctx, const b381_g2_affine *d_q, size_t n, b381_g2_prepared *d_prep) {
    if (!ctx || (n && (!d_q || !d_prep))) return B381_ERR_ARG;
    if (!n) return B381_OK;
    k_g2_prepare<<<grid_for(n, PAIRING_BLOCK), PAIRING_BLOCK, 0, ctx->stream>>>((const g2_affine_pod *)d_q, n, (g2_prepared_pod *)d_prep);
    ctx->launches++;
    CK(cudaGetLastError());
    return B381_OK;
}
int b381_miller_loop_prepared_batch_dev(b381_ctx *ctx, const b381_g1_affine *d_p, const b381_g2_prepared *d_prep, const uint32_t *d_prep_idx,
                                        size_t n, b381_fp12 *d_out) {
    if (!ctx || (n && (!d_p || !d_prep || !d_out))) return B381_ERR_ARG;
    if (!n) return B381_OK;
    k_miller_loop_prepared<<<grid_for(n, PAIRING_BLOCK), PAIRING_BLOCK, 0, ctx->stream>>>((const g1_affine_pod *)d_p, (const g2_prepared_pod *)d_prep,
                                                                                        d_prep_idx, n, (uint64_t *)d_out);
    ctx->launches++;
    CK(cudaGetLastError());
    return B381_OK;
}
int b381_g2_prepare_batch(b381_ctx *ctx, const b381_g2_affine *q, size_t n, b381_g2_prepared *prep) {
    if (!ctx || (n && (!q || !prep))) return B381_ERR_ARG;
    if (!n) return B381_OK;
    CK(cudaSetDevice(ctx->device));
    void *dq, *dp;
    int rc = scratch_get(ctx, 3, n * sizeof(b381_g2_affine), &dq); if (rc) return rc;
    rc = scratch_get(ctx, 27, n * sizeof(b381_g2_prepared), &dp); if (rc) return rc;
    CK(cudaMemcpyAsync(dq, q, n * sizeof(b381_g2_affine), cudaMemcpyHostToDevice, ctx->stream));
    rc = b381_g2_prepare_batch_dev(ctx, (const b381_g2_affine *)dq, n, (b381_g2_prepared *)dp); if (rc) return rc;
    CK(cudaMemcpyAsync(prep, dp, n * sizeof(b381_g2_prepared), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B381_OK;
}
// out[i] = MillerLoop({p[i], prep[prep_idx ? prep_idx[i] : i]}); nprep = number of prepared points behind `prep`
int b381_miller_loop_prepared_batch(b381_ctx *ctx, const b381_g1_affine *p, const b381_g2_prepared *prep, size_t nprep, const uint32_t *prep_idx,
                                    size_t n, b381_fp12 *out) {
    if (!ctx || (n && (!p || !prep || !out || !nprep)) || (!prep_idx && nprep < n)) return B381_ERR_ARG;
    if (!n) return B381_OK;
    if (prep_idx) for (size_t i = 0; i < n; i++) if (prep_idx[i] >= nprep) return B381_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    void *dp, *dprep, *didx = nullptr, *dout;
    int rc = scratch_get(ctx, 2, n * sizeof(b381_g1_affine), &dp); if (rc) return rc;
    rc = scratch_get(ctx, 27, nprep * sizeof(b381_g2_prepared), &dprep); if (rc) return rc;
    rc = scratch_get(ctx, 0, n * sizeof(b381_fp12), &dout); if (rc) return rc;
    if (prep_idx) { rc = scratch_get(ctx, 28, n * sizeof(uint32_t), &didx); if (rc) return rc; }
    CK(cudaMemcpyAsync(dp, p, n * sizeof(b381_g1_affine), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dprep, prep, nprep * sizeof(b381_g2_prepared), cudaMemcpyHostToDevice, ctx->stream));
    if (prep_idx) CK(cudaMemcpyAsync(didx, prep_idx, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    rc = b381_miller_loop_prepared_batch_dev(ctx, (const b381_g1_affine *)dp, (const b381_g2_prepared *)dprep, (const uint32_t *)didx, n, (b381_fp12 *)dout);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, dout, n * sizeof(b381_fp12), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B381_OK;
}

// prod[g] = product of the Fq12 values ml[group_off[g] .. group_off[g+1]) (ml is overwritten when groups are folded as trees)
static int group_products(b381_ctx *ctx, void *ml, size_t nvals, const uint32_t *d_group_off, size_t ngroups, void *prod) {
    int tree = nvals > 2 * ngroups;               // some group has more than two factors: fold groups as trees
    if (tree) {
        void *mx;
        int rc = scratch_get(ctx, 21, sizeof(uint32_t), &mx);
        if (rc) return rc;
        CK(cudaMemsetAsync(mx, 0, sizeof(uint32_t), ctx->stream));
        k_group_maxsize<<<grid_for(ngroups, 256), 256, 0, ctx->stream>>>(d_group_off, ngroups, (uint32_t *)mx);
        ctx->launches++;
        for (int level = 0; ((size_t)1 << level) < nvals; level++) {
            k_group_tree<<<grid_for(nvals, PAIRING_BLOCK), PAIRING_BLOCK, 0, ctx->stream>>>((uint64_t *)ml, d_group_off, ngroups, nvals,
                                                                                           level, (const uint32_t *)mx);
            ctx->launches++;
        }
    }
    k_group_product<<<grid_for(ngroups, PAIRING_BLOCK), PAIRING_BLOCK, 0, ctx->stream>>>(
        (const uint64_t *)ml, d_group_off, ngroups, (uint64_t *)prod, tree);
    ctx->launches++;
    CK(cudaGetLastError());
    return B381_OK;
}
// ok[g] = FinalExponentiation(prod[g]) == 1 (prod is overwritten)
static int final_exp_is_one(b381_ctx *ctx, void *prod, size_t ngroups, uint8_t *d_ok) {
    if (vm_for(ctx, ngroups)) {          // few groups: the warp-cooperative final exponentiation has 3x lower latency
        int rc = b381_final_exp_batch_dev(ctx, (const b381_fp12 *)prod, ngroups, (b381_fp12 *)prod, d_ok);
        if (rc) return rc;
        k_fp12_is_one<<<grid_for(ngroups, 128), 128, 0, ctx->stream>>>((const uint64_t *)prod, ngroups, d_ok);
        ctx->launches++;
        CK(cudaGetLastError());
        return B381_OK;
    }
    if (duo_for(ctx, ngroups))
        k_duo_final_exp<1><<<duo_grid(ctx, ngroups), DUO_BLOCK, 0, ctx->stream>>>((const uint64_t *)prod, ngroups, nullptr, d_ok);
    else if (quad_for(ctx, ngroups))
        k_quad_final_exp<1><<<quad_grid(ctx, ngroups), QUAD_BLOCK, 0, ctx->stream>>>((const uint64_t *)prod, ngroups, nullptr, d_ok);
    else
        k_final_exp_is_one<<<grid_for(ngroups, PAIRING_BLOCK), PAIRING_BLOCK, 0, ctx->stream>>>((const uint64_t *)prod,
                                                                                              ngroups, d_ok);
    ctx->launches++;
    CK(cudaGetLastError());
    return B381_OK;
}
int b381_pairing_product_is_one_dev(b381_ctx *ctx, const b381_g1_affine *d_p, const b381_g2_affine *d_q,
                                    size_t npairs, const uint32_t *d_group_off, size_t ngroups, uint8_t *d_ok) {
    if (!ctx || (ngroups && (!d_group_off || !d_ok)) || (npairs && (!d_p || !d_q))) return B381_ERR_ARG;
    if (!ngroups) return B381_OK;
    void *ml = nullptr, *prod = nullptr;
    int rc = scratch_get(ctx, 0, (npairs ? npairs : 1) * sizeof(b381_fp12), &ml);
    if (rc) return rc;
    rc = scratch_get(ctx, 1, ngroups * sizeof(b381_fp12), &prod);
    if (rc) return rc;
    rc = b381_miller_loop_batch_dev(ctx, d_p, d_q, npairs, (b381_fp12 *)ml);
    if (rc) return rc;
    rc = group_products(ctx, ml, npairs, d_group_off, ngroups, prod);
    if (rc) return rc;
    return final_exp_is_one(ctx, prod, ngroups, d_ok);
}
// ok[g] = [ FE(ML((p,q)[2g], (p,q)[2g+1])) == 1 ]: the engine's CompareTwoPairings for n checks whose pairs are laid out two by
// two (the verify paths build them that way); large batches use the shared-accumulator Miller loop
static int pairs2_product_is_one(b381_ctx *ctx, const b381_g1_affine *d_p, const b381_g2_affine *d_q, size_t ngroups,
                                 const uint32_t *d_group_off, uint8_t *d_ok) {
    if (!ngroups) return B381_OK;
    // few checks: one unit per PAIR, so that the two Miller loops of a check run side by side (latency), then the group products
    // and one final exponentiation per check; many checks: the shared-accumulator Miller loop (throughput)
    if (vm_for(ctx, 2 * ngroups) || (ctx->path < 0 && 2 * ngroups <= QUAD_MAX_UNITS))
        return b381_pairing_product_is_one_dev(ctx, d_p, d_q, 2 * ngroups, d_group_off, ngroups, d_ok);
    void *prod;
    int rc = scratch_get(ctx, 1, ngroups * sizeof(b381_fp12), &prod);
    if (rc) return rc;
    if (duo_for(ctx, ngroups))
        k_duo_miller_loop<2><<<duo_grid(ctx, ngroups), DUO_BLOCK, 0, ctx->stream>>>((const g1_affine_pod *)d_p, (const g2_affine_pod *)d_q,
                                                                                    ngroups, (uint64_t *)prod);
    else if (quad_for(ctx, ngroups))
        k_quad_miller_loop<2><<<quad_grid(ctx, ngroups), QUAD_BLOCK, 0, ctx->stream>>>((const g1_affine_pod *)d_p, (const g2_affine_pod *)d_q,
                                                                                      ngroups, (uint64_t *)prod);
    else
        k_miller_loop2<<<grid_for(ngroups, PAIRING_BLOCK), PAIRING_BLOCK, 0, ctx->stream>>>((const g1_affine_pod *)d_p, (const g2_affine_pod *)d_q,
                                                                                           ngroups, (uint64_t *)prod);
    ctx->launches++;
    CK(cudaGetLastError());
    return final_exp_is_one(ctx, prod, ngroups, d_ok);
}
// out = prod_i MillerLoop(p[i], q[i]) WITHOUT the final exponentiation: one rank's factor of a product that is finished
// elsewhere (the multi-GPU random-linear-combination check: all-gather of 576-byte partials)
int b381_miller_product_dev(b381_ctx *ctx, const b381_g1_affine *d_p, const b381_g2_affine *d_q, size_t npairs, b381_fp12 *d_out) {
    if (!ctx || !d_out || npairs > 0x7FFFFFF0u || (npairs && (!d_p || !d_q))) return B381_ERR_ARG;
    void *ml = nullptr, *off = nullptr;
    int rc = scratch_get(ctx, 0, (npairs ? npairs : 1) * sizeof(b381_fp12), &ml);
    if (rc) return rc;
    rc = scratch_get(ctx, 6, 2 * sizeof(uint32_t), &off);
    if (rc) return rc;
    k_one_group<<<1, 1, 0, ctx->stream>>>((uint32_t *)off, (uint32_t)npairs);
    ctx->launches++;
    rc = b381_miller_loop_batch_dev(ctx, d_p, d_q, npairs, (b381_fp12 *)ml);
    if (rc) return rc;
    return group_products(ctx, ml, npairs, (const uint32_t *)off, 1, d_out);
}
// *ok = [ FinalExponentiation(prod_i parts[i]) == 1 ]: the finishing step on every rank after the all-gather
int b381_fp12_product_final_exp_is_one_dev(b381_ctx *ctx, const b381_fp12 *d_parts, size_t n, uint8_t *d_ok) {
    if (!ctx || !d_ok || n > 0x7FFFFFF0u || (n && !d_parts)) return B381_ERR_ARG;
    void *ml = nullptr, *prod = nullptr, *off = nullptr;
    int rc = scratch_get(ctx, 0, (n ? n : 1) * sizeof(b381_fp12), &ml);
    if (rc) return rc;
    rc = scratch_get(ctx, 1, sizeof(b381_fp12), &prod);
    if (rc) return rc;
    rc = scratch_get(ctx, 6, 2 * sizeof(uint32_t), &off);
    if (rc) return rc;
    k_one_group<<<1, 1, 0, ctx->stream>>>((uint32_t *)off, (uint32_t)n);
    ctx->launches++;
    if (n) CK(cudaMemcpyAsync(ml, d_parts, n * sizeof(b381_fp12), cudaMemcpyDeviceToDevice, ctx->stream));
    rc = group_products(ctx, ml, n, (const uint32_t *)off, 1, prod);
    if (rc) return rc;
    return final_exp_is_one(ctx, prod, 1, d_ok);
}

// ---- pairing, host buffers (copies inside) -------------------------------------------------------
static int staged_pq(b381_ctx *ctx, const b381_g1_affine *p, const b381_g2_affine *q, size_t n, void **dp, void **dq) {
    int rc = scratch_get(ctx, 2, n * sizeof(b381_g1_affine), dp);
    if (rc) return rc;
    rc = scratch_get(ctx, 3, n * sizeof(b381_g2_affine), dq);
    if (rc) return rc;
    CK(cudaMemcpyAsync(*dp, p, n * sizeof(b381_g1_affine), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(*dq, q, n * sizeof(b381_g2_affine), cudaMemcpyHostToDevice, ctx->stream));
    return B381_OK;
}
static int pairing_host(b381_ctx *ctx, const b381_g1_affine *p, const b381_g2_affine *q, size_t n, b381_fp12 *out,
                        bool final_exp) {
    if (!ctx || (n && (!p || !q || !out))) return B381_ERR_ARG;
    if (!n) return B381_OK;
    CK(cudaSetDevice(ctx->device));
    void *dp, *dq, *dout;
    int rc = staged_pq(ctx, p, q, n, &dp, &dq);
    if (rc) return rc;
    rc = scratch_get(ctx, 0, n * sizeof(b381_fp12), &dout);
    if (rc) return rc;
    rc = final_exp ? b381_pairing_batch_dev(ctx, (b381_g1_affine *)dp, (b381_g2_affine *)dq, n, (b381_fp12 *)dout)
                   : b381_miller_loop_batch_dev(ctx, (b381_g1_affine *)dp, (b381_g2_affine *)dq, n, (b381_fp12 *)dout);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, dout, n * sizeof(b381_fp12), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B381_OK;
}
int b381_pairing_batch(b381_ctx *ctx, const b381_g1_affine *p, const b381_g2_affine *q, size_t n, b381_fp12 *out) {
    return pairing_host(ctx, p, q, n, out, true);
}
int b381_miller_loop_batch(b381_ctx *ctx, const b381_g1_affine *p, const b381_g2_affine *q, size_t n, b381_fp12 *out) {
    return pairing_host(ctx, p, q, n, out, false);
}
// A stream of batches from host memory: out[i] = Pairing(p[i], q[i]) for n pairs, computed `batch` pairs at a time with the
// copies of neighbouring batches overlapped with the kernels (two device buffers per array; the inputs of batch k + 1 go up and
// the results of batch k - 1 come down on the ctx's side stream while batch k computes).  With page-locked host memory the copies
// cost nothing after the first batch in and the last batch out; the values are those of b381_pairing_batch.
int b381_pairing_batch_stream(b381_ctx *ctx, const b381_g1_affine *p, const b381_g2_affine *q, size_t n, size_t batch, b381_fp12 *out) {
    if (!ctx || !batch || (n && (!p || !q || !out))) return B381_ERR_ARG;
    if (!n) return B381_OK;
    if (batch > n) batch = n;
    CK(cudaSetDevice(ctx->device));
    void *dp, *dq, *dout;
    int rc = scratch_get(ctx, 37, 2 * batch * sizeof(b381_g1_affine), &dp); if (rc) return rc;
    rc = scratch_get(ctx, 38, 2 * batch * sizeof(b381_g2_affine), &dq); if (rc) return rc;
    rc = scratch_get(ctx, 39, 2 * batch * sizeof(b381_fp12), &dout); if (rc) return rc;
    b381_g1_affine *dP[2] = {(b381_g1_affine *)dp, (b381_g1_affine *)dp + batch};
    b381_g2_affine *dQ[2] = {(b381_g2_affine *)dq, (b381_g2_affine *)dq + batch};
    b381_fp12 *dO[2] = {(b381_fp12 *)dout, (b381_fp12 *)dout + batch};
    const size_t K = (n + batch - 1) / batch;
    cudaStream_t M = ctx->stream, S = ctx->side_stream;
    auto len = [&](size_t k) { return k + 1 < K ? batch : n - k * batch; };
    auto upload = [&](size_t k) -> int {
        const int b = (int)(k & 1);
        CK(cudaMemcpyAsync(dP[b], p + k * batch, len(k) * sizeof(b381_g1_affine), cudaMemcpyHostToDevice, S));
        CK(cudaMemcpyAsync(dQ[b], q + k * batch, len(k) * sizeof(b381_g2_affine), cudaMemcpyHostToDevice, S));
        CK(cudaEventRecord(ctx->pipe_in[b], S));
        return B381_OK;
    };
    CK(cudaEventRecord(ctx->side_fork, M));                    // the side stream starts after whatever the main stream holds
    CK(cudaStreamWaitEvent(S, ctx->side_fork, 0));
    rc = upload(0); if (rc) return rc;
    for (size_t k = 0; k < K; k++) {
        const int b = (int)(k & 1);
        if (k + 1 < K) {
            if (k >= 1) CK(cudaStreamWaitEvent(S, ctx->pipe_done[b ^ 1], 0));   // batch k - 1 has read the buffers batch k + 1 goes into
            rc = upload(k + 1); if (rc) return rc;
        }
        CK(cudaStreamWaitEvent(M, ctx->pipe_in[b], 0));
        if (k >= 2) CK(cudaStreamWaitEvent(M, ctx->pipe_out[b], 0));            // the results of batch k - 2 have left this buffer
        rc = b381_pairing_batch_dev(ctx, dP[b], dQ[b], len(k), dO[b]); if (rc) return rc;
        CK(cudaEventRecord(ctx->pipe_done[b], M));
        CK(cudaStreamWaitEvent(S, ctx->pipe_done[b], 0));
        CK(cudaMemcpyAsync(out + k * batch, dO[b], len(k) * sizeof(b381_fp12), cudaMemcpyDeviceToHost, S));
        CK(cudaEventRecord(ctx->pipe_out[b], S));
    }
    CK(cudaStreamSynchronize(S));
    CK(cudaStreamSynchronize(M));
    return B381_OK;
}
int b381_final_exp_batch(b381_ctx *ctx, const b381_fp12 *in, size_t n, b381_fp12 *out, uint8_t *ok) {
    if (!ctx || (n && (!in || !out))) return B381_ERR_ARG;
    if (!n) return B381_OK;
    CK(cudaSetDevice(ctx->device));
    void *d, *dok;
    int rc = scratch_get(ctx, 0, n * sizeof(b381_fp12), &d);
    if (rc) return rc;
    rc = scratch_get(ctx, 1, n, &dok);
    if (rc) return rc;
    CK(cudaMemcpyAsync(d, in, n * sizeof(b381_fp12), cudaMemcpyHostToDevice, ctx->stream));
    rc = b381_final_exp_batch_dev(ctx, (b381_fp12 *)d, n, (b381_fp12 *)d, (uint8_t *)dok);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, d, n * sizeof(b381_fp12), cudaMemcpyDeviceToHost, ctx->stream));
    if (ok) CK(cudaMemcpyAsync(ok, dok, n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B381_OK;
}
int b381_pairing_product_is_one(b381_ctx *ctx, const b381_g1_affine *p, const b381_g2_affine *q, size_t npairs,
                                const uint32_t *group_off, size_t ngroups, uint8_t *ok) {
    if (!ctx || (ngroups && (!group_off || !ok)) || (npairs && (!p || !q))) return B381_ERR_ARG;
    if (!ngroups) return B381_OK;
    if (group_off[0] != 0 || group_off[ngroups] != npairs) return B381_ERR_ARG;
    for (size_t g = 0; g < ngroups; g++) if (group_off[g] > group_off[g + 1]) return B381_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    void *dp = nullptr, *dq = nullptr;
    int rc = staged_pq(ctx, p, q, npairs ? npairs : 1, &dp, &dq);
    if (rc) return rc;
    // offsets and result flags in grow-only scratch (a cudaMalloc / cudaFree pair per call would synchronise the device on
    // what is the latency path of a single Verify)
    void *doff = nullptr, *dok = nullptr;
    rc = scratch_get(ctx, 25, (ngroups + 1) * sizeof(uint32_t), &doff);
    if (rc) return rc;
    rc = scratch_get(ctx, 26, ngroups, &dok);
    if (rc) return rc;
    CK(cudaMemcpyAsync(doff, group_off, (ngroups + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    rc = b381_pairing_product_is_one_dev(ctx, (b381_g1_affine *)dp, (b381_g2_affine *)dq, npairs, (const uint32_t *)doff, ngroups, (uint8_t *)dok);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ok, dok, ngroups, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B381_OK;
}

// ---- aggregation, device-resident ----------------------------------------------------------------
#define SUM_BLOCK_G1 128
#define SUM_BLOCK_G2 64
static int sum_grid(b381_ctx *ctx, size_t n, int block) {
    int sms = ctx->sms > 0 ? ctx->sms : 148;
    size_t want = (n + (size_t)block * 8 - 1) / ((size_t)block * 8);      // >= 8 points per thread
    size_t cap = (size_t)sms * 4;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}
int b381_g1_sum_dev(b381_ctx *ctx, const b381_g1_affine *d_p, size_t n, b381_g1_jac *d_out) {
    if (!ctx || !d_out || (n && !d_p)) return B381_ERR_ARG;
    int grid = sum_grid(ctx, n, SUM_BLOCK_G1);
    void *part;
    int rc = scratch_get(ctx, 4, (size_t)grid * sizeof(xyzz<FpInl>), &part);
    if (rc) return rc;
    k_sum_partial<FpInl, g1_affine_pod, SUM_BLOCK_G1><<<grid, SUM_BLOCK_G1, 0, ctx->stream>>>((const g1_affine_pod *)d_p, n, (xyzz<FpInl> *)part);
    k_sum_final<FpInl, g1_jac_pod, SUM_BLOCK_G1><<<1, SUM_BLOCK_G1, 0, ctx->stream>>>((const xyzz<FpInl> *)part, grid, (g1_jac_pod *)d_out);
    ctx->launches += 2;
    CK(cudaGetLastError());
    return B381_OK;
}
int b381_g2_sum_dev(b381_ctx *ctx, const b381_g2_affine *d_p, size_t n, b381_g2_jac *d_out) {
    if (!ctx || !d_out || (n && !d_p)) return B381_ERR_ARG;
    int grid = sum_grid(ctx, n, SUM_BLOCK_G2);
    void *part;
    int rc = scratch_get(ctx, 4, (size_t)grid * sizeof(xyzz<Fp2Out>), &part);
    if (rc) return rc;
    k_sum_partial<Fp2Out, g2_affine_pod, SUM_BLOCK_G2><<<grid, SUM_BLOCK_G2, 0, ctx->stream>>>((const g2_affine_pod *)d_p, n, (xyzz<Fp2Out> *)part);
    k_sum_final<Fp2Out, g2_jac_pod, SUM_BLOCK_G2><<<1, SUM_BLOCK_G2, 0, ctx->stream>>>((const xyzz<Fp2Out> *)part, grid, (g2_jac_pod *)d_out);
    ctx->launches += 2;
    CK(cudaGetLastError());
    return B381_OK;
}
int b381_g1_fold_dev(b381_ctx *ctx, const b381_g1_jac *d_parts, size_t n, b381_g1_jac *d_out) {
    if (!ctx || !d_out || (n && !d_parts)) return B381_ERR_ARG;
    k_g1_fold<<<1, 128, 0, ctx->stream>>>((const g1_jac_pod *)d_parts, n, (g1_jac_pod *)d_out);
    ctx->launches++;
    CK(cudaGetLastError());
    return B381_OK;
}

static inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }
}   // extern "C"
// Pippenger over G1 (F = FpInl) or G2 (F = Fp2Out); nbits = scalar bits that can be non-zero (255 for field scalars, 64 for the
// weights of the random-linear-combination check): only ceil(nbits / c) windows are formed
// phase_ms (optional, host, 5 floats): sort (histogram + scan + scatter), chunk sums, chunk tree, bucket reduction (segments +
// window sums), window combine -- CUDA events on the stream; asking for them synchronises the stream at the end
// buckets (optional): stop after the bucket sums of ONE window of group_c bits over the low bits of the scalars and hand them
// out (chunk 0 of bucket d, at chunks[chunk_off[d]] when chunk_off[d + 1] > chunk_off[d], holds the sum of the points whose
// scalar is d): the group-by-key sum the attestation-level random-linear-combination check needs (key = message index + 1).
template <class F> struct msm_buckets { const xyzz<F> *chunks; const uint32_t *chunk_off; msm_geom g; };
template <class F, class APOD, class JPOD>
static int msm_shard_dev(b381_ctx *ctx, const APOD *d_p, const b381_scalar *d_k, size_t n, int nbits, int rank, int nranks, JPOD *d_partial,
                         float *phase_ms = nullptr, int group_c = 0, msm_buckets<F> *buckets = nullptr, int scratch_slot = 5) {
    if (!ctx || (!d_partial && !buckets) || (n && (!d_p || !d_k)) || nranks < 1 || rank < 0 || rank >= nranks || n > 0x7FFFFFF0u || nbits < 1 ||
        nbits > 255 || (buckets && (group_c < 2 || group_c > 24 || phase_ms)))
        return B381_ERR_ARG;                         // (bit 31 of an index entry is the sign of the digit)
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (phase_ms) for (int i = 0; i < 6; i++) CK(cudaEventCreate(&ev[i]));
#define MSM_MARK(i) do { if (phase_ms) CK(cudaEventRecord(ev[i], ctx->stream)); } while (0)
    bool whole = nranks == 1;
    msm_geom g;
    g.c = buckets ? group_c : msm_window_bits(n);
    int W = buckets ? 1 : msm_signed_windows(nbits, g.c);   // signed digits: 2^(c-1) buckets per window, one more bit for the last carry
    g.w0 = rank; g.wstep = nranks; g.nw = rank < W ? (W - rank + nranks - 1) / nranks : 0;
    g.nb = (1u << (g.c - 1)) + MSM_SEG; g.n = n;     // digit magnitudes 1 .. 2^(c-1), rounded up to whole segments
    g.maxchunks = (uint32_t)(n / MSM_CHUNK) + g.nb + 2;      // runs of the sorted list + one more chunk per bucket boundary
    g.seg = (size_t)g.maxchunks * (size_t)(g.nw > 0 ? g.nw : 1) <= MSM_SMALL_CHUNKS ? MSM_SEG_SMALL : MSM_SEG;   // (MSM_SEG_SMALL divides MSM_SEG)
    uint32_t nseg = g.nb / g.seg;
    int nw = g.nw > 0 ? g.nw : 1;
    // carve one scratch block
    size_t o_count = 0, o_boff = o_count + up256((size_t)nw * g.nb * 4), o_coff = o_boff + up256((size_t)nw * (g.nb + 1) * 4);
    size_t o_max = o_coff + up256((size_t)nw * (g.nb + 1) * 4), o_idx = o_max + 256, o_cb = o_idx + up256((size_t)nw * (n ? n : 1) * 4);
    size_t o_chunks = o_cb + up256((size_t)nw * g.maxchunks * 4), o_seg = o_chunks + up256((size_t)nw * g.maxchunks * sizeof(xyzz<F>));
    size_t o_win = o_seg + up256((size_t)nw * nseg * sizeof(xyzz<F>)), total = o_win + up256((size_t)nw * sizeof(xyzz<F>));
    char *base;
    int rc = scratch_get(ctx, scratch_slot, total, (void **)&base);
    if (rc) return rc;
    uint32_t *count = (uint32_t *)(base + o_count), *boff = (uint32_t *)(base + o_boff), *coff = (uint32_t *)(base + o_coff);
    uint32_t *maxch = (uint32_t *)(base + o_max), *idx = (uint32_t *)(base + o_idx), *cb = (uint32_t *)(base + o_cb);
    xyzz<F> *chunks = (xyzz<F> *)(base + o_chunks), *seg = (xyzz<F> *)(base + o_seg), *win = (xyzz<F> *)(base + o_win);
    if (g.nw > 0 && n > 0) {
        MSM_MARK(0);
        CK(cudaMemsetAsync(base, 0, o_idx, ctx->stream));        // count, offsets, maxch
        unsigned pg = grid_for(n, 256);
        k_msm_hist<<<pg, 256, 0, ctx->stream>>>((const uint64_t *)d_k, g, count);
        k_msm_scan<<<g.nw, 1024, 0, ctx->stream>>>(count, g, boff, coff, maxch);
        k_msm_scatter<<<pg, 256, 0, ctx->stream>>>((const uint64_t *)d_k, g, boff, count, idx);
        dim3 cg(grid_for(g.maxchunks, 128), g.nw);
        MSM_MARK(1);
        k_msm_chunk_sum<F><<<cg, 128, 0, ctx->stream>>>(d_p, idx, g, boff, coff, chunks, cb);
        ctx->launches += 4;
        MSM_MARK(2);
        const uint32_t fold_max = (size_t)g.maxchunks * (size_t)g.nw <= MSM_SMALL_CHUNKS ? MSM_FOLD_SMALL : MSM_FOLD_MAX;
        dim3 tg((unsigned)ctx->sms * 8u < cg.x ? (unsigned)ctx->sms * 8u : cg.x, g.nw);
        for (int r = 0; ((size_t)MSM_CHUNK << r) < n + MSM_CHUNK; r++) {
            k_msm_chunk_tree<F><<<tg, 128, 0, ctx->stream>>>(chunks, cb, coff, g, r, maxch, fold_max);
            ctx->launches++;
        }
        dim3 fg(grid_for(g.nb, 128), g.nw);
        k_msm_bucket_fold<F><<<fg, 128, 0, ctx->stream>>>(chunks, coff, g, maxch, fold_max);
        ctx->launches++;
        if (buckets) { buckets->chunks = chunks; buckets->chunk_off = coff; buckets->g = g; CK(cudaGetLastError()); return B381_OK; }
        MSM_MARK(3);
        bool lanes = false;
        if constexpr (lane_shift<F>::value) lanes = (size_t)g.nw * nseg <= MSM_LANE_REDUCE_MAX_SEGS;   // latency-bound: four lanes per segment
        if constexpr (lane_shift<F>::value) {
            if (lanes) {
                dim3 sg(grid_for((size_t)nseg * 4, 128), g.nw);
                k_msm_segment_reduce<F, true><<<sg, 128, 0, ctx->stream>>>(chunks, coff, g, seg);
                k_msm_window_sum<F, true><<<g.nw, 512, 0, ctx->stream>>>(seg, nseg, win);
            }
        }
        if (!lanes) {
            dim3 sg(grid_for(nseg, 128), g.nw);
            k_msm_segment_reduce<F, false><<<sg, 128, 0, ctx->stream>>>(chunks, coff, g, seg);
            k_msm_window_sum<F, false><<<g.nw, 128, 0, ctx->stream>>>(seg, nseg, win);
        }
        ctx->launches += 2;
        MSM_MARK(4);
    } else {
        if (buckets) {                               // no points: every bucket is empty
            CK(cudaMemsetAsync(base, 0, o_idx, ctx->stream));
            buckets->chunks = chunks; buckets->chunk_off = coff; buckets->g = g;
            return B381_OK;
        }
        g.nw = 0;
        for (int i = 0; i < 5; i++) MSM_MARK(i);
    }
    constexpr int CB = lane_shift<F>::value ? 128 : 64;        // G1: four lanes per window, up to 32 windows in one pass
    k_msm_combine<F, JPOD, CB><<<1, CB, 0, ctx->stream>>>(win, g, whole ? 1 : 0, d_partial);
    ctx->launches++;
    MSM_MARK(5);
    CK(cudaGetLastError());
    if (phase_ms) {
        CK(cudaEventSynchronize(ev[5]));
        for (int i = 0; i < 5; i++) { CK(cudaEventElapsedTime(&phase_ms[i], ev[i], ev[i + 1])); }
        for (int i = 0; i < 6; i++) cudaEventDestroy(ev[i]);
    }
#undef MSM_MARK
    return B381_OK;
}
extern "C" {
int b381_g1_msm_shard_dev(b381_ctx *ctx, const b381_g1_affine *d_p, const b381_scalar *d_k, size_t n, int rank, int nranks,
                          b381_g1_jac *d_partial) {
    return msm_shard_dev<FpInl>(ctx, (const g1_affine_pod *)d_p, d_k, n, 255, rank, nranks, (g1_jac_pod *)d_partial);
}
int b381_g1_msm_shard_phases_dev(b381_ctx *ctx, const b381_g1_affine *d_p, const b381_scalar *d_k, size_t n, int rank, int nranks,
                                 b381_g1_jac *d_partial, float *phase_ms) {
    if (!phase_ms) return B381_ERR_ARG;
    return msm_shard_dev<FpInl>(ctx, (const g1_affine_pod *)d_p, d_k, n, 255, rank, nranks, (g1_jac_pod *)d_partial, phase_ms);
}
int b381_g1_msm_dev(b381_ctx *ctx, const b381_g1_affine *d_p, const b381_scalar *d_k, size_t n, b381_g1_jac *d_out) {
    return b381_g1_msm_shard_dev(ctx, d_p, d_k, n, 0, 1, d_out);
}
int b381_g2_msm_dev(b381_ctx *ctx, const b381_g2_affine *d_p, const b381_scalar *d_k, size_t n, b381_g2_jac *d_out) {
    return msm_shard_dev<Fp2Out>(ctx, (const g2_affine_pod *)d_p, d_k, n, 255, 0, 1, (g2_jac_pod *)d_out);
}

int b381_verify_aggregate_common_batch_dev(b381_ctx *ctx, const b381_g1_affine *d_registry, const uint32_t *d_key_idx,
                                           const uint32_t *d_key_off, const b381_g2_affine *d_sig,
                                           const b381_g2_affine *d_msg_hash, const uint32_t *d_msg_idx, size_t nattest,
                                           size_t nkeys, size_t nmsg, uint8_t *d_ok) {
    if (!ctx || nattest > 0x7FFFFFF0u) return B381_ERR_ARG;
    if (!nattest) return B381_OK;
    if (!d_registry || !d_key_idx || !d_key_off || !d_sig || !d_msg_hash || !d_msg_idx || !d_ok || !nkeys || !nmsg) return B381_ERR_ARG;
    void *P, *Q, *off, *valid;
    int rcv = scratch_get(ctx, 24, nattest, &valid);
    if (rcv) return rcv;
    int rc = scratch_get(ctx, 2, 2 * nattest * sizeof(b381_g1_affine), &P);
    if (rc) return rc;
    rc = scratch_get(ctx, 3, 2 * nattest * sizeof(b381_g2_affine), &Q);
    if (rc) return rc;
    rc = scratch_get(ctx, 6, (nattest + 1) * sizeof(uint32_t), &off);
    if (rc) return rc;
    k_attest_pairs<<<grid_for(nattest, 128), 128, 0, ctx->stream>>>((const g1_affine_pod *)d_registry, d_key_idx, d_key_off,
                                                                    (const g2_affine_pod *)d_sig, (const g2_affine_pod *)d_msg_hash,
                                                                    d_msg_idx, nattest, nkeys, nmsg, (g1_affine_pod *)P, (g2_affine_pod *)Q,
                                                                    (uint32_t *)off, (uint8_t *)valid);
    ctx->launches++;
    CK(cudaGetLastError());
    if (path_for(ctx, nattest) == B381_PATH_THREAD && 2 * nmsg <= nattest && ctx->prepared_attest) {
        // every message point serves many attestations: G2AffineToPrepared once per message, and the pair (-pk, H(m)) of each check
        // reads the 68 coefficient triples instead of recomputing them (1 760 of the 11 600 Fq multiplications of a check)
        void *prep, *prod;
        rc = scratch_get(ctx, 27, nmsg * sizeof(b381_g2_prepared), &prep); if (rc) return rc;
        rc = scratch_get(ctx, 1, nattest * sizeof(b381_fp12), &prod); if (rc) return rc;
        rc = b381_g2_prepare_batch_dev(ctx, d_msg_hash, nmsg, (b381_g2_prepared *)prep); if (rc) return rc;
        k_miller_loop_fused_prepared<<<grid_for(nattest, PAIRING_BLOCK), PAIRING_BLOCK, 0, ctx->stream>>>(
            (const g1_affine_pod *)P, (const g2_affine_pod *)Q, (const g2_prepared_pod *)prep, d_msg_idx, nattest, (uint64_t *)prod);
        ctx->launches++;
        CK(cudaGetLastError());
        rc = final_exp_is_one(ctx, prod, nattest, d_ok);
    } else {
        rc = pairs2_product_is_one(ctx, (b381_g1_affine *)P, (b381_g2_affine *)Q, nattest, (uint32_t *)off, d_ok);
    }
    if (rc) return rc;
    k_and_bytes2<<<grid_for(nattest, 256), 256, 0, ctx->stream>>>(d_ok, (const uint8_t *)valid, nattest);
    ctx->launches++;
    CK(cudaGetLastError());
    return B381_OK;
}

// VerifyAggregateCommonWithDomain (g1pubs/bls.go:294-297) for a batch of attestations given as they arrive: compressed
// aggregate signatures and 32-byte message hashes; the committee keys come from a resident, already validated registry.
// Device: DeserializeSignature (subgroup-checked) per attestation, HashG2WithDomain per DISTINCT message, per-attestation key
// aggregation, one CompareTwoPairings per attestation.
int b381_verify_aggregate_common_with_domain_batch_dev(b381_ctx *ctx, const b381_g1_affine *d_registry, const uint32_t *d_key_idx,
                                                       const uint32_t *d_key_off, const uint8_t *d_sig96, const uint8_t *d_msg32, size_t nmsg,
                                                       const uint8_t *d_domain8, const uint32_t *d_msg_idx, size_t nattest, size_t nkeys,
                                                       uint8_t *d_ok) {
    if (!ctx || nattest > 0x7FFFFFF0u) return B381_ERR_ARG;
    if (!nattest) return B381_OK;
    if (!d_registry || !d_key_idx || !d_key_off || !d_sig96 || !d_msg32 || !nmsg || !d_domain8 || !d_msg_idx || !d_ok) return B381_ERR_ARG;
    void *sig, *H, *st;
    int rc = scratch_get(ctx, 16, nattest * sizeof(b381_g2_affine), &sig); if (rc) return rc;
    rc = scratch_get(ctx, 17, nmsg * sizeof(b381_g2_affine), &H); if (rc) return rc;
    rc = scratch_get(ctx, 18, nattest, &st); if (rc) return rc;
    rc = b381_g2_decompress_batch_dev(ctx, d_sig96, nattest, 1, (b381_g2_affine *)sig, (uint8_t *)st); if (rc) return rc;
    rc = b381_hash_g2_with_domain_batch_dev(ctx, d_msg32, d_domain8, 0, nmsg, (b381_g2_affine *)H); if (rc) return rc;
    rc = b381_verify_aggregate_common_batch_dev(ctx, d_registry, d_key_idx, d_key_off, (const b381_g2_affine *)sig, (const b381_g2_affine *)H,
                                                d_msg_idx, nattest, nkeys, nmsg, d_ok);
    if (rc) return rc;
    k_and_status_g2<<<grid_for(nattest, 256), 256, 0, ctx->stream>>>(d_ok, (const uint8_t *)st, (const g2_affine_pod *)sig, nattest);
    ctx->launches++;
    CK(cudaGetLastError());
    return B381_OK;
}

// ---- Verify / VerifyWithDomain from wire bytes: deserialise + hash + 2-pair check per item, all on the device -----------------
}   // extern "C"
enum { WIRE_G1PUBS_DOMAIN = 0, WIRE_G1PUBS = 1, WIRE_G2PUBS = 2 };
// mode WIRE_G1PUBS_DOMAIN: msg = n x 32 bytes, aux = domain (8 bytes x (stride ? n : 1)); otherwise msg = packed messages and
// aux = n + 1 u64 offsets.  g1pubs: keys 48 B / signatures 96 B; g2pubs: keys 96 B / signatures 48 B.
static int verify_rlc_core(b381_ctx *ctx, const b381_g1_affine *d_pub, const b381_g2_affine *d_h, const b381_g2_affine *d_sig,
                           const uint8_t *d_pub_status, const uint8_t *d_sig_status, const b381_scalar *d_r, size_t n, uint8_t *d_ok,
                           b381_fp12 *d_partial);
static int verify_wire_dev(b381_ctx *ctx, int mode, const uint8_t *d_pub, const uint8_t *d_msg, const void *d_aux, size_t domain_stride,
                           const uint8_t *d_sig, size_t n, uint8_t *d_ok, const b381_scalar *d_rlc = nullptr) {
    if (!ctx || domain_stride > 1 || n > 0x7FFFFFF0u || (n && (!d_pub || !d_msg || !d_aux || !d_sig || !d_ok))) return B381_ERR_ARG;
    if (!n) return B381_OK;
    const bool g2p = mode == WIRE_G2PUBS;
    const size_t pub_pod = g2p ? sizeof(b381_g2_affine) : sizeof(b381_g1_affine), sig_pod = g2p ? sizeof(b381_g1_affine) : sizeof(b381_g2_affine);
    void *pub, *sig, *H, *st, *P, *Q, *off;
    int rc = scratch_get(ctx, 15, n * pub_pod, &pub); if (rc) return rc;
    rc = scratch_get(ctx, 16, n * sig_pod, &sig); if (rc) return rc;
    rc = scratch_get(ctx, 17, n * sig_pod, &H); if (rc) return rc;          // the message point lives in the signature's group
    rc = scratch_get(ctx, 18, 3 * n, &st); if (rc) return rc;
    rc = scratch_get(ctx, 2, 2 * n * sizeof(b381_g1_affine), &P); if (rc) return rc;
    rc = scratch_get(ctx, 3, 2 * n * sizeof(b381_g2_affine), &Q); if (rc) return rc;
    rc = scratch_get(ctx, 6, (n + 1) * sizeof(uint32_t), &off); if (rc) return rc;
    uint8_t *st_pub = (uint8_t *)st, *st_sig = st_pub + n, *valid = st_pub + 2 * n;
    if (!g2p) {
        rc = b381_g1_decompress_batch_dev(ctx, d_pub, n, 1, (b381_g1_affine *)pub, st_pub); if (rc) return rc;
        rc = b381_g2_decompress_batch_dev(ctx, d_sig, n, 1, (b381_g2_affine *)sig, st_sig); if (rc) return rc;
        if (mode == WIRE_G1PUBS_DOMAIN)
            rc = b381_hash_g2_with_domain_batch_dev(ctx, d_msg, (const uint8_t *)d_aux, domain_stride, n, (b381_g2_affine *)H);
        else
            rc = b381_hash_g2_batch_dev(ctx, d_msg, (const uint64_t *)d_aux, n, (b381_g2_affine *)H);
        if (rc) return rc;
        if (d_rlc)       // one boolean for the whole batch
            return verify_rlc_core(ctx, (const b381_g1_affine *)pub, (const b381_g2_affine *)H, (const b381_g2_affine *)sig, st_pub, st_sig, d_rlc, n, d_ok, nullptr);
        k_verify_pairs<<<grid_for(n, 128), 128, 0, ctx->stream>>>((const g1_affine_pod *)pub, st_pub, (const g2_affine_pod *)sig, st_sig,
                                                                  (const g2_affine_pod *)H, n, (g1_affine_pod *)P, (g2_affine_pod *)Q,
                                                                  (uint32_t *)off, valid);
    } else {
        if (d_rlc) return B381_ERR_ARG;
        rc = b381_g2_decompress_batch_dev(ctx, d_pub, n, 1, (b381_g2_affine *)pub, st_pub); if (rc) return rc;
        rc = b381_g1_decompress_batch_dev(ctx, d_sig, n, 1, (b381_g1_affine *)sig, st_sig); if (rc) return rc;
        rc = b381_hash_g1_batch_dev(ctx, d_msg, (const uint64_t *)d_aux, n, (b381_g1_affine *)H); if (rc) return rc;
        k_verify_pairs_g2pubs<<<grid_for(n, 128), 128, 0, ctx->stream>>>((const g2_affine_pod *)pub, st_pub, (const g1_affine_pod *)sig, st_sig,
                                                                         (const g1_affine_pod *)H, n, (g1_affine_pod *)P, (g2_affine_pod *)Q,
                                                                         (uint32_t *)off, valid);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    rc = pairs2_product_is_one(ctx, (b381_g1_affine *)P, (b381_g2_affine *)Q, n, (uint32_t *)off, d_ok);
    if (rc) return rc;
    k_and_bytes<<<grid_for(n, 256), 256, 0, ctx->stream>>>(d_ok, valid, n);
    ctx->launches++;
    CK(cudaGetLastError());
    return B381_OK;
}
// S = sum_i r_i sig_i on the ctx's SIDE stream, forked from the main stream here and joined by the caller with side_join():
// the G2 MSM of a random-linear-combination check depends on nothing the G1 side computes, and both are chains of small,
// latency-bound launches at these sizes, so they overlap.  Its work area is a slot of its own (36), not the main MSM's.
static int g2_msm_on_side_stream(b381_ctx *ctx, const b381_g2_affine *d_sig, const b381_scalar *d_r, size_t n, int bits, g2_jac_pod *S) {
    CK(cudaEventRecord(ctx->side_fork, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->side_stream, ctx->side_fork, 0));
    cudaStream_t main_stream = ctx->stream;
    ctx->stream = ctx->side_stream;
    int rc = msm_shard_dev<Fp2Out>(ctx, (const g2_affine_pod *)d_sig, d_r, n, bits, 0, 1, S, nullptr, 0, nullptr, 36);
    cudaError_t e = cudaEventRecord(ctx->side_join, ctx->side_stream);
    ctx->stream = main_stream;
    if (rc) return rc;
    CK(e);
    return B381_OK;
}
static int side_join(b381_ctx *ctx) {
    CK(cudaStreamWaitEvent(ctx->stream, ctx->side_join, 0));
    return B381_OK;
}
// ---- attestation batches as ONE random-linear-combination check, grouped by message -------------------------------------------
// ok = [ e(-G1One, sum_a r_a sig_a) * prod_m e(sum_{a: msg(a) = m} r_a pk_a, H_m) == 1 ]: with independent random r_a this accepts
// iff every e(G1One, sig_a) == e(pk_a, H_msg(a)) holds (VerifyAggregateCommon, g1pubs/bls.go:287-297), except with probability
// ~2^-bits(r).  nmsg + 1 Miller loops and one final exponentiation for the whole batch instead of two loops and one
// exponentiation per attestation; the work that scales with the batch is the committee sums, one short scalar multiplication
// per attestation, a group-by-message sum (the bucket machinery of the MSM, keyed by message) and a 64-bit-weight G2 MSM.
extern "C" int b381_verify_aggregate_common_rlc_dev(b381_ctx *ctx, const b381_g1_affine *d_registry, const uint32_t *d_key_idx,
                                                    const uint32_t *d_key_off, const b381_g2_affine *d_sig, const b381_g2_affine *d_msg_hash,
                                                    const uint32_t *d_msg_idx, const b381_scalar *d_r, size_t nattest, size_t nkeys, size_t nmsg,
                                                    uint8_t *d_ok) {
    if (!ctx || !d_ok || nattest > 0x7FFFFFF0u || nmsg > (1u << 22)) return B381_ERR_ARG;
    if (!nattest) { CK(cudaMemsetAsync(d_ok, 1, 1, ctx->stream)); return B381_OK; }       // the empty product
    if (!d_registry || !d_key_idx || !d_key_off || !d_sig || !d_msg_hash || !d_msg_idx || !d_r || !nkeys || !nmsg) return B381_ERR_ARG;
    const int rlc_bits = ctx->rlc_bits;
    void *P2, *Q2, *off2, *valid, *W, *keys, *P, *Q, *S, *off, *bad;
    int rc = scratch_get(ctx, 24, nattest, &valid); if (rc) return rc;
    rc = scratch_get(ctx, 2, 2 * nattest * sizeof(b381_g1_affine), &P2); if (rc) return rc;
    rc = scratch_get(ctx, 3, 2 * nattest * sizeof(b381_g2_affine), &Q2); if (rc) return rc;
    rc = scratch_get(ctx, 6, (nattest + 1) * sizeof(uint32_t), &off2); if (rc) return rc;
    rc = scratch_get(ctx, 32, nattest * sizeof(b381_g1_affine), &W); if (rc) return rc;
    rc = scratch_get(ctx, 33, nattest * sizeof(b381_scalar), &keys); if (rc) return rc;
    rc = scratch_get(ctx, 34, (nmsg + 1) * sizeof(b381_g1_affine), &P); if (rc) return rc;
    rc = scratch_get(ctx, 35, (nmsg + 1) * sizeof(b381_g2_affine), &Q); if (rc) return rc;
    rc = scratch_get(ctx, 23, sizeof(b381_g2_jac) + 64, &S); if (rc) return rc;
    rc = scratch_get(ctx, 22, 2 * sizeof(uint32_t), &off); if (rc) return rc;
    bad = (char *)S + sizeof(b381_g2_jac);
    CK(cudaMemsetAsync(bad, 0, sizeof(uint32_t), ctx->stream));
    // T = sum_a r_a sig_a runs beside everything up to the closing pair
    rc = g2_msm_on_side_stream(ctx, d_sig, d_r, nattest, rlc_bits, (g2_jac_pod *)S); if (rc) return rc;
    // committee sums: P2[2a + 1] = -pk_a, valid[a] (indices in range, committee not empty, pk_a and sig_a finite)
    k_attest_pairs<<<grid_for(nattest, 128), 128, 0, ctx->stream>>>((const g1_affine_pod *)d_registry, d_key_idx, d_key_off,
                                                                    (const g2_affine_pod *)d_sig, (const g2_affine_pod *)d_msg_hash,
                                                                    d_msg_idx, nattest, nkeys, nmsg, (g1_affine_pod *)P2, (g2_affine_pod *)Q2,
                                                                    (uint32_t *)off2, (uint8_t *)valid);
    k_attest_rlc_valid<<<grid_for(nattest, 256), 256, 0, ctx->stream>>>((const uint8_t *)valid, (const uint64_t *)d_r, rlc_bits, nattest, (uint32_t *)bad);
    k_group_keys<<<grid_for(nattest, 256), 256, 0, ctx->stream>>>(d_msg_idx, nattest, nmsg, (uint64_t *)keys);
    ctx->launches += 3;
    CK(cudaGetLastError());
    // W_a = r_a (-pk_a): leading zero windows of a short weight cost nothing (the ladder doubles infinity)
    k_point_mul<G1Codec, false><<<grid_for(nattest, 64), 64, 0, ctx->stream>>>((const g1_affine_pod *)P2 + 1, 2, (const uint64_t *)d_r, 1, nattest,
                                                                               (g1_affine_pod *)W);
    ctx->launches++;
    CK(cudaGetLastError());
    // group by message: the bucket sums of one window wide enough for nmsg + 1 keys
    int gc = 2;
    while ((1ull << (gc - 1)) < nmsg + 1) gc++;
    msm_buckets<FpInl> bk;
    rc = msm_shard_dev<FpInl, g1_affine_pod, g1_jac_pod>(ctx, (const g1_affine_pod *)W, (const b381_scalar *)keys, nattest, 64, 0, 1, nullptr, nullptr, gc, &bk);
    if (rc) return rc;
    k_group_pairs<<<grid_for(nmsg, 64), 64, 0, ctx->stream>>>(bk.chunks, bk.chunk_off, (uint32_t)nmsg, (const g2_affine_pod *)d_msg_hash,
                                                              (g1_affine_pod *)P, (g2_affine_pod *)Q);
    ctx->launches++;
    CK(cudaGetLastError());
    // closing pair (-G1One, T)
    rc = side_join(ctx); if (rc) return rc;
    k_rlc_close<<<1, 128, 0, ctx->stream>>>((const g2_jac_pod *)S, (g1_affine_pod *)P + nmsg, (g2_affine_pod *)Q + nmsg, (uint32_t *)off, (uint32_t)nmsg);
    ctx->launches++;
    rc = b381_pairing_product_is_one_dev(ctx, (const b381_g1_affine *)P, (const b381_g2_affine *)Q, nmsg + 1, (const uint32_t *)off, 1, d_ok);
    if (rc) return rc;
    k_rlc_finish<<<1, 32, 0, ctx->stream>>>(d_ok, (const uint32_t *)bad);
    ctx->launches++;
    CK(cudaGetLastError());
    return B381_OK;
}
// ---- random-linear-combination batch verification: ONE final exponentiation for n signatures ---------------------------------
// ok = [ prod_i e(r_i pk_i, H_i) * e(-G1One, sum_i r_i sig_i) == 1 ]: with independent random r_i this accepts iff every
// e(G1One, sig_i) == e(pk_i, H_i) holds (g1pubs.Verify*, g1pubs/bls.go:165-174), except with probability ~2^-bits(r).
// n + 1 Miller loops, one tree product, one final exponentiation, n short scalar multiplications in G1 and G2.
// d_partial (optional) receives prod_i ML(r_i pk_i, H_i) * ML(-G1One, sum_i r_i sig_i) without the final exponentiation and
// *d_ok the validity flag only (1 = no infinite / undecodable input); with d_partial == nullptr the check is finished here.
static int verify_rlc_core(b381_ctx *ctx, const b381_g1_affine *d_pub, const b381_g2_affine *d_h, const b381_g2_affine *d_sig,
                           const uint8_t *d_pub_status, const uint8_t *d_sig_status, const b381_scalar *d_r, size_t n, uint8_t *d_ok,
                           b381_fp12 *d_partial = nullptr) {
    if (!ctx || n > 0x7FFFFFF0u || !d_ok || (n && (!d_pub || !d_h || !d_sig || !d_r))) return B381_ERR_ARG;
    const int rlc_bits = ctx->rlc_bits;     // the weights' bit length (b381_set_rlc_weight_bits): MSM windows and ladder length follow it
    void *P, *Q, *S, *off, *bad;
    int rc = scratch_get(ctx, 2, (n + 1) * sizeof(b381_g1_affine), &P); if (rc) return rc;
    rc = scratch_get(ctx, 3, (n + 1) * sizeof(b381_g2_affine), &Q); if (rc) return rc;
    rc = scratch_get(ctx, 23, sizeof(b381_g2_jac) + 64, &S); if (rc) return rc;
    rc = scratch_get(ctx, 22, 2 * sizeof(uint32_t), &off); if (rc) return rc;
    bad = (char *)S + sizeof(b381_g2_jac);          // (slot 21 is the tree product's own scalar)
    CK(cudaMemsetAsync(bad, 0, sizeof(uint32_t), ctx->stream));
    // S = sum_i r_i sig_i as one Pippenger MSM over G2 (the weights are scalars like any other), beside the G1 scalar multiplications
    rc = g2_msm_on_side_stream(ctx, d_sig, d_r, n, rlc_bits, (g2_jac_pod *)S); if (rc) return rc;
    if (n) {
        k_rlc_valid<<<grid_for(n, 256), 256, 0, ctx->stream>>>((const g1_affine_pod *)d_pub, (const g2_affine_pod *)d_sig, d_pub_status,
                                                               d_sig_status, (const uint64_t *)d_r, rlc_bits, n, (uint32_t *)bad);
        ctx->launches++;
        rc = b381_g1_mul_batch_dev(ctx, d_pub, 1, d_r, 1, n, (b381_g1_affine *)P); if (rc) return rc;
        CK(cudaMemcpyAsync(Q, d_h, n * sizeof(b381_g2_affine), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    rc = side_join(ctx); if (rc) return rc;
    k_rlc_close<<<1, 128, 0, ctx->stream>>>((const g2_jac_pod *)S, (g1_affine_pod *)P + n, (g2_affine_pod *)Q + n, (uint32_t *)off, (uint32_t)n);
    ctx->launches++;
    if (d_partial) {
        void *ml;
        rc = scratch_get(ctx, 0, (n + 1) * sizeof(b381_fp12), &ml); if (rc) return rc;
        rc = b381_miller_loop_batch_dev(ctx, (const b381_g1_affine *)P, (const b381_g2_affine *)Q, n + 1, (b381_fp12 *)ml); if (rc) return rc;
        rc = group_products(ctx, ml, n + 1, (const uint32_t *)off, 1, d_partial); if (rc) return rc;
        CK(cudaMemsetAsync(d_ok, 1, 1, ctx->stream));
    } else {
        rc = b381_pairing_product_is_one_dev(ctx, (const b381_g1_affine *)P, (const b381_g2_affine *)Q, n + 1, (const uint32_t *)off, 1, d_ok);
        if (rc) return rc;
    }
    k_rlc_finish<<<1, 32, 0, ctx->stream>>>(d_ok, (const uint32_t *)bad);
    ctx->launches++;
    CK(cudaGetLastError());
    return B381_OK;
}
static int verify_wire_host(b381_ctx *ctx, int mode, const uint8_t *pub, const uint8_t *msg, size_t msg_bytes, const void *aux, size_t aux_bytes,
                            size_t domain_stride, const uint8_t *sig, size_t n, uint8_t *ok, const b381_scalar *rlc = nullptr) {
    if (!ctx || domain_stride > 1 || (n && (!pub || !msg || !aux || !sig || !ok))) return B381_ERR_ARG;
    if (!n) return B381_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t pb = mode == WIRE_G2PUBS ? 96 : 48, sb = mode == WIRE_G2PUBS ? 48 : 96;
    void *in, *dok;
    size_t aux_off = (n * (pb + sb) + msg_bytes + 15) & ~(size_t)15;
    int rc = scratch_get(ctx, 19, aux_off + aux_bytes, &in); if (rc) return rc;
    rc = scratch_get(ctx, 20, n, &dok); if (rc) return rc;
    uint8_t *dp = (uint8_t *)in, *ds = dp + pb * n, *dm = ds + sb * n, *da = dp + aux_off;
    CK(cudaMemcpyAsync(dp, pub, pb * n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ds, sig, sb * n, cudaMemcpyHostToDevice, ctx->stream));
    if (msg_bytes) CK(cudaMemcpyAsync(dm, msg, msg_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(da, aux, aux_bytes, cudaMemcpyHostToDevice, ctx->stream));
    void *dr = nullptr;
    if (rlc) {
        rc = scratch_get(ctx, 14, n * sizeof(b381_scalar), &dr); if (rc) return rc;
        CK(cudaMemcpyAsync(dr, rlc, n * sizeof(b381_scalar), cudaMemcpyHostToDevice, ctx->stream));
    }
    rc = verify_wire_dev(ctx, mode, dp, dm, da, domain_stride, ds, n, (uint8_t *)dok, (const b381_scalar *)dr); if (rc) return rc;
    CK(cudaMemcpyAsync(ok, dok, rlc ? 1 : n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B381_OK;
}
template <class S, class APOD>
static int hash_curve_dev(b381_ctx *ctx, const uint8_t *d_msgs, const uint64_t *d_off, size_t n, APOD *d_out) {
    if (!ctx || (n && (!d_msgs || !d_off || !d_out))) return B381_ERR_ARG;
    if (!n) return B381_OK;
    k_hash_to_curve<S><<<grid_for(n, 64), 64, 0, ctx->stream>>>(d_msgs, d_off, n, (typename S::APOD *)d_out);
    ctx->launches++;
    CK(cudaGetLastError());
    return B381_OK;
}
template <class S, class APOD>
static int hash_curve_host(b381_ctx *ctx, const uint8_t *msgs, const uint64_t *off, size_t n, APOD *out) {
    if (!ctx || (n && (!msgs || !off || !out))) return B381_ERR_ARG;
    if (!n) return B381_OK;
    if (off[0] != 0) return B381_ERR_ARG;
    for (size_t i = 0; i < n; i++) if (off[i] > off[i + 1]) return B381_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    void *dm, *doff, *dout;
    size_t mb = off[n];
    int rc = scratch_get(ctx, 12, mb + 16, &dm); if (rc) return rc;
    rc = scratch_get(ctx, 13, n * sizeof(APOD), &dout); if (rc) return rc;
    rc = scratch_get(ctx, 14, (n + 1) * sizeof(uint64_t), &doff); if (rc) return rc;
    if (mb) CK(cudaMemcpyAsync(dm, msgs, mb, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(doff, off, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    rc = hash_curve_dev<S, APOD>(ctx, (const uint8_t *)dm, (const uint64_t *)doff, n, (APOD *)dout); if (rc) return rc;
    CK(cudaMemcpyAsync(out, dout, n * sizeof(APOD), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B381_OK;
}
static int check_offsets(const uint64_t *off, size_t n) {
    if (!off || off[0] != 0) return B381_ERR_ARG;
    for (size_t i = 0; i < n; i++) if (off[i] > off[i + 1]) return B381_ERR_ARG;
    return B381_OK;
}
extern "C" {
int b381_hash_g1_batch(b381_ctx *ctx, const uint8_t *msgs, const uint64_t *msg_off, size_t n, b381_g1_affine *out) { return hash_curve_host<G1Swu>(ctx, msgs, msg_off, n, out); }
int b381_hash_g1_batch_dev(b381_ctx *ctx, const uint8_t *d_msgs, const uint64_t *d_msg_off, size_t n, b381_g1_affine *d_out) { return hash_curve_dev<G1Swu>(ctx, d_msgs, d_msg_off, n, d_out); }
int b381_hash_g2_batch(b381_ctx *ctx, const uint8_t *msgs, const uint64_t *msg_off, size_t n, b381_g2_affine *out) { return hash_curve_host<G2Swu>(ctx, msgs, msg_off, n, out); }
int b381_hash_g2_batch_dev(b381_ctx *ctx, const uint8_t *d_msgs, const uint64_t *d_msg_off, size_t n, b381_g2_affine *d_out) { return hash_curve_dev<G2Swu>(ctx, d_msgs, d_msg_off, n, d_out); }
int b381_verify_with_domain_batch_dev(b381_ctx *ctx, const uint8_t *d_pub48, const uint8_t *d_msg32, const uint8_t *d_domain8,
                                      size_t domain_stride, const uint8_t *d_sig96, size_t n, uint8_t *d_ok) {
    return verify_wire_dev(ctx, WIRE_G1PUBS_DOMAIN, d_pub48, d_msg32, d_domain8, domain_stride, d_sig96, n, d_ok);
}
int b381_verify_with_domain_batch(b381_ctx *ctx, const uint8_t *pub48, const uint8_t *msg32, const uint8_t *domain8, size_t domain_stride,
                                  const uint8_t *sig96, size_t n, uint8_t *ok) {
    return verify_wire_host(ctx, WIRE_G1PUBS_DOMAIN, pub48, msg32, 32 * n, domain8, 8 * (domain_stride ? n : 1), domain_stride, sig96, n, ok);
}
int b381_verify_rlc_dev(b381_ctx *ctx, const b381_g1_affine *d_pub, const b381_g2_affine *d_msg_point, const b381_g2_affine *d_sig,
                        const b381_scalar *d_r, size_t n, uint8_t *d_ok) {
    return verify_rlc_core(ctx, d_pub, d_msg_point, d_sig, nullptr, nullptr, d_r, n, d_ok);
}
int b381_verify_rlc_partial_dev(b381_ctx *ctx, const b381_g1_affine *d_pub, const b381_g2_affine *d_msg_point, const b381_g2_affine *d_sig,
                                const b381_scalar *d_r, size_t n, b381_fp12 *d_partial, uint8_t *d_valid) {
    if (!d_partial) return B381_ERR_ARG;
    return verify_rlc_core(ctx, d_pub, d_msg_point, d_sig, nullptr, nullptr, d_r, n, d_valid, d_partial);
}
int b381_verify_with_domain_rlc_batch_dev(b381_ctx *ctx, const uint8_t *d_pub48, const uint8_t *d_msg32, const uint8_t *d_domain8,
                                          size_t domain_stride, const uint8_t *d_sig96, const b381_scalar *d_r, size_t n, uint8_t *d_ok) {
    if (!d_r) return B381_ERR_ARG;
    return verify_wire_dev(ctx, WIRE_G1PUBS_DOMAIN, d_pub48, d_msg32, d_domain8, domain_stride, d_sig96, n, d_ok, d_r);
}
int b381_verify_with_domain_rlc_batch(b381_ctx *ctx, const uint8_t *pub48, const uint8_t *msg32, const uint8_t *domain8, size_t domain_stride,
                                      const uint8_t *sig96, const b381_scalar *r, size_t n, uint8_t *ok) {
    if (!r) return B381_ERR_ARG;
    return verify_wire_host(ctx, WIRE_G1PUBS_DOMAIN, pub48, msg32, 32 * n, domain8, 8 * (domain_stride ? n : 1), domain_stride, sig96, n, ok, r);
}
int b381_g1pubs_verify_batch_dev(b381_ctx *ctx, const uint8_t *d_pub48, const uint8_t *d_msgs, const uint64_t *d_msg_off, const uint8_t *d_sig96,
                                 size_t n, uint8_t *d_ok) {
    return verify_wire_dev(ctx, WIRE_G1PUBS, d_pub48, d_msgs, d_msg_off, 0, d_sig96, n, d_ok);
}
int b381_g1pubs_verify_batch(b381_ctx *ctx, const uint8_t *pub48, const uint8_t *msgs, const uint64_t *msg_off, const uint8_t *sig96, size_t n,
                             uint8_t *ok) {
    if (n && check_offsets(msg_off, n)) return B381_ERR_ARG;
    return verify_wire_host(ctx, WIRE_G1PUBS, pub48, msgs, n ? msg_off[n] : 0, msg_off, (n + 1) * sizeof(uint64_t), 0, sig96, n, ok);
}
int b381_g2pubs_verify_batch_dev(b381_ctx *ctx, const uint8_t *d_pub96, const uint8_t *d_msgs, const uint64_t *d_msg_off, const uint8_t *d_sig48,
                                 size_t n, uint8_t *d_ok) {
    return verify_wire_dev(ctx, WIRE_G2PUBS, d_pub96, d_msgs, d_msg_off, 0, d_sig48, n, d_ok);
}
int b381_g2pubs_verify_batch(b381_ctx *ctx, const uint8_t *pub96, const uint8_t *msgs, const uint64_t *msg_off, const uint8_t *sig48, size_t n,
                             uint8_t *ok) {
    if (n && check_offsets(msg_off, n)) return B381_ERR_ARG;
    return verify_wire_host(ctx, WIRE_G2PUBS, pub96, msgs, n ? msg_off[n] : 0, msg_off, (n + 1) * sizeof(uint64_t), 0, sig48, n, ok);
}

// ---- aggregation, host buffers -------------------------------------------------------------------
}  // extern "C"
template <class AFF, class JAC, class FN>
static int sum_host(b381_ctx *ctx, const AFF *p, size_t n, JAC *out, FN fn) {
    if (!ctx || !out || (n && !p)) return B381_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    void *dp, *dout;
    int rc = scratch_get(ctx, 2, (n ? n : 1) * sizeof(AFF), &dp);
    if (rc) return rc;
    rc = scratch_get(ctx, 7, sizeof(b381_g2_jac), &dout);
    if (rc) return rc;
    if (n) CK(cudaMemcpyAsync(dp, p, n * sizeof(AFF), cudaMemcpyHostToDevice, ctx->stream));
    rc = fn(ctx, (const AFF *)dp, n, (JAC *)dout);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, dout, sizeof(JAC), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B381_OK;
}
extern "C" {
int b381_g1_sum(b381_ctx *ctx, const b381_g1_affine *p, size_t n, b381_g1_jac *out) { return sum_host(ctx, p, n, out, b381_g1_sum_dev); }
int b381_g2_sum(b381_ctx *ctx, const b381_g2_affine *p, size_t n, b381_g2_jac *out) { return sum_host(ctx, p, n, out, b381_g2_sum_dev); }
int b381_g1_msm(b381_ctx *ctx, const b381_g1_affine *p, const b381_scalar *k, size_t n, b381_g1_jac *out) {
    if (!ctx || !out || (n && (!p || !k))) return B381_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    void *dp, *dk, *dout;
    int rc = scratch_get(ctx, 2, (n ? n : 1) * sizeof(b381_g1_affine), &dp);
    if (rc) return rc;
    rc = scratch_get(ctx, 3, (n ? n : 1) * sizeof(b381_scalar), &dk);
    if (rc) return rc;
    rc = scratch_get(ctx, 7, sizeof(b381_g2_jac), &dout);
    if (rc) return rc;
    if (n) {
        CK(cudaMemcpyAsync(dp, p, n * sizeof(b381_g1_affine), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(dk, k, n * sizeof(b381_scalar), cudaMemcpyHostToDevice, ctx->stream));
    }
    rc = b381_g1_msm_dev(ctx, (const b381_g1_affine *)dp, (const b381_scalar *)dk, n, (b381_g1_jac *)dout);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, dout, sizeof(b381_g1_jac), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B381_OK;
}
int b381_g2_msm(b381_ctx *ctx, const b381_g2_affine *p, const b381_scalar *k, size_t n, b381_g2_jac *out) {
    if (!ctx || !out || (n && (!p || !k))) return B381_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    void *dp, *dk, *dout;
    int rc = scratch_get(ctx, 2, (n ? n : 1) * sizeof(b381_g2_affine), &dp);
    if (rc) return rc;
    rc = scratch_get(ctx, 3, (n ? n : 1) * sizeof(b381_scalar), &dk);
    if (rc) return rc;
    rc = scratch_get(ctx, 7, sizeof(b381_g2_jac), &dout);
    if (rc) return rc;
    if (n) {
        CK(cudaMemcpyAsync(dp, p, n * sizeof(b381_g2_affine), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(dk, k, n * sizeof(b381_scalar), cudaMemcpyHostToDevice, ctx->stream));
    }
    rc = b381_g2_msm_dev(ctx, (const b381_g2_affine *)dp, (const b381_scalar *)dk, n, (b381_g2_jac *)dout);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, dout, sizeof(b381_g2_jac), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B381_OK;
}

}  // extern "C"
