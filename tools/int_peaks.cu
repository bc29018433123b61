// int_peaks.cu -- integer-pipe microbenchmarks for the roofline denominator.
// Measures sustained per-SM issue rates (thread-ops per clock per SM) and chip rates (ops/s) of the
// instructions a 384-bit Montgomery multiplication is made of.  One wave, 8 blocks x 256 threads per
// SM (full occupancy), 8 independent dependency chains per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/int_peaks tools/int_peaks.cu
// Run under gpurun: ./tools/int_peaks > gpurun_out/int_peaks.json
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include "../bls_b200/csrc/fp.cuh"

#define CHAINS 8
#define KERNEL(NAME, DECL, INIT, BODY, FOLD)                                                         \
    __global__ void __launch_bounds__(256) NAME(uint32_t *out, long long *cyc, int iters) {         \
        uint32_t b = blockIdx.x * 40503u + threadIdx.x * 2654435761u + 7u;                          \
        DECL;                                                                                        \
        _Pragma("unroll") for (int j = 0; j < CHAINS; j++) { INIT; }                                 \
        __syncthreads();                                                                             \
        long long t0 = clock64();                                                                    \
        _Pragma("unroll 1") for (int i = 0; i < iters; i++) {                                        \
            _Pragma("unroll") for (int j = 0; j < CHAINS; j++) { BODY; }                             \
        }                                                                                            \
        long long t1 = clock64();                                                                    \
        uint32_t r = 0;                                                                              \
        _Pragma("unroll") for (int j = 0; j < CHAINS; j++) { FOLD; }                                 \
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;                                              \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                             \
    }

// IMAD.WIDE.U32 Rd64 = Ra * Rb + Rc64
KERNEL(k_mad_wide, uint64_t acc[CHAINS], acc[j] = (uint64_t)(b + j) * 0x9E3779B97F4A7C15ull,
       asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"((uint32_t)acc[(j + 1) & 7]), "r"(b)),
       r ^= (uint32_t)acc[j] ^ (uint32_t)(acc[j] >> 32))
// IMAD.WIDE.U32 with an immediate multiplier (the modulus rows of the reduction)
KERNEL(k_mad_wide_imm, uint64_t acc[CHAINS], acc[j] = (uint64_t)(b + j) * 0x9E3779B97F4A7C15ull,
       asm volatile("mad.wide.u32 %0, %1, 0x4b1ba7b6, %0;" : "+l"(acc[j]) : "r"((uint32_t)acc[(j + 1) & 7])),
       r ^= (uint32_t)acc[j] ^ (uint32_t)(acc[j] >> 32))
// IMAD (32-bit low product)
KERNEL(k_mad_lo, uint32_t acc[CHAINS], acc[j] = b + j,
       asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(acc[(j + 1) & 7]), "r"(b)),
       r ^= acc[j])
// IMAD.HI.U32
KERNEL(k_mad_hi, uint32_t acc[CHAINS], acc[j] = b + j,
       asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(acc[(j + 1) & 7] | 0x80000000u), "r"(b | 0x80000000u)),
       r ^= acc[j])
// carry-chained wide MAC pair as emitted for Fq: mad.lo.cc + madc.hi (one IMAD.WIDE.U32.X-class op per pair)
KERNEL(k_mad_cc_pair, uint32_t lo[CHAINS]; uint32_t hi[CHAINS], lo[j] = b + j; hi[j] = b ^ j,
       asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[j]), "+r"(hi[j]) : "r"(lo[(j + 1) & 7]), "r"(b)),
       r ^= lo[j] ^ hi[j])
// IADD3 (ALU pipe)
KERNEL(k_iadd3, uint32_t acc[CHAINS], acc[j] = b + j,
       asm volatile("add.u32 %0, %0, %1;" : "+r"(acc[j]) : "r"(acc[(j + 1) & 7])),
       r ^= acc[j])
// 1:1 mix IMAD.WIDE + IADD3: do the two pipes issue concurrently?
KERNEL(k_mix_wide_iadd, uint64_t acc[CHAINS]; uint32_t s[CHAINS], acc[j] = (uint64_t)(b + j) * 0x9E3779B97F4A7C15ull; s[j] = b ^ j,
       asm volatile("mad.wide.u32 %0, %2, %3, %0;\n\tadd.u32 %1, %1, %2;" : "+l"(acc[j]), "+r"(s[j]) : "r"((uint32_t)acc[(j + 1) & 7]), "r"(b)),
       r ^= (uint32_t)acc[j] ^ (uint32_t)(acc[j] >> 32) ^ s[j])
// 1:1 mix IMAD(lo) + IMAD.HI: the two-instruction form of a 32x32->64 product
KERNEL(k_mix_lo_hi, uint32_t lo[CHAINS]; uint32_t hi[CHAINS], lo[j] = b + j; hi[j] = b ^ j,
       asm volatile("mad.lo.u32 %0, %2, %3, %0;\n\tmad.hi.u32 %1, %2, %3, %1;" : "+r"(lo[j]), "+r"(hi[j]) : "r"(lo[(j + 1) & 7] | 0x80000000u), "r"(b | 0x80000000u)),
       r ^= lo[j] ^ hi[j])

// one "row" of an Fq multiplication: 6 carry-chained wide MACs (mad.lo.cc/madc.hi.cc pairs -> IMAD.WIDE.U32.X),
// 4 independent rows per iteration = 24 wide MACs, nothing else in the loop
__global__ void __launch_bounds__(256) k_cmad_rows(uint32_t *out, long long *cyc, int iters) {
    uint32_t b = blockIdx.x * 40503u + threadIdx.x * 2654435761u + 7u;
    uint32_t acc[4][12], a[6];
#pragma unroll
    for (int j = 0; j < 6; j++) a[j] = b * (j + 3) + 11u;
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int j = 0; j < 12; j++) acc[r][j] = b + r * 12 + j;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            uint32_t m = acc[(r + 1) & 3][0];
            asm volatile(
                "mad.lo.cc.u32 %0, %12, %18, %0;\n\tmadc.hi.cc.u32 %1, %12, %18, %1;\n\t"
                "madc.lo.cc.u32 %2, %13, %18, %2;\n\tmadc.hi.cc.u32 %3, %13, %18, %3;\n\t"
                "madc.lo.cc.u32 %4, %14, %18, %4;\n\tmadc.hi.cc.u32 %5, %14, %18, %5;\n\t"
                "madc.lo.cc.u32 %6, %15, %18, %6;\n\tmadc.hi.cc.u32 %7, %15, %18, %7;\n\t"
                "madc.lo.cc.u32 %8, %16, %18, %8;\n\tmadc.hi.cc.u32 %9, %16, %18, %9;\n\t"
                "madc.lo.cc.u32 %10, %17, %18, %10;\n\tmadc.hi.u32 %11, %17, %18, %11;"
                : "+r"(acc[r][0]), "+r"(acc[r][1]), "+r"(acc[r][2]), "+r"(acc[r][3]), "+r"(acc[r][4]), "+r"(acc[r][5]),
                  "+r"(acc[r][6]), "+r"(acc[r][7]), "+r"(acc[r][8]), "+r"(acc[r][9]), "+r"(acc[r][10]), "+r"(acc[r][11])
                : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(m));
        }
    }
    long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int j = 0; j < 12; j++) r ^= acc[q][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// the engine's own Fq Montgomery multiplication in a dependent chain, registers only: the practical ceiling
// of any kernel built from it.  Counted as 300 wide MACs per multiplication (SURVEY.md 8d unit).
__global__ void __launch_bounds__(256) k_fp_mul_chain(uint32_t *out, long long *cyc, int iters) {
    b381::fp x, y;
#pragma unroll
    for (int j = 0; j < 12; j++) { x.l[j] = (blockIdx.x * 977u + threadIdx.x * 131u + j) & 0x0fffffffu; y.l[j] = (threadIdx.x * 7919u + j * 13u) & 0x0fffffffu; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) { b381::fp_mul(x, x, y); }
    long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int j = 0; j < 12; j++) r ^= x.l[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
typedef void (*kern_t)(uint32_t *, long long *, int);

int main() {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { printf("{\"error\": \"no device\"}\n"); return 1; }
    const int sms = prop.multiProcessorCount, bps = 8, threads = 256, iters = 4096;
    const int blocks = sms * bps;
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaMalloc(&cyc, (size_t)blocks * 8);
    struct { const char *name; kern_t k; double ops_per_body; const char *sass; } ks[] = {
        {"cmad_rows_imad_wide_x", k_cmad_rows, 24.0 / CHAINS, "IMAD.WIDE.U32.X carry rows (4 x 6 wide MACs per iteration)"},
        {"fp_mul_chain", k_fp_mul_chain, 300.0 / CHAINS, "engine fp_mul, 300 wide MACs each (dependent chain, registers only)"},
        {"mad_wide_u32", k_mad_wide, 1, "IMAD.WIDE.U32 + IADD3 + IADD3.X (ptxas splits the 64-bit accumulate)"},
        {"mad_wide_u32_imm", k_mad_wide_imm, 1, "IMAD.WIDE.U32 imm + IADD3 + IADD3.X"},
        {"mad_lo_u32", k_mad_lo, 1, "IMAD"},
        {"mad_hi_u32", k_mad_hi, 1, "IMAD.HI.U32"},
        {"mad_lo_cc_madc_hi_pair", k_mad_cc_pair, 1, "one 32x32->64 MAC with carry (pair counted once)"},
        {"iadd3", k_iadd3, 1, "IADD3"},
        {"mix_wide_plus_iadd3", k_mix_wide_iadd, 2, "IMAD.WIDE.U32 + IADD3, counted as 2 ops"},
        {"mix_lo_plus_hi", k_mix_lo_hi, 2, "IMAD + IMAD.HI.U32, counted as 2 ops"},
    };
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, \"config\": \"%d blocks/SM x %d threads, %d chains/thread, %d iters\", \"results\": {",
           prop.name, sms, prop.clockRate, bps, threads, CHAINS, iters);
    for (size_t t = 0; t < sizeof ks / sizeof ks[0]; t++) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 6; rep++) {
            cudaEventRecord(e0);
            ks[t].k<<<blocks, threads>>>(out, cyc, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0) best = std::min(best, ms);
        }
        std::vector<long long> h(blocks);
        cudaMemcpy(h.data(), cyc, blocks * 8, cudaMemcpyDeviceToHost);
        std::sort(h.begin(), h.end());
        double med = (double)h[blocks / 2];
        double ops_block = (double)threads * iters * CHAINS * ks[t].ops_per_body;
        double per_clk_sm = ops_block * bps / med;
        double ops_s = ops_block * blocks / (best * 1e-3);
        printf("%s\"%s\": {\"sass\": \"%s\", \"thread_ops_per_clk_per_sm\": %.2f, \"tera_ops_per_s\": %.3f, \"ms\": %.4f, \"median_block_cycles\": %.0f, \"implied_mhz\": %.0f}",
               t ? ", " : "", ks[t].name, ks[t].sass, per_clk_sm, ops_s / 1e12, best, med, med / (best * 1e-3) / 1e6);
    }
    printf("}}\n");
    return 0;
}
