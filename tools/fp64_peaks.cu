// fp64_peaks.cu -- is the FP64 pipe of B200 a second multiplier for 384-bit arithmetic?
// Measures the issue rates of DFMA / DADD / 64-bit integer add, whether DFMA overlaps IMAD.WIDE, and the limb-product
// step of a double-precision big-integer multiplication (two round-to-zero FMAs split a 48x48-bit product into
// two 48-bit halves; the halves are accumulated as integers on the bit patterns).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peaks tools/fp64_peaks.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

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
#define FOLD64(x) r ^= (uint32_t)(x) ^ (uint32_t)((x) >> 32)

KERNEL(k_dfma, double acc[CHAINS]; double m = 1.0 + 1e-9 * (b & 255), acc[j] = 1.0 + j + (b & 15),
       asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(acc[j]) : "d"(m), "d"(acc[(j + 1) & 7])),
       FOLD64((uint64_t)__double_as_longlong(acc[j])))
KERNEL(k_dadd, double acc[CHAINS], acc[j] = 1.0 + j + (b & 15),
       asm volatile("add.rz.f64 %0, %0, %1;" : "+d"(acc[j]) : "d"(acc[(j + 1) & 7])),
       FOLD64((uint64_t)__double_as_longlong(acc[j])))
KERNEL(k_add64, uint64_t acc[CHAINS], acc[j] = (uint64_t)(b + j) * 0x9E3779B97F4A7C15ull,
       asm volatile("add.u64 %0, %0, %1;" : "+l"(acc[j]) : "l"(acc[(j + 1) & 7])),
       FOLD64(acc[j]))
KERNEL(k_mad_wide_cc, uint32_t lo[CHAINS]; uint32_t hi[CHAINS], lo[j] = b + j; hi[j] = b ^ j,
       asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[j]), "+r"(hi[j]) : "r"(lo[(j + 1) & 7]), "r"(b)),
       r ^= lo[j] ^ hi[j])
// DFMA + wide MAC in the same loop: separate pipes?
KERNEL(k_mix_dfma_wide, double acc[CHAINS]; uint32_t lo[CHAINS]; uint32_t hi[CHAINS]; double m = 1.0 + 1e-9 * (b & 255), acc[j] = 1.0 + j + (b & 15); lo[j] = b + j; hi[j] = b ^ j,
       asm volatile("fma.rz.f64 %0, %0, %3, %4;\n\tmad.lo.cc.u32 %1, %5, %6, %1;\n\tmadc.hi.u32 %2, %5, %6, %2;"
                    : "+d"(acc[j]), "+r"(lo[j]), "+r"(hi[j]) : "d"(m), "d"(acc[(j + 1) & 7]), "r"(lo[(j + 1) & 7]), "r"(b)),
       FOLD64((uint64_t)__double_as_longlong(acc[j])); r ^= lo[j] ^ hi[j])
// DFMA + 64-bit integer add
KERNEL(k_mix_dfma_add64, double acc[CHAINS]; uint64_t s[CHAINS]; double m = 1.0 + 1e-9 * (b & 255), acc[j] = 1.0 + j + (b & 15); s[j] = b + j,
       asm volatile("fma.rz.f64 %0, %0, %2, %3;\n\tadd.u64 %1, %1, %4;" : "+d"(acc[j]), "+l"(s[j]) : "d"(m), "d"(acc[(j + 1) & 7]), "l"(s[(j + 1) & 7])),
       FOLD64((uint64_t)__double_as_longlong(acc[j])); FOLD64(s[j]))
// the limb-product step: hi = rz(a*b + 2^100); lo = rz(a*b + (2^100 + 2^52 - hi)); accumulate the two bit patterns
__device__ __forceinline__ void limb_step(double x, double y, uint64_t &acch, uint64_t &accl) {
    double h, l, sub;
    asm volatile("fma.rz.f64 %0, %1, %2, 0d4630000000000000;" : "=d"(h) : "d"(x), "d"(y));          /* + 2^100 */
    asm volatile("sub.rz.f64 %0, 0d4630000000000010, %1;" : "=d"(sub) : "d"(h));                   /* 2^100 + 2^52 - h */
    asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(l) : "d"(x), "d"(y), "d"(sub));
    acch += (uint64_t)__double_as_longlong(h);
    accl += (uint64_t)__double_as_longlong(l);
}
#define LIMB_DECL uint64_t accl[CHAINS]; uint64_t acch[CHAINS]; double x[CHAINS]; double y = (double)((((uint64_t)b << 16) | 0x1234u) & 0xFFFFFFFFFFFFull)
#define LIMB_INIT accl[j] = j; acch[j] = b; x[j] = (double)(((uint64_t)(b + j) * 0x9E3779B9ull) & 0xFFFFFFFFFFFFull)
KERNEL(k_limb_step, LIMB_DECL, LIMB_INIT, limb_step(x[j], y, acch[j], accl[j]), FOLD64(accl[j]); FOLD64(acch[j]))

typedef void (*kern_t)(uint32_t *, long long *, int);
int main() {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("{\"error\": \"no device\"}\n"); return 1; }
    const int sms = prop.multiProcessorCount, bps = 8, threads = 256, iters = 4096, blocks = sms * bps;
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaMalloc(&cyc, (size_t)blocks * 8);
    struct { const char *name; kern_t k; double ops_per_body; const char *what; } ks[] = {
        {"dfma_rz", k_dfma, 1, "DFMA"},
        {"dadd_rz", k_dadd, 1, "DADD"},
        {"add_u64", k_add64, 1, "64-bit integer add"},
        {"mad_wide_cc", k_mad_wide_cc, 1, "32x32->64 MAC with carry (reference)"},
        {"mix_dfma_plus_wide", k_mix_dfma_wide, 2, "DFMA + wide MAC, counted as 2 ops"},
        {"mix_dfma_plus_add64", k_mix_dfma_add64, 2, "DFMA + add.u64, counted as 2 ops"},
        {"limb_step_48x48", k_limb_step, 1, "one 48x48-bit limb product: 2 DFMA + DADD + 2 add.u64 (counted once)"},
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
        printf("%s\"%s\": {\"what\": \"%s\", \"thread_ops_per_clk_per_sm\": %.2f, \"tera_ops_per_s\": %.3f, \"ms\": %.4f}",
               t ? ", " : "", ks[t].name, ks[t].what, ops_block * bps / med, ops_block * blocks / (best * 1e-3) / 1e12, best);
    }
    printf("}}\n");
    return 0;
}
