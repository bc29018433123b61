// occ_sweep.cu -- how many resident warps does the engine's Fq multiplier need to saturate the
// IMAD.WIDE pipe?  Register-only dependent chains of fp_mul (300 wide MACs) and fp_dot2 (444 wide
// MACs) at 1..16 warps per SM sub-partition, plus two independent chains per thread (ILP 2).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/occ_sweep tools/occ_sweep.cu
// Run under gpurun: ./tools/occ_sweep > gpurun_out/occ_sweep.json
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../bls_b200/csrc/fp.cuh"
using namespace b381;

__device__ __forceinline__ void seed(fp &x, uint32_t s) {
#pragma unroll
    for (int j = 0; j < 12; j++) x.l[j] = (s * 2654435761u + j * 40503u) & 0x0fffffffu;
}
__device__ __forceinline__ uint32_t fold(const fp &x) {
    uint32_t r = 0;
#pragma unroll
    for (int j = 0; j < 12; j++) r ^= x.l[j];
    return r;
}

__global__ void __launch_bounds__(512) k_mul1(uint32_t *out, int iters) {
    fp x, y; seed(x, threadIdx.x + 1); seed(y, blockIdx.x + 77);
#pragma unroll 1
    for (int i = 0; i < iters; i++) fp_mul_inl(x, x, y);
    out[blockIdx.x * blockDim.x + threadIdx.x] = fold(x);
}
__global__ void __launch_bounds__(512) k_mul2(uint32_t *out, int iters) {
    fp x, y, z; seed(x, threadIdx.x + 1); seed(y, blockIdx.x + 77); seed(z, threadIdx.x + 999);
#pragma unroll 1
    for (int i = 0; i < iters; i++) { fp_mul_inl(x, x, y); fp_mul_inl(z, z, y); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = fold(x) ^ fold(z);
}
__global__ void __launch_bounds__(512) k_dot1(uint32_t *out, int iters) {
    fp x, y, z; seed(x, threadIdx.x + 1); seed(y, blockIdx.x + 77); seed(z, threadIdx.x + 999);
#pragma unroll 1
    for (int i = 0; i < iters; i++) fp_dot2_inl(x, x, y, z, y);
    out[blockIdx.x * blockDim.x + threadIdx.x] = fold(x);
}
__global__ void __launch_bounds__(512) k_mulv(uint32_t *out, int iters) {
    fp x, y; seed(x, threadIdx.x + 1); seed(y, blockIdx.x + 77);
#pragma unroll 1
    for (int i = 0; i < iters; i++) x = fp_mul_v(x, y);
    out[blockIdx.x * blockDim.x + threadIdx.x] = fold(x);
}

typedef void (*kern_t)(uint32_t *, int);

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    uint32_t *out; cudaMalloc(&out, (size_t)sms * 2048 * 4 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct { const char *name; kern_t k; double macs; } ks[] = {
        {"fp_mul_chain", k_mul1, 300.0}, {"fp_mul_2chains", k_mul2, 600.0}, {"fp_dot2_chain", k_dot1, 444.0},
        {"fp_mul_v_call_chain", k_mulv, 300.0}};
    int warps_per_sm[] = {4, 8, 16, 32, 64};
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"results\": [\n", p.name, sms);
    bool first = true;
    for (auto &k : ks)
        for (int w : warps_per_sm) {
            // w warps per SM: blocks of 128 threads (one warp per sub-partition), w/4 blocks per SM
            int block = 128, grid = sms * (w / 4), iters = 4000;
            k.k<<<grid, block>>>(out, 200);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            k.k<<<grid, block>>>(out, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double t = (double)grid * block * iters * k.macs / (ms * 1e-3) / 1e12;
            printf("%s{\"kernel\": \"%s\", \"warps_per_sm\": %d, \"ms\": %.3f, \"tera_mac_per_s\": %.3f}", first ? "" : ",\n", k.name, w, ms, t);
            first = false;
        }
    printf("\n]}\n");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "cuda error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
