// vm_probe.cu -- how many warps per SM does a shared-memory-resident, warp-synchronous Fq schedule
// need to saturate the IMAD.WIDE pipe?  Each warp owns SLOTS 48-byte slots in shared memory; a step =
// every lane loads two operands (lane-dependent slots), multiplies (300 wide MACs), [adds a third],
// stores to its slot, __syncwarp.  Reports wide-MAC/s for 1..8 warps per scheduler.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/vm_probe tools/vm_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../bls_b200/csrc/fp.cuh"
using namespace b381;

#define SLOTS 96
__device__ __forceinline__ void lds(fp &r, const uint4 *s) { uint4 a = s[0], b = s[1], c = s[2];
    r.l[0]=a.x; r.l[1]=a.y; r.l[2]=a.z; r.l[3]=a.w; r.l[4]=b.x; r.l[5]=b.y; r.l[6]=b.z; r.l[7]=b.w; r.l[8]=c.x; r.l[9]=c.y; r.l[10]=c.z; r.l[11]=c.w; }
__device__ __forceinline__ void sts(uint4 *s, const fp &r) { s[0]=make_uint4(r.l[0],r.l[1],r.l[2],r.l[3]); s[1]=make_uint4(r.l[4],r.l[5],r.l[6],r.l[7]); s[2]=make_uint4(r.l[8],r.l[9],r.l[10],r.l[11]); }

template <int MODE>   // 0: registers only (chain), 1: smem operands + store, 2: smem + one extra add per mul
__global__ void k_probe(uint32_t *out, int iters) {
    extern __shared__ uint4 sm[];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint4 *slots = sm + (size_t)warp * SLOTS * 3;
    fp x, y;
    for (int j = 0; j < 12; j++) { x.l[j] = (blockIdx.x * 977u + threadIdx.x * 131u + j) & 0x0fffffffu; y.l[j] = (threadIdx.x * 7919u + j * 13u) & 0x0fffffffu; }
    for (int s = lane; s < SLOTS; s += 32) sts(slots + 3 * s, x);
    __syncwarp();
    int ia = (lane * 7 + 3) % SLOTS, ib = (lane * 13 + 5) % SLOTS, id = 32 + lane;
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) { fp_mul_inl(x, x, y); }
        else {
            fp a, b; lds(a, slots + 3 * ia); lds(b, slots + 3 * ib);
            fp_mul_inl(a, a, b);
            if (MODE == 2) { fp c; lds(c, slots + 3 * ((ia + 9) % SLOTS)); fp_add(a, a, c); }
            __syncwarp();
            sts(slots + 3 * id, a);
            __syncwarp();
            ia = (ia + 1) % SLOTS; ib = (ib + 3) % SLOTS;
        }
    }
    uint32_t r = 0;
    if (MODE == 0) { for (int j = 0; j < 12; j++) r ^= x.l[j]; } else { fp a; lds(a, slots + 3 * id); for (int j = 0; j < 12; j++) r ^= a.l[j]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE> void run(const char *name, int sms, uint32_t *out) {
    int iters = 2000;
    printf("\"%s\": {", name);
    for (int wps = 1; wps <= 8; wps++) {       // warps per scheduler; block = 4 * wps warps, one block per SM
        int threads = 128 * wps; if (threads > 1024) break;
        size_t smem = (size_t)(threads / 32) * SLOTS * 48;
        cudaFuncSetAttribute(k_probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0); k_probe<MODE><<<sms, threads, smem>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) best = best < ms ? best : ms;
        }
        cudaError_t e = cudaGetLastError();
        double macs = (double)sms * threads * iters * 300.0;
        printf("%s\"%d\": %.3f", wps > 1 ? ", " : "", wps, e == cudaSuccess ? macs / (best * 1e-3) / 1e12 : -1.0);
    }
    printf("}");
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    uint32_t *out; cudaMalloc(&out, (size_t)prop.multiProcessorCount * 1024 * 4);
    printf("{\"unit\": \"T wide-MAC/s by warps per scheduler (one block per SM)\", ");
    run<0>("regs_chain", prop.multiProcessorCount, out); printf(", ");
    run<1>("smem_step", prop.multiProcessorCount, out); printf(", ");
    run<2>("smem_step_plus_add", prop.multiProcessorCount, out);
    printf("}\n");
    return 0;
}
