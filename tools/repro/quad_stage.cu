// Standalone reproducer: q_exp_by_x_main stage outputs, device (4 lanes per unit) vs the host emulation of the same source.
#include "../../bls_b200/csrc/quad.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
using namespace b381; using namespace b381::quad;
__global__ void k(const uint64_t *a, uint64_t *o, int stage, int n) {
    int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    if (i >= n) i = n - 1;
    q6 x, r; bool bad[1];
    q12_load(&x, a + 72 * i);
    q_exp_by_x_main(&r, bad, &x, 0xd201000000010000ULL, stage);
    q12_store(o + 72 * i, &r);
}
int main(int argc, char **argv) {
    const int n = 8;
    uint64_t h[72 * n], ho[72 * n], he[72 * n];
    uint64_t s = 88172645463325252ULL;
    for (int i = 0; i < 72 * n; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = s; if (i % 6 == 5) h[i] &= 0x0fffffffffffffffULL; }
    uint64_t *da, *dout;
    cudaMalloc(&da, sizeof h); cudaMalloc(&dout, sizeof h);
    cudaMemcpy(da, h, sizeof h, cudaMemcpyHostToDevice);
    int bad_total = 0;
    for (int stage = 1; stage <= 15; stage++) {
        k<<<1, 32>>>(da, dout, stage, n);
        cudaError_t e = cudaMemcpy(ho, dout, sizeof h, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 2; }
        for (int i = 0; i < n; i++) {
            q6 x, r; bool bad[QL];
            q12_load(&x, h + 72 * i);
            q_exp_by_x_main(&r, bad, &x, 0xd201000000010000ULL, stage);
            q12_store(he + 72 * i, &r);
        }
        printf("stage %2d:", stage);
        for (int c = 0; c < 6; c++) {
            int ok = 1;
            for (int i = 0; i < n; i++) ok &= memcmp(ho + 72 * i + 12 * c, he + 72 * i + 12 * c, 96) == 0;
            printf(" %s", ok ? "ok" : "XX");
            bad_total += !ok;
        }
        printf("\n");
    }
    return bad_total ? 1 : 0;
}
