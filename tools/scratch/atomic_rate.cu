// microbenchmark: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o atomic_rate atomic_rate.cu (profiles/README.md: 0.7 ns per same-address returning atomic)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_same(unsigned long long *ctr, unsigned *out, int per_warp, int naddr)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (lane == 0) {
        unsigned long long acc = 0;
        for (int i = 0; i < per_warp; ++i) acc += atomicAdd(ctr + 16 * ((warp + i) % naddr), 1ull);
        out[warp] = (unsigned)acc;
    }
}
int main()
{
    unsigned long long *ctr; unsigned *out;
    cudaMalloc(&ctr, 16 * 8 * 4096); cudaMalloc(&out, 4 * 1184 * 8);
    cudaMemset(ctr, 0, 16 * 8 * 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int naddrs[] = {1, 3, 64, 1024};
    for (int na : naddrs)
        for (int pw : {1, 2, 8}) {
            k_same<<<1184, 256>>>(ctr, out, pw, na);
            cudaEventRecord(e0);
            for (int r = 0; r < 20; ++r) k_same<<<1184, 256>>>(ctr, out, pw, na);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("addresses %4d, %d returning atomics per warp (9472 warps): %.2f us per kernel, %.2f ns per atomic\n", na, pw,
                   ms * 1e3 / 20, ms * 1e6 / 20 / (9472.0 * pw));
        }
    return 0;
}
