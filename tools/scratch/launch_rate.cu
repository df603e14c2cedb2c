// microbenchmark: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o launch_rate launch_rate.cu (2.8 us per back-to-back launch on the GPU box)
#include <cstdio>
#include <chrono>
#include <cuda_runtime.h>
struct Big { double a[96]; };
__global__ void k_empty(Big b, int *x) { if (b.a[0] == 123.0 && threadIdx.x == 0) x[0] = 1; }
__global__ void k_small(int *x) { if (threadIdx.x == 1025) x[0] = 1; }
int main() {
    int *d; cudaMalloc(&d, 4);
    Big b; for (int i = 0; i < 96; ++i) b.a[i] = i;
    cudaStream_t st; cudaStreamCreate(&st);
    for (int rep = 0; rep < 3; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < 10000; ++i) k_empty<<<1184, 256, 0, st>>>(b, d);
        auto t1 = std::chrono::steady_clock::now();
        cudaStreamSynchronize(st);
        auto t2 = std::chrono::steady_clock::now();
        printf("big-param 1184x256: issue %.2f us/launch, total %.2f us/launch\n",
               std::chrono::duration<double, std::micro>(t1 - t0).count() / 1e4, std::chrono::duration<double, std::micro>(t2 - t0).count() / 1e4);
        t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < 10000; ++i) k_small<<<1, 32, 0, st>>>(d);
        t1 = std::chrono::steady_clock::now();
        cudaStreamSynchronize(st);
        t2 = std::chrono::steady_clock::now();
        printf("small 1x32: issue %.2f us/launch, total %.2f us/launch\n",
               std::chrono::duration<double, std::micro>(t1 - t0).count() / 1e4, std::chrono::duration<double, std::micro>(t2 - t0).count() / 1e4);
    }
    return 0;
}
