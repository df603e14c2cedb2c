// micro-benchmark: latencies inside a 1024-thread CTA with 219 KB of dynamic shared memory (the crater kernel's launch shape)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ long long clk(unsigned dep)
{
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "r"(dep) : "memory");
    return t;
}
__global__ void __launch_bounds__(1024, 1) k(long long *out, int nact)
{
    extern __shared__ unsigned sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 50000; i += 1024) sm[i] = (i * 7919u + 13u) % 50000u;
    __syncthreads();
    long long t0, t1;
    long long res[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // T1/T4: chain of 32 dependent LDS, warps < nact active
    __syncthreads();
    if (warp < nact) {
        unsigned p = lane * 131u + warp;
        t0 = clk(p);
#pragma unroll
        for (int i = 0; i < 32; ++i) p = sm[p];
        t1 = clk(p);
        res[0] = (t1 - t0);
        if (p == 0xFFFFFFFFu) out[0] = 1;
    }
    __syncthreads();
    // T2: barrier, everybody arrives together, 8 in a row
    t0 = clk(0);
#pragma unroll
    for (int i = 0; i < 8; ++i) __syncthreads();
    t1 = clk(0);
    res[1] = (t1 - t0);
    // T3: 8 ballots + popc chain (dependent)
    __syncthreads();
    if (warp < nact) {
        unsigned v = lane * 2654435761u;
        t0 = clk(v);
#pragma unroll
        for (int i = 0; i < 8; ++i) { unsigned m = __ballot_sync(0xffffffffu, (v >> i) & 1u); v += __popc(m); }
        t1 = clk(v);
        res[2] = t1 - t0;
        if (v == 0x12345u) out[0] = 2;
    }
    __syncthreads();
    // T5: 8 independent gathers LDS.64 then a dependent min tree of doubles
    if (warp < nact) {
        const double *d = reinterpret_cast<const double *>(sm);
        unsigned p = (lane * 977u + warp * 31u) % 20000u;
        t0 = clk(p);
        double h[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = d[(p + i * 1237u) % 20000u];
        double m = fmin(fmin(fmin(h[0], h[1]), fmin(h[2], h[3])), fmin(fmin(h[4], h[5]), fmin(h[6], h[7])));
        t1 = clk((unsigned)__double2loint(m));
        res[3] = t1 - t0;
        if (m == 1.2345) out[0] = 3;
    }
    __syncthreads();
    // T6: barrier where half of the warps arrive late (after a 2000-cycle spin): cost seen by an early warp beyond the spin
    t0 = clk(0);
    if (warp >= 16) { while (clk(0) - t0 < 2000) { } }
    __syncthreads();
    t1 = clk(0);
    res[4] = t1 - t0;
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 17)) {
        long long *o = out + 8 + (warp ? 8 : 0);
        for (int i = 0; i < 8; ++i) o[i] = res[i];
    }
}
int main()
{
    long long *d, h[32];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 219 * 1024);
    for (int nact : {1, 4, 16, 32}) {
        cudaMemset(d, 0, sizeof(h));
        k<<<148, 1024, 219 * 1024>>>(d, nact);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("nact %2d (%s): 32 dependent LDS %lld cyc (%.1f each) | 8 barriers %lld (%.1f each) | 8 ballot+popc %lld (%.1f each) | "
               "8 gathers + fmin tree %lld | barrier with late half %lld (spin 2000)\n",
               nact, cudaGetErrorString(e), h[8], h[8] / 32.0, h[9], h[9] / 8.0, h[10], h[10] / 8.0, h[11], h[12]);
    }
    return 0;
}
