// Dependent-issue latency of FP64 / MUFU / shared-memory operations on one warp
// (cycles per dependent operation), and the same with 2 and 4 independent chains.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int CH>
__global__ void dfma_chain(double *out, long long *cyc, double m, double c)
{
    double a[CH];
    for (int k = 0; k < CH; ++k) a[k] = threadIdx.x * 1e-9 + k;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int k = 0; k < CH; ++k) a[k] = fma(a[k], m, c);
    long long t1 = clock64();
    double s = 0; for (int k = 0; k < CH; ++k) s += a[k];
    out[threadIdx.x] = s; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void dadd_chain(double *out, long long *cyc, double c)
{
    double a = threadIdx.x * 1e-9;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = a + c;
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void rcp_chain(double *out, long long *cyc)
{
    double a = 1.5 + threadIdx.x * 1e-9;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(a));
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void lds_chain(double *out, long long *cyc)
{
    __shared__ int idx[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i + 33) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) p = idx[p];
    long long t1 = clock64();
    out[threadIdx.x] = p; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void ffma_chain(double *out, long long *cyc, float m, float c)
{
    float a = threadIdx.x * 1e-9f;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = fmaf(a, m, c);
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
#define RUN(name, launch, per) for (int r = 0; r < 2; ++r) { launch; cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); if (r) printf("%-28s %.2f cycles per step (%d ops/step)\n", name, (double)h / N, per); }
    RUN("DFMA 1 chain, 1 warp", (dfma_chain<1><<<1, 32>>>(out, cyc, 0.999999, 1e-7)), 1)
    RUN("DFMA 2 chains, 1 warp", (dfma_chain<2><<<1, 32>>>(out, cyc, 0.999999, 1e-7)), 2)
    RUN("DFMA 4 chains, 1 warp", (dfma_chain<4><<<1, 32>>>(out, cyc, 0.999999, 1e-7)), 4)
    RUN("DFMA 8 chains, 1 warp", (dfma_chain<8><<<1, 32>>>(out, cyc, 0.999999, 1e-7)), 8)
    RUN("DFMA 1 chain, 4 warps", (dfma_chain<1><<<1, 128>>>(out, cyc, 0.999999, 1e-7)), 1)
    RUN("DFMA 1 chain, 8 warps", (dfma_chain<1><<<1, 256>>>(out, cyc, 0.999999, 1e-7)), 1)
    RUN("DFMA 1 chain, 16 warps", (dfma_chain<1><<<1, 512>>>(out, cyc, 0.999999, 1e-7)), 1)
    RUN("DADD 1 chain, 1 warp", (dadd_chain<<<1, 32>>>(out, cyc, 1e-7)), 1)
    RUN("MUFU.RCP64H chain, 1 warp", (rcp_chain<<<1, 32>>>(out, cyc)), 1)
    RUN("LDS pointer chase, 1 warp", (lds_chain<<<1, 32>>>(out, cyc)), 1)
    RUN("FFMA 1 chain, 1 warp", (ffma_chain<<<1, 32>>>(out, cyc, 0.999f, 1e-7f)), 1)
    return 0;
}
