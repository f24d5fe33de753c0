// micro-benchmarks behind DESIGN.md's latency figures (B200, sm_100a): dependent-issue latency and issue rate of the
// FP64 pipe, shared-memory round trip, barrier and shuffle costs.  nvcc -arch=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dep_chain(double* out, long long* cyc, int n, int kind)
{
    double a = out[0], b = out[1], c = out[2];
    long long t0 = clock64();
    if (kind == 0) for (int i = 0; i < n; ++i) a = __fma_rn(a, b, c);
    else if (kind == 1) for (int i = 0; i < n; ++i) a = __dadd_rn(a, b);
    else if (kind == 2) for (int i = 0; i < n; ++i) a = __dmul_rn(a, b);
    else if (kind == 3) for (int i = 0; i < n; ++i) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a)); a = r; }
    else if (kind == 4) for (int i = 0; i < n; ++i) a = __shfl_down_sync(0xffffffffu, a, 1);
    else if (kind == 5) for (int i = 0; i < n; ++i) a = (a > b) ? a + c : a - c;     // DSETP + select + DADD
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[3 + threadIdx.x % 4] = a;
}

// throughput: every thread runs ILP independent DFMA chains; blocks of `threads`
template <int ILP>
__global__ void thr_chain(double* out, long long* cyc, int n)
{
    double a[ILP];
    const double b = out[1], c = out[2];
#pragma unroll
    for (int k = 0; k < ILP; ++k) a[k] = out[0] + k;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) a[k] = __fma_rn(a[k], b, c);
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += a[k];
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[8 + (threadIdx.x & 7)] = s;
}

__global__ void smem_chain(double* out, long long* cyc, int n)
{
    __shared__ double sh[1024];
    sh[threadIdx.x] = out[0];
    __syncthreads();
    double a = 0.0;
    int idx = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { a += sh[idx]; idx = (idx + (int)a) & 1023; }     // load -> address dependency
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[4] = a;
}

__global__ void sync_cost(double* out, long long* cyc, int n, int warp_only)
{
    __shared__ double sh[1024];
    sh[threadIdx.x] = threadIdx.x;
    __syncthreads();
    double a = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        sh[threadIdx.x] = a + 1.0;
        if (warp_only) __syncwarp(); else __syncthreads();
        a = sh[(threadIdx.x + 1) & (blockDim.x - 1) & (warp_only ? 31 : 1023)];
        if (warp_only) __syncwarp(); else __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[5] = a;
}

int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 64 * 8); cudaMalloc(&cyc, 8);
    double h[64]; for (int i = 0; i < 64; ++i) h[i] = 1.0 + 1e-9 * i;
    h[1] = 1.0000001; h[2] = 1e-9;
    cudaMemcpy(out, h, sizeof h, cudaMemcpyHostToDevice);
    long long c;
    const int n = 20000;
    const char* names[] = {"DFMA", "DADD", "DMUL", "MUFU.RCP64H(+move)", "SHFL.f64 (2 x SHFL)", "DSETP+FSEL+DADD"};
    for (int kind = 0; kind < 6; ++kind) {
        dep_chain<<<1, 32>>>(out, cyc, n, kind); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        dep_chain<<<1, 32>>>(out, cyc, n, kind); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("dependent %-22s %.2f cycles/op\n", names[kind], (double)c / n);
    }
    int thr[] = {32, 128, 256, 512, 1024};
    for (int t : thr) {
        thr_chain<1><<<1, t>>>(out, cyc, n); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        double r1 = (double)t * n / c;
        thr_chain<4><<<1, t>>>(out, cyc, n); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        double r4 = (double)t * n * 4 / c;
        thr_chain<8><<<1, t>>>(out, cyc, n); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        double r8 = (double)t * n * 8 / c;
        printf("DFMA issue, 1 CTA of %4d threads: ILP1 %.1f  ILP4 %.1f  ILP8 %.1f thread-DFMA / clk / SM\n", t, r1, r4, r8);
    }
    smem_chain<<<1, 32>>>(out, cyc, n); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("LDS.64 -> DADD -> address chain: %.1f cycles / trip\n", (double)c / n);
    int st[] = {32, 128, 512, 1024};
    for (int t : st) {
        sync_cost<<<1, t>>>(out, cyc, 2000, 0); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("STS + __syncthreads + LDS + __syncthreads, %4d threads: %.1f cycles / round\n", t, (double)c / 2000);
    }
    sync_cost<<<1, 32>>>(out, cyc, 2000, 1); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("STS + __syncwarp + LDS + __syncwarp, 1 warp: %.1f cycles / round\n", (double)c / 2000);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("clock rate attribute %d kHz\n", clk);
    return 0;
}
