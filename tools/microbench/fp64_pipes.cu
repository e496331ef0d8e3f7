// Microbenchmark behind DESIGN.md "FP64 tensor cores (DMMA)": is mma.sync.*.f64 on B200 a second FP64 pipe next to DFMA,
// and how fast is it?  Three kernels of the same length: DFMA only, DMMA only (m8n8k4 and m16n8k8), and both interleaved
// in every warp.  If the mixed kernel takes max(t_dfma, t_dmma) the pipes are independent; if it takes the sum they share
// the FP64 datapath.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int NCHAIN = 8;     // independent accumulators per thread

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// MODE 0: DFMA, 1: DMMA m8n8k4, 2: both interleaved, 3: DMMA m16n8k8, 4: DFMA + m16n8k8
template <int MODE>
__global__ void __launch_bounds__(256) pipes_kernel(double *out, int iters, double seed)
{
    double f[NCHAIN], c0[NCHAIN], c1[NCHAIN], c4[NCHAIN / 2][4];
    const double a = seed + threadIdx.x * 1e-9, b = 1.0 - 1e-9 * threadIdx.x;
    double a4[4] = {a, b, a, b}, b2[2] = {b, a};
#pragma unroll
    for (int i = 0; i < NCHAIN; i++) { f[i] = i; c0[i] = i; c1[i] = -i; }
#pragma unroll
    for (int i = 0; i < NCHAIN / 2; i++) for (int k = 0; k < 4; k++) c4[i][k] = i + k;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NCHAIN; i++) {
            if (MODE == 0 || MODE == 2 || MODE == 4) f[i] = fma(f[i], a, b);
            if (MODE == 1 || MODE == 2) dmma884(c0[i], c1[i], a, b);
            if ((MODE == 3 || MODE == 4) && (i & 1) == 0) dmma1688(c4[i / 2], a4, b2);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NCHAIN; i++) s += f[i] + c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < NCHAIN / 2; i++) for (int k = 0; k < 4; k++) s += c4[i][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
float run(double *out, int grid, int iters)
{
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    pipes_kernel<MODE><<<grid, 256>>>(out, iters / 10, 0.5);
    CHECK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        CHECK(cudaEventRecord(e0));
        pipes_kernel<MODE><<<grid, 256>>>(out, iters, 0.5);
        CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1));
        float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    int sms = 0;
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = sms * 4, iters = 20000;
    double *out;
    CHECK(cudaMalloc(&out, sizeof(double) * grid * 256));
    const double thr = (double)grid * 256, n = (double)iters * NCHAIN;
    const float t0 = run<0>(out, grid, iters), t1 = run<1>(out, grid, iters), t2 = run<2>(out, grid, iters),
                t3 = run<3>(out, grid, iters), t4 = run<4>(out, grid, iters);
    // flops: DFMA 2 per thread-instruction; m8n8k4 = 8*8*4*2 = 512 per warp-instruction = 16 per thread;
    // m16n8k8 = 16*8*8*2 = 2048 per warp = 64 per thread (issued for every second chain)
    printf("SMs %d, grid %d x 256 threads, %d iterations x %d chains\n", sms, grid, iters, NCHAIN);
    printf("DFMA only            : %8.3f ms  %7.2f TFLOP/s\n", t0, thr * n * 2 / t0 / 1e9);
    printf("DMMA m8n8k4 only     : %8.3f ms  %7.2f TFLOP/s\n", t1, thr * n * 16 / t1 / 1e9);
    printf("DFMA + m8n8k4 mixed  : %8.3f ms  (sum %.3f, max %.3f)  %7.2f TFLOP/s combined\n", t2, t0 + t1, t0 > t1 ? t0 : t1, thr * n * 18 / t2 / 1e9);
    printf("DMMA m16n8k8 only    : %8.3f ms  %7.2f TFLOP/s\n", t3, thr * (n / 2) * 64 / t3 / 1e9);
    printf("DFMA + m16n8k8 mixed : %8.3f ms  (sum %.3f, max %.3f)  %7.2f TFLOP/s combined\n", t4, t0 + t3, t0 > t3 ? t0 : t3, thr * (n * 2 + n / 2 * 64) / t4 / 1e9);
    return 0;
}
