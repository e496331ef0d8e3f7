// Microbenchmark behind DESIGN.md "persistent cooperative PCG": what does a grid-wide barrier cost on B200 against the gap
// between two kernels of a CUDA graph?  A CG iteration is three dependent phases (update + dot, direction, apply); a persistent
// kernel would replace three graph launches by three grid barriers.
//   (a) cooperative kernel, G CTAs x 128 threads, N x cooperative_groups::grid.sync()            -> us per barrier
//   (b) CUDA graph of N empty kernels of the same geometry (stream-ordered dependencies)         -> us per launch
//   (c) the same with programmatic dependent launch between the kernels                          -> us per launch
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o grid_sync grid_sync.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
namespace cg = cooperative_groups;
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void sync_kernel(int n, double *out)
{
    cg::grid_group g = cg::this_grid();
    double s = threadIdx.x;
    for (int i = 0; i < n; i++) { s = s * 1.0000001 + 1.0; g.sync(); }
    if (threadIdx.x == 0) out[blockIdx.x] = s;
}
__global__ void tiny_kernel(double *out, int pdl)
{
    if (pdl) { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
    if (threadIdx.x == 0) out[blockIdx.x] += 1.0;
}

static float time_graph(int grid, int n, double *out, bool pdl)
{
    cudaStream_t s; CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaGraph_t graph; cudaGraphExec_t exec;
    CHECK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < n; i++) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
        CHECK(cudaLaunchKernelEx(&cfg, tiny_kernel, out, (int)pdl));
    }
    CHECK(cudaStreamEndCapture(s, &graph));
    CHECK(cudaGraphInstantiate(&exec, graph, 0));
    cudaEvent_t e0, e1; CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    CHECK(cudaGraphLaunch(exec, s)); CHECK(cudaStreamSynchronize(s));
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        CHECK(cudaEventRecord(e0, s)); CHECK(cudaGraphLaunch(exec, s)); CHECK(cudaEventRecord(e1, s)); CHECK(cudaEventSynchronize(e1));
        float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    int sms = 0; CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    double *out; CHECK(cudaMalloc(&out, sizeof(double) * 4096)); CHECK(cudaMemset(out, 0, sizeof(double) * 4096));
    const int n = 2000;
    for (int per_sm : {1, 2, 4}) {
        const int grid = sms * per_sm;
        cudaEvent_t e0, e1; CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
        int nn = n; void *args[] = {&nn, &out};
        CHECK(cudaLaunchCooperativeKernel((void *)sync_kernel, dim3(grid), dim3(128), args, 0, 0)); CHECK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < 5; r++) {
            CHECK(cudaEventRecord(e0)); CHECK(cudaLaunchCooperativeKernel((void *)sync_kernel, dim3(grid), dim3(128), args, 0, 0));
            CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1));
            float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        const float tg = time_graph(grid, n, out, false), tp = time_graph(grid, n, out, true);
        printf("grid %4d x 128 threads: grid.sync %.2f us   graph launch of an empty kernel %.2f us   with PDL %.2f us\n", grid, 1e3 * best / n, 1e3 * tg / n, 1e3 * tp / n);
    }
    return 0;
}
