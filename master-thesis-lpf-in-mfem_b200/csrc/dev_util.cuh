// Small helpers shared by every translation unit of the device layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/lpf_b200.h"
#include "../host/lpf_common.hpp"

#define LPF_MAXP LPF_MAX_ORDER

#define CUDA_TRY(call)                                                                              \
    do {                                                                                            \
        cudaError_t err__ = (call);                                                                 \
        if (err__ != cudaSuccess) {                                                                 \
            lpf::set_error(std::string(#call) + ": " + cudaGetErrorString(err__) + " (" __FILE__ ":" + \
                           std::to_string(__LINE__) + ")");                                         \
            return LPF_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

#define LPF_TRY(call)                    \
    do {                                 \
        int rc__ = (call);               \
        if (rc__ != LPF_OK) return rc__; \
    } while (0)

// Kernel launch with (optionally) the programmatic-dependent-launch attribute; see griddep_wait() below.
template <class... KArgs, class... Args>
inline cudaError_t launch_ex(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

#ifdef __CUDACC__
// Programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start while its predecessor in the stream is still draining; griddep_wait() blocks until the predecessor has
// completed and its writes are visible, griddep_launch() lets the successor's CTAs be scheduled early.  Both are
// no-ops for a kernel launched without the attribute.  Every kernel of the PCG iteration calls them, which hides
// most of the 2-3 us launch + prologue latency between the kernels of an iteration (small meshes are launch-bound).
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Scatter-add without a return value: always the fire-and-forget reduction (SASS REDG), never the round-trip ATOMG.
// (With plain atomicAdd the compiler switched to ATOMG as soon as the kernel also read y elsewhere -- the fused
// halo tail -- which cost 30 % of the apply kernel's bandwidth; profiles/r01_apply_ncu.md.)
__device__ __forceinline__ void red_add_f64(double *addr, double v)
{
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

// Sum over the lanes of a warp that may be only partly populated (block sizes here are Q^2 E, rarely a multiple of
// 32): the mask names exactly the lanes that exist, absent partners contribute 0.
__device__ __forceinline__ double warp_sum_partial(double v, int nthreads)
{
    const int lane = threadIdx.x & 31, base = threadIdx.x & ~31;
    const int live = min(32, nthreads - base);
    const unsigned mask = live >= 32 ? 0xffffffffu : ((1u << live) - 1u);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int src = lane ^ o;
        const double t = __shfl_sync(mask, v, src < live ? src : lane);
        if (src < live) v += t;
    }
    return v;
}
#endif
