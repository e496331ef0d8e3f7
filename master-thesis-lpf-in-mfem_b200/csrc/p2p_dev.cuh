// Device-side pieces of the peer-memory (NVLink / NVSwitch) exchange that are inlined into other kernels: LL lines,
// the warp-level scalar all-reduce and the halo-sum that rides on the apply kernel.  Protocol: see p2p.cuh.
#pragma once
#include "dev_util.cuh"
#include "p2p_types.hpp"

// ---- LL ("low latency") lines: 8 data bytes + two copies of a 32-bit sequence flag in ONE 16-byte store.  The
// receiver polls the line itself; when both flags carry the expected sequence number both data halves have
// landed (NVLink guarantees 8-byte atomicity), so no fence, no separate flag write and no second round trip is
// needed -- the exchange costs one one-way NVLink latency (same idea as NCCL's LL protocol).
__device__ __forceinline__ void ll_store(uint4 *p, double v, uint32_t flag)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((uint32_t)b), "r"(flag),
                 "r"((uint32_t)(b >> 32)), "r"(flag) : "memory");
}
__device__ __forceinline__ double ll_poll(const uint4 *p, uint32_t flag, int *err)
{
    uint32_t a, fa, b, fb;
    long long n = 0;
    for (;;) {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(fa), "=r"(b), "=r"(fb) : "l"(p) : "memory");
        if (fa == flag && fb == flag) break;
        if (++n > LPF_P2P_SPIN_LIMIT) { *err = 1; return 0.0; }
    }
    return __longlong_as_double((long long)(((unsigned long long)b << 32) | a));
}
__device__ __forceinline__ uint32_t ll_flag(unsigned long long seq) { return (uint32_t)(seq % 0xFFFFFFFFull) + 1u; }

// A timed-out wait must not go unnoticed (the sums it fed are wrong): besides P2PLocal::error, which the host reads with
// lpf_p2p_error(), the PCG state is put into a failure status so that every later kernel of the solve is a no-op and
// pcg_run() returns LPF_ERR_COMM.
__device__ __forceinline__ void p2p_flag_error(const P2PDev &d, PcgState *st)
{
    if (st != nullptr && *(volatile int *)&d.local->error) { st->comm_error = 1; st->status = PCG_COMM_ERROR; }
}

// Sum of one double over all ranks, executed by one full warp: lane r writes this rank's value as one LL line into
// rank r's box and polls the line rank r writes into mine.  Returns the sum (identical bits on every rank: fixed
// rank order) in every lane.  Two parities: a rank can start exchange k+2 only after every peer finished k.
__device__ __forceinline__ double p2p_allreduce_warp(const P2PDev &d, double local_val)
{
    const int lane = threadIdx.x & 31;
    const unsigned long long seq = d.local->red_seq + 1;
    const int par = (int)(seq & 1);
    const uint32_t flag = ll_flag(seq);
    __syncwarp();
    double got = 0.0;
    if (lane < d.nranks) {
        ll_store(&d.peers[lane]->red_ll[par][d.rank], local_val, flag);
        got = ll_poll(&d.mine->red_ll[par][lane], flag, &d.local->error);
    }
    double s = 0.0;
    for (int r = 0; r < d.nranks; r++) s += __shfl_sync(0xffffffffu, got, r);
    __syncwarp();
    if (lane == 0) d.local->red_seq = seq;
    __syncwarp();
    return s;
}

__device__ __forceinline__ void p2p_send_entry(const P2PPlanDev &h, int par, uint32_t flag, int i, double own)
{
    for (int k = h.snd_off[i]; k < h.snd_off[i + 1]; k++) ll_store(h.ll_dst[par * h.n_nbr + h.snd_nbr[k]] + h.snd_pos[k], own, flag);
}
__device__ __forceinline__ double p2p_recv_entry(const P2PPlanDev &h, const uint4 *recv, uint32_t flag, int i, double own, int *err)
{
    double s = 0.0;
    for (int j = h.red_off[i]; j < h.red_off[i + 1]; j++) s += (h.red_src[j] < 0) ? own : ll_poll(recv + h.red_src[j], flag, err);
    return s;
}

// ---- halo-sum + (d, A d) all-reduce fused into the apply kernel, mode 1: everything in the CTA that arrives last ----
// The CTA that arrives last (all scatter-adds of this rank are then globally visible) packs the interface values straight
// into the neighbours' mailboxes, waits for theirs, adds the partial sums in rank order and all-reduces the PCG
// denominator -- the collective rides on the compute kernel, no extra launch, no NCCL.  One CTA works through the whole
// interface, so this pays only for small interfaces (option p2p_fuse = 1).
// (sflag: one int of the caller's dynamic shared memory -- the apply kernels keep their static shared memory at zero, the
// order-8 CTA fills its third of the SM to within a few bytes)
__device__ __forceinline__ void p2p_apply_tail_last(const P2PTail &t, double *__restrict__ y, volatile int *sflag)
{
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int n = atomicInc(t.done, gridDim.x - 1);
        *sflag = (n == gridDim.x - 1);
    }
    __syncthreads();
    if (!*sflag) return;
    __threadfence();
    const int tid = threadIdx.x, nt = blockDim.x;
    const P2PPlanDev &h = t.h;
    // same LL lines and sequence numbers as p2p_halo_ll_kernel, so a rank with a small interface (tail) and a
    // neighbour with a large one (separate kernel) interoperate
    const unsigned long long seq = t.d.local->ll_seq[0] + 1;
    const int par = (int)(seq & 1);
    const uint32_t flag = ll_flag(seq);
    for (int i = tid; i < h.n_shared; i += nt) p2p_send_entry(h, par, flag, i, __ldcg(y + h.shared[i]));
    if (t.with_den && tid < 32) {
        double s = 0.0;
        for (int i = tid; i < LPF_DEN_SLOTS; i += 32) { s += __ldcg(t.den_slots + i); t.den_slots[i] = 0.0; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const double tot = p2p_allreduce_warp(t.d, s);
        if (tid == 0) t.st->red[1] = tot;
    }
    const uint4 *recv = h.ll_recv[par];
    for (int i = tid; i < h.n_shared; i += nt) {
        const int dof = h.shared[i];
        y[dof] = p2p_recv_entry(h, recv, flag, i, __ldcg(y + dof), &t.d.local->error);
    }
    __syncthreads();
    if (tid == 0) {
        if (h.n_nbr > 0) t.d.local->ll_seq[0] = seq;
        p2p_flag_error(t.d, t.st);
    }
}

// ---- mode 2: halo exchange OVERLAPPED with the interior elements -------------------------------------------------
// Elements holding shared dofs are the first `n_if_batches` batches of the element order (lpf_space_create puts them
// first; lpf_create measures how far they reach).  Persistent CTAs take batches in increasing order, so the interface
// is finished within the first wave:
//   p2p_if_batch_done()  after a CTA finished an interface batch: fence + count;
//   p2p_if_try_send()    per thread, once the count is complete: send its share of the interface as LL lines -- the
//                        NVLink latency now elapses while the CTAs work through the interior batches;
//   p2p_if_finish()      after the CTA's last batch: (send, if not done yet,) poll the neighbours' lines and add the
//                        partial sums in rank order; the CTA that finishes last all-reduces (d, A d) and closes the exchange.
// Interior elements never touch a shared dof, so y[shared] is final once the interface count is complete.
struct P2POverlap {
    unsigned long long seq;
    uint32_t flag;
    int par;
    bool sent;
};

__device__ __forceinline__ void p2p_if_begin(const P2PTail &t, P2POverlap &o)
{
    o.seq = t.d.local->ll_seq[0] + 1;       // written only by the closing CTA of the previous exchange (a previous kernel)
    o.par = (int)(o.seq & 1);
    o.flag = ll_flag(o.seq);
    o.sent = false;
}
__device__ __forceinline__ void p2p_if_batch_done(const P2PTail &t)      // call by all threads after an interface batch
{
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(&t.d.local->if_done, 1u);
}
__device__ __forceinline__ void p2p_if_send_share(const P2PTail &t, P2POverlap &o, const double *__restrict__ y)
{
    __threadfence();
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < t.h.n_shared; i += stride) p2p_send_entry(t.h, o.par, o.flag, i, __ldcg(y + t.h.shared[i]));
    o.sent = true;
}
__device__ __forceinline__ void p2p_if_try_send(const P2PTail &t, P2POverlap &o, const double *__restrict__ y)
{
    if (o.sent) return;
    if (*(volatile unsigned int *)&t.d.local->if_done >= (unsigned int)t.n_if_batches) p2p_if_send_share(t, o, y);
}
__device__ __forceinline__ void p2p_if_finish(const P2PTail &t, P2POverlap &o, double *__restrict__ y, volatile int *sflag)
{
    if (!o.sent) {                         // few batches per CTA: the interface may still be in flight on other CTAs
        long long n = 0;
        while (*(volatile unsigned int *)&t.d.local->if_done < (unsigned int)t.n_if_batches) {
            if (++n > LPF_P2P_SPIN_LIMIT) { t.d.local->error = 1; break; }
        }
        p2p_if_send_share(t, o, y);
    }
    const uint4 *recv = t.h.ll_recv[o.par];
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < t.h.n_shared; i += stride) {
        const int dof = t.h.shared[i];
        y[dof] = p2p_recv_entry(t.h, recv, o.flag, i, __ldcg(y + dof), &t.d.local->error);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int n = atomicInc(t.done, gridDim.x - 1);
        *sflag = (n == gridDim.x - 1);
    }
    __syncthreads();
    if (!*sflag) return;
    __threadfence();
    if (t.with_den && threadIdx.x < 32) {
        double s = 0.0;
        for (int i = threadIdx.x; i < LPF_DEN_SLOTS; i += 32) { s += __ldcg(t.den_slots + i); t.den_slots[i] = 0.0; }
#pragma unroll
        for (int w = 16; w > 0; w >>= 1) s += __shfl_xor_sync(0xffffffffu, s, w);
        const double tot = p2p_allreduce_warp(t.d, s);
        if (threadIdx.x == 0) t.st->red[1] = tot;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        t.d.local->if_done = 0;
        if (t.h.n_nbr > 0) t.d.local->ll_seq[0] = o.seq;
        p2p_flag_error(t.d, t.st);
    }
}
