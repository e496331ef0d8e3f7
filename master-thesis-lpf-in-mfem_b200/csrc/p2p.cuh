// Peer-memory (NVLink / NVSwitch) exchange for the multi-GPU PCG: halo-sum and scalar all-reduce done by our
// own kernels with plain 16-byte stores into the neighbours' mailboxes, instead of NCCL calls.  At the sizes the
// reference runs (10^5 - 10^6 dofs per rank) one CG iteration is latency-bound: two ncclAllReduce + one
// ncclSend/ncclRecv group cost ~20 us per iteration on top of ~17 us of compute (profiles/r01_scaling_summary.md); an
// LL line costs one one-way NVLink latency.  Replaces GroupCommunicator::Reduce/Bcast and the MPI_Allreduce inside
// CGSolver::Dot of the reference's MFEM path.
//
// Mailbox (one cudaMalloc per rank, exported to the peers by CUDA IPC or, for one-process-many-threads
// drivers, as a raw pointer with peer access enabled):
//     P2PBox header (LL lines of the all-reduce) | LL halo recv [2 parities][total] | LL surface halo recv [2][s_total]
// Writers own the slots indexed by THEIR rank, so no two ranks ever write the same line.  Every exchange k
// uses parity k & 1; a rank can only reach exchange k+2 after its neighbour finished reading exchange k
// (each exchange is a pairwise rendezvous), so two buffers are enough and lines never need resetting:
// they carry the monotonically increasing sequence number.
// Spin loops are bounded: on time-out they raise P2PLocal::error instead of hanging the GPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "vec_kernels.cuh"

#define LPF_P2P_MAXR 16
#define LPF_P2P_SPIN_LIMIT (1ll << 27)

struct __align__(16) P2PBox {
    uint4 red_ll[2][LPF_P2P_MAXR];                      // [parity][source rank] LL lines of the scalar all-reduce
    long long ll_byte_off[2][2];                        // [plan][parity] byte offset of the LL halo receive areas
    int off_for_src[2][LPF_P2P_MAXR];                   // [plan][source rank]: where that rank writes in MY recv buffer
};

struct P2PLocal {
    unsigned long long red_seq;
    int error;
    unsigned long long ll_seq[2];      // LL halo exchanges done, per plan
    unsigned int ll_counter[2];
};

struct P2PPlanDev {                    // device-side view of one halo plan
    int n_nbr, total, n_shared;
    const int *nbr_rank;               // [n_nbr]
    const int *nbr_offset;             // [n_nbr+1]
    const int *send_dofs;              // [total]
    const int *send_nbr;               // [total] neighbour index of each send entry
    const int *shared, *red_off, *red_src;
    // LL protocol (p2p_halo_ll_kernel): per shared dof i the (neighbour, position) pairs it is sent to
    const int *snd_off, *snd_nbr, *snd_pos;    // [n_shared+1], [total], [total]
    uint4 *const *ll_dst;              // [2 parities][n_nbr] where I write LL lines in each neighbour's box
    const uint4 *ll_recv[2];           // my LL receive areas
};

struct P2PDev {
    int nranks, rank;
    P2PBox *const *peers;              // [nranks] device pointers to every rank's box (incl. mine)
    P2PBox *mine;
    P2PLocal *local;
};

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

// ---- scalar all-reduce (sum) + what the PCG does with the result -----------------------------------------
enum { P2P_RED_PLAIN = 0, P2P_RED_NOM = 1, P2P_RED_BETA = 2, P2P_RED_DEN = 3 };

// launch <<<1, 32>>>; lane r sends this rank's value to rank r and waits for rank r's value
__global__ void p2p_allreduce_kernel(P2PDev d, double *val, int mode, PcgState *st, double *den_slots)
{
    __shared__ double local_val;
    const int lane = threadIdx.x;
    if (mode == P2P_RED_DEN) {          // (d, A d): sum this rank's slots first (and clear them for the next apply)
        double s = 0.0;
        for (int i = lane; i < LPF_DEN_SLOTS; i += 32) { s += den_slots[i]; den_slots[i] = 0.0; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) local_val = s;
    } else if (lane == 0) local_val = *val;
    __syncwarp();
    const double s = p2p_allreduce_warp(d, local_val);
    if (lane == 0) {
        if (mode == P2P_RED_NOM) pcg_finalize_nom(st, s);
        else if (mode == P2P_RED_BETA) { if (st->status == PCG_RUNNING) pcg_finalize_beta(st, s); }
        else if (mode == P2P_RED_DEN) st->red[1] = s;
        else *val = s;
    }
}

// ---- halo-sum over LL lines + the (d, A d) all-reduce, one multi-CTA kernel, no intra-rank synchronisation ---------
// One thread per shared dof: phase 1 sends this rank's partial sum to every sharer (plain 16-byte stores into their
// boxes), phase 2 polls the lines the sharers wrote and adds the partials in rank order (every sharer ends with
// bit-identical copies).  All sends of a thread precede its first poll, so two ranks can never wait on each other.
// Warp 0 of block 0 all-reduces the PCG denominator in between: (d, A d) = sum over ranks of the element-local
// products x_e . y_e, which does not need the halo-summed vector -- both exchanges share one NVLink latency.
__global__ void p2p_halo_ll_kernel(P2PDev d, P2PPlanDev h, int plan, double *__restrict__ v, PcgState *st,
                                   double *den_slots, int with_den)
{
    griddep_wait();
    griddep_launch();
    if (st != nullptr && st->status != PCG_RUNNING) return;      // same decision on every rank (status is global)
    const unsigned long long seq = d.local->ll_seq[plan] + 1;
    const int par = (int)(seq & 1);
    const uint32_t flag = ll_flag(seq);
    const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = t0; i < h.n_shared; i += stride) {
        const double own = v[h.shared[i]];
        for (int t = h.snd_off[i]; t < h.snd_off[i + 1]; t++) ll_store(h.ll_dst[par * h.n_nbr + h.snd_nbr[t]] + h.snd_pos[t], own, flag);
    }
    if (with_den && blockIdx.x == 0 && threadIdx.x < 32) {
        double s = 0.0;
        for (int i = threadIdx.x; i < LPF_DEN_SLOTS; i += 32) { s += den_slots[i]; den_slots[i] = 0.0; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const double tot = p2p_allreduce_warp(d, s);
        if (threadIdx.x == 0) st->red[1] = tot;
    }
    const uint4 *recv = h.ll_recv[par];
    for (int i = t0; i < h.n_shared; i += stride) {
        const int dof = h.shared[i];
        const double own = v[dof];
        double s = 0.0;
        for (int j = h.red_off[i]; j < h.red_off[i + 1]; j++) s += (h.red_src[j] < 0) ? own : ll_poll(recv + h.red_src[j], flag, &d.local->error);
        v[dof] = s;
    }
    __shared__ bool is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicInc(&d.local->ll_counter[plan], gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) d.local->ll_seq[plan] = seq;
}

// ---- halo-sum + (d, A d) all-reduce fused into the tail of the apply kernel ---------------------------------
// Every CTA of the persistent apply kernel calls p2p_apply_tail() after its last batch.  The CTA that arrives
// last (all scatter-adds of this rank are then globally visible) packs the interface values straight into the
// neighbours' mailboxes, raises their flags, waits for theirs, adds the partial sums in rank order and finally
// all-reduces the PCG denominator -- the collective rides on the compute kernel, no extra launch, no NCCL.
struct P2PTail {
    int enabled, with_den;
    P2PDev d;
    P2PPlanDev h;
    PcgState *st;
    double *den_slots;
    unsigned int *done;
};

__device__ __forceinline__ void p2p_apply_tail(const P2PTail &t, double *__restrict__ y)
{
    __shared__ bool tail_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int n = atomicInc(t.done, gridDim.x - 1);
        tail_last = (n == gridDim.x - 1);
    }
    __syncthreads();
    if (!tail_last) return;
    __threadfence();
    const int tid = threadIdx.x, nt = blockDim.x;
    const P2PPlanDev &h = t.h;
    // same LL lines and sequence numbers as p2p_halo_ll_kernel, so a rank with a small interface (tail) and a
    // neighbour with a large one (separate kernel) interoperate
    const unsigned long long seq = t.d.local->ll_seq[0] + 1;
    const int par = (int)(seq & 1);
    const uint32_t flag = ll_flag(seq);
    for (int i = tid; i < h.n_shared; i += nt) {
        const double own = __ldcg(y + h.shared[i]);
        for (int k = h.snd_off[i]; k < h.snd_off[i + 1]; k++) ll_store(h.ll_dst[par * h.n_nbr + h.snd_nbr[k]] + h.snd_pos[k], own, flag);
    }
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
        const double own = __ldcg(y + dof);
        double s = 0.0;
        for (int j = h.red_off[i]; j < h.red_off[i + 1]; j++) s += (h.red_src[j] < 0) ? own : ll_poll(recv + h.red_src[j], flag, &t.d.local->error);
        y[dof] = s;
    }
    __syncthreads();
    if (tid == 0 && h.n_nbr > 0) t.d.local->ll_seq[0] = seq;
}

// ---- PCG vector kernels with the cross-rank reduction inside (last block = one flag round trip, no extra launch) ----
__global__ void pcg_init_p2p_kernel(int n, const double *__restrict__ b, const double *__restrict__ t,
                                    const double *__restrict__ dinv, const uint8_t *__restrict__ owned,
                                    double *__restrict__ r, double *__restrict__ z, double *__restrict__ d,
                                    double *__restrict__ ad, PcgState *st, double *partials, P2PDev pd)
{
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double ri = t ? b[i] - t[i] : b[i];
        const double zi = dinv[i] * ri;
        r[i] = ri; z[i] = zi; d[i] = zi; ad[i] = 0.0;
        if (owned[i]) acc = fma(zi, ri, acc);
    }
    __shared__ double loc;
    const bool last = grid_sum_finalize(acc, partials, &st->counter, [&](double s) { loc = s; });
    if (last) {
        __syncthreads();
        if (threadIdx.x < 32) {
            const double tot = p2p_allreduce_warp(pd, loc);
            if (threadIdx.x == 0) pcg_finalize_nom(st, tot);
        }
    }
}

__global__ void pcg_update_p2p_kernel(int n, double *__restrict__ x, double *__restrict__ r, double *__restrict__ z,
                                      const double *__restrict__ d, const double *__restrict__ ad,
                                      const double *__restrict__ dinv, const uint8_t *__restrict__ owned, PcgState *st,
                                      double *partials, P2PDev pd)
{
    griddep_wait();
    griddep_launch();      // after the wait: at most ONE successor kernel is resident ahead of time
    if (st->status != PCG_RUNNING) return;
    const double den = st->red[1];                  // all-reduced by the tail of the previous apply
    if (den == 0.0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) { st->den = den; st->status = PCG_BREAKDOWN; st->final_iter = st->iter > 1 ? st->iter : 0; }
        return;
    }
    const double alpha = st->nom / den;
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        __stcs(x + i, fma(alpha, d[i], __ldcs(x + i)));       // streaming accesses, see pcg_update_kernel
        const double ri = fma(-alpha, ad[i], __ldcs(r + i));
        const double zi = __ldcs(dinv + i) * ri;
        __stcs(r + i, ri); z[i] = zi;
        if (owned[i]) acc = fma(ri, zi, acc);
    }
    __shared__ double loc;
    const bool last = grid_sum_finalize(acc, partials, &st->counter, [&](double s) { loc = s; });
    if (last) {
        __syncthreads();
        if (threadIdx.x < 32) {
            const double tot = p2p_allreduce_warp(pd, loc);
            if (threadIdx.x == 0) { st->den = den; pcg_finalize_beta(st, tot); }
        }
    }
}
