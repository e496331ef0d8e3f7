// Peer-memory (NVLink / NVSwitch) exchange for the multi-GPU PCG: halo-sum and scalar all-reduce done by our
// own kernels with plain stores into the neighbours' mailboxes and sequence-number flags, instead of NCCL
// calls.  At the sizes the reference runs (10^5 - 10^6 dofs per rank) one CG iteration is latency-bound:
// two ncclAllReduce + one ncclSend/ncclRecv group cost ~50 us per iteration (profiles/r01_scaling_summary.md),
// a flag round trip over NVLink costs a few us.  Replaces GroupCommunicator::Reduce/Bcast and the
// MPI_Allreduce inside CGSolver::Dot of the reference's MFEM path.
//
// Mailbox (one cudaMalloc per rank, exported to the peers by CUDA IPC or, for one-process-many-threads
// drivers, as a raw pointer with peer access enabled):
//     P2PBox header | halo recv [2 parities][total] | surface halo recv [2][s_total]
// Writers own the slots indexed by THEIR rank, so no two ranks ever write the same word.  Every exchange k
// uses parity k & 1; a rank can only reach exchange k+2 after its neighbour finished reading exchange k
// (each exchange is a pairwise rendezvous), so two buffers are enough and flags never need resetting:
// they carry the monotonically increasing sequence number.
// Spin loops are bounded: on time-out they raise P2PLocal::error instead of hanging the GPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "vec_kernels.cuh"

#define LPF_P2P_MAXR 16
#define LPF_P2P_SPIN_LIMIT (1ll << 27)

struct P2PBox {
    unsigned long long red_flag[2][LPF_P2P_MAXR];
    double red_val[2][LPF_P2P_MAXR];
    unsigned long long halo_flag[2][2][LPF_P2P_MAXR];   // [plan][parity][source rank]
    int off_for_src[2][LPF_P2P_MAXR];                   // [plan][source rank]: where that rank writes in MY recv buffer
    long long recv_byte_off[2][2];                      // [plan][parity] byte offset of the recv buffers in this box
};

struct P2PLocal {
    unsigned long long red_seq;
    unsigned long long halo_seq[2];
    unsigned int counter[2][2];        // [plan][pack/unpack]
    int error;
};

struct P2PPlanDev {                    // device-side view of one halo plan
    int n_nbr, total, n_shared;
    const int *nbr_rank;               // [n_nbr]
    const int *nbr_offset;             // [n_nbr+1]
    const int *send_dofs;              // [total]
    const int *send_nbr;               // [total] neighbour index of each send entry
    double *const *dst;                // [2 parities][n_nbr] where I write in each neighbour's recv buffer
    const int *shared, *red_off, *red_src;
    const double *recv[2];             // my recv buffers
};

struct P2PDev {
    int nranks, rank;
    P2PBox *const *peers;              // [nranks] device pointers to every rank's box (incl. mine)
    P2PBox *mine;
    P2PLocal *local;
};

__device__ __forceinline__ void p2p_store_flag(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long p2p_load_flag(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool p2p_wait(const unsigned long long *flag, unsigned long long seq, int *err)
{
    long long n = 0;
    while (p2p_load_flag(flag) < seq) {
        if (++n > LPF_P2P_SPIN_LIMIT) { *err = 1; return false; }
    }
    return true;
}

// Sum of one double over all ranks, executed by one full warp; lane r sends this rank's value to rank r and
// waits for rank r's value.  Returns the sum (identical bits on every rank: fixed rank order) in every lane.
__device__ __forceinline__ double p2p_allreduce_warp(const P2PDev &d, double local_val)
{
    const int lane = threadIdx.x & 31;
    const unsigned long long seq = d.local->red_seq + 1;
    const int par = (int)(seq & 1);
    __syncwarp();
    if (lane < d.nranks) {
        P2PBox *peer = d.peers[lane];
        peer->red_val[par][d.rank] = local_val;
        __threadfence_system();
        p2p_store_flag(&peer->red_flag[par][d.rank], seq);
        p2p_wait(&d.mine->red_flag[par][lane], seq, &d.local->error);
    }
    __syncwarp();
    double s = 0.0;
    for (int r = 0; r < d.nranks; r++) s += ((volatile double *)d.mine->red_val[par])[r];
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

// ---- halo-sum --------------------------------------------------------------------------------------------
__global__ void p2p_pack_kernel(P2PDev d, P2PPlanDev h, int plan, const double *__restrict__ v)
{
    const unsigned long long seq = d.local->halo_seq[plan] + 1;
    const int par = (int)(seq & 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < h.total; i += gridDim.x * blockDim.x) {
        const int k = h.send_nbr[i];
        h.dst[par * h.n_nbr + k][i - h.nbr_offset[k]] = v[h.send_dofs[i]];
    }
    __threadfence_system();
    __shared__ bool is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicInc(&d.local->counter[plan][0], gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && (int)threadIdx.x < h.n_nbr) {
        __threadfence_system();
        p2p_store_flag(&d.peers[h.nbr_rank[threadIdx.x]]->halo_flag[plan][par][d.rank], seq);
    }
}

__global__ void p2p_unpack_kernel(P2PDev d, P2PPlanDev h, int plan, double *__restrict__ v)
{
    const unsigned long long seq = d.local->halo_seq[plan] + 1;
    const int par = (int)(seq & 1);
    if ((int)threadIdx.x < h.n_nbr) p2p_wait(&d.mine->halo_flag[plan][par][h.nbr_rank[threadIdx.x]], seq, &d.local->error);
    __syncthreads();
    const volatile double *recv = h.recv[par];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < h.n_shared; i += gridDim.x * blockDim.x) {
        const int dof = h.shared[i];
        const double own = v[dof];
        double s = 0.0;
        for (int j = h.red_off[i]; j < h.red_off[i + 1]; j++) s += (h.red_src[j] < 0) ? own : recv[h.red_src[j]];
        v[dof] = s;
    }
    __shared__ bool is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicInc(&d.local->counter[plan][1], gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) d.local->halo_seq[plan] = seq;
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
    if (h.n_nbr > 0) {
        const unsigned long long seq = t.d.local->halo_seq[0] + 1;
        const int par = (int)(seq & 1);
        for (int i = tid; i < h.total; i += nt) {
            const int k = h.send_nbr[i];
            h.dst[par * h.n_nbr + k][i - h.nbr_offset[k]] = __ldcg(y + h.send_dofs[i]);
        }
        __threadfence_system();
        __syncthreads();
        if (tid < h.n_nbr) {
            p2p_store_flag(&t.d.peers[h.nbr_rank[tid]]->halo_flag[0][par][t.d.rank], seq);
            p2p_wait(&t.d.mine->halo_flag[0][par][h.nbr_rank[tid]], seq, &t.d.local->error);
        }
        __syncthreads();
        const volatile double *recv = h.recv[par];
        for (int i = tid; i < h.n_shared; i += nt) {
            const int dof = h.shared[i];
            const double own = __ldcg(y + dof);
            double s = 0.0;
            for (int j = h.red_off[i]; j < h.red_off[i + 1]; j++) s += (h.red_src[j] < 0) ? own : recv[h.red_src[j]];
            y[dof] = s;
        }
        __syncthreads();
        if (tid == 0) t.d.local->halo_seq[0] = seq;
    }
    if (t.with_den && tid < 32) {
        double s = 0.0;
        for (int i = tid; i < LPF_DEN_SLOTS; i += 32) { s += __ldcg(t.den_slots + i); t.den_slots[i] = 0.0; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const double tot = p2p_allreduce_warp(t.d, s);
        if (tid == 0) t.st->red[1] = tot;
    }
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
        x[i] = fma(alpha, d[i], x[i]);
        const double ri = fma(-alpha, ad[i], r[i]);
        const double zi = dinv[i] * ri;
        r[i] = ri; z[i] = zi;
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
