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
    const unsigned long long seq = d.local->red_seq + 1;
    const int par = (int)(seq & 1);
    if (lane < d.nranks) {
        P2PBox *peer = d.peers[lane];
        peer->red_val[par][d.rank] = local_val;
        __threadfence_system();
        p2p_store_flag(&peer->red_flag[par][d.rank], seq);
        p2p_wait(&d.mine->red_flag[par][lane], seq, &d.local->error);
    }
    __syncwarp();
    if (lane == 0) {
        double s = 0.0;
        for (int r = 0; r < d.nranks; r++) s += ((volatile double *)d.mine->red_val[par])[r];   // rank order: identical on all ranks
        d.local->red_seq = seq;
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
