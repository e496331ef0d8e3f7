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
#include "p2p_dev.cuh"
#include "vec_kernels.cuh"

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
        p2p_flag_error(d, st);
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
    if (is_last && threadIdx.x == 0) { d.local->ll_seq[plan] = seq; p2p_flag_error(d, st); }
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
            if (threadIdx.x == 0) { pcg_finalize_nom(st, tot); p2p_flag_error(pd, st); }
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
            if (threadIdx.x == 0) { st->den = den; pcg_finalize_beta(st, tot); p2p_flag_error(pd, st); }
        }
    }
}
