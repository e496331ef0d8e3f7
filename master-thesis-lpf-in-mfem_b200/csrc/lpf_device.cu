// Device context and C-ABI of the B200 hot path (include/lpf_b200.h, "Device context" section).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/lpf_b200.h"
#include "../host/lpf_common.hpp"
#include "../host/lpf_host.hpp"
#include "apply_api.hpp"
#include "dev_util.cuh"
#include "lpf_comm.hpp"
#include "pa_kernels.cuh"
#include "vec_kernels.cuh"
#include "p2p.cuh"

namespace {

template <class T>
int upload(T *&dst, const T *src, size_t n, size_t *bytes)
{
    dst = nullptr;
    if (n == 0) return LPF_OK;
    CUDA_TRY(cudaMalloc((void **)&dst, n * sizeof(T)));
    if (src) CUDA_TRY(cudaMemcpy(dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    else CUDA_TRY(cudaMemset(dst, 0, n * sizeof(T)));
    if (bytes) *bytes += n * sizeof(T);
    return LPF_OK;
}

inline int vec_grid(int n, int sm_count)
{
    // >= 4 entries per thread: on small vectors the cost of a reducing kernel is its tail (one atomic per block on the
    // "last block" counter + the final sum over the block partials), so fewer, fatter blocks; large vectors fill 8 x SMs
    const int need = (n + 1023) / 1024;
    return std::max(1, std::min(need, std::min(sm_count * 8, LPF_MAX_PARTIALS)));
}

}  // namespace

struct lpf_ctx {
    int dev = 0, sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int p = 0, D = 0, Q = 0, ne = 0, ndof = 0, ness = 0, nsurf = 0;
    int nranks = 1, rank = 0;
    long n_true_global = 0;
    size_t bytes = 0;
    long launches = 0;
    bool setup_done = false, jacobi_done = false, rhs_done = false;
    // options
    int variant = 0;          // apply kernel variant (elements per CTA / prefetch), see apply_launch
    int use_graph = 1, chunk = 16, skip_zero_apply = 1;
    int pdl = 0;              // programmatic dependent launch between the kernels of a PCG iteration (measured slower in
                              // round 2 on every size and GPU count, profiles/r02_pcg_ab.txt: off by default)
    bool pdl_now = false;     // set while pcg_iteration() enqueues
    int max_ctas = 0;
    int verbose = 0;
    std::vector<std::pair<std::string, long>> env_opts;   // LPF_DEBUG_ENV builds only
    // deterministic mode (option "deterministic"): the restriction-transpose is an ordered gather of an E-vector instead
    // of red.global.add, the diagonal likewise -- bit-identical results (and CG iteration counts) from run to run
    int deterministic = 0;
    double *yE = nullptr;
    int *det_off = nullptr, *det_idx = nullptr;
    int ess_general = 0;      // lpf_pcg: search directions may be non-zero on essential dofs
    // geometry / maps
    double *corners = nullptr, *jac = nullptr, *jinv_z = nullptr, *qd = nullptr;
    double *qa = nullptr;     // affine fast path: [ne][6] element tensors
    int affine = 1;           // option: use the fast path when every element is affine
    bool affine_ok = false;   // decided by lpf_pa_setup
    int *gmap = nullptr, *gmap_c = nullptr, *ess = nullptr;
    uint8_t *essmask = nullptr, *owned = nullptr, *surf_owned = nullptr;
    int ntrue[2] = {0, 0};            // owned (true) dofs of this rank: volume, surface
    int *tdof[2] = {nullptr, nullptr}; // [ntrue] L-dof of true dof t (ascending); nullptr: T == L (serial)
    // solver vectors
    double *dinv = nullptr, *r = nullptr, *z = nullptr, *d = nullptr, *ad = nullptr, *X = nullptr, *Bv = nullptr, *tmp = nullptr;
    double *zdad = nullptr;   // backing store of z, d, ad
    size_t zdad_bytes = 0;
    int l2_persist = 0;       // option: persisting-L2 window over z, d, A d during the PCG (measured: < 1 % at 2.25 M dofs, off)
    double *den_slots = nullptr, *partials = nullptr;
    PcgState *st = nullptr;
    PcgState *st_host = nullptr;      // pinned, two slots (status polling runs one chunk behind)
    cudaEvent_t poll_ev[2] = {nullptr, nullptr};
    int *bad = nullptr;
    // surface
    int *surf2vol = nullptr, *surf_mult = nullptr, *sd_off = nullptr, *sd_elem = nullptr, *sd_node = nullptr;
    double *surf_xy = nullptr, *cgen = nullptr, *cabs = nullptr, *cabsy = nullptr, *env = nullptr, *wsum = nullptr;
    bool use_cabsy = false;
    double *rk_k = nullptr, *rk_y = nullptr, *rk_z = nullptr, *state_dev = nullptr;
    double *state_pinned = nullptr;
    RhsDev rhs{};
    double rel_tol = 1e-12, abs_tol = 0.0;
    int max_iter = 1000;
    lpf_pcg_info last_info[4]{};
    int n_last = 0;
    // halo (multi-GPU)
    lpf::Comm comm;
    lpf::HaloPlan halo, shalo;
    // peer-memory exchange (p2p.cuh); when connected it replaces every NCCL call of the solver
    bool p2p_on = false;
    P2PBox *box = nullptr;
    size_t box_bytes = 0;
    P2PLocal *p2p_local = nullptr;
    P2PBox **peers_dev = nullptr;
    std::vector<P2PBox *> peers;
    std::vector<char> peer_opened;
    P2PPlanDev p2p_plan[2]{};
    int *p2p_send_nbr[2] = {nullptr, nullptr};
    P2PDev p2p{};
    P2PTail tail{};           // set around PCG applies when the exchange rides on the apply kernel
    unsigned int *p2p_done = nullptr;
    int p2p_fuse = 2;         // option: how the halo-sum + (d, A d) all-reduce of a PCG apply run (p2p_types.hpp, P2PTail):
                              // 0 = separate multi-CTA LL kernel after the apply; 1 = in the LAST CTA of the apply kernel
                              // (measured slower: 32 vs 21 us per iteration at 2 x 150k dofs); 2 = inside the apply kernel,
                              // overlapped with the interior elements
    int p2p_fuse_max = 2048;  // mode 1 only: while the interface has at most this many entries (one CTA runs the tail)
    int n_if_elems = 0;       // elements [0, n_if_elems) cover every element that touches a dof shared with another rank
    // host staging for *_host entry points
    double *hx = nullptr, *hy = nullptr;
    // pipelined host entry point (lpf_apply_T_host): element chunks / dof ranges, copy streams, events
    int sub_e0 = 0, sub_ne = -1;               // element sub-range for the next apply launch (-1 = all)
    int host_pipeline = 1;                     // option
    std::vector<int> hp_elem_begin, hp_elem_end; // [K] element range of chunk k
    std::vector<int> hp_x_end;                 // [K] chunk k reads x[0, hp_x_end[k]): ONE H2D copy per chunk brings the new part of the prefix
    struct HpRun { int a, b, ea, eb; };        // dofs [a, b) of y, positions [ea, eb) of the (sorted) ess list inside them
    std::vector<std::vector<HpRun>> hp_final;  // [K + 1] runs of y that are final once chunk k is done (K: after the halo-sum): one D2H copy each
    // what the plan is built from (kept so that options hp_ranges / hp_chunks can rebuild it): which of 128 fine dof
    // ranges every element touches (2 x 64-bit mask per element), the fine ranges holding shared dofs, the sorted ess list
    int hp_R = 128, hp_K = 8, hp_fine_rs = 0;
    std::vector<uint64_t> hp_elem_mask;
    uint64_t hp_shared_mask[2] = {0, 0};
    std::vector<int> hp_ess_host;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    std::vector<cudaEvent_t> hp_ev_x, hp_ev_c;
    cudaEvent_t hp_ev_start = nullptr, hp_ev_done = nullptr;
    // CUDA graph of one PCG chunk
    cudaGraphExec_t pcg_graph = nullptr;
    int pcg_graph_chunk = 0, pcg_graph_general = 0, pcg_graph_launches = 0;
    // timing
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

// ------------------------------------------------------------------------------------------------
// apply kernel dispatch (kernels: one translation unit per order, apply_api.hpp)
// ------------------------------------------------------------------------------------------------
namespace {

// Persistent L-vector apply y (+)= G^T A_E G x on the elements [sub_e0, sub_e0 + sub_ne) (default: all).  With yE != nullptr
// (deterministic mode) the kernel writes the E-vector instead of scatter-adding and the caller assembles y.
int apply_launch(lpf_ctx *c, const int *gmap, const double *x, double *y, double *den, const int *status, double *yE = nullptr)
{
    const int D3 = c->D * c->D * c->D, DP3 = (D3 + 3) & ~3, QE = 6 * c->Q * c->Q * c->Q;
    // optional element sub-range: one element's data is contiguous, so a sub-range is a pointer offset
    const int e0 = c->sub_ne >= 0 ? c->sub_e0 : 0, ne = c->sub_ne >= 0 ? c->sub_ne : c->ne;
    LpfApplyArgs a;
    a.dev = c->dev; a.sm_count = c->sm_count; a.stream = c->stream; a.variant = c->variant; a.max_ctas = c->max_ctas;
    a.pdl = c->pdl_now; a.verbose = c->verbose != 0;
    a.affine = yE == nullptr && c->variant == 0 && c->affine && c->affine_ok;
    a.k.qd = a.affine ? c->qa + (size_t)e0 * 6 : c->qd + (size_t)e0 * QE;
    a.k.gmap = gmap + (size_t)e0 * DP3;
    a.k.x = x; a.k.y = y; a.k.yE = yE ? yE + (size_t)e0 * D3 : nullptr; a.k.ne = ne;
    a.k.den_slots = den; a.k.status = status; a.k.tail = c->tail;
    int rc = LPF_ERR_UNSUPPORTED;
#define LPF_CALL(P) rc = lpf_apply_L_p##P(a)
    LPF_ORDER_SWITCH(c->p, LPF_CALL, lpf::set_error("unsupported order"))
#undef LPF_CALL
    if (rc == LPF_OK && a.grid > 0) c->launches++;
    return rc;
}

// deterministic restriction-transpose of the E-vector c->yE into y (rows with skip != 0 stay 0)
int det_gather(lpf_ctx *c, const uint8_t *skip, double *y)
{
    det_gather_kernel<<<(c->ndof + 255) / 256, 256, 0, c->stream>>>(c->ndof, c->det_off, c->det_idx, c->yE, skip, y);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return LPF_OK;
}

// y = G^T A_E G x on this rank's elements: scatter-add into the zeroed y, or (deterministic) E-vector + ordered gather.
// `constrained`: essential rows / columns masked (gmap_c).  zero_y: y is not known to be zero yet.
int apply_full(lpf_ctx *c, bool constrained, const double *x, double *y, double *den, const int *status, bool zero_y)
{
    const int *gm = constrained ? c->gmap_c : c->gmap;
    if (c->deterministic) {
        LPF_TRY(apply_launch(c, gm, x, y, den, status, c->yE));
        return det_gather(c, constrained ? c->essmask : nullptr, y);
    }
    if (zero_y) CUDA_TRY(cudaMemsetAsync(y, 0, sizeof(double) * c->ndof, c->stream));
    return apply_launch(c, gm, x, y, den, status);
}

// Plan of the pipelined host entry point (lpf_apply_T_host).  Dependencies are tracked on hp_R dof ranges (128 fine ranges by
// default), but copies are issued per element chunk: the interior elements [n_if, ne) go first in hp_K chunks (their dofs ascend
// with the element index, so chunk k needs a prefix x[0, hp_x_end[k]) and ONE H2D copy brings what chunk k - 1 did not need
// yet), then the interface block [0, n_if) -- it reads both ends of the slab and its rows are final only after the halo-sum.
// The ranges of y whose last-touching chunk is k leave as merged contiguous runs (one or two D2H copies per chunk; group K:
// after the halo-sum).  Every copy costs ~4 us of DMA time on top of its bytes (profiles/r02_e2e_host_ab.txt), so the number of
// copies follows the number of chunks, not the resolution of the dependency tracking.
void host_plan_build(lpf_ctx *c)
{
    c->hp_elem_begin.clear(); c->hp_elem_end.clear(); c->hp_x_end.clear(); c->hp_final.clear();
    if (c->hp_elem_mask.empty()) return;
    const int fine_per = 128 / c->hp_R;                       // fine ranges per tracked range
    const long rs = (long)c->hp_fine_rs * fine_per;
    std::vector<int> range_end;
    for (int j = 0; j < c->hp_R; j++) if (j * rs < c->ndof) range_end.push_back((int)std::min<long>((j + 1) * rs, c->ndof));
    const int nr = (int)range_end.size();
    const int nif = (c->n_if_elems < c->ne) ? c->n_if_elems : 0, KI = c->hp_K;
    // (uniform chunks: sizes ramping up and down geometrically -- short H2D head, short D2H tail -- measured no faster)
    for (int k = 0; k < KI; k++) {
        c->hp_elem_begin.push_back(nif + (int)((long)(c->ne - nif) * k / KI));
        c->hp_elem_end.push_back(nif + (int)((long)(c->ne - nif) * (k + 1) / KI));
    }
    if (nif > 0) { c->hp_elem_begin.push_back(0); c->hp_elem_end.push_back(nif); }
    const int K = (int)c->hp_elem_end.size();
    std::vector<int> last_chunk(nr, 0);
    int run_max = 0;
    auto each_fine = [&](const uint64_t m[2], auto &&fn) {
        for (int w = 0; w < 2; w++)
            for (uint64_t b = m[w]; b; b &= b - 1) fn(64 * w + __builtin_ctzll(b));
    };
    for (int k = 0; k < K; k++) {
        uint64_t m[2] = {0, 0};
        for (int e = c->hp_elem_begin[k]; e < c->hp_elem_end[k]; e++) { m[0] |= c->hp_elem_mask[(size_t)2 * e]; m[1] |= c->hp_elem_mask[(size_t)2 * e + 1]; }
        each_fine(m, [&](int r) { const int j = std::min(r / fine_per, nr - 1); run_max = std::max(run_max, j); last_chunk[j] = k; });
        c->hp_x_end.push_back(range_end[run_max]);
    }
    c->hp_x_end[K - 1] = c->ndof;          // whatever no element reads still has to arrive (essential rows copy x)
    // multi-GPU: ranges holding dofs shared with other ranks are final only after the halo-sum that follows the last chunk
    // (group K); with local dofs numbered by global id these are the few ranges at both ends of a slab
    each_fine(c->hp_shared_mask, [&](int r) { last_chunk[std::min(r / fine_per, nr - 1)] = K; });
    c->hp_final.assign(K + 1, {});
    const auto &ess = c->hp_ess_host;
    for (int j = 0; j < nr; j++) {
        auto &runs = c->hp_final[last_chunk[j]];
        const int a = j ? range_end[j - 1] : 0, b = range_end[j];
        if (!runs.empty() && runs.back().b == a) runs.back().b = b;
        else runs.push_back({a, b, 0, 0});
    }
    for (auto &runs : c->hp_final)
        for (auto &r : runs) {
            r.ea = (int)(std::lower_bound(ess.begin(), ess.end(), r.a) - ess.begin());
            r.eb = (int)(std::lower_bound(ess.begin(), ess.end(), r.b) - ess.begin());
        }
}

// How the halo-sum of an apply runs on this rank (P2PTail::mode): 0 = separate kernel(s) after the element kernel,
// 1 = in the last CTA of the element kernel, 2 = inside the element kernel, overlapped with the interior elements.
int tail_mode(const lpf_ctx *c)
{
    if (c->nranks == 1 || !c->p2p_on || c->ess_general || c->deterministic || c->variant != 0 || c->halo.n_nbr > 32) return 0;
    if (c->p2p_fuse == 1) return (c->halo.n_nbr > 0 && c->halo.total <= c->p2p_fuse_max) ? 1 : 0;
    return c->p2p_fuse == 2 ? 2 : 0;
}

// constrained apply with the halo-sum (and, inside the PCG, the all-reduce of (d, A d)) riding on the element kernel
int apply_with_tail(lpf_ctx *c, const double *x, double *y, bool pcg)
{
    c->tail.mode = tail_mode(c); c->tail.with_den = pcg ? 1 : 0; c->tail.n_if_batches = c->n_if_elems;
    c->tail.d = c->p2p; c->tail.h = c->p2p_plan[0];
    c->tail.st = pcg ? c->st : nullptr; c->tail.den_slots = c->den_slots; c->tail.done = c->p2p_done;
    const int rc = apply_launch(c, c->gmap_c, x, y, pcg ? c->den_slots : nullptr, pcg ? &c->st->status : nullptr);
    c->tail.mode = 0;
    return rc;
}

int halo_sum(lpf_ctx *c, lpf::HaloPlan &h, double *v)
{
    if (c->nranks == 1 || h.n_nbr == 0) return LPF_OK;
    if (c->p2p_on) {
        // one multi-CTA kernel: LL lines into the neighbours' boxes, poll, rank-ordered sum (p2p.cuh)
        const int plan = (&h == &c->halo) ? 0 : 1;
        const P2PPlanDev &pd = c->p2p_plan[plan];
        const int g = std::max(1, std::min((h.n_shared + 127) / 128, 2 * c->sm_count));
        PcgState *no_state = nullptr;
        double *no_slots = nullptr;
        CUDA_TRY(launch_ex(false, p2p_halo_ll_kernel, dim3(g), dim3(128), 0, c->stream, c->p2p, pd, plan, v, no_state, no_slots, 0));
        c->launches++;
        return LPF_OK;
    }
    if (!c->comm.ready()) { lpf::set_error("multi-rank context used before lpf_comm_init / lpf_p2p_connect"); return LPF_ERR_STATE; }
    const int ns = h.total;
    halo_pack_kernel<<<(ns + 255) / 256, 256, 0, c->stream>>>(ns, h.send_dofs, v, h.sendbuf);
    c->launches++;
    LPF_TRY(c->comm.exchange(h, c->stream));
    halo_unpack_kernel<<<(h.n_shared + 255) / 256, 256, 0, c->stream>>>(h.n_shared, h.shared, h.red_off, h.red_src, h.recvbuf, v);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return LPF_OK;
}

int upload_halo(lpf::HaloPlan &h, int n_nbr, const int *nbr_rank, const int *nbr_offset, const int *send, int n_shared,
                const int *shared, const int *red_off, const int *red_src, size_t *bytes)
{
    h.n_nbr = n_nbr;
    h.nbr_rank.assign(nbr_rank, nbr_rank + n_nbr);
    h.nbr_offset.assign(nbr_offset, nbr_offset + n_nbr + 1);
    h.total = n_nbr ? nbr_offset[n_nbr] : 0;
    h.n_shared = n_shared;
    if (h.total == 0) return LPF_OK;
    h.h_send.assign(send, send + h.total);
    h.h_shared.assign(shared, shared + n_shared);
    LPF_TRY(upload(h.send_dofs, send, (size_t)h.total, bytes));
    LPF_TRY(upload(h.shared, shared, (size_t)n_shared, bytes));
    LPF_TRY(upload(h.red_off, red_off, (size_t)n_shared + 1, bytes));
    LPF_TRY(upload(h.red_src, red_src, (size_t)red_off[n_shared], bytes));
    LPF_TRY(upload(h.sendbuf, (const double *)nullptr, (size_t)h.total, bytes));
    LPF_TRY(upload(h.recvbuf, (const double *)nullptr, (size_t)h.total, bytes));
    return LPF_OK;
}

// Mailbox of this rank: header + double-buffered receive areas of both halo plans (see p2p.cuh)
int p2p_create(lpf_ctx *c)
{
    const size_t hdr = (sizeof(P2PBox) + 255) & ~(size_t)255;
    const size_t n0 = (size_t)c->halo.total, n1 = (size_t)c->shalo.total;
    const size_t ll0 = hdr;                                                               // LL areas: 16-byte lines
    c->box_bytes = ll0 + 2 * (n0 + n1) * sizeof(uint4) + 256;
    CUDA_TRY(cudaMalloc((void **)&c->box, c->box_bytes));
    CUDA_TRY(cudaMemset(c->box, 0, c->box_bytes));
    P2PBox h;
    std::memset(&h, 0, sizeof(h));
    h.ll_byte_off[0][0] = (long long)ll0;
    h.ll_byte_off[0][1] = (long long)(ll0 + n0 * 16);
    h.ll_byte_off[1][0] = (long long)(ll0 + 2 * n0 * 16);
    h.ll_byte_off[1][1] = (long long)(ll0 + 2 * n0 * 16 + n1 * 16);
    for (int k = 0; k < c->halo.n_nbr; k++) h.off_for_src[0][c->halo.nbr_rank[k]] = c->halo.nbr_offset[k];
    for (int k = 0; k < c->shalo.n_nbr; k++) h.off_for_src[1][c->shalo.nbr_rank[k]] = c->shalo.nbr_offset[k];
    CUDA_TRY(cudaMemcpy(c->box, &h, sizeof(h), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc((void **)&c->p2p_local, sizeof(P2PLocal)));
    CUDA_TRY(cudaMemset(c->p2p_local, 0, sizeof(P2PLocal)));
    CUDA_TRY(cudaMalloc((void **)&c->p2p_done, sizeof(unsigned int)));
    CUDA_TRY(cudaMemset(c->p2p_done, 0, sizeof(unsigned int)));
    c->bytes += c->box_bytes;
    return LPF_OK;
}

int p2p_connect_impl(lpf_ctx *c, const void *handles, const uint64_t *raws, const int *devices, int same_process)
{
    const int R = c->nranks;
    c->peers.assign(R, nullptr);
    c->peer_opened.assign(R, 0);
    for (int r = 0; r < R; r++) {
        if (r == c->rank) { c->peers[r] = c->box; continue; }
        if (same_process) {
            if (!raws || !devices) { lpf::set_error("lpf_p2p_connect: same_process needs raw pointers and devices"); return LPF_ERR_ARG; }
            int can = 0;
            CUDA_TRY(cudaDeviceCanAccessPeer(&can, c->dev, devices[r]));
            if (!can) { lpf::set_error("lpf_p2p_connect: no peer access between the GPUs"); return LPF_ERR_COMM; }
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[r], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_TRY(e);
            cudaGetLastError();
            c->peers[r] = (P2PBox *)(uintptr_t)raws[r];
        } else {
            cudaIpcMemHandle_t mh;
            std::memcpy(&mh, (const char *)handles + (size_t)r * sizeof(mh), sizeof(mh));
            void *p = nullptr;
            CUDA_TRY(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
            c->peers[r] = (P2PBox *)p;
            c->peer_opened[r] = 1;
        }
    }
    CUDA_TRY(cudaMalloc((void **)&c->peers_dev, sizeof(P2PBox *) * R));
    CUDA_TRY(cudaMemcpy(c->peers_dev, c->peers.data(), sizeof(P2PBox *) * R, cudaMemcpyHostToDevice));
    lpf::HaloPlan *plans[2] = {&c->halo, &c->shalo};
    for (int pl = 0; pl < 2; pl++) {
        lpf::HaloPlan &h = *plans[pl];
        P2PPlanDev &pd = c->p2p_plan[pl];
        pd.n_nbr = h.n_nbr; pd.total = h.total; pd.n_shared = h.n_shared;
        if (h.n_nbr == 0) continue;
        std::vector<int> send_nbr(h.total);
        for (int k = 0; k < h.n_nbr; k++) for (int i = h.nbr_offset[k]; i < h.nbr_offset[k + 1]; i++) send_nbr[i] = k;
        std::vector<uint4 *> ll_dst(2 * (size_t)h.n_nbr);
        for (int k = 0; k < h.n_nbr; k++) {
            const int s = h.nbr_rank[k];
            P2PBox ph;      // the neighbour's header tells where this rank writes inside its receive area
            CUDA_TRY(cudaMemcpy(&ph, c->peers[s], sizeof(ph), cudaMemcpyDefault));
            for (int par = 0; par < 2; par++) {
                ll_dst[(size_t)par * h.n_nbr + k] = (uint4 *)((char *)c->peers[s] + ph.ll_byte_off[pl][par]) + ph.off_for_src[pl][c->rank];
            }
        }
        // LL send lists: for shared dof i the (neighbour, position in that neighbour's block) pairs it goes to
        std::vector<int> snd_off(h.n_shared + 1, 0), snd_nbr(h.total), snd_pos(h.total);
        {
            std::vector<int> where((size_t)std::max(1, std::max(c->ndof, c->nsurf)), -1);
            for (int i = 0; i < h.n_shared; i++) where[h.h_shared[i]] = i;
            for (int i = 0; i < h.total; i++) {
                const int w = where[h.h_send[i]];
                if (w < 0) { lpf::set_error("lpf_p2p_connect: a sent dof is missing from the shared list"); return LPF_ERR_ARG; }
                snd_off[w + 1]++;
            }
            for (int i = 0; i < h.n_shared; i++) snd_off[i + 1] += snd_off[i];
            std::vector<int> pos(snd_off.begin(), snd_off.end() - 1);
            for (int i = 0; i < h.total; i++) {
                const int w = where[h.h_send[i]], k = send_nbr[i];
                snd_nbr[pos[w]] = k; snd_pos[pos[w]] = i - h.nbr_offset[k]; pos[w]++;
            }
        }
        int *so = nullptr, *sn = nullptr, *spz = nullptr;
        uint4 **lld = nullptr;
        LPF_TRY(upload(so, snd_off.data(), snd_off.size(), &c->bytes));
        LPF_TRY(upload(sn, snd_nbr.data(), snd_nbr.size(), &c->bytes));
        LPF_TRY(upload(spz, snd_pos.data(), snd_pos.size(), &c->bytes));
        LPF_TRY(upload(lld, ll_dst.data(), ll_dst.size(), &c->bytes));
        pd.snd_off = so; pd.snd_nbr = sn; pd.snd_pos = spz; pd.ll_dst = lld;
        int *nr = nullptr, *no = nullptr;
        LPF_TRY(upload(c->p2p_send_nbr[pl], send_nbr.data(), send_nbr.size(), &c->bytes));
        LPF_TRY(upload(nr, h.nbr_rank.data(), h.nbr_rank.size(), &c->bytes));
        LPF_TRY(upload(no, h.nbr_offset.data(), h.nbr_offset.size(), &c->bytes));
        pd.nbr_rank = nr; pd.nbr_offset = no;          // small tables, freed with the context's device at exit
        pd.send_dofs = h.send_dofs; pd.send_nbr = c->p2p_send_nbr[pl];
        pd.shared = h.shared; pd.red_off = h.red_off; pd.red_src = h.red_src;
        P2PBox mine;
        CUDA_TRY(cudaMemcpy(&mine, c->box, sizeof(mine), cudaMemcpyDeviceToHost));
        for (int par = 0; par < 2; par++) {
            pd.ll_recv[par] = (const uint4 *)((char *)c->box + mine.ll_byte_off[pl][par]);
        }
    }
    c->p2p.nranks = R; c->p2p.rank = c->rank; c->p2p.peers = c->peers_dev; c->p2p.mine = c->box; c->p2p.local = c->p2p_local;
    c->p2p_on = true;
    if (c->pcg_graph) { cudaGraphExecDestroy(c->pcg_graph); c->pcg_graph = nullptr; }
    return LPF_OK;
}

int create_impl(lpf_ctx *c, const lpf_space_desc *d, int device, void *stream)
{
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { lpf::set_error("lpf_create: no such CUDA device"); return LPF_ERR_CUDA; }
    CUDA_TRY(cudaSetDevice(device));
    c->dev = device;
    CUDA_TRY(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    if (stream) c->stream = (cudaStream_t)stream;
    else { CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
#ifdef LPF_DEBUG_ENV      // developer builds only: environment overrides of the options (the product reads no environment)
    for (const char *name : {"pdl", "pcg_chunk", "p2p_fuse", "p2p_fuse_max", "affine", "apply_variant", "max_ctas", "l2_persist", "deterministic", "verbose"}) {
        std::string env = "LPF_";
        for (const char *q = name; *q; q++) env += (char)std::toupper(*q);
        if (const char *e = std::getenv(env.c_str())) c->env_opts.emplace_back(name, std::atol(e));
    }
#endif
    c->p = d->order; c->D = d->order + 1; c->Q = d->order + 2;
    c->ne = d->ne; c->ndof = d->ndof; c->ness = d->n_ess; c->nsurf = d->n_surf;
    c->nranks = d->nranks > 0 ? d->nranks : 1; c->rank = d->rank;
    c->ntrue[0] = d->ndof; c->ntrue[1] = d->n_surf;
    c->n_true_global = d->n_true_global ? d->n_true_global : d->ndof;
    const int D3 = c->D * c->D * c->D, Q3 = c->Q * c->Q * c->Q;

    // basis tables -> constant memory: slot p of the generic kernels' table, and the apply kernels' own per-order tables
    {
        lpf::Basis1D bs(c->p);
        LpfBasisTab t;
        std::memset(&t, 0, sizeof(t));
        std::copy(bs.B.begin(), bs.B.end(), t.B);
        std::copy(bs.G.begin(), bs.G.end(), t.G);
        std::copy(bs.Dhat.begin(), bs.Dhat.end(), t.Dhat);
        std::copy(bs.nodes.begin(), bs.nodes.end(), t.nodes);
        std::copy(bs.qpts.begin(), bs.qpts.end(), t.qpts);
        std::copy(bs.qwts.begin(), bs.qwts.end(), t.qwts);
        CUDA_TRY(cudaMemcpyToSymbol(c_tab, &t, sizeof(t), sizeof(LpfBasisTab) * c->p));
        int rc = LPF_ERR_UNSUPPORTED;
#define LPF_CALL(P) rc = lpf_apply_tables_p##P(bs.B.data(), bs.G.data(), bs.qwts.data())
        LPF_ORDER_SWITCH(c->p, LPF_CALL, lpf::set_error("unsupported order"))
#undef LPF_CALL
        LPF_TRY(rc);
    }
    if (d->corners) LPF_TRY(upload(c->corners, d->corners, (size_t)c->ne * 24, &c->bytes));
    if (d->jac) LPF_TRY(upload(c->jac, d->jac, (size_t)c->ne * Q3 * 9, &c->bytes));
    if (d->jinv_z) LPF_TRY(upload(c->jinv_z, d->jinv_z, (size_t)c->ne * D3 * 3, &c->bytes));
    const int DP3 = (D3 + 3) & ~3;       // rows padded to 16 bytes (bulk-copy granularity)
    {
        std::vector<int> gp((size_t)c->ne * DP3, 0);
        for (int e = 0; e < c->ne; e++)
            for (int k = 0; k < D3; k++) gp[(size_t)e * DP3 + k] = d->gather[(size_t)e * D3 + k];
        LPF_TRY(upload(c->gmap, gp.data(), gp.size(), &c->bytes));
    }
    if (c->nranks > 1 && d->n_shared > 0 && d->shared_dofs) {
        // how far the elements touching shared dofs reach in the element order (lpf_space_create puts them first): the
        // overlapped halo exchange sends the interface as soon as these are done
        std::vector<uint8_t> sh((size_t)c->ndof, 0);
        for (int i = 0; i < d->n_shared; i++) sh[d->shared_dofs[i]] = 1;
        for (int e = 0; e < c->ne; e++)
            for (int k = 0; k < D3; k++) if (sh[d->gather[(size_t)e * D3 + k]]) { c->n_if_elems = e + 1; break; }
    }
    {   // constrained map: essential dofs encoded as ~dof (gather reads 0, scatter skips)
        std::vector<uint8_t> em((size_t)c->ndof, 0);
        for (int i = 0; i < c->ness; i++) {
            if (d->ess[i] < 0 || d->ess[i] >= c->ndof) { lpf::set_error("lpf_create: essential dof out of range"); return LPF_ERR_ARG; }
            em[d->ess[i]] = 1;
        }
        std::vector<int> gc((size_t)c->ne * DP3, 0);
        for (int e = 0; e < c->ne; e++)
            for (int k = 0; k < D3; k++) {
                const int g = d->gather[(size_t)e * D3 + k];
                if (g < 0 || g >= c->ndof) { lpf::set_error("lpf_create: gather index out of range"); return LPF_ERR_ARG; }
                gc[(size_t)e * DP3 + k] = em[g] ? ~g : g;
            }
        LPF_TRY(upload(c->gmap_c, gc.data(), gc.size(), &c->bytes));
        // pipelined host entry point: remember which dof ranges every element touches, then build the plan (host_plan_build)
        const bool ess_sorted = std::is_sorted(d->ess, d->ess + c->ness);
        if (c->ndof >= (1 << 18) && c->ne >= 64 && ess_sorted) {
            c->hp_fine_rs = ((c->ndof + 127) / 128 + 511) & ~511;             // 128 fine ranges, 4 KB aligned
            c->hp_elem_mask.assign((size_t)2 * c->ne, 0);
            for (int e = 0; e < c->ne; e++)
                for (int q = 0; q < D3; q++) {
                    const int r = d->gather[(size_t)e * D3 + q] / c->hp_fine_rs;
                    c->hp_elem_mask[(size_t)2 * e + (r >> 6)] |= 1ull << (r & 63);
                }
            for (int i = 0; i < d->n_shared; i++) { const int r = d->shared_dofs[i] / c->hp_fine_rs; c->hp_shared_mask[r >> 6] |= 1ull << (r & 63); }
            c->hp_ess_host.assign(d->ess, d->ess + c->ness);
            // chunks by vector size (profiles/r02_e2e_host_ab.txt: 8 chunks are fastest at 139 MB, 4 at 18 MB; every additional
            // chunk costs ~13 us); dependencies tracked on all 128 fine ranges
            const size_t vbytes = sizeof(double) * (size_t)c->ndof;
            c->hp_R = 128;
            c->hp_K = vbytes >= ((size_t)64 << 20) ? 8 : 4;
            host_plan_build(c);
        }
        LPF_TRY(upload(c->essmask, em.data(), em.size(), &c->bytes));
        LPF_TRY(upload(c->ess, d->ess, (size_t)c->ness, &c->bytes));
    }
    LPF_TRY(upload(c->qd, (const double *)nullptr, (size_t)c->ne * 6 * Q3, &c->bytes));
    if (d->corners) LPF_TRY(upload(c->qa, (const double *)nullptr, (size_t)c->ne * 6, &c->bytes));
    const size_t n = (size_t)c->ndof;
    LPF_TRY(upload(c->dinv, (const double *)nullptr, n, &c->bytes));
    LPF_TRY(upload(c->r, (const double *)nullptr, n, &c->bytes));
    {   // z, d, A d in ONE allocation: the vectors the kernels of a CG iteration hand to each other; one L2 access-policy
        // window can then keep them resident in L2 across the iteration (pcg_l2_window)
        const size_t np = (n + 31) & ~(size_t)31;
        LPF_TRY(upload(c->zdad, (const double *)nullptr, 3 * np, &c->bytes));
        c->z = c->zdad; c->d = c->zdad + np; c->ad = c->zdad + 2 * np;
        c->zdad_bytes = 3 * np * sizeof(double);
    }
    LPF_TRY(upload(c->X, (const double *)nullptr, n, &c->bytes));
    LPF_TRY(upload(c->Bv, (const double *)nullptr, n, &c->bytes));
    LPF_TRY(upload(c->tmp, (const double *)nullptr, n, &c->bytes));
    LPF_TRY(upload(c->den_slots, (const double *)nullptr, LPF_DEN_SLOTS, &c->bytes));
    LPF_TRY(upload(c->partials, (const double *)nullptr, LPF_MAX_PARTIALS, &c->bytes));
    LPF_TRY(upload(c->st, (const PcgState *)nullptr, 1, &c->bytes));
    LPF_TRY(upload(c->bad, (const int *)nullptr, 1, &c->bytes));
    CUDA_TRY(cudaMallocHost((void **)&c->st_host, 2 * sizeof(PcgState)));
    if (c->nranks > 1) {
        if (!d->owned) { lpf::set_error("lpf_create: multi-rank descriptor without ownership mask"); return LPF_ERR_ARG; }
        LPF_TRY(upload(c->owned, d->owned, n, &c->bytes));
        {   // true-dof lists: T-vectors (MFEM's GetTrueVSize ordering: owned L-dofs in ascending order) <-> L-vectors
            std::vector<int> t;
            for (int i = 0; i < c->ndof; i++) if (d->owned[i]) t.push_back(i);
            c->ntrue[0] = (int)t.size();
            LPF_TRY(upload(c->tdof[0], t.data(), t.size(), &c->bytes));
            t.clear();
            if (d->surf_owned) for (int i = 0; i < c->nsurf; i++) if (d->surf_owned[i]) t.push_back(i);
            c->ntrue[1] = (int)t.size();
            LPF_TRY(upload(c->tdof[1], t.data(), t.size(), &c->bytes));
        }
        LPF_TRY(upload_halo(c->halo, d->n_nbr, d->nbr_rank, d->nbr_offset, d->send_dofs, d->n_shared, d->shared_dofs,
                            d->red_off, d->red_src, &c->bytes));
        LPF_TRY(upload_halo(c->shalo, d->s_n_nbr, d->s_nbr_rank, d->s_nbr_offset, d->s_send, d->s_n_shared, d->s_shared,
                            d->s_red_off, d->s_red_src, &c->bytes));
        if (c->nranks <= LPF_P2P_MAXR) LPF_TRY(p2p_create(c));
    }
    // surface tables
    if (c->nsurf > 0) {
        const size_t ns = (size_t)c->nsurf;
        LPF_TRY(upload(c->surf2vol, d->surf2vol, ns, &c->bytes));
        LPF_TRY(upload(c->surf_xy, d->surf_xy, 2 * ns, &c->bytes));
        std::vector<int> vol2surf((size_t)c->ndof, -1);
        for (int s = 0; s < c->nsurf; s++) vol2surf[d->surf2vol[s]] = s;
        // per surface dof: the (element, local node) pairs contributing to GetDerivative, element order
        std::vector<int> cnt(ns + 1, 0);
        std::vector<int> elems;
        if (d->surf_elems) elems.assign(d->surf_elems, d->surf_elems + d->n_surf_elems);
        else { elems.resize(c->ne); for (int e = 0; e < c->ne; e++) elems[e] = e; }
        for (int e : elems)
            for (int k = 0; k < D3; k++) { const int s = vol2surf[d->gather[(size_t)e * D3 + k]]; if (s >= 0) cnt[s + 1]++; }
        for (size_t s = 0; s < ns; s++) cnt[s + 1] += cnt[s];
        std::vector<int> se(cnt[ns]), sn(cnt[ns]), pos(cnt.begin(), cnt.end() - 1);
        for (int e : elems)
            for (int k = 0; k < D3; k++) {
                const int s = vol2surf[d->gather[(size_t)e * D3 + k]];
                if (s >= 0) { se[pos[s]] = e; sn[pos[s]] = k; pos[s]++; }
            }
        std::vector<int> mult(ns);
        for (size_t s = 0; s < ns; s++) mult[s] = d->surf_mult ? d->surf_mult[s] : cnt[s + 1] - cnt[s];
        LPF_TRY(upload(c->sd_off, cnt.data(), ns + 1, &c->bytes));
        LPF_TRY(upload(c->sd_elem, se.data(), se.size(), &c->bytes));
        LPF_TRY(upload(c->sd_node, sn.data(), sn.size(), &c->bytes));
        LPF_TRY(upload(c->surf_mult, mult.data(), ns, &c->bytes));
        LPF_TRY(upload(c->wsum, (const double *)nullptr, ns, &c->bytes));
        LPF_TRY(upload(c->cgen, (const double *)nullptr, ns, &c->bytes));
        LPF_TRY(upload(c->cabs, (const double *)nullptr, ns, &c->bytes));
        LPF_TRY(upload(c->cabsy, (const double *)nullptr, ns, &c->bytes));
        LPF_TRY(upload(c->env, (const double *)nullptr, ns, &c->bytes));
        LPF_TRY(upload(c->rk_k, (const double *)nullptr, 2 * ns, &c->bytes));
        LPF_TRY(upload(c->rk_y, (const double *)nullptr, 2 * ns, &c->bytes));
        LPF_TRY(upload(c->rk_z, (const double *)nullptr, 2 * ns, &c->bytes));
        LPF_TRY(upload(c->state_dev, (const double *)nullptr, 2 * ns, &c->bytes));
        CUDA_TRY(cudaMallocHost((void **)&c->state_pinned, 2 * ns * sizeof(double)));
        if (c->nranks > 1 && d->surf_owned) LPF_TRY(upload(c->surf_owned, d->surf_owned, ns, &c->bytes));
    }
    CUDA_TRY(cudaEventCreate(&c->ev0));
    CUDA_TRY(cudaEventCreate(&c->ev1));
    return LPF_OK;
}

// Deterministic mode: transposed element map (offsets / indices of ElementRestriction, indices ascending per dof) and
// the E-vector the apply kernel writes instead of scatter-adding.  Built on first use from the device copy of the map.
int det_setup(lpf_ctx *c)
{
    if (c->det_off) return LPF_OK;
    CUDA_TRY(cudaSetDevice(c->dev));
    const int D3 = c->D * c->D * c->D, DP3 = (D3 + 3) & ~3;
    std::vector<int> gm((size_t)c->ne * DP3);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (!gm.empty()) CUDA_TRY(cudaMemcpy(gm.data(), c->gmap, gm.size() * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<int> off((size_t)c->ndof + 1, 0), idx((size_t)c->ne * D3);
    for (int e = 0; e < c->ne; e++) for (int k = 0; k < D3; k++) off[gm[(size_t)e * DP3 + k] + 1]++;
    for (int i = 0; i < c->ndof; i++) off[i + 1] += off[i];
    std::vector<int> pos(off.begin(), off.end() - 1);
    for (int e = 0; e < c->ne; e++) for (int k = 0; k < D3; k++) idx[pos[gm[(size_t)e * DP3 + k]]++] = e * D3 + k;
    LPF_TRY(upload(c->det_off, off.data(), off.size(), &c->bytes));
    LPF_TRY(upload(c->det_idx, idx.data(), idx.size(), &c->bytes));
    LPF_TRY(upload(c->yE, (const double *)nullptr, (size_t)c->ne * D3, &c->bytes));
    return LPF_OK;
}

void free_halo(lpf::HaloPlan &h)
{
    cudaFree(h.send_dofs); cudaFree(h.shared); cudaFree(h.red_off); cudaFree(h.red_src); cudaFree(h.sendbuf); cudaFree(h.recvbuf);
}

}  // namespace

extern "C" {

lpf_ctx *lpf_create(const lpf_space_desc *desc, int device, void *stream)
{
    if (!desc || !desc->gather || (!desc->corners && !desc->jac) || desc->ne < 0 || desc->ndof <= 0) {
        lpf::set_error("lpf_create: incomplete descriptor");
        return nullptr;
    }
    if (desc->order < 1 || desc->order > LPF_MAXP) { lpf::set_error("lpf_create: order must be in 1..10"); return nullptr; }
    if (desc->n_ess > 0 && !desc->ess) { lpf::set_error("lpf_create: n_ess > 0 but ess == NULL"); return nullptr; }
    auto *c = new lpf_ctx;
    if (create_impl(c, desc, device, stream) != LPF_OK) { lpf_destroy(c); return nullptr; }
    for (auto &o : c->env_opts) if (lpf_set_option(c, o.first.c_str(), o.second) != LPF_OK) { lpf_destroy(c); return nullptr; }
    return c;
}

void lpf_destroy(lpf_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->dev);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->pcg_graph) cudaGraphExecDestroy(c->pcg_graph);
    if (c->jacobi_done && c->l2_persist) cudaCtxResetPersistingL2Cache();
    c->comm.destroy();
    void *ptrs[] = {c->qa, c->corners, c->jac, c->jinv_z, c->qd, c->gmap, c->gmap_c, c->ess, c->essmask, c->owned, c->surf_owned, c->dinv, c->r,
                    c->zdad, c->X, c->Bv, c->tmp, c->den_slots, c->partials, c->st, c->bad, c->surf2vol,
                    c->yE, c->det_off, c->det_idx, c->surf_mult, c->sd_off, c->sd_elem, c->sd_node, c->surf_xy, c->cgen, c->cabs, c->cabsy, c->env, c->wsum, c->rk_k,
                    c->rk_y, c->rk_z, c->state_dev};
    for (void *p : ptrs) if (p) cudaFree(p);
    free_halo(c->halo); free_halo(c->shalo);
    for (size_t r = 0; r < c->peers.size(); r++) if (c->peer_opened[r]) cudaIpcCloseMemHandle(c->peers[r]);
    void *pp[] = {c->box, c->p2p_local, c->p2p_done, c->peers_dev, c->p2p_send_nbr[0], c->p2p_send_nbr[1], c->tdof[0], c->tdof[1]};
    for (void *p : pp) if (p) cudaFree(p);
    if (c->st_host) cudaFreeHost(c->st_host);
    if (c->state_pinned) cudaFreeHost(c->state_pinned);
    if (c->hx) cudaFreeHost(c->hx);
    if (c->hy) cudaFreeHost(c->hy);
    for (auto e : c->hp_ev_x) cudaEventDestroy(e);
    for (auto e : c->hp_ev_c) cudaEventDestroy(e);
    if (c->hp_ev_start) cudaEventDestroy(c->hp_ev_start);
    if (c->hp_ev_done) cudaEventDestroy(c->hp_ev_done);
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
    for (auto e : c->poll_ev) if (e) cudaEventDestroy(e);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

void *lpf_stream(lpf_ctx *c) { return c ? (void *)c->stream : nullptr; }
int lpf_sync(lpf_ctx *c)
{
    if (!c) { lpf::set_error("null context"); return LPF_ERR_ARG; }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return LPF_OK;
}
int lpf_ndof(const lpf_ctx *c) { return c ? c->ndof : LPF_ERR_ARG; }
int lpf_nsurf(const lpf_ctx *c) { return c ? c->nsurf : LPF_ERR_ARG; }
int lpf_ntrue(const lpf_ctx *c, int which) { return (c && (which == 0 || which == 1)) ? c->ntrue[which] : LPF_ERR_ARG; }

// P: true dofs -> L-vector (the owner's value on every sharer).  Non-owned entries start at 0, the halo-sum then carries
// the owner's value to them -- ConformingProlongationOperator::Mult + GroupCommunicator::Bcast of the reference's MFEM path.
int lpf_prolong(lpf_ctx *c, int which, const double *xT, double *xL)
{
    if (!c || !xT || !xL || (which != 0 && which != 1)) { lpf::set_error("lpf_prolong: bad argument"); return LPF_ERR_ARG; }
    const int nl = which == 0 ? c->ndof : c->nsurf, nt = c->ntrue[which];
    if (nl == 0) return LPF_OK;
    if (!c->tdof[which]) { CUDA_TRY(cudaMemcpyAsync(xL, xT, sizeof(double) * nl, cudaMemcpyDeviceToDevice, c->stream)); return LPF_OK; }
    CUDA_TRY(cudaMemsetAsync(xL, 0, sizeof(double) * nl, c->stream));
    if (nt) {
        scatter_kernel<<<(nt + 255) / 256, 256, 0, c->stream>>>(nt, c->tdof[which], xT, xL);
        c->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    return halo_sum(c, which == 0 ? c->halo : c->shalo, xL);
}

// R: L-vector -> true dofs (the owned entries)
int lpf_restrict(lpf_ctx *c, int which, const double *xL, double *xT)
{
    if (!c || !xT || !xL || (which != 0 && which != 1)) { lpf::set_error("lpf_restrict: bad argument"); return LPF_ERR_ARG; }
    const int nl = which == 0 ? c->ndof : c->nsurf, nt = c->ntrue[which];
    if (!c->tdof[which]) { if (nl) CUDA_TRY(cudaMemcpyAsync(xT, xL, sizeof(double) * nl, cudaMemcpyDeviceToDevice, c->stream)); return LPF_OK; }
    if (nt) {
        halo_pack_kernel<<<(nt + 255) / 256, 256, 0, c->stream>>>(nt, c->tdof[which], xL, xT);
        c->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    return LPF_OK;
}

long lpf_launch_count(lpf_ctx *c) { return c ? c->launches : 0; }
size_t lpf_device_bytes(const lpf_ctx *c) { return c ? c->bytes : 0; }
int lpf_affine_active(const lpf_ctx *c) { return c ? (int)(c->affine && c->affine_ok) : 0; }

int lpf_set_option(lpf_ctx *c, const char *name, long value)
{
    if (!c || !name) { lpf::set_error("lpf_set_option: null argument"); return LPF_ERR_ARG; }
    const std::string k(name);
    if (k == "apply_variant") c->variant = (int)value;
    else if (k == "use_graph") c->use_graph = (int)value;
    else if (k == "pcg_chunk") { if (value < 1) { lpf::set_error("pcg_chunk must be >= 1"); return LPF_ERR_ARG; } c->chunk = (int)value; }
    else if (k == "skip_zero_apply") c->skip_zero_apply = (int)value;
    else if (k == "p2p_fuse") c->p2p_fuse = (int)value;
    else if (k == "pdl") c->pdl = (int)value;
    else if (k == "affine") c->affine = (int)value;
    else if (k == "host_pipeline") c->host_pipeline = (int)value;
    else if (k == "hp_ranges" || k == "hp_chunks") {      // granularity of the pipelined host entry point
        const int v = (int)value;
        if (k == "hp_ranges") { if (v != 16 && v != 32 && v != 64 && v != 128) { lpf::set_error("hp_ranges must be 16, 32, 64 or 128"); return LPF_ERR_ARG; } c->hp_R = v; }
        else { if (v < 1 || v > 256) { lpf::set_error("hp_chunks must be in [1, 256]"); return LPF_ERR_ARG; } c->hp_K = v; }
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        for (auto e : c->hp_ev_x) cudaEventDestroy(e);
        for (auto e : c->hp_ev_c) cudaEventDestroy(e);
        c->hp_ev_x.clear(); c->hp_ev_c.clear();
        host_plan_build(c);
    }
    else if (k == "l2_persist") c->l2_persist = (int)value;
    else if (k == "verbose") c->verbose = (int)value;
    else if (k == "deterministic") { if (value) LPF_TRY(det_setup(c)); c->deterministic = value ? 1 : 0; }
    else if (k == "p2p_fuse_max") c->p2p_fuse_max = (int)value;
    else if (k == "max_ctas") c->max_ctas = (int)value;      // persistent kernels: cap the grid (tests force many batches per CTA)
    else { lpf::set_error("lpf_set_option: unknown option " + k); return LPF_ERR_ARG; }
    if (c->pcg_graph) { cudaGraphExecDestroy(c->pcg_graph); c->pcg_graph = nullptr; }
    return LPF_OK;
}

int lpf_comm_unique_id(void *id128) { return lpf::Comm::unique_id(id128); }
int lpf_comm_init(lpf_ctx *c, const void *id128)
{
    if (!c || !id128) { lpf::set_error("lpf_comm_init: null argument"); return LPF_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(c->dev));
    return c->comm.init(id128, c->nranks, c->rank);
}

int lpf_p2p_export(lpf_ctx *c, void *handle64, uint64_t *raw_ptr, int *device)
{
    if (!c || !handle64) { lpf::set_error("lpf_p2p_export: null argument"); return LPF_ERR_ARG; }
    if (!c->box) { lpf::set_error("lpf_p2p_export: single-rank context (or more than 16 ranks) has no mailbox"); return LPF_ERR_STATE; }
    CUDA_TRY(cudaSetDevice(c->dev));
    cudaIpcMemHandle_t mh;
    static_assert(sizeof(mh) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CUDA_TRY(cudaIpcGetMemHandle(&mh, c->box));
    std::memcpy(handle64, &mh, sizeof(mh));
    if (raw_ptr) *raw_ptr = (uint64_t)(uintptr_t)c->box;
    if (device) *device = c->dev;
    return LPF_OK;
}

int lpf_p2p_connect(lpf_ctx *c, const void *handles, const uint64_t *raw_ptrs, const int *devices, int same_process)
{
    if (!c || (!handles && !same_process)) { lpf::set_error("lpf_p2p_connect: null argument"); return LPF_ERR_ARG; }
    if (!c->box) { lpf::set_error("lpf_p2p_connect: context has no mailbox"); return LPF_ERR_STATE; }
    CUDA_TRY(cudaSetDevice(c->dev));
    return p2p_connect_impl(c, handles, raw_ptrs, devices, same_process);
}

int lpf_p2p_error(lpf_ctx *c)
{
    if (!c || !c->p2p_on) return 0;
    int e = 0;
    if (cudaMemcpy(&e, &c->p2p_local->error, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return e;
}

// ---- a1 -------------------------------------------------------------------------------------------
int lpf_pa_setup(lpf_ctx *c)
{
    if (!c) { lpf::set_error("null context"); return LPF_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(c->dev));
    const size_t nq = (size_t)c->ne * c->Q * c->Q * c->Q;
    if (nq) {
        pa_setup_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, c->stream>>>(c->p, c->ne, c->corners, c->jac, c->qd);
        c->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    c->affine_ok = false;
    if (c->ne && c->corners && !c->jac && c->qa) {      // affine fast path: element tensors + "is every element affine?"
        CUDA_TRY(cudaMemsetAsync(c->bad, 0, sizeof(int), c->stream));
        pa_affine_setup_kernel<<<(c->ne + 127) / 128, 128, 0, c->stream>>>(c->ne, c->corners, c->qa, c->bad);
        c->launches++;
        int not_affine = 0;
        CUDA_TRY(cudaMemcpyAsync(&not_affine, c->bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        c->affine_ok = !not_affine;
    }
    c->setup_done = true;
    return LPF_OK;
}

int lpf_pa_qdata(lpf_ctx *c, double *out)
{
    if (!c || !out) { lpf::set_error("lpf_pa_qdata: null argument"); return LPF_ERR_ARG; }
    if (!c->setup_done) { lpf::set_error("lpf_pa_qdata before lpf_pa_setup"); return LPF_ERR_STATE; }
    const size_t nq = (size_t)c->ne * c->Q * c->Q * c->Q;
    if (nq) {
        pa_qdata_export_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, c->stream>>>(c->p, c->ne, c->qd, out);
        c->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    return LPF_OK;
}

// ---- a8 -------------------------------------------------------------------------------------------
int lpf_pa_apply_E(lpf_ctx *c, const double *xE, double *yE)
{
    if (!c || !xE || !yE) { lpf::set_error("lpf_pa_apply_E: null argument"); return LPF_ERR_ARG; }
    if (!c->setup_done) { lpf::set_error("lpf_pa_apply_E before lpf_pa_setup"); return LPF_ERR_STATE; }
    CUDA_TRY(cudaSetDevice(c->dev));
    int rc = LPF_ERR_UNSUPPORTED;
#define LPF_CALL(P) rc = lpf_apply_E_p##P(c->qd, xE, yE, c->ne, c->stream)
    LPF_ORDER_SWITCH(c->p, LPF_CALL, lpf::set_error("unsupported order"))
#undef LPF_CALL
    if (rc == LPF_OK && c->ne) c->launches++;
    return rc;
}

int lpf_apply_L(lpf_ctx *c, const double *x, double *y)
{
    if (!c || !x || !y) { lpf::set_error("lpf_apply_L: null argument"); return LPF_ERR_ARG; }
    if (!c->setup_done) { lpf::set_error("lpf_apply_L before lpf_pa_setup"); return LPF_ERR_STATE; }
    return apply_full(c, false, x, y, nullptr, nullptr, true);
}

int lpf_apply_T(lpf_ctx *c, const double *x, double *y)
{
    if (!c || !x || !y) { lpf::set_error("lpf_apply_T: null argument"); return LPF_ERR_ARG; }
    if (!c->setup_done) { lpf::set_error("lpf_apply_T before lpf_pa_setup"); return LPF_ERR_STATE; }
    if (tail_mode(c) == 2 && c->sub_ne < 0) {
        // multi-GPU: the halo-sum rides on the element kernel, overlapped with the interior elements (p2p_dev.cuh)
        CUDA_TRY(cudaMemsetAsync(y, 0, sizeof(double) * c->ndof, c->stream));
        LPF_TRY(apply_with_tail(c, x, y, false));
    } else {
        LPF_TRY(apply_full(c, true, x, y, nullptr, nullptr, true));
        LPF_TRY(halo_sum(c, c->halo, y));
    }
    if (c->ness) {
        copy_at_kernel<<<(c->ness + 255) / 256, 256, 0, c->stream>>>(c->ness, c->ess, x, y);
        c->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    return LPF_OK;
}

// Pipelined host-buffer apply: x streams in over the H2D copy engine in dof ranges, element chunks start as soon as the
// ranges they read have landed, and every range of y leaves over the D2H engine as soon as its last-touching chunk is
// done -- PCIe is full duplex, so the call costs about ONE transfer instead of H2D + apply + D2H back to back.
static int apply_T_host_pipelined(lpf_ctx *c, const double *xh, double *yh)
{
    const int K = (int)c->hp_elem_end.size();
    if (!c->s_h2d) {
        CUDA_TRY(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&c->hp_ev_start, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&c->hp_ev_done, cudaEventDisableTiming));
    }
    if ((int)c->hp_ev_x.size() != K) {
        for (auto e : c->hp_ev_x) cudaEventDestroy(e);
        for (auto e : c->hp_ev_c) cudaEventDestroy(e);
        c->hp_ev_x.assign(K, nullptr); c->hp_ev_c.assign(K, nullptr);
        for (auto &e : c->hp_ev_x) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto &e : c->hp_ev_c) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    double *x = c->X, *y = c->tmp;
    // everything enqueued so far on the context stream (earlier users of X / tmp) precedes the copies
    CUDA_TRY(cudaEventRecord(c->hp_ev_start, c->stream));
    CUDA_TRY(cudaStreamWaitEvent(c->s_h2d, c->hp_ev_start, 0));
    CUDA_TRY(cudaStreamWaitEvent(c->s_d2h, c->hp_ev_start, 0));
    for (int k = 0, a = 0; k < K; k++) {                 // x: one copy per chunk, the part of its prefix that is new
        const int b = std::max(a, c->hp_x_end[k]);
        if (b > a) CUDA_TRY(cudaMemcpyAsync(x + a, xh + a, sizeof(double) * (b - a), cudaMemcpyHostToDevice, c->s_h2d));
        CUDA_TRY(cudaEventRecord(c->hp_ev_x[k], c->s_h2d));
        a = b;
    }
    CUDA_TRY(cudaMemsetAsync(y, 0, sizeof(double) * c->ndof, c->stream));
    int waited = 0, rc = LPF_OK;                         // the context stream has waited for the x pieces [0, waited)
    auto x_landed = [&](int dof_end) {                   // ... extend that to cover x[0, dof_end)
        while (waited < K && (waited == 0 || c->hp_x_end[waited - 1] < dof_end)) {
            const cudaError_t e = cudaStreamWaitEvent(c->stream, c->hp_ev_x[waited], 0);
            if (e != cudaSuccess) return e;
            waited++;
        }
        return cudaSuccess;
    };
    auto ess_rows = [&](const std::vector<lpf_ctx::HpRun> &runs) {      // essential rows of runs that are final now: y = x there
        for (const auto &r : runs) {
            if (r.eb <= r.ea) continue;
            const cudaError_t e = x_landed(r.b);
            if (e != cudaSuccess) return e;
            copy_at_kernel<<<(r.eb - r.ea + 255) / 256, 256, 0, c->stream>>>(r.eb - r.ea, c->ess + r.ea, x, y);
            c->launches++;
        }
        return cudaSuccess;
    };
    auto send_runs = [&](const std::vector<lpf_ctx::HpRun> &runs) {
        for (const auto &r : runs) {
            const cudaError_t e = cudaMemcpyAsync(yh + r.a, y + r.a, sizeof(double) * (r.b - r.a), cudaMemcpyDeviceToHost, c->s_d2h);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };
    for (int k = 0; k < K && rc == LPF_OK; k++) {
        CUDA_TRY(x_landed(c->hp_x_end[k]));
        c->sub_e0 = c->hp_elem_begin[k];
        c->sub_ne = c->hp_elem_end[k] - c->sub_e0;
        rc = apply_launch(c, c->gmap_c, x, y, nullptr, nullptr);
        c->sub_ne = -1;
        if (rc != LPF_OK) break;
        if (c->hp_final[k].empty()) continue;
        CUDA_TRY(ess_rows(c->hp_final[k]));
        CUDA_TRY(cudaEventRecord(c->hp_ev_c[k], c->stream));
        CUDA_TRY(cudaStreamWaitEvent(c->s_d2h, c->hp_ev_c[k], 0));
        CUDA_TRY(send_runs(c->hp_final[k]));
    }
    if (rc == LPF_OK && !c->hp_final[K].empty()) {   // multi-GPU: halo-sum, then the runs that hold shared dofs
        rc = halo_sum(c, c->halo, y);
        if (rc == LPF_OK) CUDA_TRY(ess_rows(c->hp_final[K]));
        CUDA_TRY(cudaEventRecord(c->hp_ev_done, c->stream));
        CUDA_TRY(cudaStreamWaitEvent(c->s_d2h, c->hp_ev_done, 0));
        CUDA_TRY(send_runs(c->hp_final[K]));
    } else if (rc == LPF_OK && c->nranks > 1) {
        rc = halo_sum(c, c->halo, y);                // a rank without shared dofs still takes part in the exchange
    }
    // join: the context stream continues only after both copy streams are drained
    CUDA_TRY(cudaEventRecord(c->hp_ev_done, c->s_d2h));
    CUDA_TRY(cudaStreamWaitEvent(c->stream, c->hp_ev_done, 0));
    CUDA_TRY(cudaEventRecord(c->hp_ev_done, c->s_h2d));
    CUDA_TRY(cudaStreamWaitEvent(c->stream, c->hp_ev_done, 0));
    LPF_TRY(rc);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return LPF_OK;
}

int lpf_apply_T_host(lpf_ctx *c, const double *xh, double *yh)
{
    if (!c || !xh || !yh) { lpf::set_error("lpf_apply_T_host: null argument"); return LPF_ERR_ARG; }
    if (!c->setup_done) { lpf::set_error("lpf_apply_T_host before lpf_pa_setup"); return LPF_ERR_STATE; }
    int rc;
    if (c->host_pipeline && !c->hp_elem_end.empty() && !c->deterministic) rc = apply_T_host_pipelined(c, xh, yh);
    else {
        const size_t nb = sizeof(double) * c->ndof;
        CUDA_TRY(cudaMemcpyAsync(c->X, xh, nb, cudaMemcpyHostToDevice, c->stream));
        LPF_TRY(lpf_apply_T(c, c->X, c->tmp));
        CUDA_TRY(cudaMemcpyAsync(yh, c->tmp, nb, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        rc = LPF_OK;
    }
    if (rc == LPF_OK && lpf_p2p_error(c) != 0) { lpf::set_error("lpf_apply_T_host: peer-memory exchange timed out (a neighbour rank stalled)"); return LPF_ERR_COMM; }
    return rc;
}

// ---- a2 -------------------------------------------------------------------------------------------
int lpf_diag(lpf_ctx *c, double *diag)
{
    if (!c || !diag) { lpf::set_error("lpf_diag: null argument"); return LPF_ERR_ARG; }
    if (!c->setup_done) { lpf::set_error("lpf_diag before lpf_pa_setup"); return LPF_ERR_STATE; }
    const size_t n = (size_t)c->ne * c->D * c->D * c->D;
    if (c->deterministic) {      // E-vector diagonal + ordered gather: the summation order of MFEM's CPU restriction-transpose
        if (n) {
            CUDA_TRY(cudaMemsetAsync(c->yE, 0, sizeof(double) * n, c->stream));
            pa_diag_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(c->p, c->ne, c->qd, nullptr, c->yE);
            c->launches++;
        }
        LPF_TRY(det_gather(c, nullptr, diag));
        return halo_sum(c, c->halo, diag);
    }
    CUDA_TRY(cudaMemsetAsync(diag, 0, sizeof(double) * c->ndof, c->stream));
    if (n) {
        pa_diag_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(c->p, c->ne, c->qd, c->gmap, diag);
        c->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    return halo_sum(c, c->halo, diag);
}

/* DiffusionIntegrator::AssembleDiagonalPA(Vector &diag): E-vector diagonal, ACCUMULATED into diagE[ne][D^3] */
int lpf_pa_diag_E(lpf_ctx *c, double *diagE)
{
    if (!c || !diagE) { lpf::set_error("lpf_pa_diag_E: null argument"); return LPF_ERR_ARG; }
    if (!c->setup_done) { lpf::set_error("lpf_pa_diag_E before lpf_pa_setup"); return LPF_ERR_STATE; }
    const size_t n = (size_t)c->ne * c->D * c->D * c->D;
    if (n) {
        pa_diag_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(c->p, c->ne, c->qd, nullptr, diagE);
        c->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    return LPF_OK;
}

int lpf_jacobi_setup(lpf_ctx *c)
{
    if (!c) { lpf::set_error("null context"); return LPF_ERR_ARG; }
    LPF_TRY(lpf_diag(c, c->tmp));
    CUDA_TRY(cudaMemsetAsync(c->bad, 0, sizeof(int), c->stream));
    dinv_kernel<<<vec_grid(c->ndof, c->sm_count), 256, 0, c->stream>>>(c->ndof, c->tmp, c->essmask, c->dinv, c->bad);
    c->launches++;
    int bad = 0;
    CUDA_TRY(cudaMemcpyAsync(&bad, c->bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (bad) { lpf::set_error("lpf_jacobi_setup: non-positive diagonal entry (MFEM_VERIFY in OperatorJacobiSmoother)"); return LPF_ERR_STATE; }
    c->jacobi_done = true;
    return LPF_OK;
}

int lpf_jacobi_dinv(lpf_ctx *c, double *out)
{
    if (!c || !out) { lpf::set_error("lpf_jacobi_dinv: null argument"); return LPF_ERR_ARG; }
    if (!c->jacobi_done) { lpf::set_error("lpf_jacobi_dinv before lpf_jacobi_setup"); return LPF_ERR_STATE; }
    CUDA_TRY(cudaMemcpyAsync(out, c->dinv, sizeof(double) * c->ndof, cudaMemcpyDeviceToDevice, c->stream));
    return LPF_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// PCG
// ------------------------------------------------------------------------------------------------
namespace {

int ess_fix(lpf_ctx *c)
{
    if (!c->ess_general || c->ness == 0) return LPF_OK;
    const int g = std::max(1, std::min((c->ness + 255) / 256, 256));
    pcg_ess_fix_kernel<<<g, 256, 0, c->stream>>>(c->ness, c->ess, c->d, c->ad, c->owned, c->den_slots, c->st);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return LPF_OK;
}

// cross-rank sum of a PCG scalar and what follows it (mode = P2P_RED_NOM / _BETA / _DEN)
int multi_reduce(lpf_ctx *c, int mode)
{
    if (c->p2p_on) {
        p2p_allreduce_kernel<<<1, 32, 0, c->stream>>>(c->p2p, &c->st->red[0], mode, c->st, c->den_slots);
        c->launches++;
        CUDA_TRY(cudaGetLastError());
        return LPF_OK;
    }
    if (!c->comm.ready()) { lpf::set_error("multi-rank context used before lpf_comm_init / lpf_p2p_connect"); return LPF_ERR_STATE; }
    if (mode == P2P_RED_DEN) {
        pcg_den_local_kernel<<<1, LPF_DEN_SLOTS, 0, c->stream>>>(c->st, c->den_slots);
        c->launches++;
        return c->comm.allreduce_sum(&c->st->red[1], 1, c->stream);
    }
    LPF_TRY(c->comm.allreduce_sum(&c->st->red[0], 1, c->stream));
    if (mode == P2P_RED_NOM) pcg_fin_nom_kernel<<<1, 1, 0, c->stream>>>(c->st);
    else pcg_fin_beta_kernel<<<1, 1, 0, c->stream>>>(c->st);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return LPF_OK;
}

// LL halo-sum of the PCG apply (+ the (d, A d) all-reduce) in one PDL-chained kernel
int halo_ll(lpf_ctx *c, bool pdl, double *v, bool with_den)
{
    const P2PPlanDev &pd = c->p2p_plan[0];
    const int g = std::max(1, std::min((c->halo.n_shared + 127) / 128, 2 * c->sm_count));
    CUDA_TRY(launch_ex(pdl, p2p_halo_ll_kernel, dim3(g), dim3(128), 0, c->stream, c->p2p, pd, 0, v, c->st, c->den_slots, (int)with_den));
    c->launches++;
    return LPF_OK;
}

// one CG iteration body: update (+ dot), direction, apply (+ den); all skip themselves once status != 0
int pcg_iteration(lpf_ctx *c)
{
    const int n = c->ndof, g = vec_grid(n, c->sm_count);
    const bool multi = c->nranks > 1;
    const uint8_t *no_mask = nullptr;
    const bool pdl = c->pdl != 0 && !c->ess_general && !c->deterministic;
    if (tail_mode(c) != 0) {
        // three launches, as on one GPU: the betanom all-reduce runs in the last block of the update kernel, the
        // halo-sum and the (d, A d) all-reduce inside the apply kernel (peer-memory stores + flags)
        CUDA_TRY(launch_ex(pdl, pcg_update_p2p_kernel, dim3(g), dim3(256), 0, c->stream, n, c->X, c->r, c->z, c->d, c->ad, c->dinv, c->owned, c->st, c->partials, c->p2p));
        CUDA_TRY(launch_ex(pdl, pcg_dir_kernel, dim3(g), dim3(256), 0, c->stream, n, c->z, c->d, c->ad, c->st, c->den_slots));
        c->launches += 2;
        c->pdl_now = pdl;
        const int rc = apply_with_tail(c, c->d, c->ad, true);
        c->pdl_now = false;
        return rc;
    }
    if (multi && c->p2p_on && !c->ess_general) {
        // four PDL-chained launches: update (+ betanom all-reduce in its last block), direction, apply, LL halo-sum
        // (+ (d, A d) all-reduce).  Two NVLink one-way latencies per iteration, no NCCL, no host involvement.
        CUDA_TRY(launch_ex(pdl, pcg_update_p2p_kernel, dim3(g), dim3(256), 0, c->stream, n, c->X, c->r, c->z, c->d, c->ad, c->dinv, c->owned, c->st, c->partials, c->p2p));
        CUDA_TRY(launch_ex(pdl, pcg_dir_kernel, dim3(g), dim3(256), 0, c->stream, n, c->z, c->d, c->ad, c->st, c->den_slots));
        c->launches += 2;
        c->pdl_now = pdl;
        const int rc = apply_full(c, true, c->d, c->ad, c->den_slots, &c->st->status, false);
        c->pdl_now = false;
        LPF_TRY(rc);
        return halo_ll(c, pdl, c->ad, true);
    }
    if (multi) {
        pcg_update_kernel<true><<<g, 256, 0, c->stream>>>(n, c->X, c->r, c->z, c->d, c->ad, c->dinv, c->owned, c->st, c->den_slots, c->partials);
        c->launches++;
        LPF_TRY(multi_reduce(c, P2P_RED_BETA));
        pcg_dir_kernel<<<g, 256, 0, c->stream>>>(n, c->z, c->d, c->ad, c->st, c->den_slots);
        c->launches++;
        LPF_TRY(apply_full(c, true, c->d, c->ad, c->den_slots, &c->st->status, false));
    } else {
        // single GPU: update -> direction -> apply chained by programmatic dependent launches (the ess_fix kernel of
        // general right-hand sides has no griddep_wait, so that path keeps plain stream order)
        CUDA_TRY(launch_ex(pdl, pcg_update_kernel<false>, dim3(g), dim3(256), 0, c->stream, n, c->X, c->r, c->z, c->d, c->ad, c->dinv, no_mask, c->st, c->den_slots, c->partials));
        CUDA_TRY(launch_ex(pdl, pcg_dir_kernel, dim3(g), dim3(256), 0, c->stream, n, c->z, c->d, c->ad, c->st, c->den_slots));
        c->launches += 2;
        c->pdl_now = pdl;
        const int rc = apply_full(c, true, c->d, c->ad, c->den_slots, &c->st->status, false);
        c->pdl_now = false;
        LPF_TRY(rc);
    }
    LPF_TRY(ess_fix(c));
    if (multi) {
        LPF_TRY(halo_sum(c, c->halo, c->ad));
        LPF_TRY(multi_reduce(c, P2P_RED_DEN));
    }
    CUDA_TRY(cudaGetLastError());
    return LPF_OK;
}

int pcg_graph_key(const lpf_ctx *c) { return c->ess_general + 2 * tail_mode(c) + 8 * c->deterministic; }

int pcg_chunk(lpf_ctx *c)
{
    if (!c->use_graph) {
        for (int i = 0; i < c->chunk; i++) LPF_TRY(pcg_iteration(c));
        return LPF_OK;
    }
    if (!c->pcg_graph || c->pcg_graph_chunk != c->chunk || c->pcg_graph_general != pcg_graph_key(c)) {
        if (c->pcg_graph) { cudaGraphExecDestroy(c->pcg_graph); c->pcg_graph = nullptr; }
        cudaGraph_t graph = nullptr;
        const long l0 = c->launches;
        CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        int rc = LPF_OK;
        for (int i = 0; i < c->chunk && rc == LPF_OK; i++) rc = pcg_iteration(c);
        cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
        c->pcg_graph_launches = (int)(c->launches - l0);
        c->launches = l0;
        if (rc != LPF_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        CUDA_TRY(ce);
        CUDA_TRY(cudaGraphInstantiate(&c->pcg_graph, graph, 0));
        cudaGraphDestroy(graph);
        c->pcg_graph_chunk = c->chunk;
        c->pcg_graph_general = pcg_graph_key(c);
    }
    CUDA_TRY(cudaGraphLaunch(c->pcg_graph, c->stream));
    c->launches += c->pcg_graph_launches;
    return LPF_OK;
}

// Persisting-L2 window over [z | d | A d] for the kernels enqueued on the context stream from here on (captured into the
// PCG graph's kernel nodes).  Per iteration these vectors are written and re-read by the next kernel (z: update ->
// direction; d: direction -> apply, update; A d: direction -> apply -> update): 72 of the iteration's 282 B/dof.  While
// they fit the persisting carve-out (<= 75 % of the 126 MB L2) that traffic never reaches HBM; beyond it a fraction does.
int pcg_l2_window(lpf_ctx *c, bool on)
{
    static int max_persist[16], max_window[16] = {0};
    static bool queried[16] = {false};
    int &mp = max_persist[c->dev & 15];
    if (!queried[c->dev & 15]) {
        queried[c->dev & 15] = true;
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, c->dev));
        mp = prop.persistingL2CacheMaxSize;
        max_window[c->dev & 15] = prop.accessPolicyMaxWindowSize;
    }
    if (mp <= 0) return LPF_OK;
    cudaStreamAttrValue attr;
    std::memset(&attr, 0, sizeof(attr));
    if (on && c->l2_persist) {
        const size_t carve = std::min((size_t)mp, c->zdad_bytes);
        static size_t carve_set[16] = {0};
        if (carve_set[c->dev & 15] != carve) { CUDA_TRY(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve)); carve_set[c->dev & 15] = carve; }
        const size_t win = std::min(c->zdad_bytes, (size_t)max_window[c->dev & 15]);
        attr.accessPolicyWindow.base_ptr = c->zdad;
        attr.accessPolicyWindow.num_bytes = win;
        attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)win);
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    } else {
        attr.accessPolicyWindow.num_bytes = 0;
    }
    CUDA_TRY(cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    return LPF_OK;
}

// Solve A_c X = B with X (ctx-owned c->X) as initial guess, B in c->Bv.  zero_guess_interior: the caller
// guarantees X is zero off the essential dofs, so A_c X == [0 ; X_ess] exactly and the initial-residual
// apply of CGSolver::Mult can be skipped without changing a single bit of r.
int pcg_run(lpf_ctx *c, double rel_tol, double abs_tol, int max_iter, bool zero_guess_interior, lpf_pcg_info *info)
{
    const int n = c->ndof, g = vec_grid(n, c->sm_count);
    const bool multi = c->nranks > 1;
    int applies = 0;
    LPF_TRY(pcg_l2_window(c, true));
    pcg_reset_kernel<<<1, 256, 0, c->stream>>>(c->st, c->den_slots, rel_tol, abs_tol, max_iter);
    c->launches++;
    const double *t = nullptr;
    if (!(zero_guess_interior && c->skip_zero_apply)) {
        LPF_TRY(lpf_apply_T(c, c->X, c->tmp));
        applies++;
        t = c->tmp;
    } else if (c->ness) {
        // r = B - [0 ; X_ess]: B_ess == X_ess by construction, so r_ess = 0: do it through tmp = [0 ; X_ess]
        CUDA_TRY(cudaMemsetAsync(c->tmp, 0, sizeof(double) * n, c->stream));
        copy_at_kernel<<<(c->ness + 255) / 256, 256, 0, c->stream>>>(c->ness, c->ess, c->X, c->tmp);
        c->launches++;
        t = c->tmp;
    }
    const bool fused = tail_mode(c) != 0;
    const bool ll_path = multi && c->p2p_on && !c->ess_general && !fused;
    if (fused) {
        pcg_init_p2p_kernel<<<g, 256, 0, c->stream>>>(n, c->Bv, t, c->dinv, c->owned, c->r, c->z, c->d, c->ad, c->st, c->partials, c->p2p);
        c->launches++;
        LPF_TRY(apply_with_tail(c, c->d, c->ad, true));
    } else if (ll_path) {
        pcg_init_p2p_kernel<<<g, 256, 0, c->stream>>>(n, c->Bv, t, c->dinv, c->owned, c->r, c->z, c->d, c->ad, c->st, c->partials, c->p2p);
        c->launches++;
        LPF_TRY(apply_full(c, true, c->d, c->ad, c->den_slots, &c->st->status, false));
        LPF_TRY(halo_ll(c, false, c->ad, true));
    } else if (multi) {
        pcg_init_kernel<true><<<g, 256, 0, c->stream>>>(n, c->Bv, t, c->dinv, c->owned, c->r, c->z, c->d, c->ad, c->st, c->partials);
        c->launches++;
        LPF_TRY(multi_reduce(c, P2P_RED_NOM));
    } else {
        pcg_init_kernel<false><<<g, 256, 0, c->stream>>>(n, c->Bv, t, c->dinv, nullptr, c->r, c->z, c->d, c->ad, c->st, c->partials);
        c->launches++;
    }
    if (!fused && !ll_path) {
        LPF_TRY(apply_full(c, true, c->d, c->ad, c->den_slots, &c->st->status, false));
        LPF_TRY(ess_fix(c));
        if (multi) {
            LPF_TRY(halo_sum(c, c->halo, c->ad));
            LPF_TRY(multi_reduce(c, P2P_RED_DEN));
        }
    }
    CUDA_TRY(cudaGetLastError());
    // Iterate in chunks of `pcg_chunk` graph-captured iterations.  The host polls one pinned status word per chunk, but
    // one chunk BEHIND: chunk k+1 is enqueued before the status after chunk k is awaited, so the GPU never idles for the
    // host round trip (~10 us per chunk, 3-4 % of a small-mesh solve).  Once the stopping rule has fired every kernel of
    // the extra chunk returns at its first instruction; iteration counts are exactly those of the sequential loop.
    if (!c->poll_ev[0]) {
        CUDA_TRY(cudaEventCreateWithFlags(&c->poll_ev[0], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&c->poll_ev[1], cudaEventDisableTiming));
    }
    int slot = 0;
    CUDA_TRY(cudaMemcpyAsync(&c->st_host[slot], c->st, sizeof(PcgState), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaEventRecord(c->poll_ev[slot], c->stream));
    for (;;) {
        // speculative next chunk + its status copy, then look at the previous status
        LPF_TRY(pcg_chunk(c));
        CUDA_TRY(cudaMemcpyAsync(&c->st_host[slot ^ 1], c->st, sizeof(PcgState), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaEventRecord(c->poll_ev[slot ^ 1], c->stream));
        CUDA_TRY(cudaEventSynchronize(c->poll_ev[slot]));
        if (c->st_host[slot].status != PCG_RUNNING) break;        // the chunk just enqueued degenerates to empty launches
        slot ^= 1;
    }
    CUDA_TRY(cudaEventSynchronize(c->poll_ev[slot ^ 1]));         // final state (unchanged by the no-op chunk)
    if (slot ^ 1) c->st_host[0] = c->st_host[1];
    LPF_TRY(pcg_l2_window(c, false));
    const PcgState &s = *c->st_host;
    if (s.comm_error || s.status == PCG_COMM_ERROR) {
        lpf::set_error("PCG: a flag wait of the peer-memory exchange timed out (a neighbour rank stalled or left the solve early)");
        return LPF_ERR_COMM;
    }
    // applies: 1 for the first A d (if the solve got that far) + one per completed direction update
    if (!(s.status == PCG_CONVERGED && s.final_iter == 0) && s.status != PCG_NOT_PD) applies += 1;
    applies += std::max(0, s.iter - 1) - (s.status == PCG_MAXITER ? 1 : 0);
    if (info) {
        info->iterations = s.final_iter;
        info->converged = s.status == PCG_CONVERGED;
        info->final_norm = std::sqrt(std::max(s.betanom, 0.0));
        info->initial_norm = std::sqrt(std::max(s.nom0, 0.0));
        info->applies = applies;
    }
    return LPF_OK;
}

int solve_from_X(lpf_ctx *c, double rel_tol, double abs_tol, int max_iter, lpf_pcg_info *info)
{
    // FormLinearSystem: c->X holds [0 ; x_ess].  B = -A X (unconstrained), B[ess] = X[ess]   (EliminateRHS)
    const int n = c->ndof;
    LPF_TRY(apply_full(c, false, c->X, c->tmp, nullptr, nullptr, true));
    LPF_TRY(halo_sum(c, c->halo, c->tmp));
    negate_kernel<<<vec_grid(n, c->sm_count), 256, 0, c->stream>>>(n, c->tmp, c->Bv);
    c->launches++;
    if (c->ness) {
        copy_at_kernel<<<(c->ness + 255) / 256, 256, 0, c->stream>>>(c->ness, c->ess, c->X, c->Bv);
        c->launches++;
    }
    c->ess_general = 0;
    LPF_TRY(pcg_run(c, rel_tol, abs_tol, max_iter, true, info));
    if (info) info->applies += 1;
    return LPF_OK;
}

}  // namespace

extern "C" {

int lpf_pcg(lpf_ctx *c, const double *B, double *X, double rel_tol, double abs_tol, int max_iter, lpf_pcg_info *info)
{
    if (!c || !B || !X) { lpf::set_error("lpf_pcg: null argument"); return LPF_ERR_ARG; }
    if (!c->jacobi_done) { lpf::set_error("lpf_pcg before lpf_jacobi_setup"); return LPF_ERR_STATE; }
    if (max_iter < 0) { lpf::set_error("lpf_pcg: negative max_iter"); return LPF_ERR_ARG; }
    const size_t nb = sizeof(double) * c->ndof;
    CUDA_TRY(cudaMemcpyAsync(c->X, X, nb, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->Bv, B, nb, cudaMemcpyDeviceToDevice, c->stream));
    c->ess_general = 1;
    LPF_TRY(pcg_run(c, rel_tol, abs_tol, max_iter, false, info));
    CUDA_TRY(cudaMemcpyAsync(X, c->X, nb, cudaMemcpyDeviceToDevice, c->stream));
    return LPF_OK;
}

int lpf_laplace_solve(lpf_ctx *c, double *phi, double rel_tol, double abs_tol, int max_iter, lpf_pcg_info *info)
{
    if (!c || !phi) { lpf::set_error("lpf_laplace_solve: null argument"); return LPF_ERR_ARG; }
    if (!c->jacobi_done) { lpf::set_error("lpf_laplace_solve before lpf_jacobi_setup"); return LPF_ERR_STATE; }
    const size_t nb = sizeof(double) * c->ndof;
    CUDA_TRY(cudaMemsetAsync(c->X, 0, nb, c->stream));
    if (c->ness) {
        copy_at_kernel<<<(c->ness + 255) / 256, 256, 0, c->stream>>>(c->ness, c->ess, phi, c->X);
        c->launches++;
    }
    LPF_TRY(solve_from_X(c, rel_tol, abs_tol, max_iter, info));
    CUDA_TRY(cudaMemcpyAsync(phi, c->X, nb, cudaMemcpyDeviceToDevice, c->stream));
    return LPF_OK;
}

int lpf_surface_dz(lpf_ctx *c, const double *phi, double *wt)
{
    if (!c || !phi || !wt) { lpf::set_error("lpf_surface_dz: null argument"); return LPF_ERR_ARG; }
    if (c->nsurf == 0) return LPF_OK;
    const int ns = c->nsurf;
    surface_dz_kernel<<<(ns + 127) / 128, 128, 0, c->stream>>>(c->p, ns, c->sd_off, c->sd_elem, c->sd_node, c->gmap, c->corners, c->jinv_z, phi, c->wsum);
    c->launches++;
    LPF_TRY(halo_sum(c, c->shalo, c->wsum));
    surface_div_kernel<<<(ns + 127) / 128, 128, 0, c->stream>>>(ns, c->surf_mult, c->wsum, wt);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return LPF_OK;
}

int lpf_rhs_setup(lpf_ctx *c, const lpf_rhs_params *p, const double *cgen, const double *cabs)
{
    if (!c || !p) { lpf::set_error("lpf_rhs_setup: null argument"); return LPF_ERR_ARG; }
    if (!c->corners && !c->jinv_z) { lpf::set_error("lpf_rhs_setup: surface derivative needs `corners` or `jinv_z` in the descriptor"); return LPF_ERR_STATE; }
    if (p->use_relaxation && (!cgen || !cabs || !(p->tau > 0.0))) { lpf::set_error("lpf_rhs_setup: relaxation needs cgen, cabs and tau > 0"); return LPF_ERR_ARG; }
    c->rhs.g = p->g; c->rhs.H = p->H; c->rhs.omega = p->omega; c->rhs.k = p->k; c->rhs.kx = p->kx_dir; c->rhs.ky = p->ky_dir;
    c->rhs.cwave = p->cwave; c->rhs.coth_kh = std::cosh(p->kh) / std::sinh(p->kh); c->rhs.T = p->T;
    c->rhs.inv_tau = p->use_relaxation ? 1.0 / p->tau : 0.0; c->rhs.n_ramp = p->n_ramp; c->rhs.relax = p->use_relaxation;
    c->rel_tol = p->rel_tol; c->abs_tol = p->abs_tol; c->max_iter = p->max_iter;
    if (p->use_relaxation && c->nsurf) {
        CUDA_TRY(cudaMemcpyAsync(c->cgen, cgen, sizeof(double) * c->nsurf, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(c->cabs, cabs, sizeof(double) * c->nsurf, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    c->use_cabsy = false;
    c->rhs_done = true;
    return LPF_OK;
}

int lpf_rhs_set_cabsy(lpf_ctx *c, const double *cabsy)
{
    if (!c) { lpf::set_error("lpf_rhs_set_cabsy: null context"); return LPF_ERR_ARG; }
    if (!c->rhs_done) { lpf::set_error("lpf_rhs_set_cabsy before lpf_rhs_setup"); return LPF_ERR_STATE; }
    c->use_cabsy = cabsy != nullptr;
    if (cabsy && c->nsurf) {
        CUDA_TRY(cudaMemcpyAsync(c->cabsy, cabsy, sizeof(double) * c->nsurf, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    return LPF_OK;
}

int lpf_envelope_reset(lpf_ctx *c)
{
    if (!c) { lpf::set_error("lpf_envelope_reset: null context"); return LPF_ERR_ARG; }
    if (c->nsurf) {
        fill_kernel<<<(c->nsurf + 255) / 256, 256, 0, c->stream>>>(c->nsurf, -1e300, c->env);
        c->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    return LPF_OK;
}

int lpf_envelope_update(lpf_ctx *c, const double *state)
{
    if (!c || !state) { lpf::set_error("lpf_envelope_update: null argument"); return LPF_ERR_ARG; }
    if (c->nsurf) {
        envelope_kernel<<<(c->nsurf + 255) / 256, 256, 0, c->stream>>>(c->nsurf, state, c->env);
        c->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    return LPF_OK;
}

int lpf_envelope_get(lpf_ctx *c, double *env_host, double scale)
{
    if (!c || !env_host) { lpf::set_error("lpf_envelope_get: null argument"); return LPF_ERR_ARG; }
    if (c->nsurf) {
        CUDA_TRY(cudaMemcpyAsync(env_host, c->env, sizeof(double) * c->nsurf, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        for (int i = 0; i < c->nsurf; i++) env_host[i] *= scale;
    }
    return LPF_OK;
}

static int rhs_impl(lpf_ctx *c, double t, const double *state, double *dstate, lpf_pcg_info *info)
{
    const int ns = c->nsurf;
    // eta_stage/phi_fs_stage -> phi on the free surface, interior zero (Transfer + FormLinearSystem)
    CUDA_TRY(cudaMemsetAsync(c->X, 0, sizeof(double) * c->ndof, c->stream));
    if (ns) {
        scatter_kernel<<<(ns + 255) / 256, 256, 0, c->stream>>>(ns, c->surf2vol, state + ns, c->X);
        c->launches++;
    }
    LPF_TRY(solve_from_X(c, c->rel_tol, c->abs_tol, c->max_iter, info));
    if (ns) {
        surface_dz_kernel<<<(ns + 127) / 128, 128, 0, c->stream>>>(c->p, ns, c->sd_off, c->sd_elem, c->sd_node, c->gmap, c->corners, c->jinv_z, c->X, c->wsum);
        c->launches++;
        LPF_TRY(halo_sum(c, c->shalo, c->wsum));
        surface_rhs_kernel<<<(ns + 127) / 128, 128, 0, c->stream>>>(ns, c->rhs, t, c->surf_mult, c->wsum, state, c->surf_xy, c->cgen, c->cabs,
                                                                        c->use_cabsy ? c->cabsy : nullptr, dstate);
        c->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return LPF_OK;
}

int lpf_rhs(lpf_ctx *c, double t, const double *state, double *dstate)
{
    if (!c || !state || !dstate) { lpf::set_error("lpf_rhs: null argument"); return LPF_ERR_ARG; }
    if (!c->rhs_done || !c->jacobi_done) { lpf::set_error("lpf_rhs before lpf_rhs_setup / lpf_jacobi_setup"); return LPF_ERR_STATE; }
    c->n_last = 1;
    return rhs_impl(c, t, state, dstate, &c->last_info[0]);
}

int lpf_rk4_step(lpf_ctx *c, double *x, double *t, double dt)
{
    if (!c || !x || !t) { lpf::set_error("lpf_rk4_step: null argument"); return LPF_ERR_ARG; }
    if (!c->rhs_done || !c->jacobi_done) { lpf::set_error("lpf_rk4_step before lpf_rhs_setup / lpf_jacobi_setup"); return LPF_ERR_STATE; }
    const int n2 = 2 * c->nsurf;
    const int g = std::max(1, (n2 + 255) / 256);
    const double ts[4] = {*t, *t + dt / 2, *t + dt / 2, *t + dt};
    c->n_last = 4;
    for (int s = 0; s < 4; s++) {
        LPF_TRY(rhs_impl(c, ts[s], s == 0 ? x : c->rk_y, c->rk_k, &c->last_info[s]));
        if (n2) {
            rk4_stage_kernel<<<g, 256, 0, c->stream>>>(n2, s, dt, x, c->rk_k, c->rk_y, c->rk_z);
            c->launches++;
        }
    }
    CUDA_TRY(cudaGetLastError());
    *t += dt;
    return LPF_OK;
}

int lpf_rk4_step_host(lpf_ctx *c, double *xh, double *t, double dt)
{
    if (!c || !xh || !t) { lpf::set_error("lpf_rk4_step_host: null argument"); return LPF_ERR_ARG; }
    const size_t nb = sizeof(double) * 2 * c->nsurf;
    if (nb) {
        std::memcpy(c->state_pinned, xh, nb);
        CUDA_TRY(cudaMemcpyAsync(c->state_dev, c->state_pinned, nb, cudaMemcpyHostToDevice, c->stream));
    }
    LPF_TRY(lpf_rk4_step(c, c->state_dev, t, dt));
    if (nb) {
        CUDA_TRY(cudaMemcpyAsync(c->state_pinned, c->state_dev, nb, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        std::memcpy(xh, c->state_pinned, nb);
    }
    return LPF_OK;
}

int lpf_last_solve_info(lpf_ctx *c, lpf_pcg_info info[4], int *nsolves)
{
    if (!c || !info || !nsolves) { lpf::set_error("lpf_last_solve_info: null argument"); return LPF_ERR_ARG; }
    for (int i = 0; i < c->n_last; i++) info[i] = c->last_info[i];
    *nsolves = c->n_last;
    return LPF_OK;
}

const double *lpf_phi_dev(lpf_ctx *c) { return c ? c->X : nullptr; }

int lpf_time_apply(lpf_ctx *c, const double *x, double *y, int reps, float *ms_total, float *ms_kernel, long *n_launches)
{
    if (!c || !x || !y || reps < 1) { lpf::set_error("lpf_time_apply: bad argument"); return LPF_ERR_ARG; }
    if (!c->setup_done) { lpf::set_error("lpf_time_apply before lpf_pa_setup"); return LPF_ERR_STATE; }
    const long l0 = c->launches;
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    for (int i = 0; i < reps; i++) LPF_TRY(lpf_apply_T(c, x, y));
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(cudaEventSynchronize(c->ev1));
    if (ms_total) CUDA_TRY(cudaEventElapsedTime(ms_total, c->ev0, c->ev1));
    if (n_launches) *n_launches = c->launches - l0;
    if (ms_kernel) {
        CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
        for (int i = 0; i < reps; i++) LPF_TRY(apply_launch(c, c->gmap_c, x, y, nullptr, nullptr));
        CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
        CUDA_TRY(cudaEventSynchronize(c->ev1));
        CUDA_TRY(cudaEventElapsedTime(ms_kernel, c->ev0, c->ev1));
    }
    return LPF_OK;
}

void *lpf_dev_alloc(size_t bytes)
{
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 8);
    if (e != cudaSuccess) { lpf::set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e)); return nullptr; }
    return p;
}
int lpf_dev_free(void *p) { CUDA_TRY(cudaFree(p)); return LPF_OK; }
int lpf_memcpy_h2d(void *d, const void *s, size_t n) { CUDA_TRY(cudaMemcpy(d, s, n, cudaMemcpyHostToDevice)); return LPF_OK; }
int lpf_memcpy_d2h(void *d, const void *s, size_t n) { CUDA_TRY(cudaMemcpy(d, s, n, cudaMemcpyDeviceToHost)); return LPF_OK; }
int lpf_memset_dev(void *d, int v, size_t n) { CUDA_TRY(cudaMemset(d, v, n)); return LPF_OK; }
void *lpf_host_alloc_pinned(size_t bytes)
{
    void *p = nullptr;
    cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 8);
    if (e != cudaSuccess) { lpf::set_error(std::string("cudaMallocHost: ") + cudaGetErrorString(e)); return nullptr; }
    return p;
}
int lpf_host_free_pinned(void *p) { CUDA_TRY(cudaFreeHost(p)); return LPF_OK; }
int lpf_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

}  // extern "C"
