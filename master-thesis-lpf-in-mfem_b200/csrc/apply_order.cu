// One order of the PA apply kernels: compiled once per order with -DLPF_ORDER=p (build.sh), see apply_api.hpp.
#ifndef LPF_ORDER
#error "compile with -DLPF_ORDER=<p>"
#endif
#include <cstdio>
#include <cstring>
#include <vector>

#include "apply_api.hpp"
#include "pa_apply_eo.cuh"
#include "pa_apply_evec.cuh"
#include "pa_apply_tma.cuh"

#define LPF_CAT_(a, b) a##b
#define LPF_CAT(a, b) LPF_CAT_(a, b)

namespace {

constexpr int P = LPF_ORDER;

// TABS < 0: the order's default (one coefficient-table copy per stage from order 7 up, from order 6 up in the kernels that
// carry the overlapped halo exchange -- the variants where ptxas otherwise falls back to per-thread LDC loads,
// profiles/r02_sass_opcodes.md)
template <int E, int MINB, bool EO, bool AFF = false, bool DET = false, int TABS = -1, bool WOVL = false, int LAY = 0, bool EQ = false>
int launch_persistent(LpfApplyArgs &a)
{
    using C = TmaCfg<P, E, AFF, LAY>;
    static_assert(EO || (LAY == 0 && !EQ), "layout / early-release options exist in the even-odd kernel only");
    constexpr int TS = TABS >= 0 ? TABS : (P >= 7 ? 1 : 0), TSO = TABS >= 0 ? TABS : (P >= 6 ? 1 : 0);
    static int blocks_per_sm[2][16] = {{0}};
    // WOVL: also instantiate the kernels that carry the multi-GPU exchange hooks (tuned default and affine kernels only; the
    // alternative variants of the tuning sweep exchange through the separate LL kernel, lpf_device.cu tail_mode)
    void (*kd)(const ApplyKArgs), (*kn)(const ApplyKArgs), (*kod)(const ApplyKArgs) = nullptr, (*kon)(const ApplyKArgs) = nullptr;
    if constexpr (EO) {
        kd = pa_apply_eo_kernel<P, E, true, MINB, AFF, DET, false, TS, LAY, EQ>; kn = pa_apply_eo_kernel<P, E, false, MINB, AFF, DET, false, TS, LAY, EQ>;
        if constexpr (WOVL) { kod = pa_apply_eo_kernel<P, E, true, MINB, AFF, DET, true, TSO, LAY, EQ>; kon = pa_apply_eo_kernel<P, E, false, MINB, AFF, DET, true, TSO, LAY, EQ>; }
    } else {
        kd = pa_apply_tma_kernel<P, E, true, MINB, DET, false>; kn = pa_apply_tma_kernel<P, E, false, MINB, DET, false>;
        if constexpr (WOVL) { kod = pa_apply_tma_kernel<P, E, true, MINB, DET, true>; kon = pa_apply_tma_kernel<P, E, false, MINB, DET, true>; }
    }
    if (!WOVL && a.k.tail.mode != 0) { lpf::set_error("this apply_variant has no exchange hooks"); return LPF_ERR_UNSUPPORTED; }
    const bool ovl = a.k.tail.mode != 0;             // multi-GPU: the halo-sum rides on this launch
    // resident CTAs per SM of the kernel that is actually launched (the kernels with the exchange hooks need a few more
    // registers: sizing the plain kernel's grid by them cost one CTA per SM at order 4 -- 14 % of a CG iteration)
    int &bps = blocks_per_sm[ovl ? 1 : 0][a.dev & 15];
    if (bps == 0) {
        for (auto k : {kd, kn, kod, kon}) if (k) CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        int b0 = 0, b1 = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b0, ovl ? kod : kd, C::NT, C::SMEM_BYTES));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, ovl ? kon : kn, C::NT, C::SMEM_BYTES));
        bps = std::max(1, std::min(b0, b1));
        if (a.verbose) fprintf(stderr, "lpf: apply kernel p=%d E=%d MINB=%d eo=%d aff=%d det=%d tabs=%d/%d lay=%d eq=%d ovl=%d: %d threads, %zu B dynamic smem, %d CTAs/SM\n", P, E, MINB, (int)EO, (int)AFF, (int)DET, TS, TSO, LAY, (int)EQ, (int)ovl, C::NT, (size_t)C::SMEM_BYTES, bps);
    }
    const int nb = (a.k.ne + E - 1) / E;
    a.threads = C::NT; a.smem = C::SMEM_BYTES; a.grid = 0;
    if (nb == 0) return LPF_OK;
    int grid = std::min(nb, a.max_ctas > 0 ? a.max_ctas : bps * a.sm_count);
    grid = std::min(grid, LPF_DEN_SLOTS);            // one (d, A d) slot per CTA
    a.grid = grid;
    if (a.k.tail.mode == 2) a.k.tail.n_if_batches = (a.k.tail.n_if_batches + E - 1) / E;     // elements -> batches
    CUDA_TRY(launch_ex(a.pdl, a.k.den_slots ? (ovl ? kod : kd) : (ovl ? kon : kn), dim3(grid), dim3(C::NT), C::SMEM_BYTES, a.stream, a.k));
    CUDA_TRY(cudaGetLastError());
    return LPF_OK;
}

// (E, CTAs/SM) of the tuned kernels, profiles/r02_sweep_orders.txt.  Even-odd contractions from order 3 up, plain
// contractions below; the affine fast path always uses the even-odd kernel.
template <bool AFF, bool DET>
int launch_default(LpfApplyArgs &a)
{
    if constexpr (AFF) {
        if constexpr (P == 1) return launch_persistent<16, 4, true, true, false, -1, true>(a);
        else if constexpr (P == 2) return launch_persistent<8, 4, true, true, false, -1, true>(a);
        else if constexpr (P == 3) return launch_persistent<8, 2, true, true, false, -1, true>(a);
        else if constexpr (P == 4) return launch_persistent<3, 4, true, true, false, -1, true>(a);
        else if constexpr (P == 5) return launch_persistent<3, 2, true, true, false, -1, true>(a);
        else if constexpr (P == 6) return launch_persistent<2, 3, true, true, false, -1, true>(a);
        else return launch_persistent<1, 2, true, true, false, -1, true>(a);
    } else {
        if constexpr (P == 1) return launch_persistent<16, 3, false, false, DET, -1, !DET>(a);
        else if constexpr (P == 2) return launch_persistent<8, 3, false, false, DET, -1, !DET>(a);
        else if constexpr (P == 3) return launch_persistent<8, 2, true, false, DET, -1, !DET>(a);
        else if constexpr (P == 4) return launch_persistent<3, 4, true, false, DET, -1, !DET>(a);
        else if constexpr (P == 5) return launch_persistent<3, 2, true, false, DET, -1, !DET>(a);
        else if constexpr (P == 6) return launch_persistent<2, 3, true, false, DET, -1, !DET>(a);
        // orders 7-9: stage buffers aliased (A inside B, apply_cfg.cuh) -> 4 / 3 / 2 CTAs per SM instead of 3 / 2 / 1;
        // order 7 also releases the q-data buffer early (the only order where that measured faster, profiles/r02_sweep_orders.txt)
        else if constexpr (P == 7) return launch_persistent<1, 4, true, false, DET, -1, !DET, 1, 1>(a);
        else if constexpr (P == 8) return launch_persistent<1, 3, true, false, DET, -1, !DET, 1, 0>(a);
        else if constexpr (P == 9) return launch_persistent<1, 2, true, false, DET, -1, !DET, 1, 0>(a);
        else return launch_persistent<1, 1, true, false, DET, -1, !DET>(a);
    }
}

}  // namespace

int LPF_CAT(lpf_apply_L_p, LPF_ORDER)(LpfApplyArgs &a)
{
    const int v = a.variant;
    if (a.k.yE != nullptr) return launch_default<false, true>(a);       // deterministic: E-vector output
    if (a.affine) return launch_default<true, false>(a);
    if (v == 0) return launch_default<false, false>(a);
    if (v == 20) {          // plain contractions on the TMA pipeline
        if constexpr (P == 1) return launch_persistent<16, 3, false>(a);
        else if constexpr (P == 2) return launch_persistent<8, 3, false>(a);
        else if constexpr (P == 3) return launch_persistent<5, 3, false>(a);
        else if constexpr (P == 4) return launch_persistent<3, 3, false>(a);
        else if constexpr (P == 5) return launch_persistent<2, 3, false>(a);
        else if constexpr (P == 6) return launch_persistent<2, 2, false>(a);
        else if constexpr (P <= 8) return launch_persistent<2, 1, false>(a);
        else return launch_persistent<1, 1, false>(a);
    }
    if (v >= 30 && v < 40) {      // even-odd contractions: one alternative (E, CTAs/SM) per order (tests, tuning sweep);
                                  // 33: the round-1 kernels of orders 7 / 8 (per-thread LDC coefficient loads) as evidence
        if constexpr (P == 1) return launch_persistent<16, 3, true>(a);
        else if constexpr (P == 2) return launch_persistent<8, 3, true>(a);
        else if constexpr (P == 3) return launch_persistent<5, 3, true>(a);
        else if constexpr (P == 4) return launch_persistent<2, 5, true>(a);
        else if constexpr (P == 5) return launch_persistent<2, 3, true>(a);
        else if constexpr (P == 6) return launch_persistent<2, 2, true, false, false, 1>(a);
        else if constexpr (P == 7) { if (v == 33) return launch_persistent<1, 2, true, false, false, 0>(a); return launch_persistent<1, 3, true>(a); }
        else if constexpr (P == 8) { if (v == 33) return launch_persistent<1, 2, true, false, false, 0>(a); return launch_persistent<1, 2, true>(a); }
        else return launch_default<false, false>(a);
    }
    if (v == 40 || v == 41) {     // the round-2 layout experiment, kept re-measurable (profiles/r02_sweep_orders.txt): stage buffer A aliased into B
                                  // (40), plus early q-data release (41), with the (E, CTAs/SM) the smaller CTA allows
        if constexpr (P < 5 || P > 9) return launch_default<false, false>(a);
#define LPF_EXP(E_, M_, T_) { if (v == 40) return launch_persistent<E_, M_, true, false, false, T_, false, 1, false>(a); return launch_persistent<E_, M_, true, false, false, T_, false, 1, true>(a); }
        else if constexpr (P == 5) LPF_EXP(3, 2, -1)
        else if constexpr (P == 6) LPF_EXP(2, 2, 1)
        else if constexpr (P == 7) LPF_EXP(1, 4, 1)
        else if constexpr (P == 8) LPF_EXP(1, 3, 1)
        else LPF_EXP(1, 2, 1)
#undef LPF_EXP
    }
    lpf::set_error("unknown apply_variant " + std::to_string(v));
    return LPF_ERR_ARG;
}

int LPF_CAT(lpf_apply_E_p, LPF_ORDER)(const double *qd, const double *xE, double *yE, int ne, cudaStream_t stream)
{
    constexpr int E = P == 1 ? 16 : P == 2 ? 16 : P == 3 ? 8 : P <= 5 ? 4 : P == 6 ? 4 : P <= 8 ? 2 : 1;
    constexpr bool PF = P <= 5;
    using C = ApplyCfg<P, E>;
    auto kern = pa_apply_evec_kernel<P, E, PF>;
    static bool attr_set[16] = {false};
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (!attr_set[dev & 15]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        attr_set[dev & 15] = true;
    }
    const int grid = (ne + E - 1) / E;
    if (grid == 0) return LPF_OK;
    kern<<<grid, C::NT, C::SMEM_BYTES, stream>>>(qd, xE, yE, ne);
    CUDA_TRY(cudaGetLastError());
    return LPF_OK;
}

// B, G: [Q][D] row-major; builds the interleaved {B, G} table and the even / odd half tables (pa_apply_eo.cuh) and uploads
// them to the CURRENT device
int LPF_CAT(lpf_apply_tables_p, LPF_ORDER)(const double *B, const double *G, const double *qwts)
{
    constexpr int D = P + 1, Q = P + 2, DC = (D + 1) / 2, DH = D / 2, QC = (Q + 1) / 2, QH = Q / 2, RS = LpfOrderTab<P>::RS;
    std::vector<LpfOrderTab<P>> tabs(LPF_TAB_COPIES);
    LpfOrderTab<P> &t = tabs[0];
    std::memset(&t, 0, sizeof(t));
    auto Bm = [&](int q, int d) { return B[(size_t)q * D + d]; };
    auto Gm = [&](int q, int d) { return G[(size_t)q * D + d]; };
    for (int i = 0; i < Q * D; i++) { t.BG[2 * i] = B[i]; t.BG[2 * i + 1] = G[i]; }
    for (int q = 0; q < Q; q++) t.qwts[q] = qwts[q];
    for (int q = 0; q < QC; q++) {
        for (int d = 0; d < DC; d++) {
            t.BeF[q * RS + d] = d < DH ? 0.5 * (Bm(q, d) + Bm(q, D - 1 - d)) : Bm(q, d);
            t.GeF[q * RS + d] = d < DH ? 0.5 * (Gm(q, d) + Gm(q, D - 1 - d)) : Gm(q, d);
        }
        for (int d = 0; d < DH; d++) {
            t.BoF[q * RS + d] = 0.5 * (Bm(q, d) - Bm(q, D - 1 - d));
            t.GoF[q * RS + d] = 0.5 * (Gm(q, d) - Gm(q, D - 1 - d));
        }
    }
    for (int d = 0; d < DC; d++) {
        for (int q = 0; q < QC; q++) {
            t.BeT[d * RS + q] = q < QH ? 0.5 * (Bm(q, d) + Bm(Q - 1 - q, d)) : Bm(q, d);
            t.GeT[d * RS + q] = q < QH ? 0.5 * (Gm(q, d) + Gm(Q - 1 - q, d)) : Gm(q, d);
        }
        for (int q = 0; q < QH; q++) {
            t.BoT[d * RS + q] = 0.5 * (Bm(q, d) - Bm(Q - 1 - q, d));
            t.GoT[d * RS + q] = 0.5 * (Gm(q, d) - Gm(Q - 1 - q, d));
        }
    }
    for (int i = 1; i < LPF_TAB_COPIES; i++) tabs[i] = tabs[0];
    CUDA_TRY(cudaMemcpyToSymbol(c_ot, tabs.data(), LPF_TAB_COPIES * sizeof(LpfOrderTab<P>)));
    return LPF_OK;
}
