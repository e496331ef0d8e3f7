// Configuration, per-order coefficient tables and TMA helpers shared by the PA apply kernels.  One translation unit per
// order (apply_order.cu compiled with -DLPF_ORDER=p) includes this file, so the __constant__ tables below exist once
// per order and hold that order only.
#pragma once
#include "apply_api.hpp"
#include "dev_util.cuh"
#include "p2p_dev.cuh"

__host__ __device__ constexpr int lpf_pad_to(int v, int target_mod16)
{
    // smallest w >= v with w % 16 == target_mod16 % 16
    int w = v;
    while ((w & 15) != (target_mod16 & 15)) w++;
    return w;
}

// Strides of the two stage buffers, per order, picked by the bank-conflict model tools/smem_layout_sim.py (shared-memory
// wavefronts per element, old formula-padded layout -> this table):  p1 48 -> 32, p2 90 -> 63, p3 230 -> 167, p4 319 -> 322
// (smaller: 4 CTAs per SM), p5 560 -> 496, p6 680 -> 642, p7 1006 -> 862, p8 1394 -> 1394 (smaller).  Measured at order 3:
// 77 % -> 89 % of the HBM roofline from the layout alone.  Orders 9, 10: formula (odd strides), not tuned.
//   {SAY, SAZ, SBZ, PAD}: A = [arr 2][dz][dy][qx] with strides SAZ, SAY, 1; B = [arr 3][dz][qy][qx] with strides SBZ, Q, 1
__host__ __device__ constexpr int lpf_smem_stride(int p, int which)
{
    constexpr int T[11][4] = {{0, 0, 0, 0}, {4, 8, 12, 1}, {5, 20, 20, 0}, {5, 28, 26, 1}, {7, 35, 38, 12},
                              {7, 42, 55, 4}, {9, 72, 72, 0}, {10, 89, 89, 0}, {10, 90, 106, 0}, {11, 111, 123, 0}, {13, 145, 145, 0}};
    return T[p][which];
}

// Aliased layout (LAY = 1, round 2): the A buffer lives INSIDE the B buffer, A[k][dz][dy][qx] = B[k][dz][qy = dy][qx].  The
// shared-memory words a Y-stage thread (dz, qx) reads from A (2 D of them) are then a subset of the words it writes to B
// (3 Q) -- and the other way round in the Yt stage -- so every thread overwrites only what it has itself read before, no
// extra barrier is needed and the A buffer's 2 D SAZ doubles per element disappear: 28.5 -> 17.1 KB at order 7 (four CTAs
// per SM instead of three), 35.9 -> 22.9 KB at order 8 (three instead of two).  B rows get their own stride SBY >= Q (odd
// where Q is even, or the X stage's row stores collide).  {SBY, SBZ, PAD} from tools/smem_layout_sim.py <p> <E> alias:
// wavefronts per element p4 322, p5 525, p6 810, p7 1006, p8 1394 (separate buffers: 322 / 496 / 642 / 862 / 1394).  Measured
// (profiles/r02_sweep_orders.txt): orders 7 / 8 / 9 gain 10-60 %; orders 4-6 lose 5-13 % (same CTAs per SM, more conflicts --
// order 6 even with the most compact variant that fits three CTAs per SM), so they keep separate buffers.
__host__ __device__ constexpr int lpf_smem_alias(int p, int which)
{
    constexpr int T[11][3] = {{0, 0, 0}, {3, 9, 0}, {4, 16, 0}, {5, 28, 9}, {7, 54, 4}, {7, 55, 12}, {9, 72, 0},
                              {9, 89, 0}, {10, 106, 0}, {11, 123, 0}, {13, 156, 0}};
    return T[p][which];
}
// the layout the tuned kernel of an order uses (apply_order.cu launch_default)
__host__ __device__ constexpr int lpf_default_layout(int p) { return p >= 7 && p <= 9 ? 1 : 0; }

template <int P, int E, int LAY = 0>
struct ApplyCfg {
    static constexpr int D = P + 1, Q = P + 2;
    static constexpr int LX = D * D, LY = D * Q, LZ = Q * Q;
    static constexpr int NT = E * LZ;
    static constexpr int D3 = D * D * D;
    static constexpr int DP3 = (D * D * D + 3) & ~3;    // gather-map row stride (rows padded to 16 bytes for bulk copies)
    static constexpr bool ALIAS = LAY == 1;
    static constexpr int SBY = ALIAS ? lpf_smem_alias(P, 0) : Q;               // B = [arr 3][dz][qy][qx], strides SBA, SBZ, SBY, 1
    static constexpr int SBZ = ALIAS ? lpf_smem_alias(P, 1) : lpf_smem_stride(P, 2);
    static constexpr int SBA = D * SBZ;
    static constexpr int SAY = ALIAS ? SBY : lpf_smem_stride(P, 0);            // A = [arr 2][dz][dy][qx], strides SAA, SAZ, SAY, 1
    static constexpr int SAZ = ALIAS ? SBZ : lpf_smem_stride(P, 1);
    static constexpr int SAA = ALIAS ? SBA : D * SAZ;
    static constexpr int OFFB = ALIAS ? 0 : 2 * SAA;
    static constexpr int ES = (ALIAS ? 3 * SBA + lpf_smem_alias(P, 2) : 2 * SAA + 3 * SBA + lpf_smem_stride(P, 3));      // element stride
    static constexpr size_t SMEM_BYTES = (size_t)E * ES * sizeof(double);
    static_assert(SAY >= Q && SAZ >= D * SAY - (SAY - Q) && SBZ >= Q * SBY - (SBY - Q), "stage-buffer strides too small");
    // Z-stage column (qy, qx) of thread q2 = qy Q + qx inside a B slab
    __device__ static __forceinline__ int zcol(int q2) { if constexpr (SBY == Q) return q2; else return (q2 / Q) * SBY + (q2 % Q); }
};

template <int P, int E, bool AFF = false, int LAY = 0>      // AFF: affine fast path, no q-data staging area (pa_apply_eo.cuh)
struct TmaCfg : ApplyCfg<P, E, LAY> {
    using B = ApplyCfg<P, E, LAY>;
    static constexpr int QE = 6 * B::Q * B::Q * B::Q;               // doubles of q-data per element
    // byte offsets inside dynamic shared memory
    static constexpr size_t OFF_Q = 0;                                                  // [E][QE] doubles (16B aligned)
    static constexpr size_t OFF_IDX = OFF_Q + (AFF ? (size_t)0 : (size_t)E * QE * 8);   // [2][E][DP3] ints
    static constexpr size_t OFF_WORK = (OFF_IDX + (size_t)2 * E * B::DP3 * 4 + 15) & ~(size_t)15;
    static constexpr size_t OFF_BAR = (OFF_WORK + (size_t)E * B::ES * 8 + 15) & ~(size_t)15;   // 3 mbarriers + the q-data release counter
    static constexpr size_t SMEM_BYTES = OFF_BAR + 32;
};

// ---- per-order coefficient tables ----------------------------------------------------------------------------------
// {B, G} interleaved per (q, d): one 16-byte uniform load (LDCU.128) feeds both FMAs that a stage issues on the same
// (q, d) pair.  Even / odd half tables of the (anti)symmetric B and G: see pa_apply_eo.cuh.
template <int P>
struct __align__(16) LpfOrderTab {
    static constexpr int D = P + 1, Q = P + 2, MH = (Q + 1) / 2;
    // rows of the half tables are padded to an even number of doubles and every array has an even length, so that two
    // neighbouring coefficients of a row can be fetched by ONE 16-byte uniform load (LDCU.128)
    static constexpr int RS = (MH + 1) & ~1;
    double BG[2 * Q * D];
    double qwts[Q + (Q & 1)];
    double BeF[MH * RS], BoF[MH * RS], GeF[MH * RS], GoF[MH * RS];   // [QC][RS] (DC used), [QC][RS] (DH used)
    double BeT[MH * RS], BoT[MH * RS], GeT[MH * RS], GoT[MH * RS];   // [DC][RS] (QC used), [DC][RS] (QH used)
};

#ifdef LPF_ORDER
// Identical copies of the table.  The kernels address copy `s + z` where z is a loop-variant value that is always zero
// (the batch counter >> 30, or a dedicated loop-carried variable): without it the compiler hoists all coefficients out of
// the batch loop, overflows the uniform register file and shuffles them back with R2UR / MOV (pa_apply_tma.cuh).  With
// TABS = 1 (orders >= 7) every stage s = 0..5 reads its own copy, so the coefficient loads of different stages cannot be
// merged into values that stay live across the whole batch (that is what pushed orders 7, 8 to 250 registers and
// per-thread LDC loads, profiles/r02_sass_opcodes.md).
#define LPF_TAB_COPIES 7
static __constant__ LpfOrderTab<LPF_ORDER> c_ot[LPF_TAB_COPIES];
#endif

// ---- TMA 1-D bulk copies + mbarriers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// L2 evict-first policy for the streamed data: q-data and gather maps are read exactly once per apply, so they should not
// push the vectors (x gathered / y scatter-added by up to 8 elements each, and reused by the next PCG kernels) out of L2
// (measured: -3 % per CG iteration at 2.25 M dofs).
__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_stream(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint64_t pol)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
