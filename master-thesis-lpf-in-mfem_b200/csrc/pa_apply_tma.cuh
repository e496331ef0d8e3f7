// Persistent PA apply kernel with the quadrature data and the gather map staged in shared memory by bulk
// asynchronous copies (TMA 1-D: cp.async.bulk ... mbarrier::complete_tx::bytes; SASS UBLKCP).
//
// a7+a8 of SURVEY.md 8a: ElementRestriction::Mult + DiffusionIntegrator::AddMultPA +
// ElementRestriction::MultTranspose (+ the essential-dof masks of ConstrainedOperator) in one kernel.
//
// Apply kernel: E elements per CTA, Q^2 threads per element, "one thread per 1-D line":
//   X  stage: thread (dz,dy) gathers D dofs from x_L, forms B_x u and G_x u (Q values each) -> smem A
//   Y  stage: thread (dz,qx) contracts along y -> BB, BG, GB (Q values each)               -> smem B
//   Z  stage: thread (qy,qx) contracts along z, applies the 3x3 symmetric q-data tensor level by
//             level and contracts back along z entirely in registers                        -> smem B
//   Yt stage: thread (dz,qx) -> smem A;   Xt stage: thread (dz,dy) -> D results, scatter-added to y_L.
// Each thread loads D (or Q) values and performs D*Q (or 2-3x that) FMAs on them, so shared-memory
// traffic per FMA is ~4x lower than in a thread-per-point scheme.
//
// Each CTA loops over batches of E consecutive elements (grid = resident CTAs x SMs).  One batch's q-data
// is ONE contiguous 48 Q^3 E-byte block in HBM (layout [e][qz][c2][q2][2], see pa_kernels.cuh), so a single
// elected thread moves it with one bulk copy; the copy for batch b' = b + gridDim.x is issued as soon as
// the Z stage of batch b has consumed the buffer and lands while the CTA runs Yt(b), Xt(b), X(b'), Y(b').
// The (padded) gather map of b' is double buffered the same way, and the x gathers of b' are issued one
// stage ahead into registers.  No thread ever waits on a global load it has just issued, so the kernel does
// not depend on occupancy to hide HBM latency (ncu of the non-pipelined kernel: 34 % long-scoreboard, 26 %
// barrier stalls -- profiles/r01_apply_ncu.md).
#pragma once
#include "apply_cfg.cuh"

// ---- pieces shared by the persistent kernels ----
// x_e . y_e partial of this CTA -> its (d, A d) slot (one slot per CTA: grids are capped at LPF_DEN_SLOTS)
// (wsum: the stage buffers, free after the last batch -- no static shared memory, the order-8 CTA fills its third of the SM
// to within a few bytes)
template <int NT>
__device__ __forceinline__ void apply_den_epilogue(double part, double *__restrict__ den_slots, double *wsum)
{
    part = warp_sum_partial(part, NT);
    const int tid = threadIdx.x, w = tid >> 5, nw = (NT + 31) >> 5;
    __syncthreads();      // every thread is done with the stage buffers
    if ((tid & 31) == 0) wsum[w] = part;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int i = 0; i < nw; i++) s += wsum[i];
        red_add_f64(den_slots + (blockIdx.x & (LPF_DEN_SLOTS - 1)), s);
    }
}

// End of one batch.  The single-GPU kernels need NO barrier here: the X stage of the next batch writes exactly the
// shared-memory line its own thread has just read in the Xt stage (same thread <-> line map), the B buffer is not touched
// before the next Y stage, and the gather-map copy that reuses this batch's index buffer is issued behind the next
// barrier (one CTA-wide barrier less per batch: the barrier was the largest stall of the order-4 kernel).  The kernels that
// carry the overlapped halo exchange keep it for the bookkeeping of the interface batches.
template <bool OVL>
__device__ __forceinline__ void apply_batch_end(const P2PTail &tail, P2POverlap &ov, int b, const double *__restrict__ y)
{
    if (OVL && tail.mode == 2) {
        const bool ifb = b < tail.n_if_batches;
        if (ifb) __threadfence();
        __syncthreads();
        if (ifb && threadIdx.x == 0) atomicAdd(&tail.d.local->if_done, 1u);
        p2p_if_try_send(tail, ov, y);
    } else if (OVL) {
        __syncthreads();
    }
}

template <int P, int E, bool DEN, int MINB, bool DET = false, bool OVL = false>
__global__ void __launch_bounds__(ApplyCfg<P, E>::NT, MINB)
pa_apply_tma_kernel(const ApplyKArgs ka)
{
    using C = TmaCfg<P, E>;
    constexpr int D = C::D, Q = C::Q, LX = C::LX, LY = C::LY, LZ = C::LZ;
    constexpr int DP3 = C::DP3, QE = C::QE, D3 = C::D3;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sq = reinterpret_cast<double *>(smem_raw + C::OFF_Q);
    int *sidx = reinterpret_cast<int *>(smem_raw + C::OFF_IDX);
    double *smem = reinterpret_cast<double *>(smem_raw + C::OFF_WORK);
    uint64_t *bar_q = reinterpret_cast<uint64_t *>(smem_raw + C::OFF_BAR);
    uint64_t *bar_i = bar_q + 1;      // two of them
    volatile int *sflag = reinterpret_cast<volatile int *>(bar_q + 3) + 1;      // "this CTA finished last" (multi-GPU tails)
    const double *__restrict__ qd = ka.qd;
    const int *__restrict__ gmap = ka.gmap;
    const double *__restrict__ x = ka.x;
    double *__restrict__ y = ka.y;
    const int ne = ka.ne;

    const int tid = threadIdx.x;
    const int nb = (ne + E - 1) / E;
    const uint64_t l2pol = l2_evict_first_policy();

    const int ez = tid / LZ, q2 = tid - ez * LZ;
    const int ex = tid / LX, lx = tid - ex * LX;
    const int xdz = lx / D, xdy = lx - xdz * D;
    const bool xrole = tid < E * LX;
    const int ey = tid / LY, ly = tid - ey * LY;
    const int ydz = ly / Q, yqx = ly - ydz * Q;
    const bool yrole = tid < E * LY;

    if (tid == 0) {
        mbar_init(bar_q, 1); mbar_init(bar_i, 1); mbar_init(bar_i + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    int b = blockIdx.x;
    if (b >= nb) return;
    auto batch_elems = [&](int bb) { return min(E, ne - bb * E); };
    // ---- prologue: stage gather map and q-data of the first batch ----
    if (tid == 0) {
        const int n0 = batch_elems(b);
        mbar_expect_tx(bar_i, (uint32_t)(n0 * DP3 * 4));
        bulk_g2s_stream(sidx, gmap + (size_t)b * E * DP3, (uint32_t)(n0 * DP3 * 4), bar_i, l2pol);
        mbar_expect_tx(bar_q, (uint32_t)(n0 * QE * 8));
        bulk_g2s_stream(sq, qd + (size_t)b * E * QE, (uint32_t)(n0 * QE * 8), bar_q, l2pol);
    }
    // Everything above touches only constant data (gather map, q-data): under programmatic dependent launch it overlaps
    // the tail of the previous kernel.  x, y and the PCG status are produced by that kernel: wait for it here.
    griddep_wait();
    griddep_launch();      // after the wait: at most ONE successor kernel is resident ahead of time
    if (ka.status != nullptr && *ka.status != 0) {      // solve already finished: drain the copies issued above and leave
        mbar_wait(bar_i, 0);
        mbar_wait(bar_q, 0);
        return;
    }
    P2POverlap ov;
    if (OVL && ka.tail.mode == 2) p2p_if_begin(ka.tail, ov);
    double xs[D], xsn[D];
    double part = 0.0;
    mbar_wait(bar_i, 0);
    if (xrole && b * E + ex < ne) {
        const int *gi = sidx + ex * DP3 + lx * D;
#pragma unroll
        for (int i = 0; i < D; i++) { const int g = gi[i]; xs[i] = g >= 0 ? x[g] : 0.0; }
    }

    uint32_t it = 0;
    for (; b < nb; b += gridDim.x, it++) {
        const int e0 = b * E;
        const int bn = b + gridDim.x;
        const int e0n = bn * E;
        const bool has_next = bn < nb;
        const bool xvalid = xrole && (e0 + ex) < ne;
        const bool yvalid = yrole && (e0 + ey) < ne;
        const bool zvalid = (e0 + ez) < ne;
        const bool xnext = xrole && has_next && (e0n + ex) < ne;
        const int cur = it & 1, nxt = cur ^ 1;
        // The basis tables are addressed through a loop-variant (always zero) offset: without it the compiler
        // hoists all 60 B/G coefficients out of the batch loop, overflows the uniform register file and pays
        // ~480 R2UR/MOV instructions per element shuffling them back (ncu source page, profiles/r01_apply_ncu.md).
        const LpfOrderTab<P> &T = c_ot[it >> 30];
#define BGL(q, i) (reinterpret_cast<const double2 *>(T.BG)[(q) * D + (i)])     /* {B, G}[q][i]: one LDCU.128 */

        // ---- X stage ----
        if (xvalid) {
            double *a = smem + ex * C::ES + xdz * C::SAZ + xdy * C::SAY;
#pragma unroll
            for (int q = 0; q < Q; q++) {
                double sb = 0.0, sg = 0.0;
#pragma unroll
                for (int i = 0; i < D; i++) { const double2 c = BGL(q, i); sb = fma(c.x, xs[i], sb); sg = fma(c.y, xs[i], sg); }
                a[q] = sb;
                a[C::SAA + q] = sg;
            }
        }
        __syncthreads();
        // gather map of the next batch -> the other index buffer.  Its last readers were the Xt threads of the PREVIOUS batch
        // (their scatter indices); there is no barrier at the end of a batch any more (see apply_batch_end), so the copy is
        // issued here, behind the first barrier every thread passes after that Xt stage.
        if (tid == 0 && has_next) {
            const int n1 = batch_elems(bn);
            fence_proxy_async();
            mbar_expect_tx(bar_i + nxt, (uint32_t)(n1 * DP3 * 4));
            bulk_g2s_stream(sidx + nxt * E * DP3, gmap + (size_t)bn * E * DP3, (uint32_t)(n1 * DP3 * 4), bar_i + nxt, l2pol);
        }

        // ---- Y stage ----
        if (yvalid) {
            const double *a = smem + ey * C::ES + ydz * C::SAZ + yqx;
            double ua[D], ub[D];
#pragma unroll
            for (int i = 0; i < D; i++) { ua[i] = a[i * C::SAY]; ub[i] = a[C::SAA + i * C::SAY]; }
            double *bb = smem + ey * C::ES + C::OFFB + ydz * C::SBZ + yqx;
#pragma unroll
            for (int q = 0; q < Q; q++) {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
                for (int i = 0; i < D; i++) {
                    const double2 c = BGL(q, i);
                    s0 = fma(c.x, ua[i], s0);
                    s1 = fma(c.y, ua[i], s1);
                    s2 = fma(c.x, ub[i], s2);
                }
                bb[q * C::SBY] = s0;
                bb[C::SBA + q * C::SBY] = s1;
                bb[2 * C::SBA + q * C::SBY] = s2;
            }
        }
        __syncthreads();

        // ---- Z stage: q-data from shared memory ----
        mbar_wait(bar_q, it & 1);
        if (zvalid) {
            double *bb = smem + ez * C::ES + C::OFFB + C::zcol(q2);
            const double2 *sqv = reinterpret_cast<const double2 *>(sq) + (size_t)ez * (QE / 2) + q2;
            double ubb[D], ubg[D], ugb[D], cbb[D], cbg[D], cgb[D];
#pragma unroll
            for (int i = 0; i < D; i++) {
                ubb[i] = bb[i * C::SBZ]; ubg[i] = bb[C::SBA + i * C::SBZ]; ugb[i] = bb[2 * C::SBA + i * C::SBZ];
                cbb[i] = 0.0; cbg[i] = 0.0; cgb[i] = 0.0;
            }
#pragma unroll
            for (int qz = 0; qz < Q; qz++) {
                const double2 d0 = sqv[(qz * 3 + 0) * LZ], d1 = sqv[(qz * 3 + 1) * LZ], d2 = sqv[(qz * 3 + 2) * LZ];
                double g0 = 0.0, g1 = 0.0, g2 = 0.0;
#pragma unroll
                for (int i = 0; i < D; i++) {
                    const double2 c = BGL(qz, i);
                    g0 = fma(c.x, ugb[i], g0);
                    g1 = fma(c.x, ubg[i], g1);
                    g2 = fma(c.y, ubb[i], g2);
                }
                const double f0 = d0.x * g0 + d0.y * g1 + d1.x * g2;
                const double f1 = d0.y * g0 + d1.y * g1 + d2.x * g2;
                const double f2 = d1.x * g0 + d2.x * g1 + d2.y * g2;
#pragma unroll
                for (int i = 0; i < D; i++) {
                    const double2 c = BGL(qz, i);
                    cgb[i] = fma(c.x, f0, cgb[i]);
                    cbg[i] = fma(c.x, f1, cbg[i]);
                    cbb[i] = fma(c.y, f2, cbb[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < D; i++) {
                bb[i * C::SBZ] = cbb[i]; bb[C::SBA + i * C::SBZ] = cbg[i]; bb[2 * C::SBA + i * C::SBZ] = cgb[i];
            }
        }
        __syncthreads();

        // the q buffer is free: stream the next batch's q-data into it while Yt, Xt, X', Y' run
        if (tid == 0 && has_next) {
            const int n1 = batch_elems(bn);
            fence_proxy_async();
            mbar_expect_tx(bar_q, (uint32_t)(n1 * QE * 8));
            bulk_g2s_stream(sq, qd + (size_t)bn * E * QE, (uint32_t)(n1 * QE * 8), bar_q, l2pol);
        }
        // x gathers of the next batch (its gather map landed long ago)
        if (has_next) mbar_wait(bar_i + nxt, ((it + 1) >> 1) & 1);
        if (xnext) {
            const int *gi = sidx + nxt * E * DP3 + ex * DP3 + lx * D;
#pragma unroll
            for (int i = 0; i < D; i++) { const int g = gi[i]; xsn[i] = g >= 0 ? x[g] : 0.0; }
        }

        // ---- Yt stage ----
        if (yvalid) {
            const double *bb = smem + ey * C::ES + C::OFFB + ydz * C::SBZ + yqx;
            double vbb[Q], vbg[Q], vgb[Q];
#pragma unroll
            for (int q = 0; q < Q; q++) { vbb[q] = bb[q * C::SBY]; vbg[q] = bb[C::SBA + q * C::SBY]; vgb[q] = bb[2 * C::SBA + q * C::SBY]; }
            double *a = smem + ey * C::ES + ydz * C::SAZ + yqx;
#pragma unroll
            for (int i = 0; i < D; i++) {
                double ta = 0.0, tb = 0.0;
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    const double2 c = BGL(q, i);
                    ta = fma(c.x, vbb[q], ta);
                    ta = fma(c.y, vbg[q], ta);
                    tb = fma(c.x, vgb[q], tb);
                }
                a[i * C::SAY] = ta;
                a[C::SAA + i * C::SAY] = tb;
            }
        }
        __syncthreads();

        // ---- Xt stage + scatter-add ----
        if (xvalid) {
            const double *a = smem + ex * C::ES + xdz * C::SAZ + xdy * C::SAY;
            const int *gi = sidx + cur * E * DP3 + ex * DP3 + lx * D;
            double ta[Q], tb[Q];
#pragma unroll
            for (int q = 0; q < Q; q++) { ta[q] = a[q]; tb[q] = a[C::SAA + q]; }
#pragma unroll
            for (int i = 0; i < D; i++) {
                double s = 0.0;
#pragma unroll
                for (int q = 0; q < Q; q++) { const double2 c = BGL(q, i); s = fma(c.x, ta[q], s); s = fma(c.y, tb[q], s); }
                const int g = gi[i];
                if (DET) ka.yE[(size_t)(e0 + ex) * D3 + lx * D + i] = s;
                if (g >= 0) {
                    if (!DET) red_add_f64(y + g, s);
                    if (DEN) part = fma(xs[i], s, part);
                }
            }
        }
        if (xnext) {
#pragma unroll
            for (int i = 0; i < D; i++) xs[i] = xsn[i];
        }
        apply_batch_end<OVL>(ka.tail, ov, b, y);
#undef BGL
    }

    if (DEN && ka.den_slots != nullptr) apply_den_epilogue<C::NT>(part, ka.den_slots, smem);
    // multi-GPU: halo-sum (+ PCG denominator all-reduce) over NVLink peer memory, riding on this kernel
    if (OVL && ka.tail.mode == 1) p2p_apply_tail_last(ka.tail, y, sflag);
    else if (OVL && ka.tail.mode == 2) p2p_if_finish(ka.tail, ov, y, sflag);
}
