// Even-odd variant of the persistent TMA-staged PA apply kernel (same pipeline, staging and shared-memory
// layout as pa_apply_tma_kernel; only the 1-D contractions differ).
//
// Gauss-Lobatto nodes and Gauss-Legendre points are symmetric about 1/2, so B[Q-1-q][D-1-d] = B[q][d] and
// G[Q-1-q][D-1-d] = -G[q][d].  Splitting a line into its even and odd parts e = u + rev(u), o = u - rev(u)
// turns one Q x D contraction into a ceil(Q/2) x ceil(D/2) and a ceil(Q/2) x floor(D/2) one plus D + Q
// additions: 26 instead of 30 FP64 operations per line at p = 4, 64 instead of 90 at p = 8, and half as many
// coefficient loads (LDCU) -- the two instruction classes that made the kernel issue-bound (ncu: 46 % DFMA,
// 25 % LDCU, profiles/r01_apply_ncu.md).  The half-size tables are built on the host from B and G
// (apply_order.cu, lpf_apply_tables) directly from their definitions:
//   forward   out[q] = sum_d M[q][d] in[d]:  MeF[q][d] = (M[q][d] + M[q][D-1-d]) / 2, MoF[q][d] = (M[q][d] - M[q][D-1-d]) / 2
//   transpose out[d] = sum_q M[q][d] in[q]:  MeT[d][q] = (M[q][d] + M[Q-1-q][d]) / 2, MoT[d][q] = (M[q][d] - M[Q-1-q][d]) / 2
// (middle columns un-halved), and  out[j] = se + so,  out[N-1-j] = sigma (se - so)  with sigma = +1 for B, -1 for G.
#pragma once
#include "pa_apply_tma.cuh"

// out (+)= M in for a (anti)symmetric NO x NI matrix given by its even / odd half tables.
template <int NI, int NO, int SIGN, bool ACC>
__device__ __forceinline__ void eo_contract(const double *__restrict__ Me, const double *__restrict__ Mo,
                                            const double (&e)[(NI + 1) / 2], const double (&o)[NI / 2 > 0 ? NI / 2 : 1],
                                            double (&out)[NO])
{
    constexpr int NIC = (NI + 1) / 2, NIH = NI / 2, NOC = (NO + 1) / 2, NOH = NO / 2;
    constexpr int RS = ((((NI > NO ? NI : NO) + 1) / 2) + 1) & ~1;      // row stride of the half tables: LpfOrderTab::RS
#pragma unroll
    for (int j = 0; j < NOC; j++) {
        const bool mid = (j >= NOH);                  // middle output row (NO odd): only one parity survives
        double se = 0.0, so = 0.0;
        if (!(mid && SIGN < 0)) {
#pragma unroll
            for (int i = 0; i < NIC; i++) se = fma(Me[j * RS + i], e[i], se);
        }
        if (!(mid && SIGN > 0)) {
#pragma unroll
            for (int i = 0; i < NIH; i++) so = fma(Mo[j * RS + i], o[i], so);
        }
        if (ACC) out[j] += se + so; else out[j] = se + so;
        if (!mid) {
            const double r = SIGN > 0 ? se - so : so - se;
            if (ACC) out[NO - 1 - j] += r; else out[NO - 1 - j] = r;
        }
    }
}

template <int N>
__device__ __forceinline__ void eo_split(const double (&in)[N], double (&e)[(N + 1) / 2], double (&o)[N / 2 > 0 ? N / 2 : 1])
{
#pragma unroll
    for (int i = 0; i < N / 2; i++) { e[i] = in[i] + in[N - 1 - i]; o[i] = in[i] - in[N - 1 - i]; }
    if (N & 1) e[N / 2] = in[N / 2];
}

// AFF = true: affine fast path.  On an affine hex J is constant, so D(q) = w_q * (detJ J^-1 J^-T) and `qd` holds just the
// six entries of that element tensor ([ne][6], pa_affine_setup_kernel) instead of 6 Q^3 doubles: the q-data stream --
// 87 % of the kernel's HBM bytes at order 4 -- disappears and the kernel becomes FP64-issue-bound.  Used automatically
// when every element of the mesh is affine (all wave tanks of the reference); reported separately from the graded
// stored-q-data number.
// LAY: stage-buffer layout (apply_cfg.cuh; 1 = A buffer aliased into the B buffer).
// EQ: early release of the q-data buffer.  The buffer is dead as soon as every thread has multiplied its column by the D
// tensor -- before the backward z contraction, a third of the Z stage -- so the warps count themselves out on a shared-memory
// counter and the LAST one issues the bulk copy of the next batch's q-data right there instead of behind the stage barrier:
// the copy is in flight for a larger part of the batch (one q-data buffer per CTA is all that fits at orders 7, 8).  Measured:
// +1.5 % at order 7, -4...-10 % at orders 4, 8, 9 (the release point splits the Z stage's instruction schedule), so order 7 only.
// Also measured and NOT adopted: dropping the fence.proxy.async in front of the refills (these buffers are only ever read by
// the threads) -- 5-10 % slower at every order.
template <int P, int E, bool DEN, int MINB, bool AFF = false, bool DET = false, bool OVL = false, int TABS = 0, int LAY = 0, bool EQ = false>
__global__ void __launch_bounds__(ApplyCfg<P, E>::NT, MINB)
pa_apply_eo_kernel(const ApplyKArgs ka)
{
    using C = TmaCfg<P, E, AFF, LAY>;
    constexpr bool EQR = EQ && !AFF;
    constexpr int D = C::D, Q = C::Q, LX = C::LX, LY = C::LY, LZ = C::LZ;
    constexpr int DP3 = C::DP3, QE = C::QE, D3 = C::D3;
    const double *__restrict__ qd = ka.qd;
    const int *__restrict__ gmap = ka.gmap;
    const double *__restrict__ x = ka.x;
    double *__restrict__ y = ka.y;
    const int ne = ka.ne;
    constexpr int DC = (D + 1) / 2, DH = D / 2 > 0 ? D / 2 : 1, QC = (Q + 1) / 2, QH = Q / 2 > 0 ? Q / 2 : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sq = reinterpret_cast<double *>(smem_raw + C::OFF_Q);
    int *sidx = reinterpret_cast<int *>(smem_raw + C::OFF_IDX);
    double *smem = reinterpret_cast<double *>(smem_raw + C::OFF_WORK);
    uint64_t *bar_q = reinterpret_cast<uint64_t *>(smem_raw + C::OFF_BAR);
    uint64_t *bar_i = bar_q + 1;
    uint32_t *qcnt = reinterpret_cast<uint32_t *>(bar_q + 3);      // EQ: warps that are done with the q-data buffer
    volatile int *sflag = reinterpret_cast<volatile int *>(qcnt + 1);      // "this CTA finished last" (multi-GPU tails)

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
        *qcnt = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    int b = blockIdx.x;
    if (b >= nb) return;
    auto batch_elems = [&](int bb) { return min(E, ne - bb * E); };
    if (tid == 0) {
        const int n0 = batch_elems(b);
        mbar_expect_tx(bar_i, (uint32_t)(n0 * DP3 * 4));
        bulk_g2s_stream(sidx, gmap + (size_t)b * E * DP3, (uint32_t)(n0 * DP3 * 4), bar_i, l2pol);
        if (!AFF) {
            mbar_expect_tx(bar_q, (uint32_t)(n0 * QE * 8));
            bulk_g2s_stream(sq, qd + (size_t)b * E * QE, (uint32_t)(n0 * QE * 8), bar_q, l2pol);
        }
    }
    // Everything above touches only constant data (gather map, q-data): under programmatic dependent launch it overlaps
    // the tail of the previous kernel.  x, y and the PCG status are produced by that kernel: wait for it here.
    griddep_wait();
    griddep_launch();      // after the wait: at most ONE successor kernel is resident ahead of time
    if (ka.status != nullptr && *ka.status != 0) {      // solve already finished: drain the copies issued above and leave
        mbar_wait(bar_i, 0);
        if (!AFF) mbar_wait(bar_q, 0);
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

    uint32_t it = 0, zoff = 0;     // zoff: always zero, but only the loop knows (TABS)
    for (; b < nb; b += gridDim.x, it++, zoff = (zoff + gridDim.x) >> 31) {
        const int e0 = b * E;
        const int bn = b + gridDim.x;
        const int e0n = bn * E;
        const bool has_next = bn < nb;
        const bool xvalid = xrole && (e0 + ex) < ne;
        const bool yvalid = yrole && (e0 + ey) < ne;
        const bool zvalid = (e0 + ez) < ne;
        const bool xnext = xrole && has_next && (e0n + ex) < ne;
        const int cur = it & 1, nxt = cur ^ 1;
        // loop-variant (always zero) table offset: keeps the compiler from hoisting the coefficients out of the
        // batch loop into (too few) uniform registers, see pa_apply_tma.cuh; TABS: one table copy per stage (apply_cfg.cuh)
        const uint32_t zv = TABS ? zoff : (it >> 30);
        const LpfOrderTab<P> &T = c_ot[zv], &TY = c_ot[(TABS ? 1 : 0) + zv], &TZ = c_ot[(TABS ? 2 : 0) + zv],
                             &TZb = c_ot[(TABS ? 3 : 0) + zv], &TYt = c_ot[(TABS ? 4 : 0) + zv], &TXt = c_ot[(TABS ? 5 : 0) + zv];
        double da[6];                                   // AFF: element tensor, loaded two stages before its use
        if (AFF && zvalid) {
            const double *de = qd + (size_t)(e0 + ez) * 6;
#pragma unroll
            for (int i = 0; i < 6; i++) da[i] = __ldg(de + i);
        }

        // ---- X stage: (B_x u, G_x u) on the line (dz,dy) ----
        if (xvalid) {
            double e[DC], o[DH], sb[Q], sg[Q];
            eo_split<D>(xs, e, o);
            eo_contract<D, Q, +1, false>(T.BeF, T.BoF, e, o, sb);
            eo_contract<D, Q, -1, false>(T.GeF, T.GoF, e, o, sg);
            double *a = smem + ex * C::ES + xdz * C::SAZ + xdy * C::SAY;
#pragma unroll
            for (int q = 0; q < Q; q++) { a[q] = sb[q]; a[C::SAA + q] = sg[q]; }
        }
        __syncthreads();
        // gather map of the next batch (issued behind the first barrier after the previous batch's Xt stage, pa_apply_tma.cuh)
        if (tid == 0 && has_next) {
            const int n1 = batch_elems(bn);
            fence_proxy_async();
            mbar_expect_tx(bar_i + nxt, (uint32_t)(n1 * DP3 * 4));
            bulk_g2s_stream(sidx + nxt * E * DP3, gmap + (size_t)bn * E * DP3, (uint32_t)(n1 * DP3 * 4), bar_i + nxt, l2pol);
        }

        // ---- Y stage: line (dz,qx) ----
        if (yvalid) {
            const double *a = smem + ey * C::ES + ydz * C::SAZ + yqx;
            double ua[D], ub[D];
#pragma unroll
            for (int i = 0; i < D; i++) { ua[i] = a[i * C::SAY]; ub[i] = a[C::SAA + i * C::SAY]; }
            double e[DC], o[DH], s0[Q], s1[Q], s2[Q];
            eo_split<D>(ua, e, o);
            eo_contract<D, Q, +1, false>(TY.BeF, TY.BoF, e, o, s0);      // B_y B_x u
            eo_contract<D, Q, -1, false>(TY.GeF, TY.GoF, e, o, s1);      // G_y B_x u
            eo_split<D>(ub, e, o);
            eo_contract<D, Q, +1, false>(TY.BeF, TY.BoF, e, o, s2);      // B_y G_x u
            double *bb = smem + ey * C::ES + C::OFFB + ydz * C::SBZ + yqx;
#pragma unroll
            for (int q = 0; q < Q; q++) { bb[q * C::SBY] = s0[q]; bb[C::SBA + q * C::SBY] = s1[q]; bb[2 * C::SBA + q * C::SBY] = s2[q]; }
        }
        __syncthreads();

        // ---- Z stage: column (qy,qx): forward z for all levels, D tensor per level, backward z ----
        if (!AFF) mbar_wait(bar_q, it & 1);
        double *bb = smem + ez * C::ES + C::OFFB + C::zcol(q2);
        double g0[Q], g1[Q], g2[Q];
        if (zvalid) {
            const double2 *sqv = reinterpret_cast<const double2 *>(sq) + (size_t)ez * (QE / 2) + q2;
            {
                double u[D], e[DC], o[DH];
#pragma unroll
                for (int i = 0; i < D; i++) u[i] = bb[2 * C::SBA + i * C::SBZ];       // G_x B_y u  -> B_z
                eo_split<D>(u, e, o);
                eo_contract<D, Q, +1, false>(TZ.BeF, TZ.BoF, e, o, g0);
#pragma unroll
                for (int i = 0; i < D; i++) u[i] = bb[C::SBA + i * C::SBZ];           // B_x G_y u  -> B_z
                eo_split<D>(u, e, o);
                eo_contract<D, Q, +1, false>(TZ.BeF, TZ.BoF, e, o, g1);
#pragma unroll
                for (int i = 0; i < D; i++) u[i] = bb[i * C::SBZ];                    // B_x B_y u  -> G_z
                eo_split<D>(u, e, o);
                eo_contract<D, Q, -1, false>(TZ.GeF, TZ.GoF, e, o, g2);
            }
            if (AFF) {
                const int qy = q2 / Q, qx = q2 - qy * Q;
                const double wxy = T.qwts[qx] * T.qwts[qy];
#pragma unroll
                for (int qz = 0; qz < Q; qz++) {
                    const double w = wxy * T.qwts[qz];
                    const double a0 = g0[qz], a1 = g1[qz], a2 = g2[qz];
                    g0[qz] = w * (da[0] * a0 + da[1] * a1 + da[2] * a2);
                    g1[qz] = w * (da[1] * a0 + da[3] * a1 + da[4] * a2);
                    g2[qz] = w * (da[2] * a0 + da[4] * a1 + da[5] * a2);
                }
            } else {
#pragma unroll
                for (int qz = 0; qz < Q; qz++) {
                    const double2 d0 = sqv[(qz * 3 + 0) * LZ], d1 = sqv[(qz * 3 + 1) * LZ], d2 = sqv[(qz * 3 + 2) * LZ];
                    const double a0 = g0[qz], a1 = g1[qz], a2 = g2[qz];
                    g0[qz] = d0.x * a0 + d0.y * a1 + d1.x * a2;
                    g1[qz] = d0.y * a0 + d1.y * a1 + d2.x * a2;
                    g2[qz] = d1.x * a0 + d2.x * a1 + d2.y * a2;
                }
            }
        }
        if (EQR) {      // this warp is done with the q-data buffer; the last warp to say so refills it
            constexpr int NW = (C::NT + 31) / 32;
            const int wbase = tid & ~31, live = C::NT - wbase;
            __syncwarp(live >= 32 ? 0xffffffffu : ((1u << live) - 1u));
            if ((tid & 31) == 0) {
                uint32_t old;
                // relaxed on purpose: an acq_rel atomic costs a MEMBAR.ALL.CTA that also waits for the scatter REDs still in
                // flight (measured: -9 % at order 8).  Ordering comes from the shared-memory pipeline itself: a warp's LDS of
                // the q-data are processed before its own ATOMS, so whoever reads NW - 1 here knows every read is done.
                asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(qcnt)) : "memory");
                if (old == NW - 1) {
                    *qcnt = 0;      // next increment: behind at least one CTA-wide barrier
                    if (has_next) {
                        const int n1 = batch_elems(bn);
                        fence_proxy_async();
                        mbar_expect_tx(bar_q, (uint32_t)(n1 * QE * 8));
                        bulk_g2s_stream(sq, qd + (size_t)bn * E * QE, (uint32_t)(n1 * QE * 8), bar_q, l2pol);
                    }
                }
            }
        }
        if (zvalid) {
            {
                double c[D], e[QC], o[QH];
                eo_split<Q>(g0, e, o);
                eo_contract<Q, D, +1, false>(TZb.BeT, TZb.BoT, e, o, c);
#pragma unroll
                for (int i = 0; i < D; i++) bb[2 * C::SBA + i * C::SBZ] = c[i];
                eo_split<Q>(g1, e, o);
                eo_contract<Q, D, +1, false>(TZb.BeT, TZb.BoT, e, o, c);
#pragma unroll
                for (int i = 0; i < D; i++) bb[C::SBA + i * C::SBZ] = c[i];
                eo_split<Q>(g2, e, o);
                eo_contract<Q, D, -1, false>(TZb.GeT, TZb.GoT, e, o, c);
#pragma unroll
                for (int i = 0; i < D; i++) bb[i * C::SBZ] = c[i];
            }
        }
        __syncthreads();

        if (!AFF && !EQR && tid == 0 && has_next) {
            const int n1 = batch_elems(bn);
            fence_proxy_async();
            mbar_expect_tx(bar_q, (uint32_t)(n1 * QE * 8));
            bulk_g2s_stream(sq, qd + (size_t)bn * E * QE, (uint32_t)(n1 * QE * 8), bar_q, l2pol);
        }
        if (has_next) mbar_wait(bar_i + nxt, ((it + 1) >> 1) & 1);
        if (xnext) {
            const int *gi = sidx + nxt * E * DP3 + ex * DP3 + lx * D;
#pragma unroll
            for (int i = 0; i < D; i++) { const int g = gi[i]; xsn[i] = g >= 0 ? x[g] : 0.0; }
        }

        // ---- Yt stage: line (dz,qx) ----
        if (yvalid) {
            const double *bb = smem + ey * C::ES + C::OFFB + ydz * C::SBZ + yqx;
            double v[Q], e[QC], o[QH], ta[D], tb[D];
#pragma unroll
            for (int q = 0; q < Q; q++) v[q] = bb[q * C::SBY];
            eo_split<Q>(v, e, o);
            eo_contract<Q, D, +1, false>(TYt.BeT, TYt.BoT, e, o, ta);
#pragma unroll
            for (int q = 0; q < Q; q++) v[q] = bb[C::SBA + q * C::SBY];
            eo_split<Q>(v, e, o);
            eo_contract<Q, D, -1, true>(TYt.GeT, TYt.GoT, e, o, ta);
#pragma unroll
            for (int q = 0; q < Q; q++) v[q] = bb[2 * C::SBA + q * C::SBY];
            eo_split<Q>(v, e, o);
            eo_contract<Q, D, +1, false>(TYt.BeT, TYt.BoT, e, o, tb);
            double *a = smem + ey * C::ES + ydz * C::SAZ + yqx;
#pragma unroll
            for (int i = 0; i < D; i++) { a[i * C::SAY] = ta[i]; a[C::SAA + i * C::SAY] = tb[i]; }
        }
        __syncthreads();

        // ---- Xt stage + scatter-add ----
        if (xvalid) {
            const double *a = smem + ex * C::ES + xdz * C::SAZ + xdy * C::SAY;
            const int *gi = sidx + cur * E * DP3 + ex * DP3 + lx * D;
            double v[Q], e[QC], o[QH], yv[D];
#pragma unroll
            for (int q = 0; q < Q; q++) v[q] = a[q];
            eo_split<Q>(v, e, o);
            eo_contract<Q, D, +1, false>(TXt.BeT, TXt.BoT, e, o, yv);
#pragma unroll
            for (int q = 0; q < Q; q++) v[q] = a[C::SAA + q];
            eo_split<Q>(v, e, o);
            eo_contract<Q, D, -1, true>(TXt.GeT, TXt.GoT, e, o, yv);
#pragma unroll
            for (int i = 0; i < D; i++) {
                const int g = gi[i];
                if (DET) ka.yE[(size_t)(e0 + ex) * D3 + lx * D + i] = yv[i];
                if (g >= 0) {
                    if (!DET) red_add_f64(y + g, yv[i]);
                    if (DEN) part = fma(xs[i], yv[i], part);
                }
            }
        }
        if (xnext) {
#pragma unroll
            for (int i = 0; i < D; i++) xs[i] = xsn[i];
        }
        apply_batch_end<OVL>(ka.tail, ov, b, y);
    }

    if (DEN && ka.den_slots != nullptr) apply_den_epilogue<C::NT>(part, ka.den_slots, smem);
    // multi-GPU: halo-sum (+ PCG denominator all-reduce) over NVLink peer memory, riding on this kernel
    if (OVL && ka.tail.mode == 1) p2p_apply_tail_last(ka.tail, y, sflag);
    else if (OVL && ka.tail.mode == 2) p2p_if_finish(ka.tail, ov, y, sflag);
}
