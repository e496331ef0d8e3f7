// Persistent, software-pipelined variant of the PA apply kernel (a7+a8).
//
// Same five "one thread per 1-D line" stages and shared-memory layout as pa_apply_kernel, but each CTA
// loops over element batches (grid = resident CTAs per SM x SM count) and keeps the global loads of the
// NEXT batch in flight while it computes the current one:
//
//     iteration b:   X(b)  |  Y(b) + idx(b') loads  |  Z(b), then q-data(b') loads into the registers Z(b)
//                    just freed  |  Yt(b) + x(b') gathers  |  Xt(b) + scatter-add
//
// so the q-data stream (87 % of the bytes at p = 4) has ~half an iteration (the Yt, Xt, X, Y stages, several
// thousand cycles) to arrive, without any shared-memory staging and without needing high occupancy.  ncu on
// the one-batch-per-CTA kernel showed 34 % long-scoreboard + 26 % barrier stalls at 2.7 warps/SMSP
// (profiles/r01_apply_v3_ncu.txt); this kernel removes the exposed global latency from the critical path.
#pragma once
#include "pa_kernels.cuh"

template <int P, int E, bool DEN, int MINB>
__global__ void __launch_bounds__(ApplyCfg<P, E>::NT, MINB)
pa_apply_pipe_kernel(const double *__restrict__ qd, const int *__restrict__ gmap, const double *__restrict__ x,
                     double *__restrict__ y, int ne, double *__restrict__ den_slots, const int *__restrict__ status)
{
    using C = ApplyCfg<P, E>;
    constexpr int D = C::D, Q = C::Q, LX = C::LX, LY = C::LY, LZ = C::LZ;
    constexpr int D3 = D * D * D;
    extern __shared__ double smem[];
    if (status != nullptr && *status != 0) return;
    const LpfBasisTab &T = c_tab[P];
    const int tid = threadIdx.x;
    const int nb = (ne + E - 1) / E;

    // static roles
    const int ez = tid / LZ, q2 = tid - ez * LZ;
    const int ex = tid / LX, lx = tid - ex * LX;
    const int xdz = lx / D, xdy = lx - xdz * D;
    const bool xrole = tid < E * LX;
    const int ey = tid / LY, ly = tid - ey * LY;
    const int ydz = ly / Q, yqx = ly - ydz * Q;
    const bool yrole = tid < E * LY;

    int idx[D], idxn[D];
    double xs[D], xsn[D];
    double2 qv[Q][3];
    double part = 0.0;

    int b = blockIdx.x;
    // ---- prologue: loads of the first batch ----
    if (b < nb) {
        const int e0 = b * E;
        if (xrole && e0 + ex < ne) {
            const int *gi = gmap + (size_t)(e0 + ex) * C::DP3 + lx * D;
#pragma unroll
            for (int i = 0; i < D; i++) idx[i] = gi[i];
#pragma unroll
            for (int i = 0; i < D; i++) xs[i] = idx[i] >= 0 ? x[idx[i]] : 0.0;
        }
        if (e0 + ez < ne) {
            const double2 *qsrc = reinterpret_cast<const double2 *>(qd) + ((size_t)(e0 + ez) * Q * 3) * LZ + q2;
#pragma unroll
            for (int qz = 0; qz < Q; qz++)
#pragma unroll
                for (int c = 0; c < 3; c++) qv[qz][c] = ldg_stream2(qsrc + (qz * 3 + c) * LZ);
        }
    }

    for (; b < nb; b += gridDim.x) {
        const int e0 = b * E;
        const int bn = b + gridDim.x;
        const int e0n = bn * E;
        const bool xvalid = xrole && (e0 + ex) < ne;
        const bool yvalid = yrole && (e0 + ey) < ne;
        const bool zvalid = (e0 + ez) < ne;
        const bool xnext = xrole && bn < nb && (e0n + ex) < ne;
        const bool znext = bn < nb && (e0n + ez) < ne;

        // ---- X stage ----
        if (xvalid) {
            double *a = smem + ex * C::ES + xdz * C::SAZ + xdy * C::SAY;
#pragma unroll
            for (int q = 0; q < Q; q++) {
                double sb = 0.0, sg = 0.0;
#pragma unroll
                for (int i = 0; i < D; i++) { sb = fma(T.BG[2 * (q * D + i)], xs[i], sb); sg = fma(T.BG[2 * (q * D + i) + 1], xs[i], sg); }
                a[q] = sb;
                a[C::SAA + q] = sg;
            }
        }
        __syncthreads();

        // ---- Y stage (+ gather-map loads of the next batch) ----
        if (xnext) {
            const int *gi = gmap + (size_t)(e0n + ex) * C::DP3 + lx * D;
#pragma unroll
            for (int i = 0; i < D; i++) idxn[i] = gi[i];
        }
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
                    s0 = fma(T.BG[2 * (q * D + i)], ua[i], s0);
                    s1 = fma(T.BG[2 * (q * D + i) + 1], ua[i], s1);
                    s2 = fma(T.BG[2 * (q * D + i)], ub[i], s2);
                }
                bb[q * Q] = s0;
                bb[C::SBA + q * Q] = s1;
                bb[2 * C::SBA + q * Q] = s2;
            }
        }
        __syncthreads();

        // ---- Z stage ----
        if (zvalid) {
            double *bb = smem + ez * C::ES + C::OFFB + q2;
            double ubb[D], ubg[D], ugb[D], cbb[D], cbg[D], cgb[D];
#pragma unroll
            for (int i = 0; i < D; i++) {
                ubb[i] = bb[i * C::SBZ]; ubg[i] = bb[C::SBA + i * C::SBZ]; ugb[i] = bb[2 * C::SBA + i * C::SBZ];
                cbb[i] = 0.0; cbg[i] = 0.0; cgb[i] = 0.0;
            }
#pragma unroll
            for (int qz = 0; qz < Q; qz++) {
                double g0 = 0.0, g1 = 0.0, g2 = 0.0;
#pragma unroll
                for (int i = 0; i < D; i++) {
                    g0 = fma(T.BG[2 * (qz * D + i)], ugb[i], g0);
                    g1 = fma(T.BG[2 * (qz * D + i)], ubg[i], g1);
                    g2 = fma(T.BG[2 * (qz * D + i) + 1], ubb[i], g2);
                }
                const double2 d0 = qv[qz][0], d1 = qv[qz][1], d2 = qv[qz][2];
                const double f0 = d0.x * g0 + d0.y * g1 + d1.x * g2;
                const double f1 = d0.y * g0 + d1.y * g1 + d2.x * g2;
                const double f2 = d1.x * g0 + d2.x * g1 + d2.y * g2;
#pragma unroll
                for (int i = 0; i < D; i++) {
                    cgb[i] = fma(T.BG[2 * (qz * D + i)], f0, cgb[i]);
                    cbg[i] = fma(T.BG[2 * (qz * D + i)], f1, cbg[i]);
                    cbb[i] = fma(T.BG[2 * (qz * D + i) + 1], f2, cbb[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < D; i++) {
                bb[i * C::SBZ] = cbb[i]; bb[C::SBA + i * C::SBZ] = cbg[i]; bb[2 * C::SBA + i * C::SBZ] = cgb[i];
            }
        }
        // q-data of the next batch into the registers this stage just released
        if (znext) {
            const double2 *qsrc = reinterpret_cast<const double2 *>(qd) + ((size_t)(e0n + ez) * Q * 3) * LZ + q2;
#pragma unroll
            for (int qz = 0; qz < Q; qz++)
#pragma unroll
                for (int c = 0; c < 3; c++) qv[qz][c] = ldg_stream2(qsrc + (qz * 3 + c) * LZ);
        }
        __syncthreads();

        // ---- Yt stage (+ x gathers of the next batch) ----
        if (xnext) {
#pragma unroll
            for (int i = 0; i < D; i++) xsn[i] = idxn[i] >= 0 ? x[idxn[i]] : 0.0;
        }
        if (yvalid) {
            const double *bb = smem + ey * C::ES + C::OFFB + ydz * C::SBZ + yqx;
            double vbb[Q], vbg[Q], vgb[Q];
#pragma unroll
            for (int q = 0; q < Q; q++) { vbb[q] = bb[q * Q]; vbg[q] = bb[C::SBA + q * Q]; vgb[q] = bb[2 * C::SBA + q * Q]; }
            double *a = smem + ey * C::ES + ydz * C::SAZ + yqx;
#pragma unroll
            for (int i = 0; i < D; i++) {
                double ta = 0.0, tb = 0.0;
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    ta = fma(T.BG[2 * (q * D + i)], vbb[q], ta);
                    ta = fma(T.BG[2 * (q * D + i) + 1], vbg[q], ta);
                    tb = fma(T.BG[2 * (q * D + i)], vgb[q], tb);
                }
                a[i * C::SAY] = ta;
                a[C::SAA + i * C::SAY] = tb;
            }
        }
        __syncthreads();

        // ---- Xt stage + scatter-add ----
        if (xvalid) {
            const double *a = smem + ex * C::ES + xdz * C::SAZ + xdy * C::SAY;
            double ta[Q], tb[Q];
#pragma unroll
            for (int q = 0; q < Q; q++) { ta[q] = a[q]; tb[q] = a[C::SAA + q]; }
#pragma unroll
            for (int i = 0; i < D; i++) {
                double s = 0.0;
#pragma unroll
                for (int q = 0; q < Q; q++) { s = fma(T.BG[2 * (q * D + i)], ta[q], s); s = fma(T.BG[2 * (q * D + i) + 1], tb[q], s); }
                if (idx[i] >= 0) {
                    red_add_f64(y + idx[i], s);
                    if (DEN) part = fma(xs[i], s, part);
                }
            }
        }
        if (xnext) {
#pragma unroll
            for (int i = 0; i < D; i++) { idx[i] = idxn[i]; xs[i] = xsn[i]; }
        }
        __syncthreads();     // smem A is rewritten by X(b') next
    }

    if (DEN && den_slots != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        __shared__ double wsum[32];
        const int w = tid >> 5, nw = (C::NT + 31) >> 5;
        if ((tid & 31) == 0) wsum[w] = part;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int i = 0; i < nw; i++) s += wsum[i];
            atomicAdd(den_slots + (blockIdx.x & 255), s);
        }
    }
}
