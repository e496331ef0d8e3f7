// AddMultPA on E-vectors (a8: DiffusionIntegrator::AddMultPA(x_E, y_E), y_E += B^T G^T D (G, B) x_E) -- the adapter entry
// point lpf_pa_apply_E.  One batch of E elements per CTA, q-data read straight from global memory into registers ahead of
// the X / Y stages; same five stages and stage buffers as the persistent L-vector kernels, plain contractions.
#pragma once
#include "apply_cfg.cuh"

__device__ __forceinline__ double2 ldg_stream2(const double2 *p)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

template <int P, int E, bool PREFETCH>
__global__ void __launch_bounds__(ApplyCfg<P, E>::NT, 1)
pa_apply_evec_kernel(const double *__restrict__ qd, const double *__restrict__ x, double *__restrict__ y, int ne)
{
    using C = ApplyCfg<P, E>;
    constexpr int D = C::D, Q = C::Q, LX = C::LX, LY = C::LY, LZ = C::LZ;
    constexpr int D3 = D * D * D;
    extern __shared__ double smem[];
    const LpfOrderTab<P> &T = c_ot[0];
    const int tid = threadIdx.x;
    const int e0 = blockIdx.x * E;

    // ---- Z-role bookkeeping + early q-data loads (latency hidden behind the X and Y stages) ----
    const int ez = tid / LZ, q2 = tid - ez * LZ;
    const bool zvalid = (e0 + ez) < ne;
    const double2 *qsrc = reinterpret_cast<const double2 *>(qd) + ((size_t)(e0 + ez) * Q * 3) * LZ + q2;
    double2 qv[PREFETCH ? Q : 1][3];
    if (PREFETCH && zvalid) {
#pragma unroll
        for (int qz = 0; qz < Q; qz++)
#pragma unroll
            for (int c = 0; c < 3; c++) qv[qz][c] = ldg_stream2(qsrc + (qz * 3 + c) * LZ);
    }

    // ---- X stage: line (dz,dy) ----
    double xs[D];
    const int ex = tid / LX, lx = tid - ex * LX;
    const int xdz = lx / D, xdy = lx - xdz * D;
    const bool xvalid = (tid < E * LX) && (e0 + ex) < ne;
    if (xvalid) {
        const double *src = x + (size_t)(e0 + ex) * D3 + lx * D;
#pragma unroll
        for (int i = 0; i < D; i++) xs[i] = src[i];
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

    // ---- Y stage: line (dz,qx) ----
    const int ey = tid / LY, ly = tid - ey * LY;
    const int ydz = ly / Q, yqx = ly - ydz * Q;
    const bool yvalid = (tid < E * LY) && (e0 + ey) < ne;
    if (yvalid) {
        const double *a = smem + ey * C::ES + ydz * C::SAZ + yqx;
        double ua[D], ub[D];
#pragma unroll
        for (int i = 0; i < D; i++) { ua[i] = a[i * C::SAY]; ub[i] = a[C::SAA + i * C::SAY]; }
        double *b = smem + ey * C::ES + C::OFFB + ydz * C::SBZ + yqx;
#pragma unroll
        for (int q = 0; q < Q; q++) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int i = 0; i < D; i++) {
                s0 = fma(T.BG[2 * (q * D + i)], ua[i], s0);     // B_y B_x u
                s1 = fma(T.BG[2 * (q * D + i) + 1], ua[i], s1);     // G_y B_x u
                s2 = fma(T.BG[2 * (q * D + i)], ub[i], s2);     // B_y G_x u
            }
            b[q * C::SBY] = s0;
            b[C::SBA + q * C::SBY] = s1;
            b[2 * C::SBA + q * C::SBY] = s2;
        }
    }
    __syncthreads();

    // ---- Z stage: column (qy,qx): forward z, q-data, backward z, all in registers ----
    if (zvalid) {
        double *b = smem + ez * C::ES + C::OFFB + C::zcol(q2);
        double ubb[D], ubg[D], ugb[D], cbb[D], cbg[D], cgb[D];
#pragma unroll
        for (int i = 0; i < D; i++) {
            ubb[i] = b[i * C::SBZ]; ubg[i] = b[C::SBA + i * C::SBZ]; ugb[i] = b[2 * C::SBA + i * C::SBZ];
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
            double2 d0, d1, d2;
            if (PREFETCH) { d0 = qv[qz][0]; d1 = qv[qz][1]; d2 = qv[qz][2]; }
            else {
                d0 = ldg_stream2(qsrc + (qz * 3 + 0) * LZ);
                d1 = ldg_stream2(qsrc + (qz * 3 + 1) * LZ);
                d2 = ldg_stream2(qsrc + (qz * 3 + 2) * LZ);
            }
            // (D11, D21) (D31, D22) (D32, D33)
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
            b[i * C::SBZ] = cbb[i]; b[C::SBA + i * C::SBZ] = cbg[i]; b[2 * C::SBA + i * C::SBZ] = cgb[i];
        }
    }
    __syncthreads();

    // ---- Yt stage: line (dz,qx) ----
    if (yvalid) {
        const double *b = smem + ey * C::ES + C::OFFB + ydz * C::SBZ + yqx;
        double vbb[Q], vbg[Q], vgb[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) { vbb[q] = b[q * C::SBY]; vbg[q] = b[C::SBA + q * C::SBY]; vgb[q] = b[2 * C::SBA + q * C::SBY]; }
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

    // ---- Xt stage: line (dz,dy) + scatter-add ----
    if (xvalid) {
        const double *a = smem + ex * C::ES + xdz * C::SAZ + xdy * C::SAY;
        double ta[Q], tb[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) { ta[q] = a[q]; tb[q] = a[C::SAA + q]; }
        double *dst = y + (size_t)(e0 + ex) * D3 + lx * D;
#pragma unroll
        for (int i = 0; i < D; i++) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < Q; q++) { s = fma(T.BG[2 * (q * D + i)], ta[q], s); s = fma(T.BG[2 * (q * D + i) + 1], tb[q], s); }
            dst[i] += s;
        }
    }
}
