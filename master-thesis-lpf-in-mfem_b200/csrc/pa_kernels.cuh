// sm_100a kernels for the partial-assembly DiffusionIntegrator on H1 hexes.
//
//   pa_setup_kernel   a1  DiffusionIntegrator::AssemblePA  (Solvers/PF_linear_par_partial.cpp:118-121)
//   pa_apply_kernel   a7+a8 (+a5 masks)  ElementRestriction::Mult + AddMultPA + MultTranspose fused
//   pa_diag_kernel    a2  AssembleDiagonalPA + restriction-transpose (:124)
//
// q-data layout in HBM (internal, produced by pa_setup_kernel):
//     qd[e][qz][c2][q2][2]   with q2 = qx + Q*qy, c2 in {0,1,2} holding components (2*c2, 2*c2+1) of
//     (D11, D21, D31, D22, D32, D33): one element is one contiguous 48*Q^3-byte block, and the thread
//     that owns quadrature column (qx,qy) reads three 16-byte vectors per z-level, fully coalesced
//     across the CTA.
//
// Apply kernel: E elements per CTA, Q^2 threads per element, "one thread per 1-D line":
//   X  stage: thread (dz,dy) gathers D dofs from x_L, forms B_x u and G_x u (Q values each) -> smem A
//   Y  stage: thread (dz,qx) contracts along y -> BB, BG, GB (Q values each)               -> smem B
//   Z  stage: thread (qy,qx) contracts along z, applies the 3x3 symmetric q-data tensor level by
//             level and contracts back along z entirely in registers                        -> smem B
//   Yt stage: thread (dz,qx) -> smem A;   Xt stage: thread (dz,dy) -> D results, scatter-added to y_L.
// Each thread loads D (or Q) values and performs D*Q (or 2-3x that) FMAs on them, so shared-memory
// traffic per FMA is ~4x lower than in a thread-per-point scheme; B/G come from constant memory as
// immediate operands of the unrolled FMAs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Scatter-add without a return value: always the fire-and-forget reduction (SASS REDG), never the round-trip ATOMG.
// (With plain atomicAdd the compiler switched to ATOMG as soon as the kernel also read y elsewhere -- the fused
// halo tail -- which cost 30 % of the apply kernel's bandwidth; profiles/r01_apply_ncu.md.)
__device__ __forceinline__ void red_add_f64(double *addr, double v)
{
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

#define LPF_MAXP 8

struct __align__(16) LpfBasisTab {
    // {B, G} interleaved per (q, d): one 16-byte uniform load (LDCU.128) feeds both FMAs that every stage
    // issues on the same (q, d) pair -- halves the coefficient loads of the apply kernels
    double BG[2 * (LPF_MAXP + 2) * (LPF_MAXP + 1)];
    double B[(LPF_MAXP + 2) * (LPF_MAXP + 1)];     // [Q][D]
    double G[(LPF_MAXP + 2) * (LPF_MAXP + 1)];
    double Dhat[(LPF_MAXP + 1) * (LPF_MAXP + 1)];  // [D][D]
    double nodes[LPF_MAXP + 1];
    double qpts[LPF_MAXP + 2];
    double qwts[LPF_MAXP + 2];
};

__constant__ LpfBasisTab c_tab[LPF_MAXP + 1];

__host__ __device__ constexpr int lpf_pad_to(int v, int target_mod16)
{
    // smallest w >= v with w % 16 == target_mod16 % 16
    int w = v;
    while ((w & 15) != (target_mod16 & 15)) w++;
    return w;
}

// Affine fast path set-up: one thread per element.  qa[e][6] = (D11, D21, D31, D22, D32, D33) of detJ J^-1 J^-T (no
// quadrature weight) from the constant Jacobian of an affine hex; *not_affine is raised if any element's trilinear map
// has a bilinear / trilinear part larger than 1e-13 of its edge length (then the fast path stays off).
__global__ void pa_affine_setup_kernel(int ne, const double *__restrict__ corners, double *__restrict__ qa, int *not_affine)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    const double *C = corners + (size_t)e * 24;
    double J[3][3], dev = 0.0, len = 0.0;
    for (int c = 0; c < 3; c++) {
        const double c0 = C[c], c1 = C[3 + c], c2 = C[6 + c], c3 = C[9 + c], c4 = C[12 + c], c5 = C[15 + c], c6 = C[18 + c], c7 = C[21 + c];
        J[c][0] = c1 - c0; J[c][1] = c2 - c0; J[c][2] = c4 - c0;
        dev = fmax(dev, fmax(fmax(fabs(c3 - (c1 + c2 - c0)), fabs(c5 - (c1 + c4 - c0))),
                             fmax(fabs(c6 - (c2 + c4 - c0)), fabs(c7 - (c1 + c2 + c4 - 2.0 * c0)))));
        len = fmax(len, fmax(fabs(J[c][0]), fmax(fabs(J[c][1]), fabs(J[c][2]))));
    }
    if (dev > 1e-13 * len) atomicOr(not_affine, 1);
    const double A11 = J[1][1] * J[2][2] - J[1][2] * J[2][1], A12 = J[2][1] * J[0][2] - J[0][1] * J[2][2], A13 = J[0][1] * J[1][2] - J[1][1] * J[0][2];
    const double A21 = J[2][0] * J[1][2] - J[1][0] * J[2][2], A22 = J[0][0] * J[2][2] - J[0][2] * J[2][0], A23 = J[1][0] * J[0][2] - J[0][0] * J[1][2];
    const double A31 = J[1][0] * J[2][1] - J[2][0] * J[1][1], A32 = J[2][0] * J[0][1] - J[0][0] * J[2][1], A33 = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double det = J[0][0] * A11 + J[0][1] * A21 + J[0][2] * A31;
    const double s = 1.0 / det;
    double *q = qa + (size_t)e * 6;
    q[0] = s * (A11 * A11 + A12 * A12 + A13 * A13);
    q[1] = s * (A11 * A21 + A12 * A22 + A13 * A23);
    q[2] = s * (A11 * A31 + A12 * A32 + A13 * A33);
    q[3] = s * (A21 * A21 + A22 * A22 + A23 * A23);
    q[4] = s * (A21 * A31 + A22 * A32 + A23 * A33);
    q[5] = s * (A31 * A31 + A32 * A32 + A33 * A33);
}

// Strides of the two stage buffers, per order, picked by the bank-conflict model tools/smem_layout_sim.py (shared-memory
// wavefronts per element, old formula-padded layout -> this table):  p1 48 -> 32, p2 90 -> 63, p3 230 -> 167, p4 319 -> 322
// (smaller: 4 CTAs per SM), p5 560 -> 496, p6 680 -> 642, p7 1006 -> 862, p8 1394 -> 1394 (smaller).  Measured at order 3:
// 77 % -> 89 % of the HBM roofline from the layout alone.
//   {SAY, SAZ, SBZ, PAD}: A = [arr 2][dz][dy][qx] with strides SAZ, SAY, 1; B = [arr 3][dz][qy][qx] with strides SBZ, Q, 1
__host__ __device__ constexpr int lpf_smem_stride(int p, int which)
{
    constexpr int T[9][4] = {{0, 0, 0, 0}, {4, 8, 12, 1}, {5, 20, 20, 0}, {5, 28, 26, 1}, {7, 35, 38, 12},
                             {7, 42, 55, 4}, {9, 72, 72, 0}, {10, 89, 89, 0}, {10, 90, 106, 0}};
    return T[p][which];
}

template <int P, int E>
struct ApplyCfg {
    static constexpr int D = P + 1, Q = P + 2;
    static constexpr int LX = D * D, LY = D * Q, LZ = Q * Q;
    static constexpr int NT = E * LZ;
    static constexpr int DP3 = (D * D * D + 3) & ~3;    // gather-map row stride (rows padded to 16 bytes for bulk copies)
    static constexpr int SAY = lpf_smem_stride(P, 0);
    static constexpr int SAZ = lpf_smem_stride(P, 1);
    static constexpr int SAA = D * SAZ;
    static constexpr int SBZ = lpf_smem_stride(P, 2);
    static constexpr int SBA = D * SBZ;
    static constexpr int ES = 2 * SAA + 3 * SBA + lpf_smem_stride(P, 3);      // element stride
    static constexpr int OFFB = 2 * SAA;
    static constexpr size_t SMEM_BYTES = (size_t)E * ES * sizeof(double);
    static_assert(SAY >= Q && SAZ >= D * SAY - (SAY - Q) && SBZ >= Q * Q, "stage-buffer strides too small");
};

__device__ __forceinline__ double2 ldg_stream2(const double2 *p)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

// MODE bit 0: accumulate per-CTA x_e . y_e into den_slots (PCG denominator)
// EVEC: x/y are E-vectors (AddMultPA semantics: y_E += ...), gmap unused.
template <int P, int E, bool PREFETCH, bool EVEC, int MINB>
__global__ void __launch_bounds__(ApplyCfg<P, E>::NT, MINB)
pa_apply_kernel(const double *__restrict__ qd, const int *__restrict__ gmap, const double *__restrict__ x,
                double *__restrict__ y, int ne, double *__restrict__ den_slots, const int *__restrict__ status)
{
    using C = ApplyCfg<P, E>;
    constexpr int D = C::D, Q = C::Q, LX = C::LX, LY = C::LY, LZ = C::LZ;
    constexpr int D3 = D * D * D;
    extern __shared__ double smem[];
    if (status != nullptr && *status != 0) return;     // PCG already finished: graph-replayed launch is a no-op
    const LpfBasisTab &T = c_tab[P];
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
    int idx[D];
    double xs[D];
    const int ex = tid / LX, lx = tid - ex * LX;
    const int xdz = lx / D, xdy = lx - xdz * D;
    const bool xvalid = (tid < E * LX) && (e0 + ex) < ne;
    if (xvalid) {
        if (EVEC) {
            const double *src = x + (size_t)(e0 + ex) * D3 + lx * D;
#pragma unroll
            for (int i = 0; i < D; i++) { idx[i] = 0; xs[i] = src[i]; }
        } else {
            const int *gi = gmap + (size_t)(e0 + ex) * C::DP3 + lx * D;
#pragma unroll
            for (int i = 0; i < D; i++) idx[i] = gi[i];
#pragma unroll
            for (int i = 0; i < D; i++) xs[i] = idx[i] >= 0 ? x[idx[i]] : 0.0;
        }
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
            b[q * Q] = s0;
            b[C::SBA + q * Q] = s1;
            b[2 * C::SBA + q * Q] = s2;
        }
    }
    __syncthreads();

    // ---- Z stage: column (qy,qx): forward z, q-data, backward z, all in registers ----
    if (zvalid) {
        double *b = smem + ez * C::ES + C::OFFB + q2;
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
        for (int q = 0; q < Q; q++) { vbb[q] = b[q * Q]; vbg[q] = b[C::SBA + q * Q]; vgb[q] = b[2 * C::SBA + q * Q]; }
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
    double part = 0.0;
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
            if (EVEC) dst[i] += s;
            else if (idx[i] >= 0) { red_add_f64(y + idx[i], s); part = fma(xs[i], s, part); }
        }
    }
    if (den_slots != nullptr) {
        // CTA-level sum of x_e . (A_e x_e): (d, A d) of the PCG without a second pass over the vectors
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

// ------------------------------------------------------------------------------------------------
// q-data setup: one thread per quadrature point.  Geometry either from the trilinear
// corners or from a Jacobian array in MFEM's GeometricFactors layout [Q^3][3][3][ne].
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void trilinear_jac(const double *__restrict__ c, double x, double y, double z, double J[3][3])
{
#pragma unroll
    for (int a = 0; a < 3; a++) { J[a][0] = 0.0; J[a][1] = 0.0; J[a][2] = 0.0; }
#pragma unroll
    for (int n = 0; n < 8; n++) {
        const int cx = n & 1, cy = (n >> 1) & 1, cz = (n >> 2) & 1;
        const double fx = cx ? x : 1.0 - x, fy = cy ? y : 1.0 - y, fz = cz ? z : 1.0 - z;
        const double sx = cx ? 1.0 : -1.0, sy = cy ? 1.0 : -1.0, sz = cz ? 1.0 : -1.0;
        const double d0 = sx * fy * fz, d1 = fx * sy * fz, d2 = fx * fy * sz;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const double ca = c[3 * n + a];
            J[a][0] = fma(ca, d0, J[a][0]);
            J[a][1] = fma(ca, d1, J[a][1]);
            J[a][2] = fma(ca, d2, J[a][2]);
        }
    }
}

__global__ void pa_setup_kernel(int p, int ne, const double *__restrict__ corners, const double *__restrict__ jac,
                                double *__restrict__ qd)
{
    const int Q = p + 2, Q2 = Q * Q, Q3 = Q2 * Q;
    const LpfBasisTab &T = c_tab[p];
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)ne * Q3) return;
    const int e = (int)(gid / Q3), q = (int)(gid - (size_t)e * Q3);
    const int qz = q / Q2, q2 = q - qz * Q2, qy = q2 / Q, qx = q2 - qy * Q;
    double J[3][3];
    if (jac != nullptr) {
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int k = 0; k < 3; k++) J[c][k] = jac[q + (size_t)Q3 * (c + 3 * (k + 3 * (size_t)e))];
    } else {
        trilinear_jac(corners + (size_t)e * 24, T.qpts[qx], T.qpts[qy], T.qpts[qz], J);
    }
    const double J11 = J[0][0], J12 = J[0][1], J13 = J[0][2];
    const double J21 = J[1][0], J22 = J[1][1], J23 = J[1][2];
    const double J31 = J[2][0], J32 = J[2][1], J33 = J[2][2];
    const double det = J11 * (J22 * J33 - J32 * J23) - J21 * (J12 * J33 - J32 * J13) + J31 * (J12 * J23 - J22 * J13);
    const double cw = T.qwts[qx] * T.qwts[qy] * T.qwts[qz] / det;
    const double A11 = (J22 * J33) - (J23 * J32), A12 = (J32 * J13) - (J12 * J33), A13 = (J12 * J23) - (J22 * J13);
    const double A21 = (J31 * J23) - (J21 * J33), A22 = (J11 * J33) - (J13 * J31), A23 = (J21 * J13) - (J11 * J23);
    const double A31 = (J21 * J32) - (J31 * J22), A32 = (J31 * J12) - (J11 * J32), A33 = (J11 * J22) - (J12 * J21);
    double v[6];
    v[0] = cw * (A11 * A11 + A12 * A12 + A13 * A13);
    v[1] = cw * (A11 * A21 + A12 * A22 + A13 * A23);
    v[2] = cw * (A11 * A31 + A12 * A32 + A13 * A33);
    v[3] = cw * (A21 * A21 + A22 * A22 + A23 * A23);
    v[4] = cw * (A21 * A31 + A22 * A32 + A23 * A33);
    v[5] = cw * (A31 * A31 + A32 * A32 + A33 * A33);
    double2 *o = reinterpret_cast<double2 *>(qd) + ((size_t)(e * Q + qz) * 3) * Q2 + q2;
    o[0] = make_double2(v[0], v[1]);
    o[Q2] = make_double2(v[2], v[3]);
    o[2 * Q2] = make_double2(v[4], v[5]);
}

// internal layout -> MFEM pa_data layout [Q^3][6][ne]
__global__ void pa_qdata_export_kernel(int p, int ne, const double *__restrict__ qd, double *__restrict__ out)
{
    const int Q = p + 2, Q2 = Q * Q, Q3 = Q2 * Q;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)ne * Q3) return;
    const int e = (int)(gid / Q3), q = (int)(gid - (size_t)e * Q3);
    const int qz = q / Q2, q2 = q - qz * Q2;
    const double2 *s = reinterpret_cast<const double2 *>(qd) + ((size_t)(e * Q + qz) * 3) * Q2 + q2;
    const double2 a = s[0], b = s[Q2], c = s[2 * Q2];
    double *o = out + (size_t)e * 6 * Q3 + q;
    o[0] = a.x; o[Q3] = a.y; o[2 * Q3] = b.x; o[3 * Q3] = b.y; o[4 * Q3] = c.x; o[5 * Q3] = c.y;
}

// PA diagonal: one thread per element dof, scatter-added into the L-vector.
__global__ void pa_diag_kernel(int p, int ne, const double *__restrict__ qd, const int *__restrict__ gmap,
                               double *__restrict__ diag)
{
    const int D = p + 1, Q = p + 2, Q2 = Q * Q, D3 = D * D * D, DP3 = (D3 + 3) & ~3;
    const LpfBasisTab &T = c_tab[p];
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)ne * D3) return;
    const int e = (int)(gid / D3), d = (int)(gid - (size_t)e * D3);
    const int dz = d / (D * D), dy = (d / D) % D, dx = d % D;
    const double2 *s = reinterpret_cast<const double2 *>(qd) + ((size_t)e * Q * 3) * Q2;
    double acc = 0.0;
    for (int qz = 0; qz < Q; qz++) {
        const double bz = T.B[qz * D + dz], gz = T.G[qz * D + dz];
        for (int qy = 0; qy < Q; qy++) {
            const double by = T.B[qy * D + dy], gy = T.G[qy * D + dy];
            for (int qx = 0; qx < Q; qx++) {
                const double bx = T.B[qx * D + dx], gx = T.G[qx * D + dx];
                const double2 *t = s + (size_t)(qz * 3) * Q2 + qy * Q + qx;
                const double2 a = t[0], b = t[Q2], c = t[2 * Q2];
                const double p0 = gx * by * bz, p1 = bx * gy * bz, p2 = bx * by * gz;
                acc += a.x * p0 * p0 + b.y * p1 * p1 + c.y * p2 * p2
                     + 2.0 * (a.y * p0 * p1 + b.x * p0 * p2 + c.x * p1 * p2);
            }
        }
    }
    const int g = gmap[(size_t)e * DP3 + d];
    atomicAdd(diag + (g >= 0 ? g : ~g), acc);
}
