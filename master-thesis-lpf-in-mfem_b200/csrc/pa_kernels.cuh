// sm_100a kernels for the partial-assembly DiffusionIntegrator on H1 hexes.
//
//   pa_setup_kernel   a1  DiffusionIntegrator::AssemblePA  (Solvers/PF_linear_par_partial.cpp:118-121)
//   (apply kernels: pa_apply_tma.cuh / pa_apply_eo.cuh / pa_apply_evec.cuh, one translation unit per order)
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
#include "dev_util.cuh"

// 1-D tables of one order for the generic (order as a run-time argument) kernels of this file and of vec_kernels.cuh.
// The apply kernels keep their own per-order tables (apply_cfg.cuh).
struct __align__(16) LpfBasisTab {
    double B[(LPF_MAXP + 2) * (LPF_MAXP + 1)];     // [Q][D]
    double G[(LPF_MAXP + 2) * (LPF_MAXP + 1)];
    double Dhat[(LPF_MAXP + 1) * (LPF_MAXP + 1)];  // [D][D]
    double nodes[LPF_MAXP + 1];
    double qpts[LPF_MAXP + 2];
    double qwts[LPF_MAXP + 2];
};

static __constant__ LpfBasisTab c_tab[LPF_MAXP + 1];

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

// PA diagonal: one thread per element dof.  gmap != nullptr: scatter-added into the L-vector `diag`
// (a2, restriction-transpose fused); gmap == nullptr: E-vector output diagE[e][D^3] += ... -- the semantics of
// DiffusionIntegrator::AssembleDiagonalPA(Vector &diag) (accumulates), also the first half of the deterministic route.
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
    if (gmap == nullptr) { diag[gid] += acc; return; }
    const int g = gmap[(size_t)e * DP3 + d];
    atomicAdd(diag + (g >= 0 ? g : ~g), acc);
}

// Deterministic restriction-transpose (option "deterministic"): y[i] = sum of the E-vector entries of dof i, taken in
// ascending (element, local node) order -- the order of MFEM's ElementRestriction::MultTranspose on the CPU, which walks
// the same offsets / indices table.  `skip` (may be NULL) marks rows that stay 0 (essential dofs of the constrained map).
__global__ void det_gather_kernel(int n, const int *__restrict__ offsets, const int *__restrict__ indices,
                                  const double *__restrict__ yE, const uint8_t *__restrict__ skip, double *__restrict__ y)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    if (skip == nullptr || !skip[i])
        for (int j = offsets[i]; j < offsets[i + 1]; j++) s += yE[indices[j]];
    y[i] = s;
}
