/* CPU oracle (plain C + OpenMP) for the LPF Laplace hot path -- TEST INFRASTRUCTURE ONLY.
 *
 * Restates, for timing and cross-checking, what the reference executes inside MFEM's CPU
 * partial-assembly path (the reference itself has no native code; every driver is
 * `#include "mfem.hpp"`, e.g. Solvers/PF_linear_par_partial.cpp:1, and MFEM is absent from
 * /root/reference and from this image => PARITY UNPINNED at the MFEM boundary, see
 * oracle/lpf_oracle.py header and DESIGN.md).
 *
 *   lpf_or_setup      DiffusionIntegrator::AssemblePA  (PF_linear_par_partial.cpp:118-121)  SURVEY A.3
 *   lpf_or_gather     ElementRestriction::Mult                                               SURVEY 3.5
 *   lpf_or_apply_E    DiffusionIntegrator::AddMultPA  (sum factorisation)                    SURVEY A.4
 *   lpf_or_scatter    ElementRestriction::MultTranspose (offset/index form, deterministic)   SURVEY 3.5
 *   lpf_or_diag_E     DiffusionIntegrator::AssembleDiagonalPA (PF_linear_par_partial.cpp:124) SURVEY A.5
 *   lpf_or_pcg        FormLinearSystem + CGSolver::Mult with OperatorJacobiSmoother
 *                     (PF_linear_par_partial.cpp:155-164)                                    SURVEY A.6, 3.4
 *   lpf_or_deriv_z    GridFunction::GetDerivative(1,2,w) (PF_linear_par_partial.cpp:169)     SURVEY A.7
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  q-data layout: [e][6][Q^3]; E-vectors: [e][D^3] (x fastest).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXD 11
#define MAXQ 12

void lpf_or_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int lpf_or_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Jacobian of the trilinear map at reference point (x,y,z); corners lexicographic [8][3]. */
static void trilinear_jac(const double *c, double x, double y, double z, double J[3][3])
{
    for (int a = 0; a < 3; a++) { J[a][0] = J[a][1] = J[a][2] = 0.0; }
    for (int n = 0; n < 8; n++) {
        const int cx = n & 1, cy = (n >> 1) & 1, cz = (n >> 2) & 1;
        const double fx = cx ? x : 1.0 - x, fy = cy ? y : 1.0 - y, fz = cz ? z : 1.0 - z;
        const double sx = cx ? 1.0 : -1.0, sy = cy ? 1.0 : -1.0, sz = cz ? 1.0 : -1.0;
        const double d0 = sx * fy * fz, d1 = fx * sy * fz, d2 = fx * fy * sz;
        for (int a = 0; a < 3; a++) {
            J[a][0] += c[3 * n + a] * d0;
            J[a][1] += c[3 * n + a] * d1;
            J[a][2] += c[3 * n + a] * d2;
        }
    }
}

void lpf_or_setup(int ne, int Q, const double *corners, const double *qpts, const double *qwts, double *qd)
{
    const int Q3 = Q * Q * Q;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < ne; e++) {
        double *o = qd + (size_t)e * 6 * Q3;
        for (int qz = 0; qz < Q; qz++)
            for (int qy = 0; qy < Q; qy++)
                for (int qx = 0; qx < Q; qx++) {
                    double J[3][3];
                    trilinear_jac(corners + (size_t)e * 24, qpts[qx], qpts[qy], qpts[qz], J);
                    const double J11 = J[0][0], J12 = J[0][1], J13 = J[0][2];
                    const double J21 = J[1][0], J22 = J[1][1], J23 = J[1][2];
                    const double J31 = J[2][0], J32 = J[2][1], J33 = J[2][2];
                    const double det = J11 * (J22 * J33 - J32 * J23) - J21 * (J12 * J33 - J32 * J13)
                                     + J31 * (J12 * J23 - J22 * J13);
                    const double cw = qwts[qx] * qwts[qy] * qwts[qz] / det;
                    const double A11 = J22 * J33 - J23 * J32, A12 = J32 * J13 - J12 * J33, A13 = J12 * J23 - J22 * J13;
                    const double A21 = J31 * J23 - J21 * J33, A22 = J11 * J33 - J13 * J31, A23 = J21 * J13 - J11 * J23;
                    const double A31 = J21 * J32 - J31 * J22, A32 = J31 * J12 - J11 * J32, A33 = J11 * J22 - J12 * J21;
                    const int q = qx + Q * (qy + Q * qz);
                    o[0 * Q3 + q] = cw * (A11 * A11 + A12 * A12 + A13 * A13);
                    o[1 * Q3 + q] = cw * (A11 * A21 + A12 * A22 + A13 * A23);
                    o[2 * Q3 + q] = cw * (A11 * A31 + A12 * A32 + A13 * A33);
                    o[3 * Q3 + q] = cw * (A21 * A21 + A22 * A22 + A23 * A23);
                    o[4 * Q3 + q] = cw * (A21 * A31 + A22 * A32 + A23 * A33);
                    o[5 * Q3 + q] = cw * (A31 * A31 + A32 * A32 + A33 * A33);
                }
    }
}

void lpf_or_gather(int ne, int D3, const int *gather, const double *x, double *xE)
{
    const size_t n = (size_t)ne * D3;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) xE[i] = x[gather[i]];
}

/* y[i] = sum over (offsets[i] .. offsets[i+1]) of yE[indices[.]] */
void lpf_or_scatter(int ndof, const int *offsets, const int *indices, const double *yE, double *y)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < ndof; i++) {
        double s = 0.0;
        for (int j = offsets[i]; j < offsets[i + 1]; j++) s += yE[indices[j]];
        y[i] = s;
    }
}

/* One element of the sum-factorised action; D, Q compile-time constants after inlining. */
static inline __attribute__((always_inline)) void
apply_element(const int D, const int Q, const double *restrict B, const double *restrict G,
              const double *restrict qd, const double *restrict u, double *restrict y)
{
    const int Q3 = Q * Q * Q;
    /* D, Q are literal constants at every call site, so these are fixed-size arrays after inlining */
    double bx[D][D][Q], gx[D][D][Q];
    double bb[D][Q][Q], gb[D][Q][Q], bg[D][Q][Q];
    double f0[Q][Q][Q], f1[Q][Q][Q], f2[Q][Q][Q];
    /* x contraction */
    for (int dz = 0; dz < D; dz++)
        for (int dy = 0; dy < D; dy++)
            for (int qx = 0; qx < Q; qx++) {
                double sb = 0.0, sg = 0.0;
                for (int dx = 0; dx < D; dx++) {
                    const double v = u[dx + D * (dy + D * dz)];
                    sb += B[qx * D + dx] * v;
                    sg += G[qx * D + dx] * v;
                }
                bx[dz][dy][qx] = sb; gx[dz][dy][qx] = sg;
            }
    /* y contraction */
    for (int dz = 0; dz < D; dz++)
        for (int qy = 0; qy < Q; qy++)
            for (int qx = 0; qx < Q; qx++) {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0;
                for (int dy = 0; dy < D; dy++) {
                    s0 += B[qy * D + dy] * bx[dz][dy][qx];
                    s1 += B[qy * D + dy] * gx[dz][dy][qx];
                    s2 += G[qy * D + dy] * bx[dz][dy][qx];
                }
                bb[dz][qy][qx] = s0; gb[dz][qy][qx] = s1; bg[dz][qy][qx] = s2;
            }
    /* z contraction + q-point scaling */
    for (int qz = 0; qz < Q; qz++)
        for (int qy = 0; qy < Q; qy++)
            for (int qx = 0; qx < Q; qx++) {
                double g0 = 0.0, g1 = 0.0, g2 = 0.0;
                for (int dz = 0; dz < D; dz++) {
                    g0 += B[qz * D + dz] * gb[dz][qy][qx];
                    g1 += B[qz * D + dz] * bg[dz][qy][qx];
                    g2 += G[qz * D + dz] * bb[dz][qy][qx];
                }
                const int q = qx + Q * (qy + Q * qz);
                const double O11 = qd[q], O12 = qd[Q3 + q], O13 = qd[2 * Q3 + q];
                const double O22 = qd[3 * Q3 + q], O23 = qd[4 * Q3 + q], O33 = qd[5 * Q3 + q];
                f0[qz][qy][qx] = O11 * g0 + O12 * g1 + O13 * g2;
                f1[qz][qy][qx] = O12 * g0 + O22 * g1 + O23 * g2;
                f2[qz][qy][qx] = O13 * g0 + O23 * g1 + O33 * g2;
            }
    /* transposed z */
    for (int dz = 0; dz < D; dz++)
        for (int qy = 0; qy < Q; qy++)
            for (int qx = 0; qx < Q; qx++) {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0;
                for (int qz = 0; qz < Q; qz++) {
                    s0 += B[qz * D + dz] * f0[qz][qy][qx];
                    s1 += B[qz * D + dz] * f1[qz][qy][qx];
                    s2 += G[qz * D + dz] * f2[qz][qy][qx];
                }
                gb[dz][qy][qx] = s0; bg[dz][qy][qx] = s1; bb[dz][qy][qx] = s2;
            }
    /* transposed y */
    for (int dz = 0; dz < D; dz++)
        for (int dy = 0; dy < D; dy++)
            for (int qx = 0; qx < Q; qx++) {
                double sa = 0.0, sb = 0.0;
                for (int qy = 0; qy < Q; qy++) {
                    sa += B[qy * D + dy] * bb[dz][qy][qx] + G[qy * D + dy] * bg[dz][qy][qx];
                    sb += B[qy * D + dy] * gb[dz][qy][qx];
                }
                bx[dz][dy][qx] = sa; gx[dz][dy][qx] = sb;
            }
    /* transposed x */
    for (int dz = 0; dz < D; dz++)
        for (int dy = 0; dy < D; dy++)
            for (int dx = 0; dx < D; dx++) {
                double s = 0.0;
                for (int qx = 0; qx < Q; qx++)
                    s += B[qx * D + dx] * bx[dz][dy][qx] + G[qx * D + dx] * gx[dz][dy][qx];
                y[dx + D * (dy + D * dz)] += s;
            }
}

#define APPLY_CASE(P)                                                                         \
    case P: {                                                                                 \
        _Pragma("omp parallel for schedule(static)")                                          \
        for (int e = 0; e < ne; e++)                                                          \
            apply_element(P + 1, P + 2, B, G, qd + (size_t)e * 6 * (P + 2) * (P + 2) * (P + 2), \
                          xE + (size_t)e * (P + 1) * (P + 1) * (P + 1),                       \
                          yE + (size_t)e * (P + 1) * (P + 1) * (P + 1));                      \
    } break;

/* yE += A_E xE  (AddMultPA semantics).  B, G: [Q][D] row-major. */
int lpf_or_apply_E(int ne, int p, const double *B, const double *G, const double *qd,
                   const double *xE, double *yE)
{
    switch (p) {
        APPLY_CASE(1) APPLY_CASE(2) APPLY_CASE(3) APPLY_CASE(4)
        APPLY_CASE(5) APPLY_CASE(6) APPLY_CASE(7) APPLY_CASE(8) APPLY_CASE(9) APPLY_CASE(10)
        default: return -1;
    }
    return 0;
}

void lpf_or_diag_E(int ne, int p, const double *B, const double *G, const double *qd, double *dE)
{
    const int D = p + 1, Q = p + 2, Q3 = Q * Q * Q, D3 = D * D * D;
    static const int sym[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
#pragma omp parallel for schedule(static)
    for (int e = 0; e < ne; e++) {
        const double *o = qd + (size_t)e * 6 * Q3;
        for (int dz = 0; dz < D; dz++)
            for (int dy = 0; dy < D; dy++)
                for (int dx = 0; dx < D; dx++) {
                    double s = 0.0;
                    for (int qz = 0; qz < Q; qz++)
                        for (int qy = 0; qy < Q; qy++)
                            for (int qx = 0; qx < Q; qx++) {
                                const int q = qx + Q * (qy + Q * qz);
                                const double psi[3] = {
                                    G[qx * D + dx] * B[qy * D + dy] * B[qz * D + dz],
                                    B[qx * D + dx] * G[qy * D + dy] * B[qz * D + dz],
                                    B[qx * D + dx] * B[qy * D + dy] * G[qz * D + dz]};
                                for (int a = 0; a < 3; a++)
                                    for (int b = 0; b < 3; b++) s += o[sym[a][b] * Q3 + q] * psi[a] * psi[b];
                            }
                    dE[(size_t)e * D3 + dx + D * (dy + D * dz)] = s;
                }
    }
}

/* ---- constrained operator and PCG on L(=T)-vectors, serial rank ---- */
typedef struct {
    int ne, p, ndof, ness;
    const double *B, *G, *qd;
    const int *gather, *offsets, *indices, *ess;
    double *xE, *yE, *tmp;
    long applies;
} lpf_or_op;

static void op_mult_raw(lpf_or_op *op, const double *x, double *y)
{
    const int D3 = (op->p + 1) * (op->p + 1) * (op->p + 1);
    lpf_or_gather(op->ne, D3, op->gather, x, op->xE);
    memset(op->yE, 0, sizeof(double) * (size_t)op->ne * D3);
    lpf_or_apply_E(op->ne, op->p, op->B, op->G, op->qd, op->xE, op->yE);
    lpf_or_scatter(op->ndof, op->offsets, op->indices, op->yE, y);
    op->applies++;
}

/* ConstrainedOperator::Mult (DIAG_ONE) */
static void op_mult_constrained(lpf_or_op *op, const double *x, double *y)
{
    memcpy(op->tmp, x, sizeof(double) * op->ndof);
    for (int i = 0; i < op->ness; i++) op->tmp[op->ess[i]] = 0.0;
    op_mult_raw(op, op->tmp, y);
    for (int i = 0; i < op->ness; i++) y[op->ess[i]] = x[op->ess[i]];
}

static double dot(int n, const double *a, const double *b)
{
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}

/* y = A x  (unconstrained P^T A P on one rank) */
void lpf_or_mult(int ne, int p, int ndof, const double *B, const double *G, const double *qd,
                 const int *gather, const int *offsets, const int *indices, const double *x, double *y)
{
    const int D3 = (p + 1) * (p + 1) * (p + 1);
    lpf_or_op op = {ne, p, ndof, 0, B, G, qd, gather, offsets, indices, NULL, NULL, NULL, NULL, 0};
    op.xE = (double *)malloc(sizeof(double) * (size_t)ne * D3);
    op.yE = (double *)malloc(sizeof(double) * (size_t)ne * D3);
    op_mult_raw(&op, x, y);
    free(op.xE); free(op.yE);
}

/* reps applies with the E-vector work buffers allocated ONCE (what MFEM's operator does): the timing entry of the CPU
 * baseline -- lpf_or_mult() above pays two 32 MB malloc + first-touch page faults per call, which is not the algorithm */
void lpf_or_mult_n(int ne, int p, int ndof, const double *B, const double *G, const double *qd,
                   const int *gather, const int *offsets, const int *indices, const double *x, double *y, int reps)
{
    const int D3 = (p + 1) * (p + 1) * (p + 1);
    lpf_or_op op = {ne, p, ndof, 0, B, G, qd, gather, offsets, indices, NULL, NULL, NULL, NULL, 0};
    op.xE = (double *)malloc(sizeof(double) * (size_t)ne * D3);
    op.yE = (double *)malloc(sizeof(double) * (size_t)ne * D3);
    for (int r = 0; r < reps; r++) op_mult_raw(&op, x, y);
    free(op.xE); free(op.yE);
}

/* FormLinearSystem(ess, x, b=0) + Jacobi-PCG.  On entry x holds the essential values on `ess`
 * (other entries ignored: the interior of the initial guess is zeroed, copy_interior = 0);
 * on exit x is the solution.  info[0]=iterations, info[1]=converged, info[2]=final (r,Mr)^1/2,
 * info[3]=initial, info[4]=#operator applies (incl. EliminateRHS). */
int lpf_or_pcg(int ne, int p, int ndof, const double *B, const double *G, const double *qd,
               const int *gather, const int *offsets, const int *indices, int ness, const int *ess,
               const double *dinv, double *x, double rel_tol, double abs_tol, int max_iter, double *info)
{
    const int D3 = (p + 1) * (p + 1) * (p + 1);
    lpf_or_op op = {ne, p, ndof, ness, B, G, qd, gather, offsets, indices, ess, NULL, NULL, NULL, 0};
    op.xE = (double *)malloc(sizeof(double) * (size_t)ne * D3);
    op.yE = (double *)malloc(sizeof(double) * (size_t)ne * D3);
    op.tmp = (double *)malloc(sizeof(double) * ndof);
    double *b = (double *)calloc(ndof, sizeof(double));
    double *r = (double *)malloc(sizeof(double) * ndof);
    double *z = (double *)malloc(sizeof(double) * ndof);
    double *d = (double *)malloc(sizeof(double) * ndof);
    double *X = (double *)calloc(ndof, sizeof(double));
    for (int i = 0; i < ness; i++) X[ess[i]] = x[ess[i]];
    /* EliminateRHS: b -= A w, w = X restricted to ess; b[ess] = X[ess] */
    op_mult_raw(&op, X, z);
    for (int i = 0; i < ndof; i++) b[i] -= z[i];
    for (int i = 0; i < ness; i++) b[ess[i]] = X[ess[i]];
    /* CGSolver::Mult, iterative_mode = true */
    int final_iter = 0, converged = 0;
    double nom0, nom, betanom = 0.0, den, r0;
    op_mult_constrained(&op, X, r);
    for (int i = 0; i < ndof; i++) { r[i] = b[i] - r[i]; z[i] = dinv[i] * r[i]; d[i] = z[i]; }
    nom0 = nom = dot(ndof, d, r);
    betanom = nom;
    if (nom < 0.0) goto done;
    r0 = fmax(nom * rel_tol * rel_tol, abs_tol * abs_tol);
    if (nom <= r0) { converged = 1; goto done; }
    op_mult_constrained(&op, d, z);
    den = dot(ndof, z, d);
    if (den <= 0.0 && den == 0.0) goto done;
    final_iter = max_iter;
    for (int i = 1;;) {
        const double alpha = nom / den;
#pragma omp parallel for schedule(static)
        for (int j = 0; j < ndof; j++) { X[j] += alpha * d[j]; r[j] -= alpha * z[j]; z[j] = dinv[j] * r[j]; }
        betanom = dot(ndof, r, z);
        if (betanom < 0.0) { final_iter = i; break; }
        if (betanom <= r0) { converged = 1; final_iter = i; break; }
        if (++i > max_iter) break;
        const double beta = betanom / nom;
#pragma omp parallel for schedule(static)
        for (int j = 0; j < ndof; j++) d[j] = z[j] + beta * d[j];
        op_mult_constrained(&op, d, z);
        den = dot(ndof, d, z);
        if (den <= 0.0 && den == 0.0) { final_iter = i; break; }
        nom = betanom;
    }
done:
    memcpy(x, X, sizeof(double) * ndof);
    info[0] = final_iter; info[1] = converged; info[2] = sqrt(fmax(betanom, 0.0));
    info[3] = sqrt(fmax(nom0, 0.0)); info[4] = (double)op.applies;
    free(op.xE); free(op.yE); free(op.tmp); free(b); free(r); free(z); free(d); free(X);
    return 0;
}

/* GetDerivative(1,2,w) over the listed elements: accumulates a = sum_k Jinv(k,z) dphi/dxi_k at every
 * node into w and counts into cnt (caller divides).  Dhat: [D][D] collocation derivative. */
void lpf_or_deriv_z(int nel, const int *elems, int p, const double *nodes1d, const double *Dhat,
                    const double *corners, const int *gather, const double *phi, double *w, double *cnt)
{
    const int D = p + 1, D3 = D * D * D;
    for (int ie = 0; ie < nel; ie++) {
        const int e = elems ? elems[ie] : ie;
        const int *g = gather + (size_t)e * D3;
        for (int k = 0; k < D; k++)
            for (int j = 0; j < D; j++)
                for (int i = 0; i < D; i++) {
                    double g0 = 0.0, g1 = 0.0, g2 = 0.0;
                    for (int m = 0; m < D; m++) {
                        g0 += Dhat[i * D + m] * phi[g[m + D * (j + D * k)]];
                        g1 += Dhat[j * D + m] * phi[g[i + D * (m + D * k)]];
                        g2 += Dhat[k * D + m] * phi[g[i + D * (j + D * m)]];
                    }
                    double J[3][3];
                    trilinear_jac(corners + (size_t)e * 24, nodes1d[i], nodes1d[j], nodes1d[k], J);
                    const double det = J[0][0] * (J[1][1] * J[2][2] - J[2][1] * J[1][2])
                                     - J[1][0] * (J[0][1] * J[2][2] - J[2][1] * J[0][2])
                                     + J[2][0] * (J[0][1] * J[1][2] - J[1][1] * J[0][2]);
                    /* third column of J^{-1}: Jinv(k, z) = cofactor(z, k) / det */
                    const double i0 = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
                    const double i1 = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
                    const double i2 = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
                    const int n = g[i + D * (j + D * k)];
                    w[n] += i0 * g0 + i1 * g1 + i2 * g2;
                    cnt[n] += 1.0;
                }
    }
}
