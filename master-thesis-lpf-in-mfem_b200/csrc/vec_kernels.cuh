// Device-resident Jacobi-PCG, surface and RK4 kernels.
//
// PCG state lives in device memory (PcgState); every kernel of the iteration starts with
// `if (st->status) return;` so a CUDA graph holding a fixed chunk of iterations can be replayed
// without host round trips -- once the stopping rule of CGSolver::Mult fires, the rest of the chunk
// degenerates to empty launches and the iteration count stays exactly MFEM's.
#pragma once
#include "dev_util.cuh"
#include "p2p_types.hpp"

__device__ __forceinline__ double block_sum(double v)
{
    __shared__ double sh[33];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double s = 0.0;
    if (w == 0) {
        s = l < nw ? sh[l] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (l == 0) sh[32] = s;
    }
    __syncthreads();
    return sh[32];
}

// this thread's share of the (d, A d) slots written by the apply kernel, in a fixed order (reproducible)
__device__ __forceinline__ double den_slot_sum(const double *den_slots)
{
    double s = 0.0;
    for (int i = threadIdx.x; i < LPF_DEN_SLOTS; i += blockDim.x) s += den_slots[i];
    return s;
}

// Block partial -> global array; the last block to arrive sums the array in a fixed order and calls fin(sum).
template <class F>
__device__ __forceinline__ bool grid_sum_finalize(double v, double *partials, unsigned int *counter, F fin)
{
    const double bs = block_sum(v);
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = bs;
        __threadfence();
        const unsigned int t = atomicInc(counter, gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double s = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += ((volatile double *)partials)[i];
        s = block_sum(s);
        if (threadIdx.x == 0) fin(s);
    }
    return is_last;       // true in every thread of the block that arrived last (all other blocks are done)
}

__device__ __forceinline__ void pcg_finalize_nom(PcgState *st, double nom)
{
    st->nom0 = nom; st->nom = nom; st->betanom = nom;
    if (nom < 0.0) { st->status = PCG_NOT_PD; st->final_iter = 0; return; }
    const double r0 = fmax(nom * st->rel2, st->abs2);
    st->r0 = r0;
    if (nom <= r0) { st->status = PCG_CONVERGED; st->final_iter = 0; }
}

__device__ __forceinline__ void pcg_finalize_beta(PcgState *st, double betanom)
{
    st->betanom = betanom;
    if (betanom < 0.0) { st->status = PCG_NOT_PD; st->final_iter = st->iter; return; }
    if (betanom <= st->r0) { st->status = PCG_CONVERGED; st->final_iter = st->iter; return; }
    const int i = st->iter + 1;
    st->iter = i;
    if (i > st->max_iter) { st->status = PCG_MAXITER; st->final_iter = st->max_iter; return; }
    st->beta = betanom / st->nom;
    st->nom = betanom;
}

__global__ void pcg_reset_kernel(PcgState *st, double *den_slots, double rel_tol, double abs_tol, int max_iter)
{
    const int t = threadIdx.x;
    for (int i = t; i < LPF_DEN_SLOTS; i += blockDim.x) den_slots[i] = 0.0;
    if (t == 0) {
        st->comm_error = 0;
        st->nom = st->den = st->betanom = st->r0 = st->nom0 = st->beta = 0.0;
        st->rel2 = rel_tol * rel_tol; st->abs2 = abs_tol * abs_tol;
        st->iter = 1; st->status = PCG_RUNNING; st->max_iter = max_iter; st->final_iter = 0; st->counter = 0;
    }
}

// r = b - t (t = A_c x or nothing when t == nullptr), z = dinv r, d = z, nom = (d, r) over owned dofs; ad = 0
template <bool MULTI>
__global__ void pcg_init_kernel(int n, const double *__restrict__ b, const double *__restrict__ t,
                                const double *__restrict__ dinv, const uint8_t *__restrict__ owned,
                                double *__restrict__ r, double *__restrict__ z, double *__restrict__ d,
                                double *__restrict__ ad, PcgState *st, double *partials)
{
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double ri = t ? b[i] - t[i] : b[i];
        const double zi = dinv[i] * ri;
        r[i] = ri; z[i] = zi; d[i] = zi; ad[i] = 0.0;
        if (!owned || owned[i]) acc = fma(zi, ri, acc);
    }
    grid_sum_finalize(acc, partials, &st->counter, [&](double s) {
        if (MULTI) st->red[0] = s; else pcg_finalize_nom(st, s);
    });
}

__global__ void pcg_fin_nom_kernel(PcgState *st) { pcg_finalize_nom(st, st->red[0]); }
__global__ void pcg_fin_beta_kernel(PcgState *st) { if (st->status == PCG_RUNNING) pcg_finalize_beta(st, st->red[0]); }

// multi-GPU: den_slots -> st->red[1] (then all-reduced into st->den by the host-enqueued collective)
__global__ void pcg_den_local_kernel(PcgState *st, double *den_slots)
{
    const double s = block_sum(den_slot_sum(den_slots));
    for (int i = threadIdx.x; i < LPF_DEN_SLOTS; i += blockDim.x) den_slots[i] = 0.0;
    if (threadIdx.x == 0) st->red[1] = s;
}

// alpha = nom/den; x += alpha d; r -= alpha ad; z = dinv r; betanom = (r, z)
template <bool MULTI>
__global__ void pcg_update_kernel(int n, double *__restrict__ x, double *__restrict__ r, double *__restrict__ z,
                                  const double *__restrict__ d, const double *__restrict__ ad,
                                  const double *__restrict__ dinv, const uint8_t *__restrict__ owned, PcgState *st,
                                  const double *__restrict__ den_slots, double *partials)
{
    griddep_wait();
    griddep_launch();      // after the wait: at most ONE successor kernel is resident ahead of time
    if (st->status != PCG_RUNNING) return;
    double den;
    if (MULTI) den = st->red[1];
    else den = block_sum(den_slot_sum(den_slots));
    if (den == 0.0) {                                  // CGSolver: den == 0 -> stop, final_iter = i
        if (blockIdx.x == 0 && threadIdx.x == 0) { st->den = den; st->status = PCG_BREAKDOWN; st->final_iter = st->iter > 1 ? st->iter : 0; }
        return;
    }
    const double alpha = st->nom / den;
    double acc = 0.0;
    // x, r, dinv are touched once per iteration: streaming (evict-first) accesses, so that z, d and A d -- handed from kernel
    // to kernel inside the iteration and covered by the persisting-L2 window (pcg_l2_window) -- stay in L2
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        __stcs(x + i, fma(alpha, d[i], __ldcs(x + i)));
        const double ri = fma(-alpha, ad[i], __ldcs(r + i));
        const double zi = __ldcs(dinv + i) * ri;
        __stcs(r + i, ri); z[i] = zi;
        if (!owned || owned[i]) acc = fma(ri, zi, acc);
    }
    grid_sum_finalize(acc, partials, &st->counter, [&](double s) {
        st->den = den;
        if (MULTI) st->red[0] = s; else pcg_finalize_beta(st, s);
    });
}

// d = z + beta d; ad = 0; den slots cleared for the next apply
__global__ void pcg_dir_kernel(int n, const double *__restrict__ z, double *__restrict__ d, double *__restrict__ ad,
                               const PcgState *st, double *den_slots)
{
    griddep_wait();
    griddep_launch();      // after the wait: at most ONE successor kernel is resident ahead of time
    if (st->status != PCG_RUNNING) return;
    const double beta = st->beta;
    if (blockIdx.x == 0) for (int i = threadIdx.x; i < LPF_DEN_SLOTS; i += blockDim.x) den_slots[i] = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        d[i] = fma(beta, d[i], z[i]);
        ad[i] = 0.0;
    }
}

// General right-hand sides (lpf_pcg): the search direction can be non-zero on essential dofs, where the
// constrained operator is the identity: ad[ess] = d[ess] and (d, A_c d) gains sum d[ess]^2.  (The Laplace
// solve of the RHS never needs this: there r[ess] == 0 exactly, so d[ess] == 0 for all iterations.)
__global__ void pcg_ess_fix_kernel(int ness, const int *__restrict__ ess, const double *__restrict__ d,
                                   double *__restrict__ ad, const uint8_t *__restrict__ owned,
                                   double *__restrict__ den_slots, const PcgState *st)
{
    if (st->status != PCG_RUNNING) return;
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ness; i += gridDim.x * blockDim.x) {
        const int e = ess[i];
        const double v = d[e];
        ad[e] = v;
        if (!owned || owned[e]) acc = fma(v, v, acc);
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) atomicAdd(den_slots + (blockIdx.x & (LPF_DEN_SLOTS - 1)), acc);
}

// ---- small helpers ------------------------------------------------------------------------------
__global__ void copy_at_kernel(int n, const int *__restrict__ idx, const double *__restrict__ src, double *__restrict__ dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[idx[i]] = src[idx[i]];
}
__global__ void scatter_kernel(int n, const int *__restrict__ idx, const double *__restrict__ src, double *__restrict__ dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[idx[i]] = src[i];
}
__global__ void negate_kernel(int n, const double *__restrict__ t, double *__restrict__ b)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) b[i] = -t[i];
}
__global__ void dinv_kernel(int n, const double *__restrict__ diag, const uint8_t *__restrict__ essmask, double *__restrict__ dinv, int *bad)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double dg = diag[i];
        if (!(dg > 0.0)) atomicAdd(bad, 1);
        dinv[i] = essmask[i] ? 1.0 : 1.0 / dg;
    }
}

// halo-sum pack / rank-ordered unpack (replaces GroupCommunicator::Reduce + Bcast)
__global__ void halo_pack_kernel(int n, const int *__restrict__ dofs, const double *__restrict__ v, double *__restrict__ buf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) buf[i] = v[dofs[i]];
}
__global__ void halo_unpack_kernel(int n, const int *__restrict__ shared, const int *__restrict__ off,
                                   const int *__restrict__ src, const double *__restrict__ recv, double *__restrict__ v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int dof = shared[i];
    const double own = v[dof];
    double s = 0.0;
    for (int j = off[i]; j < off[i + 1]; j++) s += (src[j] < 0) ? own : recv[src[j]];
    v[dof] = s;
}

// ---- surface kernels ------------------------------------------------------------
// wsum[s] = sum over the elements touching surface dof s of  sum_k Jinv(k,z) dphi_e/dxi_k  at that node.
// One thread per surface dof walks its (element, local node) list in a fixed order: deterministic.
__global__ void surface_dz_kernel(int p, int ns, const int *__restrict__ sd_off, const int *__restrict__ sd_elem,
                                  const int *__restrict__ sd_node, const int *__restrict__ gmap,
                                  const double *__restrict__ corners, const double *__restrict__ jinv_z,
                                  const double *__restrict__ phi, double *__restrict__ wsum)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= ns) return;
    const int D = p + 1, D3 = D * D * D, DP3 = (D3 + 3) & ~3;
    const LpfBasisTab &T = c_tab[p];
    double acc = 0.0;
    for (int j = sd_off[s]; j < sd_off[s + 1]; j++) {
        const int e = sd_elem[j], n = sd_node[j];
        const int k = n / (D * D), jj = (n / D) % D, i = n % D;
        const int *g = gmap + (size_t)e * DP3;
        double g0 = 0.0, g1 = 0.0, g2 = 0.0;
        for (int m = 0; m < D; m++) {
            g0 = fma(T.Dhat[i * D + m], phi[g[m + D * (jj + D * k)]], g0);
            g1 = fma(T.Dhat[jj * D + m], phi[g[i + D * (m + D * k)]], g1);
            g2 = fma(T.Dhat[k * D + m], phi[g[i + D * (jj + D * m)]], g2);
        }
        double i0, i1, i2;
        if (jinv_z != nullptr) {
            const double *ji = jinv_z + ((size_t)e * D3 + n) * 3;
            i0 = ji[0]; i1 = ji[1]; i2 = ji[2];
        } else {
            double J[3][3];
            trilinear_jac(corners + (size_t)e * 24, T.nodes[i], T.nodes[jj], T.nodes[k], J);
            const double det = J[0][0] * (J[1][1] * J[2][2] - J[2][1] * J[1][2]) - J[1][0] * (J[0][1] * J[2][2] - J[2][1] * J[0][2])
                             + J[2][0] * (J[0][1] * J[1][2] - J[1][1] * J[0][2]);
            i0 = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
            i1 = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
            i2 = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
        }
        acc += i0 * g0 + i1 * g1 + i2 * g2;
    }
    wsum[s] = acc;
}

__global__ void surface_div_kernel(int ns, const int *__restrict__ mult, const double *__restrict__ wsum, double *__restrict__ wt)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < ns) wt[s] = wsum[s] / (double)mult[s];
}

struct RhsDev {
    double g, H, omega, k, kx, ky, cwave, coth_kh, T, inv_tau, n_ramp;
    int relax;
};

// dstate = [w~ ; -g eta] (+ relaxation-zone forcing)   Solvers/PF_linear_par_partial.cpp:169-239
__global__ void surface_rhs_kernel(int ns, RhsDev prm, double t, const int *__restrict__ mult,
                                   const double *__restrict__ wsum, const double *__restrict__ state,
                                   const double *__restrict__ xy, const double *__restrict__ cgen,
                                   const double *__restrict__ cabs, const double *__restrict__ cabsy,
                                   double *__restrict__ dstate)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= ns) return;
    const double eta = state[s], pfs = state[ns + s];
    double deta = wsum[s] / (double)mult[s];
    double dphi = -prm.g * eta;
    if (prm.relax) {
        const double phase = prm.omega * t - prm.k * (prm.kx * xy[2 * s] + prm.ky * xy[2 * s + 1]);
        const double eta_e = 0.5 * prm.H * cos(phase);
        const double phi_e = -0.5 * prm.H * prm.cwave * prm.coth_kh * sin(phase);
        double alpha = t / (prm.n_ramp * prm.T);
        alpha = fmin(1.0, fmax(0.0, alpha));
        const double gw = alpha * cgen[s];
        deta += (gw * prm.inv_tau) * (eta_e - eta);
        dphi += (gw * prm.inv_tau) * (phi_e - pfs);
        deta += (cabs[s] * prm.inv_tau) * (0.0 - eta);
        dphi += (cabs[s] * prm.inv_tau) * (0.0 - pfs);
        if (cabsy != nullptr) {       // third weight of the cylinder driver (cylinder-diffraction.cpp:199-210), same order of adds
            deta += (cabsy[s] * prm.inv_tau) * (0.0 - eta);
            dphi += (cabsy[s] * prm.inv_tau) * (0.0 - pfs);
        }
    }
    dstate[s] = deta;
    dstate[ns + s] = dphi;
}

// eta envelope: env = max(env, eta) over the steps of the last period (cylinder-diffraction.cpp:410-432)
__global__ void envelope_kernel(int ns, const double *__restrict__ state, double *__restrict__ env)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < ns) env[s] = fmax(env[s], state[s]);
}
__global__ void fill_kernel(int n, double v, double *__restrict__ dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = v;
}

// RK4Solver::Step vector updates ([MFEM] linalg/ode.cpp), stage = 0..3
__global__ void rk4_stage_kernel(int n, int stage, double dt, double *__restrict__ x, const double *__restrict__ k,
                                 double *__restrict__ y, double *__restrict__ z)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double xi = x[i], ki = k[i];
    if (stage == 0) { y[i] = xi + (dt / 2) * ki; z[i] = xi + (dt / 6) * ki; }
    else if (stage == 1) { y[i] = xi + (dt / 2) * ki; z[i] += (dt / 3) * ki; }
    else if (stage == 2) { y[i] = xi + dt * ki; z[i] += (dt / 3) * ki; }
    else { x[i] = z[i] + (dt / 6) * ki; }
}
