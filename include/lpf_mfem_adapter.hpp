// lpf_mfem_adapter.hpp -- the thin MFEM-side classes a maintainer of the reference would add to run its
// hot path on liblpf_b200.so without touching the drivers' call sites.
//
// The reference has no FFI: its hot path sits behind MFEM's virtual interfaces.  Each class below derives
// from the MFEM base class the reference uses and forwards to C-ABI calls of include/lpf_b200.h:
//
//   SpaceDescBuilder        (Par)FiniteElementSpace + GroupCommunicator -> lpf_space_desc      :268-285, :407-412
//   B200DiffusionIntegrator : mfem::BilinearFormIntegrator   AssemblePA / AddMultPA / AssembleDiagonalPA
//        replaces `new DiffusionIntegrator`                  Solvers/PF_linear_par_partial.cpp:119, :124 (diagonal)
//   B200LaplaceOperator     : mfem::Operator                 Mult = constrained P^T A P on TRUE dofs   :155 (A_loc)
//   B200JacobiPCG           : mfem::Solver                   Mult = CGSolver + OperatorJacobiSmoother   :124,:157-164
//   B200RhsLinear           : mfem::TimeDependentOperator    Mult = rhs_linear::Mult             :130-244
//   B200RK4Solver           : mfem::ODESolver                Step = RK4Solver::Step              :472-494
//
// Vector sizes are MFEM's: the solver-level classes take T-VECTORS (pfes.GetTrueVSize(): the dofs this rank owns) and
// translate to the library's L-vectors (all local dofs, consistent copies of shared ones) with lpf_prolong / lpf_restrict;
// the integrator takes E-vectors.  With a real MFEM (built with CUDA, `mfem::Device device("cuda")`) Vector::Read() /
// ReadWrite() are device pointers and no copy happens.  In this repository MFEM is not available: the header is compiled
// AND EXECUTED against drivers/stub/mfem.hpp, a minimal functional stand-in with the upstream signatures
// (drivers/adapter_check.cpp, tests/test_gpu_adapter.py); [MFEM] marks calls to re-verify against a real checkout.
#pragma once
#include <algorithm>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "lpf_b200.h"
#include "mfem.hpp"

namespace lpf_mfem {

inline void check(int rc, const char *what)
{
    // MFEM reports errors through MFEM_ABORT / mfem_error; interface methods return void
    if (rc != LPF_OK) mfem::mfem_error((std::string(what) + ": " + lpf_last_error()).c_str());
}

/// Neighbour exchange plan of one rank in the library's form (lpf_space_desc: nbr_rank, nbr_offset, send_dofs, shared_dofs,
/// red_off, red_src, owned), translated from MFEM's GroupCommunicator.
struct HaloArrays {
    std::vector<int> nbr_rank, nbr_offset, send, shared, red_off, red_src;
    std::vector<uint8_t> owned;
};

/// GroupCommunicator (groups of ranks sharing L-dofs, ParFiniteElementSpace::GroupComm(), PF_linear_par_partial.cpp:277)
/// -> halo plan.  `gid(ldof)` must be a number that is the same on every rank sharing the dof (the global true-dof number,
/// [MFEM] ParFiniteElementSpace::GetGlobalTDofNumber): both sides of a pair order their common list by it.  A dof is owned
/// by its group's master rank; partial sums are added in ascending rank order on every sharer (bit-identical copies).
template <class GlobalId>
inline void halo_from_groups(const mfem::GroupCommunicator &gc, int nldof, GlobalId gid, HaloArrays &out)
{
    const mfem::GroupTopology &gt = gc.GetGroupTopology();                 // [MFEM] general/communication.hpp
    const mfem::Table &g2l = gc.GroupLDofTable();                          // [MFEM] group -> L-dofs
    const int me = gt.MyRank();
    out.owned.assign(nldof, 1);
    std::map<int, std::vector<int>> per_nbr;                               // neighbour rank -> shared L-dofs
    std::map<int, std::vector<int>> ranks_of;                              // shared L-dof -> sharing ranks (ascending)
    for (int g = 1; g < gt.NGroups(); g++) {                               // group 0 is the local group
        std::vector<int> ranks;
        for (int i = 0; i < gt.GetGroupSize(g); i++) ranks.push_back(gt.GetNeighborRank(gt.GetGroup(g)[i]));
        std::sort(ranks.begin(), ranks.end());
        const bool mine = gt.GetGroupMasterRank(g) == me;
        for (int j = 0; j < g2l.RowSize(g); j++) {
            const int l = g2l.GetRow(g)[j];
            if (!mine) out.owned[l] = 0;
            ranks_of[l] = ranks;
            for (int r : ranks) if (r != me) per_nbr[r].push_back(l);
        }
    }
    out.nbr_rank.clear(); out.nbr_offset.assign(1, 0); out.send.clear();
    std::map<int, std::map<int, int>> pos;                                  // neighbour -> (L-dof -> index in the receive buffer)
    for (auto &pr : per_nbr) {
        std::vector<int> &l = pr.second;
        std::sort(l.begin(), l.end(), [&](int a, int b) { return gid(a) < gid(b); });
        out.nbr_rank.push_back(pr.first);
        for (size_t i = 0; i < l.size(); i++) { pos[pr.first][l[i]] = out.nbr_offset.back() + (int)i; out.send.push_back(l[i]); }
        out.nbr_offset.push_back((int)out.send.size());
    }
    out.shared.clear(); out.red_off.assign(1, 0); out.red_src.clear();
    for (auto &pr : ranks_of) {                                             // ascending L-dof
        out.shared.push_back(pr.first);
        for (int r : pr.second) out.red_src.push_back(r == me ? -1 : pos[r][pr.first]);
        out.red_off.push_back((int)out.red_src.size());
    }
}

/// Builds the plain descriptor the device context needs from MFEM objects (SURVEY.md 8b last row).
struct SpaceDescBuilder {
    std::vector<double> jac, corners, surf_xy;
    std::vector<int> gather, ess, surf2vol, surf_mult;
    std::vector<uint8_t> surf_owned;
    HaloArrays halo, shalo;
    lpf_space_desc desc{};

    /// Serial space.  ess_ldofs: the essential dofs (serial: ess_tdof_list of :411-412); surf_vdofs: the parent L-dof of every
    /// free-surface dof (what ParSubMesh::Transfer maps through, :147,:170), surf_coords [n_surf][2] their (x, y).
    SpaceDescBuilder(const mfem::FiniteElementSpace &fes, const mfem::Array<int> &ess_ldofs,
                     const mfem::Array<int> &surf_vdofs, const mfem::Vector &surf_coords)
    {
        common(fes, ess_ldofs, surf_vdofs, surf_coords);
        desc.nranks = 1; desc.rank = 0;
        desc.n_true_global = fes.GetTrueVSize();
        desc.n_surf_global = desc.n_surf;
    }

    /// Parallel space: one rank of ParFiniteElementSpace (:277).  ess_ldofs must hold EVERY local essential L-dof, owned or
    /// not ([MFEM] ParFiniteElementSpace::GetEssentialVDofs synchronises the markers across ranks: a rank can hold a copy of
    /// a surface dof without holding a surface face).  The free-surface space's exchange plan comes from ITS communicator
    /// (surf_gc / surf_gid: [MFEM] fespace_fs.GroupComm(), GetGlobalTDofNumber); surf_mult_global[s] = number of elements
    /// touching the dof over ALL ranks (the divisor of GetDerivative's average, :169).
    template <class SurfGid>
    SpaceDescBuilder(const mfem::ParFiniteElementSpace &pfes, const mfem::Array<int> &ess_ldofs,
                     const mfem::Array<int> &surf_vdofs, const mfem::Vector &surf_coords,
                     const mfem::GroupCommunicator &surf_gc, SurfGid surf_gid, const mfem::Array<int> &surf_mult_global,
                     long n_surf_global)
    {
        common(pfes, ess_ldofs, surf_vdofs, surf_coords);
        halo_from_groups(pfes.GroupComm(), desc.ndof, [&](int l) { return pfes.GetGlobalTDofNumber(l); }, halo);
        halo_from_groups(surf_gc, desc.n_surf, surf_gid, shalo);
        surf_mult.assign(surf_mult_global.GetData(), surf_mult_global.GetData() + surf_mult_global.Size());
        surf_owned = shalo.owned;
        desc.nranks = pfes.GetNRanks(); desc.rank = pfes.GetMyRank();
        desc.owned = halo.owned.data(); desc.surf_owned = surf_owned.data(); desc.surf_mult = surf_mult.data();
        desc.n_nbr = (int)halo.nbr_rank.size(); desc.nbr_rank = halo.nbr_rank.data(); desc.nbr_offset = halo.nbr_offset.data();
        desc.send_dofs = halo.send.data(); desc.n_shared = (int)halo.shared.size(); desc.shared_dofs = halo.shared.data();
        desc.red_off = halo.red_off.data(); desc.red_src = halo.red_src.data();
        desc.s_n_nbr = (int)shalo.nbr_rank.size(); desc.s_nbr_rank = shalo.nbr_rank.data(); desc.s_nbr_offset = shalo.nbr_offset.data();
        desc.s_send = shalo.send.data(); desc.s_n_shared = (int)shalo.shared.size(); desc.s_shared = shalo.shared.data();
        desc.s_red_off = shalo.red_off.data(); desc.s_red_src = shalo.red_src.data();
        desc.n_true_global = pfes.GlobalTrueVSize();
        desc.n_surf_global = n_surf_global;
    }

private:
    void common(const mfem::FiniteElementSpace &fes, const mfem::Array<int> &ess_ldofs, const mfem::Array<int> &surf_vdofs,
                const mfem::Vector &surf_coords)
    {
        const mfem::FiniteElement &el = *fes.GetFE(0);
        const int p = el.GetOrder(), D = p + 1, Q = p + 2, D3 = D * D * D, Q3 = Q * Q * Q;
        const int ne = fes.GetNE();
        // [MFEM] the rule DiffusionIntegrator::GetRule picks for tensor elements: order 2p + dim - 1
        const mfem::IntegrationRule &ir = mfem::IntRules.Get(el.GetGeomType(), 2 * p + 2);
        if (ir.GetNPoints() != Q3) mfem::mfem_error("lpf_mfem: unexpected quadrature rule");
        // [MFEM] GeometricFactors::J layout: [Q^3][3][3][ne], q fastest -- exactly lpf_space_desc::jac
        const mfem::GeometricFactors *geom = fes.GetMesh()->GetGeometricFactors(ir, mfem::GeometricFactors::JACOBIANS);
        const double *J = geom->J.HostRead();
        jac.assign(J, J + (size_t)Q3 * 9 * ne);
        // [MFEM] ElementRestriction (LEXICOGRAPHIC) gather map: [D^3][ne], d fastest; H1 hexes have no signs
        const mfem::Operator *R = fes.GetElementRestriction(mfem::ElementDofOrdering::LEXICOGRAPHIC);
        const auto *er = dynamic_cast<const mfem::ElementRestriction *>(R);
        if (!er) mfem::mfem_error("lpf_mfem: expected an ElementRestriction");
        const int *g = er->GatherMap().HostRead();
        gather.assign(g, g + (size_t)D3 * ne);
        ess.assign(ess_ldofs.GetData(), ess_ldofs.GetData() + ess_ldofs.Size());
        std::sort(ess.begin(), ess.end());
        surf2vol.assign(surf_vdofs.GetData(), surf_vdofs.GetData() + surf_vdofs.Size());
        surf_xy.assign(surf_coords.GetData(), surf_coords.GetData() + surf_coords.Size());
        // Geometry of GetDerivative (and of the affine fast path): the 8 corners of every hex, lexicographic, from the mesh
        // transformation -- valid for the trilinear meshes the reference ships (vertex coordinates alone would be wrong on
        // its periodic meshes, whose geometry lives in the L2 nodes).  [MFEM] ElementTransformation::Transform
        corners.resize((size_t)ne * 24);
        {
            mfem::Vector pt(3);
            for (int e = 0; e < ne; e++) {
                mfem::ElementTransformation *T = fes.GetMesh()->GetElementTransformation(e);
                for (int c = 0; c < 8; c++) {
                    mfem::IntegrationPoint ip;
                    ip.Set3(c & 1, (c >> 1) & 1, (c >> 2) & 1);
                    T->Transform(ip, pt);
                    for (int a = 0; a < 3; a++) corners[(size_t)e * 24 + c * 3 + a] = pt(a);
                }
            }
        }
        desc.order = p; desc.ne = ne; desc.ndof = fes.GetVSize();
        // q-data from MFEM's own Jacobians (curved meshes included); corners serve the surface derivative.  Pass
        // desc.jac = nullptr instead to let the library build q-data from the corners and enable its affine fast path.
        desc.corners = corners.data(); desc.jac = jac.data(); desc.gather = gather.data();
        desc.n_ess = (int)ess.size(); desc.ess = ess.data();
        desc.n_surf = (int)surf2vol.size(); desc.surf2vol = surf2vol.data(); desc.surf_xy = surf_xy.data();
    }
};

/// Owns one lpf_ctx; shared by the adapter classes of one rank.
class B200Context {
public:
    explicit B200Context(const lpf_space_desc &d, int device = 0, void *stream = nullptr) : ctx_(lpf_create(&d, device, stream))
    {
        if (!ctx_) mfem::mfem_error((std::string("lpf_create: ") + lpf_last_error()).c_str());
    }
    ~B200Context() { lpf_destroy(ctx_); }
    B200Context(const B200Context &) = delete;
    B200Context &operator=(const B200Context &) = delete;
    lpf_ctx *get() const { return ctx_; }
    /// The library enqueues on its own stream (CUDA graphs cannot be captured on MFEM's default stream).  MFEM's interfaces
    /// are synchronous for the caller -- the next line of the driver may read the result on the host or launch an MFEM
    /// kernel on the default stream -- so every adapter method ends with done().  [MFEM] With a real MFEM whose inputs were
    /// produced by asynchronous device kernels, call MFEM_STREAM_SYNC before the adapter method (the stand-in's copies are
    /// synchronous).
    void done() const { check(lpf_sync(ctx_), "lpf_sync"); }
    /// T-vector -> device L-vector in a work buffer (which = 0 volume, 1 surface; two buffers per space)
    double *to_L(int which, int slot, const double *xT_dev) const
    {
        double *xl = work(which, slot);
        check(lpf_prolong(ctx_, which, xT_dev, xl), "lpf_prolong");
        return xl;
    }
    double *work(int which, int slot) const
    {
        mfem::Vector &w = work_[which][slot];
        const int n = which == 0 ? lpf_ndof(ctx_) : 2 * lpf_nsurf(ctx_);
        if (w.Size() != std::max(n, 1)) w.SetSize(std::max(n, 1));
        return w.Write();
    }

private:
    lpf_ctx *ctx_;
    mutable mfem::Vector work_[2][2];
};

/// Drop-in for `new DiffusionIntegrator` under AssemblyLevel::PARTIAL (PF_linear_par_partial.cpp:119-121).
class B200DiffusionIntegrator : public mfem::BilinearFormIntegrator {
public:
    explicit B200DiffusionIntegrator(B200Context &c) : c_(c) {}
    void AssemblePA(const mfem::FiniteElementSpace &) override { check(lpf_pa_setup(c_.get()), "lpf_pa_setup"); c_.done(); }
    void AddMultPA(const mfem::Vector &x, mfem::Vector &y) const override   // E-vectors, y += A_E x
    {
        check(lpf_pa_apply_E(c_.get(), x.Read(), y.ReadWrite()), "lpf_pa_apply_E");
        c_.done();
    }
    /// E-vector diagonal, accumulated: what OperatorJacobiSmoother(*a_loc_cach, ess_tdof) (:124) reaches through
    /// BilinearForm::AssembleDiagonal -> AssembleDiagonalPA; MFEM then applies G^T and P^T itself.
    void AssembleDiagonalPA(mfem::Vector &diag) override { check(lpf_pa_diag_E(c_.get(), diag.ReadWrite()), "lpf_pa_diag_E"); c_.done(); }

private:
    B200Context &c_;
};

/// The constrained operator FormLinearSystem returns (A_loc, PF_linear_par_partial.cpp:152-155): TRUE-dof sized.
class B200LaplaceOperator : public mfem::Operator {
public:
    explicit B200LaplaceOperator(B200Context &c) : mfem::Operator(lpf_ntrue(c.get(), 0)), c_(c)
    {
        check(lpf_pa_setup(c_.get()), "lpf_pa_setup");
    }
    void Mult(const mfem::Vector &x, mfem::Vector &y) const override
    {
        double *xl = c_.to_L(0, 0, x.Read()), *yl = c_.work(0, 1);
        check(lpf_apply_T(c_.get(), xl, yl), "lpf_apply_T");
        check(lpf_restrict(c_.get(), 0, yl, y.Write()), "lpf_restrict");
        c_.done();
    }
    void AssembleDiagonal(mfem::Vector &diag) const override
    {
        double *dl = c_.work(0, 0);
        check(lpf_diag(c_.get(), dl), "lpf_diag");
        check(lpf_restrict(c_.get(), 0, dl, diag.Write()), "lpf_restrict");
        c_.done();
    }

private:
    B200Context &c_;
};

/// CGSolver + OperatorJacobiSmoother in one object (PF_linear_par_partial.cpp:124,157-164), T-vectors in and out.
class B200JacobiPCG : public mfem::Solver {
public:
    explicit B200JacobiPCG(B200Context &c) : mfem::Solver(lpf_ntrue(c.get(), 0)), c_(c) { iterative_mode = true; }
    void SetOperator(const mfem::Operator &) override { check(lpf_jacobi_setup(c_.get()), "lpf_jacobi_setup"); }
    void SetRelTol(double r) { rel_ = r; }
    void SetAbsTol(double a) { abs_ = a; }
    void SetMaxIter(int m) { max_iter_ = m; }
    void SetPrintLevel(int) {}
    void Mult(const mfem::Vector &b, mfem::Vector &x) const override
    {
        double *bl = c_.to_L(0, 0, b.Read()), *xl = c_.to_L(0, 1, x.Read());
        check(lpf_pcg(c_.get(), bl, xl, rel_, abs_, max_iter_, &info_), "lpf_pcg");
        check(lpf_restrict(c_.get(), 0, xl, x.Write()), "lpf_restrict");
        c_.done();
    }
    int GetNumIterations() const { return info_.iterations; }
    bool GetConverged() const { return info_.converged != 0; }
    double GetFinalNorm() const { return info_.final_norm; }

private:
    B200Context &c_;
    double rel_ = 1e-12, abs_ = 0.0;
    int max_iter_ = 1000;
    mutable lpf_pcg_info info_{};
};

/// rhs_linear (PF_linear_par_partial.cpp:36-245): state = [eta ; phi_fs] in surface TRUE dofs (:385-390).
class B200RhsLinear : public mfem::TimeDependentOperator {
public:
    /// cgen / cabs / cabsy: relaxation weights on the LOCAL surface dofs (Cgen_gf / Cabs_gf data, :415-447); cabsy is the third
    /// weight of Solvers/cylinder-diffraction.cpp:373-389, NULL for the wave-tank drivers
    B200RhsLinear(B200Context &c, const lpf_rhs_params &prm, const mfem::Vector *cgen, const mfem::Vector *cabs,
                  const mfem::Vector *cabsy = nullptr)
        : mfem::TimeDependentOperator(2 * lpf_ntrue(c.get(), 1)), c_(c), nt_(lpf_ntrue(c.get(), 1)), nl_(lpf_nsurf(c.get()))
    {
        check(lpf_pa_setup(c_.get()), "lpf_pa_setup");
        check(lpf_jacobi_setup(c_.get()), "lpf_jacobi_setup");
        check(lpf_rhs_setup(c_.get(), &prm, cgen ? cgen->HostRead() : nullptr, cabs ? cabs->HostRead() : nullptr), "lpf_rhs_setup");
        if (cabsy) check(lpf_rhs_set_cabsy(c_.get(), cabsy->HostRead()), "lpf_rhs_set_cabsy");
    }
    /// eta envelope of cylinder-diffraction.cpp:410-444: call after every Step once t >= t_last_start
    void EnvelopeReset() { check(lpf_envelope_reset(c_.get()), "lpf_envelope_reset"); }
    void EnvelopeUpdate(const mfem::Vector &state) { check(lpf_envelope_update(c_.get(), state_to_L(state, 0)), "lpf_envelope_update"); c_.done(); }
    void EnvelopeGet(mfem::Vector &env, double scale) { check(lpf_envelope_get(c_.get(), env.HostWrite(), scale), "lpf_envelope_get"); }
    void Mult(const mfem::Vector &x, mfem::Vector &dxdt) const override
    {
        double *xl = state_to_L(x, 0), *dl = c_.work(1, 1);
        check(lpf_rhs(c_.get(), GetTime(), xl, dl), "lpf_rhs");
        state_to_T(dl, dxdt);
        c_.done();
    }
    /// [eta_T ; phi_T] -> [eta_L ; phi_L] in work buffer `slot` of the surface space, and back
    double *state_to_L(const mfem::Vector &xT, int slot) const
    {
        double *xl = c_.work(1, slot);
        const double *xt = xT.Read();
        check(lpf_prolong(c_.get(), 1, xt, xl), "lpf_prolong");
        check(lpf_prolong(c_.get(), 1, xt + nt_, xl + nl_), "lpf_prolong");
        return xl;
    }
    void state_to_T(const double *xl, mfem::Vector &xT) const
    {
        double *xt = xT.Write();
        check(lpf_restrict(c_.get(), 1, xl, xt), "lpf_restrict");
        check(lpf_restrict(c_.get(), 1, xl + nl_, xt + nt_), "lpf_restrict");
    }
    lpf_ctx *ctx() const { return c_.get(); }
    const B200Context &context() const { return c_; }

private:
    B200Context &c_;
    int nt_, nl_;
};

/// RK4Solver (PF_linear_par_partial.cpp:472,483,494); the stage vectors stay on the device.
class B200RK4Solver : public mfem::ODESolver {
public:
    void Init(mfem::TimeDependentOperator &f) override
    {
        mfem::ODESolver::Init(f);
        rhs_ = dynamic_cast<B200RhsLinear *>(&f);
        if (!rhs_) mfem::mfem_error("B200RK4Solver needs a B200RhsLinear");
    }
    void Step(mfem::Vector &x, double &t, double &dt) override
    {
        double *xl = rhs_->state_to_L(x, 0);
        check(lpf_rk4_step(rhs_->ctx(), xl, &t, dt), "lpf_rk4_step");
        rhs_->state_to_T(xl, x);
        rhs_->context().done();
    }

private:
    B200RhsLinear *rhs_ = nullptr;
};

}  // namespace lpf_mfem
