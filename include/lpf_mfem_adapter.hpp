// lpf_mfem_adapter.hpp -- the thin MFEM-side classes a maintainer of the reference would add to run its
// hot path on liblpf_b200.so without touching the drivers' call sites.
//
// The reference has no FFI: its hot path sits behind MFEM's virtual interfaces.  Each class below derives
// from the MFEM base class the reference uses and forwards to one C-ABI call of include/lpf_b200.h:
//
//   B200DiffusionIntegrator : mfem::BilinearFormIntegrator   AssemblePA / AddMultPA / AssembleDiagonalPA
//        replaces `new DiffusionIntegrator`                  Solvers/PF_linear_par_partial.cpp:119
//   B200LaplaceOperator     : mfem::Operator                 Mult = constrained P^T A P          :155 (A_loc)
//   B200JacobiPCG           : mfem::Solver                   Mult = CGSolver + OperatorJacobiSmoother   :124,:157-164
//   B200RhsLinear           : mfem::TimeDependentOperator    Mult = rhs_linear::Mult             :130-244
//   B200RK4Solver           : mfem::ODESolver                Step = RK4Solver::Step              :472-494
//
// With a real MFEM (mfem.hpp on the include path, built with CUDA and `mfem::Device device("cuda")`) the
// Vector::Read()/ReadWrite() pointers are device pointers and no copy happens.  In this repository MFEM is
// not available, so the header is compile-checked against drivers/stub/mfem.hpp, which declares only the
// MFEM signatures used here ([MFEM] marks calls that must be re-verified against a real checkout).
#pragma once
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

/// Builds the plain descriptor the device context needs from MFEM objects (SURVEY.md 8b last row).
struct SpaceDescBuilder {
    std::vector<double> jac, corners;
    std::vector<int> gather, ess, surf2vol, surf_elems, surf_mult;
    std::vector<double> surf_xy;
    lpf_space_desc desc{};

    SpaceDescBuilder(const mfem::FiniteElementSpace &fes, const mfem::Array<int> &ess_tdof_list,
                     const mfem::Array<int> &surf_vdofs /* parent vdofs of the ParSubMesh space */,
                     const mfem::Vector &surf_coords /* [n_surf][2] */)
    {
        const mfem::FiniteElement &el = *fes.GetFE(0);
        const int p = el.GetOrder(), D = p + 1, Q = p + 2, D3 = D * D * D, Q3 = Q * Q * Q;
        const int ne = fes.GetNE();
        // [MFEM] the rule DiffusionIntegrator::GetRule picks for tensor elements: order 2p + dim - 1
        const mfem::IntegrationRule &ir = mfem::IntRules.Get(el.GetGeomType(), 2 * p + 2);
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
        ess.assign(ess_tdof_list.GetData(), ess_tdof_list.GetData() + ess_tdof_list.Size());
        surf2vol.assign(surf_vdofs.GetData(), surf_vdofs.GetData() + surf_vdofs.Size());
        surf_xy.assign(surf_coords.GetData(), surf_coords.GetData() + surf_coords.Size());
        // Geometry of GetDerivative (and of the affine fast path): the 8 corners of every hex, lexicographic, from the mesh
        // transformation -- valid for the trilinear meshes the reference ships (vertex coordinates alone would be wrong on
        // its periodic meshes, whose geometry lives in the L2 nodes).  [MFEM] ElementTransformation::Transform
        corners.resize((size_t)ne * 24);
        {
            double buf[3] = {0, 0, 0};
            mfem::Vector pt(buf, 3);
            for (int e = 0; e < ne; e++) {
                mfem::ElementTransformation *T = fes.GetMesh()->GetElementTransformation(e);
                for (int c = 0; c < 8; c++) {
                    mfem::IntegrationPoint ip;
                    ip.Set3(c & 1, (c >> 1) & 1, (c >> 2) & 1);
                    T->Transform(ip, pt);
                    for (int a = 0; a < 3; a++) corners[(size_t)e * 24 + c * 3 + a] = buf[a];
                }
            }
        }
        desc.order = p; desc.ne = ne; desc.ndof = fes.GetVSize();
        // q-data from MFEM's own Jacobians (curved meshes included); corners serve the surface derivative.  Pass
        // desc.jac = nullptr instead to let the library build q-data from the corners and enable its affine fast path.
        desc.corners = corners.data(); desc.jac = jac.data(); desc.gather = gather.data();
        desc.n_ess = (int)ess.size(); desc.ess = ess.data();
        desc.n_surf = (int)surf2vol.size(); desc.surf2vol = surf2vol.data(); desc.surf_xy = surf_xy.data();
        desc.nranks = 1; desc.rank = 0;
        desc.n_true_global = fes.GetTrueVSize();
        desc.n_surf_global = desc.n_surf;
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

private:
    lpf_ctx *ctx_;
};

/// Drop-in for `new DiffusionIntegrator` under AssemblyLevel::PARTIAL (PF_linear_par_partial.cpp:119-121).
class B200DiffusionIntegrator : public mfem::BilinearFormIntegrator {
public:
    explicit B200DiffusionIntegrator(B200Context &c) : c_(c) {}
    void AssemblePA(const mfem::FiniteElementSpace &) override { check(lpf_pa_setup(c_.get()), "lpf_pa_setup"); }
    void AddMultPA(const mfem::Vector &x, mfem::Vector &y) const override   // E-vectors, y += A_E x
    {
        check(lpf_pa_apply_E(c_.get(), x.Read(), y.ReadWrite()), "lpf_pa_apply_E");
    }
    void AssembleDiagonalPA(mfem::Vector &diag) override
    {
        // MFEM hands an E-vector here and applies G^T itself; the fused L-vector diagonal is lpf_diag()
        mfem::mfem_error("B200DiffusionIntegrator: use B200LaplaceOperator::AssembleDiagonal (fused L-vector diagonal)");
        (void)diag;
    }

private:
    B200Context &c_;
};

/// The constrained operator FormLinearSystem returns (A_loc, PF_linear_par_partial.cpp:152-155).
class B200LaplaceOperator : public mfem::Operator {
public:
    explicit B200LaplaceOperator(B200Context &c) : mfem::Operator(lpf_ndof(c.get())), c_(c)
    {
        check(lpf_pa_setup(c_.get()), "lpf_pa_setup");
    }
    void Mult(const mfem::Vector &x, mfem::Vector &y) const override
    {
        check(lpf_apply_T(c_.get(), x.Read(), y.Write()), "lpf_apply_T");
    }
    void AssembleDiagonal(mfem::Vector &diag) const override { check(lpf_diag(c_.get(), diag.Write()), "lpf_diag"); }

private:
    B200Context &c_;
};

/// CGSolver + OperatorJacobiSmoother in one object (PF_linear_par_partial.cpp:124,157-164).
class B200JacobiPCG : public mfem::Solver {
public:
    explicit B200JacobiPCG(B200Context &c) : mfem::Solver(lpf_ndof(c.get())), c_(c) { iterative_mode = true; }
    void SetOperator(const mfem::Operator &) override { check(lpf_jacobi_setup(c_.get()), "lpf_jacobi_setup"); }
    void SetRelTol(double r) { rel_ = r; }
    void SetAbsTol(double a) { abs_ = a; }
    void SetMaxIter(int m) { max_iter_ = m; }
    void SetPrintLevel(int) {}
    void Mult(const mfem::Vector &b, mfem::Vector &x) const override
    {
        check(lpf_pcg(c_.get(), b.Read(), x.ReadWrite(), rel_, abs_, max_iter_, &info_), "lpf_pcg");
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

/// rhs_linear (PF_linear_par_partial.cpp:36-245): state = [eta ; phi_fs] in surface true dofs.
class B200RhsLinear : public mfem::TimeDependentOperator {
public:
    /// cabsy: the third weight of Solvers/cylinder-diffraction.cpp:373-389 (Cabsy_gf), NULL for the wave-tank drivers
    B200RhsLinear(B200Context &c, const lpf_rhs_params &prm, const mfem::Vector *cgen, const mfem::Vector *cabs,
                  const mfem::Vector *cabsy = nullptr)
        : mfem::TimeDependentOperator(2 * lpf_nsurf(c.get())), c_(c)
    {
        check(lpf_pa_setup(c_.get()), "lpf_pa_setup");
        check(lpf_jacobi_setup(c_.get()), "lpf_jacobi_setup");
        check(lpf_rhs_setup(c_.get(), &prm, cgen ? cgen->HostRead() : nullptr, cabs ? cabs->HostRead() : nullptr), "lpf_rhs_setup");
        if (cabsy) check(lpf_rhs_set_cabsy(c_.get(), cabsy->HostRead()), "lpf_rhs_set_cabsy");
    }
    /// eta envelope of cylinder-diffraction.cpp:410-444: call after every Step once t >= t_last_start
    void EnvelopeReset() { check(lpf_envelope_reset(c_.get()), "lpf_envelope_reset"); }
    void EnvelopeUpdate(const mfem::Vector &state) { check(lpf_envelope_update(c_.get(), state.Read()), "lpf_envelope_update"); }
    void EnvelopeGet(mfem::Vector &env, double scale) { check(lpf_envelope_get(c_.get(), env.HostWrite(), scale), "lpf_envelope_get"); }
    void Mult(const mfem::Vector &x, mfem::Vector &dxdt) const override
    {
        check(lpf_rhs(c_.get(), GetTime(), x.Read(), dxdt.Write()), "lpf_rhs");
    }
    lpf_ctx *ctx() const { return c_.get(); }

private:
    B200Context &c_;
};

/// RK4Solver (PF_linear_par_partial.cpp:472,483,494); keeps the stage vectors on the device.
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
        check(lpf_rk4_step(rhs_->ctx(), x.ReadWrite(), &t, dt), "lpf_rk4_step");
    }

private:
    B200RhsLinear *rhs_ = nullptr;
};

}  // namespace lpf_mfem
