// Compile check of include/lpf_mfem_adapter.hpp against the stub MFEM declarations (drivers/stub/mfem.hpp).
// Instantiates every adapter class so that a signature drift in include/lpf_b200.h breaks the build.
#include "lpf_mfem_adapter.hpp"

int main()
{
    // never executed with the stub: only has to compile and link against liblpf_b200.so
    if (lpf_version() < 0) {
        mfem::FiniteElementSpace fes;
        mfem::Array<int> ess, surf;
        mfem::Vector xy;
        lpf_mfem::SpaceDescBuilder b(fes, ess, surf, xy);
        lpf_mfem::B200Context ctx(b.desc);
        lpf_mfem::B200DiffusionIntegrator integ(ctx);
        lpf_mfem::B200LaplaceOperator A(ctx);
        lpf_mfem::B200JacobiPCG cg(ctx);
        cg.SetOperator(A);
        lpf_rhs_params prm{};
        lpf_mfem::B200RhsLinear rhs(ctx, prm, nullptr, nullptr);
        lpf_mfem::B200RK4Solver rk;
        rk.Init(rhs);
        mfem::Vector x, y;
        double t = 0, dt = 1;
        integ.AssemblePA(fes);
        integ.AddMultPA(x, y);
        A.Mult(x, y);
        cg.Mult(x, y);
        rhs.Mult(x, y);
        rk.Step(x, t, dt);
    }
    return 0;
}
