// mfem_crosscheck.cpp -- the one program that can PIN this repository's oracle and kernels against MFEM itself.
//
// MFEM is not available in the build image (no mfem.hpp, no hypre / MPI / METIS, no network), so this file is
// NOT compiled by build() and has never been run here; it is written against the MFEM 4.x API the reference's drivers use
// (Solvers/PF_linear_par_partial.cpp:113-166, Solvers/laplace_solver.cpp:99-113).  On a machine with MFEM:
//
//     mpicxx -O2 -I$MFEM_DIR tools/mfem_crosscheck.cpp -L$MFEM_DIR -lmfem $MFEM_LIBS -o mfem_crosscheck
//     ./mfem_crosscheck ../reference/Meshes/wave-tank-big8.mesh 4 0 dump_big8_p4.txt
//     python tools/mfem_crosscheck_compare.py dump_big8_p4.txt --mesh tank:128,2,16 --order 4      (needs a B200)
//
// It dumps, for every vector entry, the node coordinates and the values, so the comparison is independent of any dof
// numbering (tools/mfem_crosscheck_compare.py matches nodes geometrically, like tests/util.py dof_map):
//   section OPERATOR   x y z  u  (A_pa u)  (A_fa u)  diag_pa     -- unconstrained P^T A P, partial and full assembly
//   section SOLVE      x y z  phi  w                                -- FormLinearSystem + Jacobi-PCG (rel 1e-12) + GetDerivative(1,2,w)
//   section INFO       CG iterations, final norm
// u is a closed-form function of the coordinates, periodic in x with the mesh length, so both sides can evaluate it.
#include <fstream>
#include <iomanip>
#include <iostream>

#include "mfem.hpp"

using namespace mfem;

static double Lx_global = 1.0;
static double u_fun(const Vector &X)
{
    const double a = 2.0 * M_PI / Lx_global;
    return std::cos(3.0 * a * X(0) + 0.3) * std::sin(17.0 * X(1) + 0.4) * std::cosh(2.5 * X(2)) + 0.37 * std::sin(5.0 * a * X(0)) * std::cos(11.0 * X(2));
}

int main(int argc, char *argv[])
{
    Mpi::Init(argc, argv);
    Hypre::Init();
    if (argc < 5) { std::cerr << "usage: mfem_crosscheck <mesh> <order> <ref_levels> <out.txt>\n"; return 1; }
    const char *mesh_file = argv[1];
    const int order = std::atoi(argv[2]), ref_levels = std::atoi(argv[3]);
    Mesh mesh_serial(mesh_file, 1, 1);
    for (int i = 0; i < ref_levels; i++) mesh_serial.UniformRefinement();
    Vector bbmin, bbmax;
    mesh_serial.GetBoundingBox(bbmin, bbmax);
    Lx_global = bbmax(0) - bbmin(0);
    ParMesh mesh(MPI_COMM_WORLD, mesh_serial);
    H1_FECollection fec(order, mesh.Dimension());
    ParFiniteElementSpace fespace(&mesh, &fec);
    if (Mpi::WorldSize() != 1) { if (Mpi::Root()) std::cerr << "run on ONE rank: the dump is written from L-vectors\n"; return 1; }

    // node coordinates of the dofs
    ParGridFunction xc(&fespace), yc(&fespace), zc(&fespace);
    FunctionCoefficient cx([](const Vector &X) { return X(0); }), cy([](const Vector &X) { return X(1); }), cz([](const Vector &X) { return X(2); });
    xc.ProjectCoefficient(cx); yc.ProjectCoefficient(cy); zc.ProjectCoefficient(cz);
    ParGridFunction u(&fespace);
    FunctionCoefficient uc(u_fun);
    u.ProjectCoefficient(uc);

    // partial assembly (the hot path) and full assembly (the FA twin), unconstrained
    ParBilinearForm a_pa(&fespace), a_fa(&fespace);
    a_pa.SetAssemblyLevel(AssemblyLevel::PARTIAL);
    a_pa.AddDomainIntegrator(new DiffusionIntegrator);
    a_pa.Assemble();
    a_fa.AddDomainIntegrator(new DiffusionIntegrator);
    a_fa.Assemble();
    a_fa.Finalize();
    Vector y_pa(fespace.GetVSize()), y_fa(fespace.GetVSize()), diag(fespace.GetVSize());
    a_pa.Mult(u, y_pa);
    a_fa.Mult(u, y_fa);
    a_pa.AssembleDiagonal(diag);

    // the constrained solve of the drivers: essential data = u on boundary attribute 2, interior zero, b = 0
    Array<int> ess_bdr(mesh.bdr_attributes.Max()), ess_tdof;
    ess_bdr = 0; ess_bdr[2 - 1] = 1;
    fespace.GetEssentialTrueDofs(ess_bdr, ess_tdof);
    ParGridFunction phi(&fespace);
    phi = 0.0;
    phi.ProjectBdrCoefficient(uc, ess_bdr);
    ParLinearForm b(&fespace);
    b.Assemble();
    OperatorPtr A;
    Vector X, B;
    a_pa.FormLinearSystem(ess_tdof, phi, b, A, X, B);
    OperatorJacobiSmoother jacobi(a_pa, ess_tdof);
    CGSolver cg(MPI_COMM_WORLD);
    cg.SetRelTol(1e-12); cg.SetAbsTol(0.0); cg.SetMaxIter(5000); cg.SetPrintLevel(0);
    cg.SetPreconditioner(jacobi);
    cg.SetOperator(*A);
    cg.Mult(B, X);
    a_pa.RecoverFEMSolution(X, b, phi);
    ParGridFunction w(&fespace);
    phi.GetDerivative(1, 2, w);

    std::ofstream out(argv[4]);
    out << std::setprecision(17);
    out << "OPERATOR " << fespace.GetVSize() << "\n";
    for (int i = 0; i < fespace.GetVSize(); i++)
        out << xc(i) << " " << yc(i) << " " << zc(i) << " " << u(i) << " " << y_pa(i) << " " << y_fa(i) << " " << diag(i) << "\n";
    out << "SOLVE " << fespace.GetVSize() << "\n";
    for (int i = 0; i < fespace.GetVSize(); i++) out << xc(i) << " " << yc(i) << " " << zc(i) << " " << phi(i) << " " << w(i) << "\n";
    out << "INFO " << cg.GetNumIterations() << " " << cg.GetConverged() << " " << cg.GetFinalNorm() << " " << Lx_global << "\n";
    std::cout << "mfem_crosscheck: " << fespace.GetVSize() << " dofs, CG " << cg.GetNumIterations() << " iterations -> " << argv[4] << std::endl;
    return 0;
}
