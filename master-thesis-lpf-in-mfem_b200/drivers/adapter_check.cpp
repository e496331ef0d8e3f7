// Runs include/lpf_mfem_adapter.hpp against the functional MFEM stand-in (drivers/stub/mfem.hpp):
//
//   adapter_check gpu <mesh: file | tank:nx,ny,nz> <order> <out.bin>
//       every adapter class on GPU 0, written the way the reference's call sites use them
//       (Solvers/PF_linear_par_partial.cpp:118-124,155-166,472-494); results go to out.bin, tests/test_gpu_adapter.py
//       recomputes them with direct C-ABI calls and compares.
//   adapter_check host-par <nranks> <mesh> <order>
//       no GPU: for every rank, the halo plan that SpaceDescBuilder derives from the (stand-in) ParFiniteElementSpace's
//       GroupCommunicator must equal the library's own plan (lpf_space_desc_get) -- neighbour lists, ownership, and the
//       rank-ordered reduction sources.
#include <cstdio>
#include <cstring>
#include <memory>

#include "lpf_mfem_adapter.hpp"

using namespace mfem;

static Mesh *make_mesh(const std::string &spec)
{
    if (spec.rfind("tank:", 0) == 0) {
        int nx, ny, nz;
        if (std::sscanf(spec.c_str() + 5, "%d,%d,%d", &nx, &ny, &nz) != 3) mfem_error("tank:nx,ny,nz");
        return new Mesh(lpf_mesh_make_wave_tank(nx, ny, nz, 1.0, 0.1, 1.0 / (2.0 * M_PI), 1));
    }
    return new Mesh(spec.c_str(), 1, 1);
}

static void write_vec(FILE *f, const char *name, const Vector &v)
{
    const int n = v.Size(), len = (int)std::strlen(name);
    std::fwrite(&len, sizeof(int), 1, f); std::fwrite(name, 1, len, f);
    std::fwrite(&n, sizeof(int), 1, f); std::fwrite(v.HostRead(), sizeof(double), n, f);
}

static double noise(int i) { unsigned long long z = (i + 0x9E3779B97F4A7C15ull) * 0xBF58476D1CE4E5B9ull; z ^= z >> 29; return (double)(z % 2000001ull) / 1e6 - 1.0; }

static int run_gpu(const std::string &mesh_spec, int order, const char *out)
{
    Device device("cuda");
    std::unique_ptr<Mesh> mesh(make_mesh(mesh_spec));
    H1_FECollection fec(order, 3);
    FiniteElementSpace fespace(mesh.get(), &fec);
    Array<int> ess_bdr(mesh->bdr_attributes.Max());
    ess_bdr = 0; ess_bdr[2 - 1] = 1;
    Array<int> ess_tdof;
    fespace.GetEssentialTrueDofs(ess_bdr, ess_tdof);                              // :407-412
    // what the driver derives from the free-surface SubMesh (:279-285): parent dof and (x, y) of every surface dof
    const lpf_space_desc &sd = fespace.shim_desc();
    Array<int> surf_vdofs(sd.surf2vol, sd.n_surf);
    Vector surf_xy(2 * sd.n_surf);
    std::copy(sd.surf_xy, sd.surf_xy + 2 * sd.n_surf, surf_xy.HostWrite());

    lpf_mfem::SpaceDescBuilder builder(fespace, ess_tdof, surf_vdofs, surf_xy);
    lpf_mfem::B200Context ctx(builder.desc, 0);
    FILE *f = std::fopen(out, "wb");
    if (!f) mfem_error("cannot open output file");
    const int nT = fespace.GetTrueVSize(), D3 = (order + 1) * (order + 1) * (order + 1), ne = fespace.GetNE();

    // a_loc_cach->AddDomainIntegrator(new DiffusionIntegrator); Assemble()  (:118-121)
    lpf_mfem::B200DiffusionIntegrator integ(ctx);
    integ.AssemblePA(fespace);
    Vector xE(ne * D3), yE(ne * D3), dE(ne * D3);
    for (int i = 0; i < ne * D3; i++) xE(i) = noise(i);
    yE = 1.0; dE = 0.0;
    integ.AddMultPA(xE, yE);                                                      // accumulates
    integ.AssembleDiagonalPA(dE);                                                 // :124 through the integrator
    write_vec(f, "AddMultPA", yE);
    write_vec(f, "AssembleDiagonalPA", dE);

    // FormLinearSystem -> A_loc; CGSolver + OperatorJacobiSmoother  (:124, :152-166)
    lpf_mfem::B200LaplaceOperator A(ctx);
    Vector x(nT), y(nT), diag(nT);
    for (int i = 0; i < nT; i++) x(i) = noise(i + 17);
    A.Mult(x, y);
    A.AssembleDiagonal(diag);
    write_vec(f, "Mult", y);
    write_vec(f, "AssembleDiagonal", diag);
    lpf_mfem::B200JacobiPCG cg(ctx);
    cg.SetRelTol(1e-12); cg.SetAbsTol(0.0); cg.SetMaxIter(2000); cg.SetPrintLevel(0);
    cg.SetOperator(A);
    Vector B(nT), X(nT);
    X = 0.0;
    for (int i = 0; i < ess_tdof.Size(); i++) X(ess_tdof[i]) = std::cos(6.0 * surf_xy(2 * i)) ;     // essential data (order of ess == surface here is irrelevant: any data)
    A.Mult(X, B);                      // B = A_c X has B[ess] = X[ess]; off the essential rows it is a generic right-hand side
    for (int i = 0; i < nT; i++) B(i) = 0.5 * B(i) + 1e-3 * noise(i + 5);
    for (int i = 0; i < ess_tdof.Size(); i++) B(ess_tdof[i]) = X(ess_tdof[i]);
    cg.Mult(B, X);
    write_vec(f, "CG_X", X);
    Vector its(3);
    its(0) = cg.GetNumIterations(); its(1) = cg.GetConverged(); its(2) = cg.GetFinalNorm();
    write_vec(f, "CG_info", its);

    // rhs_linear + RK4Solver (:36-245, :472-494), lambda-mode wave of :298-306, no relaxation zones (ss.cpp RHS)
    lpf_rhs_params prm{};
    const double g = 9.81, lambda = 1.0, kh = 1.0, k = 2.0 * M_PI / lambda, cw = std::sqrt((g / k) * std::tanh(kh)), T = lambda / cw;
    prm.g = g; prm.H = 0.01; prm.k = k; prm.kh = kh; prm.cwave = cw; prm.T = T; prm.omega = 2.0 * M_PI / T; prm.kx_dir = 1.0; prm.ky_dir = 0.0;
    prm.tau = T / 150; prm.n_ramp = 3.0; prm.use_relaxation = 0; prm.rel_tol = 1e-12; prm.abs_tol = 0.0; prm.max_iter = 2000;
    lpf_mfem::B200RhsLinear surface(ctx, prm, nullptr, nullptr);
    lpf_mfem::B200RK4Solver ode;
    ode.Init(surface);
    const int ns = sd.n_surf;
    Vector state(2 * ns), dstate(2 * ns);
    for (int s = 0; s < ns; s++) {
        const double ph = -k * surf_xy(2 * s);
        state(s) = 0.5 * prm.H * std::cos(ph);
        state(ns + s) = -0.5 * prm.H * cw * std::cosh(kh) / std::sinh(kh) * std::sin(ph);
    }
    surface.SetTime(0.0);
    surface.Mult(state, dstate);
    write_vec(f, "rhs", dstate);
    double t = 0.0, dt = T / 150;
    for (int step = 0; step < 2; step++) ode.Step(state, t, dt);
    write_vec(f, "state_after_2_steps", state);
    Vector tt(1); tt(0) = t;
    write_vec(f, "t", tt);
    std::fclose(f);
    std::printf("adapter_check gpu: %d hexes, order %d, %d true dofs, %d surface dofs, CG %d iterations -> %s\n", ne, order, nT, ns, cg.GetNumIterations(), out);
    return 0;
}

template <class T>
static bool same(const char *what, int rank, const std::vector<T> &a, const T *b, size_t n)
{
    if (a.size() == n && std::equal(a.begin(), a.end(), b)) return true;
    std::fprintf(stderr, "rank %d: %s differs (sizes %zu / %zu)\n", rank, what, a.size(), n);
    return false;
}

static bool same_reduction(const char *what, int rank, const lpf_mfem::HaloArrays &h, int n_shared, const int *shared, const int *red_off, const int *red_src)
{
    std::map<int, std::vector<int>> a, b;
    for (size_t i = 0; i < h.shared.size(); i++) a[h.shared[i]].assign(h.red_src.begin() + h.red_off[i], h.red_src.begin() + h.red_off[i + 1]);
    for (int i = 0; i < n_shared; i++) b[shared[i]].assign(red_src + red_off[i], red_src + red_off[i + 1]);
    if (a == b) return true;
    std::fprintf(stderr, "rank %d: %s differs\n", rank, what);
    return false;
}

static int run_host_par(int nranks, const std::string &mesh_spec, int order)
{
    std::unique_ptr<Mesh> mesh(make_mesh(mesh_spec));
    H1_FECollection fec(order, 3);
    bool ok = true;
    long shared_total = 0;
    for (int rank = 0; rank < nranks; rank++) {
        ParFiniteElementSpace pfes(mesh.get(), &fec, nranks, rank);               // ParMesh(MPI_COMM_WORLD, mesh) + ParFiniteElementSpace (:268,:277)
        Array<int> ess_bdr(mesh->bdr_attributes.Max());
        ess_bdr = 0; ess_bdr[2 - 1] = 1;
        Array<int> marker, ess_ldofs;
        pfes.GetEssentialVDofs(ess_bdr, marker);
        for (int i = 0; i < marker.Size(); i++) if (marker[i]) ess_ldofs.Append(i);
        const lpf_space_desc &d = pfes.shim_desc();
        Array<int> surf_vdofs(d.surf2vol, d.n_surf), surf_mult(d.surf_mult, d.n_surf);
        Vector surf_xy(std::max(1, 2 * d.n_surf));
        if (d.n_surf) std::copy(d.surf_xy, d.surf_xy + 2 * d.n_surf, surf_xy.HostWrite());
        lpf_mfem::SpaceDescBuilder b(pfes, ess_ldofs, surf_vdofs, surf_xy, pfes.SurfaceGroupComm(),
                                     [&](int s) { return d.surf_g[s]; }, surf_mult, d.n_surf_global);
        ok &= b.desc.ndof == d.ndof && b.desc.n_ess == d.n_ess && b.desc.nranks == nranks && b.desc.rank == rank;
        ok &= same("essential dofs", rank, b.ess, d.ess, (size_t)d.n_ess);
        ok &= same("owned", rank, b.halo.owned, d.owned, (size_t)d.ndof);
        ok &= same("nbr_rank", rank, b.halo.nbr_rank, d.nbr_rank, (size_t)d.n_nbr);
        ok &= same("nbr_offset", rank, b.halo.nbr_offset, d.nbr_offset, (size_t)d.n_nbr + 1);
        ok &= same("send_dofs", rank, b.halo.send, d.send_dofs, (size_t)(d.n_nbr ? d.nbr_offset[d.n_nbr] : 0));
        ok &= same_reduction("reduction sources", rank, b.halo, d.n_shared, d.shared_dofs, d.red_off, d.red_src);
        ok &= same("surface owned", rank, b.surf_owned, d.surf_owned, (size_t)d.n_surf);
        ok &= same("surface nbr_rank", rank, b.shalo.nbr_rank, d.s_nbr_rank, (size_t)d.s_n_nbr);
        ok &= same("surface send", rank, b.shalo.send, d.s_send, (size_t)(d.s_n_nbr ? d.s_nbr_offset[d.s_n_nbr] : 0));
        ok &= same_reduction("surface reduction sources", rank, b.shalo, d.s_n_shared, d.s_shared, d.s_red_off, d.s_red_src);
        ok &= pfes.GetTrueVSize() == (int)std::count(d.owned, d.owned + d.ndof, (uint8_t)1);
        shared_total += d.n_shared;
    }
    std::printf("adapter_check host-par: %d ranks, order %d, %ld shared dofs in total: %s\n", nranks, order, shared_total, ok ? "OK" : "MISMATCH");
    return ok ? 0 : 1;
}

int main(int argc, char **argv)
{
    if (argc >= 5 && std::string(argv[1]) == "gpu") return run_gpu(argv[2], std::atoi(argv[3]), argv[4]);
    if (argc >= 5 && std::string(argv[1]) == "host-par") return run_host_par(std::atoi(argv[2]), argv[3], std::atoi(argv[4]));
    std::fprintf(stderr, "usage: adapter_check gpu <mesh|tank:nx,ny,nz> <order> <out.bin> | adapter_check host-par <nranks> <mesh> <order>\n");
    return 2;
}
