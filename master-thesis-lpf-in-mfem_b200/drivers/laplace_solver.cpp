// B200 twin of Solvers/laplace_solver.cpp / laplace_solver_parallel_partial.cpp and of the convergence
// drivers Convergence_and_Scaling/laplace-parallel-{pconv,hconv}.cpp: one Laplace solve with Dirichlet data
// from the Airy potential on the free surface, errors of phi and of w = dphi/dz against the analytic solution
// (laplace_solver.cpp:70-81,126-138; laplace-parallel-pconv.cpp:197-213).
//   --mesh wave-tank.mesh  --orders 3 (or 1,2,..,8 for the p-convergence table)  --ref 2  --rel-tol 1e-12  --max-iter 1000
#include <sstream>

#include "lpf_drivers.hpp"

using namespace lpfd;

int main(int argc, char *argv[])
{
    Args a{argc, argv};
    try {
        const int ref_levels = a.geti("--ref", 2);                     // laplace_solver.cpp:12
        const double rel_tol = a.getd("--rel-tol", 1e-12);
        const int max_iter = a.geti("--max-iter", 1000);
        std::vector<int> orders;
        { std::stringstream ss(a.get("--orders", "3")); std::string t; while (std::getline(ss, t, ',')) orders.push_back(std::atoi(t.c_str())); }
        std::unique_ptr<Mesh> mesh(Mesh::FromName(a.get("--mesh", "wave-tank.mesh")));
        for (int i = 0; i < ref_levels; i++) mesh->UniformRefinement();
        double lo[3], hi[3];
        mesh->GetBoundingBox(lo, hi);
        Wave w;
        printf("order  dofs  iterations  converged  max|phi-phi_exact|  max|w-w_exact|\n");
        for (int order : orders) {
            RankSpace fes(*mesh, order, 1, 0);
            const lpf_space_desc &d = fes.desc;
            lpf_ctx *ctx = lpf_create(&d, 0, nullptr);
            if (!ctx) throw std::runtime_error(lpf_last_error());
            check(lpf_pa_setup(ctx), "Assemble");
            check(lpf_jacobi_setup(ctx), "Jacobi");
            auto xyz = fes.NodeCoordinates();
            std::vector<double> phi(d.ndof, 0.0), ex(d.ndof);
            for (int i = 0; i < d.ndof; i++) ex[i] = w.phi(0.0, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2] - lo[2]);
            for (int i = 0; i < d.n_ess; i++) phi[d.ess[i]] = ex[d.ess[i]];
            double *dphi = (double *)lpf_dev_alloc(sizeof(double) * d.ndof), *dw = (double *)lpf_dev_alloc(sizeof(double) * (d.n_surf + 1));
            check(lpf_memcpy_h2d(dphi, phi.data(), sizeof(double) * d.ndof), "h2d");
            lpf_pcg_info info{};
            check(lpf_laplace_solve(ctx, dphi, rel_tol, 0.0, max_iter, &info), "cg.Mult");
            check(lpf_surface_dz(ctx, dphi, dw), "GetDerivative");
            check(lpf_sync(ctx), "sync");
            std::vector<double> wt(d.n_surf);
            check(lpf_memcpy_d2h(phi.data(), dphi, sizeof(double) * d.ndof), "d2h");
            check(lpf_memcpy_d2h(wt.data(), dw, sizeof(double) * d.n_surf), "d2h");
            double e_phi = 0.0, e_w = 0.0;
            for (int i = 0; i < d.ndof; i++) e_phi = std::max(e_phi, std::fabs(phi[i] - ex[i]));
            for (int s = 0; s < d.n_surf; s++) e_w = std::max(e_w, std::fabs(wt[s] - w.w_surface(0.0, d.surf_xy[2 * s], d.surf_xy[2 * s + 1])));
            printf("%d  %d  %d  %d  %.6e  %.6e\n", order, d.ndof, info.iterations, info.converged, e_phi, e_w);
            lpf_dev_free(dphi); lpf_dev_free(dw);
            lpf_destroy(ctx);
        }
    } catch (const std::exception &e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
