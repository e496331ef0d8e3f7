// B200 twin of the reference's verification drivers Convergence_and_Scaling/convergence-parallel-partial.cpp (p-convergence
// of the time stepper, orders 1..8 on the 3-element periodic wave-tank.mesh) and convergence-parallel-partial-hconv.cpp
// (h-convergence at order 4): one wave period in 150 (+1, as in the reference's loop :255) RK4 steps, no relaxation
// zones, then the nodal max-norm error of eta and phi_fs against the Airy wave at the final time.  Same output columns
// as the reference's data/pf-parallel-pconv-eta.txt:  order dofs eta_inf_error  (+ phi_fs error, the ||w~ - w_exact||_inf of
// convergence-parallel-partial-hconv.cpp:331-350 evaluated on the final state, CG iterations).
// Defaults are the reference's constants (CG rel-tol 1e-24 / 2000 iterations: every solve runs to the cap, :96-99).
//   --mode p|h  --orders 1,2,..,8  --levels 0,1,2  --nsteps 150  --rel-tol 1e-24  --max-iter 2000  --gpus 1  --out <file>
#include <mutex>
#include <sstream>

#include "lpf_drivers.hpp"

using namespace lpfd;

int main(int argc, char *argv[])
{
    Args a{argc, argv};
    try {
        const bool hconv = std::string(a.get("--mode", "p")) == "h";
        const int num_procs = a.geti("--gpus", 1);
        const int nsteps = a.geti("--nsteps", 150);
        const double rel_tol = a.getd("--rel-tol", 1e-24);
        const int max_iter = a.geti("--max-iter", hconv ? 5000 : 2000);
        std::vector<int> list;
        { std::stringstream ss(a.get(hconv ? "--levels" : "--orders", hconv ? "0,1,2" : "1,2,3,4,5,6,7,8")); std::string t; while (std::getline(ss, t, ',')) list.push_back(std::atoi(t.c_str())); }
        const std::string out = a.get("--out", "");
        FILE *fout = out.empty() ? nullptr : fopen(out.c_str(), "w");
        Wave w;                                                       // lambda = 1, kh = 1, H = 0.01
        const double dt = w.T / nsteps;
        if (fout) fprintf(fout, "# order dofs eta_inf_error k=%g, kh=%g, omega=%g, T=%g, cwave=%g dt=%g\n", w.k, w.kh, w.omega, w.T, w.cwave, dt);
        printf("%s  dofs  ||eta - eta_exact||_inf  ||phi_fs - phi_fs_exact||_inf  ||w - w_exact||_inf  CG iterations of the last step\n", hconv ? "level" : "order");
        World world(num_procs, std::string(a.get("--comm", "p2p")) == "nccl");
        world.parse_options(a.get("--opt", ""));   // e.g. --opt affine=0,deterministic=1 (lpf_set_option)
        for (int it : list) {
            const int order = hconv ? a.geti("--order", 4) : it;
            std::unique_ptr<Mesh> mesh(Mesh::FromName(a.get("--mesh", "wave-tank.mesh")));
            for (int i = 0; i < (hconv ? it : a.geti("--ref", 0)); i++) mesh->UniformRefinement();
            std::mutex io;
            double err_eta = 0.0, err_phi = 0.0, err_w = 0.0;
            long dofs = 0;
            std::vector<int> its;
            world.run([&](int myid) {
                RankSpace fespace(*mesh, order, num_procs, myid);
                const lpf_space_desc &d = fespace.desc;
                const int ns = d.n_surf;
                std::vector<double> state(2 * (size_t)ns);
                for (int s = 0; s < ns; s++) {
                    state[s] = w.eta(0.0, d.surf_xy[2 * s], d.surf_xy[2 * s + 1]);
                    state[ns + s] = w.phi_fs(0.0, d.surf_xy[2 * s], d.surf_xy[2 * s + 1]);
                }
                RhsLinear surface(fespace, myid, world);
                surface.Setup(w.params(dt, false, rel_tol, max_iter), nullptr, nullptr);
                surface.SetState(state);
                double t = 0.0;
                for (int step = 0; step < nsteps + 1; step++) surface.Step(t, dt);      // nsteps + 1 steps: t = T + dt
                surface.GetState(state);
                const std::vector<double> wt = surface.WTilde(t);                     // hconv.cpp:331: surface.GetWTilde()
                double ee = 0.0, ep = 0.0, ew = 0.0;
                for (int s = 0; s < ns; s++) {
                    ee = std::max(ee, std::fabs(state[s] - w.eta(t, d.surf_xy[2 * s], d.surf_xy[2 * s + 1])));
                    ep = std::max(ep, std::fabs(state[ns + s] - w.phi_fs(t, d.surf_xy[2 * s], d.surf_xy[2 * s + 1])));
                    ew = std::max(ew, std::fabs(wt[s] - w.w_surface(t, d.surf_xy[2 * s], d.surf_xy[2 * s + 1])));
                }
                std::lock_guard<std::mutex> lk(io);
                err_eta = std::max(err_eta, ee); err_phi = std::max(err_phi, ep); err_w = std::max(err_w, ew);      // MPI_Allreduce(MAX) (:291, hconv :334-336)
                if (myid == 0) { dofs = d.n_true_global; its = surface.LastIterations(); }
            });
            printf("%d  %ld  %.6e  %.6e  %.6e  %d %d %d %d\n", it, dofs, err_eta, err_phi, err_w, its[0], its[1], its[2], its[3]);
            if (fout) fprintf(fout, "%d %ld %.12e %.12e\n", it, dofs, err_eta, err_w);
        }
        if (fout) fclose(fout);
    } catch (const std::exception &e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
