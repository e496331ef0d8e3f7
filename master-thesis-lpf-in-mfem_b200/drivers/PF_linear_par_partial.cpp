// B200 twin of the reference's canonical driver Solvers/PF_linear_par_partial.cpp (and, with --cylinder, of
// Solvers/cylinder-diffraction.cpp up to the envelope): linear potential-flow free-surface waves, PA Laplace +
// Jacobi-PCG per RK4 stage, relaxation zones.  Defaults are the reference's source constants
// (PF_linear_par_partial.cpp:255-261, 298-306, 359-360, 415-447); every one can be overridden:
//   --mesh wave-tank-finite.mesh|<file>  --order 4  --ref 0  --nsteps 180  --periods 5  --gpus 1
//   --no-relax   --cylinder (mesh tests/meshes/cylinder_half.mesh, third absorption weight, eta envelope)
//   --serial-params: the constants of Solvers/PF_linear_serial.cpp (BASELINE config 2: order 5, 1 refinement, H = 0.05,
//                    wave chosen by its period T = 1.13392/3, Ng = Ns = 2, 800 steps over 8 T, PCG rel 1e-12 / 400 it) --
//                    with the PA + Jacobi-PCG solver of this library in place of the assembled matrix + GS-PCG
#include <algorithm>
#include <chrono>
#include <mutex>

#include "lpf_drivers.hpp"

using namespace lpfd;

int main(int argc, char *argv[])
{
    Args a{argc, argv};
    try {
        const bool cyl = a.has("--cylinder");
        const bool serial = a.has("--serial-params");
        const int order = a.geti("--order", serial ? 5 : 4);                    // PF_linear_serial.cpp:266-267
        const int ref_levels = a.geti("--ref", serial ? 1 : 0);
        const int num_procs = a.geti("--gpus", 1);
        const bool relax = !a.has("--no-relax");
        const std::string mesh_file = a.get("--mesh", cyl ? "../../../tests/meshes/cylinder_half.mesh" : "wave-tank-finite.mesh");
        std::unique_ptr<Mesh> mesh(Mesh::FromName(mesh_file));
        for (int i = 0; i < ref_levels; i++) mesh->UniformRefinement();

        Wave w;                                   // H = 0.01, g = 9.81, lambda = 1, kh = 1  (:287-306)
        double lo[3], hi[3];
        mesh->GetBoundingBox(lo, hi);
        if (serial) { w.H = 0.05; w.period_mode(1.13392 / 3, hi[2] - lo[2], 40); }   // PF_linear_serial.cpp:307-327
        const int nsteps = a.geti("--nsteps", serial ? 800 : cyl ? 350 : 180);        // :359 / cylinder-diffraction.cpp:252 / serial :341-344
        const double t_final = a.getd("--periods", serial ? 8.0 : cyl ? 10.0 : 5.0) * w.T;
        const double dt = t_final / nsteps;
        const double t_last_start = t_final - w.T;
        printf("Wave parameters:\n  Lx     = %g\n  lwave  = %g\n  kh     = %g\n  k      = %g\n  cwave  = %g\n  T      = %g\n  omega  = %g\n  H      = %g\n",
               hi[0] - lo[0], w.lambda, w.kh, w.k, w.cwave, w.T, w.omega, w.H);
        printf("%g\nRunning on %d GPUs\nStarting time integration with %d steps\n", dt, num_procs, nsteps);

        World world(num_procs, std::string(a.get("--comm", "p2p")) == "nccl");
        world.parse_options(a.get("--opt", ""));   // e.g. --opt affine=0,deterministic=1 (lpf_set_option)
        std::mutex io;
        double eta_max_global = 0.0;
        world.run([&](int myid) {
            RankSpace fespace(*mesh, order, num_procs, myid);
            const lpf_space_desc &d = fespace.desc;
            const int ns = d.n_surf;
            // ICs: eta = H/2 cos(phase), phi_fs = -H/2 c coth(kh) sin(phase) at t = 0  (:374-400)
            std::vector<double> state(2 * (size_t)ns), cgen(ns), cabs(ns);
            for (int s = 0; s < ns; s++) {
                const double x = d.surf_xy[2 * s], y = d.surf_xy[2 * s + 1];
                state[s] = w.eta(0.0, x, y);
                state[ns + s] = w.phi_fs(0.0, x, y);
                // relaxation functions Cgen / Cabs (:415-447); cylinder: Ng 2.5, Ns 4, plus Cabsy over the last
                // 3 lambda in y (cylinder-diffraction.cpp:373-389) -- absorption weights add up in the RHS
                const double Ng = serial ? 2.0 : 2.5, Ns = serial ? 2.0 : 4.0, xg0 = lo[0], xg1 = xg0 + Ng * w.lambda, x1 = hi[0], x0 = x1 - Ns * w.lambda;
                double cg;
                if (x <= xg0) cg = 1.0; else if (x >= xg1) cg = 0.0;
                else { const double xi = (x - xg0) / (xg1 - xg0); cg = 1 - (-2.0 * xi * xi * xi + 3.0 * xi * xi); }
                double ca;
                if (x <= x0) ca = 0.0; else if (x >= x1) ca = 1.0; else ca = std::pow((x - x0) / (x1 - x0), 5.0);
                if (cyl) {
                    const double y1 = hi[1], y0 = y1 - 3.0 * w.lambda;
                    if (y >= y1) ca += 1.0; else if (y > y0) ca += std::pow((y - y0) / (y1 - y0), 5.0);
                }
                cgen[s] = cg; cabs[s] = ca;
            }
            RhsLinear surface(fespace, myid, world);
            surface.Setup(w.params(dt, relax, 1e-12, serial ? 400 : cyl ? 2000 : 1000), cgen.data(), cabs.data());   // :157-164, tau = dt :470
            surface.SetState(state);
            std::vector<double> env(ns, -1e300);
            // ParaView output of eta / phi_fs every 5 steps (:453-467, 505-514); off unless --paraview <name> is given
            const std::string pv_name = a.get("--paraview", "");
            std::vector<double> eta_host(ns), phi_host(ns);
            ParaViewDataCollection pv_fs(pv_name, fespace, myid, num_procs, hi[2]);
            pv_fs.SetPrefixPath("ParaView");
            pv_fs.SetLevelsOfDetail(order);
            pv_fs.SetHighOrderOutput(true);
            pv_fs.RegisterField("eta", &eta_host);
            pv_fs.RegisterField("phi_fs", &phi_host);
            double t = 0.0;
            const auto t0 = std::chrono::steady_clock::now();
            for (int step = 0; step < nsteps + 1; step++) {               // nsteps + 1 steps as in the reference (:492)
                surface.Step(t, dt);
                if (cyl && t >= t_last_start) {                            // eta envelope over the last period
                    surface.GetState(state);
                    for (int s = 0; s < ns; s++) env[s] = std::max(env[s], state[s]);
                }
                if (!pv_name.empty() && step % 5 == 0) {
                    surface.GetState(state);
                    std::copy(state.begin(), state.begin() + ns, eta_host.begin());
                    std::copy(state.begin() + ns, state.end(), phi_host.begin());
                    pv_fs.SetCycle(step); pv_fs.SetTime(t); pv_fs.Save();
                }
                if (myid == 0 && step % 10 == 0) {
                    auto it = surface.LastIterations();
                    printf("Step %d / %d, t = %g   (CG iterations per stage: %d %d %d %d)\n", step, nsteps, t, it[0], it[1], it[2], it[3]);
                }
            }
            surface.GetState(state);
            const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            double emax = 0.0;
            for (int s = 0; s < ns; s++) emax = std::max(emax, std::fabs(state[s]));
            std::lock_guard<std::mutex> lk(io);
            eta_max_global = std::max(eta_max_global, emax);
            if (myid == 0) printf("rank 0: %d local dofs of %ld, %d surface dofs, %.3f s total, %.3f ms per RK4 step\n",
                                  d.ndof, d.n_true_global, ns, sec, 1e3 * sec / (nsteps + 1));
            if (cyl && myid == 0) {
                double rmax = 0.0;
                for (int s = 0; s < ns; s++) rmax = std::max(rmax, env[s] * 2.0 / w.H);
                printf("max run-up envelope 2 eta_max / H on rank 0 = %g\n", rmax);
            }
        });
        printf("max |eta| at t_final = %g (H/2 = %g)\n", eta_max_global, 0.5 * w.H);
    } catch (const std::exception &e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
