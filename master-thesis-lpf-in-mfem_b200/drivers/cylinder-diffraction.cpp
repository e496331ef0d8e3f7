// B200 twin of the reference's Solvers/cylinder-diffraction.cpp (BASELINE config 4): a plane Airy wave diffracted
// by a vertical circular cylinder on mesh_cylinder_half.msh, generation zone at the inlet, absorption zones towards
// x_max and y_max, RK4 + Jacobi-PCG per stage; the eta envelope over the last period is sampled on the cylinder rim
// and written as "theta eta" (2 eta_max / H), next to the MacCamy-Fuchs series of Solvers/cylinder-exact.cpp:53-115
// (lpf_maccamy_fuchs: restated with the C++17 special functions std::cyl_bessel_j / std::cyl_neumann; Boost is absent here).
// Defaults are the reference's source constants (cylinder-diffraction.cpp:225-256, 341-389); all can be overridden:
//   --mesh <file> --order 4 --nsteps 350 --periods 10 --gpus 1 --comm p2p|nccl --out data/cylinder-diffraction.txt
//   --paraview <name> [--pv-every 1]  (the reference saves eta every step, :434-439; off by default here)
#include <algorithm>
#include <chrono>
#include <fstream>
#include <mutex>

#include "lpf_drivers.hpp"

using namespace lpfd;

int main(int argc, char *argv[])
{
    Args a{argc, argv};
    try {
        const std::string mesh_file = a.get("--mesh", "../../../tests/meshes/cylinder_half.mesh");
        const int order = a.geti("--order", 4);                                  // :226
        const int num_procs = a.geti("--gpus", 1);
        const double cx = a.getd("--cx", 4.0), cy = a.getd("--cy", 4.0), rad = a.getd("--radius", 0.5);   // :229-231
        Wave w;                                                                  // lambda = 1, kh = 1, H = 0.01 (:236-246)
        const int nsteps = a.geti("--nsteps", 350);                              // :250
        const double t_final = a.getd("--periods", 10.0) * w.T;
        const double dt = t_final / nsteps;
        const double t_last_start = t_final - w.T;
        const std::string out_file = a.get("--out", "data/cylinder-diffraction.txt");
        printf("%g\n", dt);

        std::unique_ptr<Mesh> mesh(Mesh::FromName(mesh_file));
        double lo[3], hi[3];
        mesh->GetBoundingBox(lo, hi);

        World world(num_procs, std::string(a.get("--comm", "p2p")) == "nccl");
        world.parse_options(a.get("--opt", ""));   // e.g. --opt affine=0,deterministic=1 (lpf_set_option)
        std::mutex io;
        std::vector<std::pair<double, double>> cyl_eta;           // gathered over the ranks (MPI_Allgatherv in the reference)
        long its_total = 0;
        world.run([&](int myid) {
            RankSpace fespace(*mesh, order, num_procs, myid);
            const lpf_space_desc &d = fespace.desc;
            const int ns = d.n_surf;
            std::vector<double> state(2 * (size_t)ns), cgen(ns), cabs(ns), cabsy(ns);
            const double Ng = 2.5, xg0 = lo[0], xg1 = xg0 + Ng * w.lambda;                  // :341-343
            const double Ns = 4.0, x1 = hi[0], x0 = x1 - Ns * w.lambda;                     // :357-359
            const double Ns_y = 3.0, y1 = hi[1], y0 = y1 - Ns_y * w.lambda;                 // :374-377
            for (int s = 0; s < ns; s++) {
                const double x = d.surf_xy[2 * s], y = d.surf_xy[2 * s + 1];
                state[s] = w.eta(0.0, x, y);                                                // :306-319
                state[ns + s] = w.phi_fs(0.0, x, y);
                if (x <= xg0) cgen[s] = 1.0; else if (x >= xg1) cgen[s] = 0.0;
                else { const double xi = (x - xg0) / (xg1 - xg0), q = 1.0 - xi; cgen[s] = -2.0 * q * q * q + 3.0 * q * q; }
                if (x <= x0) cabs[s] = 0.0; else if (x >= x1) cabs[s] = 1.0; else cabs[s] = std::pow((x - x0) / (x1 - x0), 5.0);
                if (y <= y0) cabsy[s] = 0.0; else if (y >= y1) cabsy[s] = 1.0; else cabsy[s] = std::pow((y - y0) / (y1 - y0), 5.0);
            }
            RhsLinear surface(fespace, myid, world);
            surface.Setup(w.params(dt, true, 1e-12, 2000), cgen.data(), cabs.data());       // cg: rel 1e-12, 2000 it (:136-141)
            surface.SetCabsy(cabsy.data());
            surface.SetState(state);
            surface.EnvelopeReset();
            // ParaView output of eta (:392-397, 434-439)
            const std::string pv_name = a.get("--paraview", "");
            const int pv_every = a.geti("--pv-every", 1);
            std::vector<double> eta_host(ns);
            ParaViewDataCollection pv_fs(pv_name, fespace, myid, num_procs, hi[2]);
            pv_fs.SetPrefixPath("ParaView");
            pv_fs.SetLevelsOfDetail(5 * order);
            pv_fs.SetHighOrderOutput(true);
            pv_fs.RegisterField("eta", &eta_host);
            double t = 0.0;
            long its = 0;
            const auto t0 = std::chrono::steady_clock::now();
            for (int step = 0; step < nsteps + 1; step++) {                                 // :416
                surface.Step(t, dt);
                if (t >= t_last_start) surface.EnvelopeUpdate();                            // :421
                if (!pv_name.empty() && step % pv_every == 0) {
                    surface.GetState(state);
                    std::copy(state.begin(), state.begin() + ns, eta_host.begin());
                    pv_fs.SetCycle(step); pv_fs.SetTime(t); pv_fs.Save();
                }
                if (step % 10 == 0) {
                    auto it = surface.LastIterations();
                    for (int v : it) its += v;
                    if (myid == 0) printf("step %d/%d t=%g   (CG iterations per stage: %d %d %d %d)\n", step, nsteps, t, it[0], it[1], it[2], it[3]);
                }
            }
            std::vector<double> env = surface.Envelope(2.0 / w.H);                          // :444
            const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            // rim vertices of this rank; shared vertices appear on several ranks and are dropped after the sort
            std::vector<std::pair<double, double>> mine;
            for (auto &pr : fespace.Rim(3, cx, cy, rad, 5e-3)) {                            // tol = 5e-3 (:477)
                const double val = env[pr.second];
                if (val == 0.0) continue;                                                   // :493
                mine.emplace_back(pr.first, val);
            }
            std::lock_guard<std::mutex> lk(io);
            cyl_eta.insert(cyl_eta.end(), mine.begin(), mine.end());
            if (myid == 0) {
                its_total = its;
                printf("rank 0: %d local dofs of %ld, %d surface dofs, %.3f s total, %.3f ms per RK4 step\n", d.ndof, d.n_true_global, ns,
                       sec, 1e3 * sec / (nsteps + 1));
            }
        });
        std::sort(cyl_eta.begin(), cyl_eta.end(), [](const auto &p, const auto &q) { return p.first < q.first; });   // :566-568
        std::ofstream fout(out_file);
        if (!fout) fprintf(stderr, "warning: cannot open %s, printing to stdout only\n", out_file.c_str());
        fout << "# theta(rad)  eta  eta_maccamy_fuchs\n";
        fout.precision(12);
        double prev_th = -1.0, err2 = 0.0, emax = 0.0;
        int n = 0;
        for (auto &p : cyl_eta) {
            if (p.first - prev_th < 1e-10) continue;                                        // :589
            const double ex = lpf_maccamy_fuchs(w.k, rad, rad, p.first, a.has("--exact-robust") ? -1e-10 : 1e-10, 400);
            fout << p.first << " " << p.second << " " << ex << "\n";
            printf("rim theta = %.6f  2 eta_max / H = %.6f   MacCamy-Fuchs %.6f\n", p.first, p.second, ex);
            err2 += (p.second - ex) * (p.second - ex);
            emax = std::max(emax, std::fabs(p.second - ex));
            prev_th = p.first;
            n++;
        }
        printf("Extracted %d points on cylinder rim\n", n);
        if (n) printf("run-up vs MacCamy-Fuchs: rms difference %.4f, max difference %.4f (in units of H/2)\n", std::sqrt(err2 / n), emax);
        (void)its_total;
    } catch (const std::exception &e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
