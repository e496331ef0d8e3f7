// B200 twin of the reference's scaling drivers Convergence_and_Scaling/ss.cpp (mode 0 strong / 1 weak),
// ws.cpp (mesh chosen by rank count) and strongscaling.cpp: periodic tank, RHS without relaxation zones,
// 1 warm-up RK4 step + nsteps timed steps, max over ranks (ss.cpp:253-272).  Same output columns:
//     procs  order  par_ref  dofs  time_total[s]  time_per_step[s]
// and the same append-mode table "# mode order par_ref_level ranks dofs runtime[s]" (ss.cpp:143-144,284-285).
//   --gpus N  --mode 0|1|ws  --orders 3,4  --nsteps 10  --rel-tol 1e-8  --max-iter 300  --par-ref 1  --out data/strong-scaling.txt
#include <chrono>
#include <mutex>
#include <sstream>

#include "lpf_drivers.hpp"

using namespace lpfd;

int main(int argc, char *argv[])
{
    Args a{argc, argv};
    try {
        const int num_procs = a.geti("--gpus", 1);
        const std::string mode_s = a.get("--mode", "1");
        const bool ws = mode_s == "ws";
        const int mode = ws ? 1 : std::atoi(mode_s.c_str());
        const int nsteps = a.geti("--nsteps", 10);                    // ss.cpp:138
        const double rel_tol = a.getd("--rel-tol", 1e-8);             // ss.cpp:90
        const int max_iter = a.geti("--max-iter", 300);               // ss.cpp:93
        int par_ref_levels = a.geti("--par-ref", 1);                  // par_ref_levels_fixed / par_ref_base (ss.cpp:132-135)
        if (mode == 1 && !ws) par_ref_levels += (int)std::llround(std::log2((double)num_procs) / 3.0);   // ss.cpp:172-180
        std::string mesh_name = "wave-tank-big.mesh";                  // ss.cpp:129
        if (ws) {                                                      // ws.cpp:116-128
            mesh_name = num_procs == 1 ? "wave-tank-big.mesh" : num_procs == 2 ? "wave-tank-big2.mesh"
                      : num_procs == 4 ? "wave-tank-big4.mesh" : "wave-tank-big8.mesh";
            par_ref_levels = a.geti("--par-ref", 0);
        }
        mesh_name = a.get("--mesh", mesh_name.c_str());
        std::vector<int> orders;
        { std::stringstream ss(a.get("--orders", "3,4")); std::string t; while (std::getline(ss, t, ',')) orders.push_back(std::atoi(t.c_str())); }
        const std::string out = a.get("--out", "");
        FILE *fout = out.empty() ? nullptr : fopen(out.c_str(), "a");
        if (fout) fprintf(fout, "# mode(0=strong,1=weak)  order  par_ref_level  ranks  dofs  runtime[s]\n");
        printf("%s\nprocs  order  par_ref  dofs  time_total[s]  time_per_step[s]\n------------------------------------------------------------\n",
               mode == 0 ? "Strong scaling test" : "Weak scaling test");
        Wave w;
        const double dt = w.T / nsteps;                                // ss.cpp:170
        for (int order : orders) {
            std::unique_ptr<Mesh> mesh(Mesh::FromName(mesh_name));
            for (int i = 0; i < par_ref_levels; i++) mesh->UniformRefinement();
            World world(num_procs, std::string(a.get("--comm", "p2p")) == "nccl");
            world.parse_options(a.get("--opt", ""));   // e.g. --opt affine=0,deterministic=1 (lpf_set_option)
            std::mutex mu;
            double max_time = 0.0;
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
                surface.Setup(w.params(0.0, false, rel_tol, max_iter), nullptr, nullptr);
                surface.SetState(state);
                double t = 0.0;
                surface.Step(t, dt);                                   // warm-up step, not timed (ss.cpp:253)
                surface.Sync();
                const auto t0 = std::chrono::steady_clock::now();
                for (int step = 0; step < nsteps; step++) surface.Step(t, dt);
                surface.Sync();
                const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                std::lock_guard<std::mutex> lk(mu);
                max_time = std::max(max_time, sec);                    // MPI_Reduce(MAX) (ss.cpp:270-272)
                dofs = d.n_true_global;
                if (myid == 0) its = surface.LastIterations();
            });
            printf("%d  %d  %d  %ld  %g  %g    (CG its last step: %d %d %d %d)\n", num_procs, order, par_ref_levels, dofs, max_time,
                   max_time / nsteps, its[0], its[1], its[2], its[3]);
            if (fout) fprintf(fout, "%d  %d  %d  %d  %ld  %g\n", mode, order, par_ref_levels, num_procs, dofs, max_time);
        }
        if (fout) fclose(fout);
    } catch (const std::exception &e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
