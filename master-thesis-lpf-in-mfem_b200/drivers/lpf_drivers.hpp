// Host-side C++ layer of the drivers: thin RAII classes over the C-ABI (include/lpf_b200.h) named after
// the MFEM objects the reference drivers use, so that drivers/*.cpp read like Solvers/*.cpp and
// Convergence_and_Scaling/*.cpp of the reference.  One std::thread per GPU stands in for one MPI rank.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "lpf_b200.h"

namespace lpfd {

inline void check(int rc, const char *what)
{
    if (rc != LPF_OK) throw std::runtime_error(std::string(what) + ": " + lpf_last_error());
}

// ---- command line: --key value pairs with defaults equal to the reference's source constants ----
struct Args {
    int argc; char **argv;
    const char *get(const char *key, const char *def) const
    {
        for (int i = 1; i + 1 < argc; i++) if (std::strcmp(argv[i], key) == 0) return argv[i + 1];
        return def;
    }
    int geti(const char *key, int def) const { return std::atoi(get(key, std::to_string(def).c_str())); }
    double getd(const char *key, double def) const { const char *v = get(key, nullptr); return v ? std::atof(v) : def; }
    bool has(const char *key) const { for (int i = 1; i < argc; i++) if (std::strcmp(argv[i], key) == 0) return true; return false; }
};

// ---- mfem::Mesh ----
class Mesh {
public:
    explicit Mesh(const std::string &file) : h_(lpf_mesh_read(file.c_str())) { if (!h_) throw std::runtime_error(lpf_last_error()); }
    Mesh(int nx, int ny, int nz, double Lx, double Ly, double H, bool periodic_x)
        : h_(lpf_mesh_make_wave_tank(nx, ny, nz, Lx, Ly, H, periodic_x)) { if (!h_) throw std::runtime_error(lpf_last_error()); }
    ~Mesh() { lpf_mesh_destroy(h_); }
    Mesh(const Mesh &) = delete;
    void UniformRefinement() { check(lpf_mesh_refine(h_, 1), "UniformRefinement"); }
    void GetBoundingBox(double lo[3], double hi[3]) const { check(lpf_mesh_bounding_box(h_, lo, hi), "GetBoundingBox"); }
    int GetNE() const { return lpf_mesh_num_elements(h_); }
    lpf_mesh *handle() const { return h_; }
    /// "wave-tank.mesh" style names map onto the generator (the reference's mesh files are its outputs)
    static Mesh *FromName(const std::string &name)
    {
        const double H = 1.0 / (2.0 * M_PI);
        if (name == "wave-tank.mesh") return new Mesh(3, 1, 1, 1.0, 0.1, H, true);
        if (name == "wave-tank-big.mesh") return new Mesh(32, 2, 8, 1.0, 0.1, H, true);
        if (name == "wave-tank-big2.mesh") return new Mesh(64, 2, 8, 1.0, 0.1, H, true);
        if (name == "wave-tank-big4.mesh") return new Mesh(64, 2, 16, 1.0, 0.1, H, true);
        if (name == "wave-tank-big8.mesh") return new Mesh(128, 2, 16, 1.0, 0.1, H, true);
        if (name == "wave-tank-finite.mesh") return new Mesh(36, 1, 1, 12.0, 1.0, H, false);
        return new Mesh(name);
    }
private:
    lpf_mesh *h_;
};

// ---- H1 ParFiniteElementSpace of one rank + its device context (what rhs_linear owns) ----
class RankSpace {
public:
    RankSpace(const Mesh &mesh, int order, int nranks, int rank)
        : s_(lpf_space_create(mesh.handle(), order, 2, nranks, rank))
    {
        if (!s_) throw std::runtime_error(lpf_last_error());
        check(lpf_space_desc_get(s_, &desc), "lpf_space_desc_get");
    }
    ~RankSpace() { lpf_space_destroy(s_); }
    RankSpace(const RankSpace &) = delete;
    int GetTrueVSize() const { int n = 0; for (int i = 0; i < desc.ndof; i++) n += desc.owned ? desc.owned[i] : 1; return n; }
    long GlobalTrueVSize() const { return desc.n_true_global; }
    std::vector<double> NodeCoordinates() const
    {
        std::vector<double> xyz((size_t)desc.ndof * 3);
        check(lpf_space_node_coordinates(s_, xyz.data()), "node coordinates");
        return xyz;
    }
    /// (theta, local surface dof) of the mesh vertices on the cylinder rim (cylinder-diffraction.cpp:476-496)
    std::vector<std::pair<double, int>> Rim(int wall_attr, double cx, double cy, double a, double tol) const
    {
        const int n = lpf_space_rim(s_, wall_attr, cx, cy, a, tol, nullptr, nullptr, 0);
        if (n < 0) throw std::runtime_error(lpf_last_error());
        std::vector<int> idx(n); std::vector<double> th(n);
        lpf_space_rim(s_, wall_attr, cx, cy, a, tol, idx.data(), th.data(), n);
        std::vector<std::pair<double, int>> out;
        for (int i = 0; i < n; i++) out.emplace_back(th[i], idx[i]);
        return out;
    }
    lpf_space *handle() const { return s_; }
    lpf_space_desc desc{};
private:
    lpf_space *s_;
};

// ---- wave parameters (lambda-mode, PF_linear_par_partial.cpp:298-306) ----
struct Wave {
    double H = 0.01, g = 9.81, lambda = 1.0, kh = 1.0, theta = 0.0;
    double k, cwave, T, omega, kx_dir, ky_dir;
    Wave() { finish(); }
    void finish()
    {
        k = 2.0 * M_PI / lambda;
        cwave = std::sqrt((g / k) * std::tanh(kh));
        T = lambda / cwave;
        omega = 2.0 * M_PI / T;
        kx_dir = std::cos(theta); ky_dir = std::sin(theta);
    }
    /// "period mode" of Solvers/PF_linear_serial.cpp:22-34,320-327: the wave is chosen by its period; kh from the
    /// fixed-point iteration kh <- sqrt(w^2 h / g * kh * coth(kh)), n_iter times
    void period_mode(double T_input, double h, int n_iter = 40)
    {
        const double wq = 2.0 * M_PI / T_input;
        double x = std::max(wq * wq * h / g, 1e-8);
        for (int i = 0; i < n_iter; i++) {
            const double xs = std::max(x, 1e-12);
            x = std::max(std::sqrt((wq * wq / g) * h * x * (std::cosh(xs) / std::sinh(xs))), 1e-8);
        }
        kh = x; T = T_input; omega = wq; k = kh / h; cwave = omega / k; lambda = 2.0 * M_PI / k;
        kx_dir = std::cos(theta); ky_dir = std::sin(theta);
    }
    double phase(double t, double x, double y) const { return omega * t - k * (kx_dir * x + ky_dir * y); }
    double eta(double t, double x, double y) const { return 0.5 * H * std::cos(phase(t, x, y)); }
    double phi_fs(double t, double x, double y) const { return -0.5 * H * cwave * std::cosh(kh) / std::sinh(kh) * std::sin(phase(t, x, y)); }
    // volume potential with z measured from the bottom (laplace_solver.cpp:70-81)
    double phi(double t, double x, double y, double zb) const { return -0.5 * H * cwave * std::cosh(k * zb) / std::sinh(kh) * std::sin(phase(t, x, y)); }
    double w_surface(double t, double x, double y) const { return -0.5 * H * cwave * k * std::sin(phase(t, x, y)); }
    lpf_rhs_params params(double tau, bool relax, double rel_tol, int max_iter) const
    {
        lpf_rhs_params p{};
        p.g = g; p.H = H; p.omega = omega; p.k = k; p.kx_dir = kx_dir; p.ky_dir = ky_dir; p.cwave = cwave; p.kh = kh; p.T = T;
        p.tau = tau; p.n_ramp = 3.0; p.use_relaxation = relax; p.rel_tol = rel_tol; p.abs_tol = 0.0; p.max_iter = max_iter;
        return p;
    }
};

// ---- one std::thread per GPU ("mpirun -np N") ----
class ThreadBarrier {
public:
    explicit ThreadBarrier(int n) : n_(n) {}
    void wait()
    {
        std::unique_lock<std::mutex> lk(m_);
        const int gen = gen_;
        if (++count_ == n_) { count_ = 0; gen_++; cv_.notify_all(); }
        else cv_.wait(lk, [&] { return gen != gen_; });
    }
private:
    std::mutex m_;
    std::condition_variable cv_;
    int n_, count_ = 0, gen_ = 0;
};

struct World {
    int nranks;
    bool use_nccl;                     // --comm nccl: NCCL send/recv + all-reduce instead of the peer-memory exchange
    unsigned char nccl_id[128];
    std::vector<uint64_t> raw_ptrs;    // mailboxes of all ranks (threads of one process share the address space)
    std::vector<int> devices;
    ThreadBarrier barrier;
    std::vector<std::pair<std::string, long>> options;   // lpf_set_option pairs for every context (--opt name=value,name=value)
    void parse_options(const char *spec)
    {
        std::string s(spec ? spec : "");
        size_t a = 0;
        while (a < s.size()) {
            const size_t b = s.find(',', a), e = s.find('=', a);
            if (e != std::string::npos) options.emplace_back(s.substr(a, e - a), std::atol(s.substr(e + 1, b == std::string::npos ? b : b - e - 1).c_str()));
            if (b == std::string::npos) break;
            a = b + 1;
        }
    }
    explicit World(int n, bool nccl = false) : nranks(n), use_nccl(nccl), raw_ptrs(n, 0), devices(n, 0), barrier(n)
    {
        if (n > 1 && use_nccl) check(lpf_comm_unique_id(nccl_id), "lpf_comm_unique_id");
        const int ndev = lpf_device_count();
        if (ndev < n) throw std::runtime_error("need " + std::to_string(n) + " GPUs, found " + std::to_string(ndev));
    }
    void run(const std::function<void(int)> &fn)
    {
        std::vector<std::thread> th;
        std::vector<std::string> err(nranks);
        for (int r = 0; r < nranks; r++)
            th.emplace_back([&, r] { try { fn(r); } catch (const std::exception &e) { err[r] = e.what(); fprintf(stderr, "rank %d: %s\n", r, e.what()); if (nranks > 1) std::abort(); } });
        for (auto &t : th) t.join();
        for (int r = 0; r < nranks; r++) if (!err[r].empty()) throw std::runtime_error("rank " + std::to_string(r) + ": " + err[r]);
    }
};

// ---- rhs_linear : TimeDependentOperator + RK4Solver, device resident ----
class RhsLinear {
public:
    RhsLinear(const RankSpace &sp, int device, World &world) : ns_(sp.desc.n_surf), ndof_(sp.desc.ndof)
    {
        ctx_ = lpf_create(&sp.desc, device, nullptr);
        if (!ctx_) throw std::runtime_error(std::string("lpf_create: ") + lpf_last_error());
        for (auto &o : world.options) check(lpf_set_option(ctx_, o.first.c_str(), o.second), "lpf_set_option");
        if (sp.desc.nranks > 1 && world.use_nccl) check(lpf_comm_init(ctx_, world.nccl_id), "lpf_comm_init");
        else if (sp.desc.nranks > 1) {
            // replaces MPI_COMM_WORLD inside CGSolver / GroupCommunicator: peer-memory mailboxes over NVLink
            unsigned char h[64];
            check(lpf_p2p_export(ctx_, h, &world.raw_ptrs[sp.desc.rank], &world.devices[sp.desc.rank]), "lpf_p2p_export");
            world.barrier.wait();
            check(lpf_p2p_connect(ctx_, nullptr, world.raw_ptrs.data(), world.devices.data(), 1), "lpf_p2p_connect");
            world.barrier.wait();
        }
        check(lpf_pa_setup(ctx_), "a_loc_cach->Assemble()");           // PF_linear_par_partial.cpp:118-121
        check(lpf_jacobi_setup(ctx_), "OperatorJacobiSmoother");        // :124
        state_ = (double *)lpf_dev_alloc(sizeof(double) * 2 * (ns_ ? ns_ : 1));
        if (!state_) throw std::runtime_error(lpf_last_error());
    }
    ~RhsLinear() { lpf_dev_free(state_); lpf_destroy(ctx_); }
    RhsLinear(const RhsLinear &) = delete;
    void Setup(const lpf_rhs_params &p, const double *cgen, const double *cabs) { check(lpf_rhs_setup(ctx_, &p, cgen, cabs), "lpf_rhs_setup"); }
    void SetCabsy(const double *cabsy) { check(lpf_rhs_set_cabsy(ctx_, cabsy), "lpf_rhs_set_cabsy"); }         // cylinder-diffraction.cpp:373-389
    void EnvelopeReset() { check(lpf_envelope_reset(ctx_), "lpf_envelope_reset"); }                              // eta_env = -1e300  (:412)
    void EnvelopeUpdate() { check(lpf_envelope_update(ctx_, state_), "lpf_envelope_update"); }                   // env = max(env, eta) (:421-430)
    std::vector<double> Envelope(double scale) { std::vector<double> e((size_t)ns_); check(lpf_envelope_get(ctx_, e.data(), scale), "lpf_envelope_get"); return e; }
    void SetState(const std::vector<double> &s) { if (ns_) check(lpf_memcpy_h2d(state_, s.data(), sizeof(double) * 2 * ns_), "h2d"); }
    void GetState(std::vector<double> &s) { s.resize(2 * (size_t)ns_); check(lpf_sync(ctx_), "sync"); if (ns_) check(lpf_memcpy_d2h(s.data(), state_, sizeof(double) * 2 * ns_), "d2h"); }
    void Step(double &t, double dt) { check(lpf_rk4_step(ctx_, state_, &t, dt), "ode_solver->Step"); }   // :494
    /// w~ = d(phi)/dz on the free surface for the CURRENT state at time t (rhs_linear::GetWTilde() of
    /// convergence-parallel-partial-hconv.cpp:331: the first half of rhs_linear::Mult's output without relaxation zones)
    std::vector<double> WTilde(double t)
    {
        std::vector<double> w((size_t)ns_);
        double *d = (double *)lpf_dev_alloc(sizeof(double) * 2 * (ns_ ? ns_ : 1));
        if (!d) throw std::runtime_error(lpf_last_error());
        check(lpf_rhs(ctx_, t, state_, d), "rhs_linear::Mult");
        check(lpf_sync(ctx_), "sync");
        if (ns_) check(lpf_memcpy_d2h(w.data(), d, sizeof(double) * ns_), "d2h");
        lpf_dev_free(d);
        return w;
    }
    void Sync() { check(lpf_sync(ctx_), "sync"); }
    std::vector<int> LastIterations()
    {
        lpf_pcg_info info[4]; int n = 0;
        check(lpf_last_solve_info(ctx_, info, &n), "info");
        std::vector<int> it; for (int i = 0; i < n; i++) it.push_back(info[i].iterations);
        return it;
    }
    lpf_ctx *ctx() const { return ctx_; }
    int nsurf() const { return ns_; }
    int ndof() const { return ndof_; }
private:
    lpf_ctx *ctx_ = nullptr;
    double *state_ = nullptr;
    int ns_, ndof_;
};

// ---- mfem::ParaViewDataCollection for the free-surface fields (PF_linear_par_partial.cpp:453-467, 505-514) ----
// <prefix>/<name>/Cycle000123/proc000000.vtu per rank and cycle + <prefix>/<name>/<name>.pvd listing every piece.
class ParaViewDataCollection {
public:
    ParaViewDataCollection(const std::string &name, const RankSpace &fs, int rank, int nranks, double z)
        : name_(name), fs_(fs), rank_(rank), nranks_(nranks), z_(z) {}
    void SetPrefixPath(const std::string &p) { prefix_ = p; }
    void SetLevelsOfDetail(int) {}                       // refinement of the VTK cells: the Lagrange cells carry order p already
    void SetHighOrderOutput(bool h) { high_order_ = h; }
    void RegisterField(const std::string &field, const std::vector<double> *values) { names_.push_back(field); fields_.push_back(values); }
    void SetCycle(int c) { cycle_ = c; }
    void SetTime(double t) { time_ = t; }
    void Save()
    {
        char cyc[64];
        snprintf(cyc, sizeof(cyc), "Cycle%06d", cycle_);
        const std::string dir = prefix_ + "/" + name_ + "/" + cyc;
        if (std::system(("mkdir -p '" + dir + "'").c_str()) != 0) throw std::runtime_error("cannot create " + dir);
        char piece[64];
        snprintf(piece, sizeof(piece), "proc%06d.vtu", rank_);
        std::vector<const char *> nm;
        std::vector<const double *> fl;
        for (size_t k = 0; k < names_.size(); k++) { nm.push_back(names_[k].c_str()); fl.push_back(fields_[k]->data()); }
        check(lpf_write_surface_vtu(fs_.handle(), (dir + "/" + piece).c_str(), z_, (int)nm.size(), nm.data(), fl.data(), high_order_ ? 1 : 0), "pv.Save");
        if (rank_ == 0) {
            saved_.emplace_back(time_, std::string(cyc));
            FILE *f = fopen((prefix_ + "/" + name_ + "/" + name_ + ".pvd").c_str(), "w");
            if (!f) throw std::runtime_error("cannot write the .pvd file");
            fprintf(f, "<?xml version=\"1.0\"?>\n<VTKFile type=\"Collection\" version=\"0.1\">\n<Collection>\n");
            for (auto &sv : saved_)
                for (int r = 0; r < nranks_; r++)
                    fprintf(f, "<DataSet timestep=\"%.12g\" part=\"%d\" file=\"%s/proc%06d.vtu\"/>\n", sv.first, r, sv.second.c_str(), r);
            fprintf(f, "</Collection>\n</VTKFile>\n");
            fclose(f);
        }
    }
private:
    std::string name_, prefix_ = "ParaView";
    const RankSpace &fs_;
    int rank_, nranks_, cycle_ = 0;
    double z_, time_ = 0.0;
    bool high_order_ = true;
    std::vector<std::string> names_;
    std::vector<const std::vector<double> *> fields_;
    std::vector<std::pair<double, std::string>> saved_;
};

}  // namespace lpfd
