// Minimal FUNCTIONAL stand-in for the MFEM classes that include/lpf_mfem_adapter.hpp touches.
//
// MFEM itself is not available in this image, so the adapter cannot be run against the real library
// here.  This shim implements, with the upstream signatures ([MFEM] general/array.hpp, general/table.hpp,
// general/communication.hpp, linalg/vector.hpp, linalg/operator.hpp, linalg/ode.hpp, fem/intrules.hpp,
// fem/bilininteg.hpp, fem/fespace.hpp, fem/pfespace.hpp, fem/restriction.hpp, mesh/mesh.hpp), exactly the behaviour the
// adapter relies on, so that drivers/adapter_check.cpp EXECUTES every adapter class on the GPU:
//   * Vector with host / device mirrors and MFEM's Read / Write / ReadWrite / HostRead / HostWrite semantics
//     (device pointers once Device("cuda") is constructed);
//   * IntRules.Get(Geometry::CUBE, order): tensor Gauss-Legendre rule, n = order/2 + 1 points per direction, x fastest;
//   * Mesh / FiniteElementSpace / ParFiniteElementSpace on top of the library's host mini-FEM (lpf_mesh / lpf_space):
//     GetGeometricFactors(...)->J in MFEM's [Q^3][3][3][NE] layout (computed HERE from the trilinear corners, independent
//     of the device code), ElementRestriction::GatherMap(), GetEssentialTrueDofs, and for the parallel space the
//     GroupCommunicator / GroupTopology / Table views of the shared dofs that ParSpaceDescBuilder translates.
// Everything a real MFEM build must be re-verified against is marked [MFEM] in the adapter header.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "lpf_b200.h"

namespace mfem {

using real_t = double;
inline void mfem_error(const char *msg) { std::fprintf(stderr, "MFEM abort: %s\n", msg); std::abort(); }

class Device {
public:
    explicit Device(const char *spec) { cuda() = std::strstr(spec, "cuda") != nullptr; }
    static bool &cuda() { static bool on = false; return on; }
};

template <class T>
class Array {
public:
    Array() = default;
    explicit Array(int n) : v_(n) {}
    Array(const T *d, int n) : v_(d, d + n) {}
    int Size() const { return (int)v_.size(); }
    void SetSize(int n) { v_.resize(n); }
    T *GetData() { return v_.data(); }
    const T *GetData() const { return v_.data(); }
    const T *HostRead() const { return v_.data(); }
    T &operator[](int i) { return v_[i]; }
    const T &operator[](int i) const { return v_[i]; }
    void Append(const T &t) { v_.push_back(t); }
    Array &operator=(const T &t) { std::fill(v_.begin(), v_.end(), t); return *this; }
    T Max() const { return *std::max_element(v_.begin(), v_.end()); }
private:
    std::vector<T> v_;
};

// Host + device mirrors with validity flags (what mfem::Memory<double> does).  Non-owning views over a host pointer are
// supported as in MFEM (Vector(double *, int), Solvers/PF_linear_par_partial.cpp:136-140).
class Vector {
public:
    Vector() = default;
    explicit Vector(int n) { SetSize(n); }
    Vector(double *d, int n) : ext_(d), n_(n) {}
    Vector(const Vector &o) { *this = o; }
    ~Vector() { if (dev_) lpf_dev_free(dev_); }
    Vector &operator=(const Vector &o)
    {
        if (this == &o) return *this;
        SetSize(o.n_);
        const double *s = o.HostRead();
        std::copy(s, s + n_, HostWrite());
        return *this;
    }
    Vector &operator=(double v) { double *h = HostWrite(); std::fill(h, h + n_, v); return *this; }
    void SetSize(int n)
    {
        if (n == n_ && !ext_) return;
        if (dev_) { lpf_dev_free(dev_); dev_ = nullptr; }
        ext_ = nullptr; own_.assign(n, 0.0); n_ = n; host_ok_ = true; dev_ok_ = false;
    }
    int Size() const { return n_; }
    double *GetData() { return HostReadWrite(); }
    const double *GetData() const { return HostRead(); }
    double &operator()(int i) { return HostReadWrite()[i]; }
    const double &operator()(int i) const { return HostRead()[i]; }
    // device-or-host access, MFEM semantics
    const double *Read() const { if (!Device::cuda()) return HostRead(); to_dev(); return dev_; }
    double *Write() { if (!Device::cuda()) return HostWrite(); alloc_dev(); dev_ok_ = true; host_ok_ = false; return dev_; }
    double *ReadWrite() { if (!Device::cuda()) return HostReadWrite(); to_dev(); host_ok_ = false; return dev_; }
    const double *HostRead() const { to_host(); return host(); }
    double *HostWrite() { host_ok_ = true; dev_ok_ = false; return host(); }
    double *HostReadWrite() { to_host(); dev_ok_ = false; return host(); }
    double Norml2() const { const double *h = HostRead(); double s = 0; for (int i = 0; i < n_; i++) s += h[i] * h[i]; return std::sqrt(s); }

private:
    double *host() const { return ext_ ? ext_ : const_cast<double *>(own_.data()); }
    void alloc_dev() const
    {
        if (!dev_ && n_) { dev_ = (double *)lpf_dev_alloc(sizeof(double) * n_); if (!dev_) mfem_error(lpf_last_error()); }
    }
    void to_dev() const
    {
        alloc_dev();
        if (!dev_ok_ && n_) { if (lpf_memcpy_h2d(dev_, host(), sizeof(double) * n_)) mfem_error(lpf_last_error()); }
        dev_ok_ = true;
    }
    void to_host() const
    {
        if (!host_ok_ && n_) { if (lpf_memcpy_d2h(host(), dev_, sizeof(double) * n_)) mfem_error(lpf_last_error()); }
        host_ok_ = true;
    }
    std::vector<double> own_;
    double *ext_ = nullptr;
    int n_ = 0;
    mutable double *dev_ = nullptr;
    mutable bool host_ok_ = true, dev_ok_ = false;
};

class Operator {
public:
    explicit Operator(int s = 0) : height(s), width(s) {}
    virtual ~Operator() = default;
    int Height() const { return height; }
    int Width() const { return width; }
    virtual void Mult(const Vector &x, Vector &y) const = 0;
    virtual void AssembleDiagonal(Vector &) const { mfem_error("not supported"); }
protected:
    int height, width;
};

class Solver : public Operator {
public:
    explicit Solver(int s = 0, bool iter_mode = false) : Operator(s), iterative_mode(iter_mode) {}
    virtual void SetOperator(const Operator &op) = 0;
    bool iterative_mode;
};

class TimeDependentOperator : public Operator {
public:
    explicit TimeDependentOperator(int n = 0, double t_ = 0.0) : Operator(n), t(t_) {}
    virtual double GetTime() const { return t; }
    virtual void SetTime(const double t_) { t = t_; }
protected:
    double t;
};

class ODESolver {
public:
    virtual ~ODESolver() = default;
    virtual void Init(TimeDependentOperator &f_) { f = &f_; }
    virtual void Step(Vector &x, double &t, double &dt) = 0;
protected:
    TimeDependentOperator *f = nullptr;
};

// ---- quadrature ----
struct Geometry { enum Type { SQUARE = 3, CUBE = 5 }; };
class IntegrationPoint {
public:
    double x = 0, y = 0, z = 0, weight = 0;
    void Set3(double a, double b, double c) { x = a; y = b; z = c; }
};
class IntegrationRule {
public:
    int GetNPoints() const { return (int)pts_.size(); }
    const IntegrationPoint &IntPoint(int i) const { return pts_[i]; }
    int n1d = 0;
    std::vector<IntegrationPoint> pts_;
};
class IntegrationRules {
public:
    // tensor Gauss-Legendre rule exact to `order`: n = order/2 + 1 points per direction, index = ix + n (iy + n iz)
    const IntegrationRule &Get(int geom, int order)
    {
        if (geom != Geometry::CUBE) mfem_error("shim: only Geometry::CUBE");
        const int n = order / 2 + 1;
        auto it = cache_.find(n);
        if (it != cache_.end()) return it->second;
        std::vector<double> qp(n), qw(n);
        if (n < 3 || lpf_basis_tables(n - 2, nullptr, qp.data(), qw.data(), nullptr, nullptr, nullptr)) mfem_error("shim: quadrature rule");
        IntegrationRule ir;
        ir.n1d = n;
        for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
            IntegrationPoint ip; ip.x = qp[i]; ip.y = qp[j]; ip.z = qp[k]; ip.weight = qw[i] * qw[j] * qw[k];
            ir.pts_.push_back(ip);
        }
        return cache_[n] = ir;
    }
private:
    std::map<int, IntegrationRule> cache_;
};
static IntegrationRules IntRules;

class GeometricFactors {
public:
    enum FactorFlags { COORDINATES = 1, JACOBIANS = 2, DETERMINANTS = 4 };
    Vector J;
};

// ---- mesh ----
class ElementTransformation {
public:
    const double *corners = nullptr;       // [8][3] lexicographic
    void Transform(const IntegrationPoint &ip, Vector &pt)
    {
        double *o = pt.HostWrite();
        for (int a = 0; a < 3; a++) o[a] = 0.0;
        for (int c = 0; c < 8; c++) {
            const double w = ((c & 1) ? ip.x : 1 - ip.x) * (((c >> 1) & 1) ? ip.y : 1 - ip.y) * (((c >> 2) & 1) ? ip.z : 1 - ip.z);
            for (int a = 0; a < 3; a++) o[a] += w * corners[3 * c + a];
        }
    }
};

class Mesh {
public:
    Mesh(const char *file, int = 1, int = 1) : h_(lpf_mesh_read(file)) { if (!h_) mfem_error(lpf_last_error()); init(); }
    explicit Mesh(lpf_mesh *h) : h_(h) { if (!h_) mfem_error(lpf_last_error()); init(); }      // shim only: generator meshes
    ~Mesh() { lpf_mesh_destroy(h_); for (auto &g : geom_) delete g.second; }
    Mesh(const Mesh &) = delete;
    int Dimension() const { return 3; }
    int GetNE() const { return lpf_mesh_num_elements(h_); }
    void UniformRefinement() { if (lpf_mesh_refine(h_, 1)) mfem_error(lpf_last_error()); for (auto &g : geom_) delete g.second; geom_.clear(); }
    void GetBoundingBox(Vector &lo, Vector &hi) { lo.SetSize(3); hi.SetSize(3); lpf_mesh_bounding_box(h_, lo.HostWrite(), hi.HostWrite()); }
    ElementTransformation *GetElementTransformation(int e) { tr_.corners = lpf_mesh_corners(h_) + (size_t)e * 24; return &tr_; }
    // J(q, c, k, e) = d x_c / d xi_k at quadrature point q of element e, layout [NQ][3][3][NE] (q fastest)
    const GeometricFactors *GetGeometricFactors(const IntegrationRule &ir, int)
    {
        auto it = geom_.find(ir.n1d);
        if (it != geom_.end()) return it->second;
        const int nq = ir.GetNPoints(), ne = GetNE();
        auto *g = new GeometricFactors;
        g->J.SetSize(nq * 9 * ne);
        double *J = g->J.HostWrite();
        const double *C = lpf_mesh_corners(h_);
        for (int e = 0; e < ne; e++)
            for (int q = 0; q < nq; q++) {
                const IntegrationPoint &ip = ir.IntPoint(q);
                const double xi[3] = {ip.x, ip.y, ip.z};
                for (int c = 0; c < 3; c++)
                    for (int k = 0; k < 3; k++) {
                        double s = 0.0;
                        for (int n = 0; n < 8; n++) {
                            double w = 1.0;
                            for (int a = 0; a < 3; a++) {
                                const int bit = (n >> a) & 1;
                                w *= (a == k) ? (bit ? 1.0 : -1.0) : (bit ? xi[a] : 1.0 - xi[a]);
                            }
                            s += w * C[(size_t)e * 24 + 3 * n + c];
                        }
                        J[q + (size_t)nq * (c + 3 * (k + 3 * (size_t)e))] = s;
                    }
            }
        return geom_[ir.n1d] = g;
    }
    Array<int> bdr_attributes;
    lpf_mesh *handle() const { return h_; }
private:
    void init()
    {
        const int nb = lpf_mesh_num_bdr(h_);
        const int *at = lpf_mesh_bdr_attr(h_);
        std::set<int> s(at, at + nb);
        for (int a : s) bdr_attributes.Append(a);
    }
    lpf_mesh *h_;
    ElementTransformation tr_;
    std::map<int, GeometricFactors *> geom_;
};

// ---- spaces ----
class FiniteElement {
public:
    explicit FiniteElement(int p = 1) : p_(p) {}
    int GetOrder() const { return p_; }
    Geometry::Type GetGeomType() const { return Geometry::CUBE; }
private:
    int p_;
};
class FiniteElementCollection { public: virtual ~FiniteElementCollection() = default; int order = 1; };
class H1_FECollection : public FiniteElementCollection { public: H1_FECollection(int p, int /*dim*/) { order = p; } };

enum class ElementDofOrdering { NATIVE, LEXICOGRAPHIC };

class ElementRestriction : public Operator {
public:
    void Mult(const Vector &, Vector &) const override { mfem_error("shim: ElementRestriction::Mult is not needed by the adapter"); }
    const Array<int> &GatherMap() const { return gather_; }     // [D^3][NE], d fastest
    Array<int> gather_;
};

class Table {
public:
    int Size() const { return (int)rows_.size(); }
    int RowSize(int i) const { return (int)rows_[i].size(); }
    const int *GetRow(int i) const { return rows_[i].data(); }
    std::vector<std::vector<int>> rows_;
};

// Groups of ranks sharing dofs.  Group 0 is the local group {me}; inside a group the ranks are referred to by their
// index in the neighbour list (lproc), neighbour 0 being this rank itself ([MFEM] general/communication.hpp).
class GroupTopology {
public:
    int NGroups() const { return (int)groups_.size(); }
    int GetGroupSize(int g) const { return (int)groups_[g].size(); }
    const int *GetGroup(int g) const { return groups_[g].data(); }           // lproc indices
    int GetNeighborRank(int lproc) const { return lproc_rank_[lproc]; }
    int GetGroupMasterRank(int g) const { return master_[g]; }
    bool IAmMaster(int g) const { return master_[g] == my_rank_; }
    int MyRank() const { return my_rank_; }
    int NRanks() const { return nranks_; }
    std::vector<std::vector<int>> groups_;
    std::vector<int> lproc_rank_, master_;
    int my_rank_ = 0, nranks_ = 1;
};
class GroupCommunicator {
public:
    const GroupTopology &GetGroupTopology() const { return gtopo_; }
    const Table &GroupLDofTable() const { return group_ldof_; }
    GroupTopology gtopo_;
    Table group_ldof_;
};

class FiniteElementSpace {
public:
    FiniteElementSpace(Mesh *m, const FiniteElementCollection *fec, int nranks = 1, int rank = 0)
        : mesh_(m), fe_(fec->order), s_(lpf_space_create(m->handle(), fec->order, 2, nranks, rank))
    {
        if (!s_) mfem_error(lpf_last_error());
        if (lpf_space_desc_get(s_, &d_)) mfem_error(lpf_last_error());
        const int D3 = (fec->order + 1) * (fec->order + 1) * (fec->order + 1);
        restr_.gather_ = Array<int>(d_.gather, d_.ne * D3);
    }
    virtual ~FiniteElementSpace() { lpf_space_destroy(s_); }
    FiniteElementSpace(const FiniteElementSpace &) = delete;
    const FiniteElement *GetFE(int) const { return &fe_; }
    int GetNE() const { return d_.ne; }
    int GetVSize() const { return d_.ndof; }
    virtual int GetTrueVSize() const { return d_.ndof; }
    Mesh *GetMesh() const { return mesh_; }
    const Operator *GetElementRestriction(ElementDofOrdering) const { return &restr_; }
    // essential TRUE dofs of the boundary attributes marked in ess_bdr (the library's space was built for attribute 2)
    virtual void GetEssentialTrueDofs(const Array<int> &ess_bdr, Array<int> &list) const
    {
        if (ess_bdr.Size() < 2 || !ess_bdr[1]) mfem_error("shim: the space knows the essential dofs of boundary attribute 2 only");
        list = Array<int>(d_.ess, d_.n_ess);
    }
    const lpf_space_desc &shim_desc() const { return d_; }      // shim only: what a driver takes from ParSubMesh (surface maps)
    lpf_space *shim_handle() const { return s_; }
protected:
    Mesh *mesh_;
    FiniteElement fe_;
    lpf_space *s_;
    lpf_space_desc d_{};
    ElementRestriction restr_;
};

class ParFiniteElementSpace : public FiniteElementSpace {
public:
    ParFiniteElementSpace(Mesh *m, const FiniteElementCollection *fec, int nranks, int rank) : FiniteElementSpace(m, fec, nranks, rank)
    {
        // true-dof numbering: owned L-dofs in ascending order
        ltdof_.assign(d_.ndof, -1);
        for (int i = 0; i < d_.ndof; i++) if (!d_.owned || d_.owned[i]) ltdof_[i] = ntrue_++;
        build_groups(gc_, rank, nranks, d_.n_nbr, d_.nbr_rank, d_.nbr_offset, d_.send_dofs);
        build_groups(sgc_, rank, nranks, d_.s_n_nbr, d_.s_nbr_rank, d_.s_nbr_offset, d_.s_send);
    }
    // groups from the neighbour lists: ranks(dof) = {me} + {neighbour k : dof in its exchange list}; lowest rank is master
    static void build_groups(GroupCommunicator &gc, int rank, int nranks, int n_nbr, const int *nbr_rank, const int *nbr_offset, const int *send)
    {
        std::map<int, std::set<int>> ranks;
        for (int k = 0; k < n_nbr; k++)
            for (int i = nbr_offset[k]; i < nbr_offset[k + 1]; i++) { ranks[send[i]].insert(nbr_rank[k]); ranks[send[i]].insert(rank); }
        GroupTopology &gt = gc.gtopo_;
        gt.my_rank_ = rank; gt.nranks_ = nranks;
        gt.lproc_rank_.push_back(rank);
        std::map<int, int> lproc{{rank, 0}};
        for (int k = 0; k < n_nbr; k++) { lproc[nbr_rank[k]] = (int)gt.lproc_rank_.size(); gt.lproc_rank_.push_back(nbr_rank[k]); }
        std::map<std::set<int>, int> gid;
        gt.groups_.push_back({0}); gt.master_.push_back(rank);
        gc.group_ldof_.rows_.push_back({});
        for (auto &pr : ranks) {
            auto it = gid.find(pr.second);
            if (it == gid.end()) {
                it = gid.emplace(pr.second, (int)gt.groups_.size()).first;
                std::vector<int> g;
                for (int r : pr.second) g.push_back(lproc[r]);
                gt.groups_.push_back(g);
                gt.master_.push_back(*pr.second.begin());
                gc.group_ldof_.rows_.push_back({});
            }
            gc.group_ldof_.rows_[it->second].push_back(pr.first);
        }
    }
    int GetTrueVSize() const override { return ntrue_; }
    long GlobalTrueVSize() const { return d_.n_true_global; }
    int GetLocalTDofNumber(int ldof) const { return ltdof_[ldof]; }           // -1: owned by another rank
    long GetGlobalTDofNumber(int ldof) const { return d_.l2g ? d_.l2g[ldof] : ldof; }
    const GroupCommunicator &GroupComm() const { return gc_; }
    // shim only: the communicator of the free-surface space (with MFEM: fespace_fs.GroupComm() of the ParSubMesh space, :281-285)
    const GroupCommunicator &SurfaceGroupComm() const { return sgc_; }
    // [MFEM] ParFiniteElementSpace::GetEssentialVDofs: markers on ALL local dofs, synchronised across ranks
    void GetEssentialVDofs(const Array<int> &ess_bdr, Array<int> &marker) const
    {
        if (ess_bdr.Size() < 2 || !ess_bdr[1]) mfem_error("shim: the space knows the essential dofs of boundary attribute 2 only");
        marker.SetSize(d_.ndof); marker = 0;
        for (int i = 0; i < d_.n_ess; i++) marker[d_.ess[i]] = -1;
    }
    int GetMyRank() const { return gc_.gtopo_.my_rank_; }
    int GetNRanks() const { return gc_.gtopo_.nranks_; }
    void GetEssentialTrueDofs(const Array<int> &ess_bdr, Array<int> &list) const override
    {
        if (ess_bdr.Size() < 2 || !ess_bdr[1]) mfem_error("shim: the space knows the essential dofs of boundary attribute 2 only");
        list = Array<int>();
        for (int i = 0; i < d_.n_ess; i++) if (ltdof_[d_.ess[i]] >= 0) list.Append(ltdof_[d_.ess[i]]);
    }
private:
    std::vector<int> ltdof_;
    int ntrue_ = 0;
    GroupCommunicator gc_, sgc_;
};

class BilinearFormIntegrator {
public:
    virtual ~BilinearFormIntegrator() = default;
    virtual void AssemblePA(const FiniteElementSpace &) { mfem_error("not supported"); }
    virtual void AddMultPA(const Vector &, Vector &) const { mfem_error("not supported"); }
    virtual void AssembleDiagonalPA(Vector &) { mfem_error("not supported"); }
};

}  // namespace mfem
