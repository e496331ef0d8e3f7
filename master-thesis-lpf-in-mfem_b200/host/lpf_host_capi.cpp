// C-ABI of the host mini-FEM layer (include/lpf_b200.h, "Host mini-FEM" section).
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>

#include "../../include/lpf_b200.h"
#include "lpf_common.hpp"
#include "lpf_host.hpp"

struct lpf_mesh {
    lpf::Mesh m;
};

struct lpf_space {
    int order = 0;
    lpf::H1Space space;      // global space (numbering of the whole mesh)
    lpf::Partition part;     // this rank's view
    std::vector<double> xyz; // lazily built node coordinates of local dofs
    const lpf::Mesh *mesh = nullptr;
};

namespace lpf {
thread_local std::string g_last_error;
void set_error(const std::string &s) { g_last_error = s; }
}  // namespace lpf

extern "C" {

const char *lpf_last_error(void) { return lpf::g_last_error.c_str(); }
int lpf_version(void) { return LPF_B200_VERSION; }

lpf_mesh *lpf_mesh_read(const char *path)
{
    try {
        if (!path) throw std::runtime_error("lpf_mesh_read: null path");
        auto *m = new lpf_mesh;
        m->m = lpf::Mesh::Read(path);
        return m;
    } catch (const std::exception &e) { lpf::set_error(e.what()); return nullptr; }
}

lpf_mesh *lpf_mesh_make_wave_tank(int nx, int ny, int nz, double Lx, double Ly, double H, int periodic_x)
{
    try {
        if (nx < 1 || ny < 1 || nz < 1 || !(Lx > 0) || !(Ly > 0) || !(H > 0))
            throw std::runtime_error("lpf_mesh_make_wave_tank: non-positive size");
        auto *m = new lpf_mesh;
        m->m = lpf::Mesh::MakeWaveTank(nx, ny, nz, Lx, Ly, H, periodic_x != 0);
        return m;
    } catch (const std::exception &e) { lpf::set_error(e.what()); return nullptr; }
}

int lpf_mesh_refine(lpf_mesh *m, int levels)
{
    if (!m || levels < 0) { lpf::set_error("lpf_mesh_refine: bad argument"); return LPF_ERR_ARG; }
    try { for (int i = 0; i < levels; i++) m->m.UniformRefinement(); }
    catch (const std::exception &e) { lpf::set_error(e.what()); return LPF_ERR_ARG; }
    return LPF_OK;
}

int lpf_mesh_perturb(lpf_mesh *m, double amp)
{
    if (!m) { lpf::set_error("lpf_mesh_perturb: null mesh"); return LPF_ERR_ARG; }
    m->m.Perturb(amp);
    return LPF_OK;
}

int lpf_mesh_num_elements(const lpf_mesh *m) { return m ? m->m.ne() : LPF_ERR_ARG; }
int lpf_mesh_num_vertices(const lpf_mesh *m) { return m ? m->m.nv : LPF_ERR_ARG; }
int lpf_mesh_num_bdr(const lpf_mesh *m) { return m ? m->m.nb() : LPF_ERR_ARG; }
int lpf_mesh_bounding_box(const lpf_mesh *m, double lo[3], double hi[3])
{
    if (!m || !lo || !hi) { lpf::set_error("lpf_mesh_bounding_box: null argument"); return LPF_ERR_ARG; }
    m->m.GetBoundingBox(lo, hi);
    return LPF_OK;
}
const double *lpf_mesh_corners(const lpf_mesh *m) { return m ? m->m.corners.data() : nullptr; }
const int *lpf_mesh_elements(const lpf_mesh *m) { return m ? m->m.elems.data() : nullptr; }
const int *lpf_mesh_bdr(const lpf_mesh *m) { return m ? m->m.bdr.data() : nullptr; }
const int *lpf_mesh_bdr_attr(const lpf_mesh *m) { return m ? m->m.bdr_attr.data() : nullptr; }
void lpf_mesh_destroy(lpf_mesh *m) { delete m; }

lpf_space *lpf_space_create(const lpf_mesh *m, int order, int ess_attr, int nranks, int rank)
{
    try {
        if (!m) throw std::runtime_error("lpf_space_create: null mesh");
        if (order < 1 || order > LPF_MAX_ORDER) throw std::runtime_error("lpf_space_create: order must be in 1..8");
        auto s = std::make_unique<lpf_space>();
        s->order = order;
        s->mesh = &m->m;
        s->space = lpf::H1Space(m->m, order, ess_attr);
        s->part = lpf::Partition(m->m, s->space, nranks, rank);
        if (nranks > 1) {   // the global tables are no longer needed once the local view exists
            std::vector<int>().swap(s->space.gather);
        }
        return s.release();
    } catch (const std::exception &e) { lpf::set_error(e.what()); return nullptr; }
}

void lpf_space_destroy(lpf_space *s) { delete s; }

int lpf_basis_tables(int order, double *nodes, double *qpts, double *qwts, double *B, double *G, double *Dhat)
{
    if (order < 1 || order > 16) { lpf::set_error("lpf_basis_tables: bad order"); return LPF_ERR_ARG; }
    lpf::Basis1D b(order);
    if (nodes) std::memcpy(nodes, b.nodes.data(), sizeof(double) * b.D);
    if (qpts) std::memcpy(qpts, b.qpts.data(), sizeof(double) * b.Q);
    if (qwts) std::memcpy(qwts, b.qwts.data(), sizeof(double) * b.Q);
    if (B) std::memcpy(B, b.B.data(), sizeof(double) * b.Q * b.D);
    if (G) std::memcpy(G, b.G.data(), sizeof(double) * b.Q * b.D);
    if (Dhat) std::memcpy(Dhat, b.Dhat.data(), sizeof(double) * b.D * b.D);
    return LPF_OK;
}

int lpf_space_desc_get(const lpf_space *s, lpf_space_desc *d)
{
    if (!s || !d) { lpf::set_error("lpf_space_desc_get: null argument"); return LPF_ERR_ARG; }
    const lpf::Partition &p = s->part;
    std::memset(d, 0, sizeof(*d));
    d->order = s->order;
    d->ne = (int)p.elems.size();
    d->ndof = (int)p.l2g.size();
    d->corners = p.corners.data();
    d->jac = nullptr;
    d->jinv_z = nullptr;
    d->gather = p.gather.data();
    d->n_ess = (int)p.ess.size();
    d->ess = p.ess.data();
    d->n_surf = (int)p.surf2vol.size();
    d->surf2vol = p.surf2vol.data();
    d->surf_xy = p.surf_xy.data();
    d->n_surf_elems = (int)p.surf_elems.size();
    d->surf_elems = p.surf_elems.data();
    d->surf_mult = p.surf_mult.data();
    d->nranks = p.nranks;
    d->rank = p.rank;
    d->owned = p.owned.data();
    d->surf_owned = p.surf_owned.data();
    d->n_nbr = (int)p.nbr_rank.size();
    d->nbr_rank = p.nbr_rank.data();
    d->nbr_offset = p.nbr_offset.data();
    d->send_dofs = p.send_dofs.data();
    d->n_shared = (int)p.shared_dofs.size();
    d->shared_dofs = p.shared_dofs.data();
    d->red_off = p.red_off.data();
    d->red_src = p.red_src.data();
    d->s_n_nbr = (int)p.s_nbr_rank.size();
    d->s_nbr_rank = p.s_nbr_rank.data();
    d->s_nbr_offset = p.s_nbr_offset.data();
    d->s_send = p.s_send.data();
    d->s_n_shared = (int)p.s_shared.size();
    d->s_shared = p.s_shared.data();
    d->s_red_off = p.s_red_off.data();
    d->s_red_src = p.s_red_src.data();
    d->n_true_global = p.n_true_global;
    d->n_surf_global = p.n_surf_global;
    d->l2g = p.l2g.data();
    d->surf_g = p.surf_g.data();
    return LPF_OK;
}

int lpf_space_node_coordinates(const lpf_space *s, double *xyz)
{
    if (!s || !xyz) { lpf::set_error("lpf_space_node_coordinates: null argument"); return LPF_ERR_ARG; }
    const lpf::Partition &p = s->part;
    const int D = s->order + 1, D3 = D * D * D;
    lpf::Basis1D bs(s->order);
    const int nel = (int)p.elems.size();
    for (int e = 0; e < nel; e++) {
        const double *C = &p.corners[(size_t)e * 24];
        for (int k = 0; k < D; k++)
            for (int j = 0; j < D; j++)
                for (int i = 0; i < D; i++) {
                    const double x = bs.nodes[i], y = bs.nodes[j], z = bs.nodes[k];
                    double pt[3] = {0, 0, 0};
                    for (int c = 0; c < 8; c++) {
                        const double w = ((c & 1) ? x : 1 - x) * (((c >> 1) & 1) ? y : 1 - y) * (((c >> 2) & 1) ? z : 1 - z);
                        for (int a = 0; a < 3; a++) pt[a] += w * C[c * 3 + a];
                    }
                    const int l = p.gather[(size_t)e * D3 + i + D * (j + D * k)];
                    for (int a = 0; a < 3; a++) xyz[(size_t)l * 3 + a] = pt[a];
                }
    }
    return LPF_OK;
}

}  // extern "C"
