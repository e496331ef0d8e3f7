// C-ABI of the host mini-FEM layer (include/lpf_b200.h, "Host mini-FEM" section).
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
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
        if (order < 1 || order > LPF_MAX_ORDER) throw std::runtime_error("lpf_space_create: order must be in 1..10");
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

// |eta|_max / (H/2) on r >= a at angle phi from the direction of propagation (cylinder-exact.cpp:53-115).
// tol > 0: the reference's stopping rule verbatim (:104-110): stop when |Re(term)| of two consecutive terms is below tol,
//          the "previous term" starting at 0.
// tol < 0: robust rule with tolerance |tol|: the phi-independent coefficient of two consecutive terms is tested instead.
//          (The reference's rule carries cos(m phi) inside the tested term, so at phi = pi/2 it fires at m = 1 -- cos = 0,
//          previous term 0 -- and truncates the series to its m = 0 term; everywhere else the two rules agree to ~tol.)
double lpf_maccamy_fuchs(double k, double a, double r, double phi, double tol, int max_iter)
{
    using cd = std::complex<double>;
    const bool robust = tol < 0.0;
    const double etol = std::fabs(tol);
    const double ka = k * a, kr = k * r;
    const double J0P = -std::cyl_bessel_j(1.0, ka);
    const cd H0P(-std::cyl_bessel_j(1.0, ka), -std::cyl_neumann(1.0, ka));
    const cd H0r(std::cyl_bessel_j(0.0, kr), std::cyl_neumann(0.0, kr));
    cd E = std::cyl_bessel_j(0.0, kr) - H0r * (J0P / H0P);
    double oldterm = robust ? 1.0 : 0.0;
    for (int m = 1; m <= max_iter; m++) {
        const double JmP = 0.5 * (std::cyl_bessel_j(m - 1.0, ka) - std::cyl_bessel_j(m + 1.0, ka));
        const cd HmP(JmP, 0.5 * (std::cyl_neumann(m - 1.0, ka) - std::cyl_neumann(m + 1.0, ka)));
        if (std::abs(HmP) < 1e-14) continue;
        const double Jmr = std::cyl_bessel_j((double)m, kr);
        const cd Hmr(Jmr, std::cyl_neumann((double)m, kr));
        const double ph = m * M_PI / 2.0;
        const cd coef = 2.0 * cd(std::cos(ph), std::sin(ph)) * (Jmr - Hmr * (JmP / HmP));
        const cd term = coef * std::cos(m * phi);
        if (std::isnan(term.real())) break;
        E += term;
        const double nextterm = robust ? std::abs(coef) : term.real();
        if (std::fabs(nextterm) < etol && std::fabs(oldterm) < etol) break;
        oldterm = nextterm;
    }
    return std::abs(E);
}

int lpf_space_rim(const lpf_space *s, int wall_attr, double cx, double cy, double a, double tol,
                  int *surf_idx, double *theta, int cap)
{
    if (!s || !s->mesh || (cap > 0 && (!surf_idx || !theta))) { lpf::set_error("lpf_space_rim: null argument"); return LPF_ERR_ARG; }
    const lpf::Mesh &m = *s->mesh;
    const lpf::Partition &p = s->part;
    const int D = s->order + 1, D3 = D * D * D, P = s->order;
    std::vector<uint8_t> on_wall((size_t)m.nv, 0);
    for (int b = 0; b < m.nb(); b++)
        if (m.bdr_attr[b] == wall_attr) for (int k = 0; k < 4; k++) on_wall[m.bdr[(size_t)b * 4 + k]] = 1;
    std::vector<int> vol2surf(p.l2g.size(), -1);
    for (size_t i = 0; i < p.surf2vol.size(); i++) vol2surf[p.surf2vol[i]] = (int)i;
    std::vector<uint8_t> seen((size_t)m.nv, 0);
    int n = 0;
    for (size_t le = 0; le < p.elems.size(); le++) {
        const int ge = p.elems[le];
        for (int c = 0; c < 8; c++) {
            const int v = m.elems[(size_t)ge * 8 + lpf::kLex2Mfem[c]];
            if (!on_wall[v] || seen[v]) continue;
            const int loc = ((c & 1) ? P : 0) + D * ((((c >> 1) & 1) ? P : 0) + D * (((c >> 2) & 1) ? P : 0));
            const int si = vol2surf[p.gather[le * D3 + loc]];
            if (si < 0) continue;                                   // wall vertex below the free surface
            seen[v] = 1;
            const double dx = p.corners[le * 24 + c * 3] - cx, dy = p.corners[le * 24 + c * 3 + 1] - cy;
            if (std::fabs(std::sqrt(dx * dx + dy * dy) - a) > tol) continue;
            const double th = std::atan2(dy, dx);
            if (th < 0.0) continue;
            if (n < cap) { surf_idx[n] = si; theta[n] = th; }
            n++;
        }
    }
    return n;
}

/* free-surface faces of this rank as high-order quads: conn[nf][D*D] local surface dofs (x-fastest lexicographic
 * in the face's own (s,t) axes); returns nf (pass conn = NULL to size) */
int lpf_space_surface_quads(const lpf_space *s, int *conn, int cap_faces)
{
    if (!s) { lpf::set_error("lpf_space_surface_quads: null argument"); return LPF_ERR_ARG; }
    const lpf::Partition &p = s->part;
    const int D = s->order + 1, D3 = D * D * D, P = s->order;
    std::vector<int> g2l;
    {
        int gmax = -1;
        for (int ge : p.elems) gmax = std::max(gmax, ge);
        g2l.assign((size_t)gmax + 1, -1);
        for (size_t le = 0; le < p.elems.size(); le++) g2l[p.elems[le]] = (int)le;
    }
    std::vector<int> vol2surf(p.l2g.size(), -1);
    for (size_t i = 0; i < p.surf2vol.size(); i++) vol2surf[p.surf2vol[i]] = (int)i;
    int nf = 0;
    const std::vector<int> &sf = s->space.surf_faces;
    for (size_t f = 0; f + 1 < sf.size(); f += 2) {
        const int ge = sf[f], lf = sf[f + 1];
        if (ge >= (int)g2l.size() || g2l[ge] < 0) continue;
        const int le = g2l[ge], ax = lf / 2, fixed = (lf & 1) ? P : 0;
        if (conn && nf < cap_faces) {
            for (int t = 0; t < D; t++)
                for (int q = 0; q < D; q++) {
                    int i, j, k;
                    if (ax == 0) { i = fixed; j = q; k = t; } else if (ax == 1) { i = q; j = fixed; k = t; } else { i = q; j = t; k = fixed; }
                    conn[(size_t)nf * D * D + q + D * t] = vol2surf[p.gather[(size_t)le * D3 + i + D * (j + D * k)]];
                }
        }
        nf++;
    }
    return nf;
}

/* VTK (>= 9) node order of an order-p Lagrange quadrilateral in terms of the lexicographic lattice (i + D j):
 * 4 corners, then the edges (bottom, right, top, left; each in increasing coordinate), then the interior row by row */
static void vtk_lagrange_quad_order(int p, std::vector<int> &ord)
{
    const int D = p + 1;
    ord.clear();
    ord.push_back(0); ord.push_back(p); ord.push_back(p + D * p); ord.push_back(D * p);
    for (int i = 1; i < p; i++) ord.push_back(i);
    for (int j = 1; j < p; j++) ord.push_back(p + D * j);
    for (int i = 1; i < p; i++) ord.push_back(i + D * p);
    for (int j = 1; j < p; j++) ord.push_back(D * j);
    for (int j = 1; j < p; j++) for (int i = 1; i < p; i++) ord.push_back(i + D * j);
}

int lpf_write_surface_vtu(const lpf_space *s, const char *path, double z, int nfields, const char *const *names,
                          const double *const *fields, int high_order)
{
    (void)z;
    if (!s || !path || nfields < 0 || (nfields > 0 && (!names || !fields))) { lpf::set_error("lpf_write_surface_vtu: bad argument"); return LPF_ERR_ARG; }
    const lpf::Partition &p = s->part;
    const int P = s->order, D = P + 1, DD = D * D, D3 = D * DD;
    std::vector<int> g2l;
    {
        int gmax = -1;
        for (int ge : p.elems) gmax = std::max(gmax, ge);
        g2l.assign((size_t)gmax + 1, -1);
        for (size_t le = 0; le < p.elems.size(); le++) g2l[p.elems[le]] = (int)le;
    }
    std::vector<int> vol2surf(p.l2g.size(), -1);
    for (size_t i = 0; i < p.surf2vol.size(); i++) vol2surf[p.surf2vol[i]] = (int)i;
    // faces of this rank; like MFEM's ParaView writer every cell carries its own copy of its nodes (discontinuous
    // points), taken from the ELEMENT geometry -- so cells across a periodic seam keep their true coordinates
    lpf::Basis1D bs(P);
    std::vector<double> pts;      // [nf][DD][3]
    std::vector<int> sdof;        // [nf][DD] surface dof of each cell node
    const std::vector<int> &sf = s->space.surf_faces;
    for (size_t f = 0; f + 1 < sf.size(); f += 2) {
        const int ge = sf[f], lf = sf[f + 1];
        if (ge >= (int)g2l.size() || g2l[ge] < 0) continue;
        const int le = g2l[ge], ax = lf / 2, fixed = (lf & 1) ? P : 0;
        const double *Cn = &p.corners[(size_t)le * 24];
        for (int t = 0; t < D; t++)
            for (int q = 0; q < D; q++) {
                int i, j, k;
                if (ax == 0) { i = fixed; j = q; k = t; } else if (ax == 1) { i = q; j = fixed; k = t; } else { i = q; j = t; k = fixed; }
                sdof.push_back(vol2surf[p.gather[(size_t)le * D3 + i + D * (j + D * k)]]);
                const double x = bs.nodes[i], y = bs.nodes[j], zz = bs.nodes[k];
                double pt[3] = {0, 0, 0};
                for (int c = 0; c < 8; c++) {
                    const double w = ((c & 1) ? x : 1 - x) * (((c >> 1) & 1) ? y : 1 - y) * (((c >> 2) & 1) ? zz : 1 - zz);
                    for (int d = 0; d < 3; d++) pt[d] += w * Cn[c * 3 + d];
                }
                pts.insert(pts.end(), pt, pt + 3);
            }
    }
    const int nf = (int)(sdof.size() / DD), npts = nf * DD;
    FILE *f = fopen(path, "w");
    if (!f) { lpf::set_error(std::string("lpf_write_surface_vtu: cannot open ") + path); return LPF_ERR_ARG; }
    // high_order: one VTK_LAGRANGE_QUADRILATERAL (70) per face; otherwise p*p bilinear VTK_QUAD (9) sub-cells
    const int ncell = high_order ? nf : nf * P * P, npc = high_order ? DD : 4;
    fprintf(f, "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"2.2\" byte_order=\"LittleEndian\">\n<UnstructuredGrid>\n");
    fprintf(f, "<Piece NumberOfPoints=\"%d\" NumberOfCells=\"%d\">\n<Points>\n<DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n", npts, ncell);
    for (int i = 0; i < npts; i++) fprintf(f, "%.17g %.17g %.17g\n", pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2]);
    fprintf(f, "</DataArray>\n</Points>\n<Cells>\n<DataArray type=\"Int32\" Name=\"connectivity\" format=\"ascii\">\n");
    std::vector<int> ord;
    vtk_lagrange_quad_order(P, ord);
    for (int q = 0; q < nf; q++) {
        const int base = q * DD;
        if (high_order) { for (int k = 0; k < DD; k++) fprintf(f, "%d ", base + ord[k]); fprintf(f, "\n"); }
        else for (int j = 0; j < P; j++) for (int i = 0; i < P; i++)
            fprintf(f, "%d %d %d %d\n", base + i + D * j, base + i + 1 + D * j, base + i + 1 + D * (j + 1), base + i + D * (j + 1));
    }
    fprintf(f, "</DataArray>\n<DataArray type=\"Int32\" Name=\"offsets\" format=\"ascii\">\n");
    for (int q = 1; q <= ncell; q++) fprintf(f, "%d\n", q * npc);
    fprintf(f, "</DataArray>\n<DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">\n");
    for (int q = 0; q < ncell; q++) fprintf(f, "%d\n", high_order ? 70 : 9);
    fprintf(f, "</DataArray>\n</Cells>\n<PointData>\n");
    for (int k = 0; k < nfields; k++) {
        fprintf(f, "<DataArray type=\"Float64\" Name=\"%s\" NumberOfComponents=\"1\" format=\"ascii\">\n", names[k]);
        for (int i = 0; i < npts; i++) fprintf(f, "%.17g\n", fields[k][sdof[i]]);
        fprintf(f, "</DataArray>\n");
    }
    fprintf(f, "</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n");
    fclose(f);
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
