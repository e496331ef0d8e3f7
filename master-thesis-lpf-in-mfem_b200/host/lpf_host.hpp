// Host-side mini-FEM layer: the minimum subset of MFEM's mesh / finite-element-space set-up the
// reference drivers use before they reach the hot path (SURVEY.md 7 step 1, 8a "set-up" rows).
//
//   lpf::Mesh        <- mfem::Mesh(file,1,1), Mesh::MakeCartesian3D/MakePeriodic (Meshes/wave_tank.cpp:13-47,
//                       Meshes/wave-tank-finite.cpp:10-45), UniformRefinement, GetBoundingBox
//                       (Solvers/PF_linear_par_partial.cpp:263-266,292-293)
//   lpf::H1Space     <- H1_FECollection(order,3) + (Par)FiniteElementSpace + GetEssentialTrueDofs
//                       + ParSubMesh::CreateFromBoundary surface space (:276-285,:407-412)
//   lpf::Partition   <- ParMesh(MPI_COMM_WORLD, mesh) element partition + GroupCommunicator shared-dof
//                       groups (:268), re-designed as a rank-ordered halo-sum plan
//
// Nothing here touches the GPU.  Written from scratch; MFEM is not available in this image.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

namespace lpf {

// lexicographic corner c = cx + 2 cy + 4 cz  ->  MFEM/Gmsh hex vertex number
static constexpr int kLex2Mfem[8] = {0, 1, 3, 2, 4, 5, 7, 6};

struct Mesh {
    int nv = 0;
    std::vector<int> elems;        // [ne][8] vertex ids, MFEM ordering
    std::vector<double> corners;   // [ne][8][3] trilinear corner coordinates, lexicographic corners
    std::vector<int> bdr;          // [nb][4] vertex ids, cyclic
    std::vector<int> bdr_attr;     // [nb]

    int ne() const { return (int)(elems.size() / 8); }
    int nb() const { return (int)bdr_attr.size(); }

    static Mesh Read(const std::string &path);                 // MFEM mesh v1.0 or Gmsh 2.2 ASCII
    static Mesh MakeWaveTank(int nx, int ny, int nz, double Lx, double Ly, double H, bool periodic_x);
    void UniformRefinement();
    void Perturb(double amp);                                   // smooth seam-periodic displacement
    void GetBoundingBox(double lo[3], double hi[3]) const;
};

// 1-D tables (Gauss-Lobatto nodes, Gauss-Legendre rule with Q = p + 2 points, Lagrange B/G, collocation D)
struct Basis1D {
    int p = 0, D = 0, Q = 0;
    std::vector<double> nodes, qpts, qwts;
    std::vector<double> B, G;      // [Q][D] row-major
    std::vector<double> Dhat;      // [D][D]: Dhat[a*D+m] = l_m'(x_a)
    explicit Basis1D(int order = 1);
};

struct H1Space {
    int order = 0, D = 0;
    int ne = 0;                         // elements
    int ndof = 0;                       // dofs (first-appearance, element-major numbering)
    std::vector<int> gather;            // [ne][D^3] lexicographic element dof -> dof (ElementRestriction)
    std::vector<int> ess;               // essential dofs (closure of boundary attribute ess_attr), sorted
    std::vector<int> surf2vol;          // surface dof -> volume dof, first-appearance order
    std::vector<double> surf_xy;        // [ns][2] coordinates of surface dofs
    std::vector<int> surf_elems;        // elements that hold at least one surface dof
    std::vector<int> surf_faces;        // [nf][2] (element, local face 2*axis+side) of the boundary faces carrying ess_attr
    H1Space() = default;
    H1Space(const Mesh &mesh, int order, int ess_attr = 2);
    void NodeCoordinates(const Mesh &mesh, std::vector<double> &xyz) const;   // [ndof][3]
};

// Element partition of (mesh, space) into nranks sub-domains + everything rank `rank` needs.
struct Partition {
    int nranks = 1, rank = 0;
    std::vector<int> elem_rank;         // [global ne] owner rank of each element (RCB on centroids)
    // local view
    std::vector<int> elems;             // local element -> global element
    std::vector<int> l2g;               // local dof -> global dof
    std::vector<int> gather;            // [nloc][D^3] local dof ids
    std::vector<double> corners;        // [nloc][8][3]
    std::vector<int> ess;               // local essential dofs (decided from GLOBAL topology)
    std::vector<int> surf2vol;          // local surface dof -> local volume dof
    std::vector<int> surf_g;            // local surface dof -> global surface dof
    std::vector<double> surf_xy;
    std::vector<int> surf_elems;        // local elements with a surface dof
    std::vector<int> surf_mult;         // [nsurf local] GLOBAL number of elements touching the dof
    std::vector<uint8_t> owned;         // [nlocal dofs] 1 if this rank is the lowest rank sharing the dof
    std::vector<uint8_t> surf_owned;    // [nsurf local]
    // halo plan (neighbour exchange of partial sums on shared dofs)
    std::vector<int> nbr_rank;          // neighbour ranks, ascending
    std::vector<int> nbr_offset;        // [n_nbr+1] offsets into send_dofs / recv buffer
    std::vector<int> send_dofs;         // local dof ids, per neighbour, ascending global id
    // rank-ordered reduction: for shared local dof shared_dofs[i], sum entries
    // red_src[red_off[i] .. red_off[i+1]) in order; entry -1 = own partial, else index into recv buffer
    std::vector<int> shared_dofs, red_off, red_src;
    // same plan restricted to surface dofs (indices into the local surface vector)
    std::vector<int> s_nbr_rank, s_nbr_offset, s_send, s_shared, s_red_off, s_red_src;
    long n_true_global = 0;             // global number of true dofs
    long n_surf_global = 0;

    Partition() = default;
    Partition(const Mesh &mesh, const H1Space &space, int nranks, int rank);
};

std::vector<int> PartitionElementsRCB(const Mesh &mesh, int nparts);

}  // namespace lpf
