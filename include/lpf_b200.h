/* lpf_b200.h -- C-ABI of the B200-native LPF Laplace hot path.
 *
 * Drop-in boundary for the hot path of hirschjulien/Master-Thesis-LPF-in-MFEM: the reference has no
 * FFI of its own -- the path sits behind MFEM's virtual interfaces.  Each entry point
 * below names the MFEM interface / reference call site (file:line relative to the reference repo) it
 * replaces; include/lpf_mfem_adapter.hpp shows the thin mfem::Operator / BilinearFormIntegrator /
 * Solver / TimeDependentOperator classes a maintainer would add on top.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a negative
 * lpf_status on error (message via lpf_last_error(), thread-local); nothing throws or aborts across
 * the boundary.  `double*` arguments named *_dev are DEVICE pointers owned by the caller, *_host are
 * host pointers.  All work is enqueued on the context's stream; one context per GPU; not re-entrant.
 * Vectors are "L-vectors": local dofs of this rank with consistent copies of shared dofs (serial:
 * L == T).  E-vectors are [ne][D^3], x fastest (MFEM ElementDofOrdering::LEXICOGRAPHIC).
 */
#ifndef LPF_B200_H
#define LPF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LPF_B200_VERSION 200

typedef enum {
    LPF_OK = 0,
    LPF_ERR_ARG = -1,        /* invalid argument */
    LPF_ERR_CUDA = -2,       /* CUDA runtime error (no device, launch failure, ...) */
    LPF_ERR_STATE = -3,      /* call order violated (e.g. apply before pa_setup) */
    LPF_ERR_IO = -4,         /* mesh file / format error */
    LPF_ERR_COMM = -5,       /* NCCL error / library not found */
    LPF_ERR_UNSUPPORTED = -6 /* order out of range etc. */
} lpf_status;

const char *lpf_last_error(void);
int lpf_version(void);

/* ================================================================================================
 * Host mini-FEM (stand-in for the MFEM set-up objects the drivers build before the hot path).
 * ============================================================================================== */
typedef struct lpf_mesh lpf_mesh;     /* mfem::Mesh             Solvers/PF_linear_par_partial.cpp:263-266 */
typedef struct lpf_space lpf_space;   /* H1 (Par)FiniteElementSpace + surface ParSubMesh space  :276-285 */

/* Mesh(file,1,1): MFEM mesh v1.0 (hex, optional L2_T1_3D_P1 nodes) or Gmsh 2.2 ASCII (hex + quads). */
lpf_mesh *lpf_mesh_read(const char *path);
/* Meshes/wave_tank.cpp:13-47 (periodic_x=1) and Meshes/wave-tank-finite.cpp:10-45 (periodic_x=0). */
lpf_mesh *lpf_mesh_make_wave_tank(int nx, int ny, int nz, double Lx, double Ly, double H, int periodic_x);
int lpf_mesh_refine(lpf_mesh *m, int levels);              /* Mesh::UniformRefinement  :266 */
int lpf_mesh_perturb(lpf_mesh *m, double amp);             /* synthetic non-affine copy */
int lpf_mesh_num_elements(const lpf_mesh *m);
int lpf_mesh_num_vertices(const lpf_mesh *m);
int lpf_mesh_num_bdr(const lpf_mesh *m);
int lpf_mesh_bounding_box(const lpf_mesh *m, double lo[3], double hi[3]);   /* GetBoundingBox :292 */
const double *lpf_mesh_corners(const lpf_mesh *m);         /* [ne][8][3] lexicographic corners */
const int *lpf_mesh_elements(const lpf_mesh *m);           /* [ne][8] vertex ids (MFEM order) */
const int *lpf_mesh_bdr(const lpf_mesh *m);                /* [nb][4] */
const int *lpf_mesh_bdr_attr(const lpf_mesh *m);           /* [nb] */
void lpf_mesh_destroy(lpf_mesh *m);

/* H1_FECollection(order,3) + ParFiniteElementSpace on the part of `m` owned by `rank` of `nranks`
 * (ParMesh(MPI_COMM_WORLD, mesh) :268, element partition by recursive coordinate bisection), with
 * essential dofs = closure of boundary attribute ess_attr (GetEssentialTrueDofs :407-412) and the
 * free-surface trace space (ParSubMesh::CreateFromBoundary :279-285). */
lpf_space *lpf_space_create(const lpf_mesh *m, int order, int ess_attr, int nranks, int rank);
void lpf_space_destroy(lpf_space *s);

/* 1-D tables of order p (GLL nodes, Gauss-Legendre p+2 rule, B/G [Q][D], Dhat [D][D]); any output may be NULL */
int lpf_basis_tables(int order, double *nodes, double *qpts, double *qwts, double *B, double *G, double *Dhat);

/* Plain descriptor of one rank's space -- everything the device context needs.  An MFEM adapter fills
 * it from ElementRestriction::GatherMap(), GeometricFactors, ess_tdof_list and GroupCommunicator
 * (SURVEY.md 8b "Where an MFEM adapter gets its inputs"); lpf_space_desc_get() fills it from lpf_space. */
typedef struct {
    int order;                 /* p; D = p+1, Q = p+2 */
    int ne;                    /* local elements */
    int ndof;                  /* local L-dofs */
    const double *corners;     /* [ne][8][3] trilinear corners (lexicographic) or NULL if `jac` given */
    const double *jac;         /* optional GeometricFactors::J layout [Q^3][3][3][ne] (q fastest) */
    const double *jinv_z;      /* optional [ne][D^3][3]: third column of J^-1 at the element nodes (GetDerivative
                                  geometry); required for the surface RHS when `corners` is NULL */
    const int *gather;         /* [ne][D^3] element dof -> L-dof */
    int n_ess;
    const int *ess;            /* essential L-dofs */
    int n_surf;
    const int *surf2vol;       /* surface dof -> L-dof                       (ParSubMesh::Transfer map) */
    const double *surf_xy;     /* [n_surf][2] node coordinates of the surface dofs */
    int n_surf_elems;
    const int *surf_elems;     /* local elements holding >= 1 surface dof   (GetDerivative scope) */
    const int *surf_mult;      /* [n_surf] global #elements touching the dof (GetDerivative divisor) */
    /* ---- parallel part (all NULL / 0 when nranks == 1) ---- */
    int nranks, rank;
    const uint8_t *owned;      /* [ndof] 1 if this rank owns the dof (counts in dot products) */
    const uint8_t *surf_owned; /* [n_surf] */
    int n_nbr;
    const int *nbr_rank;       /* [n_nbr] ascending */
    const int *nbr_offset;     /* [n_nbr+1] */
    const int *send_dofs;      /* [nbr_offset[n_nbr]] L-dofs exchanged with each neighbour (same order on both sides) */
    int n_shared;
    const int *shared_dofs;    /* [n_shared] */
    const int *red_off;        /* [n_shared+1] */
    const int *red_src;        /* rank-ordered sources: -1 = own partial, else index into the receive buffer */
    int s_n_nbr;               /* same plan on the surface vector */
    const int *s_nbr_rank, *s_nbr_offset, *s_send;
    int s_n_shared;
    const int *s_shared, *s_red_off, *s_red_src;
    long n_true_global;        /* global number of true dofs (reported as `dofs`, ss.cpp:274-276) */
    long n_surf_global;
    const int *l2g;            /* [ndof] local -> global dof (diagnostics / gathers), may be NULL */
    const int *surf_g;         /* [n_surf] local surface dof -> global surface dof, may be NULL */
} lpf_space_desc;

int lpf_space_desc_get(const lpf_space *s, lpf_space_desc *out);   /* pointers stay owned by `s` */
/* f1  cylinder rim on the free surface (Solvers/cylinder-diffraction.cpp:449-505): the mesh VERTICES that lie on
 * a boundary face of attribute wall_attr AND on the free surface, with |r - a| <= tol around (cx, cy) and
 * theta = atan2(dy, dx) >= 0.  Writes up to `cap` (local surface dof, theta) pairs of this rank, returns the
 * number found (or < 0 on error).  The caller gathers the ranks, sorts by theta and drops duplicates. */
int lpf_space_rim(const lpf_space *s, int wall_attr, double cx, double cy, double a, double tol,
                  int *surf_idx, double *theta, int cap);
/* f2  MacCamy-Fuchs linear diffraction by a vertical circular cylinder: |eta|_max / (H/2) at (r >= a, phi), phi
 * measured from the direction of propagation; the series of Solvers/cylinder-exact.cpp:53-115 (tol 1e-10, <= 400
 * terms there) with std::cyl_bessel_j / std::cyl_neumann in place of Boost.Math.  tol > 0: the reference's stopping rule
 * verbatim (:104-110).  tol < 0: tolerance |tol| with a robust rule (the reference's rule truncates the series to one
 * term at phi = pi/2, where cos(m phi) makes the tested term vanish at m = 1). */
double lpf_maccamy_fuchs(double k, double a, double r, double phi, double tol, int max_iter);
/* f3  free-surface faces of this rank as order-p quads: conn[nf][(p+1)^2] local surface dofs, lexicographic in the
 * face's own axes (what ParaViewDataCollection writes for mesh_fs, :453-467).  conn = NULL: returns the count. */
int lpf_space_surface_quads(const lpf_space *s, int *conn, int cap_faces);
/* f3  one piece of the ParaView output of the free-surface fields (what pv_fs.Save() writes per rank and cycle,
 * Solvers/PF_linear_par_partial.cpp:453-467,505-514): an ASCII .vtu with the faces as
 * VTK_LAGRANGE_QUADRILATERAL cells of order p (high_order = 1, SetHighOrderOutput(true)) or p*p bilinear sub-quads
 * (0).  As in MFEM's writer every cell carries its own copy of its nodes (coordinates from the element geometry, so
 * periodic seams are drawn correctly); nfields point-data arrays are sampled from fields[k][n_surf]. `z` is unused. */
int lpf_write_surface_vtu(const lpf_space *s, const char *path, double z, int nfields, const char *const *names,
                          const double *const *fields, int high_order);
/* global (unpartitioned) node coordinates of the L-dofs of this rank: xyz[ndof][3] */
int lpf_space_node_coordinates(const lpf_space *s, double *xyz_host);

/* ================================================================================================
 * Device context
 * ============================================================================================== */
typedef struct lpf_ctx lpf_ctx;

/* Creates the per-GPU context and uploads the descriptor.  `stream` is a cudaStream_t (NULL = a new
 * non-blocking stream owned by the context).  Fails loudly (NULL + lpf_last_error) without a GPU. */
lpf_ctx *lpf_create(const lpf_space_desc *desc, int device, void *stream);
void lpf_destroy(lpf_ctx *ctx);
void *lpf_stream(lpf_ctx *ctx);
int lpf_sync(lpf_ctx *ctx);                                   /* cudaStreamSynchronize */
int lpf_ndof(const lpf_ctx *ctx);                            /* local L-dofs            (pfes.GetVSize())     */
int lpf_nsurf(const lpf_ctx *ctx);                           /* local surface L-dofs                          */
/* T-vectors <-> L-vectors.  MFEM's solver-level vectors (FormLinearSystem, CGSolver, the ODE state) are T-vectors: the dofs
 * a rank OWNS (pfes.GetTrueVSize()); this library computes on L-vectors (all local dofs, consistent copies of shared ones).
 * which = 0: volume space, 1: free-surface space.  True dof t = the t-th owned L-dof in ascending order.  Serial: T == L.
 *   lpf_prolong   P   (ConformingProlongationOperator::Mult / GroupCommunicator::Bcast)
 *   lpf_restrict  R   (the owned entries; GetTrueDofs) */
int lpf_ntrue(const lpf_ctx *ctx, int which);
int lpf_prolong(lpf_ctx *ctx, int which, const double *xT_dev, double *xL_dev);
int lpf_restrict(lpf_ctx *ctx, int which, const double *xL_dev, double *xT_dev);

/* Multi-GPU plumbing: rank 0 calls lpf_comm_unique_id and ships the 128 bytes to the other ranks
 * (MPI_Bcast / torch.distributed); every rank then calls lpf_comm_init.  Replaces the MPI communicator
 * inside ParFiniteElementSpace::GroupComm() and CGSolver(MPI_COMM_WORLD) (:157). */
int lpf_comm_unique_id(void *id128);
int lpf_comm_init(lpf_ctx *ctx, const void *id128);
/* Peer-memory exchange over NVLink (csrc/p2p.cuh): halo-sums and the PCG's scalar all-reduces are done by our
 * own kernels with stores into the neighbours' mailboxes -- no NCCL call inside the solver, a few us per
 * exchange.  Every rank exports its mailbox (64-byte CUDA IPC handle; raw pointer + device for ranks that are
 * threads of ONE process), the host gathers the nranks handles (MPI_Allgather / torch.distributed), then every
 * rank connects.  Once connected the context no longer needs lpf_comm_init.  Ranks <= 16, one node. */
int lpf_p2p_export(lpf_ctx *ctx, void *handle64, uint64_t *raw_ptr, int *device);
int lpf_p2p_connect(lpf_ctx *ctx, const void *handles /* [nranks][64] */, const uint64_t *raw_ptrs /* [nranks] */,
                    const int *devices /* [nranks] */, int same_process);
int lpf_p2p_error(lpf_ctx *ctx);                              /* 1 if a bounded flag wait timed out */

/* a1  ParBilinearForm::Assemble (PARTIAL) -> DiffusionIntegrator::AssemblePA         (:118-121) */
int lpf_pa_setup(lpf_ctx *ctx);
/* export q-data in MFEM's pa_data layout [Q^3][6][ne] (q fastest) for verification */
int lpf_pa_qdata(lpf_ctx *ctx, double *qdata_dev);
/* a8  DiffusionIntegrator::AddMultPA(x, y): y_E += B^T G^T D (G,B) x_E */
int lpf_pa_apply_E(lpf_ctx *ctx, const double *xE_dev, double *yE_dev);
/* a7+a8  PABilinearFormExtension::Mult: y_L = G^T A_E G x_L on this rank's elements, NO halo-sum
 * (y is overwritten). */
int lpf_apply_L(lpf_ctx *ctx, const double *x_dev, double *y_dev);
/* a5+a6  ConstrainedOperator(RAPOperator(P, A, P))::Mult: z = x, z[ess] = 0, y = A z (+ halo-sum over
 * ranks), y[ess] = x[ess].  (:155; used by CGSolver and EliminateRHS) */
int lpf_apply_T(lpf_ctx *ctx, const double *x_dev, double *y_dev);
/* same call with HOST buffers (copies inside): the end-to-end path of a host-memory MFEM build */
int lpf_apply_T_host(lpf_ctx *ctx, const double *x_host, double *y_host);
/* a2  BilinearForm::AssembleDiagonal (DiffusionIntegrator::AssembleDiagonalPA + G^T + P^T) */
int lpf_diag(lpf_ctx *ctx, double *diag_dev);
/* a2  DiffusionIntegrator::AssembleDiagonalPA(Vector &diag): the E-vector diagonal [ne][D^3], ACCUMULATED into diagE_dev
 * (MFEM's semantics; the form then applies the restriction-transpose and P^T).  This is the call
 * OperatorJacobiSmoother(*a_loc_cach, ess_tdof) reaches through the integrator                        (:124) */
int lpf_pa_diag_E(lpf_ctx *ctx, double *diagE_dev);
/* a2  OperatorJacobiSmoother(a, ess_tdof): dinv = 1/diag, dinv[ess] = 1              (:124) */
int lpf_jacobi_setup(lpf_ctx *ctx);
int lpf_jacobi_dinv(lpf_ctx *ctx, double *dinv_dev);

typedef struct {
    int iterations;       /* CGSolver::GetNumIterations */
    int converged;        /* CGSolver::GetConverged */
    double final_norm;    /* sqrt((r, M r)) at exit */
    double initial_norm;
    int applies;          /* operator applies spent (incl. EliminateRHS when called through solve) */
} lpf_pcg_info;

/* a4  CGSolver::Mult(B, X) with OperatorJacobiSmoother, iterative_mode = true (X is the initial guess),
 * stop on (r,Mr) <= max(rel^2 (r0,Mr0), abs^2)                                      (:157-164) */
int lpf_pcg(lpf_ctx *ctx, const double *B_dev, double *X_dev, double rel_tol, double abs_tol, int max_iter,
            lpf_pcg_info *info);
/* a3+a4+a10  FormLinearSystem(ess, phi, b=0) + CG + RecoverFEMSolution: on entry phi_dev carries the
 * essential values on ess dofs (interior ignored, zeroed: copy_interior = 0); on exit the solution. */
int lpf_laplace_solve(lpf_ctx *ctx, double *phi_dev, double rel_tol, double abs_tol, int max_iter,
                      lpf_pcg_info *info);
/* a11  GridFunction::GetDerivative(1, 2, w) restricted to the free surface: wt[s] = averaged nodal
 * d(phi)/dz at surface dof s                                                        (:169-170) */
int lpf_surface_dz(lpf_ctx *ctx, const double *phi_dev, double *wt_dev);

/* wave + relaxation-zone parameters of rhs_linear (:36-112, :298-306, :415-447, :470) */
typedef struct {
    double g, H, omega, k, kx_dir, ky_dir, cwave, kh, T;
    double tau;            /* relaxation time scale (= dt in the reference) */
    double n_ramp;         /* ramp length in periods (3.0) */
    int use_relaxation;    /* 0: ss.cpp-style RHS without zones */
    double rel_tol, abs_tol;
    int max_iter;
} lpf_rhs_params;

int lpf_rhs_setup(lpf_ctx *ctx, const lpf_rhs_params *prm, const double *cgen_host, const double *cabs_host);
/* a14, cylinder driver: third relaxation weight C_absy (absorption towards y_max), added after C_abs in the
 * same order as Solvers/cylinder-diffraction.cpp:199-210.  NULL switches it off again.  Call after lpf_rhs_setup. */
int lpf_rhs_set_cabsy(lpf_ctx *ctx, const double *cabsy_host);
/* a16  rhs_linear::Mult at stage time t: state = [eta ; phi_fs] (2 n_surf), dstate likewise (:130-244) */
int lpf_rhs(lpf_ctx *ctx, double t, const double *state_dev, double *dstate_dev);
/* a15  RK4Solver::Step(x, t, dt): advances state in place, *t += dt                  (:494) */
int lpf_rk4_step(lpf_ctx *ctx, double *state_dev, double *t, double dt);
/* same with HOST state (H2D + step + D2H inside) */
int lpf_rk4_step_host(lpf_ctx *ctx, double *state_host, double *t, double dt);
/* f1  eta envelope of Solvers/cylinder-diffraction.cpp:410-432: env = -1e300; env = max(env, eta) per step
 * (state_dev = [eta ; phi_fs], device); _get copies it to the host and multiplies by `scale` (2/H, :444). */
int lpf_envelope_reset(lpf_ctx *ctx);
int lpf_envelope_update(lpf_ctx *ctx, const double *state_dev);
int lpf_envelope_get(lpf_ctx *ctx, double *env_host, double scale);
/* iteration counts of the (up to 4) solves of the last rhs / rk4 call */
int lpf_last_solve_info(lpf_ctx *ctx, lpf_pcg_info info[4], int *nsolves);
/* volume potential of the last solve (device pointer owned by the context, ndof doubles) */
const double *lpf_phi_dev(lpf_ctx *ctx);

/* ---- measurement helpers (used by bench.py; not part of the MFEM-facing surface) ---- */
/* runs `reps` constrained applies back to back and returns the CUDA-event time of the whole batch and
 * of the element kernel alone (ms); launches counted into *n_launches */
int lpf_time_apply(lpf_ctx *ctx, const double *x_dev, double *y_dev, int reps, float *ms_total,
                   float *ms_kernel, long *n_launches);
long lpf_launch_count(lpf_ctx *ctx);                          /* kernels launched by this context so far */
/* Options (defaults in brackets; every default is the measured best):
 *   "deterministic" [0]   1 = bit-reproducible results: the restriction-transpose is an ordered gather (element order, as
 *                         MFEM's CPU ElementRestriction::MultTranspose) of an E-vector instead of red.global.add, the
 *                         diagonal likewise; CG iteration counts are then identical from run to run.  ~20 % slower apply.
 *   "apply_variant" [0]   0 = tuned kernel per order; 20 = plain contractions; 30 = an alternative (E, CTAs/SM) pair; 40 / 41 = stage
 *                         buffers aliased into each other (/ + early q-data release) at orders 5-9 (the default at orders 7-9);
 *                         33 = the round-1 kernels of orders 7 / 8 (kept as evidence for profiles/r02_sweep_orders.txt)
 *   "affine" [1]          affine fast path when every element is affine (lpf_affine_active)
 *   "use_graph" [1], "pcg_chunk" [16]   CUDA graph of pcg_chunk CG iterations, status polled once per chunk
 *   "pdl" [0]             programmatic dependent launch between the kernels of a CG iteration (round 2: slower, off)
 *   "skip_zero_apply" [1] skip the initial-residual apply when the guess is zero off the essential dofs (exact)
 *   "p2p_fuse" [2]        multi-GPU halo-sum of an apply: 0 = separate LL kernel, 1 = last CTA of the apply kernel (only while
 *                         the interface has <= "p2p_fuse_max" [2048] entries), 2 = inside the apply kernel, overlapped with the
 *                         interior elements
 *   "host_pipeline" [1]   lpf_apply_T_host overlaps H2D / element chunks / D2H
 *   "hp_chunks" [8 from 64 MB per vector, else 4]   element chunks of that pipeline = copies per direction;
 *   "hp_ranges" [128]     dof ranges its dependencies are tracked on (16, 32, 64 or 128)
 *   "l2_persist" [0]      persisting-L2 window over z, d, A d
 *   "max_ctas" [0 = resident CTAs x SMs]   caps the persistent grid (tests)
 *   "verbose" [0]         print the launch geometry of every apply kernel once
 * The library reads no environment variables (a -DLPF_DEBUG_ENV developer build maps LPF_<OPTION> onto these). */
int lpf_set_option(lpf_ctx *ctx, const char *name, long value);
size_t lpf_device_bytes(const lpf_ctx *ctx);
/* 1 if the affine fast path (element tensor instead of stored q-data; option "affine", on by default) is in use:
 * decided by lpf_pa_setup -- every element of the rank must be an affine hex */
int lpf_affine_active(const lpf_ctx *ctx);

/* plain device memory helpers so hosts without a CUDA runtime binding can drive the library */
void *lpf_dev_alloc(size_t bytes);
int lpf_dev_free(void *p);
int lpf_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes);
int lpf_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes);
int lpf_memset_dev(void *dst_dev, int value, size_t bytes);
void *lpf_host_alloc_pinned(size_t bytes);
int lpf_host_free_pinned(void *p);
int lpf_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* LPF_B200_H */
