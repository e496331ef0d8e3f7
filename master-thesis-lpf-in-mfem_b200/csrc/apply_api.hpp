// Host-side interface between the context (lpf_device.cu) and the per-order apply translation units
// (apply_order.cu compiled once per order with -DLPF_ORDER=p): kernels of different orders share no code at run time,
// so they are compiled in parallel and each keeps only its own coefficient tables in constant memory.
#pragma once
#include <cuda_runtime.h>

#include "p2p_types.hpp"

// Everything one apply launch needs (kernel argument, by value)
struct ApplyKArgs {
    const double *qd;          // stored q-data [ne][Q][3][Q^2][2] (or [ne][6] element tensors for the affine kernels)
    const int *gmap;           // [ne][DP3], essential dofs as ~dof in the constrained map
    const double *x;
    double *y;                 // L-vector, scatter-added; unused when yE != nullptr
    double *yE;                // deterministic mode: E-vector output [ne][D^3] (plain stores), y is assembled by det_gather_kernel
    int ne;
    double *den_slots;         // (d, A d) partials, one slot per CTA (DEN kernels)
    const int *status;         // PCG status word: != 0 -> the launch is a no-op
    P2PTail tail;              // multi-GPU: halo-sum riding on this kernel
};

struct LpfApplyArgs {
    int dev = 0, sm_count = 148;
    cudaStream_t stream = nullptr;
    int variant = 0;           // 0 = tuned kernel of the order; 20 = plain contractions; 30.. = alternative (E, CTAs/SM) pairs
    int max_ctas = 0;          // cap of the persistent grid (0 = resident CTAs x SMs)
    bool pdl = false;          // launch with the programmatic-dependent-launch attribute
    bool affine = false;       // k.qd holds [ne][6] element tensors (affine fast path)
    bool verbose = false;
    ApplyKArgs k{};
    int grid = 0, threads = 0; // out: launch geometry
    size_t smem = 0;
};

#define LPF_DECL_ORDER(P)                                                          \
    int lpf_apply_L_p##P(LpfApplyArgs &a);                                         \
    int lpf_apply_E_p##P(const double *qd, const double *xE, double *yE, int ne, cudaStream_t stream); \
    int lpf_apply_tables_p##P(const double *B, const double *G, const double *qwts);
LPF_DECL_ORDER(1) LPF_DECL_ORDER(2) LPF_DECL_ORDER(3) LPF_DECL_ORDER(4) LPF_DECL_ORDER(5)
LPF_DECL_ORDER(6) LPF_DECL_ORDER(7) LPF_DECL_ORDER(8) LPF_DECL_ORDER(9) LPF_DECL_ORDER(10)
#undef LPF_DECL_ORDER

#define LPF_ORDER_SWITCH(p, CALL, BAD)                   \
    switch (p) {                                          \
        case 1: CALL(1); break; case 2: CALL(2); break; case 3: CALL(3); break; case 4: CALL(4); break;   \
        case 5: CALL(5); break; case 6: CALL(6); break; case 7: CALL(7); break; case 8: CALL(8); break;   \
        case 9: CALL(9); break; case 10: CALL(10); break;                                                \
        default: BAD;                                                                                     \
    }
