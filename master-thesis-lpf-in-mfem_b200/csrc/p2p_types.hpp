// Plain data shared by the device layer's translation units: PCG state and the peer-memory exchange descriptors
// (see p2p.cuh for the protocol).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

enum { PCG_RUNNING = 0, PCG_CONVERGED = 1, PCG_MAXITER = 2, PCG_BREAKDOWN = 3, PCG_NOT_PD = 4, PCG_COMM_ERROR = 5 };

// (d, A d) of the PCG is accumulated by the apply kernel, one slot per CTA (grids are capped at this many CTAs, so every
// slot receives exactly one addend and the sum over the slots, taken in slot order, is run-to-run reproducible)
#define LPF_DEN_SLOTS 1024
#define LPF_MAX_PARTIALS 2048

struct PcgState {
    double nom, den, betanom, r0, nom0, beta, rel2, abs2;
    double red[4];              // staging for cross-rank reductions
    int iter, status, max_iter, final_iter;
    unsigned int counter;
    int comm_error;             // a bounded flag wait of the peer-memory exchange timed out during this solve
};

#define LPF_P2P_MAXR 16
#define LPF_P2P_SPIN_LIMIT (1ll << 27)

struct __align__(16) P2PBox {
    uint4 red_ll[2][LPF_P2P_MAXR];                      // [parity][source rank] LL lines of the scalar all-reduce
    long long ll_byte_off[2][2];                        // [plan][parity] byte offset of the LL halo receive areas
    int off_for_src[2][LPF_P2P_MAXR];                   // [plan][source rank]: where that rank writes in MY recv buffer
};

struct P2PLocal {
    unsigned long long red_seq;
    int error;
    unsigned long long ll_seq[2];      // LL halo exchanges done, per plan
    unsigned int ll_counter[2];
    unsigned int if_done;              // overlapped apply: interface batches finished in the current launch
    unsigned int ctas_done;            // overlapped apply: CTAs that finished their receive slice
};

struct P2PPlanDev {                    // device-side view of one halo plan
    int n_nbr, total, n_shared;
    const int *nbr_rank;               // [n_nbr]
    const int *nbr_offset;             // [n_nbr+1]
    const int *send_dofs;              // [total]
    const int *send_nbr;               // [total] neighbour index of each send entry
    const int *shared, *red_off, *red_src;
    // LL protocol (p2p_halo_ll_kernel): per shared dof i the (neighbour, position) pairs it is sent to
    const int *snd_off, *snd_nbr, *snd_pos;    // [n_shared+1], [total], [total]
    uint4 *const *ll_dst;              // [2 parities][n_nbr] where I write LL lines in each neighbour's box
    const uint4 *ll_recv[2];           // my LL receive areas
};

struct P2PDev {
    int nranks, rank;
    P2PBox *const *peers;              // [nranks] device pointers to every rank's box (incl. mine)
    P2PBox *mine;
    P2PLocal *local;
};

// Halo-sum (+ PCG denominator all-reduce) riding on the apply kernel (p2p_dev.cuh).
//   mode 0: off
//   mode 1: the last CTA of the kernel sends, polls and sums the whole interface (small interfaces)
//   mode 2: OVERLAPPED -- the elements holding shared dofs come first in the element order (n_if_batches batches); the
//           CTA that finishes the last of them sends the interface as LL lines while every CTA goes on with interior
//           batches; after its last batch each CTA polls and sums its slice of the interface.  The NVLink latency is
//           hidden behind the interior elements.
struct P2PTail {
    int mode, with_den;
    int n_if_batches;
    P2PDev d;
    P2PPlanDev h;
    PcgState *st;
    double *den_slots;
    unsigned int *done;
};
