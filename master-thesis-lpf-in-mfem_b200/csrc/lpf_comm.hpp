// NCCL plumbing for the multi-GPU path: halo exchange of partial sums on shared dofs and the scalar
// all-reduces of the PCG.  Replaces ParFiniteElementSpace's GroupCommunicator (MPI_Isend/Irecv per
// neighbour) and CGSolver's MPI_Allreduce.  NCCL is dlopen'ed at lpf_comm_init so the
// single-GPU path has no NCCL dependency.
#pragma once
#include <cuda_runtime.h>

#include <vector>

namespace lpf {

struct HaloPlan {
    int n_nbr = 0, total = 0, n_shared = 0;
    std::vector<int> nbr_rank, nbr_offset;      // host copies drive the send/recv group
    std::vector<int> h_send, h_shared;          // host copies of send_dofs / shared (LL send lists are built from them)
    int *send_dofs = nullptr, *shared = nullptr, *red_off = nullptr, *red_src = nullptr;   // device
    double *sendbuf = nullptr, *recvbuf = nullptr;                                          // device
};

class Comm {
public:
    static int unique_id(void *id128);
    int init(const void *id128, int nranks, int rank);
    bool ready() const { return comm_ != nullptr; }
    int exchange(const HaloPlan &h, cudaStream_t s);          // grouped ncclSend/ncclRecv with every neighbour
    int allreduce_sum(double *buf_dev, int count, cudaStream_t s);   // in place
    void destroy();

private:
    void *comm_ = nullptr;
};

}  // namespace lpf
