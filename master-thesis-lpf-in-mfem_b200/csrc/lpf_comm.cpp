#include "lpf_comm.hpp"

#include <dlfcn.h>
#include <nccl.h>

#include <mutex>
#include <string>

#include "../../include/lpf_b200.h"
#include "../host/lpf_common.hpp"

namespace lpf {
namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_api;
std::once_flag g_once;
std::string g_load_error;

void load_api()
{
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_api.handle) break;
    }
    if (!g_api.handle) { g_load_error = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return; }
    auto sym = [&](const char *s) {
        void *p = dlsym(g_api.handle, s);
        if (!p) g_load_error = std::string("missing NCCL symbol ") + s;
        return p;
    };
    g_api.GetUniqueId = (decltype(g_api.GetUniqueId))sym("ncclGetUniqueId");
    g_api.CommInitRank = (decltype(g_api.CommInitRank))sym("ncclCommInitRank");
    g_api.CommDestroy = (decltype(g_api.CommDestroy))sym("ncclCommDestroy");
    g_api.Send = (decltype(g_api.Send))sym("ncclSend");
    g_api.Recv = (decltype(g_api.Recv))sym("ncclRecv");
    g_api.AllReduce = (decltype(g_api.AllReduce))sym("ncclAllReduce");
    g_api.GroupStart = (decltype(g_api.GroupStart))sym("ncclGroupStart");
    g_api.GroupEnd = (decltype(g_api.GroupEnd))sym("ncclGroupEnd");
    g_api.GetErrorString = (decltype(g_api.GetErrorString))sym("ncclGetErrorString");
}

int ensure_api()
{
    std::call_once(g_once, load_api);
    if (!g_load_error.empty()) { set_error(g_load_error); return LPF_ERR_COMM; }
    return LPF_OK;
}

#define NCCL_TRY(call)                                                                        \
    do {                                                                                      \
        ncclResult_t r__ = (call);                                                            \
        if (r__ != ncclSuccess) {                                                             \
            set_error(std::string(#call) + ": " + g_api.GetErrorString(r__));                 \
            return LPF_ERR_COMM;                                                              \
        }                                                                                     \
    } while (0)

}  // namespace

int Comm::unique_id(void *id128)
{
    if (!id128) { set_error("lpf_comm_unique_id: null buffer"); return LPF_ERR_ARG; }
    if (int rc = ensure_api()) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    NCCL_TRY(g_api.GetUniqueId((ncclUniqueId *)id128));
    return LPF_OK;
}

int Comm::init(const void *id128, int nranks, int rank)
{
    if (int rc = ensure_api()) return rc;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    NCCL_TRY(g_api.CommInitRank(&c, nranks, id, rank));
    comm_ = c;
    return LPF_OK;
}

int Comm::exchange(const HaloPlan &h, cudaStream_t s)
{
    NCCL_TRY(g_api.GroupStart());
    for (int i = 0; i < h.n_nbr; i++) {
        const int off = h.nbr_offset[i], cnt = h.nbr_offset[i + 1] - off;
        NCCL_TRY(g_api.Send(h.sendbuf + off, (size_t)cnt, ncclDouble, h.nbr_rank[i], (ncclComm_t)comm_, s));
        NCCL_TRY(g_api.Recv(h.recvbuf + off, (size_t)cnt, ncclDouble, h.nbr_rank[i], (ncclComm_t)comm_, s));
    }
    NCCL_TRY(g_api.GroupEnd());
    return LPF_OK;
}

int Comm::allreduce_sum(double *buf, int count, cudaStream_t s)
{
    NCCL_TRY(g_api.AllReduce(buf, buf, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)comm_, s));
    return LPF_OK;
}

void Comm::destroy()
{
    if (comm_ && g_api.CommDestroy) g_api.CommDestroy((ncclComm_t)comm_);
    comm_ = nullptr;
}

}  // namespace lpf
