// nccl_shim.h -- NCCL bound at run time (dlopen) instead of at link time.
//
// The library must coexist with whatever NCCL the host process already carries (PyTorch bundles its own
// libnccl.so.2) and must load on machines where only the single-GPU path is used.  The first call that needs
// NCCL resolves libnccl.so.2 -- the copy already mapped into the process if there is one -- and the handful of
// entry points the row-partitioned PCG uses.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

namespace mofem {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    const char *(*GetErrorString)(ncclResult_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
};

// nullptr when no NCCL can be loaded
inline const NcclApi *nccl_api()
{
    static NcclApi api;
    static int state = 0;  // 0 untried, 1 ok, -1 failed
    if (state == 0) {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        state = -1;
        if (h) {
            bool ok = true;
#define MOFEM_NCCL_SYM(field, name) ok = ok && ((*(void **)(&api.field) = dlsym(h, name)) != nullptr)
            MOFEM_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
            MOFEM_NCCL_SYM(CommInitRank, "ncclCommInitRank");
            MOFEM_NCCL_SYM(CommDestroy, "ncclCommDestroy");
            MOFEM_NCCL_SYM(GetErrorString, "ncclGetErrorString");
            MOFEM_NCCL_SYM(GroupStart, "ncclGroupStart");
            MOFEM_NCCL_SYM(GroupEnd, "ncclGroupEnd");
            MOFEM_NCCL_SYM(Send, "ncclSend");
            MOFEM_NCCL_SYM(Recv, "ncclRecv");
            MOFEM_NCCL_SYM(AllReduce, "ncclAllReduce");
#undef MOFEM_NCCL_SYM
            if (ok) state = 1;
        }
    }
    return state == 1 ? &api : nullptr;
}

}  // namespace mofem
