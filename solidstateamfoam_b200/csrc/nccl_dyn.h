// NCCL bound at run time (dlopen), so libssam_gpu.so loads on a machine without NCCL/GPU and shares
// the libnccl.so.2 a host process (e.g. torch) has already loaded.  Types come from nccl.h.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <string>

namespace ssam {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;

    bool load() {
        if (handle) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) {
            error = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
            return false;
        }
#define SSAM_SYM(field, name)                                                     \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, name));               \
    if (!field) {                                                                 \
        error = std::string("NCCL symbol missing: ") + name;                      \
        return false;                                                             \
    }
        SSAM_SYM(GetUniqueId, "ncclGetUniqueId")
        SSAM_SYM(CommInitRank, "ncclCommInitRank")
        SSAM_SYM(CommDestroy, "ncclCommDestroy")
        SSAM_SYM(GroupStart, "ncclGroupStart")
        SSAM_SYM(GroupEnd, "ncclGroupEnd")
        SSAM_SYM(Send, "ncclSend")
        SSAM_SYM(Recv, "ncclRecv")
        SSAM_SYM(AllReduce, "ncclAllReduce")
        SSAM_SYM(GetErrorString, "ncclGetErrorString")
#undef SSAM_SYM
        return true;
    }
};

inline NcclApi& nccl() {
    static NcclApi api;
    return api;
}

}  // namespace ssam
