// nccl_dyn.h -- NCCL resolved at run time (dlopen) so that the library loads on hosts without NCCL
// and binds to the copy PyTorch already mapped when running under torch.distributed.
// Replaces the reference's MPI send/recv + cudaMemcpyPeer plumbing
// (/root/reference/src/qubit_backend/circuit_distributed.rs:14-39,
//  /root/reference/damavand-gpu/rust_communication.cu:106-141).
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <string>

namespace dvd {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    int (*GetVersion)(int*) = nullptr;

    // returns empty string on success, else the reason
    std::string load() {
        if (handle) return "";
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // already mapped (e.g. by torch)?
            if (handle) break;
        }
        if (!handle)
            for (const char* n : names) {
                handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
                if (handle) break;
            }
        if (!handle) return std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
#define DVD_SYM(field, name)                                                        \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, name));                 \
    if (!field) return std::string("NCCL symbol missing: ") + name;
        DVD_SYM(GetUniqueId, "ncclGetUniqueId")
        DVD_SYM(CommInitRank, "ncclCommInitRank")
        DVD_SYM(CommDestroy, "ncclCommDestroy")
        DVD_SYM(Send, "ncclSend")
        DVD_SYM(Recv, "ncclRecv")
        DVD_SYM(GroupStart, "ncclGroupStart")
        DVD_SYM(GroupEnd, "ncclGroupEnd")
        DVD_SYM(AllReduce, "ncclAllReduce")
        DVD_SYM(AllGather, "ncclAllGather")
        DVD_SYM(GetErrorString, "ncclGetErrorString")
#undef DVD_SYM
        GetVersion = reinterpret_cast<int (*)(int*)>(dlsym(handle, "ncclGetVersion"));
        return "";
    }
};

}  // namespace dvd
