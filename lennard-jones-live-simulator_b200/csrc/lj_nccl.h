// lj_nccl.h - NCCL entry points resolved at run time (dlopen), so that libljgpu.so has no link-time
// dependency on NCCL: the single-device paths work without it, the slab path loads whichever
// libnccl.so.2 the process already has (torch's bundled one under bench.py) or the system one.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstddef>
#include <stdexcept>
#include <string>

namespace ljnccl {

struct UniqueId { char internal[128]; };                 // ncclUniqueId, NCCL_UNIQUE_ID_BYTES = 128
typedef struct ncclComm* Comm;
enum { kSum = 0, kMax = 2 };                             // ncclRedOp_t
enum { kUint8 = 1, kUint32 = 3, kUint64 = 5, kFloat64 = 8 };   // ncclDataType_t

struct Api {
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*CommAbort)(Comm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, Comm, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    bool ok = false;
    std::string error;
};

inline Api& api() {
    static Api a;
    static bool tried = false;
    if (tried) return a;
    tried = true;
    void* h = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { a.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?"); return a; }
    auto sym = [&](const char* n) { void* p = dlsym(h, n); if (!p) a.error += std::string(" missing ") + n; return p; };
    a.GetUniqueId = (int (*)(UniqueId*))sym("ncclGetUniqueId");
    a.CommInitRank = (int (*)(Comm*, int, UniqueId, int))sym("ncclCommInitRank");
    a.CommDestroy = (int (*)(Comm))sym("ncclCommDestroy");
    a.CommAbort = (int (*)(Comm))sym("ncclCommAbort");
    a.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    a.AllReduce = (int (*)(const void*, void*, size_t, int, int, Comm, cudaStream_t))sym("ncclAllReduce");
    a.AllGather = (int (*)(const void*, void*, size_t, int, Comm, cudaStream_t))sym("ncclAllGather");
    a.Send = (int (*)(const void*, size_t, int, int, Comm, cudaStream_t))sym("ncclSend");
    a.Recv = (int (*)(void*, size_t, int, int, Comm, cudaStream_t))sym("ncclRecv");
    a.GroupStart = (int (*)())sym("ncclGroupStart");
    a.GroupEnd = (int (*)())sym("ncclGroupEnd");
    a.ok = a.error.empty();
    return a;
}

} // namespace ljnccl
