// nccl_emul.cpp - TEST INFRASTRUCTURE ONLY.  A stand-in libnccl.so.2 for the CPU emulation (tests/cuda_emul): the
// "ranks" of a communicator are threads of one process, collectives are synchronous and rank-ordered.  Implements
// exactly the entry points lj_nccl.h resolves.  Never loaded on a machine with GPUs: the emulation tests put its
// directory on LD_LIBRARY_PATH of their own subprocess only.
#include <barrier>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace {

constexpr int kMaxRanks = 16;
struct Msg { const void* ptr; size_t bytes; };
struct World {
    int nranks;
    std::barrier<> bar;
    const void* ptrs[kMaxRanks] = {};
    std::mutex mu;
    std::deque<Msg> box[kMaxRanks][kMaxRanks];      // [src][dst], FIFO like NCCL's matching order
    int joined = 0;
    explicit World(int n) : nranks(n), bar(n) {}
};
struct Comm { std::shared_ptr<World> w; int rank; };
struct Pending { bool send; void* ptr; size_t bytes; int peer; Comm* comm; };

std::mutex g_mu;
std::map<std::string, std::shared_ptr<World>> g_worlds;
unsigned long long g_next_id = 1;
thread_local int t_group_depth = 0;
thread_local std::vector<Pending> t_pending;

size_t dtype_size(int t) { switch (t) { case 1: return 1; case 3: return 4; case 5: return 8; case 8: return 8; default: return 0; } }

template <class T> void reduce_into(T* out, const World& w, size_t count, int op) {
    for (size_t i = 0; i < count; ++i) {
        T acc = static_cast<const T*>(w.ptrs[0])[i];
        for (int r = 1; r < w.nranks; ++r) {
            const T v = static_cast<const T*>(w.ptrs[r])[i];
            acc = op == 0 ? (T)(acc + v) : (v > acc ? v : acc);
        }
        out[i] = acc;
    }
}

int flush_group() {
    if (t_pending.empty()) return 0;
    Comm* c = t_pending.front().comm;
    World& w = *c->w;
    {
        std::lock_guard<std::mutex> lk(w.mu);
        for (auto& p : t_pending) if (p.send) w.box[c->rank][p.peer].push_back(Msg{p.ptr, p.bytes});
    }
    w.bar.arrive_and_wait();                                  // every rank has posted its sends
    for (auto& p : t_pending) {
        if (p.send) continue;
        Msg m;
        {
            std::lock_guard<std::mutex> lk(w.mu);
            if (w.box[p.peer][c->rank].empty()) return 3;     // unmatched receive
            m = w.box[p.peer][c->rank].front();
            w.box[p.peer][c->rank].pop_front();
        }
        if (m.bytes > p.bytes) return 3;
        std::memcpy(p.ptr, m.ptr, m.bytes);
    }
    w.bar.arrive_and_wait();                                  // send buffers may be reused from here on
    t_pending.clear();
    return 0;
}

}  // namespace

extern "C" {

struct ncclUniqueId { char internal[128]; };

int ncclGetUniqueId(ncclUniqueId* id) {
    std::lock_guard<std::mutex> lk(g_mu);
    std::memset(id, 0, sizeof *id);
    std::snprintf(id->internal, sizeof id->internal, "ljemul-nccl-%llu", g_next_id++);
    return 0;
}

int ncclCommInitRank(void** comm, int nranks, ncclUniqueId id, int rank) {
    if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) return 4;
    std::shared_ptr<World> w;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        const std::string key(id.internal, strnlen(id.internal, sizeof id.internal));
        auto it = g_worlds.find(key);
        if (it == g_worlds.end()) it = g_worlds.emplace(key, std::make_shared<World>(nranks)).first;
        w = it->second;
        if (w->nranks != nranks) return 4;
        if (++w->joined == nranks) g_worlds.erase(key);       // complete: the id is spent
    }
    w->bar.arrive_and_wait();
    *comm = new Comm{w, rank};
    return 0;
}

int ncclCommDestroy(void* comm) { delete static_cast<Comm*>(comm); return 0; }
int ncclCommAbort(void* comm) { delete static_cast<Comm*>(comm); return 0; }
const char* ncclGetErrorString(int e) { return e == 0 ? "no error" : e == 3 ? "unmatched point-to-point operation" : "invalid argument"; }

int ncclAllReduce(const void* send, void* recv, size_t count, int dtype, int op, void* comm, void*) {
    Comm* c = static_cast<Comm*>(comm);
    World& w = *c->w;
    const size_t es = dtype_size(dtype);
    if (!es || (op != 0 && op != 2)) return 4;
    w.ptrs[c->rank] = send;
    w.bar.arrive_and_wait();
    std::vector<char> tmp(count * es);
    switch (dtype) {
        case 1: reduce_into(reinterpret_cast<uint8_t*>(tmp.data()), w, count, op); break;
        case 3: reduce_into(reinterpret_cast<uint32_t*>(tmp.data()), w, count, op); break;
        case 5: reduce_into(reinterpret_cast<uint64_t*>(tmp.data()), w, count, op); break;
        case 8: reduce_into(reinterpret_cast<double*>(tmp.data()), w, count, op); break;
    }
    w.bar.arrive_and_wait();                                  // everybody has read every input (in-place is allowed)
    std::memcpy(recv, tmp.data(), tmp.size());
    return 0;
}

int ncclAllGather(const void* send, void* recv, size_t sendcount, int dtype, void* comm, void*) {
    Comm* c = static_cast<Comm*>(comm);
    World& w = *c->w;
    const size_t bytes = sendcount * dtype_size(dtype);
    if (!dtype_size(dtype)) return 4;
    std::vector<char> mine(bytes);
    std::memcpy(mine.data(), send, bytes);                    // the send buffer may lie inside the receive buffer
    w.ptrs[c->rank] = mine.data();
    w.bar.arrive_and_wait();
    for (int r = 0; r < w.nranks; ++r) std::memcpy(static_cast<char*>(recv) + (size_t)r * bytes, w.ptrs[r], bytes);
    w.bar.arrive_and_wait();
    return 0;
}

int ncclGroupStart() { ++t_group_depth; return 0; }
int ncclGroupEnd() {
    if (--t_group_depth > 0) return 0;
    t_group_depth = 0;
    return flush_group();
}
int ncclSend(const void* ptr, size_t count, int dtype, int peer, void* comm, void*) {
    t_pending.push_back(Pending{true, const_cast<void*>(ptr), count * dtype_size(dtype), peer, static_cast<Comm*>(comm)});
    return t_group_depth ? 0 : flush_group();
}
int ncclRecv(void* ptr, size_t count, int dtype, int peer, void* comm, void*) {
    t_pending.push_back(Pending{false, ptr, count * dtype_size(dtype), peer, static_cast<Comm*>(comm)});
    return t_group_depth ? 0 : flush_group();
}

}  // extern "C"
