// cuda_emul.h - TEST INFRASTRUCTURE ONLY.  A minimal CUDA-on-CPU shim: the device headers of
// lennard-jones-live-simulator_b200/csrc are compiled by g++ with -DLJ_EMUL and their kernels run with the
// threads of a block as cooperatively scheduled fibers on one host thread (barriers and warp collectives yield to
// the next fiber), blocks spread over a few host threads.  It exists to check kernel LOGIC (indexing, list layout, staging tables) where no GPU is
// available; it is never linked into libljgpu.so and the product has no CPU path.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <functional>
#include <sys/mman.h>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static thread_local      /* one copy per host thread = per running block */
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct uint3 { unsigned x = 0, y = 0, z = 0; };
using cudaStream_t = void*;
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct double2 { double x, y; };
struct alignas(32) double4 { double x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
inline double4 make_double4(double a, double b, double c, double d) { return double4{a, b, c, d}; }
inline double2 make_double2(double a, double b) { return double2{a, b}; }

inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

// Minimal x86-64 context switch (callee-saved registers + stack pointer; no signal-mask system call like swapcontext).
extern "C" void ljemul_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.weak ljemul_switch
.type ljemul_switch, @function
ljemul_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size ljemul_switch, .-ljemul_switch
)");

namespace ljemul {

constexpr size_t kFiberStack = 128 * 1024;

struct WarpState { unsigned live = 0, count = 0, gen = 0; unsigned long long xch[32]; };
struct Fiber { void* sp = nullptr; bool done = false; unsigned tid = 0; };

// One running block: its fibers, barrier state and dynamic shared memory.  Lives on the host thread that runs it.
struct BlockRun {
    unsigned nthreads = 0, live = 0, bar_count = 0, bar_gen = 0;
    unsigned nbar_count[16] = {}, nbar_gen[16] = {};   // named barriers (bar.sync id, count)
    std::vector<WarpState> warps;
    std::vector<Fiber> fibers;
    std::vector<char> dyn_storage;
    char* dyn = nullptr;
    char* stacks = nullptr;
    size_t stacks_bytes = 0;
    void* sched_sp = nullptr;
    Fiber* cur = nullptr;
    const std::function<void()>* kernel = nullptr;
    unsigned long long progress = 0;
    // optional shared-memory gather profile (LJEMUL_LDS_PROFILE=1): byte offsets of the k-th profiled load of every lane
    std::vector<std::vector<uint32_t>> lds_addr;          // per warp: [call][lane], 0xffffffff = lane made no such call
    std::vector<unsigned> lds_calls;                      // per thread: profiled loads so far
    ~BlockRun() { if (stacks) munmap(stacks, stacks_bytes); }
};
struct LdsProfile {
    std::atomic<unsigned long long> instr{0}, wavefronts{0}, lanes{0};
    bool on = std::getenv("LJEMUL_LDS_PROFILE") != nullptr;
    ~LdsProfile() {
        if (on && instr.load())
            std::fprintf(stderr, "[ljemul lds] 64-bit gathers: %llu warp instructions, %.2f active lanes, %.3f wavefronts per instruction "
                                 "(conflict-free: 2 for a full warp)\n", instr.load(), (double)lanes.load() / (double)instr.load(),
                         (double)wavefronts.load() / (double)instr.load());
    }
};
inline LdsProfile g_lds;
inline thread_local BlockRun* t_run = nullptr;
inline thread_local unsigned t_tid = 0;

inline void* dynamic_smem() { return t_run->dyn; }
inline unsigned smem_offset(const void* p) { return (unsigned)(static_cast<const char*>(p) - t_run->dyn); }
inline void smem_copy(unsigned off, const void* g, unsigned bytes) { std::memcpy(t_run->dyn + off, g, bytes); }
// Called by tile_ld for every 64-bit shared-memory load of a gathered record (three per record).
inline void lds_note(const void* p);
inline double rcp_seed(double a) { return (double)(1.0f / (float)a); }        // ~2^-23 relative error, like MUFU.RCP64H
inline unsigned long long atomic_load_u64(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void atomic_store_u64(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline double atomic_load_f64(const double* p) {
    unsigned long long b = __atomic_load_n(reinterpret_cast<const unsigned long long*>(p), __ATOMIC_ACQUIRE);
    double d; std::memcpy(&d, &b, 8); return d;
}

inline void lds_note(const void* p) {
    if (!g_lds.on) return;
    BlockRun& r = *t_run;
    if (r.lds_calls.size() < r.nthreads) { r.lds_calls.assign(r.nthreads, 0); r.lds_addr.assign((r.nthreads + 31) / 32, {}); }
    const unsigned k = r.lds_calls[t_tid]++;
    auto& tab = r.lds_addr[t_tid >> 5];
    if (tab.size() < (size_t)(k + 1) * 32) tab.resize((size_t)(k + 1) * 32, 0xffffffffu);
    tab[(size_t)k * 32 + (t_tid & 31u)] = (uint32_t)(static_cast<const char*>(p) - r.dyn);
}
inline void lds_flush(BlockRun& r) {               // end of a block: conflict degree of every profiled warp instruction
    if (!g_lds.on) return;
    unsigned long long ni = 0, nw = 0, nl = 0;
    for (auto& tab : r.lds_addr)
        for (size_t k = 0; k + 32 <= tab.size(); k += 32) {
            unsigned wf = 0, lanes = 0;
            for (int half = 0; half < 2; ++half) {
                uint32_t seen[16][16]; unsigned cnt[16] = {};
                unsigned mx = 0;
                for (int l = 0; l < 16; ++l) {
                    const uint32_t a = tab[k + 16 * half + l];
                    if (a == 0xffffffffu) continue;
                    ++lanes;
                    const unsigned bank = (a >> 3) & 15u;
                    bool dup = false;
                    for (unsigned q = 0; q < cnt[bank]; ++q) dup = dup || seen[bank][q] == a;
                    if (!dup) { seen[bank][cnt[bank]++] = a; mx = std::max(mx, cnt[bank]); }
                }
                wf += mx;
            }
            if (lanes) { ++ni; nw += wf; nl += lanes; }
        }
    g_lds.instr += ni; g_lds.wavefronts += nw; g_lds.lanes += nl;
    r.lds_addr.clear(); r.lds_calls.clear();
}
inline void yield_fiber() {
    BlockRun& r = *t_run;
    Fiber* me = r.cur;
    ljemul_switch(&me->sp, r.sched_sp);
}
inline void sync_block() {
    BlockRun& r = *t_run;
    const unsigned gen = r.bar_gen;
    if (++r.bar_count == r.live) { r.bar_count = 0; ++r.bar_gen; ++r.progress; }
    else while (r.bar_gen == gen) yield_fiber();
}
inline void note_progress() { ++t_run->progress; }
inline void named_sync(unsigned id, unsigned count) {
    BlockRun& r = *t_run;
    const unsigned gen = r.nbar_gen[id];
    if (++r.nbar_count[id] == count) { r.nbar_count[id] = 0; ++r.nbar_gen[id]; ++r.progress; }
    else while (r.nbar_gen[id] == gen) yield_fiber();
}
inline void sync_warp() {
    BlockRun& r = *t_run;
    WarpState& w = r.warps[t_tid >> 5];
    const unsigned gen = w.gen;
    if (++w.count == w.live) { w.count = 0; ++w.gen; ++r.progress; }
    else while (w.gen == gen) yield_fiber();
}
template <class T>
inline T warp_exchange(T v, int src_lane) {           // every live lane of the warp calls this together
    static_assert(sizeof(T) <= 8, "8-byte lanes");
    WarpState& w = t_run->warps[t_tid >> 5];
    const int lane = (int)(t_tid & 31u);
    unsigned long long bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    w.xch[lane] = bits;
    sync_warp();
    const unsigned long long got = w.xch[(src_lane >= 0 && src_lane < 32) ? src_lane : lane];
    sync_warp();
    T r; std::memcpy(&r, &got, sizeof(T));
    return r;
}

inline void fiber_main() {
    BlockRun& r = *t_run;
    Fiber* me = r.cur;
    (*r.kernel)();
    // a finished thread no longer takes part in barriers: release the ones it was the last straggler of
    me->done = true;
    ++r.progress;
    --r.live;
    if (r.bar_count && r.bar_count == r.live) { r.bar_count = 0; ++r.bar_gen; }
    WarpState& w = r.warps[me->tid >> 5];
    --w.live;
    if (w.count && w.count == w.live) { w.count = 0; ++w.gen; }
    ljemul_switch(&me->sp, r.sched_sp);
    std::abort();                                   // a finished fiber is never resumed
}

// Runs one block to completion on the calling host thread.
inline void run_block(BlockRun& r, unsigned nthreads, size_t smem, uint3 bidx, dim3 grid, const std::function<void()>& kernel) {
    r.nthreads = r.live = nthreads; r.bar_count = 0; r.kernel = &kernel; r.progress = 0;
    for (int k = 0; k < 16; ++k) { r.nbar_count[k] = 0; r.nbar_gen[k] = 0; }
    r.warps.assign((nthreads + 31) / 32, WarpState());
    for (unsigned w = 0; w < r.warps.size(); ++w) r.warps[w].live = std::min(32u, nthreads - 32 * w);
    if (r.dyn_storage.size() < smem + 64) r.dyn_storage.resize(smem + 64);
    r.dyn = r.dyn_storage.data();
    r.dyn += (64 - (reinterpret_cast<uintptr_t>(r.dyn) & 63)) & 63;
    if (r.stacks_bytes < (size_t)nthreads * kFiberStack) {
        if (r.stacks) munmap(r.stacks, r.stacks_bytes);
        r.stacks_bytes = (size_t)nthreads * kFiberStack;
        r.stacks = static_cast<char*>(mmap(nullptr, r.stacks_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
        if (r.stacks == MAP_FAILED) { std::fprintf(stderr, "ljemul: cannot map fiber stacks\n"); std::abort(); }
    }
    r.fibers.assign(nthreads, Fiber());
    t_run = &r;
    blockIdx = bidx; blockDim = dim3(nthreads); gridDim = grid;
    for (unsigned t = 0; t < nthreads; ++t) {
        Fiber& f = r.fibers[t];
        f.tid = t;
        // initial frame: six callee-saved registers, the entry point for `ret`, a null return address
        char* top = r.stacks + (size_t)(t + 1) * kFiberStack;
        void** sp = reinterpret_cast<void**>(top) - 8;
        for (int k = 0; k < 6; ++k) sp[k] = nullptr;
        sp[6] = reinterpret_cast<void*>(&fiber_main);
        sp[7] = nullptr;
        f.sp = sp;
    }
    unsigned long long seen = ~0ull;
    while (r.live) {
        if (seen == r.progress) { std::fprintf(stderr, "ljemul: deadlock in block (%u,%u): %u threads wait at a barrier nobody else reaches\n", bidx.x, bidx.y, r.live); std::abort(); }
        seen = r.progress;
        for (unsigned t = 0; t < nthreads; ++t) {
            Fiber& f = r.fibers[t];
            if (f.done) continue;
            r.cur = &f; t_tid = t; threadIdx = uint3{t, 0, 0};
            ljemul_switch(&r.sched_sp, f.sp);
        }
    }
    lds_flush(r);
    t_run = nullptr;
}

}  // namespace ljemul


// ---- synchronisation, warp collectives ------------------------------------------------------------------
inline void __syncthreads() { ljemul::sync_block(); }
inline void __syncwarp(unsigned = 0xffffffffu) { ljemul::sync_warp(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return ljemul::warp_exchange(v, src); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return ljemul::warp_exchange(v, (int)(ljemul::t_tid & 31u) ^ m); }
template <class T> inline T __shfl_up_sync(unsigned, T v, int d) { const int l = (int)(ljemul::t_tid & 31u); return ljemul::warp_exchange(v, l - d >= 0 ? l - d : l); }
template <class T> inline T __shfl_down_sync(unsigned, T v, int d) { const int l = (int)(ljemul::t_tid & 31u); return ljemul::warp_exchange(v, l + d < 32 ? l + d : l); }
inline int __any_sync(unsigned, int pred) {
    int any = 0;
    for (int l = 0; l < 32; ++l) any |= ljemul::warp_exchange(pred ? 1 : 0, l);
    return any;
}

inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) { unsigned long long v = ((unsigned long long)b << 32) | a; unsigned r = 0; for (int k = 0; k < 4; ++k) r |= (unsigned)((v >> (8 * ((sel >> (4 * k)) & 7u))) & 0xffull) << (8 * k); return r; }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) { shift &= 31u; return shift ? (lo >> shift) | (hi << (32u - shift)) : lo; }
template <class T> inline unsigned __match_any_sync(unsigned, T v) {
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) if (ljemul::warp_exchange(v, l) == v) m |= 1u << l;
    return m;
}

// ---- atomics ----------------------------------------------------------------------------------------------
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T atomic_max_impl(T* p, T v) {
    T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline unsigned atomicMax(unsigned* p, unsigned v) { return atomic_max_impl(p, v); }
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { return atomic_max_impl(p, v); }

// ---- arithmetic / bit intrinsics ------------------------------------------------------------------------------
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }
inline int __double2hiint(double d) { unsigned long long b; std::memcpy(&b, &d, 8); return (int)(b >> 32); }
inline double __hiloint2double(int hi, int lo) {
    const unsigned long long b = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
    double d; std::memcpy(&d, &b, 8); return d;
}
inline int __float_as_int(float f) { int v; std::memcpy(&v, &f, 4); return v; }
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
inline unsigned __double2uint_rz(double d) { return d <= 0.0 ? 0u : (d >= 4294967295.0 ? 4294967295u : (unsigned)d); }
inline int __double2int_rz(double d) { return (int)d; }
template <class T> inline T __ldg(const T* p) { return *p; }
inline long long clock64() { return std::chrono::steady_clock::now().time_since_epoch().count(); }
inline void __nanosleep(unsigned) { if (ljemul::t_run) { ++ljemul::t_run->progress; ljemul::yield_fiber(); } }   // waiting for another "device" is not a deadlock
using std::floor;
using std::pow;
using std::exp;
using std::fma;
using std::sqrt;
using std::fmax;
using std::fabs;
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
inline double min(double a, double b) { return a < b ? a : b; }
inline double max(double a, double b) { return a > b ? a : b; }

namespace ljemul {

// Runs `kernel()` for every block of the grid: blocks are handed to a few host threads, each runs its block's
// threads as fibers.  grid may be an unsigned or a dim3 (x, y).
template <class G, class F>
void launch_grid(G grid, unsigned nthreads, size_t smem, F&& kernel) {
    const dim3 g = dim3(grid);
    const unsigned total = g.x * g.y;
    const std::function<void()> fn = kernel;
    static const unsigned hw = [] {
        const char* e = std::getenv("LJEMUL_THREADS");
        const unsigned n = e ? (unsigned)std::atoi(e) : std::thread::hardware_concurrency();
        return std::max(1u, std::min(n, 32u));
    }();
    const unsigned nworkers = std::min(hw, total);
    std::atomic<unsigned> next{0};
    auto work = [&] {
        BlockRun r;
        for (unsigned b = next.fetch_add(1); b < total; b = next.fetch_add(1))
            run_block(r, nthreads, smem, uint3{b % g.x, b / g.x, 0}, g, fn);
    };
    if (nworkers <= 1) { work(); return; }
    std::vector<std::thread> th;
    for (unsigned w = 0; w < nworkers; ++w) th.emplace_back(work);
    for (auto& x : th) x.join();
}
template <class F>
void launch(unsigned nblocks, unsigned nthreads, size_t smem, F&& kernel) { launch_grid(nblocks, nthreads, smem, kernel); }

}  // namespace ljemul
