// cuda_emul.h - TEST INFRASTRUCTURE ONLY.  A minimal CUDA-on-CPU shim: the device headers of
// lennard-jones-live-simulator_b200/csrc are compiled by g++ with -DLJ_EMUL and their kernels run with the
// threads of a block as host threads (one block at a time, std::barrier for __syncthreads / warp
// collectives).  It exists to check kernel LOGIC (indexing, list layout, staging tables) where no GPU is
// available; it is never linked into libljgpu.so and the product has no CPU path.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
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
#define __shared__ static
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

namespace ljemul {

struct Warp {
    std::barrier<> bar;
    unsigned long long xch[32];
    explicit Warp(int lanes) : bar(lanes) {}
};
struct Block {
    unsigned nthreads;
    std::barrier<> bar;
    std::vector<std::unique_ptr<Warp>> warps;
    std::vector<char> dyn_storage;
    char* dyn = nullptr;
    Block(unsigned nt, size_t smem) : nthreads(nt), bar((std::ptrdiff_t)nt), dyn_storage(smem + 64) {
        for (unsigned w = 0; w < (nt + 31) / 32; ++w) warps.emplace_back(new Warp((int)std::min(32u, nt - 32 * w)));
        dyn = dyn_storage.data();
        dyn += (64 - (reinterpret_cast<uintptr_t>(dyn) & 63)) & 63;
    }
};
inline thread_local Block* t_block = nullptr;
inline thread_local unsigned t_tid = 0;

inline void* dynamic_smem() { return t_block->dyn; }
inline unsigned smem_offset(const void* p) { return (unsigned)(static_cast<const char*>(p) - t_block->dyn); }
inline void smem_copy(unsigned off, const void* g, unsigned bytes) { std::memcpy(t_block->dyn + off, g, bytes); }
inline double rcp_seed(double a) { return (double)(1.0f / (float)a); }        // ~2^-23 relative error, like MUFU.RCP64H
inline unsigned long long atomic_load_u64(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void atomic_store_u64(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline double atomic_load_f64(const double* p) {
    unsigned long long b = __atomic_load_n(reinterpret_cast<const unsigned long long*>(p), __ATOMIC_ACQUIRE);
    double d; std::memcpy(&d, &b, 8); return d;
}

inline Warp& warp() { return *t_block->warps[t_tid >> 5]; }
template <class T>
inline T warp_exchange(T v, int src_lane) {           // every live lane of the warp calls this together
    static_assert(sizeof(T) <= 8, "8-byte lanes");
    Warp& w = warp();
    const int lane = (int)(t_tid & 31u);
    unsigned long long bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    w.xch[lane] = bits;
    w.bar.arrive_and_wait();
    const unsigned long long got = w.xch[(src_lane >= 0 && src_lane < 32) ? src_lane : lane];
    w.bar.arrive_and_wait();
    T r; std::memcpy(&r, &got, sizeof(T));
    return r;
}

}  // namespace ljemul

inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

// ---- synchronisation, warp collectives ------------------------------------------------------------------
inline void __syncthreads() { ljemul::t_block->bar.arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { ljemul::warp().bar.arrive_and_wait(); }
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
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
inline unsigned __double2uint_rz(double d) { return d <= 0.0 ? 0u : (d >= 4294967295.0 ? 4294967295u : (unsigned)d); }
inline int __double2int_rz(double d) { return (int)d; }
template <class T> inline T __ldg(const T* p) { return *p; }
inline long long clock64() { return std::chrono::steady_clock::now().time_since_epoch().count(); }
inline void __nanosleep(unsigned) { std::this_thread::yield(); }
using std::floor;
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

// Runs `kernel()` for every block of the grid, one block at a time, one host thread per CUDA thread.
template <class F>
void launch(unsigned nblocks, unsigned nthreads, size_t smem, F&& kernel) {
    for (unsigned b = 0; b < nblocks; ++b) {
        Block blk(nthreads, smem);
        std::vector<std::thread> th;
        th.reserve(nthreads);
        for (unsigned t = 0; t < nthreads; ++t)
            th.emplace_back([&, t] {
                t_block = &blk; t_tid = t;
                threadIdx = uint3{t, 0, 0}; blockIdx = uint3{b, 0, 0};
                blockDim = dim3(nthreads); gridDim = dim3(nblocks);
                kernel();
                blk.warps[t >> 5]->bar.arrive_and_drop();      // a finished thread no longer takes part in barriers
                blk.bar.arrive_and_drop();
            });
        for (auto& x : th) x.join();
    }
}

}  // namespace ljemul
