// lj_common.cuh - shared declarations of the sm_100a Lennard-Jones integrator (libljgpu.so).
//
// Reference path being replaced: LennardJonesSimulation::integrate and its stages,
// /root/reference/src/simulator/lennardjonessimulation.cpp:70-576 (".cpp" in the comments below).
#pragma once

// LJ_EMUL: the same sources compiled by g++ against tests/cuda_emul/cuda_emul.h (threads of a block as host
// threads) - test infrastructure for checking kernel logic without a GPU, never part of libljgpu.so.
#ifndef LJ_EMUL
#include <cuda_runtime.h>
#endif
#include <cstdint>
#include <cstddef>

namespace ljgpu {

// ---------------------------------------------------------------------------------------------
// Exactly rounded, never-contracted fp64 arithmetic.  The reference's arithmetic is an x86-64
// baseline build (no FMA); wherever a result feeds an integer decision (cell id, wrap, histogram
// bin, neighbour count) or is cheap to keep identical (the HBM-bound integrator stages) we spell
// the operations with these so nvcc cannot fuse a multiply into an add.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dadd_rn(a, -b); }
__device__ __forceinline__ double xdiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double xsqrt(double a)          { return __dsqrt_rn(a); }

// x*x + y*y + z*z in the reference's association ((xx + yy) + zz), no contraction.
__device__ __forceinline__ double xnorm2(double x, double y, double z) {
    return xadd(xadd(xmul(x, x), xmul(y, y)), xmul(z, z));
}

// apply_boundary_conditions, .cpp:431-433:  x -= L * floor(x * invL)
__device__ __forceinline__ double wrap_coord(double x, double L, double invL) {
    return xsub(x, xmul(L, floor(xmul(x, invL))));
}

// One-shot minimum image, .cpp:295-302 / :523-530.
__device__ __forceinline__ double min_image(double d, double L, double halfL) {
    if (d >= halfL) d = xsub(d, L);
    else if (d < -halfL) d = xadd(d, L);
    return d;
}

// Cell coordinate along one axis, .cpp:468-474:  min((unsigned)(x / edge), nc - 1).
// The float->unsigned conversion of a negative value is undefined in C++; the reference only
// ever converts wrapped coordinates (>= 0).  __double2uint_rz saturates, which agrees there.
__device__ __forceinline__ unsigned cell_coord(double x, double edge, unsigned nc) {
    unsigned i = __double2uint_rz(xdiv(x, edge));
    return i < nc - 1u ? i : nc - 1u;
}

// ---------------------------------------------------------------------------------------------
// Constants of one simulation, passed to kernels by value.
// ---------------------------------------------------------------------------------------------
struct LJConst {
    double L, invL, halfL;        // box edge, 1/L, L*0.5
    double rc2;                   // rcut*rcut                                  (.cpp:262)
    double sigma2;                // sigma*sigma                                (.cpp:264)
    double ecut;                  // (s2/rc2)^6 - (s2/rc2)^3                    (.cpp:270-273)
    double prefctr;               // 24*epsilon                                 (.cpp:267)
    double epsilon;
    double mass;
    double ekin0;                 // 1.5*(N-1)*kT                               (.cpp:374)
    double dt_over_tau;           // set per step
    double far2;                  // a hair under (L/2)^2: raw d2 below this => no axis needs the minimum image
    double list_cut2;             // (rcut+shell)^2                             (.cpp:445)
    double disp_thresh;           // shell*shell*0.25                           (.cpp:559-560)
    double cell_edge;             // L / cells_per_axis                         (.cpp:455-457)
    double prune_cut2;            // (rcut + s_in)^2: cutoff of the pruned inner list
    double prune_half;            // s_in / 2: how far atoms may have moved since the pruning
    double dt;                    // this step's dt (set per step)
    unsigned nc;                  // cells per axis                             (.cpp:447-453)
    unsigned n;                   // atoms owned by this device
    // z-slab decomposition (slab == 0: one device owns the whole box)
    unsigned slab;                // 1 on the multi-device path
    unsigned zlo;                 // first cell layer (global iz) owned by this device
    unsigned nzl;                 // number of owned cell layers
};

constexpr int kHistCap = 4096;    // entries in the energy history ring
constexpr int kSampleCap = 8192;  // entries in the ring of sampled energies (one per sample_stride steps: the GraphWidget's series)
struct HistEntry { double ekin, epot; unsigned int step; unsigned int pad; };

// Device-resident control block: accumulators, published energies, batch bookkeeping.
struct StepCtrl {
    double pred_sum;              // sum |v + a f|^2 : kinetic sum the NEXT kick1 will see
    double ekin, epot, etot;      // published (.h:100-102)
    double epot_sum;              // sum over ordered pairs of (s12 - s6 - ecut), last force pass
    unsigned long long maxdisp2;  // bits of max_i |dij_i|^2 of the current step (when above the threshold)
    unsigned int stop;            // a re-binning is due: remaining steps of the batch do nothing
    unsigned int steps_done;      // steps completed in the current batch
    unsigned int next_step;       // the `step` argument of the integrate() call being executed (history, log cadence)
    unsigned int ticket;          // blocks of the current kick2 launch that have finished (the last one finalizes the step)
    unsigned int overflow;        // list capacity exceeded during a build
    unsigned int max_count;       // max list length seen by the last build
    unsigned int max_tile;        // largest shared-memory tile (atoms) over all home blocks
    unsigned int max_home;        // most home atoms in one block
    unsigned long long hist_head; // energy history ring: entries written so far
    unsigned long long sample_head; // sampled-energy ring (GraphWidget series): entries written so far
    // pruned inner list (single-device Verlet path): valid while drift_bound <= prune_half
    unsigned long long vmax2;     // bits of max_i |v_i|^2 of the current step's drift velocities
    double drift_bound;           // sum over the steps since the pruning of max_i |v_i| dt
    unsigned int inner_valid;     // an inner list exists
    unsigned int inner_used;      // steps that used it (statistics)
    double vmax_last;             // max_i |v_i| of the last completed step (host sizes the pruning cadence from it)
    double red[4];                // slab path: {epot_sum, ekin_sum, pred_sum, rebuild votes}, local then all-reduced
    // slab path over peer memory: sequence numbers of the halo pushes and of the mailbox all-reduces done so far.  They live on
    // the device (not in kernel arguments) so that a captured CUDA graph of a step stays valid from step to step.
    unsigned long long halo_seq, red_seq;
    // the energy series the GraphWidget plots: every sample_stride-th step's energies (threadintegrate.cpp:57-62: every 1000
    // iterations) are appended to `samples` by the kernel that publishes the step
    HistEntry* samples;
    unsigned int sample_stride;
    // the brick kernel's work queue: next brick to hand out, and how many CTAs of the current launch have run dry (the last
    // one to do so zeroes both, so every launch starts from 0 without a memset in between)
    unsigned int brick_next, brick_dry, pad1;
};

// Slab path, peer-memory exchange: every device owns one mailbox that its peers write over NVLink.
constexpr int kMaxPeers = 16;
struct Mailbox {
    unsigned long long halo_flag[2];                 // [0]: written by my lower neighbour, [1]: by my upper; value = step sequence
    unsigned long long red_flag[2][kMaxPeers];       // [parity][source rank] = sequence number of the deposited sums
    double red[2][kMaxPeers][4];                     // [parity][source rank]{epot_sum, ekin_sum, pred_sum, rebuild vote}
    double nb_vmax2[2];                              // max |v|^2 of this step's drift on my lower [0] / upper [1] neighbour (pushed with the halo)
    unsigned long long pub_flag[kMaxPeers];          // [source rank] = sequence number of the publication whose atoms that rank has scattered into my frame slice
    unsigned int error;                              // a wait timed out (peer gone): results are invalid
    unsigned int pad;
    unsigned long long timeout_cycles;               // how long a wait for a peer may take (SM clock cycles) before it is reported as gone
};



// ---------------------------------------------------------------------------------------------
// Deterministic block reduction (fixed shape => fixed summation order).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}
// All threads must call; result valid in thread 0.  blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ double block_sum(double v, double* smem32) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) smem32[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double r = 0.0;
    if (w == 0) {
        r = lane < nw ? smem32[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}
__device__ __forceinline__ double block_max(double v, double* smem32) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_max(v);
    if (lane == 0) smem32[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double r = 0.0;
    if (w == 0) {
        r = lane < nw ? smem32[lane] : 0.0;
        r = warp_max(r);
    }
    __syncthreads();
    return r;
}

// ---------------------------------------------------------------------------------------------
// The Lennard-Jones pair term, .cpp:307-317, for a pair already known to be inside the cutoff.
// Returns fr (force / 24 eps / distance-vector) and adds the pair's s12 - s6 to e; the constant
// -ecut per accepted pair is applied once per thread from the accepted-pair count.
//   inv = 1/d2 ; s2 = sigma2*inv ; s6 = s2^3 ; s12 = s6^2 ; e += s12 - s6 ; fr = (2 s12 - s6) inv
// The reciprocal is the hardware seed (MUFU.RCP64H) refined by one cubic step: relative error a few
// 1e-16 (measured against 1.0/x in tests), no slow path needed because d2 is a normal number.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double fast_rcp(double a) {
    double y;
#ifdef LJ_EMUL
    y = ljemul::rcp_seed(a);
#else
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));     // MUFU.RCP64H: relative error e0 < 2^-20
#endif
    // one third-order step: y (1 + e + e^2) with e = 1 - a y  ->  error e0^3 < 2^-60, three DFMA
    const double e = fma(-a, y, 1.0);
    const double p = fma(e, e, e);
    return fma(y, p, y);
}
__device__ __forceinline__ double lj_pair(double d2, double sigma2, double& e) {
    const double inv = fast_rcp(d2);
    const double s2 = sigma2 * inv;
    const double s6 = s2 * s2 * s2;
    const double t = fma(s6, s6, -s6);       // s12 - s6
    const double u = fma(s6, s6, t);         // 2 s12 - s6
    e += t;
    return u * inv;
}

// ---------------------------------------------------------------------------------------------
// One position = one 32-byte sector.  sm_100 has 256-bit global loads/stores (SASS LDG.E.ENL2.256 /
// STG.E.ENL2.256); a plain double4 access is split into two 128-bit halves, which doubles the L1
// sector traffic of the gather in the Verlet force kernel.  cudaMalloc'ed double4 arrays are
// 32-byte aligned element by element.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double4 ld_pos(const double4* p) {
    double4 r;
#ifdef LJ_EMUL
    r = *p;
#else
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
#endif
    return r;
}
// read-only path (positions are not written by the kernel that uses this)
__device__ __forceinline__ double4 ld_pos_nc(const double4* p) {
    double4 r;
#ifdef LJ_EMUL
    r = *p;
#else
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
#endif
    return r;
}
__device__ __forceinline__ void st_pos(double4* p, const double4& v) {
#ifdef LJ_EMUL
    *p = v;
#else
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
#endif
}

// Asynchronous global->shared copies (LDGSTS) behind helpers; shared addresses are 32-bit shared-window offsets.
#ifdef LJ_EMUL
#define LJ_DYNAMIC_SMEM(name) double* name = reinterpret_cast<double*>(ljemul::dynamic_smem())
__device__ __forceinline__ unsigned smem_addr(const void* p) { return ljemul::smem_offset(p); }
__device__ __forceinline__ void cp_async8(unsigned saddr, const void* g) { ljemul::smem_copy(saddr, g, 8); }
__device__ __forceinline__ void cp_async16(unsigned saddr, const void* g) { ljemul::smem_copy(saddr, g, 16); }
__device__ __forceinline__ void cp_async_wait_all() {}
#else
#define LJ_DYNAMIC_SMEM(name) extern __shared__ __align__(16) double name[]
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async8(unsigned saddr, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async16(unsigned saddr, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#endif

// Monotonic counters in shared memory with release / acquire semantics at CTA scope: how the staging warp and the
// computing warps of the brick kernel hand tiles over (a counter, unlike an mbarrier phase bit, cannot alias when a
// waiter is several hand-overs ahead).
__device__ __forceinline__ unsigned ld_acquire_smem(const unsigned* p) {
#ifdef LJ_EMUL
    return *reinterpret_cast<const volatile unsigned*>(p);
#else
    unsigned v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_addr(p)) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void st_release_smem(unsigned* p, unsigned v) {
#ifdef LJ_EMUL
    *reinterpret_cast<volatile unsigned*>(p) = v;
    ljemul::note_progress();
#else
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" :: "r"(smem_addr(p)), "r"(v) : "memory");
#endif
}
__device__ __forceinline__ void spin_pause() {
#ifdef LJ_EMUL
    ljemul::yield_fiber();
#else
    __nanosleep(200);      // (a polling warp takes issue slots from the warps that walk lists)
#endif
}

// The staging warps' idle wait (a ring slot to come free, the computing warps to want another brick): these waits last
// many microseconds - in the list build 15 % of all instructions issued were this poll - so they nap longer.
__device__ __forceinline__ void spin_pause_long() {
#ifdef LJ_EMUL
    ljemul::yield_fiber();
#else
    __nanosleep(800);
#endif
}

// Barrier among a subset of a CTA's warps (bar.sync id, count): the staging warps of the brick kernel use id 1.
__device__ __forceinline__ void named_barrier_sync(unsigned id, unsigned nthreads) {
#ifdef LJ_EMUL
    ljemul::named_sync(id, nthreads);
#else
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
#endif
}

// ---------------------------------------------------------------------------------------------
// Device-wide primitives implemented in lj_sort.cu
// ---------------------------------------------------------------------------------------------
struct SortScratch {
    uint32_t* keys_alt = nullptr;
    uint32_t* vals_alt = nullptr;
    uint32_t* hist = nullptr;       // 256 * nblocks
    uint32_t* scan_tmp = nullptr;   // block sums for the scan (two levels)
    size_t    cap_items = 0;
    size_t    cap_hist = 0;
    size_t    cap_scan = 0;
};
// Stable LSD radix sort of (key,value) pairs on `key_bits` low bits.  On return the sorted data is
// in keys/vals (the alt buffers are scratch).  Returns number of kernels launched.
int radix_sort_pairs(uint32_t* keys, uint32_t* vals, size_t n, int key_bits, SortScratch& scr,
                     cudaStream_t st);
// In-place exclusive prefix sum of n uint32; total written to *total_out (device pointer, may be null).
int exclusive_scan_u32(uint32_t* data, size_t n, uint32_t* total_out, SortScratch& scr, cudaStream_t st);
// The first CUDA error the two helpers above met on this thread since the last call (cudaSuccess if none), cleared.
int sort_take_error();            // (a cudaError_t value)
void sort_scratch_free(SortScratch& scr);

} // namespace ljgpu
