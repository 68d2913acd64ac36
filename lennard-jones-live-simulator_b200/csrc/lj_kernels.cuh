// lj_kernels.cuh - every per-step and re-binning kernel of the integrator (included by lj_sim.cu only).
//
// Stage order of one step, LennardJonesSimulation::integrate
// (/root/reference/src/simulator/lennardjonessimulation.cpp:70-98, ".cpp" below):
//   kick1 -> Berendsen scale -> drift -> forces (on drifted, not yet wrapped positions) -> wrap
//   -> list-rebuild check (+ rebuild) -> kick2 -> etot -> publish.
// Here:  kick_drift_kernel (the PREVIOUS step's wrap applied lazily, kick1 + scale + drift + displacement
//                           accumulation and check: one pass over HBM)
//        force_*_kernel    (on the drifted, not yet wrapped positions, as the reference)
//        kick2_kernel      (kick2 + sum v^2 + the NEXT step's kick1 kinetic sum; its last block - ticket - does the
//                           fixed-order final reductions and publishes energies, history, rebuild flag).
// Stored positions are the reference's positions BEFORE its wrap of the current step; the wrap
// (x - L*floor(x/L), .cpp:431-433) is applied wherever the reference would read a wrapped value: at the
// top of the next drift, in the re-binning, and in every query.  Values are bit-identical to wrapping
// eagerly and it saves one read+write of the positions per step.
#pragma once
#include "lj_common.cuh"

namespace ljgpu {

constexpr int kBlock = 128;            // threads per block of the per-atom kernels
constexpr int kFinalizeThreads = 1024;

// ---------------------------------------------------------------------------------------------
// K0  predict: sum |v + a f|^2, the kinetic sum kick1 of the coming step will produce
//     (.cpp:350-354 evaluated ahead of time; same operations, so the same values).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
predict_kernel(unsigned n, double a,
               const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ vz,
               const double* __restrict__ fx, const double* __restrict__ fy, const double* __restrict__ fz,
               double* __restrict__ partial_pred) {
    __shared__ double red[32];
    const unsigned i = blockIdx.x * kBlock + threadIdx.x;
    double s = 0.0;
    if (i < n) {
        const double tx = xadd(vx[i], xmul(a, fx[i]));
        const double ty = xadd(vy[i], xmul(a, fy[i]));
        const double tz = xadd(vz[i], xmul(a, fz[i]));
        s = xnorm2(tx, ty, tz);
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) partial_pred[blockIdx.x] = s;
}

// Fixed-order partial sum of one thread of a finalize block: four independent accumulators so the
// loads overlap; the order depends only on the launch shape, so results are reproducible run to run.
__device__ __forceinline__ double strided_sum(const double* __restrict__ p, unsigned n) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    unsigned k = threadIdx.x;
    for (; k + 3 * kFinalizeThreads < n; k += 4 * kFinalizeThreads) {
        a0 += p[k]; a1 += p[k + kFinalizeThreads]; a2 += p[k + 2 * kFinalizeThreads]; a3 += p[k + 3 * kFinalizeThreads];
    }
    for (; k < n; k += kFinalizeThreads) a0 += p[k];
    return (a0 + a1) + (a2 + a3);
}

__global__ void __launch_bounds__(kFinalizeThreads)
predict_finalize_kernel(StepCtrl* __restrict__ ctrl, const double* __restrict__ partial_pred, unsigned nb) {
    __shared__ double red[32];
    double s = strided_sum(partial_pred, nb);
    s = block_sum(s, red);
    if (threadIdx.x == 0) ctrl->pred_sum = s;
}

// ---------------------------------------------------------------------------------------------
// K2  (lazy wrap) + kick1 + Berendsen + drift + displacement accumulation and check
//     .cpp:431-433 (wrap of the previous step), :336-362 (kick), :368-383 (lambda from the kick1 kinetic
//     energy), :388-413 (drift, dij += inc), :557-568 (any |dij|^2 > shell^2/4).
// ---------------------------------------------------------------------------------------------
// Body shared by the two kick_drift kernels: returns the atom's new position in `p` (valid when i < c.n).
template <bool TRACK_DISP>
__device__ __forceinline__ void kick_drift_body(const LJConst& c, double dt, double a, StepCtrl* __restrict__ ctrl,
                  double4* __restrict__ pos,
                  double* __restrict__ dijx, double* __restrict__ dijy, double* __restrict__ dijz,
                  double* __restrict__ vx, double* __restrict__ vy, double* __restrict__ vz,
                  const double* __restrict__ fx, const double* __restrict__ fy, const double* __restrict__ fz,
                  unsigned i, double4& p, double* red, double* red2) {
    // lambda = sqrt(1 + dt/tau * (ekin0/ekin - 1)),  ekin = 0.5*m*sum   (.cpp:354, :373-375)
    const double ek = xmul(xmul(0.5, c.mass), ctrl->pred_sum);
    const double lambda = xsqrt(xadd(1.0, xmul(c.dt_over_tau, xsub(xdiv(c.ekin0, ek), 1.0))));

    double d2 = 0.0, v2 = 0.0;
    if (i < c.n) {
        p = ld_pos(pos + i);
        p.x = wrap_coord(p.x, c.L, c.invL);
        p.y = wrap_coord(p.y, c.L, c.invL);
        p.z = wrap_coord(p.z, c.L, c.invL);
        double ux = xadd(vx[i], xmul(a, fx[i]));
        double uy = xadd(vy[i], xmul(a, fy[i]));
        double uz = xadd(vz[i], xmul(a, fz[i]));
        ux = xmul(ux, lambda); uy = xmul(uy, lambda); uz = xmul(uz, lambda);
        const double incx = xmul(ux, dt), incy = xmul(uy, dt), incz = xmul(uz, dt);
        p.x = xadd(p.x, incx); p.y = xadd(p.y, incy); p.z = xadd(p.z, incz);
        vx[i] = ux; vy[i] = uy; vz[i] = uz;
        st_pos(pos + i, p);
        if (TRACK_DISP) v2 = fma(uz, uz, fma(uy, uy, ux * ux));
        if (TRACK_DISP) {
            const double ax = xadd(dijx[i], incx), ay = xadd(dijy[i], incy), az = xadd(dijz[i], incz);
            dijx[i] = ax; dijy[i] = ay; dijz[i] = az;
            d2 = xnorm2(ax, ay, az);
        }
    }
    if (TRACK_DISP) {
        // two maxima in one pass over the block
        d2 = warp_max(d2); v2 = warp_max(v2);
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) { red[w] = d2; red2[w] = v2; }
        __syncthreads();
        if (w == 0) {
            d2 = lane < kBlock / 32 ? red[lane] : 0.0;
            v2 = lane < kBlock / 32 ? red2[lane] : 0.0;
            d2 = warp_max(d2); v2 = warp_max(v2);
            if (lane == 0) {
                if (d2 > c.disp_thresh) atomicMax(&ctrl->maxdisp2, (unsigned long long)__double_as_longlong(d2));
                // fastest atom of this step: bounds how far any pair distance can have shrunk since the pruning
                v2 *= (1.0 + 1e-12);
                if (v2 > __longlong_as_double((long long)*(volatile unsigned long long*)&ctrl->vmax2))
                    atomicMax(&ctrl->vmax2, (unsigned long long)__double_as_longlong(v2));
            }
        }
    }
}

template <bool TRACK_DISP>
__global__ void __launch_bounds__(kBlock)
kick_drift_kernel(LJConst c, double dt, double a, StepCtrl* __restrict__ ctrl,
                  double4* __restrict__ pos,
                  double* __restrict__ dijx, double* __restrict__ dijy, double* __restrict__ dijz,
                  double* __restrict__ vx, double* __restrict__ vy, double* __restrict__ vz,
                  const double* __restrict__ fx, const double* __restrict__ fy, const double* __restrict__ fz) {
    __shared__ double red[32];
    __shared__ double red2[32];
    if (ctrl->stop) return;
    double4 p;
    kick_drift_body<TRACK_DISP>(c, dt, a, ctrl, pos, dijx, dijy, dijz, vx, vy, vz, fx, fy, fz,
                                blockIdx.x * kBlock + threadIdx.x, p, red, red2);
}

// ---------------------------------------------------------------------------------------------
// Minimum image without paying for it on every pair.  The reference shifts an axis only when
// |d| >= L/2 (.cpp:295-302); then the raw d2 is at least L^2/4, far beyond any cutoff in use
// (L >= 3 (rcut+shell)).  So: raw d2 below `far2` (a hair under L^2/4) => no axis shifts and the raw
// difference IS the reference's; otherwise redo the pair with the reference's exact comparisons.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool maybe_wrapped(double d2raw, int far2_hi) {
    return __double2hiint(d2raw) >= far2_hi;          // integer compare on the high word: no fp64 issue slot
}

// ---------------------------------------------------------------------------------------------
// K1  all-pairs force kernel (small N): calculate_forces .cpp:258-330 with the pair set taken
//     over ALL j (the Verlet list only pre-filters it), minimum image per pair as :295-302, the
//     cutoff decision d2 >= rcut^2 on a d2 formed exactly like :304.
//     grid = (i tiles, j splits); each block stages its j range through shared memory; partial
//     forces of the j splits are summed in fixed order by allpairs_reduce_kernel.
// ---------------------------------------------------------------------------------------------
constexpr int kApTile = 128;      // j positions staged per shared-memory tile (== kBlock)

__global__ void __launch_bounds__(kBlock)
force_allpairs_kernel(LJConst c, const StepCtrl* __restrict__ ctrl, const double4* __restrict__ pos,
                      unsigned jchunk,
                      double* __restrict__ pfx, double* __restrict__ pfy, double* __restrict__ pfz,
                      double* __restrict__ partial_epot) {
    __shared__ double sx[kApTile], sy[kApTile], sz[kApTile];
    __shared__ double red[32];
    if (ctrl->stop) return;
    const unsigned i = blockIdx.x * kBlock + threadIdx.x;
    const bool active = i < c.n;
    double4 pi = active ? pos[i] : make_double4(0, 0, 0, 0);
    const unsigned j0 = blockIdx.y * jchunk;
    const unsigned j1 = min(j0 + jchunk, c.n);

    double ax = 0.0, ay = 0.0, az = 0.0, e = 0.0;
    unsigned nacc = 0;
    for (unsigned jb = j0; jb < j1; jb += kApTile) {
        const unsigned jl = jb + threadIdx.x;
        __syncthreads();
        if (jl < j1) { const double4 q = pos[jl]; sx[threadIdx.x] = q.x; sy[threadIdx.x] = q.y; sz[threadIdx.x] = q.z; }
        __syncthreads();
        const unsigned cnt = min((unsigned)kApTile, j1 - jb);
        // Four j per round, branch-free: the reference's one-shot minimum image (.cpp:295-302) as a selected
        // shift (d + s with s in {-L, 0, +L} is the same rounded value as d - L / d + L / d), d2 formed
        // without contraction exactly like .cpp:304; the Lennard-Jones term is evaluated only when some lane
        // of the warp has an in-cutoff pair (1.3 % of all pairs are), and merged with selects.
        for (unsigned jj0 = 0; jj0 < cnt; jj0 += 4) {
            double dx[4], dy[4], dz[4], d2[4];
            bool in[4];
            bool any = false;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned jj = min(jj0 + u, cnt - 1);
                double ex = xsub(pi.x, sx[jj]), ey = xsub(pi.y, sy[jj]), ez = xsub(pi.z, sz[jj]);
                ex = xadd(ex, ex >= c.halfL ? -c.L : (ex < -c.halfL ? c.L : 0.0));
                ey = xadd(ey, ey >= c.halfL ? -c.L : (ey < -c.halfL ? c.L : 0.0));
                ez = xadd(ez, ez >= c.halfL ? -c.L : (ez < -c.halfL ? c.L : 0.0));
                dx[u] = ex; dy[u] = ey; dz[u] = ez;
                d2[u] = xnorm2(ex, ey, ez);
                in[u] = active && d2[u] < c.rc2 && (jb + jj0 + u) != i && (jj0 + u) < cnt;
                any = any || in[u];
            }
            if (__any_sync(0xffffffffu, any)) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const double inv = fast_rcp(in[u] ? d2[u] : 1.0);
                    const double s2 = c.sigma2 * inv;
                    const double s6 = s2 * s2 * s2;
                    const double t0 = fma(s6, s6, -s6);
                    const double fr0 = fma(s6, s6, t0) * inv;
                    const double t = in[u] ? t0 : 0.0, fr = in[u] ? fr0 : 0.0;
                    e += t;
                    ax = fma(fr, dx[u], ax); ay = fma(fr, dy[u], ay); az = fma(fr, dz[u], az);
                    nacc += in[u] ? 1u : 0u;
                }
            }
        }
    }
    if (active) {
        const size_t o = (size_t)blockIdx.y * c.n + i;
        pfx[o] = ax; pfy[o] = ay; pfz[o] = az;
    }
    e = fma(-(double)nacc, c.ecut, e);
    e = block_sum(e, red);
    if (threadIdx.x == 0) partial_epot[blockIdx.y * gridDim.x + blockIdx.x] = e;
}

__global__ void __launch_bounds__(kBlock)
allpairs_reduce_kernel(LJConst c, const StepCtrl* __restrict__ ctrl, unsigned nsplit,
                       const double* __restrict__ pfx, const double* __restrict__ pfy, const double* __restrict__ pfz,
                       double* __restrict__ fx, double* __restrict__ fy, double* __restrict__ fz) {
    if (ctrl->stop) return;
    const unsigned i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= c.n) return;
    double ax = 0.0, ay = 0.0, az = 0.0;
    for (unsigned s = 0; s < nsplit; ++s) {
        const size_t o = (size_t)s * c.n + i;
        ax += pfx[o]; ay += pfy[o]; az += pfz[o];
    }
    fx[i] = ax * c.prefctr; fy[i] = ay * c.prefctr; fz[i] = az * c.prefctr;     // .cpp:321-325
}

// ---------------------------------------------------------------------------------------------
// Cell traversal of the neighbour-count probe (the hot kernels use the tiles of lj_tile.cuh):
// one warp per home cell, lanes = home atoms, the 27 periodic stencil cells of .cpp:488-516 streamed
// through a warp-private shared-memory tile.  Distances are formed exactly as the reference forms them
// (no contraction), so acceptance decisions are bit-identical.
// ---------------------------------------------------------------------------------------------
constexpr int kCellWarps = 4;

struct CellGrid {
    const uint32_t* __restrict__ cstart;     // first slot of every cell
    const uint32_t* __restrict__ ccount;     // atoms per cell
    unsigned nc;                             // cells per axis
};

// EXACT: distances formed without contraction (bit-exact acceptance, used by the parity probe);
// otherwise fused multiply-adds (the list / cell force do not need bit-exact acceptance: the list
// keeps a skin, the force tolerance is 1e-12).
template <bool EXACT, class Op>
__device__ __forceinline__ void cell_traverse(const LJConst& c, const CellGrid& g, const double4* __restrict__ pos,
                                              double4 (*tile)[32], Op& op) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned cell = blockIdx.x * kCellWarps + w;
    const unsigned nc = g.nc;
    if (cell >= nc * nc * nc) return;
    const unsigned ix = cell % nc, iy = (cell / nc) % nc, iz = cell / (nc * nc);
    const int far2_hi = __double2hiint(c.far2);
    const unsigned hs = g.cstart[cell], hn = g.ccount[cell];
    for (unsigned hb = 0; hb < hn; hb += 32) {
        const unsigned i = hs + hb + lane;
        const bool active = hb + lane < hn;
        const double4 pi = active ? ld_pos_nc(pos + i) : make_double4(0, 0, 0, 0);
        op.begin(i, active);
#pragma unroll 1
        for (int s = 0; s < 27; ++s) {
            // stencil cell with periodic wrap of the cell coordinates (.cpp:496-516)
            const int ox = s % 3 - 1, oy = (s / 3) % 3 - 1, oz = s / 9 - 1;
            const unsigned cx = (ix + nc + ox) % nc, cy = (iy + nc + oy) % nc, cz = (iz + nc + oz) % nc;
            const unsigned cid = cz * nc * nc + cy * nc + cx;
            const unsigned js = g.cstart[cid], jn = g.ccount[cid];
            for (unsigned jb = 0; jb < jn; jb += 32) {
                __syncwarp();
                if (jb + lane < jn) tile[w][lane] = ld_pos_nc(pos + js + jb + lane);
                __syncwarp();
                const unsigned cnt = min(32u, jn - jb);
                for (unsigned jj = 0; jj < cnt; ++jj) {
                    const double4 pj = tile[w][jj];
                    double ddx = xsub(pi.x, pj.x), ddy = xsub(pi.y, pj.y), ddz = xsub(pi.z, pj.z);
                    double d2 = EXACT ? xnorm2(ddx, ddy, ddz) : fma(ddz, ddz, fma(ddy, ddy, ddx * ddx));
                    if (maybe_wrapped(d2, far2_hi)) {
                        ddx = min_image(ddx, c.L, c.halfL);
                        ddy = min_image(ddy, c.L, c.halfL);
                        ddz = min_image(ddz, c.L, c.halfL);
                        d2 = EXACT ? xnorm2(ddx, ddy, ddz) : fma(ddz, ddz, fma(ddy, ddy, ddx * ddx));
                    }
                    op.pair(i, js + jb + jj, ddx, ddy, ddz, d2, active);
                }
            }
        }
        op.end(i, active);
    }
}

// Neighbour-count parity probe of include/ljgpu.h (ljgpu_get_neighbor_counts): the acceptance test of
// build_neighbor_list (.cpp:518-540, inclusive) or of the force loop (.cpp:291-305, strict) with a
// caller-supplied cutoff, bit-exact, original atom order out.
struct CountOp {
    double cut2; int inclusive; uint32_t* out_count; const uint32_t* orig; unsigned cnt;
    __device__ __forceinline__ void begin(unsigned, bool) { cnt = 0; }
    __device__ __forceinline__ void pair(unsigned i, unsigned j, double, double, double, double d2, bool active) {
        const bool in = inclusive ? (d2 <= cut2) : (d2 < cut2);
        if (active && in && j != i) ++cnt;
    }
    __device__ __forceinline__ void end(unsigned i, bool active) { if (active) out_count[orig[i]] = cnt; }
};

__global__ void __launch_bounds__(kCellWarps * 32)
count_cells_kernel(LJConst c, CellGrid g, const double4* __restrict__ pos, double cut2, int inclusive,
                   uint32_t* __restrict__ out_count, const uint32_t* __restrict__ orig) {
    __shared__ double4 tile[kCellWarps][32];
    CountOp op{cut2, inclusive, out_count, orig, 0};
    cell_traverse<true>(c, g, pos, tile, op);
}

// ---- slab path over peer memory (NVLink loads/stores, no NCCL in the step loop) ----------------------
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
#ifdef LJ_EMUL
    v = ljemul::atomic_load_u64(p);
#else
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
#endif
    return v;
}
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
#ifdef LJ_EMUL
    ljemul::atomic_store_u64(p, v);
#else
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
#endif
}
__device__ __forceinline__ double ld_sys_f64(const double* p) {
    double v;
#ifdef LJ_EMUL
    v = ljemul::atomic_load_f64(p);
#else
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
#endif
    return v;
}
// Spin until *flag >= seq.  A peer that is merely late (its host has not made the collective call yet) is waited for; a
// peer that does not answer within mine->timeout_cycles (default ~30 s of SM clock, LJGPU_PEER_TIMEOUT_S) is reported
// through mine->error instead of hanging the device for ever.
__device__ __forceinline__ bool wait_flag(const unsigned long long* flag, unsigned long long seq, Mailbox* mine) {
    const long long t0 = clock64();
    const long long limit = (long long)mine->timeout_cycles;
    while (ld_sys_u64(flag) < seq) {
        if (clock64() - t0 > limit) { atomicExch(&mine->error, 1u); return false; }
        __nanosleep(100);
    }
    __threadfence_system();
    return true;
}

// ---------------------------------------------------------------------------------------------
// K3  kick2 + kinetic sums + the step's final reductions and publication, ONE launch.  .cpp:336-362, :95-96, :327-329.
//     Also forms sum |v' + a f|^2, which is exactly the kinetic sum the next step's kick1 will produce (same forces,
//     same operations), so the thermostat factor of the next step is known without another pass.
//     Every block reduces its atoms' two kinetic sums and a fixed slice of the force pass's energy partials; the
//     block that finishes last (ticket) adds the per-block sums in fixed order - so the result does not depend on
//     which block that is - and publishes energies, history and the re-binning flag.  On the slab path over peer
//     memory (SLAB) the last block also runs the all-reduce through the mailboxes.
// ---------------------------------------------------------------------------------------------
constexpr int kKick2Block = 256;

// three sums at once: one barrier pair instead of three.  All threads call; results valid in thread 0.
__device__ __forceinline__ void block_sum3(double& a, double& b, double& cc, double (*red)[3]) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    a = warp_sum(a); b = warp_sum(b); cc = warp_sum(cc);
    if (lane == 0) { red[w][0] = a; red[w][1] = b; red[w][2] = cc; }
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    if (w == 0) {
        a = lane < nw ? red[lane][0] : 0.0; b = lane < nw ? red[lane][1] : 0.0; cc = lane < nw ? red[lane][2] : 0.0;
        a = warp_sum(a); b = warp_sum(b); cc = warp_sum(cc);
    }
    __syncthreads();
}

// fixed-order sum of p[0..n) by the threads of one block (four accumulators per thread so the loads overlap)
__device__ __forceinline__ double strided_sum_block(const double* __restrict__ p, unsigned n) {
    const unsigned T = blockDim.x;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    unsigned k = threadIdx.x;
    for (; k + 3 * T < n; k += 4 * T) { a0 += p[k]; a1 += p[k + T]; a2 += p[k + 2 * T]; a3 += p[k + 3 * T]; }
    for (; k < n; k += T) a0 += p[k];
    return (a0 + a1) + (a2 + a3);
}

// The tail of integrate() (.cpp:95-98 and the per-step bookkeeping of this layout), one thread.
__device__ __forceinline__ void publish_step(const LJConst& c, StepCtrl* __restrict__ ctrl, HistEntry* __restrict__ hist,
                                             double se, double sk, double sp, bool rebuild_due, double vmax2) {
    const double ekin = xmul(xmul(0.5, c.mass), sk);
    const double epot = xmul(xmul(xmul(se, 4.0), c.epsilon), 0.5);
    ctrl->epot_sum = se;
    ctrl->pred_sum = sp;
    ctrl->ekin = ekin; ctrl->epot = epot; ctrl->etot = xadd(ekin, epot);
    HistEntry& h = hist[ctrl->hist_head % kHistCap];
    h.ekin = ekin; h.epot = epot; h.step = ctrl->next_step; h.pad = 0;
    if (ctrl->sample_stride && h.step % ctrl->sample_stride == 0u) {       // signal_ekin / _epot / _etot cadence, threadintegrate.cpp:57-62
        ctrl->samples[ctrl->sample_head % kSampleCap] = h;
        ctrl->sample_head += 1;
    }
    ctrl->next_step += 1;
    ctrl->hist_head += 1;
    ctrl->steps_done += 1;
    if (rebuild_due) ctrl->stop = 1;
    // pruned inner list: the displacement bound advances by this step's fastest atom
    const double vmax = sqrt(vmax2);
    ctrl->vmax_last = vmax;
    ctrl->drift_bound += vmax * c.dt;
    ctrl->vmax2 = 0ull;
}

// All-reduce of four doubles over the devices through their mailboxes (peer memory over NVLink), by one block of at
// least 4 * world threads: thread r deposits this device's sums in peer r's mailbox (its own included), fences,
// raises the flag and waits for peer r's deposit in mine; the deposits are then added in rank order, so every device
// gets identical sums.  tot / g are shared-memory arrays of four doubles; g is valid for every thread on return.
__device__ __forceinline__ void mailbox_allreduce(const double* tot, double* g, double* dep,
                                                  Mailbox* mine, Mailbox* const* __restrict__ peers,
                                                  int rank, int world, unsigned long long seq) {
    const unsigned par = (unsigned)(seq & 1ull);
    if ((int)threadIdx.x < world) {
        Mailbox* dst = peers[threadIdx.x];
        double* slot = dst->red[par][rank];
        slot[0] = tot[0]; slot[1] = tot[1]; slot[2] = tot[2]; slot[3] = tot[3];
        __threadfence_system();
        st_sys_u64(&dst->red_flag[par][rank], seq);
        wait_flag(&mine->red_flag[par][threadIdx.x], seq, mine);
    }
    __syncthreads();                                       // every deposit has arrived before anyone reads
    if ((int)threadIdx.x < world * 4) dep[threadIdx.x] = ld_sys_f64(&mine->red[par][threadIdx.x >> 2][threadIdx.x & 3]);
    __syncthreads();
    if (threadIdx.x == 0) {
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        for (int r = 0; r < world; ++r)
            for (int k = 0; k < 4; ++k) s[k] += dep[r * 4 + k];
        g[0] = s[0]; g[1] = s[1]; g[2] = s[2]; g[3] = s[3];
    }
    __syncthreads();
}

struct SlabRed { Mailbox* mine; Mailbox* const* peers; int rank, world; };

// KIND: 0 = one device, 1 = slab over peer memory (mailbox all-reduce here), 2 = slab with the NCCL step loop (local sums only:
// ncclAllReduce and slab_publish_kernel follow)
template <int KIND>
__global__ void __launch_bounds__(kKick2Block)
kick2_kernel(LJConst c, double a, StepCtrl* __restrict__ ctrl, HistEntry* __restrict__ hist,
             double* __restrict__ vx, double* __restrict__ vy, double* __restrict__ vz,
             const double* __restrict__ fx, const double* __restrict__ fy, const double* __restrict__ fz,
             const double* __restrict__ partial_epot, unsigned nb_epot,
             double* __restrict__ block_sums /* [3][gridDim.x]: epot slice, ekin, pred */, SlabRed sr) {
    __shared__ double red[kKick2Block / 32][3];
    __shared__ unsigned s_last;
    __shared__ double s_tot[4], s_g[4], s_dep[kMaxPeers * 4], s_nbv2;
    if (ctrl->stop) return;
    // two atoms per thread and round (128-bit loads and stores), a grid of one resident wave that strides over the
    // atoms: the per-block epilogue (two barriers, a ticket) is paid ~1200 times per launch, not once per 256 atoms
    double s1 = 0.0, s2 = 0.0;
    const unsigned npair = c.n >> 1;
    for (unsigned p = blockIdx.x * kKick2Block + threadIdx.x; p < npair; p += gridDim.x * kKick2Block) {
        const double2 f0 = reinterpret_cast<const double2*>(fx)[p], f1 = reinterpret_cast<const double2*>(fy)[p], f2 = reinterpret_cast<const double2*>(fz)[p];
        double2 u0 = reinterpret_cast<double2*>(vx)[p], u1 = reinterpret_cast<double2*>(vy)[p], u2 = reinterpret_cast<double2*>(vz)[p];
        const double gxa = xmul(a, f0.x), gya = xmul(a, f1.x), gza = xmul(a, f2.x);
        const double gxb = xmul(a, f0.y), gyb = xmul(a, f1.y), gzb = xmul(a, f2.y);
        u0.x = xadd(u0.x, gxa); u1.x = xadd(u1.x, gya); u2.x = xadd(u2.x, gza);
        u0.y = xadd(u0.y, gxb); u1.y = xadd(u1.y, gyb); u2.y = xadd(u2.y, gzb);
        reinterpret_cast<double2*>(vx)[p] = u0; reinterpret_cast<double2*>(vy)[p] = u1; reinterpret_cast<double2*>(vz)[p] = u2;
        s1 += xnorm2(u0.x, u1.x, u2.x);
        s1 += xnorm2(u0.y, u1.y, u2.y);
        s2 += xnorm2(xadd(u0.x, gxa), xadd(u1.x, gya), xadd(u2.x, gza));
        s2 += xnorm2(xadd(u0.y, gxb), xadd(u1.y, gyb), xadd(u2.y, gzb));
    }
    if ((c.n & 1u) && blockIdx.x == 0 && threadIdx.x == 0) {       // odd atom count: the last atom
        const unsigned i = c.n - 1;
        const double gx = xmul(a, fx[i]), gy = xmul(a, fy[i]), gz = xmul(a, fz[i]);
        const double ux = xadd(vx[i], gx), uy = xadd(vy[i], gy), uz = xadd(vz[i], gz);
        vx[i] = ux; vy[i] = uy; vz[i] = uz;
        s1 += xnorm2(ux, uy, uz);
        s2 += xnorm2(xadd(ux, gx), xadd(uy, gy), xadd(uz, gz));
    }
    // this block's slice of the force pass's energy partials (fixed slices => fixed summation order)
    const unsigned per = (nb_epot + gridDim.x - 1) / gridDim.x;
    double s0 = 0.0;
    for (unsigned k = threadIdx.x; k < per; k += kKick2Block) {
        const unsigned j = blockIdx.x * per + k;
        if (j < nb_epot) s0 += partial_epot[j];
    }
    block_sum3(s0, s1, s2, red);
    const unsigned G = gridDim.x;
    if (threadIdx.x == 0) {
        block_sums[blockIdx.x] = s0; block_sums[G + blockIdx.x] = s1; block_sums[2 * G + blockIdx.x] = s2;
        __threadfence();
        s_last = atomicAdd(&ctrl->ticket, 1u) == G - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double se = strided_sum_block(block_sums, G), sk = strided_sum_block(block_sums + G, G), sp = strided_sum_block(block_sums + 2 * G, G);
    block_sum3(se, sk, sp, red);
    if (KIND == 0) {
        if (threadIdx.x == 0) {
            ctrl->ticket = 0;
            const bool due = __longlong_as_double((long long)ctrl->maxdisp2) > c.disp_thresh;
            ctrl->maxdisp2 = 0ull;
            publish_step(c, ctrl, hist, se, sk, sp, due, __longlong_as_double((long long)ctrl->vmax2));
        }
        return;
    }
    if (KIND == 2) {
        if (threadIdx.x == 0) {
            ctrl->ticket = 0;
            ctrl->red[0] = se; ctrl->red[1] = sk; ctrl->red[2] = sp;
            ctrl->red[3] = (__longlong_as_double((long long)ctrl->maxdisp2) > c.disp_thresh) ? 1.0 : 0.0;
            ctrl->maxdisp2 = 0ull;
        }
        return;
    }
    if (threadIdx.x == 0) {
        ctrl->ticket = 0;
        s_tot[0] = se; s_tot[1] = sk; s_tot[2] = sp;
        s_tot[3] = (__longlong_as_double((long long)ctrl->maxdisp2) > c.disp_thresh) ? 1.0 : 0.0;
        ctrl->maxdisp2 = 0ull;
        // the neighbours' fastest-atom values of this step arrived with the halo the force kernel waited for;
        // read them before my deposit releases the neighbours into the next step's push
        s_nbv2 = fmax(ld_sys_f64(&sr.mine->nb_vmax2[0]), ld_sys_f64(&sr.mine->nb_vmax2[1]));
    }
    __syncthreads();
    const unsigned long long seq = ctrl->red_seq + 1ull;      // (read by every thread before thread 0 stores it back below)
    mailbox_allreduce(s_tot, s_g, s_dep, sr.mine, sr.peers, sr.rank, sr.world, seq);
    if (threadIdx.x == 0) {
        ctrl->red_seq = seq;
        ctrl->red[0] = s_g[0]; ctrl->red[1] = s_g[1]; ctrl->red[2] = s_g[2]; ctrl->red[3] = s_g[3];
        // pruned inner list: pairs of this device involve its own atoms and the two neighbours' boundary atoms, so the
        // displacement bound advances by the fastest atom among the three devices
        publish_step(c, ctrl, hist, s_g[0], s_g[1], s_g[2], s_g[3] > 0.0,
                     fmax(__longlong_as_double((long long)ctrl->vmax2), s_nbv2));
    }
}

// After a pruning pass: the inner list is exact for the current positions.
__global__ void prune_commit_kernel(StepCtrl* __restrict__ ctrl) {
    if (ctrl->stop) return;
    ctrl->inner_valid = 1; ctrl->drift_bound = 0.0; ctrl->vmax2 = 0ull;
}

// Constructor path: energies after the first force pass; ekin stays 0 (.cpp:57-62).
__global__ void __launch_bounds__(kFinalizeThreads)
init_finalize_kernel(LJConst c, StepCtrl* __restrict__ ctrl, const double* __restrict__ partial_epot, unsigned nb_epot) {
    __shared__ double red[32];
    double se = strided_sum(partial_epot, nb_epot);
    se = block_sum(se, red);
    if (threadIdx.x == 0) {
        ctrl->epot_sum = se;
        ctrl->epot = xmul(xmul(xmul(se, 4.0), c.epsilon), 0.5);
    }
}

// ---------------------------------------------------------------------------------------------
// Re-binning kernels (K4 wrap, K5 cell id, K6 permutation after the radix sort).
// ---------------------------------------------------------------------------------------------
// wrap (.cpp:431-433) + cell key (.cpp:468-476) of every owned atom; unsorted wrapped copy to pos_tmp.
__global__ void __launch_bounds__(kBlock)
wrap_key_kernel(LJConst c, const double4* __restrict__ pos, double4* __restrict__ pos_tmp,
                uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const unsigned i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= c.n) return;
    double4 p = pos[i];
    p.x = wrap_coord(p.x, c.L, c.invL);
    p.y = wrap_coord(p.y, c.L, c.invL);
    p.z = wrap_coord(p.z, c.L, c.invL);
    pos_tmp[i] = p;
    const unsigned ix = cell_coord(p.x, c.cell_edge, c.nc);
    const unsigned iy = cell_coord(p.y, c.cell_edge, c.nc);
    const unsigned iz = cell_coord(p.z, c.cell_edge, c.nc);
    if (!c.slab) {
        keys[i] = iz * c.nc * c.nc + iy * c.nc + ix;
    } else {
        // local table: layer 0 = lower halo, 1..nzl = owned, nzl+1 = upper halo; atoms that left the slab
        // get the two keys past the table (down / up movers) and are handed to the neighbour device
        const unsigned lz = (iz + c.nc - c.zlo) % c.nc;
        const unsigned ncell_loc = c.nc * c.nc * (c.nzl + 2);
        if (lz < c.nzl) keys[i] = (lz + 1) * c.nc * c.nc + iy * c.nc + ix;
        else keys[i] = (lz == c.nzl) ? ncell_loc + 1 : ncell_loc;        // one layer above : anything else is "below"
    }
    vals[i] = i;
}

// ---- slab path: atom migration, halo cell table, split finalize -------------------------------------
constexpr int kMoverDoubles = 11;      // x y z vx vy vz fx fy fz orig(as bits) pad
__global__ void __launch_bounds__(kBlock)
pack_movers_kernel(unsigned first, unsigned cnt, const double4* __restrict__ pos,
                   const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ vz,
                   const double* __restrict__ fx, const double* __restrict__ fy, const double* __restrict__ fz,
                   const uint32_t* __restrict__ orig, double* __restrict__ buf) {
    const unsigned k = blockIdx.x * kBlock + threadIdx.x;
    if (k >= cnt) return;
    const unsigned i = first + k;
    const double4 p = pos[i];
    double* b = buf + (size_t)k * kMoverDoubles;
    b[0] = p.x; b[1] = p.y; b[2] = p.z;
    b[3] = vx[i]; b[4] = vy[i]; b[5] = vz[i];
    b[6] = fx[i]; b[7] = fy[i]; b[8] = fz[i];
    b[9] = __longlong_as_double((long long)orig[i]); b[10] = 0.0;
}

__global__ void __launch_bounds__(kBlock)
unpack_movers_kernel(unsigned first, unsigned cnt, const double* __restrict__ buf, double4* __restrict__ pos,
                     double* __restrict__ vx, double* __restrict__ vy, double* __restrict__ vz,
                     double* __restrict__ fx, double* __restrict__ fy, double* __restrict__ fz,
                     uint32_t* __restrict__ orig) {
    const unsigned k = blockIdx.x * kBlock + threadIdx.x;
    if (k >= cnt) return;
    const unsigned i = first + k;
    const double* b = buf + (size_t)k * kMoverDoubles;
    pos[i] = make_double4(b[0], b[1], b[2], 0.0);
    vx[i] = b[3]; vy[i] = b[4]; vz[i] = b[5];
    fx[i] = b[6]; fy[i] = b[7]; fz[i] = b[8];
    orig[i] = (uint32_t)__double_as_longlong(b[9]);
}

// cell table of one halo layer: counts received from the neighbour, slots appended after the owned atoms
__global__ void __launch_bounds__(kBlock)
halo_cells_kernel(unsigned ncell_layer, unsigned layer_first_cell, unsigned base_slot,
                  const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets,
                  uint32_t* __restrict__ cstart, uint32_t* __restrict__ ccount) {
    const unsigned k = blockIdx.x * kBlock + threadIdx.x;
    if (k >= ncell_layer) return;
    cstart[layer_first_cell + k] = base_slot + offsets[k];
    ccount[layer_first_cell + k] = counts[k];
}

// local part of the step's reductions (fixed order) -> ctrl->red, to be all-reduced over the devices
__global__ void __launch_bounds__(kFinalizeThreads)
slab_reduce_kernel(LJConst c, StepCtrl* __restrict__ ctrl,
                   const double* __restrict__ partial_epot, unsigned nb_epot,
                   const double* __restrict__ partial_ekin, const double* __restrict__ partial_pred, unsigned nb_kick) {
    __shared__ double red[32];
    double se = strided_sum(partial_epot, nb_epot);
    double sk = partial_ekin ? strided_sum(partial_ekin, nb_kick) : 0.0;
    double sp = strided_sum(partial_pred, nb_kick);
    se = block_sum(se, red);
    sk = block_sum(sk, red);
    sp = block_sum(sp, red);
    if (threadIdx.x == 0) {
        ctrl->red[0] = se; ctrl->red[1] = sk; ctrl->red[2] = sp;
        ctrl->red[3] = (__longlong_as_double((long long)ctrl->maxdisp2) > c.disp_thresh) ? 1.0 : 0.0;
        ctrl->maxdisp2 = 0ull;
    }
}

// Halo push: my two boundary layers are contiguous slot ranges; write them straight into the halo slots of
// the ring neighbours' position arrays (peer pointers), then publish the step sequence in their mailboxes.
__global__ void __launch_bounds__(256)
halo_push_kernel(StepCtrl* __restrict__ ctrl, const double4* __restrict__ pos, unsigned n, unsigned c_lo, unsigned c_hi,
                 double4* __restrict__ down_dst, double4* __restrict__ up_dst,
                 Mailbox* down_mail, Mailbox* up_mail, unsigned int* ticket) {
    if (ctrl->stop) return;
    const unsigned total = c_lo + c_hi;
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
        if (k < c_lo) st_pos(down_dst + k, ld_pos(pos + k));
        else st_pos(up_dst + (k - c_lo), ld_pos(pos + (n - c_hi) + (k - c_lo)));
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {          // last block: every store of every block is fenced
            *ticket = 0;
            // my fastest atom of this step travels with the halo: it bounds how far my boundary atoms moved
            const double v2 = __longlong_as_double((long long)ctrl->vmax2);
            down_mail->nb_vmax2[1] = v2;
            up_mail->nb_vmax2[0] = v2;
            __threadfence_system();
            const unsigned long long seq = ctrl->halo_seq + 1ull;       // (only this block touches the counter)
            ctrl->halo_seq = seq;
            st_sys_u64(&down_mail->halo_flag[1], seq);          // I am my lower neighbour's upper neighbour
            st_sys_u64(&up_mail->halo_flag[0], seq);
        }
    }
}

// Publication on the slab path (ljgpu_step_publish): the frame of positions in ORIGINAL atom order is cut into `world`
// contiguous slices of S atoms; device r assembles slice r.  Every device scatters the atoms it owns straight into the
// slice buffers of their owners over NVLink (8-byte peer stores), then tells every owner that its part is in; each owner
// then copies its slice to the host over its own PCIe link (3 x 8 x N / world bytes per device instead of 24 N on each).
struct PubTargets { double* slice[kMaxPeers]; Mailbox* mail[kMaxPeers]; };
__global__ void __launch_bounds__(256)
slab_publish_scatter_kernel(LJConst c, const double4* __restrict__ pos, const uint32_t* __restrict__ orig, unsigned S, unsigned nglobal,
                            PubTargets tgt, int rank, int world, unsigned long long seq, unsigned int* ticket) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < c.n; i += gridDim.x * blockDim.x) {
        const double4 p = pos[i];
        const unsigned o = orig[i];
        const unsigned d = o / S, off = o - d * S;
        const unsigned len = min(S, nglobal - d * S);           // atoms of slice d (the last one may be shorter)
        double* dst = tgt.slice[d];
        dst[off] = wrap_coord(p.x, c.L, c.invL);
        dst[len + off] = wrap_coord(p.y, c.L, c.invL);
        dst[2 * (size_t)len + off] = wrap_coord(p.z, c.L, c.invL);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {          // last block: every store of every block is fenced
            *ticket = 0;
            __threadfence_system();
            for (int r = 0; r < world; ++r) st_sys_u64(&tgt.mail[r]->pub_flag[rank], seq);
        }
    }
}
// copy stream of a slice owner: wait until every device has scattered its atoms of publication `seq` into my slice
__global__ void slab_publish_wait_kernel(Mailbox* mine, int world, unsigned long long seq) {
    if ((int)threadIdx.x < world) wait_flag(&mine->pub_flag[threadIdx.x], seq, mine);
}

// Local reductions + all-reduce through the mailboxes + publication, one launch: the constructor's energy
// (what = 1, epot only) and the predicted kinetic sum (what = 2).  The per-step variant lives in kick2_kernel<true>.
__global__ void __launch_bounds__(kFinalizeThreads)
slab_exchange_kernel(LJConst c, StepCtrl* __restrict__ ctrl, int what,
                     const double* __restrict__ partial_epot, unsigned nb_epot,
                     const double* __restrict__ partial_pred, unsigned nb_kick, SlabRed sr) {
    __shared__ double red[32];
    __shared__ double tot[4], g[4], dep[kMaxPeers * 4];
    double se = strided_sum(partial_epot, nb_epot);
    double sp = strided_sum(partial_pred, nb_kick);
    se = block_sum(se, red);
    sp = block_sum(sp, red);
    if (threadIdx.x == 0) { tot[0] = se; tot[1] = 0.0; tot[2] = sp; tot[3] = 0.0; }
    __syncthreads();
    const unsigned long long seq = ctrl->red_seq + 1ull;
    mailbox_allreduce(tot, g, dep, sr.mine, sr.peers, sr.rank, sr.world, seq);
    if (threadIdx.x == 0) {
        ctrl->red_seq = seq;
        ctrl->red[0] = g[0]; ctrl->red[1] = g[1]; ctrl->red[2] = g[2]; ctrl->red[3] = g[3];
        if (what == 2) { ctrl->pred_sum = g[2]; return; }
        ctrl->epot_sum = g[0];
        ctrl->epot = xmul(xmul(xmul(g[0], 4.0), c.epsilon), 0.5);
    }
}

// what: 0 = full step (energies, history, rebuild flag), 1 = constructor (epot only), 2 = predict (pred_sum only)
__global__ void slab_publish_kernel(LJConst c, StepCtrl* __restrict__ ctrl, HistEntry* __restrict__ hist, int what) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (what == 2) { ctrl->pred_sum = ctrl->red[2]; return; }
    const double se = ctrl->red[0];
    const double epot = xmul(xmul(xmul(se, 4.0), c.epsilon), 0.5);
    ctrl->epot_sum = se;
    if (what == 1) { ctrl->epot = epot; return; }
    if (ctrl->stop) return;
    const double ekin = xmul(xmul(0.5, c.mass), ctrl->red[1]);
    ctrl->pred_sum = ctrl->red[2];
    ctrl->ekin = ekin; ctrl->epot = epot; ctrl->etot = xadd(ekin, epot);
    HistEntry& h = hist[ctrl->hist_head % kHistCap];
    h.ekin = ekin; h.epot = epot; h.step = ctrl->next_step; h.pad = 0;
        ctrl->next_step += 1;
    ctrl->hist_head += 1;
    ctrl->steps_done += 1;
    if (ctrl->red[3] > 0.0) ctrl->stop = 1;
}

__global__ void __launch_bounds__(kBlock)
permute_kernel(unsigned n, const uint32_t* __restrict__ perm, const double4* __restrict__ pos_tmp,
               double4* __restrict__ pos,
               const double* __restrict__ vxi, const double* __restrict__ vyi, const double* __restrict__ vzi,
               const double* __restrict__ fxi, const double* __restrict__ fyi, const double* __restrict__ fzi,
               const uint32_t* __restrict__ origi,
               double* __restrict__ vxo, double* __restrict__ vyo, double* __restrict__ vzo,
               double* __restrict__ fxo, double* __restrict__ fyo, double* __restrict__ fzo,
               uint32_t* __restrict__ origo) {
    const unsigned i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const unsigned s = perm[i];
    pos[i] = pos_tmp[s];
    vxo[i] = vxi[s]; vyo[i] = vyi[s]; vzo[i] = vzi[s];
    fxo[i] = fxi[s]; fyo[i] = fyi[s]; fzo[i] = fzi[s];
    origo[i] = origi[s];
}

// Cell table from the sorted keys.
__global__ void __launch_bounds__(kBlock)
cell_bounds_kernel(unsigned n, const uint32_t* __restrict__ keys, uint32_t* __restrict__ cstart, uint32_t* __restrict__ cend) {
    const unsigned i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = keys[i];
    if (i == 0 || keys[i - 1] != k) cstart[k] = i;
    if (i == n - 1 || keys[i + 1] != k) cend[k] = i + 1;
}

__global__ void __launch_bounds__(kBlock)
cell_count_kernel(unsigned ncell, const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ cend, uint32_t* __restrict__ ccount) {
    const unsigned k = blockIdx.x * kBlock + threadIdx.x;
    if (k < ncell) ccount[k] = cend[k] - cstart[k];
}

// ---------------------------------------------------------------------------------------------
// Query kernels (original atom order out).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
export_positions_kernel(LJConst c, const double4* __restrict__ pos, const uint32_t* __restrict__ orig,
                        double* __restrict__ ox, double* __restrict__ oy, double* __restrict__ oz) {
    const unsigned i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= c.n) return;
    const double4 p = pos[i];
    const unsigned o = orig[i];
    ox[o] = wrap_coord(p.x, c.L, c.invL);
    oy[o] = wrap_coord(p.y, c.L, c.invL);
    oz[o] = wrap_coord(p.z, c.L, c.invL);
}

__global__ void __launch_bounds__(kBlock)
export_soa_kernel(unsigned n, const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ cc,
                  const uint32_t* __restrict__ orig, double* __restrict__ ox, double* __restrict__ oy, double* __restrict__ oz) {
    const unsigned i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const unsigned o = orig[i];
    ox[o] = a[i]; oy[o] = b[i]; oz[o] = cc[i];
}

__global__ void __launch_bounds__(kBlock)
import_state_kernel(unsigned n, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                    double4* __restrict__ pos, uint32_t* __restrict__ orig) {
    const unsigned i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    pos[i] = make_double4(x[i], y[i], z[i], 0.0);
    orig[i] = i;
}

// values of the device's atoms to their original indices (query kernels of integer per-atom data)
__global__ void __launch_bounds__(kBlock)
scatter_u32_kernel(unsigned n, const uint32_t* __restrict__ src, const uint32_t* __restrict__ orig, uint32_t* __restrict__ out) {
    const unsigned i = blockIdx.x * kBlock + threadIdx.x;
    if (i < n) out[orig[i]] = src[i];
}

// Neighbour-count probe: cell key (.cpp:468-476) of every atom of the whole system from already wrapped positions in
// original order; vals = original index.
__global__ void __launch_bounds__(kBlock)
probe_key_kernel(LJConst c, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                 double4* __restrict__ pos_out, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const unsigned i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= c.n) return;
    const double px = x[i], py = y[i], pz = z[i];
    pos_out[i] = make_double4(px, py, pz, 0.0);
    const unsigned ix = cell_coord(px, c.cell_edge, c.nc), iy = cell_coord(py, c.cell_edge, c.nc), iz = cell_coord(pz, c.cell_edge, c.nc);
    keys[i] = iz * c.nc * c.nc + iy * c.nc + ix;
    vals[i] = i;
}
__global__ void __launch_bounds__(kBlock)
gather_pos_kernel(unsigned n, const uint32_t* __restrict__ perm, const double4* __restrict__ in, double4* __restrict__ out) {
    const unsigned i = blockIdx.x * kBlock + threadIdx.x;
    if (i < n) out[i] = in[perm[i]];
}

// GraphWidget statistics on the device (graphwidget.cpp:139-203, :235-248): minimum, maximum, mean and standard deviation
// of the sampled kinetic / potential / total energy series.  One lane per series walks the ring oldest-first with the
// reference's operations in the reference's order (no contraction), so the numbers equal a host evaluation bit for bit.
// reference_truncation reproduces the reference's std::accumulate(..., 0): an int accumulator truncates every partial sum
// (:147, :169, :191, :236); otherwise the sum is done in double, as the plot evidently intends.
struct SeriesStatsOut { double mean[3], stddev[3], vmin[3], vmax[3]; unsigned int n; unsigned int pad; };
__global__ void series_stats_kernel(const StepCtrl* __restrict__ ctrl, int reference_truncation, SeriesStatsOut* __restrict__ out) {
    const unsigned k = threadIdx.x;
    if (k >= 3) return;
    const unsigned long long head = ctrl->sample_head;
    const unsigned n = (unsigned)(head < (unsigned long long)kSampleCap ? head : (unsigned long long)kSampleCap);
    const HistEntry* ring = ctrl->samples;
    auto value = [&](unsigned idx) {
        const HistEntry& e = ring[(head - n + idx) % kSampleCap];
        return k == 0 ? e.ekin : (k == 1 ? e.epot : xadd(e.ekin, e.epot));
    };
    if (k == 0) { out->n = n; out->pad = 0; }
    if (n == 0) { out->mean[k] = 0.0; out->stddev[k] = 0.0; out->vmin[k] = 0.0; out->vmax[k] = 0.0; return; }
    double sum = 0.0, lo = value(0), hi = lo;
    int isum = 0;
    for (unsigned idx = 0; idx < n; ++idx) {
        const double v = value(idx);
        if (reference_truncation) isum = (int)xadd((double)isum, v); else sum = xadd(sum, v);
        lo = v < lo ? v : lo; hi = v > hi ? v : hi;
    }
    const double avg = xdiv(reference_truncation ? (double)isum : sum, (double)n);
    double var = 0.0;
    for (unsigned idx = 0; idx < n; ++idx) { const double d = xsub(value(idx), avg); var = xadd(var, xmul(d, d)); }
    var = xdiv(var, (double)(n - 1));                       // graphwidget.cpp:241
    if (!(var - var == 0.0)) var = 0.0;                     // not finite (one sample): graphwidget.cpp:243-245
    out->mean[k] = avg; out->stddev[k] = xsqrt(var); out->vmin[k] = lo; out->vmax[k] = hi;
}

// Maxwell-Boltzmann overlay of the speed histogram (graphwidget.cpp:55-73: f(v) N binsize at the left bin edges) and the
// chi-square of the device histogram against it over the bins that expect more than five atoms.
__global__ void __launch_bounds__(256)
speed_mb_kernel(unsigned nbins, double vmax, double kT, double mass, double natoms,
                const unsigned long long* __restrict__ counts, double* __restrict__ expected, double* __restrict__ chi2_out) {
    __shared__ double red[32];
    __shared__ double red2[32];
    double chi = 0.0, used = 0.0;
    const double binsize = vmax / (double)nbins;
    const double pref = 4.0 * 3.14159265358979323846 * pow(mass / (2.0 * 3.14159265358979323846 * kT), 1.5);
    for (unsigned b = threadIdx.x; b < nbins; b += blockDim.x) {
        const double v = binsize * (double)b;
        const double e = pref * v * v * exp(-mass * v * v / (2.0 * kT)) * natoms * binsize;
        expected[b] = e;
        if (e > 5.0) { const double d = (double)counts[b] - e; chi += d * d / e; used += 1.0; }
    }
    chi = block_sum(chi, red);
    used = block_sum(used, red2);
    if (threadIdx.x == 0) { chi2_out[0] = chi; chi2_out[1] = used; }
}

// FP64 roofline denominator, measured in place (ljgpu_measure_fp64_peak): eight independent DFMA chains per thread.
__global__ void __launch_bounds__(256)
fp64_peak_kernel(double* __restrict__ out, int iters, double a, double b) {
    double acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = threadIdx.x * 1e-9 + k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fma(acc[k], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// K8  speed histogram, graphwidget.cpp:225-230: bin = min((int)(|v| / vmax * nbins), nbins-1),
//     |v| = sqrt(x*x + y*y + z*z) (glm::length).  Shared-memory integer atomics, then one global
//     atomic per non-empty bin per block.
constexpr int kMaxBins = 1024;
__global__ void __launch_bounds__(256)
speed_histogram_kernel(unsigned n, const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ vz,
                       unsigned nbins, double vmax, unsigned long long* __restrict__ counts) {
    __shared__ unsigned int h[kMaxBins];
    for (unsigned b = threadIdx.x; b < nbins; b += blockDim.x) h[b] = 0;
    __syncthreads();
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double speed = xsqrt(xnorm2(vx[i], vy[i], vz[i]));
        int bin = __double2int_rz(xmul(xdiv(speed, vmax), (double)nbins));
        bin = min(bin, (int)nbins - 1);
        bin = max(bin, 0);
        atomicAdd(&h[bin], 1u);
    }
    __syncthreads();
    for (unsigned b = threadIdx.x; b < nbins; b += blockDim.x)
        if (h[b]) atomicAdd(&counts[b], (unsigned long long)h[b]);
}

// Cell-id probe on the published (wrapped) positions, .cpp:447-476, original order out.
__global__ void __launch_bounds__(kBlock)
cell_id_probe_kernel(LJConst c, const double4* __restrict__ pos, const uint32_t* __restrict__ orig, uint32_t* __restrict__ out) {
    const unsigned i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= c.n) return;
    const double4 p = pos[i];
    const unsigned ix = cell_coord(wrap_coord(p.x, c.L, c.invL), c.cell_edge, c.nc);
    const unsigned iy = cell_coord(wrap_coord(p.y, c.L, c.invL), c.cell_edge, c.nc);
    const unsigned iz = cell_coord(wrap_coord(p.z, c.L, c.invL), c.cell_edge, c.nc);
    out[orig[i]] = iz * c.nc * c.nc + iy * c.nc + ix;
}

// Neighbour-count probe on the all-pairs layout (no cell table): every j through shared memory,
// published (wrapped) positions, the reference's comparisons (.cpp:519-536 / :291-305).
__global__ void __launch_bounds__(kBlock)
count_allpairs_kernel(LJConst c, const double4* __restrict__ pos, const uint32_t* __restrict__ orig,
                      double cut2, int inclusive, uint32_t* __restrict__ out) {
    __shared__ double sx[kApTile], sy[kApTile], sz[kApTile];
    const unsigned i = blockIdx.x * kBlock + threadIdx.x;
    const bool active = i < c.n;
    double4 pi = active ? pos[i] : make_double4(0, 0, 0, 0);
    pi.x = wrap_coord(pi.x, c.L, c.invL); pi.y = wrap_coord(pi.y, c.L, c.invL); pi.z = wrap_coord(pi.z, c.L, c.invL);
    unsigned cnt = 0;
    for (unsigned jb = 0; jb < c.n; jb += kApTile) {
        __syncthreads();
        if (jb + threadIdx.x < c.n) {
            const double4 q = pos[jb + threadIdx.x];
            sx[threadIdx.x] = wrap_coord(q.x, c.L, c.invL);
            sy[threadIdx.x] = wrap_coord(q.y, c.L, c.invL);
            sz[threadIdx.x] = wrap_coord(q.z, c.L, c.invL);
        }
        __syncthreads();
        const unsigned m = min((unsigned)kApTile, c.n - jb);
        if (active)
            for (unsigned jj = 0; jj < m; ++jj) {
                const double dx = min_image(xsub(pi.x, sx[jj]), c.L, c.halfL);
                const double dy = min_image(xsub(pi.y, sy[jj]), c.L, c.halfL);
                const double dz = min_image(xsub(pi.z, sz[jj]), c.L, c.halfL);
                const double d2 = xnorm2(dx, dy, dz);
                const bool in = inclusive ? (d2 <= cut2) : (d2 < cut2);
                if (in && (jb + jj) != i) ++cnt;
            }
    }
    if (active) out[orig[i]] = cnt;
}

} // namespace ljgpu
