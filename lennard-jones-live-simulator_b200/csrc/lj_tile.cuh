// lj_tile.cuh - the persistent brick kernel: Verlet-list force, cell-list force and the list build.
//
// Why tiles: a per-thread gather of neighbour positions from L1/L2 costs one cache line per lane per
// load (measured: the first Verlet kernel ran at exactly one 128-byte line per clock per SM, 85-90 %
// of the L1 data pipe, FP64 pipe a third busy - profiles/r01b_*).  Shared memory serves 32 arbitrary
// addresses per clock.  So the box is cut into bricks of kTileBX x kTileBY x kTileBZ cells; the positions of
// a brick plus one halo cell on every side ("tile") are staged in shared memory once, and every home atom
// walks its list as 16-bit indices into that tile.
//
// Why persistent and warp-specialised (round 2).  The list walk is latency-bound: nothing is saturated (FP64 pipe
// 45 %, issue slots 55 %, shared-memory pipe 63 %), a warp inside the walk issues an instruction every ~6 cycles, a work
// item of 32 atoms takes ~10 us whatever the load, and the pass time follows the number of warps that are inside the
// walk at any moment (profiles/r02b_*, r02t_*, r03c_*).  With one CTA per brick only ~11 of the 24 resident warps were:
// the rest read tables, waited at barriers, staged, belonged to the 8 % of bricks with more home atoms than threads
// (a second, almost empty round) or to the partial bricks at the box edge.  Now ONE CTA per SM stays resident and
// streams bricks - handed out by a device-wide queue - through a ring of up to four tiles: four staging warps fill the
// slots ahead (asynchronous global->shared copies, LDGSTS; release / acquire counters per ring) and sixteen computing
// warps pull work items - 32 consecutive home atoms of a brick - from a CTA-wide counter, so every warp walks lists
// practically all the time, whatever the shapes of the bricks.
//
// Reference code replaced: calculate_forces (.cpp:258-330) and build_neighbor_list (.cpp:441-551) of
// /root/reference/src/simulator/lennardjonessimulation.cpp.
#pragma once
#include "lj_kernels.cuh"

namespace ljgpu {

#ifndef LJ_TILE_BX
#define LJ_TILE_BX 3
#endif
#ifndef LJ_BRICK_WARPS
#define LJ_BRICK_WARPS 16
#endif
#ifndef LJ_TILE_GROUP
#define LJ_TILE_GROUP 4
#endif
constexpr int kTileBX = LJ_TILE_BX;                    // home cells per brick along x (contiguous slots)
constexpr int kTileBY = 2;
constexpr int kTileBZ = 2;
constexpr int kTileRowsMax = (kTileBY + 2) * (kTileBZ + 2);           // x-rows in a tile
constexpr int kTileCellsMax = (kTileBX + 2) * kTileRowsMax;           // tile cells per brick
constexpr int kHomeRowsMax = kTileBY * kTileBZ;
constexpr int kTileGroup = LJ_TILE_GROUP;                             // list entries evaluated as one batch (4 or 8)
constexpr int kBrickWarps = LJ_BRICK_WARPS;                           // computing warps per CTA
#ifndef LJ_BRICK_STAGERS
#define LJ_BRICK_STAGERS 4
#endif
constexpr int kBrickStagers = LJ_BRICK_STAGERS;                       // staging warps per CTA (they share the rows of a tile)
constexpr int kBrickThreads = 32 * (kBrickWarps + kBrickStagers);
constexpr int kBrickSlotsMax = 4;                                     // tiles in flight per CTA (fewer when tiles are large)
#ifndef LJ_FILL_BATCH
#define LJ_FILL_BATCH 4
#endif
constexpr int kFillBatch = LJ_FILL_BATCH;                            // candidates per round of the list build (measured: 4 -> 1005 us per re-binning at N = 2^20, 8 -> 1093, 16 spills)
constexpr int kBrickItemsMax = 32;                                    // work items (32 home atoms each) per brick

// Shared-memory image of a tile: (x, y) pairs as 16-byte records followed by the z column.  The list walk is a
// random gather and the shared-memory data pipe is its busiest unit; a 128-bit load moves 16 bytes per lane in
// four quarter-warp phases over eight bank quads, the z column is a 64-bit load over sixteen bank pairs -
// together ~16.6 wavefronts per gathered position where three 64-bit loads out of 24-byte records took 19.5.
constexpr unsigned kTileRecBytes = 24u;
__device__ __forceinline__ void tile_ld(const double* sxy, const double* sz, unsigned j, double& x, double& y, double& z) {
    const double2 xy = reinterpret_cast<const double2*>(sxy)[j];
    x = xy.x; y = xy.y;
    z = sz[j];
}

// A list under construction: eight 16-bit entries are collected in registers (the quad is shifted down by one entry, the new
// one enters at the top: after eight pushes the first entry sits in half-word 0, so the walk sees the entries in the order
// they were found and a filtered list sums in the same order as its parent - bit-identical forces) and leave as ONE 128-bit
// store.  Entry-by-entry 16-bit stores cost a 32-byte L2 sector write each: 84 M of them per list build at N = 2^20.
struct QuadPacker {
    uint4 q;
    unsigned cnt;
    __device__ __forceinline__ void push(unsigned j, uint4* __restrict__ row, unsigned capq) {
        q.x = __funnelshift_r(q.x, q.y, 16); q.y = __funnelshift_r(q.y, q.z, 16);
        q.z = __funnelshift_r(q.z, q.w, 16); q.w = (q.w >> 16) | (j << 16);
        ++cnt;
        if ((cnt & 7u) == 0u && cnt <= 8u * capq) row[(size_t)((cnt >> 3) - 1u) * 32] = q;
    }
    // the unused tail of the last quad points at the atom itself
    __device__ __forceinline__ void finish(unsigned self, uint4* __restrict__ row, unsigned capq) {
        unsigned k = cnt;
        if (k > 8u * capq || (k & 7u) == 0u) return;
        const unsigned at = k >> 3;
        for (; k & 7u; ++k) {
            q.x = __funnelshift_r(q.x, q.y, 16); q.y = __funnelshift_r(q.y, q.z, 16);
            q.z = __funnelshift_r(q.z, q.w, 16); q.w = (q.w >> 16) | (self << 16);
        }
        row[(size_t)at * 32] = q;
    }
};

#ifdef LJ_BRICK_TRACE   /* developer build only (scripts/brick_trace.py): a timeline of the brick kernel's last launch */
constexpr unsigned kTraceCap = 2048;                       // events per CTA
struct TraceEvent { unsigned long long t; unsigned kind, a, b, warp; };     // kind: 0 cta start, 1 ticket, 2 tile staged, 3 item start, 4 item end
__device__ TraceEvent g_trace[160][kTraceCap];
__device__ unsigned g_trace_n[160];
__device__ __forceinline__ void trace_event(unsigned kind, unsigned a, unsigned b) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (blockIdx.x >= 160u) return;
    const unsigned k = atomicAdd(&g_trace_n[blockIdx.x], 1u);
    if (k < kTraceCap) g_trace[blockIdx.x][k] = TraceEvent{t, kind, a, b, threadIdx.x >> 5};
}
#define LJ_TRACE(kind, a, b) do { if (MODE == TILE_FORCE_LIST && (threadIdx.x & 31u) == 0u) trace_event(kind, a, b); } while (0)
#else
#define LJ_TRACE(kind, a, b) do { } while (0)
#endif

struct TileGrid {
    unsigned nc;                 // cells per axis in x and y (periodic)
    unsigned ncz;                // cell layers of the local table (nc, or owned layers + 2 halo layers on a slab)
    unsigned zper;               // z periodic within this table (single device) or bounded by halo layers (slab)
    unsigned hz0, nhz;           // first home layer and number of home layers
    unsigned bxd, byd, bzd;      // brick size in cells (<= kTileBX, kTileBY, kTileBZ; smaller for small systems)
    unsigned nbx, nby, nbz;      // bricks per axis
    unsigned nblocks;            // bricks
    const uint32_t* __restrict__ tsrc;   // [nblocks][kTileCellsMax]     first global slot of each tile cell
    const uint32_t* __restrict__ toff;   // [nblocks][kTileCellsMax + 1] tile-local offset of each tile cell
};

struct TileGeom { unsigned x0, y0, z0, bw, bh, bd, W, H; };

// Brick b -> its cells.  Tile cell numbering: t = ((tz*(bh+2)) + ty)*(bw+2) + tx with one halo cell on
// every side (tx = 0 and bw+1 etc.), so the cells of one x-row are consecutive in the tile.
__device__ __forceinline__ TileGeom tile_geometry(unsigned b, const TileGrid& g) {
    TileGeom t;
    const unsigned bx = b % g.nbx, by = (b / g.nbx) % g.nby;
    unsigned bz = b / (g.nbx * g.nby);
    if (!g.zper && g.nbz > 2)        // slab: the two brick layers that touch halo layers are scheduled last (they wait for the neighbours)
        bz = bz < g.nbz - 2 ? bz + 1 : (bz == g.nbz - 2 ? 0u : g.nbz - 1);
    t.x0 = bx * g.bxd; t.y0 = by * g.byd; t.z0 = g.hz0 + bz * g.bzd;
    t.bw = min(g.bxd, g.nc - t.x0);
    t.bh = min(g.byd, g.nc - t.y0);
    t.bd = min(g.bzd, g.hz0 + g.nhz - t.z0);
    t.W = t.bw + 2; t.H = t.bh + 2;
    return t;
}

// Re-binning: per brick, where each tile cell's atoms live (global slot) and where they go in the tile.
__global__ void __launch_bounds__(128)
tile_table_kernel(TileGrid g, const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ ccount,
                  uint32_t* __restrict__ tsrc, uint32_t* __restrict__ toff, StepCtrl* __restrict__ ctrl) {
    const int lane = threadIdx.x & 31;
    const unsigned b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= g.nblocks) return;
    const TileGeom t = tile_geometry(b, g);
    const unsigned nc = g.nc, ntc = t.W * t.H * (t.bd + 2);
    unsigned run = 0, home = 0;
    for (unsigned t0 = 0; t0 < ntc; t0 += 32) {
        const unsigned k = t0 + lane;
        unsigned cnt = 0, src = 0;
        bool is_home = false;
        if (k < ntc) {
            const unsigned tx = k % t.W, ty = (k / t.W) % t.H, tz = k / (t.W * t.H);
            const unsigned cx = (t.x0 + nc + tx - 1) % nc, cy = (t.y0 + nc + ty - 1) % nc;
            const unsigned cz = g.zper ? (t.z0 + g.ncz + tz - 1) % g.ncz : (t.z0 + tz - 1);   // a slab always has a layer below and above
            const unsigned cid = cz * nc * nc + cy * nc + cx;
            cnt = ccount[cid]; src = cstart[cid];
            is_home = tx >= 1 && tx <= t.bw && ty >= 1 && ty <= t.bh && tz >= 1 && tz <= t.bd;
        }
        unsigned inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        if (k < ntc) {
            tsrc[(size_t)b * kTileCellsMax + k] = src;
            toff[(size_t)b * (kTileCellsMax + 1) + k] = run + inc - cnt;
        }
        run += __shfl_sync(0xffffffffu, inc, 31);
        unsigned hc = is_home ? cnt : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) hc += __shfl_xor_sync(0xffffffffu, hc, o);
        home += hc;
    }
    if (lane == 0) {
        toff[(size_t)b * (kTileCellsMax + 1) + ntc] = run;
        atomicMax(&ctrl->max_tile, run);
        atomicMax(&ctrl->max_home, home);
    }
}

// 16-bit list entries, eight per 128-bit quad; groups of 32 consecutive slots are interleaved so that
// one warp-wide 128-bit load fetches 512 contiguous bytes: quad q of slot i starts at 32-bit word
// (((i>>5)*capq + q)*32 + (i&31))*4.  Entry k of slot i is half-word k&7 of quad k>>3.
// Replaces the reference's flat [offsets][entries..., sentinel] vector (.cpp:543-550).

enum { TILE_FORCE_LIST = 0, TILE_FORCE_CELL = 1, TILE_LIST_COUNT = 2, TILE_LIST_FILL = 3,
       TILE_FORCE_PRUNE = 5 /* force pass over the outer list that also emits the pruned inner list */ };

struct TileArgs {
    const double4* __restrict__ pos;
    uint32_t* __restrict__ nl32;           // list words (fill: written as uint16 halves)
    uint32_t* __restrict__ nl_count;
    uint32_t* __restrict__ in32;           // pruned inner list, same layout (null: none)
    uint32_t* __restrict__ in_count;
    double* __restrict__ fx; double* __restrict__ fy; double* __restrict__ fz;
    double* __restrict__ partial_epot;     // one per (brick, work item)
    StepCtrl* ctrl;
    unsigned capq;                         // list capacity in 128-bit quads (8 entries) per atom
    unsigned tcap;                         // atoms per tile slot (shared-memory capacity)
    unsigned slots;                        // tiles in flight per CTA (2..kBrickSlotsMax)
    unsigned ipb;                          // work items per brick: ceil(most home atoms of a brick / 32)
    unsigned lookahead;                    // a brick is requested once the computing warps have claimed all but this many of the work
                                           // items before it (0: only when a warp already waits for it) - see the staging loop
    // slab path over peer memory: wait until both neighbours have pushed this step's halo (null: no wait)
    Mailbox* mail;                         // (the sequence number to wait for is ctrl->halo_seq: the push of this step has already counted itself)
};

// What the staging warp leaves next to a tile for the computing warps.
struct BrickHeader {
    unsigned brick;                        // which brick the slot holds
    unsigned hn;                           // home atoms of the brick
    unsigned W, H, bw, bh, bd;
    unsigned lo_halo, hi_halo;             // slab: the tile reaches the lower / upper halo layer
    unsigned hstart[kHomeRowsMax + 1];     // prefix of home atoms per home x-row
};

// UNIT_SIGMA: sigma == 1 exactly (reduced units): s2 = sigma^2/d2 needs no multiply.
template <int MODE, bool UNIT_SIGMA>
__global__ void __launch_bounds__(kBrickThreads, 1)
tile_kernel(LJConst c, TileGrid g, TileArgs a, int guarded) {
    LJ_DYNAMIC_SMEM(sm);                                // a.slots tiles: [slot][ (x,y) x tcap | z x tcap ]
    __shared__ uint32_t s_off[kBrickSlotsMax][kTileCellsMax + 1];
    __shared__ uint32_t s_src[kBrickSlotsMax][kTileCellsMax];
    __shared__ BrickHeader s_hdr[kBrickSlotsMax];
    __shared__ unsigned s_filled;                       // bricks staged so far (written by the staging warp, release)
    __shared__ unsigned s_drained[kBrickSlotsMax];      // how many bricks have been worked off in each slot (release)
    __shared__ unsigned s_done[kBrickSlotsMax];         // work items of the slot's brick that are finished
    __shared__ unsigned s_item;                         // next work item of this CTA
    __shared__ unsigned s_ticket[2];                    // the brick the staging warps work on next (double-buffered by k & 1)
    __shared__ unsigned s_total;                        // bricks this CTA got before the queue ran dry (release; ~0u until then)
    if (guarded && a.ctrl->stop) return;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kBrickSlotsMax; ++s) { s_drained[s] = 0; s_done[s] = 0; }
        s_item = 0; s_filled = 0; s_total = 0xffffffffu;
    }
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const size_t slot_doubles = (size_t)a.tcap * 3;
#ifdef LJ_BRICK_TRACE
    if (MODE == TILE_FORCE_LIST && threadIdx.x == 0) g_trace_n[blockIdx.x] = 0;
    __syncthreads();
    if (warp == 0) LJ_TRACE(0, 0, 0);
#endif
    // Bricks are handed out by a device-wide counter (ctrl->brick_next): bricks differ in size (partial bricks at the box
    // edge, density fluctuations), and with a fixed brick -> CTA assignment the slowest SM ran 13 % longer than the
    // average one (ncu: sm__cycles_active min / avg / max 368 k / 433 k / 492 k; with the queue 372 k / 399 k / 414 k).
    // Results do not depend on who runs a brick: forces are per atom, energy partials are indexed by (brick, item).

    if (warp >= (unsigned)kBrickWarps) {
        // ---- staging warps: fill the ring ahead of the computing warps (the first one also reads the brick's tables) ----
        const unsigned stager = warp - (unsigned)kBrickWarps;
        for (unsigned k = 0;; ++k) {
            const unsigned slot = k % a.slots, use = k / a.slots;
            // The slot is free (every work item of the brick that was there is finished) AND the computing warps are about to
            // need another brick.  Taking tickets as early as the ring allows would hoard up to four bricks per CTA: at the end
            // of the queue (and from the very start when there are only a few bricks per CTA) some SMs would then sit on
            // several unstarted bricks while others idle.
            while (ld_acquire_smem(&s_drained[slot]) < use ||
                   (int)(k * a.ipb) - (int)ld_acquire_smem(&s_item) > (int)a.lookahead) spin_pause_long();
            if (lane == 0 && stager == 0) s_ticket[k & 1u] = atomicAdd(&a.ctrl->brick_next, 1u);
            if (kBrickStagers > 1) named_barrier_sync(3, 32u * kBrickStagers); else __syncwarp();
            const unsigned b = s_ticket[k & 1u];

            if (stager == 0) LJ_TRACE(1, k, b);
            if (b >= g.nblocks) {           // the queue is dry: tell the computing warps how many bricks there were
                if (lane == 0 && stager == 0) {
                    st_release_smem(&s_total, k);
                    // the last CTA to get here re-arms the queue for the next launch (nobody takes a ticket after its dry one)
                    if (atomicAdd(&a.ctrl->brick_dry, 1u) + 1u == gridDim.x) { a.ctrl->brick_dry = 0u; a.ctrl->brick_next = 0u; }
                }
                break;
            }
            const TileGeom tg = tile_geometry(b, g);
            const bool lo_halo = a.mail && tg.z0 == g.hz0, hi_halo = a.mail && tg.z0 + tg.bd == g.hz0 + g.nhz;
            if (lo_halo || hi_halo) {    // slab over peer memory: only bricks whose tile reaches a halo layer wait for that neighbour's push
                if (lane == 0) {
                    const unsigned long long seq = a.ctrl->halo_seq;
                    if (lo_halo) wait_flag(&a.mail->halo_flag[0], seq, a.mail);
                    if (hi_halo) wait_flag(&a.mail->halo_flag[1], seq, a.mail);
                }
                __syncwarp();
            }
            const unsigned W = tg.W, H = tg.H, ntc = W * H * (tg.bd + 2);
            for (unsigned t = lane + 32u * stager; t <= ntc; t += 32u * kBrickStagers) {
                s_off[slot][t] = g.toff[(size_t)b * (kTileCellsMax + 1) + t];
                if (t < ntc) s_src[slot][t] = g.tsrc[(size_t)b * kTileCellsMax + t];
            }
            if (kBrickStagers > 1) named_barrier_sync(1, 32u * kBrickStagers); else __syncwarp();
            if (stager == 0) LJ_TRACE(5, k, b);
            if (lane == 0 && stager == 0) {
                s_hdr[slot].brick = b;
                BrickHeader& h = s_hdr[slot];
                h.W = W; h.H = H; h.bw = tg.bw; h.bh = tg.bh; h.bd = tg.bd;
                h.lo_halo = lo_halo ? 1u : 0u; h.hi_halo = hi_halo ? 1u : 0u;
                unsigned run = 0, q = 0;
                for (unsigned hz = 1; hz <= tg.bd; ++hz)
                    for (unsigned hy = 1; hy <= tg.bh; ++hy) {
                        const unsigned r = (hz * H + hy) * W;
                        h.hstart[q++] = run;
                        run += s_off[slot][r + 1 + tg.bw] - s_off[slot][r + 1];
                    }
                for (; q <= (unsigned)kHomeRowsMax; ++q) h.hstart[q] = run;
                h.hn = run;
            }
            // one tile x-row (W consecutive tile cells, ~100 atoms at liquid density) at a time, lanes stride over
            // the row's atoms.  An atom's global slot is its tile index plus the (source - offset) difference of its
            // cell, picked by comparing against the row's W-1 inner cell offsets held in registers.
            // Asynchronous global->shared copies (LDGSTS): (x, y) are the first 16 bytes of the 32-byte global
            // record, z the next 8.
            {
                constexpr int kW = kTileBX + 2;
                const unsigned nrows = H * (tg.bd + 2);
                const unsigned sxy_base = smem_addr(sm + slot * slot_doubles);
                const unsigned sz_base = sxy_base + a.tcap * 16u;
                for (unsigned r = stager; r < nrows; r += (unsigned)kBrickStagers) {
                    const unsigned t0 = r * W;
                    unsigned offk[kW], delk[kW];
#pragma unroll
                    for (int kk = 0; kk < kW; ++kk) {
                        const bool on = (unsigned)kk < W;
                        const unsigned o = on ? s_off[slot][t0 + kk] : 0xffffffffu;
                        offk[kk] = o;
                        delk[kk] = on ? s_src[slot][t0 + kk] - o : 0u;      // modulo 2^32: li + delk = source slot
                    }
                    const unsigned o1 = s_off[slot][t0 + W];
                    for (unsigned li = offk[0] + lane; li < o1; li += 32) {
                        unsigned d = delk[0];
#pragma unroll
                        for (int kk = 1; kk < kW; ++kk) d = (li >= offk[kk]) ? delk[kk] : d;
                        const double* gp = reinterpret_cast<const double*>(a.pos + (li + d));
                        cp_async16(sxy_base + li * 16u, gp);
                        cp_async8(sz_base + li * 8u, gp + 2);
                    }
                }
                if (stager == 0) LJ_TRACE(6, k, b);
                cp_async_wait_all();
                if (stager == 0) LJ_TRACE(7, k, b);
            }
            // every staging lane's copies have landed, its table words are written: publish the tile
            if (kBrickStagers > 1) named_barrier_sync(2, 32u * kBrickStagers); else __syncwarp();
            if (lane == 0 && stager == 0) st_release_smem(&s_filled, k + 1u);
            if (stager == 0) LJ_TRACE(2, k, b);
        }
        return;
    }

    // ---- computing warps: pull work items (brick k of this CTA, home atoms 32 c .. 32 c + 31) until none is left -------
    const int far2_hi = __double2hiint(c.far2);
    unsigned maxcnt = 0;
    // Pruned inner list (entries within rcut + s_in at the last pruning): usable while no pair can have come
    // closer than s_in since, i.e. while the summed per-step maximum displacement stays below s_in / 2.
    // Same decision in every brick of a device (slab: per neighbour): the inputs were final before this launch,
    // so they are read once per warp, not once per work item.
    double own_v2 = 0.0, drift_bound = 0.0;
    bool inner_ok = false, use_inner_own = false;
    if (MODE == TILE_FORCE_LIST && a.in32 != nullptr) {   // (TILE_FORCE_PRUNE always walks the outer list)
        own_v2 = __longlong_as_double((long long)a.ctrl->vmax2);
        drift_bound = a.ctrl->drift_bound;
        inner_ok = a.ctrl->inner_valid != 0u;
        use_inner_own = inner_ok && drift_bound + sqrt(own_v2) * c.dt <= c.prune_half;
    }
    for (;;) {
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(&s_item, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        const unsigned k = item / a.ipb, chunk = item - k * a.ipb;
        const unsigned slot = k % a.slots, use = k / a.slots;
        bool dry = false;
        while (ld_acquire_smem(&s_filled) <= k) {
            if (ld_acquire_smem(&s_total) <= k) { dry = true; break; }     // there is no brick k: this CTA is done
            spin_pause();
        }
        if (dry) break;
        const BrickHeader& hd = s_hdr[slot];
        if (chunk * 32u < hd.hn) LJ_TRACE(3, k, chunk);
        const uint32_t* __restrict__ soff = s_off[slot];
        const uint32_t* __restrict__ ssrc = s_src[slot];
        const double* __restrict__ sxy = sm + slot * slot_doubles;
        const double* __restrict__ sz = sxy + (size_t)a.tcap * 2;
        const unsigned W = hd.W, H = hd.H, hn = hd.hn, b = hd.brick;
        bool use_inner = use_inner_own;
        if (MODE == TILE_FORCE_LIST && a.in32 != nullptr) {
            if (a.mail && (hd.lo_halo | hd.hi_halo)) {   // slab: a brick that reaches a halo layer also sees that neighbour's atoms (value pushed with the halo)
                double v2 = own_v2;
                if (hd.lo_halo) v2 = fmax(v2, ld_sys_f64(&a.mail->nb_vmax2[0]));
                if (hd.hi_halo) v2 = fmax(v2, ld_sys_f64(&a.mail->nb_vmax2[1]));
                use_inner = inner_ok && drift_bound + sqrt(v2) * c.dt <= c.prune_half;
            }
            if (use_inner && b == 0 && chunk == 0 && lane == 0) atomicAdd(&a.ctrl->inner_used, 1u);
        }
        double etot = 0.0;

        {
            const unsigned t = chunk * 32u + lane;
            if (t < hn) {
        // home row of this atom -> global slot i and tile-local index li
        unsigned q = 0;
#pragma unroll
        for (int kk = 1; kk < kHomeRowsMax; ++kk) q += (hd.hstart[kk] <= t) ? 1u : 0u;
        const unsigned hy = 1 + q % hd.bh, hz = 1 + q / hd.bh;
        const unsigned rbase = (hz * H + hy) * W;                 // tile cell index of (tx=0, hy, hz)
        const unsigned li = soff[rbase + 1] + (t - hd.hstart[q]);
        const unsigned i = ssrc[rbase + 1] + (t - hd.hstart[q]);
        double pix, piy, piz;
        tile_ld(sxy, sz, li, pix, piy, piz);
        double ax = 0.0, ay = 0.0, az = 0.0, e = 0.0;
        unsigned nacc = 0, cnt = 0;
        // entry k of slot i is half-word k&7 of quad k>>3; quads of one slot are 32 lanes x 8 half-words apart
        // the list this pass writes (pruning: the inner list; build: the outer list): quad q of slot i, see the layout above
        uint4* __restrict__ out_row = (MODE == TILE_FORCE_PRUNE || MODE == TILE_LIST_FILL)
            ? reinterpret_cast<uint4*>(MODE == TILE_FORCE_PRUNE ? a.in32 : a.nl32) + ((size_t)(i >> 5) * a.capq) * 32 + (i & 31u) : nullptr;
        QuadPacker pk; pk.q = make_uint4(0u, 0u, 0u, 0u); pk.cnt = 0;

#define LJ_TILE_PAIR(J)                                                                     \
        {                                                                                   \
            double dx, dy, dz;                                                              \
            tile_ld(sxy, sz, (J), dx, dy, dz);                                              \
            dx = pix - dx; dy = piy - dy; dz = piz - dz;                                    \
            double d2 = fma(dz, dz, fma(dy, dy, dx * dx));                                  \
            if (maybe_wrapped(d2, far2_hi)) {                                               \
                dx = min_image(dx, c.L, c.halfL); dy = min_image(dy, c.L, c.halfL);         \
                dz = min_image(dz, c.L, c.halfL);                                           \
                d2 = fma(dz, dz, fma(dy, dy, dx * dx));                                     \
            }                                                                               \
            if (d2 < c.rc2) {                                                               \
                const double fr = lj_pair(d2, c.sigma2, e);                                 \
                ax = fma(fr, dx, ax); ay = fma(fr, dy, ay); az = fma(fr, dz, az);           \
                ++nacc;                                                                     \
            }                                                                               \
        }

        if (MODE == TILE_FORCE_LIST || MODE == TILE_FORCE_PRUNE) {
            // Branch-free groups of kTileGroup entries: all the shared-memory loads and the
            // distance evaluations are issued back to back and the Lennard-Jones term is evaluated
            // for every entry, accumulated under a predicate.  (With 32 lanes and ~90 % of the entries
            // inside the cutoff a warp would execute the term for practically every entry anyway;
            // without the branches the four dependent chains overlap.)  The only branch left is the
            // rare minimum-image fix-up for pairs that straddle the box.  List quads (8 entries) are
            // fetched one iteration ahead with a single 128-bit load.
            // the length and the first quad are fetched together (the quad's address does not depend on the length;
            // capq >= 1, so it exists even for an empty list): one global round trip per work item, not two
            const uint4* __restrict__ row = reinterpret_cast<const uint4*>(use_inner ? a.in32 : a.nl32) + ((size_t)(i >> 5) * a.capq) * 32 + (i & 31u);
            uint4 cur = __ldg(row);
            const unsigned n = use_inner ? a.in_count[i] : a.nl_count[i];
            const unsigned nq = (n + 7) >> 3;
            const uint32_t selfw = li | (li << 16);
            if (nq == 0) cur = make_uint4(selfw, selfw, selfw, selfw);
            for (unsigned qq = 0; qq < nq; ++qq) {
                uint4 nxt = make_uint4(selfw, selfw, selfw, selfw);
                if (qq + 1 < nq) nxt = __ldg(row + (size_t)(qq + 1) * 32);     // (two quads ahead: measured, no gain - r03e)
                const uint32_t cw[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
                for (int part = 0; part < 8 / kTileGroup; ++part) {
                    unsigned jj[kTileGroup];
                    double dx[kTileGroup], dy[kTileGroup], dz[kTileGroup], d2[kTileGroup];
                    int far = 0;
#pragma unroll
                    for (int u = 0; u < kTileGroup; ++u) {
                        const uint32_t w = cw[(part * kTileGroup + u) >> 1];
                        // (extracting with one PRMT per index - __byte_perm - saves 16 integer instructions per quad and changed
                        // nothing for the plain pass; with it the 2-GPU slab worker produced garbage forces at N = 262 144 while this
                        // form passes, same sources otherwise: gpurun_out/r03u_bisect.log.  Not understood, so not used.)
                        jj[u] = (u & 1) ? (w >> 16) : (w & 0xffffu);
#ifdef LJ_EXP_NOCONFLICT /* timing experiment only (steps with dt == 0): every lane reads a different bank - a gather without bank conflicts */
                        if (c.dt == 0.0) jj[u] = (jj[u] & ~31u) | lane;
#endif
#ifdef LJ_EXP_NOLDS      /* timing experiment only (steps with dt == 0): no shared-memory gather */
                        if (c.dt == 0.0) {
                            const double fake = __hiloint2double(0x3ff00000 + (int)(jj[u] << 8), 0);
                            dx[u] = fake; dy[u] = fake * 0.5; dz[u] = fake * 0.25;
                        } else
#endif
                        tile_ld(sxy, sz, jj[u], dx[u], dy[u], dz[u]);
                    }
#ifdef LJ_EXP_NOFP       /* timing experiment only (steps with dt == 0): gather, almost no arithmetic */
                    if (c.dt == 0.0) {
#pragma unroll
                        for (int u = 0; u < kTileGroup; ++u) { ax += dx[u]; ay += dy[u]; az += dz[u]; }
                        continue;
                    }
#endif
#pragma unroll
                    for (int u = 0; u < kTileGroup; ++u) {
                        dx[u] = pix - dx[u]; dy[u] = piy - dy[u]; dz[u] = piz - dz[u];
                        d2[u] = fma(dz[u], dz[u], fma(dy[u], dy[u], dx[u] * dx[u]));
                        far = max(far, __double2hiint(d2[u]));
                    }
                    if (far >= far2_hi) {
#pragma unroll
                        for (int u = 0; u < kTileGroup; ++u)
                            if (maybe_wrapped(d2[u], far2_hi)) {
                                dx[u] = min_image(dx[u], c.L, c.halfL); dy[u] = min_image(dy[u], c.L, c.halfL);
                                dz[u] = min_image(dz[u], c.L, c.halfL);
                                d2[u] = fma(dz[u], dz[u], fma(dy[u], dy[u], dx[u] * dx[u]));
                            }
                    }
#pragma unroll
                    for (int u = 0; u < kTileGroup; ++u) {
                        // padding entries point at the atom itself (d2 == 0): excluded by index
                        const bool in = (d2[u] < c.rc2) && (jj[u] != li);
                        // one select, not a divergent branch: a rejected entry gets 1/d2 := 0, which turns both
                        // terms below into exact +0 (same bits as selecting them afterwards, half the selects).
                        // (Predicating the accumulations instead - inline PTX, "@p fma" - was tried: ptxas turns every
                        // predicated fp64 accumulation into the operation plus two FSEL, eight extra instructions per entry.)
                        const double inv = in ? fast_rcp(d2[u]) : 0.0;
                        const double s2 = UNIT_SIGMA ? inv : c.sigma2 * inv;
                        const double s6 = s2 * s2 * s2;
                        const double tt = fma(s6, s6, -s6);         // s12 - s6
                        const double fr = fma(s6, s6, tt) * inv;    // (2 s12 - s6) / d2
                        e += tt;
                        ax = fma(fr, dx[u], ax); ay = fma(fr, dy[u], ay); az = fma(fr, dz[u], az);
                        nacc += in ? 1u : 0u;
                        if (MODE == TILE_FORCE_PRUNE) {
                            // by-product: the entries still within rcut + s_in form the inner list of the coming steps
                            if (d2[u] <= c.prune_cut2 && jj[u] != li) pk.push(jj[u], out_row, a.capq);
                        }
                    }
                }
                cur = nxt;
            }
        } else {
            // home cell of this atom along x, then the 3x3 x-rows around it: 3 consecutive tile cells each
            unsigned h = 0;
            while (h + 1 < hd.bw && soff[rbase + 2 + h] <= li) ++h;
            for (unsigned rz = hz - 1; rz <= hz + 1; ++rz)
            for (unsigned ry = hy - 1; ry <= hy + 1; ++ry) {
                const unsigned r = (rz * H + ry) * W + h;
                const unsigned j0 = soff[r], j1 = soff[r + 3];
                if (MODE == TILE_FORCE_CELL) {
                    for (unsigned j = j0; j < j1; ++j)
                        if (j != li) LJ_TILE_PAIR(j)
                } else {
                    // list build / count.  Candidates are examined 32 at a time, kFillBatch per round with loads and distances
                    // batched (the scan is bound by the latency of its dependent chain load -> difference -> square -> sum
                    // -> compare per warp, like the list walk, but a candidate only keeps two registers alive, so many can be
                    // in flight); acceptance goes into a bit mask and only the accepted ones are pushed afterwards.
                    for (unsigned jc = j0; jc < j1; jc += 32) {
                        unsigned mask = 0;
                        const unsigned jend = min(jc + 32u, j1);
                        // (a round may read up to kFillBatch - 1 records past j1: they exist - every slot has 32 spare records,
                        // see tcap - and their bits are cleared below; the atom itself is dropped when the mask is drained)
                        for (unsigned jb = jc; jb < jend; jb += kFillBatch) {
                            double d2[kFillBatch];
                            int far = 0;
#pragma unroll
                            for (int u = 0; u < kFillBatch; ++u) {
                                double dx, dy, dz;
                                tile_ld(sxy, sz, jb + u, dx, dy, dz);
                                dx = pix - dx; dy = piy - dy; dz = piz - dz;
                                d2[u] = fma(dz, dz, fma(dy, dy, dx * dx));
                                far = max(far, __double2hiint(d2[u]));
                            }
                            if (far >= far2_hi) {
#pragma unroll
                                for (int u = 0; u < kFillBatch; ++u)
                                    if (maybe_wrapped(d2[u], far2_hi)) {
                                        double dx, dy, dz;
                                        tile_ld(sxy, sz, jb + u, dx, dy, dz);
                                        dx = min_image(pix - dx, c.L, c.halfL); dy = min_image(piy - dy, c.L, c.halfL);
                                        dz = min_image(piz - dz, c.L, c.halfL);
                                        d2[u] = fma(dz, dz, fma(dy, dy, dx * dx));
                                    }
                            }
                            unsigned mb = 0;
#pragma unroll
                            for (int u = 0; u < kFillBatch; ++u) mb |= (d2[u] <= c.list_cut2) ? (1u << u) : 0u;
                            mask |= mb << (jb - jc);
                        }
                        mask &= 0xffffffffu >> (32u - (jend - jc));          // candidates past the end of the row
                        if (li - jc < 32u) mask &= ~(1u << (li - jc));         // the atom itself
                        if (MODE == TILE_LIST_FILL) {
                            while (mask) {                       // ascending j: the list keeps the order of the scan
                                const unsigned bit = (unsigned)__ffs((int)mask) - 1u;
                                mask &= mask - 1u;
                                pk.push(jc + bit, out_row, a.capq);   // (entries beyond the capacity are counted, not stored)
                            }
                        } else {
                            cnt += (unsigned)__popc(mask);
                        }
                    }
                }
            }
        }
#undef LJ_TILE_PAIR

        if (MODE == TILE_FORCE_PRUNE) {
            pk.finish(li, out_row, a.capq);
            a.in_count[i] = pk.cnt;
        }
        if (MODE == TILE_FORCE_LIST || MODE == TILE_FORCE_CELL || MODE == TILE_FORCE_PRUNE) {
            a.fx[i] = ax * c.prefctr; a.fy[i] = ay * c.prefctr; a.fz[i] = az * c.prefctr;
            etot += fma(-(double)nacc, c.ecut, e);
        } else {
            if (MODE == TILE_LIST_FILL) { pk.finish(li, out_row, a.capq); cnt = pk.cnt; }
            a.nl_count[i] = MODE == TILE_LIST_FILL ? min(cnt, 8 * a.capq) : cnt;
            maxcnt = max(maxcnt, cnt);
        }
            }
        }

        if (MODE == TILE_FORCE_LIST || MODE == TILE_FORCE_CELL || MODE == TILE_FORCE_PRUNE) {
            // one partial per (brick, item): fixed shape => the final sum is reproducible whichever warp ran the item
            const double ew = warp_sum(etot);
            if (lane == 0) a.partial_epot[(size_t)b * a.ipb + chunk] = ew;
        }
        __syncwarp();                                    // every lane's reads of the slot are done
        if (chunk * 32u < hd.hn) LJ_TRACE(4, k, chunk);
        if (lane == 0) {
            if (atomicAdd(&s_done[slot], 1u) + 1u == a.ipb) {      // the brick's last item: hand the slot back
                s_done[slot] = 0;
                st_release_smem(&s_drained[slot], use + 1u);
            }
        }
    }

    if (MODE == TILE_LIST_COUNT || MODE == TILE_LIST_FILL) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxcnt = max(maxcnt, __shfl_xor_sync(0xffffffffu, maxcnt, o));
        if (lane == 0 && maxcnt) {
            atomicMax(&a.ctrl->max_count, maxcnt);
            if (MODE == TILE_LIST_FILL && maxcnt > 8 * a.capq) atomicExch(&a.ctrl->overflow, 1u);
        }
    }
}

} // namespace ljgpu
