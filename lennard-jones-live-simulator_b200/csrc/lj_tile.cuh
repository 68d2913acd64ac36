// lj_tile.cuh - shared-memory tile kernels: Verlet-list force, cell-list force and the list build.
//
// Why tiles: a per-thread gather of neighbour positions from L1/L2 costs one cache line per lane per
// load (measured: the first Verlet kernel ran at exactly one 128-byte line per clock per SM, 85-90 %
// of the L1 data pipe, FP64 pipe a third busy - profiles/r01b_*).  Shared memory serves 32 arbitrary
// addresses per clock.  So one CTA owns a small brick of cells ("home block", kTileBX x kTileBY x
// kTileBZ cells), stages the positions of the brick plus one halo cell on every side once with
// coalesced 256-bit loads, and every home atom walks its list as 16-bit indices into that tile.
//
// Reference code replaced: calculate_forces (.cpp:258-330) and build_neighbor_list (.cpp:441-551) of
// /root/reference/src/simulator/lennardjonessimulation.cpp.
#pragma once
#include "lj_kernels.cuh"

namespace ljgpu {

#ifndef LJ_TILE_BX
#define LJ_TILE_BX 3
#endif
#ifndef LJ_TILE_MAXTHREADS
#define LJ_TILE_MAXTHREADS 256
#endif
#ifndef LJ_TILE_MINBLOCKS
#define LJ_TILE_MINBLOCKS 3
#endif
constexpr int kTileBX = LJ_TILE_BX;                    // home cells per CTA along x (contiguous slots)
constexpr int kTileBY = 2;
constexpr int kTileBZ = 2;
constexpr int kTileRowsMax = (kTileBY + 2) * (kTileBZ + 2);           // x-rows in a tile
constexpr int kTileCellsMax = (kTileBX + 2) * kTileRowsMax;           // tile cells per CTA
constexpr int kHomeRowsMax = kTileBY * kTileBZ;

// Shared-memory record of a tile atom.  The list walk is a random gather, and the shared-memory data pipe is
// its busiest unit (ncu: 72 % of cycles, 6.5 wavefronts per 64-bit load where 2 would be conflict-free - 16
// random lanes over 16 bank pairs).  LJ_TILE_REC32: 32-byte records, (x,y) fetched with one 128-bit load
// (quarter-warps over 8 bank quads) and z with a 64-bit load; the two 16-byte halves of a record swap places
// with bit 4 of the atom index so that the (x,y) halves spread over all eight quads.  Otherwise 24-byte
// records and three 64-bit loads.  Measured at N = 2^20: 237 us per force pass with 32-byte records against
// 227 us with 24-byte ones (fewer loads, but not fewer wavefronts, and two more integer instructions per
// entry) - so 24 bytes stay the default.
#ifndef LJ_TILE_REC32
#define LJ_TILE_REC32 0
#endif
constexpr unsigned kTileRecBytes = LJ_TILE_REC32 ? 32u : 24u;

__device__ __forceinline__ void tile_ld(const double* sm, unsigned j, double& x, double& y, double& z) {
#if LJ_TILE_REC32
    const char* b = reinterpret_cast<const char*>(sm);
    const unsigned a = j * 32u, s = j & 16u;
    const double2 xy = *reinterpret_cast<const double2*>(b + (a + s));
    z = *reinterpret_cast<const double*>(b + (a + 16u - s));
    x = xy.x; y = xy.y;
#else
    const double* p = sm + 3 * (size_t)j;
#ifdef LJ_EMUL
    ljemul::lds_note(p);                   // conflict profile of the x loads (y and z behave alike: same stride)
#endif
    x = p[0]; y = p[1]; z = p[2];
#endif
}

struct TileGrid {
    unsigned nc;                 // cells per axis in x and y (periodic)
    unsigned ncz;                // cell layers of the local table (nc, or owned layers + 2 halo layers on a slab)
    unsigned zper;               // z periodic within this table (single device) or bounded by halo layers (slab)
    unsigned hz0, nhz;           // first home layer and number of home layers
    unsigned bxd, byd, bzd;      // brick size in cells (<= kTileBX, kTileBY, kTileBZ; smaller for small systems)
    unsigned nbx, nby, nbz;      // home blocks per axis
    unsigned nblocks;
    const uint32_t* __restrict__ tsrc;   // [nblocks][kTileCellsMax]     first global slot of each tile cell
    const uint32_t* __restrict__ toff;   // [nblocks][kTileCellsMax + 1] tile-local offset of each tile cell
};

struct TileGeom { unsigned x0, y0, z0, bw, bh, bd, W, H; };

// Block b -> its brick.  Tile cell numbering: t = ((tz*(bh+2)) + ty)*(bw+2) + tx with one halo cell on
// every side (tx = 0 and bw+1 etc.), so the cells of one x-row are consecutive in the tile.
__device__ __forceinline__ TileGeom tile_geometry(unsigned b, const TileGrid& g) {
    TileGeom t;
    const unsigned bx = b % g.nbx, by = (b / g.nbx) % g.nby;
    unsigned bz = b / (g.nbx * g.nby);
    if (!g.zper && g.nbz > 2)        // slab: the two bricks that touch halo layers are scheduled last (they wait for the neighbours)
        bz = bz < g.nbz - 2 ? bz + 1 : (bz == g.nbz - 2 ? 0u : g.nbz - 1);
    t.x0 = bx * g.bxd; t.y0 = by * g.byd; t.z0 = g.hz0 + bz * g.bzd;
    t.bw = min(g.bxd, g.nc - t.x0);
    t.bh = min(g.byd, g.nc - t.y0);
    t.bd = min(g.bzd, g.hz0 + g.nhz - t.z0);
    t.W = t.bw + 2; t.H = t.bh + 2;
    return t;
}

// Re-binning: per block, where each tile cell's atoms live (global slot) and where they go in the tile.
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

// Re-binning, optional (LJGPU_LIST_ORDER=1): copies every slot's list from src to dst with the entries sorted by
// bank class (j mod 16: the 24-byte records put atom j's x in bank pair 3j mod 16) rotated by the slot number, so
// that the 16 lanes of a half-warp tend to read 16 different bank pairs at the same step of their walks.  Stable,
// so the pruned list - a filtered subsequence - inherits the order.  CPU-shim gather profile (liquid, inner-list
// pass): 4.62 -> 4.04 wavefronts per gather on the per-atom path, 5.18 -> 4.23 on the pair path.  One thread per slot.
__global__ void __launch_bounds__(256)
list_class_sort_kernel(const uint32_t* __restrict__ src32, uint32_t* __restrict__ dst32,
                       const uint32_t* __restrict__ count, unsigned nslots, unsigned capq) {
    const unsigned slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= nslots) return;
    const unsigned n = min(count[slot], 8 * capq);
    const size_t rowq = ((size_t)(slot >> 5) * capq) * 32 + (slot & 31u);
    const uint4* __restrict__ src = reinterpret_cast<const uint4*>(src32) + rowq;
    unsigned short* __restrict__ dst = reinterpret_cast<unsigned short*>(dst32) + rowq * 8;
    const unsigned nq = (n + 7) >> 3, rot = slot & 15u;
    unsigned char start[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) start[k] = 0;
    for (unsigned q = 0; q < nq; ++q) {                       // histogram of the rotated classes
        const uint4 v = __ldg(src + (size_t)q * 32);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 8; ++e)
            if (8 * q + e < n) {
                const unsigned j = (e & 1) ? (w[e >> 1] >> 16) : (w[e >> 1] & 0xffffu);
                ++start[(j - rot) & 15u];
            }
    }
    unsigned run = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) { const unsigned c = start[k]; start[k] = (unsigned char)run; run += c; }
    for (unsigned q = 0; q < nq; ++q) {                       // stable scatter; the padding of the last quad is copied
        const uint4 v = __ldg(src + (size_t)q * 32);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const unsigned j = (e & 1) ? (w[e >> 1] >> 16) : (w[e >> 1] & 0xffffu);
            const unsigned k = 8 * q + e;
            const unsigned pos = k < n ? start[(j - rot) & 15u]++ : k;
            dst[(size_t)(pos >> 3) * 256 + (pos & 7u)] = (unsigned short)j;
        }
    }
}

enum { TILE_FORCE_LIST = 0, TILE_FORCE_CELL = 1, TILE_LIST_COUNT = 2, TILE_LIST_FILL = 3,
       TILE_FORCE_PRUNE = 5 /* force pass over the outer list that also emits the pruned inner list */ };

struct TileArgs {
    const double4* __restrict__ pos;
    uint32_t* __restrict__ nl32;           // list words (fill: written as uint16 halves)
    uint32_t* __restrict__ nl_count;
    uint32_t* __restrict__ in32;           // pruned inner list, same layout (null: none)
    uint32_t* __restrict__ in_count;
    double* __restrict__ fx; double* __restrict__ fy; double* __restrict__ fz;
    double* __restrict__ partial_epot;     // one per CTA
    StepCtrl* ctrl;
    unsigned capq;                         // list capacity in 128-bit quads (8 entries) per atom
    // slab path over peer memory: wait until both neighbours have pushed this step's halo (null: no wait)
    Mailbox* mail;
    unsigned long long halo_seq;
};

// UNIT_SIGMA: sigma == 1 exactly (reduced units): s2 = sigma^2/d2 needs no multiply.
// The pruning pass carries the list emission on top of the force walk and spills at the common register budget;
// LJ_TILE_PRUNE_MINBLOCKS lets it trade occupancy for registers on its own (tuning knob, default: same budget).
#ifndef LJ_TILE_PRUNE_MINBLOCKS
#define LJ_TILE_PRUNE_MINBLOCKS LJ_TILE_MINBLOCKS
#endif
template <int MODE, bool UNIT_SIGMA>
__global__ void __launch_bounds__(LJ_TILE_MAXTHREADS, MODE == 5 /* TILE_FORCE_PRUNE */ ? LJ_TILE_PRUNE_MINBLOCKS : LJ_TILE_MINBLOCKS)
tile_kernel(LJConst c, TileGrid g, TileArgs a, int guarded) {
    LJ_DYNAMIC_SMEM(sm);                                // tile positions, one record per atom (tile_ld)
    __shared__ uint32_t s_off[kTileCellsMax + 1];
    __shared__ uint32_t s_src[kTileCellsMax];
    __shared__ uint32_t s_hstart[kHomeRowsMax + 1];     // prefix of home atoms per home x-row
    __shared__ double s_epot[LJ_TILE_MAXTHREADS / 32];
    __shared__ unsigned s_ticket;
    if (guarded && a.ctrl->stop) return;
    const unsigned b = blockIdx.x;
    const TileGeom tg = tile_geometry(b, g);
    if (a.mail) {       // slab over peer memory: only bricks whose tile reaches a halo layer wait for that neighbour's push
        if (threadIdx.x == 0) {
            if (tg.z0 == g.hz0) wait_flag(&a.mail->halo_flag[0], a.halo_seq, &a.mail->error);
            if (tg.z0 + tg.bd == g.hz0 + g.nhz) wait_flag(&a.mail->halo_flag[1], a.halo_seq, &a.mail->error);
        }
        __syncthreads();
    }
    const unsigned W = tg.W, H = tg.H, ntc = W * H * (tg.bd + 2);
    for (unsigned t = threadIdx.x; t <= ntc; t += blockDim.x) {
        s_off[t] = g.toff[(size_t)b * (kTileCellsMax + 1) + t];
        if (t < ntc) s_src[t] = g.tsrc[(size_t)b * kTileCellsMax + t];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        s_ticket = 0;
        unsigned run = 0, q = 0;
        for (unsigned hz = 1; hz <= tg.bd; ++hz)
            for (unsigned hy = 1; hy <= tg.bh; ++hy) {
                const unsigned r = (hz * H + hy) * W;
                s_hstart[q++] = run;
                run += s_off[r + 1 + tg.bw] - s_off[r + 1];
            }
        for (; q <= kHomeRowsMax; ++q) s_hstart[q] = run;
    }

    // stage the neighbourhood: one warp per tile x-row (W consecutive tile cells, ~100 atoms at liquid
    // density), lanes stride over the row's atoms.  An atom's global slot is its tile index plus the
    // (source - offset) difference of its cell, picked by comparing against the row's W-1 inner cell
    // offsets held in registers (a per-cell loop cost ~4x the instructions: 21 of 32 lanes busy and the
    // loop set-up once per cell).  Asynchronous global->shared copies (LDGSTS), so no thread waits on a
    // load before issuing the next; x,y,z of an atom are 24 contiguous bytes of the 32-byte global
    // record and go to the 24-byte shared record as three 8-byte copies.
    {
        constexpr int kW = kTileBX + 2;
        const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
        const unsigned nrows = H * (tg.bd + 2);
        const unsigned smem_base = smem_addr(sm);
        for (unsigned r = w; r < nrows; r += nw) {
            const unsigned t0 = r * W;
            unsigned offk[kW], delk[kW];
#pragma unroll
            for (int k = 0; k < kW; ++k) {
                const bool on = (unsigned)k < W;
                const unsigned o = on ? s_off[t0 + k] : 0xffffffffu;
                offk[k] = o;
                delk[k] = on ? s_src[t0 + k] - o : 0u;      // modulo 2^32: li + delk = source slot
            }
            const unsigned o1 = s_off[t0 + W];
            for (unsigned li = offk[0] + lane; li < o1; li += 32) {
                unsigned d = delk[0];
#pragma unroll
                for (int k = 1; k < kW; ++k) d = (li >= offk[k]) ? delk[k] : d;
                const double* gp = reinterpret_cast<const double*>(a.pos + (li + d));
#if LJ_TILE_REC32
                const unsigned sp = smem_base + li * 32u, sw = li & 16u;
                cp_async16(sp + sw, gp);
                cp_async16(sp + 16u - sw, gp + 2);
#else
                const unsigned sp = smem_base + li * 24u;
                cp_async8(sp, gp);
                cp_async8(sp + 8, gp + 1);
                cp_async8(sp + 16, gp + 2);
#endif
            }
        }
        cp_async_wait_all();
    }
    __syncthreads();

    const unsigned hn = s_hstart[kHomeRowsMax];
    const int far2_hi = __double2hiint(c.far2);
    // Pruned inner list (entries within rcut + s_in at the last pruning): usable while no pair can have come
    // closer than s_in since, i.e. while the summed per-step maximum displacement stays below s_in / 2.
    // Same decision in every block: the inputs were final before this launch.
    bool use_inner = false;
    if (MODE == TILE_FORCE_LIST && a.in32 != nullptr) {   // (TILE_FORCE_PRUNE always walks the outer list)
        double v2 = __longlong_as_double((long long)a.ctrl->vmax2);
        if (a.mail) {   // slab: a brick that reaches a halo layer also sees that neighbour's atoms (value pushed with the halo)
            if (tg.z0 == g.hz0) v2 = fmax(v2, ld_sys_f64(&a.mail->nb_vmax2[0]));
            if (tg.z0 + tg.bd == g.hz0 + g.nhz) v2 = fmax(v2, ld_sys_f64(&a.mail->nb_vmax2[1]));
        }
        const double bound = a.ctrl->drift_bound + sqrt(v2) * c.dt;
        use_inner = a.ctrl->inner_valid != 0u && bound <= c.prune_half;
        if (use_inner && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&a.ctrl->inner_used, 1u);
    }
    double etot = 0.0;
    unsigned maxcnt = 0;

    for (unsigned t = threadIdx.x; t < hn; t += blockDim.x) {
        // home row of this atom -> global slot i and tile-local index li
        unsigned q = 0;
#pragma unroll
        for (int k = 1; k < kHomeRowsMax; ++k) q += (s_hstart[k] <= t) ? 1u : 0u;
        const unsigned hy = 1 + q % tg.bh, hz = 1 + q / tg.bh;
        const unsigned rbase = (hz * H + hy) * W;                 // tile cell index of (tx=0, hy, hz)
        const unsigned li = s_off[rbase + 1] + (t - s_hstart[q]);
        const unsigned i = s_src[rbase + 1] + (t - s_hstart[q]);
        double pix, piy, piz;
        tile_ld(sm, li, pix, piy, piz);
        double ax = 0.0, ay = 0.0, az = 0.0, e = 0.0;
        unsigned nacc = 0, cnt = 0;
        // entry k of slot i is half-word k&7 of quad k>>3; quads of one slot are 32 lanes x 8 half-words apart
        unsigned short* __restrict__ in16 = (MODE == TILE_FORCE_PRUNE)
            ? reinterpret_cast<unsigned short*>(a.in32) + (((size_t)(i >> 5) * a.capq) * 32 + (i & 31u)) * 8 : nullptr;
        unsigned short* __restrict__ fill16 = (MODE == TILE_LIST_FILL)
            ? reinterpret_cast<unsigned short*>(a.nl32) + (((size_t)(i >> 5) * a.capq) * 32 + (i & 31u)) * 8 : nullptr;

#define LJ_TILE_PAIR(J)                                                                     \
        {                                                                                   \
            double dx, dy, dz;                                                              \
            tile_ld(sm, (J), dx, dy, dz);                                                   \
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
            // Branch-free groups of four entries: all twelve shared-memory loads and the four
            // distance evaluations are issued back to back and the Lennard-Jones term is evaluated
            // for every entry, accumulated under a predicate.  (With 32 lanes and ~60 % of the entries
            // inside the cutoff a warp would execute the term for practically every entry anyway;
            // without the branches the four dependent chains overlap.)  The only branch left is the
            // rare minimum-image fix-up for pairs that straddle the box.  List quads (8 entries) are
            // fetched one iteration ahead with a single 128-bit load.
            const unsigned n = use_inner ? a.in_count[i] : a.nl_count[i];
            const uint4* __restrict__ row = reinterpret_cast<const uint4*>(use_inner ? a.in32 : a.nl32) + ((size_t)(i >> 5) * a.capq) * 32 + (i & 31u);
            const unsigned nq = (n + 7) >> 3;
            const uint32_t selfw = li | (li << 16);
            uint4 cur = make_uint4(selfw, selfw, selfw, selfw);
            if (nq > 0) cur = __ldg(row);
            for (unsigned q = 0; q < nq; ++q) {
                uint4 nxt = make_uint4(selfw, selfw, selfw, selfw);
#ifdef LJ_EXP_NOLIST     /* timing experiment only: no list traffic */
                nxt = make_uint4(cur.x + 0x00010001u, cur.y + 0x00010001u, cur.z + 0x00010001u, cur.w + 0x00010001u);
                nxt.x &= 0x03ff03ffu; nxt.y &= 0x03ff03ffu; nxt.z &= 0x03ff03ffu; nxt.w &= 0x03ff03ffu;
#else
                if (q + 1 < nq) nxt = __ldg(row + (size_t)(q + 1) * 32);
#endif
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const uint32_t w0 = half ? cur.z : cur.x, w1 = half ? cur.w : cur.y;
                    unsigned jj[4] = { w0 & 0xffffu, w0 >> 16, w1 & 0xffffu, w1 >> 16 };
                    double dx[4], dy[4], dz[4], d2[4];
                    int far = 0;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
#ifdef LJ_EXP_NOLDS      /* timing experiment only: no shared-memory gather */
                        const double fake = __hiloint2double(0x3ff00000 + (int)(jj[u] << 8), 0);
                        dx[u] = fake; dy[u] = fake * 0.5; dz[u] = fake * 0.25;
#else
                        tile_ld(sm, jj[u], dx[u], dy[u], dz[u]);
                        dx[u] = pix - dx[u]; dy[u] = piy - dy[u]; dz[u] = piz - dz[u];
#endif
                        d2[u] = fma(dz[u], dz[u], fma(dy[u], dy[u], dx[u] * dx[u]));
                        far = max(far, __double2hiint(d2[u]));
                    }
                    if (far >= far2_hi) {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (maybe_wrapped(d2[u], far2_hi)) {
                                dx[u] = min_image(dx[u], c.L, c.halfL); dy[u] = min_image(dy[u], c.L, c.halfL);
                                dz[u] = min_image(dz[u], c.L, c.halfL);
                                d2[u] = fma(dz[u], dz[u], fma(dy[u], dy[u], dx[u] * dx[u]));
                            }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        // padding entries point at the atom itself (d2 == 0): excluded by index
                        const bool in = (d2[u] < c.rc2) && (jj[u] != li);
                        // one select, not a divergent branch: a rejected entry gets 1/d2 := 0, which turns both
                        // terms below into exact +0 (same bits as selecting them afterwards, half the selects)
                        const double inv = in ? fast_rcp(d2[u]) : 0.0;
                        const double s2 = UNIT_SIGMA ? inv : c.sigma2 * inv;
                        const double s6 = s2 * s2 * s2;
                        const double t = fma(s6, s6, -s6);          // s12 - s6
                        const double fr = fma(s6, s6, t) * inv;     // (2 s12 - s6) / d2
                        e += t;
                        ax = fma(fr, dx[u], ax); ay = fma(fr, dy[u], ay); az = fma(fr, dz[u], az);
                        nacc += in ? 1u : 0u;
                        if (MODE == TILE_FORCE_PRUNE) {
                            // by-product: the entries still within rcut + s_in form the inner list of the coming steps
                            if (d2[u] <= c.prune_cut2 && jj[u] != li) {
                                in16[(size_t)(cnt >> 3) * 256 + (cnt & 7u)] = (unsigned short)jj[u];   // quad stride: 32 lanes x 8 entries
                                ++cnt;
                            }
                        }
                    }
                }
                cur = nxt;
            }
        } else {
            // home cell of this atom along x, then the 3x3 x-rows around it: 3 consecutive tile cells each
            unsigned h = 0;
            while (h + 1 < tg.bw && s_off[rbase + 2 + h] <= li) ++h;
            for (unsigned rz = hz - 1; rz <= hz + 1; ++rz)
            for (unsigned ry = hy - 1; ry <= hy + 1; ++ry) {
                const unsigned r = (rz * H + ry) * W + h;
                const unsigned j0 = s_off[r], j1 = s_off[r + 3];
                if (MODE == TILE_FORCE_CELL) {
                    for (unsigned j = j0; j < j1; ++j)
                        if (j != li) LJ_TILE_PAIR(j)
                } else {
                    // list build / count: four candidates per round, loads and distances batched
                    for (unsigned jb = j0; jb < j1; jb += 4) {
                        double d2[4];
                        int far = 0;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const unsigned j = min(jb + u, j1 - 1);
                            double dx, dy, dz;
                            tile_ld(sm, j, dx, dy, dz);
                            dx = pix - dx; dy = piy - dy; dz = piz - dz;
                            d2[u] = fma(dz, dz, fma(dy, dy, dx * dx));
                            far = max(far, __double2hiint(d2[u]));
                        }
                        if (far >= far2_hi) {
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (maybe_wrapped(d2[u], far2_hi)) {
                                    const unsigned j = min(jb + u, j1 - 1);
                                    double dx, dy, dz;
                                    tile_ld(sm, j, dx, dy, dz);
                                    dx = min_image(pix - dx, c.L, c.halfL); dy = min_image(piy - dy, c.L, c.halfL);
                                    dz = min_image(piz - dz, c.L, c.halfL);
                                    d2[u] = fma(dz, dz, fma(dy, dy, dx * dx));
                                }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const unsigned j = jb + u;
                            if (j < j1 && d2[u] <= c.list_cut2 && j != li) {
                                if (MODE == TILE_LIST_FILL && cnt < 8 * a.capq)     // plain 16-bit stores: measured cheaper than packing quads in registers
                                    fill16[(size_t)(cnt >> 3) * 256 + (cnt & 7u)] = (unsigned short)j;
                                ++cnt;
                            }
                        }
                    }
                }
            }
        }
#undef LJ_TILE_PAIR

        if (MODE == TILE_FORCE_PRUNE) {
            for (unsigned k = cnt; k & 7u; ++k)               // the unused tail of the last quad points at the atom itself
                in16[(size_t)(k >> 3) * 256 + (k & 7u)] = (unsigned short)li;
            a.in_count[i] = cnt;
        }
        if (MODE == TILE_FORCE_LIST || MODE == TILE_FORCE_CELL || MODE == TILE_FORCE_PRUNE) {
            a.fx[i] = ax * c.prefctr; a.fy[i] = ay * c.prefctr; a.fz[i] = az * c.prefctr;
            etot += fma(-(double)nacc, c.ecut, e);
        } else {
            if (MODE == TILE_LIST_FILL)                      // the unused tail of the last quad points at the atom itself
                for (unsigned k = min(cnt, 8 * a.capq); k & 7u; ++k) fill16[(size_t)(k >> 3) * 256 + (k & 7u)] = (unsigned short)li;
            a.nl_count[i] = MODE == TILE_LIST_FILL ? min(cnt, 8 * a.capq) : cnt;
            maxcnt = max(maxcnt, cnt);
        }
    }

    if (MODE == TILE_FORCE_LIST || MODE == TILE_FORCE_CELL || MODE == TILE_FORCE_PRUNE) {
        // one partial per CTA without a block barrier (a finished warp frees its issue slots at once): every
        // warp parks its sum in shared memory, the last one to arrive adds them in warp order (fixed order
        // => reproducible) and writes the CTA's partial
        const double e = warp_sum(etot);
        if ((threadIdx.x & 31) == 0) {
            s_epot[threadIdx.x >> 5] = e;
            __threadfence_block();
            const unsigned nw = blockDim.x >> 5;
            if (atomicAdd(&s_ticket, 1u) == nw - 1) {
                __threadfence_block();
                double tot = 0.0;
                for (unsigned w = 0; w < nw; ++w) tot += ((volatile double*)s_epot)[w];
                a.partial_epot[blockIdx.x] = tot;
            }
        }
    } else {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxcnt = max(maxcnt, __shfl_xor_sync(0xffffffffu, maxcnt, o));
        if ((threadIdx.x & 31) == 0 && maxcnt) {
            atomicMax(&a.ctrl->max_count, maxcnt);
            if (MODE == TILE_LIST_FILL && maxcnt > 8 * a.capq) atomicExch(&a.ctrl->overflow, 1u);
        }
    }
}

} // namespace ljgpu
