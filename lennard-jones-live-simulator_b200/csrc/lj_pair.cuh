// lj_pair.cuh - EXPERIMENTAL, off by default (LJGPU_PAIRLIST=1; single device, Verlet path).
//
// Why: in the tile kernels every lane walks its own atom's neighbour list, and the gather of neighbour
// positions from shared memory is the busiest unit of the force pass (ncu: shared-memory data pipe 72 % of the
// cycles, 6.5 wavefronts per 64-bit load, FP64 pipe 41 % - profiles/r01k_ncu_tile_force_regions.txt).  Here one
// thread owns TWO consecutive home atoms of an x-row (same or adjacent cell, ~1-2 sigma apart) and walks the
// UNION of their lists: every gathered position serves two atoms.  Expected at rho* = 0.8: union ~1.3-1.5 x one
// list, i.e. -25..35 % gathers and list bytes for +30..50 % FP64 work (entries in range of only one of the two
// atoms are evaluated for both and accumulated as exact zeros), which balances the two pipes.
// Bricks are 6x2x2 cells (~500 atoms, ~250 pairs per CTA, tile 8x4x4 cells ~ 64 KB).
//
// Same arithmetic per pair as lj_tile.cuh (fast_rcp, selects, one-shot minimum image), same list format
// (16-bit tile indices, eight per 128-bit quad, 32 slots interleaved), slots are (block, thread) pairs.
// Status: compiles; NOT yet run on a GPU (written after the round's GPU budget was spent).  Reference code
// replaced: calculate_forces (.cpp:258-330), build_neighbor_list (.cpp:441-551).
#pragma once
#include "lj_tile.cuh"

namespace ljgpu {

#ifndef LJ_PAIR_BX
#define LJ_PAIR_BX 6
#endif
#ifndef LJ_PAIR_MAXTHREADS
#define LJ_PAIR_MAXTHREADS 256
#endif
#ifndef LJ_PAIR_MINBLOCKS
#define LJ_PAIR_MINBLOCKS 2
#endif
constexpr int kPairBX = LJ_PAIR_BX;
constexpr int kPairBY = 2;
constexpr int kPairBZ = 2;
constexpr int kPairRowsMax = (kPairBY + 2) * (kPairBZ + 2);
constexpr int kPairCellsMax = (kPairBX + 2) * kPairRowsMax;
constexpr int kPairHomeRows = kPairBY * kPairBZ;
#ifndef LJ_PAIR_GROUP
#define LJ_PAIR_GROUP 2                    // list entries evaluated together (x 2 atoms = independent chains): 1, 2 or 4
#endif
constexpr int kPairGroup = LJ_PAIR_GROUP;

// Re-binning: tile tables of the pair bricks (strides kPairCellsMax); max_tile = atoms in the largest tile,
// max_home = PAIRS in the fullest brick (sum over home x-rows of ceil(atoms / 2)).
__global__ void __launch_bounds__(128)
pair_table_kernel(TileGrid g, const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ ccount,
                  uint32_t* __restrict__ tsrc, uint32_t* __restrict__ toff, StepCtrl* __restrict__ ctrl) {
    const int lane = threadIdx.x & 31;
    const unsigned b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= g.nblocks) return;
    const TileGeom t = tile_geometry(b, g);
    const unsigned nc = g.nc, ntc = t.W * t.H * (t.bd + 2);
    uint32_t* my_off = toff + (size_t)b * (kPairCellsMax + 1);
    unsigned run = 0;
    for (unsigned t0 = 0; t0 < ntc; t0 += 32) {
        const unsigned k = t0 + lane;
        unsigned cnt = 0, src = 0;
        if (k < ntc) {
            const unsigned tx = k % t.W, ty = (k / t.W) % t.H, tz = k / (t.W * t.H);
            const unsigned cx = (t.x0 + nc + tx - 1) % nc, cy = (t.y0 + nc + ty - 1) % nc;
            const unsigned cz = g.zper ? (t.z0 + g.ncz + tz - 1) % g.ncz : (t.z0 + tz - 1);
            const unsigned cid = cz * nc * nc + cy * nc + cx;
            cnt = ccount[cid]; src = cstart[cid];
        }
        unsigned inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        if (k < ntc) {
            tsrc[(size_t)b * kPairCellsMax + k] = src;
            my_off[k] = run + inc - cnt;
        }
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) my_off[ntc] = run;
    __syncwarp();                                   // the offsets written above are read back by other lanes
    unsigned pairs = 0;
    if ((unsigned)lane < t.bd * t.bh) {
        const unsigned hy = 1 + (unsigned)lane % t.bh, hz = 1 + (unsigned)lane / t.bh;
        const unsigned r = (hz * t.H + hy) * t.W;
        pairs = (my_off[r + 1 + t.bw] - my_off[r + 1] + 1u) >> 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pairs += __shfl_xor_sync(0xffffffffu, pairs, o);
    if (lane == 0) {
        atomicMax(&ctrl->max_tile, run);
        atomicMax(&ctrl->max_home, pairs);
    }
}

// Re-binning, before the state is permuted: one warp per cell re-orders the cell's atoms (the entries of `vals` in
// the cell's segment of the sorted order) into a nearest-neighbour chain - first atom, then the unplaced atom nearest
// to it, then the one nearest to that, ... - so that the two consecutive atoms a thread owns are about one sigma
// apart and their lists overlap as much as possible (oracle liquid state: union 1.36x instead of 1.48x one list).
// Deterministic (ties broken by lane); cells with more than 64 atoms keep their order.
__global__ void __launch_bounds__(128)
pair_order_kernel(unsigned ncell, const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ cend,
                  const double4* __restrict__ pos_tmp, uint32_t* __restrict__ vals) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned cell = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (cell >= ncell) return;
    const unsigned c0 = cstart[cell], cnt = cend[cell] - c0;
    if (cnt < 3 || cnt > 64) return;
    unsigned idx[2] = {0u, 0u};
    double px[2], py[2], pz[2];
    bool free_[2];
    unsigned rank[2] = {0u, 0u};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const unsigned k = lane + 32u * h;
        free_[h] = k < cnt;
        px[h] = py[h] = pz[h] = 0.0;
        if (free_[h]) {
            idx[h] = vals[c0 + k];
            const double4 p = ld_pos(pos_tmp + idx[h]);
            px[h] = p.x; py[h] = p.y; pz[h] = p.z;
        }
    }
    // the chain starts at the cell's first atom
    double cx = __shfl_sync(0xffffffffu, px[0], 0), cy = __shfl_sync(0xffffffffu, py[0], 0), cz = __shfl_sync(0xffffffffu, pz[0], 0);
    if (lane == 0) { free_[0] = false; rank[0] = 0; }
    for (unsigned placed = 1; placed < cnt; ++placed) {
        double best = 1.0e300;
        unsigned who = 0xffffffffu;                                   // lane + 32 * h of my best candidate
#pragma unroll
        for (int h = 0; h < 2; ++h)
            if (free_[h]) {
                const double dx = px[h] - cx, dy = py[h] - cy, dz = pz[h] - cz;
                const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
                if (d2 < best) { best = d2; who = lane + 32u * (unsigned)h; }
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const unsigned ow = __shfl_xor_sync(0xffffffffu, who, o);
            if (ob < best || (ob == best && ow < who)) { best = ob; who = ow; }
        }
        const unsigned wl = who & 31u, wh = who >> 5;
        const double sx = wh ? px[1] : px[0], sy = wh ? py[1] : py[0], sz = wh ? pz[1] : pz[0];
        cx = __shfl_sync(0xffffffffu, sx, (int)wl); cy = __shfl_sync(0xffffffffu, sy, (int)wl); cz = __shfl_sync(0xffffffffu, sz, (int)wl);
        if (lane == wl) { free_[wh] = false; rank[wh] = placed; }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h)
        if (lane + 32u * h < cnt) vals[c0 + rank[h]] = idx[h];
}

// Re-binning, after the tile tables: per brick, rank the pairs by the squared distance between their two atoms
// (a lone atom counts as distance zero: the shortest list) with a bitonic sort in shared memory; ties by pair
// number.  rank2pair / pair2rank are per-brick tables of pslots entries.
constexpr int kPairMapMax = 512;
__global__ void __launch_bounds__(256)
pair_map_kernel(TileGrid g, const double4* __restrict__ pos, unsigned pslots,
                unsigned short* __restrict__ rank2pair, unsigned short* __restrict__ pair2rank) {
    __shared__ unsigned long long key[kPairMapMax];          // (float bits of d2) << 32 | pair number
    __shared__ uint32_t s_row_src[kPairHomeRows], s_row_cnt[kPairHomeRows], s_pstart[kPairHomeRows + 1];
    const unsigned b = blockIdx.x;
    const TileGeom tg = tile_geometry(b, g);
    const uint32_t* off = g.toff + (size_t)b * (kPairCellsMax + 1);
    const uint32_t* src = g.tsrc + (size_t)b * kPairCellsMax;
    if (threadIdx.x == 0) {
        unsigned run = 0, q = 0;
        for (unsigned hz = 1; hz <= tg.bd; ++hz)
            for (unsigned hy = 1; hy <= tg.bh; ++hy) {
                const unsigned r = (hz * tg.H + hy) * tg.W;
                s_row_src[q] = src[r + 1];
                s_row_cnt[q] = off[r + 1 + tg.bw] - off[r + 1];
                s_pstart[q++] = run;
                run += (s_row_cnt[q - 1] + 1u) >> 1;
            }
        for (; q <= kPairHomeRows; ++q) s_pstart[q] = run;
    }
    __syncthreads();
    const unsigned np = s_pstart[kPairHomeRows];
    for (unsigned t = threadIdx.x; t < kPairMapMax; t += blockDim.x) {
        unsigned long long k = ~0ull;                        // padding sorts to the end
        if (t < np) {
            unsigned q = 0;
#pragma unroll
            for (int j = 1; j < kPairHomeRows; ++j) q += (s_pstart[j] <= t) ? 1u : 0u;
            const unsigned pl = t - s_pstart[q];
            float d2 = 0.0f;
            if (2u * pl + 1u < s_row_cnt[q]) {
                const double4 p1 = ld_pos(pos + s_row_src[q] + 2u * pl), p2 = ld_pos(pos + s_row_src[q] + 2u * pl + 1u);
                const double dx = p1.x - p2.x, dy = p1.y - p2.y, dz = p1.z - p2.z;
                d2 = (float)(dx * dx + dy * dy + dz * dz);
            }
            k = ((unsigned long long)(unsigned)__float_as_int(d2) << 32) | t;     // d2 >= 0: float bits order like the value
        }
        key[t] = k;
    }
    __syncthreads();
    for (unsigned size = 2; size <= (unsigned)kPairMapMax; size <<= 1)
        for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
            for (unsigned t = threadIdx.x; t < (unsigned)kPairMapMax; t += blockDim.x) {
                const unsigned partner = t ^ stride;
                if (partner > t) {
                    const bool up = (t & size) == 0;
                    const unsigned long long x = key[t], y = key[partner];
                    if ((x > y) == up) { key[t] = y; key[partner] = x; }
                }
            }
            __syncthreads();
        }
    for (unsigned t = threadIdx.x; t < np; t += blockDim.x) {
        const unsigned pair = (unsigned)(key[t] & 0xffffffffull);
        rank2pair[(size_t)b * pslots + t] = (unsigned short)pair;
        pair2rank[(size_t)b * pslots + pair] = (unsigned short)t;
    }
}

struct PairArgs {
    const double4* __restrict__ pos;
    uint32_t* __restrict__ nl32;           // union lists, one row set per (block, thread) slot
    uint32_t* __restrict__ nl_count;       // per slot
    uint32_t* __restrict__ in32;           // pruned inner union lists (null: none)
    uint32_t* __restrict__ in_count;
    uint32_t* __restrict__ atom_count;     // per atom: its own list length (statistics; written by the fill)
    double* __restrict__ fx; double* __restrict__ fy; double* __restrict__ fz;
    double* __restrict__ partial_epot;     // one per CTA
    StepCtrl* ctrl;
    unsigned capq;                         // list capacity in quads per slot
    unsigned pslots;                       // slots per block (>= pairs of the fullest brick, multiple of 32)
    // lane balance (pair_map_kernel): the force passes give thread t the pair rank2pair[t] - pairs sorted by the
    // distance between their two atoms, i.e. by union length, so that the lanes of a warp finish together - and
    // keep list slot t (coalesced rows); the list build runs in natural order (lanes of a warp share their
    // candidates: broadcast loads) and writes pair p's list to slot pair2rank[p].  Null: identity.
    const unsigned short* __restrict__ rank2pair;
    const unsigned short* __restrict__ pair2rank;
};

template <int MODE, bool UNIT_SIGMA>
__global__ void __launch_bounds__(LJ_PAIR_MAXTHREADS, LJ_PAIR_MINBLOCKS)
pair_kernel(LJConst c, TileGrid g, PairArgs a, int guarded) {
    LJ_DYNAMIC_SMEM(sm);
    __shared__ uint32_t s_off[kPairCellsMax + 1];
    __shared__ uint32_t s_src[kPairCellsMax];
    __shared__ uint32_t s_pstart[kPairHomeRows + 1];    // prefix of pairs per home x-row
    __shared__ double s_epot[LJ_PAIR_MAXTHREADS / 32];
    __shared__ unsigned s_ticket;
    if (guarded && a.ctrl->stop) return;
    const unsigned b = blockIdx.x;
    const TileGeom tg = tile_geometry(b, g);
    const unsigned W = tg.W, H = tg.H, ntc = W * H * (tg.bd + 2);
    for (unsigned t = threadIdx.x; t <= ntc; t += blockDim.x) {
        s_off[t] = g.toff[(size_t)b * (kPairCellsMax + 1) + t];
        if (t < ntc) s_src[t] = g.tsrc[(size_t)b * kPairCellsMax + t];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        s_ticket = 0;
        unsigned run = 0, q = 0;
        for (unsigned hz = 1; hz <= tg.bd; ++hz)
            for (unsigned hy = 1; hy <= tg.bh; ++hy) {
                const unsigned r = (hz * H + hy) * W;
                s_pstart[q++] = run;
                run += (s_off[r + 1 + tg.bw] - s_off[r + 1] + 1u) >> 1;
            }
        for (; q <= kPairHomeRows; ++q) s_pstart[q] = run;
    }
    {   // stage the neighbourhood, one warp per tile x-row (see lj_tile.cuh)
        constexpr int kW = kPairBX + 2;
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
                delk[k] = on ? s_src[t0 + k] - o : 0u;
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

    const unsigned np = s_pstart[kPairHomeRows];
    const int far2_hi = __double2hiint(c.far2);
    bool use_inner = false;
    if (MODE == TILE_FORCE_LIST && a.in32 != nullptr) {
        const double bound = a.ctrl->drift_bound + sqrt(__longlong_as_double((long long)a.ctrl->vmax2)) * c.dt;
        use_inner = a.ctrl->inner_valid != 0u && bound <= c.prune_half;
        if (use_inner && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&a.ctrl->inner_used, 1u);
    }
    double etot = 0.0;
    unsigned maxcnt = 0;

    constexpr bool kForcePass = MODE == TILE_FORCE_LIST || MODE == TILE_FORCE_PRUNE;
    for (unsigned t = threadIdx.x; t < np; t += blockDim.x) {
        // pt: the pair (natural numbering: home rows, two consecutive atoms each); slot: where its list lives
        const unsigned pt = (kForcePass && a.rank2pair) ? a.rank2pair[(size_t)b * a.pslots + t] : t;
        const unsigned slot = b * a.pslots + ((!kForcePass && a.pair2rank) ? a.pair2rank[(size_t)b * a.pslots + t] : t);
        unsigned q = 0;
#pragma unroll
        for (int k = 1; k < kPairHomeRows; ++k) q += (s_pstart[k] <= pt) ? 1u : 0u;
        const unsigned hy = 1 + q % tg.bh, hz = 1 + q / tg.bh;
        const unsigned rbase = (hz * H + hy) * W;
        const unsigned pl = pt - s_pstart[q];
        const unsigned rs = s_off[rbase + 1], rcnt = s_off[rbase + 1 + tg.bw] - rs;
        const unsigned li1 = rs + 2u * pl;
        const bool has2 = 2u * pl + 1u < rcnt;
        const unsigned li2 = has2 ? li1 + 1u : li1;
        const unsigned i1 = s_src[rbase + 1] + 2u * pl, i2 = i1 + 1u;
        const size_t rowq = ((size_t)(slot >> 5) * a.capq) * 32 + (slot & 31u);      // first quad of the slot
        double p1x, p1y, p1z, p2x, p2y, p2z;
        tile_ld(sm, li1, p1x, p1y, p1z);
        tile_ld(sm, li2, p2x, p2y, p2z);
        double a1x = 0.0, a1y = 0.0, a1z = 0.0, a2x = 0.0, a2y = 0.0, a2z = 0.0, e = 0.0;
        unsigned nacc = 0, cnt = 0, c1 = 0, c2 = 0;
        unsigned short* __restrict__ out16 = nullptr;
        if (MODE == TILE_FORCE_PRUNE) out16 = reinterpret_cast<unsigned short*>(a.in32) + rowq * 8;
        if (MODE == TILE_LIST_FILL) out16 = reinterpret_cast<unsigned short*>(a.nl32) + rowq * 8;

        if (MODE == TILE_FORCE_LIST || MODE == TILE_FORCE_PRUNE) {
            const unsigned n = use_inner ? a.in_count[slot] : a.nl_count[slot];
            const uint4* __restrict__ row = reinterpret_cast<const uint4*>(use_inner ? a.in32 : a.nl32) + rowq;
            const unsigned nq = (n + 7) >> 3;
            const uint32_t padw = li1 | (li1 << 16);
            uint4 cur = make_uint4(padw, padw, padw, padw);
            if (nq > 0) cur = __ldg(row);
            for (unsigned qd = 0; qd < nq; ++qd) {
                uint4 nxt = make_uint4(padw, padw, padw, padw);
                if (qd + 1 < nq) nxt = __ldg(row + (size_t)(qd + 1) * 32);
                const unsigned nvalid = n - 8u * qd;                 // entries of this quad that exist (tail: padding)
#pragma unroll
                for (int part = 0; part < 8 / kPairGroup; ++part) {  // kPairGroup entries x two atoms: chains in flight
                    constexpr int G = kPairGroup;
                    unsigned jj[G];
#pragma unroll
                    for (int u = 0; u < G; ++u) {
                        const int k = part * G + u;                  // half-word k of the quad
                        const uint32_t w = (k >> 1) == 0 ? cur.x : (k >> 1) == 1 ? cur.y : (k >> 1) == 2 ? cur.z : cur.w;
                        jj[u] = (k & 1) ? (w >> 16) : (w & 0xffffu);
                    }
                    double d1x[G], d1y[G], d1z[G], d2x[G], d2y[G], d2z[G], r1[G], r2[G];
                    int far = 0;
#pragma unroll
                    for (int u = 0; u < G; ++u) {
                        double jx, jy, jz;
                        tile_ld(sm, jj[u], jx, jy, jz);
                        d1x[u] = p1x - jx; d1y[u] = p1y - jy; d1z[u] = p1z - jz;
                        d2x[u] = p2x - jx; d2y[u] = p2y - jy; d2z[u] = p2z - jz;
                        r1[u] = fma(d1z[u], d1z[u], fma(d1y[u], d1y[u], d1x[u] * d1x[u]));
                        r2[u] = fma(d2z[u], d2z[u], fma(d2y[u], d2y[u], d2x[u] * d2x[u]));
                        far = max(far, max(__double2hiint(r1[u]), __double2hiint(r2[u])));
                    }
                    if (far >= far2_hi) {
#pragma unroll
                        for (int u = 0; u < G; ++u) {
                            if (maybe_wrapped(r1[u], far2_hi)) {
                                d1x[u] = min_image(d1x[u], c.L, c.halfL); d1y[u] = min_image(d1y[u], c.L, c.halfL);
                                d1z[u] = min_image(d1z[u], c.L, c.halfL);
                                r1[u] = fma(d1z[u], d1z[u], fma(d1y[u], d1y[u], d1x[u] * d1x[u]));
                            }
                            if (maybe_wrapped(r2[u], far2_hi)) {
                                d2x[u] = min_image(d2x[u], c.L, c.halfL); d2y[u] = min_image(d2y[u], c.L, c.halfL);
                                d2z[u] = min_image(d2z[u], c.L, c.halfL);
                                r2[u] = fma(d2z[u], d2z[u], fma(d2y[u], d2y[u], d2x[u] * d2x[u]));
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < G; ++u) {
                        const bool valid = (unsigned)(G * part + u) < nvalid;
                        const bool in1 = valid && (r1[u] < c.rc2) && (jj[u] != li1);
                        const bool in2 = valid && has2 && (r2[u] < c.rc2) && (jj[u] != li2);
                        {
                            const double inv = in1 ? fast_rcp(r1[u]) : 0.0;       // rejected: exact +0 contributions
                            const double s2 = UNIT_SIGMA ? inv : c.sigma2 * inv;
                            const double s6 = s2 * s2 * s2;
                            const double tt = fma(s6, s6, -s6);
                            const double fr = fma(s6, s6, tt) * inv;
                            e += tt;
                            a1x = fma(fr, d1x[u], a1x); a1y = fma(fr, d1y[u], a1y); a1z = fma(fr, d1z[u], a1z);
                        }
                        {
                            const double inv = in2 ? fast_rcp(r2[u]) : 0.0;
                            const double s2 = UNIT_SIGMA ? inv : c.sigma2 * inv;
                            const double s6 = s2 * s2 * s2;
                            const double tt = fma(s6, s6, -s6);
                            const double fr = fma(s6, s6, tt) * inv;
                            e += tt;
                            a2x = fma(fr, d2x[u], a2x); a2y = fma(fr, d2y[u], a2y); a2z = fma(fr, d2z[u], a2z);
                        }
                        nacc += (in1 ? 1u : 0u) + (in2 ? 1u : 0u);
                        if (MODE == TILE_FORCE_PRUNE) {
                            const bool k1 = (r1[u] <= c.prune_cut2) && (jj[u] != li1);
                            const bool k2 = has2 && (r2[u] <= c.prune_cut2) && (jj[u] != li2);
                            if (valid && (k1 || k2)) {
                                out16[(size_t)(cnt >> 3) * 256 + (cnt & 7u)] = (unsigned short)jj[u];
                                ++cnt;
                            }
                        }
                    }
                }
                cur = nxt;
            }
        } else {
            // list build / count: candidates are the 3x3 x-rows around the pair's home row, from the cell left of
            // atom 1's cell to the cell right of atom 2's cell (consecutive tile cells)
            unsigned h1 = 0;
            while (h1 + 1 < tg.bw && s_off[rbase + 2 + h1] <= li1) ++h1;
            unsigned h2 = h1;
            while (h2 + 1 < tg.bw && s_off[rbase + 2 + h2] <= li2) ++h2;
            for (unsigned rz = hz - 1; rz <= hz + 1; ++rz)
            for (unsigned ry = hy - 1; ry <= hy + 1; ++ry) {
                const unsigned r = (rz * H + ry) * W;
                const unsigned j0 = s_off[r + h1], j1 = s_off[r + h2 + 3];
                for (unsigned jb = j0; jb < j1; jb += 2) {
                    double r1[2], r2[2];
                    int far = 0;
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const unsigned j = min(jb + u, j1 - 1);
                        double jx, jy, jz;
                        tile_ld(sm, j, jx, jy, jz);
                        const double ax = p1x - jx, ay = p1y - jy, az = p1z - jz;
                        const double bx = p2x - jx, by = p2y - jy, bz = p2z - jz;
                        r1[u] = fma(az, az, fma(ay, ay, ax * ax));
                        r2[u] = fma(bz, bz, fma(by, by, bx * bx));
                        far = max(far, max(__double2hiint(r1[u]), __double2hiint(r2[u])));
                    }
                    if (far >= far2_hi) {
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const unsigned j = min(jb + u, j1 - 1);
                            double jx, jy, jz;
                            tile_ld(sm, j, jx, jy, jz);
                            if (maybe_wrapped(r1[u], far2_hi)) {
                                const double ax = min_image(p1x - jx, c.L, c.halfL), ay = min_image(p1y - jy, c.L, c.halfL);
                                const double az = min_image(p1z - jz, c.L, c.halfL);
                                r1[u] = fma(az, az, fma(ay, ay, ax * ax));
                            }
                            if (maybe_wrapped(r2[u], far2_hi)) {
                                const double bx = min_image(p2x - jx, c.L, c.halfL), by = min_image(p2y - jy, c.L, c.halfL);
                                const double bz = min_image(p2z - jz, c.L, c.halfL);
                                r2[u] = fma(bz, bz, fma(by, by, bx * bx));
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const unsigned j = jb + u;
                        const bool k1 = (r1[u] <= c.list_cut2) && (j != li1);
                        const bool k2 = has2 && (r2[u] <= c.list_cut2) && (j != li2);
                        if (j < j1 && (k1 || k2)) {
                            if (MODE == TILE_LIST_FILL && cnt < 8 * a.capq)
                                out16[(size_t)(cnt >> 3) * 256 + (cnt & 7u)] = (unsigned short)j;
                            ++cnt;
                            c1 += k1 ? 1u : 0u; c2 += k2 ? 1u : 0u;
                        }
                    }
                }
            }
        }

        if (MODE == TILE_FORCE_PRUNE) {
            for (unsigned k = cnt; k & 7u; ++k) out16[(size_t)(k >> 3) * 256 + (k & 7u)] = (unsigned short)li1;
            a.in_count[slot] = cnt;
        }
        if (MODE == TILE_FORCE_LIST || MODE == TILE_FORCE_PRUNE) {
            a.fx[i1] = a1x * c.prefctr; a.fy[i1] = a1y * c.prefctr; a.fz[i1] = a1z * c.prefctr;
            if (has2) { a.fx[i2] = a2x * c.prefctr; a.fy[i2] = a2y * c.prefctr; a.fz[i2] = a2z * c.prefctr; }
            etot += fma(-(double)nacc, c.ecut, e);
        } else {
            if (MODE == TILE_LIST_FILL) {
                for (unsigned k = min(cnt, 8 * a.capq); k & 7u; ++k) out16[(size_t)(k >> 3) * 256 + (k & 7u)] = (unsigned short)li1;
                a.atom_count[i1] = c1;
                if (has2) a.atom_count[i2] = c2;
            }
            a.nl_count[slot] = MODE == TILE_LIST_FILL ? min(cnt, 8 * a.capq) : cnt;
            maxcnt = max(maxcnt, cnt);
        }
    }

    if (MODE == TILE_FORCE_LIST || MODE == TILE_FORCE_PRUNE) {
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
