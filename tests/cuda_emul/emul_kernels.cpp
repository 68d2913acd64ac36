// emul_kernels.cpp - TEST INFRASTRUCTURE ONLY (see cuda_emul.h).  Runs the persistent brick kernel (lj_tile.cuh)
// of the product sources on the CPU shim for one configuration:
// binning by the reference's cell id, tile tables, list count + fill, a force pass over the outer list, a force
// pass that also prunes, a force pass over the pruned list.  The Python test compares the three force sets
// with each other (bit-identical) and with the oracle.
#include "cuda_emul.h"

#include "../../lennard-jones-live-simulator_b200/csrc/lj_tile.cuh"

#include <cstdio>
#include <numeric>

using namespace ljgpu;

namespace {

struct Problem {
    unsigned n, nc;
    LJConst c;
    std::vector<double4> pos;           // sorted by cell
    std::vector<uint32_t> orig, cstart, ccount;
};

Problem bin(unsigned n, double L, double rcut, double shell, double sigma, double epsilon, double prune_skin,
            const double* x, const double* y, const double* z) {
    Problem p;
    p.n = n;
    LJConst& c = p.c;
    std::memset(&c, 0, sizeof c);
    c.L = L; c.invL = 1.0 / L; c.halfL = L * 0.5;
    c.far2 = c.halfL * c.halfL * (1.0 - 1e-9);
    c.rc2 = rcut * rcut; c.sigma2 = sigma * sigma;
    const double sr2 = c.sigma2 / c.rc2, sr6 = sr2 * sr2 * sr2, sr12 = sr6 * sr6;
    c.ecut = sr12 - sr6; c.prefctr = 24.0 * epsilon; c.epsilon = epsilon; c.mass = 1.0;
    const double csc = rcut + shell;
    c.list_cut2 = csc * csc; c.disp_thresh = shell * shell * 0.25;
    c.prune_cut2 = (rcut + prune_skin) * (rcut + prune_skin); c.prune_half = 0.5 * prune_skin; c.dt = 0.0;
    unsigned ncells = std::max(1u, (unsigned)(L / csc));
    c.nc = ncells; c.cell_edge = L / (double)ncells; c.n = n;
    p.nc = ncells;
    const unsigned ncell = ncells * ncells * ncells;
    std::vector<uint32_t> key(n);
    for (unsigned i = 0; i < n; ++i) {                      // lennardjonessimulation.cpp:467-478
        const unsigned ix = std::min((unsigned)(x[i] / c.cell_edge), ncells - 1);
        const unsigned iy = std::min((unsigned)(y[i] / c.cell_edge), ncells - 1);
        const unsigned iz = std::min((unsigned)(z[i] / c.cell_edge), ncells - 1);
        key[i] = iz * ncells * ncells + iy * ncells + ix;
    }
    p.orig.resize(n);
    std::iota(p.orig.begin(), p.orig.end(), 0u);
    std::stable_sort(p.orig.begin(), p.orig.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
    p.pos.resize(n);
    p.cstart.assign(ncell, 0); p.ccount.assign(ncell, 0);
    for (unsigned s = 0; s < n; ++s) {
        const uint32_t i = p.orig[s];
        p.pos[s] = double4{x[i], y[i], z[i], 0.0};
        if (p.ccount[key[i]]++ == 0) p.cstart[key[i]] = s;
    }
    return p;
}

unsigned cdivu(unsigned a, unsigned b) { return (a + b - 1) / b; }

}  // namespace

// ctas: CTAs of the persistent kernel (fewer CTAs than bricks => every CTA streams several bricks through its ring).
// out_f: 3 passes x 3 components x n (original atom order): outer list, pruning pass, pruned list.
// out_epot: the three sums of (s12 - s6 - ecut) over ordered pairs.  out_counts: per-atom list length (n).
// Returns 0, or a negative code when a capacity / shape check fails.
extern "C" int ljemul_forces(unsigned ctas, unsigned n, double L, double rcut, double shell, double sigma, double epsilon,
                             double prune_skin, unsigned bx, unsigned by, unsigned bz,
                             const double* x, const double* y, const double* z,
                             double* out_f, double* out_epot, uint32_t* out_counts, uint32_t* out_info) {
    Problem p = bin(n, L, rcut, shell, sigma, epsilon, prune_skin, x, y, z);
    const unsigned nc = p.nc;
    if (bx < 1 || by < 1 || bz < 1 || bx > (unsigned)kTileBX || by > 2 || bz > 2) return -1;
    if (bx + 2 > nc || by + 2 > nc || bz + 2 > nc) return -2;              // a tile must not contain a cell twice
    const unsigned nbx = cdivu(nc, bx), nby = cdivu(nc, by), nbz = cdivu(nc, bz), nblocks = nbx * nby * nbz;
    std::vector<uint32_t> tsrc((size_t)nblocks * kTileCellsMax), toff((size_t)nblocks * (kTileCellsMax + 1));
    TileGrid g{nc, nc, 1u, 0u, nc, bx, by, bz, nbx, nby, nbz, nblocks, tsrc.data(), toff.data()};
    StepCtrl ctrl;
    std::memset(&ctrl, 0, sizeof ctrl);
    std::vector<double> f[3] = {std::vector<double>(n), std::vector<double>(n), std::vector<double>(n)};
    std::vector<double> partial;

    ljemul::launch(cdivu(nblocks, 4), 128, 0, [&] { tile_table_kernel(g, p.cstart.data(), p.ccount.data(), tsrc.data(), toff.data(), &ctrl); });
    if (ctrl.max_tile > 65535) return -3;
    const unsigned tcap = ((ctrl.max_tile + 32 + 63) / 64) * 64;
    const unsigned slots = 3, ipb = (ctrl.max_home + 31) / 32;
    const size_t smem = (size_t)tcap * kTileRecBytes * slots;
    partial.assign((size_t)nblocks * ipb, 0.0);
    const size_t ntiles = ((size_t)n + 31) / 32;
    std::vector<uint32_t> slot_count(ntiles * 32, 0), slot_in_count(ntiles * 32, 0);
    const unsigned grid = std::max(1u, std::min(ctas, nblocks));
    out_info[0] = nblocks; out_info[1] = ctrl.max_tile; out_info[2] = ctrl.max_home; out_info[3] = grid;

    auto run = [&](int mode, uint32_t* nl, uint32_t* nl_in, unsigned capq) {
        TileArgs a{p.pos.data(), nl, slot_count.data(), nl_in, slot_in_count.data(), f[0].data(), f[1].data(), f[2].data(),
                   partial.data(), &ctrl, capq, tcap, slots, ipb, (ctas & 1u) ? ipb / 2 : 0u, nullptr};
        ljemul::launch(grid, kBrickThreads, smem, [&] {
            switch (mode) {
                case TILE_LIST_COUNT:  tile_kernel<TILE_LIST_COUNT, false>(p.c, g, a, 0); break;
                case TILE_LIST_FILL:   tile_kernel<TILE_LIST_FILL, false>(p.c, g, a, 0); break;
                case TILE_FORCE_LIST:  tile_kernel<TILE_FORCE_LIST, true>(p.c, g, a, 1); break;
                case TILE_FORCE_PRUNE: tile_kernel<TILE_FORCE_PRUNE, true>(p.c, g, a, 1); break;
                case TILE_FORCE_CELL:  tile_kernel<TILE_FORCE_CELL, false>(p.c, g, a, 1); break;
            }
        });
    };
    auto collect = [&](int pass) {
        for (int k = 0; k < 3; ++k)
            for (unsigned s = 0; s < n; ++s) out_f[((size_t)pass * 3 + k) * n + p.orig[s]] = f[k][s];
        double e = 0.0;
        for (double v : partial) e += v;
        out_epot[pass] = e;
        for (auto& v : f) std::fill(v.begin(), v.end(), 0.0);
        std::fill(partial.begin(), partial.end(), 0.0);
    };

    run(TILE_LIST_COUNT, nullptr, nullptr, 0);
    const unsigned capq = (ctrl.max_count + 12 + 7) / 8;
    out_info[4] = ctrl.max_count; out_info[5] = capq;
    std::vector<uint32_t> nl(ntiles * capq * 32 * 4, 0xdeadbeefu), nl_in(ntiles * capq * 32 * 4, 0xdeadbeefu);
    ctrl.max_count = 0; ctrl.overflow = 0;
    run(TILE_LIST_FILL, nl.data(), nullptr, capq);
    if (ctrl.overflow) return -4;
    for (unsigned s = 0; s < n; ++s) out_counts[p.orig[s]] = slot_count[s];

    if (sigma != 1.0) return -5;                              // the force instantiations above are the unit-sigma ones
    run(TILE_FORCE_LIST, nl.data(), nullptr, capq);           // pass 0: outer list (no inner list offered)
    collect(0);
    run(TILE_FORCE_PRUNE, nl.data(), nl_in.data(), capq);     // pass 1: outer list, emits the inner one
    collect(1);
    ctrl.inner_valid = 1; ctrl.drift_bound = 0.0; ctrl.vmax2 = 0ull; ctrl.inner_used = 0;
    run(TILE_FORCE_LIST, nl.data(), nl_in.data(), capq);      // pass 2: inner list
    collect(2);
    out_info[6] = ctrl.inner_used;
    unsigned long long inner_total = 0;
    for (uint32_t v : slot_in_count) inner_total += v;
    out_info[7] = (uint32_t)std::min<unsigned long long>(inner_total, 0xffffffffull);
    run(TILE_FORCE_CELL, nullptr, nullptr, 0);                // pass 3: cell-list force kernel (no list)
    for (int k = 0; k < 3; ++k)
        for (unsigned s = 0; s < n; ++s) out_f[((size_t)3 * 3 + k) * n + p.orig[s]] = f[k][s];
    double e = 0.0;
    for (double v : partial) e += v;
    out_epot[3] = e;
    return 0;
}
