// lj_sim.cu - host orchestration of the sm_100a Lennard-Jones integrator and its C ABI (include/ljgpu.h).
//
// Path replaced: class LennardJonesSimulation, /root/reference/src/simulator/lennardjonessimulation.{h,cpp}
// (".cpp"/".h" below).  No CPU fallback: every entry point fails with LJGPU_ECUDA when there is no device.
#include "lj_tile.cuh"
#include "lj_nccl.h"
#include "../../include/ljgpu.h"

#include <algorithm>
#include <atomic>
#include <thread>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

using namespace ljgpu;

namespace {

thread_local std::string g_err;

struct LjError : std::runtime_error {
    int code;
    LjError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            throw LjError(LJGPU_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__) +       \
                                           " (" __FILE__ ":" + std::to_string(__LINE__) + ")");    \
    } while (0)

inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

enum Stage { ST_PREDICT = 0, ST_KICK_DRIFT, ST_FORCE_PRUNE, ST_FORCE, ST_KICK2, ST_REBUILD, ST_HALO, ST_COUNT };

template <class T>
void dfree(T*& p) { if (p) { cudaFree(p); p = nullptr; } }

#define NCK(call)                                                                                  \
    do {                                                                                           \
        int r__ = (call);                                                                          \
        if (r__ != 0)                                                                              \
            throw LjError(LJGPU_ENCCL, std::string(#call) + ": " + ljnccl::api().GetErrorString(r__)); \
    } while (0)

// z-slab decomposition state: cell layers [zlo, zlo+nzl) of the global grid belong to
// this device; one halo layer from each ring neighbour.
struct Slab {
    bool on = false;
    int rank = 0, world = 1, down = 0, up = 0;
    ljnccl::Comm comm = nullptr;
    unsigned zlo = 0, nzl = 0;
    unsigned c_lo = 0, c_hi = 0;          // atoms in my first / last owned layer (what I send every step)
    unsigned h_lo = 0, h_hi = 0;          // atoms in my lower / upper halo layer (what I receive every step)
    double* sendbuf[2] = {nullptr, nullptr};
    double* recvbuf[2] = {nullptr, nullptr};
    size_t mover_cap = 0;                 // atoms per mover buffer
    uint32_t* hcount[2] = {nullptr, nullptr};   // per-cell counts of the halo layers as received
    uint32_t* hoff[2] = {nullptr, nullptr};     // their exclusive scans
    uint32_t* xchg = nullptr;             // device scratch for count exchanges (16 words)
    uint32_t* h_xchg = nullptr;           // pinned mirror
    // peer-memory exchange (NVLink loads/stores through CUDA IPC mappings); NCCL remains for re-binning
    bool p2p = false;
    Mailbox* mail = nullptr;              // mine (peers write into it)
    Mailbox* peer_mail[kMaxPeers] = {nullptr};
    Mailbox** d_peer_mail = nullptr;      // device copy of the pointer table
    double4* peer_pos[2] = {nullptr, nullptr};    // position arrays of my lower / upper neighbour
    void* ipc_opened[3 * kMaxPeers + 2] = {nullptr};  // mappings to close
    int n_ipc_opened = 0;
    unsigned nb_n[2] = {0, 0}, nb_hlo[2] = {0, 0};   // owned atoms and lower-halo size of the two neighbours
    unsigned long long pub_seq = 0;
    unsigned int* push_ticket = nullptr;
    // publication (ljgpu_step_publish): my slice of the original-order frame, two buffers; the peers' slices
    unsigned pub_S = 0;                    // atoms per slice: ceil(N / world)
    double* pub_slice[2] = {nullptr, nullptr};
    double* peer_pub[2][kMaxPeers] = {{nullptr}};
    unsigned int* pub_ticket = nullptr;
};

} // namespace

static void set_tile_attributes(ljgpu_sim* s);

// LJGPU_TRACE_HOST=1: wall-clock spent by the stepping thread per host section (totals, maxima and every
// event above 2 ms), printed when the simulation is destroyed.  Adds no synchronisation.
struct HostTrace {
    enum { ENQUEUE, CAPTURE, GRAPH_LAUNCH, FETCH, REBUILD, NSEC };
    bool on = std::getenv("LJGPU_TRACE_HOST") != nullptr;
    double tot[NSEC] = {}, mx[NSEC] = {};
    uint64_t cnt[NSEC] = {};
    std::vector<std::pair<int, double>> slow;
    struct Scope {
        HostTrace& t; int sec; std::chrono::steady_clock::time_point t0;
        Scope(HostTrace& tr, int s) : t(tr), sec(s) { if (t.on) t0 = std::chrono::steady_clock::now(); }
        ~Scope() {
            if (!t.on) return;
            const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
            t.tot[sec] += us; t.mx[sec] = std::max(t.mx[sec], us); ++t.cnt[sec];
            if (us > 2000.0 && t.slow.size() < 64) t.slow.emplace_back(sec, us);
        }
    };
    void report(int rank) const {
        if (!on) return;
        static const char* names[NSEC] = {"enqueue", "capture", "graph_launch", "fetch", "rebuild"};
        for (int k = 0; k < NSEC; ++k)
            if (cnt[k]) std::fprintf(stderr, "[ljgpu host r%d] %-12s n=%8llu total %10.1f us  max %9.1f us\n", rank, names[k],
                                     (unsigned long long)cnt[k], tot[k], mx[k]);
        for (auto& e : slow) std::fprintf(stderr, "[ljgpu host r%d]   slow %-12s %9.1f us\n", rank, names[e.first], e.second);
    }
};

struct ljgpu_sim {
    std::mutex mu;
    // A query from another thread (the GUI's get_velocities_copy ...) must not wait for a whole ljgpu_step_n call: callers
    // announce themselves in `waiting`, and the stepping loop hands the lock over between two batches (a batch is a few
    // dozen steps: the state is complete and consistent there).
    std::atomic<int> waiting{0};
    std::unique_lock<std::mutex>* held = nullptr;       // the lock of the call that is executing (set by with_sim)
    ljgpu_params p{};
    int mode = 0;
    int device = 0;
    cudaStream_t st = nullptr;
    LJConst c{};
    unsigned n = 0;                 // atoms owned by this device
    unsigned nglobal = 0;           // atoms of the whole system
    unsigned nc = 0;                // cells per axis (global)
    unsigned ncz = 0;               // cell layers of the local table
    unsigned ncell = 0;             // cells of the local table (+2 mover bins on a slab)
    size_t cap_atoms = 0, pos_cap = 0;
    Slab slab;
    bool failed = false;            // a call on this handle threw: NCCL state is undefined
    bool energies_fresh = false;    // h_ctrl holds the energies of the last completed step (run_steps ends with a fetch): a query needs no device round trip
    // slab path: a capacity check that fails on ONE device must not make it leave a sequence of collectives the others are
    // still in.  The failing device records the reason, keeps every exchange well-formed (skips what does not fit), and all
    // devices learn about it in agree() - one small all-reduce per re-binning - and throw the same error together.
    unsigned slab_fail = 0;
    void slab_note_failure(unsigned code) { if (!slab_fail) slab_fail = code; }
    void agree();

    // state, device order (cell-sorted on the list paths, caller order on the all-pairs path)
    double4* pos = nullptr; double4* pos_tmp = nullptr;
    double* dij[3] = {nullptr, nullptr, nullptr};      // accumulated displacement since the last binning (.h:94)
    double* v[3] = {nullptr, nullptr, nullptr};
    double* f[3] = {nullptr, nullptr, nullptr};
    double* v_alt[3] = {nullptr, nullptr, nullptr};
    double* f_alt[3] = {nullptr, nullptr, nullptr};
    uint32_t* orig = nullptr; uint32_t* orig_alt = nullptr;

    // binning
    uint32_t *keys = nullptr, *vals = nullptr;
    uint32_t *cstart = nullptr, *cend = nullptr, *ccount = nullptr;
    SortScratch scr;

    // shared-memory tiles (lj_tile.cuh) and the Verlet list of 16-bit tile indices
    uint32_t *tsrc = nullptr, *toff = nullptr;
    unsigned tile_attr_tcap = 0, tile_attr_slots = 0;          // the tile shape the kernels' attributes were last set for
    unsigned bxd = kTileBX, byd = kTileBY, bzd = kTileBZ;      // brick size in cells
    unsigned nbx = 0, nby = 0, nbz = 0, nblocks_tile = 0, tcap = 0, tile_slots = 2, tile_ipb = 1;
    int num_sms = 148;
    unsigned tile_grid_ctas[8][2] = {{0}};     // CTAs of the persistent brick kernel per (mode, unit sigma): resident CTAs per SM x SMs
    uint32_t* nl = nullptr; uint32_t* nl_count = nullptr; unsigned capq = 0; size_t nl_alloc = 0;
    // pruned inner list (single-device Verlet path)
    uint32_t* nl_in = nullptr; uint32_t* nl_in_count = nullptr;
    double prune_skin = 0.0; unsigned prune_interval = 8, since_prune = 0;

    // reductions
    double *partial_epot = nullptr, *kick_sums = nullptr, *partial_pred = nullptr;     // kick_sums: [3][blocks of kick2]
    unsigned nb_force = 0;
    double* apf[3] = {nullptr, nullptr, nullptr}; unsigned nsplit = 1, jchunk = 0;

    StepCtrl* ctrl = nullptr; StepCtrl* h_ctrl = nullptr;
    HistEntry* hist = nullptr; HistEntry* samples = nullptr;
    SeriesStatsOut* d_stats = nullptr; double* d_mb = nullptr;       // device statistics of the GraphWidget series / speed histogram

    // scratch of the neighbour-count probe (allocated on first use): the probe bins a COPY of the published positions,
    // so it never touches the integrator's own binning, lists or displacement accumulators
    double4 *probe_pos = nullptr, *probe_sorted = nullptr;
    uint32_t *probe_keys = nullptr, *probe_vals = nullptr, *probe_cstart = nullptr, *probe_cend = nullptr, *probe_ccount = nullptr;
    void probe_neighbor_counts(double cut2, int inclusive, uint32_t* host_out);
    // staging for queries
    double* out3 = nullptr; uint32_t* out_u32 = nullptr; unsigned long long* hist_counts = nullptr;
    // asynchronous publication of positions (ljgpu_step_publish): staging buffers, copy stream, events
    double* pub3[2] = {nullptr, nullptr};
    cudaStream_t st_copy = nullptr;
    cudaEvent_t ev_pub_ready[2] = {nullptr, nullptr}, ev_pub_done[2] = {nullptr, nullptr};
    int pub_cur = 0; bool pub_pending = false; uint32_t pub_step = 0;
    void step_publish(unsigned step, double dt, double* x, double* y, double* z);
    void publish_wait();

    // CUDA graphs of one step (plain / pruning), re-captured after every re-binning or dt change
    // Captured steps, keyed by everything a launch bakes in (buffer pointers alternate between two sets from
    // one re-binning to the next, launch shapes only grow): after a few re-binnings every step replays a cached
    // graph and nothing is instantiated any more (an instantiation now and then took tens of milliseconds).
    struct GraphKey {
        const void* ptr[20]; unsigned shape[10]; double dt; uint64_t gen; int prune;
        bool operator==(const GraphKey& o) const { return std::memcmp(this, &o, sizeof(GraphKey)) == 0; }
    };
    struct StepGraph { cudaGraphExec_t exec = nullptr; uint64_t kernels = 0; uint64_t cnt[ST_COUNT] = {0}; GraphKey key; uint64_t used = 0; };
    std::vector<StepGraph> graphs;
    uint64_t graph_gen = 0, graph_clock = 0;          // gen: bumped when constants change (all cached graphs are stale)
    GraphKey graph_key(bool prune_now, double dt) const;
    unsigned* h_scalar = nullptr;     // pinned scratch for 4-byte uploads

    // stepping bookkeeping
    bool pred_valid = false; double pred_dt = 0.0;
    unsigned since_rebuild = 0, last_interval = 16;
    bool dirty_since_rebuild = false;

    // profiling
    bool prof = false;
    unsigned prof_every = 1;        // profile every k-th step (the others replay the CUDA graph); stage times are scaled by k
    unsigned prof_tick = 0;
    bool prof_this_step = true;
    bool in_rebuild = false;
    struct Pending { int stage; cudaEvent_t a, b; double weight; };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> pool;
    double ms[ST_COUNT] = {0};
    uint64_t cnt[ST_COUNT] = {0};
    uint64_t steps = 0, rebuilds = 0, launches = 0;

    HostTrace htrace;
    ~ljgpu_sim();

    // ---- helpers ------------------------------------------------------------------------------
    cudaEvent_t get_event() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; CK(cudaEventCreate(&e)); return e;
    }
    // stage timers nest (a re-binning contains halo exchanges): the caller keeps its start event
    cudaEvent_t stage_begin() {
        if (!prof || !prof_this_step) return nullptr;
        cudaEvent_t a = get_event();
        CK(cudaEventRecord(a, st));
        return a;
    }
    void stage_end(int stage, cudaEvent_t a) {
        cnt[stage]++;
        if (prof && a) {
            cudaEvent_t b = get_event();
            CK(cudaEventRecord(b, st));
            // step stages are sampled every prof_every-th step; everything inside a re-binning is always timed
            const bool sampled = stage != ST_REBUILD && !in_rebuild;
            pending.push_back({stage, a, b, sampled ? (double)prof_every : 1.0});
        }
    }
    void resolve_pending() {      // stream must be idle
        for (auto& q : pending) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, q.a, q.b) == cudaSuccess) ms[q.stage] += t * q.weight;
            else cudaGetLastError();
            pool.push_back(q.a); pool.push_back(q.b);
        }
        pending.clear();
    }
    // Busy-wait on the stream instead of a blocking synchronize: the step loop syncs a few times per
    // re-binning and a sleeping host thread that is slow to wake up costs more than the spin.
    void sync() {
        cudaError_t e;
        while ((e = cudaStreamQuery(st)) == cudaErrorNotReady) { }
        if (e != cudaSuccess) throw LjError(LJGPU_ECUDA, std::string("stream: ") + cudaGetErrorString(e));
        resolve_pending();
    }
    void check_launch(const char* what) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) throw LjError(LJGPU_ECUDA, std::string(what) + " launch: " + cudaGetErrorString(e));
        ++launches;
    }

    bool list_mode() const { return mode != LJGPU_KERNEL_ALLPAIRS; }
    unsigned nb_atoms() const { return std::max(1u, cdiv(n, kBlock)); }
    TileGrid tile_grid() const {
        return TileGrid{nc, ncz, slab.on ? 0u : 1u, slab.on ? 1u : 0u, slab.on ? slab.nzl : nc,
                        bxd, byd, bzd, nbx, nby, nbz, nblocks_tile, tsrc, toff};
    }

    void setup_constants();
    void allocate();
    void upload_state(const double* x, const double* y, const double* z,
                      const double* vx, const double* vy, const double* vz);
    void sort_by_cell();
    void migrate();
    void halo_setup();
    void halo_exchange();
    void allreduce_red();
    void setup_p2p();
    void slab_finalize(int what, unsigned nb_epot, unsigned nb_kick);
    SlabRed slab_red() const { return SlabRed{slab.mail, slab.d_peer_mail, slab.rank, slab.world}; }
    void rebuild();
    void build_list();
    void launch_force(bool timed);
    void initial_forces();
    void finalize_epot_only();
    void launch_predict(double dt);
    void enqueue_step(bool prune_now, double dt);
    void launch_step(double dt);
    void drop_graphs();
    void run_steps(unsigned first, unsigned nsteps, double dt);
    void fetch_ctrl() {
        CK(cudaMemcpyAsync(h_ctrl, ctrl, sizeof(StepCtrl), cudaMemcpyDeviceToHost, st));
        unsigned err = 0;
        if (slab.p2p) CK(cudaMemcpyAsync(&err, &slab.mail->error, 4, cudaMemcpyDeviceToHost, st));
        sync();
        if (err) throw LjError(LJGPU_ENCCL, "a peer device stopped answering (mailbox wait timed out)");
    }
    void export3(int what, double* a, double* b, double* cc);
};

ljgpu_sim::~ljgpu_sim() {
    htrace.report(slab.on ? slab.rank : 0);
    cudaSetDevice(device);
    if (st) cudaStreamSynchronize(st);
    // after a failure the peers may be inside a collective this rank will never join: abort, do not wait
    if (slab.comm) { if (failed) ljnccl::api().CommAbort(slab.comm); else ljnccl::api().CommDestroy(slab.comm); }
    dfree(pos); dfree(pos_tmp);
    for (int k = 0; k < 3; ++k) { dfree(dij[k]); dfree(v[k]); dfree(f[k]); dfree(v_alt[k]); dfree(f_alt[k]); dfree(apf[k]); }
    dfree(orig); dfree(orig_alt); dfree(keys); dfree(vals);
    dfree(cstart); dfree(cend); dfree(ccount);
    dfree(nl); dfree(nl_count); dfree(nl_in); dfree(nl_in_count); dfree(tsrc); dfree(toff);
    dfree(partial_epot); dfree(kick_sums); dfree(partial_pred);
    dfree(ctrl); dfree(hist); dfree(samples); dfree(d_stats); dfree(d_mb); dfree(out3); dfree(out_u32); dfree(hist_counts);
    dfree(probe_pos); dfree(probe_sorted); dfree(probe_keys); dfree(probe_vals); dfree(probe_cstart); dfree(probe_cend); dfree(probe_ccount);
    if (st_copy) { cudaStreamSynchronize(st_copy); cudaStreamDestroy(st_copy); }
    for (int b = 0; b < 2; ++b) {
        dfree(pub3[b]);
        if (ev_pub_ready[b]) cudaEventDestroy(ev_pub_ready[b]);
        if (ev_pub_done[b]) cudaEventDestroy(ev_pub_done[b]);
    }
    for (int k = 0; k < 2; ++k) { dfree(slab.sendbuf[k]); dfree(slab.recvbuf[k]); dfree(slab.hcount[k]); dfree(slab.hoff[k]); }
    dfree(slab.xchg);
    for (int k = 0; k < slab.n_ipc_opened; ++k) cudaIpcCloseMemHandle(slab.ipc_opened[k]);
    dfree(slab.mail); dfree(slab.d_peer_mail); dfree(slab.push_ticket); dfree(slab.pub_ticket); dfree(slab.pub_slice[0]); dfree(slab.pub_slice[1]);
    if (slab.h_xchg) cudaFreeHost(slab.h_xchg);
    if (h_ctrl) cudaFreeHost(h_ctrl);
    if (h_scalar) cudaFreeHost(h_scalar);
    for (auto& g : graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    sort_scratch_free(scr);
    for (auto& q : pending) { cudaEventDestroy(q.a); cudaEventDestroy(q.b); }
    for (auto e : pool) cudaEventDestroy(e);
    if (st) cudaStreamDestroy(st);
}

void ljgpu_sim::setup_constants() {
    // Every expression below is the host-side scalar set-up of the cited reference lines.
    const double L = p.cell_length;
    c.L = L; c.invL = 1.0 / L; c.halfL = L * 0.5;                       // .cpp:424-426, :295
    c.far2 = c.halfL * c.halfL * (1.0 - 1e-9);
    c.rc2 = p.rcut * p.rcut;                                            // .cpp:262
    c.sigma2 = p.sigma * p.sigma;                                       // .cpp:264
    const double sr2 = c.sigma2 / c.rc2, sr6 = sr2 * sr2 * sr2, sr12 = sr6 * sr6;
    c.ecut = sr12 - sr6;                                                // .cpp:270-273
    c.prefctr = 24.0 * p.epsilon;                                       // .cpp:267
    c.epsilon = p.epsilon;
    c.mass = p.mass;
    c.ekin0 = 1.5 * (double)((size_t)p.nr_particles - 1) * p.kT;        // .cpp:374
    const double csc = p.rcut + p.shell;                                // .cpp:443-445
    c.list_cut2 = csc * csc;
    c.disp_thresh = p.shell * p.shell * 0.25;                           // .cpp:559-560
    {   // inner-list skin: LJGPU_PRUNE_SKIN (in sigma units of length; 0 disables), default 0.12, never beyond the shell
        const char* e = std::getenv("LJGPU_PRUNE_SKIN");
        prune_skin = e ? std::atof(e) : 0.12;
        // on slabs only the peer-memory step loop carries the neighbours' speeds (NCCL fallback: no pruning)
        if (prune_skin >= p.shell || (slab.on && !slab.p2p) || mode != LJGPU_KERNEL_VERLET) prune_skin = 0.0;
        c.prune_cut2 = (p.rcut + prune_skin) * (p.rcut + prune_skin);
        c.prune_half = 0.5 * prune_skin;
        c.dt = 0.0;
    }
    unsigned ncells = (unsigned)(L / csc);                              // .cpp:447-453
    ncells = std::max(1u, ncells);
    c.nc = ncells;
    c.cell_edge = L / (double)ncells;                                   // .cpp:455-457
    c.n = n;
    c.slab = slab.on ? 1u : 0u; c.zlo = slab.zlo; c.nzl = slab.nzl;
    nc = ncells;
    nglobal = (unsigned)p.nr_particles;
    ncz = slab.on ? slab.nzl + 2 : nc;
    ncell = nc * nc * ncz + (slab.on ? 2u : 0u);
}

void ljgpu_sim::allocate() {
    const size_t N = nglobal;
    if (slab.on) {
        cap_atoms = N / slab.world + N / slab.world / 4 + 65536;
        const size_t layer = N / nc + 1;
        pos_cap = cap_atoms + 2 * (layer + layer / 2 + 4096);
    } else {
        cap_atoms = N; pos_cap = N;
    }
    const size_t A = cap_atoms;
    const unsigned nbmax = cdiv(A, kBlock);
    CK(cudaMalloc(&pos, pos_cap * sizeof(double4)));
    CK(cudaMemset(pos, 0, pos_cap * sizeof(double4)));
    for (int k = 0; k < 3; ++k) {
        CK(cudaMalloc(&v[k], A * 8)); CK(cudaMalloc(&f[k], A * 8));
        CK(cudaMemset(f[k], 0, A * 8));
    }
    CK(cudaMalloc(&orig, A * 4));
    CK(cudaMalloc(&ctrl, sizeof(StepCtrl))); CK(cudaMemset(ctrl, 0, sizeof(StepCtrl)));
    CK(cudaMallocHost(&h_ctrl, sizeof(StepCtrl))); std::memset(h_ctrl, 0, sizeof(StepCtrl));
    CK(cudaMallocHost(&h_scalar, 64));
    CK(cudaMalloc(&hist, sizeof(HistEntry) * kHistCap)); CK(cudaMemset(hist, 0, sizeof(HistEntry) * kHistCap));
    CK(cudaMalloc(&samples, sizeof(HistEntry) * kSampleCap)); CK(cudaMemset(samples, 0, sizeof(HistEntry) * kSampleCap));
    CK(cudaMalloc(&d_stats, sizeof(SeriesStatsOut))); CK(cudaMalloc(&d_mb, sizeof(double) * (kMaxBins + 2)));
    {   // the sampled series: every 1000th step, as ThreadIntegrate::run emits them (threadintegrate.cpp:57-62)
        StepCtrl init{};
        init.samples = samples; init.sample_stride = 1000u;
        CK(cudaMemcpy(ctrl, &init, sizeof(StepCtrl), cudaMemcpyHostToDevice));
    }
    CK(cudaMalloc(&out3, 3 * N * 8));
    CK(cudaMalloc(&out_u32, N * 4));
    CK(cudaMalloc(&hist_counts, kMaxBins * sizeof(unsigned long long)));
    CK(cudaMalloc(&kick_sums, (size_t)3 * cdiv(A, kKick2Block) * 8));
    CK(cudaMalloc(&partial_pred, (size_t)nbmax * 8));

    if (list_mode()) {
        CK(cudaMalloc(&pos_tmp, A * sizeof(double4)));
        for (int k = 0; k < 3; ++k) {
            CK(cudaMalloc(&dij[k], A * 8)); CK(cudaMalloc(&v_alt[k], A * 8)); CK(cudaMalloc(&f_alt[k], A * 8));
        }
        CK(cudaMalloc(&orig_alt, A * 4));
        CK(cudaMalloc(&keys, A * 4)); CK(cudaMalloc(&vals, A * 4));
        CK(cudaMalloc(&cstart, (size_t)ncell * 4)); CK(cudaMalloc(&cend, (size_t)ncell * 4)); CK(cudaMalloc(&ccount, (size_t)ncell * 4));
        CK(cudaMalloc(&nl_count, A * 4));
        if (mode == LJGPU_KERNEL_VERLET) CK(cudaMalloc(&nl_in_count, A * 4));
        const unsigned nhz = slab.on ? slab.nzl : nc;
        // brick = 3x2x2 cells when there are plenty of them; small systems (and thin slabs) get smaller bricks
        // so that the grid still covers the 148 SMs several times over
        bxd = kTileBX; byd = kTileBY; bzd = kTileBZ;
        auto count_blocks = [&] { nbx = cdiv(nc, bxd); nby = cdiv(nc, byd); nbz = cdiv(nhz, bzd); nblocks_tile = nbx * nby * nbz; };
        count_blocks();
        while (nblocks_tile < 2u * 148u && bxd * byd * bzd > 1) {     // (one persistent CTA per SM streams its bricks: two or more bricks per CTA keep its ring busy)
            if (bxd > 1 && bxd >= byd && bxd >= bzd) --bxd; else if (byd > 1 && byd >= bzd) --byd; else --bzd;
            count_blocks();
        }
        if (const char* e = std::getenv("LJGPU_BRICK")) {          // tuning knob: "bx,by,bz" in cells, within the compiled maxima
            unsigned ex = 0, ey = 0, ez = 0;
            if (std::sscanf(e, "%u,%u,%u", &ex, &ey, &ez) == 3 && ex >= 1 && ey >= 1 && ez >= 1 &&
                ex <= (unsigned)kTileBX && ey <= (unsigned)kTileBY && ez <= (unsigned)kTileBZ) {
                bxd = ex; byd = ey; bzd = ez; count_blocks();
            }
        }
        CK(cudaMalloc(&tsrc, (size_t)nblocks_tile * kTileCellsMax * 4));
        CK(cudaMalloc(&toff, (size_t)nblocks_tile * (kTileCellsMax + 1) * 4));
        nb_force = nblocks_tile * kBrickItemsMax;                  // one partial per (brick, work item): room for the most items a brick may have
    } else {
        // j splits: enough blocks to fill 148 SMs a few times over, each split at least 64 atoms wide
        const unsigned nba = cdiv(N, kBlock);
        // j splits: between four and eight blocks per SM, each split at least 64 atoms wide; among those the count whose grid
        // fills whole waves of 148 SMs best (N = 4096: 32 x 37 = 1184 = 8 x 148 blocks; measured 50.8 us per force pass against
        // 54.6 us with 32 x 18)
        static const int ap_env = std::getenv("LJGPU_AP_SPLITS") ? std::atoi(std::getenv("LJGPU_AP_SPLITS")) : 0;   // tuning knob
        const unsigned sms = (unsigned)num_sms;
        const unsigned lo = std::max(1u, (4u * sms + nba - 1) / nba), hi = std::max(lo, (8u * sms + nba - 1) / nba);
        unsigned best = lo; double best_fill = 0.0;
        for (unsigned sp = lo; sp <= hi; ++sp) {
            const unsigned blocks = nba * sp;
            const double fill = (double)blocks / (double)(cdiv(blocks, sms) * sms);
            if (fill >= best_fill) { best_fill = fill; best = sp; }
        }
        nsplit = std::max(1u, std::min(ap_env > 0 ? (unsigned)ap_env : best, (unsigned)((N + 63) / 64)));
        jchunk = cdiv(N, nsplit);
        nsplit = cdiv(N, jchunk);
        for (int k = 0; k < 3; ++k) CK(cudaMalloc(&apf[k], (size_t)nsplit * N * 8));
        nb_force = nba * nsplit;
    }
    CK(cudaMalloc(&partial_epot, (size_t)std::max(nb_force, 1u) * 8));
    CK(cudaMemset(partial_epot, 0, (size_t)std::max(nb_force, 1u) * 8));

    if (slab.on) {
        const size_t layer = N / nc + 1;
        slab.mover_cap = layer / 4 + 4096;
        for (int k = 0; k < 2; ++k) {
            CK(cudaMalloc(&slab.sendbuf[k], slab.mover_cap * kMoverDoubles * 8));
            CK(cudaMalloc(&slab.recvbuf[k], slab.mover_cap * kMoverDoubles * 8));
            CK(cudaMalloc(&slab.hcount[k], (size_t)nc * nc * 4));
            CK(cudaMalloc(&slab.hoff[k], (size_t)nc * nc * 4));
        }
        CK(cudaMalloc(&slab.xchg, 16 * 4));
        CK(cudaMallocHost(&slab.h_xchg, 16 * 4));
    }
}

// Host-side owner test of the slab path: the cell layer of a wrapped coordinate, same arithmetic as
// wrap_coord + cell_coord on the device (this translation unit is built without FMA contraction on the host).
static unsigned host_layer(double z, double L, double invL, double edge, unsigned ncells) {
    const double w = z - L * std::floor(z * invL);
    unsigned i = (unsigned)(w / edge);
    return i < ncells - 1 ? i : ncells - 1;
}

void ljgpu_sim::upload_state(const double* x, const double* y, const double* z,
                             const double* vx, const double* vy, const double* vz) {
    energies_fresh = false;
    // everything but the slab sequence counters at the end of the block: the peers' mailboxes remember the sequence numbers
    CK(cudaMemsetAsync(ctrl, 0, offsetof(StepCtrl, halo_seq), st));
    pred_valid = false;
    since_rebuild = 0;
    if (!slab.on) {
        const size_t N = nglobal, B = N * 8;
        n = nglobal; c.n = n;
        CK(cudaMemcpyAsync(out3, x, B, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(out3 + N, y, B, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(out3 + 2 * N, z, B, cudaMemcpyHostToDevice, st));
        import_state_kernel<<<nb_atoms(), kBlock, 0, st>>>(n, out3, out3 + N, out3 + 2 * N, pos, orig);
        check_launch("import_state");
        CK(cudaMemcpyAsync(v[0], vx, B, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(v[1], vy, B, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(v[2], vz, B, cudaMemcpyHostToDevice, st));
        for (int k = 0; k < 3; ++k) CK(cudaMemsetAsync(f[k], 0, B, st));
        return;
    }
    // slab: every rank is handed the whole system and keeps the atoms of its cell layers
    std::vector<double> hx, hy, hz, hvx, hvy, hvz;
    std::vector<uint32_t> ho;
    for (uint32_t i = 0; i < nglobal; ++i) {
        const unsigned lz = (host_layer(z[i], c.L, c.invL, c.cell_edge, nc) + nc - slab.zlo) % nc;
        if (lz >= slab.nzl) continue;
        hx.push_back(x[i]); hy.push_back(y[i]); hz.push_back(z[i]);
        hvx.push_back(vx[i]); hvy.push_back(vy[i]); hvz.push_back(vz[i]);
        ho.push_back(i);
    }
    n = (unsigned)ho.size(); c.n = n;
    if (n > cap_atoms) throw LjError(LJGPU_ESTATE, "slab holds more atoms than its capacity (very inhomogeneous start state)");
    const size_t B = (size_t)n * 8;
    // stage through out3 (3 N doubles): x y z, then import; velocities straight into place
    CK(cudaMemcpy(out3, hx.data(), B, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(out3 + n, hy.data(), B, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(out3 + 2 * (size_t)n, hz.data(), B, cudaMemcpyHostToDevice));
    import_state_kernel<<<nb_atoms(), kBlock, 0, st>>>(n, out3, out3 + n, out3 + 2 * (size_t)n, pos, orig);
    check_launch("import_state");
    CK(cudaMemcpyAsync(orig, ho.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(v[0], hvx.data(), B, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(v[1], hvy.data(), B, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(v[2], hvz.data(), B, cudaMemcpyHostToDevice, st));
    for (int k = 0; k < 3; ++k) CK(cudaMemsetAsync(f[k], 0, B, st));
    CK(cudaStreamSynchronize(st));        // host vectors go out of scope
}

// wrap (.cpp:419-435), cell keys (.cpp:467-478), stable radix sort, permutation of the whole state,
// cell table of the owned atoms.
void ljgpu_sim::sort_by_cell() {
    wrap_key_kernel<<<nb_atoms(), kBlock, 0, st>>>(c, pos, pos_tmp, keys, vals);
    check_launch("wrap_key");
    int bits = 1; while ((1ull << bits) < (unsigned long long)ncell) ++bits;
    launches += radix_sort_pairs(keys, vals, n, bits, scr, st);
    CK((cudaError_t)sort_take_error());
    CK(cudaMemsetAsync(cstart, 0, (size_t)ncell * 4, st));
    CK(cudaMemsetAsync(cend, 0, (size_t)ncell * 4, st));
    if (n) {
        cell_bounds_kernel<<<nb_atoms(), kBlock, 0, st>>>(n, keys, cstart, cend);
        check_launch("cell_bounds");
    }
    permute_kernel<<<nb_atoms(), kBlock, 0, st>>>(n, vals, pos_tmp, pos,
                                                  v[0], v[1], v[2], f[0], f[1], f[2], orig,
                                                  v_alt[0], v_alt[1], v_alt[2], f_alt[0], f_alt[1], f_alt[2], orig_alt);
    check_launch("permute");
    for (int k = 0; k < 3; ++k) { std::swap(v[k], v_alt[k]); std::swap(f[k], f_alt[k]); }
    std::swap(orig, orig_alt);
    cell_count_kernel<<<cdiv(ncell, kBlock), kBlock, 0, st>>>(ncell, cstart, cend, ccount);
    check_launch("cell_count");
}

// Slab path: hand the atoms that left my cell layers to the ring neighbours (they sit at the tail of the
// sorted state: keys ncell-2 = down, ncell-1 = up), take theirs.  {x,v,f,orig} travel, 88 bytes per atom.
void ljgpu_sim::migrate() {
    auto& A = ljnccl::api();
    uint32_t hc[4];
    CK(cudaMemcpyAsync(hc, ccount + (ncell - 2), 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hc + 2, cstart + (ncell - 2), 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const unsigned n_down = hc[0], n_up = hc[1];
    const unsigned first_down = n_down ? hc[2] : n, first_up = n_up ? hc[3] : n;
    const unsigned n_stay = n - n_down - n_up;
    // counts first.  A message that would not fit the mover buffers is dropped by BOTH ends of the link (each sees the
    // same count against the same capacity), the failure is recorded and reported by agree().
    if (n_down > slab.mover_cap || n_up > slab.mover_cap) slab_note_failure(1);
    slab.h_xchg[0] = n_down; slab.h_xchg[1] = n_up;
    CK(cudaMemcpyAsync(slab.xchg, slab.h_xchg, 8, cudaMemcpyHostToDevice, st));
    NCK(A.GroupStart());
    NCK(A.Send(slab.xchg + 0, 1, ljnccl::kUint32, slab.down, slab.comm, st));
    NCK(A.Send(slab.xchg + 1, 1, ljnccl::kUint32, slab.up, slab.comm, st));
    NCK(A.Recv(slab.xchg + 2, 1, ljnccl::kUint32, slab.up, slab.comm, st));      // what the upper neighbour sends down to me
    NCK(A.Recv(slab.xchg + 3, 1, ljnccl::kUint32, slab.down, slab.comm, st));    // what the lower neighbour sends up to me
    NCK(A.GroupEnd());
    CK(cudaMemcpyAsync(slab.h_xchg + 2, slab.xchg + 2, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (slab.h_xchg[2] > slab.mover_cap || slab.h_xchg[3] > slab.mover_cap) slab_note_failure(1);
    auto fits = [&](unsigned k) { return k > slab.mover_cap ? 0u : k; };
    const unsigned s_down = fits(n_down), s_up = fits(n_up);                      // what actually travels
    unsigned from_up = fits(slab.h_xchg[2]), from_down = fits(slab.h_xchg[3]);

    if (s_down) { pack_movers_kernel<<<cdiv(s_down, kBlock), kBlock, 0, st>>>(first_down, s_down, pos, v[0], v[1], v[2], f[0], f[1], f[2], orig, slab.sendbuf[0]); check_launch("pack_movers"); }
    if (s_up)   { pack_movers_kernel<<<cdiv(s_up, kBlock), kBlock, 0, st>>>(first_up, s_up, pos, v[0], v[1], v[2], f[0], f[1], f[2], orig, slab.sendbuf[1]); check_launch("pack_movers"); }
    // data; at least one record per message so both sides always post matching calls
    NCK(A.GroupStart());
    NCK(A.Send(slab.sendbuf[0], (size_t)std::max(s_down, 1u) * kMoverDoubles, ljnccl::kFloat64, slab.down, slab.comm, st));
    NCK(A.Send(slab.sendbuf[1], (size_t)std::max(s_up, 1u) * kMoverDoubles, ljnccl::kFloat64, slab.up, slab.comm, st));
    NCK(A.Recv(slab.recvbuf[1], (size_t)std::max(from_up, 1u) * kMoverDoubles, ljnccl::kFloat64, slab.up, slab.comm, st));
    NCK(A.Recv(slab.recvbuf[0], (size_t)std::max(from_down, 1u) * kMoverDoubles, ljnccl::kFloat64, slab.down, slab.comm, st));
    NCK(A.GroupEnd());
    if ((size_t)n_stay + from_up + from_down > cap_atoms) {                       // arrivals that do not fit are not unpacked
        slab_note_failure(2);
        from_down = (unsigned)std::min<size_t>(from_down, cap_atoms - n_stay);
        from_up = (unsigned)std::min<size_t>(from_up, cap_atoms - n_stay - from_down);
    }
    // arrivals overwrite the tail that held my leavers (their data has been packed)
    if (from_down) { unpack_movers_kernel<<<cdiv(from_down, kBlock), kBlock, 0, st>>>(n_stay, from_down, slab.recvbuf[0], pos, v[0], v[1], v[2], f[0], f[1], f[2], orig); check_launch("unpack_movers"); }
    if (from_up)   { unpack_movers_kernel<<<cdiv(from_up, kBlock), kBlock, 0, st>>>(n_stay + from_down, from_up, slab.recvbuf[1], pos, v[0], v[1], v[2], f[0], f[1], f[2], orig); check_launch("unpack_movers"); }
    const bool changed = n_down + n_up + from_up + from_down > 0;
    n = n_stay + from_down + from_up; c.n = n;
    if (changed) sort_by_cell();
    cnt[ST_HALO]++;
}

// Slab path, at re-binning: tell the neighbours how my boundary layers are populated (per-cell counts),
// build the cell table of my two halo layers from what they tell me.
void ljgpu_sim::halo_setup() {
    auto& A = ljnccl::api();
    const unsigned nc2 = nc * nc;
    const uint32_t* my_lo = ccount + (size_t)1 * nc2;              // first owned layer
    const uint32_t* my_hi = ccount + (size_t)slab.nzl * nc2;       // last owned layer
    NCK(A.GroupStart());
    NCK(A.Send(my_lo, nc2, ljnccl::kUint32, slab.down, slab.comm, st));
    NCK(A.Send(my_hi, nc2, ljnccl::kUint32, slab.up, slab.comm, st));
    NCK(A.Recv(slab.hcount[1], nc2, ljnccl::kUint32, slab.up, slab.comm, st));     // upper halo = up neighbour's first layer
    NCK(A.Recv(slab.hcount[0], nc2, ljnccl::kUint32, slab.down, slab.comm, st));   // lower halo = down neighbour's last layer
    NCK(A.GroupEnd());
    for (int k = 0; k < 2; ++k) {
        CK(cudaMemcpyAsync(slab.hoff[k], slab.hcount[k], (size_t)nc2 * 4, cudaMemcpyDeviceToDevice, st));
        launches += exclusive_scan_u32(slab.hoff[k], nc2, slab.xchg + 4 + k, scr, st);
        CK((cudaError_t)sort_take_error());
    }
    // my own boundary layer sizes: slots [0, c_lo) and [n - c_hi, n)
    CK(cudaMemcpyAsync(slab.h_xchg + 4, slab.xchg + 4, 8, cudaMemcpyDeviceToHost, st));
    std::vector<uint32_t> lo(nc2), hi(nc2);
    CK(cudaMemcpyAsync(lo.data(), my_lo, (size_t)nc2 * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hi.data(), my_hi, (size_t)nc2 * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    slab.h_lo = slab.h_xchg[4]; slab.h_hi = slab.h_xchg[5];
    slab.c_lo = 0; slab.c_hi = 0;
    for (uint32_t q : lo) slab.c_lo += q;
    for (uint32_t q : hi) slab.c_hi += q;
    bool halo_fits = (size_t)n + slab.h_lo + slab.h_hi <= pos_cap;
    if (const char* e = std::getenv("LJGPU_TEST_FAIL_RANK"))        // fault injection for the tests: this device "runs out of halo space"
        if (std::atoi(e) == slab.rank && rebuilds >= 1) halo_fits = false;
    if (!halo_fits) {     // keep going with empty halo layers; the neighbours are told below not to push into them
        slab_note_failure(3);
        slab.h_lo = slab.h_hi = 0;
        for (int k = 0; k < 2; ++k) { CK(cudaMemsetAsync(slab.hcount[k], 0, (size_t)nc2 * 4, st)); CK(cudaMemsetAsync(slab.hoff[k], 0, (size_t)nc2 * 4, st)); }
    }
    if (slab.p2p) {     // where my boundary layers land in the neighbours' position arrays
        slab.h_xchg[10] = halo_fits ? n : 0xffffffffu; slab.h_xchg[11] = slab.h_lo;      // n = all ones: do not push to me
        CK(cudaMemcpyAsync(slab.xchg + 10, slab.h_xchg + 10, 8, cudaMemcpyHostToDevice, st));
        NCK(A.GroupStart());
        NCK(A.Send(slab.xchg + 10, 2, ljnccl::kUint32, slab.down, slab.comm, st));
        NCK(A.Send(slab.xchg + 10, 2, ljnccl::kUint32, slab.up, slab.comm, st));
        NCK(A.Recv(slab.xchg + 12, 2, ljnccl::kUint32, slab.down, slab.comm, st));
        NCK(A.Recv(slab.xchg + 14, 2, ljnccl::kUint32, slab.up, slab.comm, st));
        NCK(A.GroupEnd());
        CK(cudaMemcpyAsync(slab.h_xchg + 12, slab.xchg + 12, 16, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        slab.nb_n[0] = slab.h_xchg[12]; slab.nb_hlo[0] = slab.h_xchg[13];
        slab.nb_n[1] = slab.h_xchg[14]; slab.nb_hlo[1] = slab.h_xchg[15];
        if (slab.nb_n[0] == 0xffffffffu || slab.nb_n[1] == 0xffffffffu) {      // a neighbour has no room for my boundary layer
            slab_note_failure(3);
            slab.nb_n[0] = slab.nb_n[1] = 0; slab.nb_hlo[0] = slab.nb_hlo[1] = 0;
            slab.c_lo = slab.c_hi = 0;                                          // nothing is pushed until everybody has thrown
        }
    }
    halo_cells_kernel<<<cdiv(nc2, kBlock), kBlock, 0, st>>>(nc2, 0, n, slab.hcount[0], slab.hoff[0], cstart, ccount);
    check_launch("halo_cells");
    halo_cells_kernel<<<cdiv(nc2, kBlock), kBlock, 0, st>>>(nc2, (slab.nzl + 1) * nc2, n + slab.h_lo, slab.hcount[1], slab.hoff[1], cstart, ccount);
    check_launch("halo_cells");
}

// Slab path, every step: my boundary layers are contiguous slot ranges of the cell-sorted positions, so
// they go out as they lie (no packing) and land in the neighbours' halo slots.
void ljgpu_sim::halo_exchange() {
    auto& A = ljnccl::api();
    cudaEvent_t ev = stage_begin();
    if (slab.p2p) {
        // my first layer -> lower neighbour's UPPER halo slots; my last layer -> upper neighbour's LOWER halo slots
        double4* down_dst = slab.peer_pos[0] + slab.nb_n[0] + slab.nb_hlo[0];
        double4* up_dst = slab.peer_pos[1] + slab.nb_n[1];
        const unsigned total = slab.c_lo + slab.c_hi;
        const unsigned blocks = std::max(1u, std::min(cdiv(total, 256), 296u));
        halo_push_kernel<<<blocks, 256, 0, st>>>(ctrl, pos, n, slab.c_lo, slab.c_hi, down_dst, up_dst,
                                                 slab.peer_mail[slab.down], slab.peer_mail[slab.up], slab.push_ticket);
        check_launch("halo_push");
        stage_end(ST_HALO, ev);
        return;
    }
    NCK(A.GroupStart());
    NCK(A.Send(pos, (size_t)std::max(slab.c_lo, 1u) * 4, ljnccl::kFloat64, slab.down, slab.comm, st));
    NCK(A.Send(pos + (n - slab.c_hi), (size_t)std::max(slab.c_hi, 1u) * 4, ljnccl::kFloat64, slab.up, slab.comm, st));
    NCK(A.Recv(pos + n + slab.h_lo, (size_t)std::max(slab.h_hi, 1u) * 4, ljnccl::kFloat64, slab.up, slab.comm, st));
    NCK(A.Recv(pos + n, (size_t)std::max(slab.h_lo, 1u) * 4, ljnccl::kFloat64, slab.down, slab.comm, st));
    NCK(A.GroupEnd());
    stage_end(ST_HALO, ev);
}

void ljgpu_sim::allreduce_red() {
    NCK(ljnccl::api().AllReduce(ctrl->red, ctrl->red, 4, ljnccl::kFloat64, ljnccl::kSum, slab.comm, st));
}

// Slab path, constructor (what = 1: epot) and prediction (what = 2: pred_sum): local sums -> global sums -> publication.
// Over peer memory this is ONE kernel (mailbox all-reduce inside); on the NCCL fallback it is reduce kernel +
// ncclAllReduce + publish kernel.  (The per-step variant is part of kick2_kernel.)
void ljgpu_sim::slab_finalize(int what, unsigned nb_epot, unsigned nb_kick) {
    if (slab.p2p) {
        slab_exchange_kernel<<<1, kFinalizeThreads, 0, st>>>(c, ctrl, what, partial_epot, nb_epot, partial_pred, nb_kick, slab_red());
        check_launch("slab_exchange");
        return;
    }
    slab_reduce_kernel<<<1, kFinalizeThreads, 0, st>>>(c, ctrl, partial_epot, nb_epot, nullptr, partial_pred, nb_kick);
    check_launch("slab_reduce");
    allreduce_red();
    slab_publish_kernel<<<1, 32, 0, st>>>(c, ctrl, hist, what);
    check_launch("slab_publish");
}

// Map the neighbours' position arrays and every rank's mailbox into this process (CUDA IPC over NVLink).
void ljgpu_sim::setup_p2p() {
    auto& A = ljnccl::api();
    if (std::getenv("LJGPU_SLAB_NCCL") || slab.world > kMaxPeers) return;
    CK(cudaMalloc(&slab.mail, sizeof(Mailbox)));
    CK(cudaMemset(slab.mail, 0, sizeof(Mailbox)));
    {   // a late peer is waited for, a dead one is not: ~30 s by default (ranks do unequal host work between collective calls)
        const char* e = std::getenv("LJGPU_PEER_TIMEOUT_S");
        const double secs = e ? std::max(0.1, std::atof(e)) : 30.0;
        const unsigned long long cycles = (unsigned long long)(secs * 2.0e9);
        CK(cudaMemcpy(&slab.mail->timeout_cycles, &cycles, sizeof cycles, cudaMemcpyHostToDevice));
    }
    CK(cudaMalloc(&slab.push_ticket, sizeof(unsigned)));
    CK(cudaMemset(slab.push_ticket, 0, sizeof(unsigned)));
    CK(cudaMalloc(&slab.pub_ticket, sizeof(unsigned)));
    CK(cudaMemset(slab.pub_ticket, 0, sizeof(unsigned)));
    slab.pub_S = cdiv(nglobal, slab.world);
    for (int b = 0; b < 2; ++b) {
        CK(cudaMalloc(&slab.pub_slice[b], (size_t)3 * slab.pub_S * 8));
        CK(cudaMemset(slab.pub_slice[b], 0, (size_t)3 * slab.pub_S * 8));
    }
    struct Handles { cudaIpcMemHandle_t pos, mail, pub[2]; };
    static_assert(sizeof(Handles) == 256, "four 64-byte IPC handles");
    Handles mine;
    CK(cudaIpcGetMemHandle(&mine.pos, pos));
    CK(cudaIpcGetMemHandle(&mine.mail, slab.mail));
    for (int b = 0; b < 2; ++b) CK(cudaIpcGetMemHandle(&mine.pub[b], slab.pub_slice[b]));
    Handles* d_all = nullptr;
    CK(cudaMalloc(&d_all, sizeof(Handles) * (slab.world + 1)));
    CK(cudaMemcpy(d_all + slab.world, &mine, sizeof(Handles), cudaMemcpyHostToDevice));
    NCK(A.AllGather(d_all + slab.world, d_all, sizeof(Handles), ljnccl::kUint8, slab.comm, st));
    CK(cudaStreamSynchronize(st));
    std::vector<Handles> all(slab.world);
    CK(cudaMemcpy(all.data(), d_all, sizeof(Handles) * slab.world, cudaMemcpyDeviceToHost));
    cudaFree(d_all);
    auto open = [&](const cudaIpcMemHandle_t& h) {
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        slab.ipc_opened[slab.n_ipc_opened++] = p;
        return p;
    };
    for (int r = 0; r < slab.world; ++r) {
        slab.peer_mail[r] = r == slab.rank ? slab.mail : static_cast<Mailbox*>(open(all[r].mail));
        for (int b = 0; b < 2; ++b)
            slab.peer_pub[b][r] = r == slab.rank ? slab.pub_slice[b] : static_cast<double*>(open(all[r].pub[b]));
    }
    slab.peer_pos[0] = static_cast<double4*>(open(all[slab.down].pos));
    slab.peer_pos[1] = slab.up == slab.down ? slab.peer_pos[0] : static_cast<double4*>(open(all[slab.up].pos));
    CK(cudaMalloc(&slab.d_peer_mail, sizeof(Mailbox*) * kMaxPeers));
    CK(cudaMemcpy(slab.d_peer_mail, slab.peer_mail, sizeof(Mailbox*) * kMaxPeers, cudaMemcpyHostToDevice));
    slab.p2p = true;
}

// Slab path: one all-reduce (max) of the devices' failure codes; when any device failed, ALL of them throw the same
// LJGPU_ESTATE error here, so nobody is left waiting in a collective for a device that has gone away.
void ljgpu_sim::agree() {
    slab.h_xchg[8] = slab_fail;
    CK(cudaMemcpyAsync(slab.xchg + 8, slab.h_xchg + 8, 4, cudaMemcpyHostToDevice, st));
    NCK(ljnccl::api().AllReduce(slab.xchg + 8, slab.xchg + 8, 1, ljnccl::kUint32, ljnccl::kMax, slab.comm, st));
    CK(cudaMemcpyAsync(slab.h_xchg + 8, slab.xchg + 8, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const unsigned code = slab.h_xchg[8];
    if (!code) return;
    static const char* what[] = {"", "too many atoms crossing a slab face at once", "a slab holds more atoms than its capacity after migration",
                                 "a halo layer does not fit the position buffer", "a brick holds more home atoms than the brick kernel's work items cover",
                                 "a cell neighbourhood does not fit in shared memory (density too inhomogeneous)"};
    throw LjError(LJGPU_ESTATE, std::string("slab path: ") + what[std::min(code, 5u)] + (slab_fail == code ? " (on this device)" : " (reported by another device)"));
}

// Re-binning: the reference's list rebuild (.cpp:441-551) on the device layout: sort by cell, dij = 0
// (.cpp:484-486), tile tables, and on the Verlet path the list itself.
void ljgpu_sim::rebuild() {
    struct InRebuild { bool& f; explicit InRebuild(bool& x) : f(x) { f = true; } ~InRebuild() { f = false; } } guard(in_rebuild);
    cudaEvent_t ev = stage_begin();
    static const bool trace = std::getenv("LJGPU_TRACE_REBUILD") != nullptr;     // developer aid: host wall time per section
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!trace) return;
        static const bool nosync = std::getenv("LJGPU_TRACE_NOSYNC") != nullptr;
        if (!nosync) cudaStreamSynchronize(st);
        auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[ljgpu rebuild r%d] %-12s %8.1f us\n", slab.rank, what, std::chrono::duration<double, std::micro>(t1 - t0).count());
        t0 = t1;
    };
    CK(cudaMemsetAsync(&ctrl->stop, 0, sizeof(unsigned), st));
    CK(cudaMemsetAsync(&ctrl->maxdisp2, 0, sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(&ctrl->inner_valid, 0, sizeof(unsigned), st));       // slots are about to be permuted
    since_prune = ~0u / 2;                                                  // prune at the next step
    sort_by_cell();
    lap("sort");
    if (slab.on) { migrate(); lap("migrate"); }
    if (n) for (int k = 0; k < 3; ++k) CK(cudaMemsetAsync(dij[k], 0, (size_t)n * 8, st));
    if (slab.on) { halo_setup(); lap("halo_setup"); halo_exchange(); lap("halo_xchg"); }

    CK(cudaMemsetAsync(&ctrl->max_tile, 0, sizeof(unsigned), st));
    CK(cudaMemsetAsync(&ctrl->max_home, 0, sizeof(unsigned), st));
    tile_table_kernel<<<cdiv(nblocks_tile, 4), 128, 0, st>>>(tile_grid(), cstart, ccount, tsrc, toff, ctrl);
    check_launch("tile_table");
    fetch_ctrl();
    // launch shapes only grow (a few spare shared-memory records cost nothing; a changed shape costs a graph capture)
    tcap = std::max(tcap, ((h_ctrl->max_tile + 32 + 63) / 64) * 64);     // (32 spare records: the list build reads whole rounds of candidates)
    tile_ipb = std::max(tile_ipb, (h_ctrl->max_home + 31) / 32);
    if (tile_ipb > (unsigned)kBrickItemsMax) {
        if (!slab.on) throw LjError(LJGPU_ESTATE, "a brick holds more home atoms than the brick kernel's work items cover");
        slab_note_failure(4); tile_ipb = kBrickItemsMax;
    }
    nb_force = nblocks_tile * tile_ipb;
    // as many tiles in flight per CTA as fit beside the kernel's static shared memory (the ring never needs more than four)
    tile_slots = (unsigned)std::min<size_t>(kBrickSlotsMax, (size_t)(220 * 1024) / ((size_t)tcap * kTileRecBytes));
    if (const char* e = std::getenv("LJGPU_BRICK_SLOTS")) tile_slots = std::max(2u, std::min(tile_slots, (unsigned)std::atoi(e)));   // tuning knob
    if (tile_slots < 2 || h_ctrl->max_tile > 65535) {
        if (!slab.on) throw LjError(LJGPU_ESTATE, "a cell neighbourhood does not fit in shared memory (density too inhomogeneous)");
        slab_note_failure(5);
    }
    if (slab.on) agree();                   // every device throws here, or none does
    set_tile_attributes(this);

    lap("tile_table");
    if (mode == LJGPU_KERNEL_VERLET) build_list();
    lap("build_list");
    stage_end(ST_REBUILD, ev);
    ++rebuilds;
    if (since_rebuild > 0) last_interval = since_rebuild;
    since_rebuild = 0;
    dirty_since_rebuild = false;
}

// The persistent brick kernel (lj_tile.cuh): as many CTAs as stay resident (occupancy x SMs), each streaming bricks
// through its ring of two to four shared-memory tiles.
static size_t tile_smem(const ljgpu_sim* s) { return (size_t)s->tcap * kTileRecBytes * s->tile_slots; }

template <int MODE, bool UNIT_SIGMA>
static void launch_tile2(ljgpu_sim* s, int guarded) {
    TileArgs a{s->pos, s->nl, s->nl_count, s->prune_skin > 0.0 ? s->nl_in : nullptr, s->nl_in_count,
               s->f[0], s->f[1], s->f[2], s->partial_epot, s->ctrl, s->capq, s->tcap, s->tile_slots, s->tile_ipb, 0u,
               s->slab.p2p ? s->slab.mail : nullptr};
    const unsigned grid = std::max(1u, std::min(s->nblocks_tile, s->tile_grid_ctas[MODE][UNIT_SIGMA ? 1 : 0]));
    // how early the staging warps fetch the next brick (in work items still unclaimed): with many bricks per CTA a
    // brick's worth hides the staging latency; with few, bricks go only to CTAs whose warps are already waiting
    static const int la_env = std::getenv("LJGPU_BRICK_LOOKAHEAD") ? std::atoi(std::getenv("LJGPU_BRICK_LOOKAHEAD")) : -1;   // tuning knob: eighths of a brick
    // (measured at N = 2^17 ... 2^20, profiles/r03a_*, r03h_*: a full brick from ~12 bricks per CTA on, half a brick from ~4, none below)
    const unsigned eighths = la_env >= 0 ? (unsigned)la_env : (s->nblocks_tile >= 12u * grid ? 8u : (s->nblocks_tile >= 4u * grid ? 4u : 0u));
    a.lookahead = s->tile_ipb * eighths / 8u;
    tile_kernel<MODE, UNIT_SIGMA><<<grid, kBrickThreads, tile_smem(s), s->st>>>(s->c, s->tile_grid(), a, guarded);
    s->check_launch("tile_kernel");
}
// Dynamic shared memory beyond 48 KB is an opt-in, and the grid is sized from the occupancy the driver reports for
// this instantiation.  Done at re-binning time (never inside a stream capture).
template <int MODE, bool US>
static void tile_prepare(ljgpu_sim* s) {
    const size_t smem = tile_smem(s);
    CK(cudaFuncSetAttribute(tile_kernel<MODE, US>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tile_kernel<MODE, US>, kBrickThreads, smem));
    if (per_sm < 1) throw LjError(LJGPU_ESTATE, "the brick kernel does not fit on an SM with this tile size");
    static const int cap = std::getenv("LJGPU_BRICK_CTAS_PER_SM") ? std::atoi(std::getenv("LJGPU_BRICK_CTAS_PER_SM")) : 0;   // tuning knob
    if (cap > 0) per_sm = std::min(per_sm, cap);
    s->tile_grid_ctas[MODE][US ? 1 : 0] = (unsigned)per_sm * (unsigned)s->num_sms;
}
static void set_tile_attributes(ljgpu_sim* s) {
    // seven instantiations x (attribute + occupancy query) cost ~70 us of host time: only when the tile shape changed
    if (s->tile_attr_tcap == s->tcap && s->tile_attr_slots == s->tile_slots) return;
    s->tile_attr_tcap = s->tcap; s->tile_attr_slots = s->tile_slots;
    tile_prepare<TILE_FORCE_LIST, true>(s); tile_prepare<TILE_FORCE_LIST, false>(s);
    tile_prepare<TILE_FORCE_PRUNE, true>(s); tile_prepare<TILE_FORCE_PRUNE, false>(s);
    tile_prepare<TILE_FORCE_CELL, false>(s); tile_prepare<TILE_LIST_COUNT, false>(s); tile_prepare<TILE_LIST_FILL, false>(s);
}

template <int MODE>
static void launch_tile(ljgpu_sim* s, int guarded) {
    if ((MODE == TILE_FORCE_LIST || MODE == TILE_FORCE_PRUNE) && s->c.sigma2 == 1.0) launch_tile2<MODE, true>(s, guarded);
    else launch_tile2<MODE, false>(s, guarded);
}

void ljgpu_sim::build_list() {
    const unsigned ntiles = cdiv(cap_atoms, 32);
    static const bool trace = std::getenv("LJGPU_TRACE_REBUILD") != nullptr;
    auto note = [&](const char* what) {
        if (trace) std::fprintf(stderr, "[ljgpu build_list r%d] %s n=%u halo=%u+%u tcap=%u capq=%u blocks=%u\n",
                                slab.rank, what, n, slab.h_lo, slab.h_hi, tcap, capq, nblocks_tile);
    };
    note("enter");
    for (int attempt = 0; attempt < 4; ++attempt) {
        CK(cudaMemsetAsync(&ctrl->max_count, 0, sizeof(unsigned), st));
        CK(cudaMemsetAsync(&ctrl->overflow, 0, sizeof(unsigned), st));
        if (capq == 0) {
            launch_tile<TILE_LIST_COUNT>(this, 0);
            fetch_ctrl();
            note("count kernel done");
            unsigned mx = h_ctrl->max_count;
            if (slab.on) {          // every device must use the same capacity decision sequence (same NCCL call order)
                slab.h_xchg[8] = mx;
                CK(cudaMemcpyAsync(slab.xchg + 8, slab.h_xchg + 8, 4, cudaMemcpyHostToDevice, st));
                NCK(ljnccl::api().AllReduce(slab.xchg + 8, slab.xchg + 8, 1, ljnccl::kUint32, ljnccl::kMax, slab.comm, st));
                CK(cudaMemcpyAsync(slab.h_xchg + 8, slab.xchg + 8, 4, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                mx = slab.h_xchg[8];
            }
            capq = (mx + 12 + 7) / 8;                             // quads of 8 entries: max + slack
            note("capacity agreed");
        }
        const size_t need = (size_t)ntiles * capq * 32 * 4;
        if (need > nl_alloc) {      // two spare quads per atom: the longest list creeps up by an entry now and then,
            const size_t want = (size_t)ntiles * (capq + 2) * 32 * 4;   // and freeing + allocating 2 x 250 MB took up to 170 ms
            dfree(nl); dfree(nl_in);
            CK(cudaMalloc(&nl, want * 4));
            if (mode == LJGPU_KERNEL_VERLET) CK(cudaMalloc(&nl_in, want * 4));
            nl_alloc = want;
            ++graph_gen;            // (cached graphs hold the old pointers; the key would catch it too)
        }
        launch_tile<TILE_LIST_FILL>(this, 0);
        fetch_ctrl();
        note("fill kernel done");
        unsigned over = h_ctrl->overflow, mx = h_ctrl->max_count;
        if (slab.on) {
            slab.h_xchg[8] = mx; slab.h_xchg[9] = over;
            CK(cudaMemcpyAsync(slab.xchg + 8, slab.h_xchg + 8, 8, cudaMemcpyHostToDevice, st));
            NCK(ljnccl::api().AllReduce(slab.xchg + 8, slab.xchg + 8, 2, ljnccl::kUint32, ljnccl::kMax, slab.comm, st));
            CK(cudaMemcpyAsync(slab.h_xchg + 8, slab.xchg + 8, 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            mx = slab.h_xchg[8]; over = slab.h_xchg[9];
        }
        if (!over) {
            return;
        }
        capq = (mx + 12 + 7) / 8;                                 // grow and refill
    }
    throw LjError(LJGPU_ESTATE, "neighbour list capacity did not converge");
}

void ljgpu_sim::launch_force(bool timed) {
    cudaEvent_t ev = timed ? stage_begin() : nullptr;
    if (mode == LJGPU_KERNEL_ALLPAIRS) {
        dim3 grid(nb_atoms(), nsplit);
        force_allpairs_kernel<<<grid, kBlock, 0, st>>>(c, ctrl, pos, jchunk, apf[0], apf[1], apf[2], partial_epot);
        check_launch("force_allpairs");
        allpairs_reduce_kernel<<<nb_atoms(), kBlock, 0, st>>>(c, ctrl, nsplit, apf[0], apf[1], apf[2], f[0], f[1], f[2]);
        check_launch("allpairs_reduce");
    } else if (mode == LJGPU_KERNEL_CELL) {
        launch_tile<TILE_FORCE_CELL>(this, 1);
    } else {
        launch_tile<TILE_FORCE_LIST>(this, 1);
    }
    if (timed) stage_end(ST_FORCE, ev);
}

void ljgpu_sim::finalize_epot_only() {
    if (slab.on) {
        slab_finalize(1, nb_force, 0);
    } else {
        init_finalize_kernel<<<1, kFinalizeThreads, 0, st>>>(c, ctrl, partial_epot, nb_force);
        check_launch("init_finalize");
    }
}

// Constructor tail, .cpp:57-62: list, forces, publish; ekin stays 0.
void ljgpu_sim::initial_forces() {
    if (list_mode()) rebuild();
    launch_force(false);
    finalize_epot_only();
    sync();
}

void ljgpu_sim::launch_predict(double dt) {
    const double a = 0.5 * dt / p.mass;                                  // .cpp:339
    cudaEvent_t ev = stage_begin();
    predict_kernel<<<nb_atoms(), kBlock, 0, st>>>(n, a, v[0], v[1], v[2], f[0], f[1], f[2], partial_pred);
    check_launch("predict");
    if (slab.on) {
        slab_finalize(2, 0, nb_atoms());
    } else {
        predict_finalize_kernel<<<1, kFinalizeThreads, 0, st>>>(ctrl, partial_pred, nb_atoms());
        check_launch("predict_finalize");
    }
    stage_end(ST_PREDICT, ev);
    pred_valid = true; pred_dt = dt;
}

// The kernels of one integrate() call (no host decisions inside: capturable into a CUDA graph).
void ljgpu_sim::enqueue_step(bool prune_now, double dt) {
    const double a = 0.5 * dt / p.mass;                                  // .cpp:339
    c.dt_over_tau = dt / p.tau;                                          // .cpp:375
    c.dt = dt;
    cudaEvent_t ev = stage_begin();
    if (list_mode())
        kick_drift_kernel<true><<<nb_atoms(), kBlock, 0, st>>>(c, dt, a, ctrl, pos, dij[0], dij[1], dij[2],
                                                               v[0], v[1], v[2], f[0], f[1], f[2]);
    else
        kick_drift_kernel<false><<<nb_atoms(), kBlock, 0, st>>>(c, dt, a, ctrl, pos, nullptr, nullptr, nullptr,
                                                                v[0], v[1], v[2], f[0], f[1], f[2]);
    check_launch("kick_drift");
    stage_end(ST_KICK_DRIFT, ev);
    if (slab.on) halo_exchange();
    if (prune_now) {       // this step's force pass walks the outer list and emits the inner one
        cudaEvent_t pv = stage_begin();
        launch_tile<TILE_FORCE_PRUNE>(this, 1);
        prune_commit_kernel<<<1, 1, 0, st>>>(ctrl);
        check_launch("prune_commit");
        stage_end(ST_FORCE_PRUNE, pv);
    } else {
        launch_force(true);
    }
    ev = stage_begin();
    // at most one resident wave, and every thread the same number of rounds (two atoms per round): no ragged last round
    const unsigned pairs2 = std::max(1u, n / 2), wave2 = (unsigned)num_sms * 2048u;
    const unsigned nb2 = cdiv(pairs2, cdiv(pairs2, wave2) * kKick2Block);
    if (!slab.on) {
        kick2_kernel<0><<<nb2, kKick2Block, 0, st>>>(c, a, ctrl, hist, v[0], v[1], v[2], f[0], f[1], f[2], partial_epot, nb_force, kick_sums, SlabRed{});
        check_launch("kick2");
    } else if (slab.p2p) {
        kick2_kernel<1><<<nb2, kKick2Block, 0, st>>>(c, a, ctrl, hist, v[0], v[1], v[2], f[0], f[1], f[2], partial_epot, nb_force, kick_sums, slab_red());
        check_launch("kick2");
    } else {
        kick2_kernel<2><<<nb2, kKick2Block, 0, st>>>(c, a, ctrl, hist, v[0], v[1], v[2], f[0], f[1], f[2], partial_epot, nb_force, kick_sums, SlabRed{});
        check_launch("kick2");
        allreduce_red();
        slab_publish_kernel<<<1, 32, 0, st>>>(c, ctrl, hist, 0);
        check_launch("slab_publish");
    }
    stage_end(ST_KICK2, ev);
}

ljgpu_sim::GraphKey ljgpu_sim::graph_key(bool prune_now, double dt) const {
    GraphKey k;
    std::memset(&k, 0, sizeof k);                      // padding included: keys are compared bytewise
    const void* ptrs[20] = {pos, v[0], v[1], v[2], f[0], f[1], f[2], dij[0], dij[1], dij[2],
                            nl, nl_in, nl_count, nl_in_count, tsrc, toff,
                            slab.p2p ? (const void*)(slab.peer_pos[0] + slab.nb_n[0] + slab.nb_hlo[0]) : nullptr,     // where my halo pushes land
                            slab.p2p ? (const void*)(slab.peer_pos[1] + slab.nb_n[1]) : nullptr, nullptr, nullptr};
    std::memcpy(k.ptr, ptrs, sizeof ptrs);
    k.shape[0] = n; k.shape[1] = tile_grid_ctas[TILE_FORCE_LIST][1]; k.shape[2] = tcap; k.shape[3] = capq; k.shape[4] = nblocks_tile; k.shape[5] = nb_force; k.shape[6] = tile_slots; k.shape[7] = tile_ipb;
    k.shape[8] = slab.c_lo; k.shape[9] = slab.c_hi;
    k.dt = dt; k.gen = graph_gen; k.prune = prune_now ? 1 : 0;
    return k;
}

void ljgpu_sim::drop_graphs() {
    for (auto& g : graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    graphs.clear();
    ++graph_gen;
}

// One integrate() call: decides on the host whether this step also prunes, then replays the matching
// CUDA graph (single device, not profiling) or launches the kernels one by one.
void ljgpu_sim::launch_step(double dt) {
    bool prune_now = false;
    if (prune_skin > 0.0) {
        prune_now = since_prune >= prune_interval;
        if (prune_now) since_prune = 0;
        ++since_prune;
    }
    static const bool no_graph = std::getenv("LJGPU_NO_GRAPH") != nullptr;
    bool timed_step = false;
    if (prof) { timed_step = (prof_tick++ % prof_every) == 0; }
    if (timed_step || (slab.on && !slab.p2p) || no_graph) {      // (the NCCL step loop is not captured)
        HostTrace::Scope hs(htrace, HostTrace::ENQUEUE);
        prof_this_step = timed_step || !prof;          // slab / no-graph steps that are not sampled carry no event records
        enqueue_step(prune_now, dt);
        prof_this_step = true;
        return;
    }
    prof_this_step = false;          // graph replay (and its capture) carries no event records
    const GraphKey key = graph_key(prune_now, dt);
    StepGraph* hit = nullptr;
    for (auto& e : graphs) if (e.key == key) { hit = &e; break; }
    cudaGraphExec_t recycle = nullptr;                   // an executable graph of the same shape whose parameters can be updated in place
    if (!hit) {
        if (graphs.size() >= 12) {                    // evict the least recently used
            size_t lru = graphs.size();
            for (size_t k = 0; k < graphs.size(); ++k)      // prefer one of the same kind (plain / pruning): same topology
                if (graphs[k].key.prune == key.prune && (lru == graphs.size() || graphs[k].used < graphs[lru].used)) lru = k;
            if (lru == graphs.size()) {
                lru = 0;
                for (size_t k = 1; k < graphs.size(); ++k) if (graphs[k].used < graphs[lru].used) lru = k;
                cudaGraphExecDestroy(graphs[lru].exec);
            } else {
                recycle = graphs[lru].exec;
            }
            graphs.erase(graphs.begin() + (long)lru);
        } else if (slab.on) {                          // slabs: every re-binning changes counts baked into the launches - keep two graphs, update them
            for (size_t k = 0; k < graphs.size(); ++k)
                if (graphs[k].key.prune == key.prune) { recycle = graphs[k].exec; graphs.erase(graphs.begin() + (long)k); break; }
        }
        graphs.emplace_back();
        hit = &graphs.back();
        hit->key = key;
    }
    StepGraph& g = *hit;
    g.used = ++graph_clock;
    if (!g.exec) {
        HostTrace::Scope hs(htrace, HostTrace::CAPTURE);
        uint64_t cnt0[ST_COUNT]; std::memcpy(cnt0, cnt, sizeof cnt0);
        const uint64_t l0 = launches;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        cudaGraph_t graph = nullptr;
        try { enqueue_step(prune_now, dt); }
        catch (...) { cudaStreamEndCapture(st, &graph); if (graph) cudaGraphDestroy(graph); if (recycle) cudaGraphExecDestroy(recycle); throw; }
        CK(cudaStreamEndCapture(st, &graph));
        if (recycle) {                                 // same kernels, new arguments: far cheaper than a fresh instantiation
            cudaGraphExecUpdateResultInfo info;
            if (cudaGraphExecUpdate(recycle, graph, &info) == cudaSuccess) g.exec = recycle;
            else { cudaGetLastError(); cudaGraphExecDestroy(recycle); }
        }
        if (!g.exec) CK(cudaGraphInstantiate(&g.exec, graph, 0));
        cudaGraphDestroy(graph);
        g.kernels = launches - l0;
        for (int k = 0; k < ST_COUNT; ++k) { g.cnt[k] = cnt[k] - cnt0[k]; cnt[k] = cnt0[k]; }
        launches = l0;
    }
    {
        HostTrace::Scope hs(htrace, HostTrace::GRAPH_LAUNCH);
        CK(cudaGraphLaunch(g.exec, st));
    }
    launches += g.kernels;
    for (int k = 0; k < ST_COUNT; ++k) cnt[k] += g.cnt[k];
    prof_this_step = true;
}

// n integrate() calls.  Steps are enqueued in batches without host round trips; once a step finds a
// displacement beyond shell/2 (.cpp:557-568) the device raises `stop`, the rest of the batch turns
// into no-ops, and the host re-bins (.cpp:89-91) before continuing.  On the slab path the flag is part
// of the per-step all-reduce, so every device stops at the same step and issues the same NCCL calls.
void ljgpu_sim::run_steps(unsigned first, unsigned nsteps, double dt) {
    unsigned step = first, remaining = nsteps;
    energies_fresh = false;
    while (remaining) {
        if (!pred_valid || pred_dt != dt) launch_predict(dt);
        unsigned batch;
        if (!list_mode()) batch = std::min(remaining, 512u);
        else if (since_rebuild < last_interval) batch = std::max(1u, ((last_interval - since_rebuild) * 3u) / 4u);
        else batch = 2;
        batch = std::min(std::min(batch, remaining), 256u);
        CK(cudaMemsetAsync(&ctrl->steps_done, 0, sizeof(unsigned), st));
        h_scalar[0] = step;                              // integrate(step, dt): the device counts on from here
        CK(cudaMemcpyAsync(&ctrl->next_step, h_scalar, sizeof(unsigned), cudaMemcpyHostToDevice, st));
        for (unsigned k = 0; k < batch; ++k) launch_step(dt);
        {
            HostTrace::Scope hs(htrace, HostTrace::FETCH);
            fetch_ctrl();
        }
        const unsigned done = h_ctrl->steps_done;
        if (prune_skin > 0.0 && h_ctrl->vmax_last * std::fabs(dt) > 0.0)     // steps until the fastest atom has covered s_in/2, with margin
            prune_interval = (unsigned)std::max(1.0, std::min(1.0e6, 0.85 * c.prune_half / (h_ctrl->vmax_last * std::fabs(dt))));
        step += done; remaining -= done; steps += done; since_rebuild += done;
        if (done) dirty_since_rebuild = true;
        if (h_ctrl->stop) { HostTrace::Scope hs(htrace, HostTrace::REBUILD); rebuild(); }
        else if (done != batch) throw LjError(LJGPU_ESTATE, "step batch ended early without a rebuild request");
        // between two batches the state is complete: let a waiting query in (not on slabs - every call there is collective,
        // a query from another thread would have to be made on every rank at the same point anyway)
        if (remaining && !slab.on && held && waiting.load(std::memory_order_acquire) > 0) {
            held->unlock();
            while (waiting.load(std::memory_order_acquire) > 0) std::this_thread::yield();
            held->lock();
            CK(cudaSetDevice(device));
        }
    }
    energies_fresh = nsteps > 0;        // (a re-binning after the last fetch does not touch the published energies)
}

// Neighbour-count parity probe (include/ljgpu.h): bins a copy of the published (wrapped) positions of the WHOLE
// system with the reference's cell assignment and counts with the reference's uncontracted comparisons.  The
// integrator's own state is only read.  On slabs every device first assembles the whole system (collective).
void ljgpu_sim::probe_neighbor_counts(double cut2, int inclusive, uint32_t* host_out) {
    const size_t N = nglobal;
    const unsigned nbN = std::max(1u, cdiv(N, kBlock));
    if (!list_mode()) {
        count_allpairs_kernel<<<nb_atoms(), kBlock, 0, st>>>(c, pos, orig, cut2, inclusive, out_u32);
        check_launch("neighbor_count");
        CK(cudaMemcpyAsync(host_out, out_u32, N * 4, cudaMemcpyDeviceToHost, st));
        sync();
        return;
    }
    const unsigned ncell_g = nc * nc * nc;
    if (!probe_pos) {
        CK(cudaMalloc(&probe_pos, N * sizeof(double4))); CK(cudaMalloc(&probe_sorted, N * sizeof(double4)));
        CK(cudaMalloc(&probe_keys, N * 4)); CK(cudaMalloc(&probe_vals, N * 4));
        CK(cudaMalloc(&probe_cstart, (size_t)ncell_g * 4)); CK(cudaMalloc(&probe_cend, (size_t)ncell_g * 4)); CK(cudaMalloc(&probe_ccount, (size_t)ncell_g * 4));
    }
    // published positions in original order (slab: summed over the devices, every atom is owned by exactly one)
    if (slab.on) CK(cudaMemsetAsync(out3, 0, 3 * N * 8, st));
    if (n) {
        export_positions_kernel<<<nb_atoms(), kBlock, 0, st>>>(c, pos, orig, out3, out3 + N, out3 + 2 * N);
        check_launch("export");
    }
    if (slab.on) NCK(ljnccl::api().AllReduce(out3, out3, 3 * N, ljnccl::kFloat64, ljnccl::kSum, slab.comm, st));
    LJConst cg = c;                     // the whole box, one device's view
    cg.n = (unsigned)N; cg.slab = 0; cg.zlo = 0; cg.nzl = nc;
    probe_key_kernel<<<nbN, kBlock, 0, st>>>(cg, out3, out3 + N, out3 + 2 * N, probe_pos, probe_keys, probe_vals);
    check_launch("probe_key");
    int bits = 1; while ((1ull << bits) < (unsigned long long)ncell_g) ++bits;
    launches += radix_sort_pairs(probe_keys, probe_vals, N, bits, scr, st);
    CK((cudaError_t)sort_take_error());
    CK(cudaMemsetAsync(probe_cstart, 0, (size_t)ncell_g * 4, st));
    CK(cudaMemsetAsync(probe_cend, 0, (size_t)ncell_g * 4, st));
    cell_bounds_kernel<<<nbN, kBlock, 0, st>>>((unsigned)N, probe_keys, probe_cstart, probe_cend);
    check_launch("cell_bounds");
    cell_count_kernel<<<cdiv(ncell_g, kBlock), kBlock, 0, st>>>(ncell_g, probe_cstart, probe_cend, probe_ccount);
    check_launch("cell_count");
    gather_pos_kernel<<<nbN, kBlock, 0, st>>>((unsigned)N, probe_vals, probe_pos, probe_sorted);
    check_launch("gather_pos");
    CellGrid g{probe_cstart, probe_ccount, nc};
    count_cells_kernel<<<cdiv(ncell_g, kCellWarps), kCellWarps * 32, 0, st>>>(cg, g, probe_sorted, cut2, inclusive, out_u32, probe_vals);
    check_launch("neighbor_count");
    CK(cudaMemcpyAsync(host_out, out_u32, N * 4, cudaMemcpyDeviceToHost, st));
    sync();
}

void ljgpu_sim::export3(int what, double* a, double* b, double* cc) {
    const size_t N = nglobal;
    if (slab.on) CK(cudaMemsetAsync(out3, 0, 3 * N * 8, st));
    if (n) {
        if (what == 0) {
            export_positions_kernel<<<nb_atoms(), kBlock, 0, st>>>(c, pos, orig, out3, out3 + N, out3 + 2 * N);
        } else {
            double** src = what == 1 ? v : f;
            export_soa_kernel<<<nb_atoms(), kBlock, 0, st>>>(n, src[0], src[1], src[2], orig, out3, out3 + N, out3 + 2 * N);
        }
        check_launch("export");
    }
    if (slab.on)     // every device contributes its atoms at their original indices; the sum is the whole system
        NCK(ljnccl::api().AllReduce(out3, out3, 3 * N, ljnccl::kFloat64, ljnccl::kSum, slab.comm, st));
    CK(cudaMemcpyAsync(a, out3, N * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(b, out3 + N, N * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(cc, out3 + 2 * N, N * 8, cudaMemcpyDeviceToHost, st));
    sync();
}

// integrate() + publish_state() (.cpp:70-98, :574-576) without stalling the integrator on the transfer: the step
// runs (and is waited for: the energies are final on return), its positions are gathered into original order on
// the integrator's stream, and the device->host copies go out on a second stream while the caller's next step
// computes.  On return the PREVIOUS delivery is complete; the current one is in flight.
void ljgpu_sim::step_publish(unsigned step, double dt, double* x, double* y, double* z) {
    run_steps(step, 1, dt);
    if (slab.on && !slab.p2p) { export3(0, x, y, z); return; }      // NCCL step loop: every rank assembles the whole system, synchronous
    if (slab.on) {
        // Peer-memory slabs: every device scatters the atoms it owns into the frame slices of their owners (NVLink), each
        // owner copies ITS slice - original indices [rank * S, rank * S + len) of x, y and z - to the host over its own
        // PCIe link, on the copy stream, while the next step computes.  x, y, z address the WHOLE frame (N doubles each);
        // for one consumer to see all of it they must be memory shared by the ranks' processes.
        if (!st_copy) {
            CK(cudaStreamCreateWithFlags(&st_copy, cudaStreamNonBlocking));
            for (int b = 0; b < 2; ++b) {
                CK(cudaEventCreateWithFlags(&ev_pub_ready[b], cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&ev_pub_done[b], cudaEventDisableTiming));
            }
        }
        const int b = pub_cur ^ 1;
        ++slab.pub_seq;
        PubTargets tgt{};
        for (int r = 0; r < slab.world; ++r) { tgt.slice[r] = slab.peer_pub[b][r]; tgt.mail[r] = slab.peer_mail[r]; }
        const unsigned blocks = std::max(1u, std::min(cdiv(n, 256), 4u * (unsigned)num_sms));
        slab_publish_scatter_kernel<<<blocks, 256, 0, st>>>(c, pos, orig, slab.pub_S, nglobal, tgt, slab.rank, slab.world, slab.pub_seq, slab.pub_ticket);
        check_launch("slab_publish_scatter");
        CK(cudaEventRecord(ev_pub_ready[b], st));
        CK(cudaStreamWaitEvent(st_copy, ev_pub_ready[b], 0));
        slab_publish_wait_kernel<<<1, 32, 0, st_copy>>>(slab.mail, slab.world, slab.pub_seq);
        check_launch("slab_publish_wait");
        const size_t first = (size_t)slab.rank * slab.pub_S;
        const size_t len = first < nglobal ? std::min<size_t>(slab.pub_S, nglobal - first) : 0;
        if (len) {
            const double* o = slab.pub_slice[b];
            CK(cudaMemcpyAsync(x + first, o, len * 8, cudaMemcpyDeviceToHost, st_copy));
            CK(cudaMemcpyAsync(y + first, o + len, len * 8, cudaMemcpyDeviceToHost, st_copy));
            CK(cudaMemcpyAsync(z + first, o + 2 * len, len * 8, cudaMemcpyDeviceToHost, st_copy));
        }
        CK(cudaEventRecord(ev_pub_done[b], st_copy));
        if (pub_pending) CK(cudaEventSynchronize(ev_pub_done[pub_cur]));
        pub_cur = b; pub_pending = true; pub_step = step;
        return;
    }
    const size_t N = nglobal;
    if (!st_copy) {
        CK(cudaStreamCreateWithFlags(&st_copy, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            CK(cudaMalloc(&pub3[b], 3 * N * 8));
            CK(cudaEventCreateWithFlags(&ev_pub_ready[b], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ev_pub_done[b], cudaEventDisableTiming));
        }
    }
    const int b = pub_cur ^ 1;                               // its last delivery (two calls ago) was waited for one call ago
    double* o = pub3[b];
    export_positions_kernel<<<nb_atoms(), kBlock, 0, st>>>(c, pos, orig, o, o + N, o + 2 * N);
    check_launch("export");
    CK(cudaEventRecord(ev_pub_ready[b], st));
    CK(cudaStreamWaitEvent(st_copy, ev_pub_ready[b], 0));
    if (y == x + N && z == y + N) {                          // one block on the host too (the adapter's mirror): one transfer
        CK(cudaMemcpyAsync(x, o, 3 * N * 8, cudaMemcpyDeviceToHost, st_copy));
    } else {
        CK(cudaMemcpyAsync(x, o, N * 8, cudaMemcpyDeviceToHost, st_copy));
        CK(cudaMemcpyAsync(y, o + N, N * 8, cudaMemcpyDeviceToHost, st_copy));
        CK(cudaMemcpyAsync(z, o + 2 * N, N * 8, cudaMemcpyDeviceToHost, st_copy));
    }
    CK(cudaEventRecord(ev_pub_done[b], st_copy));
    if (pub_pending) CK(cudaEventSynchronize(ev_pub_done[pub_cur]));
    pub_cur = b; pub_pending = true; pub_step = step;
}

void ljgpu_sim::publish_wait() {
    if (!pub_pending) return;
    CK(cudaEventSynchronize(ev_pub_done[pub_cur]));
    pub_pending = false;
}

// =================================================================================================
// C ABI
// =================================================================================================
namespace {

int fail(int code, const std::string& msg) { g_err = msg; return code; }

template <class F>
int guarded(F&& fn) {
    try { fn(); g_err.clear(); return LJGPU_OK; }
    catch (const LjError& e) { return fail(e.code, e.what()); }
    catch (const std::exception& e) { return fail(LJGPU_ECUDA, e.what()); }
}

template <class F>
int with_sim(ljgpu_sim* s, F&& fn, bool collective_state = false) {
    if (!s) return fail(LJGPU_EINVAL, "null simulation handle");
    const int rc = guarded([&] {
        s->waiting.fetch_add(1, std::memory_order_acq_rel);
        std::unique_lock<std::mutex> lk(s->mu);
        s->waiting.fetch_sub(1, std::memory_order_acq_rel);
        struct Held { ljgpu_sim* s; ~Held() { s->held = nullptr; } } h{s};
        s->held = &lk;
        CK(cudaSetDevice(s->device));
        fn();
    });
    // (an LJGPU_ESTATE refusal of a query leaves the handle usable; a failed step or re-binning does not)
    if (rc == LJGPU_ECUDA || rc == LJGPU_ENCCL || (rc == LJGPU_ESTATE && collective_state)) s->failed = true;
    return rc;
}

void validate(const ljgpu_params& p, int& mode) {
    auto bad = [](const std::string& m) { throw LjError(LJGPU_EINVAL, m); };
    if (p.nr_particles < 2) bad("nr_particles must be at least 2");
    if (!(p.cell_length > 0) || !(p.mass > 0) || !(p.sigma > 0) || !(p.rcut > 0) || !(p.shell >= 0) || !(p.tau != 0))
        bad("cell_length, mass, sigma, rcut must be positive, shell non-negative, tau non-zero");
    const unsigned ncells = std::max(1u, (unsigned)(p.cell_length / (p.rcut + p.shell)));
    if (ncells < 3)
        bad("cell_length / (rcut + shell) < 3: the reference's 27-cell stencil revisits cells there and "
            "double-counts pairs (lennardjonessimulation.cpp:496-516); such boxes are not supported");
    mode = p.force_kernel;
    if (mode == LJGPU_KERNEL_AUTO) mode = p.nr_particles <= 8192 ? LJGPU_KERNEL_ALLPAIRS : LJGPU_KERNEL_VERLET;
    if (mode < LJGPU_KERNEL_ALLPAIRS || mode > LJGPU_KERNEL_VERLET) bad("unknown force_kernel");
    if (mode == LJGPU_KERNEL_ALLPAIRS && p.nr_particles > (1 << 20)) bad("all-pairs path is limited to 2^20 atoms");
}

} // namespace

extern "C" {

const char* ljgpu_last_error(void) { return g_err.c_str(); }

int ljgpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static int create_common(const ljgpu_params* p, int rank, int world, const void* nccl_id,
                         const double* x, const double* y, const double* z,
                         const double* vx, const double* vy, const double* vz, ljgpu_sim** out) {
    if (!p || !x || !y || !z || !vx || !vy || !vz || !out) return fail(LJGPU_EINVAL, "null argument");
    *out = nullptr;
    ljgpu_sim* s = nullptr;
    int rc = guarded([&] {
        int mode = 0;
        validate(*p, mode);
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            throw LjError(LJGPU_ECUDA, "no CUDA device available (libljgpu has no CPU fallback)");
        }
        int dev = p->device;
        if (dev < 0) CK(cudaGetDevice(&dev));
        if (dev >= ndev) throw LjError(LJGPU_EINVAL, "device ordinal out of range");
        CK(cudaSetDevice(dev));
        s = new ljgpu_sim;
        s->p = *p; s->mode = mode; s->device = dev;
        if (world > 1) {
            if (mode == LJGPU_KERNEL_ALLPAIRS) { mode = LJGPU_KERNEL_VERLET; s->mode = mode; }   // slabs need the cell structure
            const unsigned ncells = std::max(1u, (unsigned)(p->cell_length / (p->rcut + p->shell)));
            if ((unsigned)world > ncells) throw LjError(LJGPU_EINVAL, "more devices than cell layers along z");
            if (!nccl_id) throw LjError(LJGPU_EINVAL, "slab path needs the NCCL unique id of rank 0");
            auto& A = ljnccl::api();
            if (!A.ok) throw LjError(LJGPU_ENCCL, A.error);
            s->slab.on = true; s->slab.rank = rank; s->slab.world = world;
            s->slab.down = (rank + world - 1) % world; s->slab.up = (rank + 1) % world;
            uint32_t lo = 0, hi = 0;
            ljgpu_slab_layers(ncells, world, rank, &lo, &hi);
            s->slab.zlo = lo; s->slab.nzl = hi - lo;
            ljnccl::UniqueId id;
            std::memcpy(id.internal, nccl_id, sizeof id.internal);
            NCK(A.CommInitRank(&s->slab.comm, world, id, rank));
        }
        CK(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));
        CK(cudaDeviceGetAttribute(&s->num_sms, cudaDevAttrMultiProcessorCount, dev));
        s->n = 0;
        s->setup_constants();
        s->allocate();
        if (s->slab.on) { s->setup_p2p(); s->setup_constants(); }      // pruning is decided once the step loop is known
        s->upload_state(x, y, z, vx, vy, vz);
        s->initial_forces();
    });
    if (rc != LJGPU_OK) { if (s) s->failed = true; delete s; return rc; }
    *out = s;
    return LJGPU_OK;
}

int ljgpu_create(const ljgpu_params* p, const double* x, const double* y, const double* z,
                 const double* vx, const double* vy, const double* vz, ljgpu_sim** out) {
    return create_common(p, 0, 1, nullptr, x, y, z, vx, vy, vz, out);
}

// ---- slab path ------------------------------------------------------------------------------------
int ljgpu_slab_layers(uint32_t cells_per_axis, int world, int rank, uint32_t* first_layer, uint32_t* end_layer) {
    if (world < 1 || rank < 0 || rank >= world || !first_layer || !end_layer) return LJGPU_EINVAL;
    const uint32_t base = cells_per_axis / (uint32_t)world, extra = cells_per_axis % (uint32_t)world;
    const uint32_t r = (uint32_t)rank;
    *first_layer = r * base + std::min(r, extra);
    *end_layer = *first_layer + base + (r < extra ? 1u : 0u);
    return LJGPU_OK;
}

int ljgpu_slab_unique_id(void* id128) {
    if (!id128) return fail(LJGPU_EINVAL, "null argument");
    return guarded([&] {
        auto& A = ljnccl::api();
        if (!A.ok) throw LjError(LJGPU_ENCCL, A.error);
        ljnccl::UniqueId id;
        NCK(A.GetUniqueId(&id));
        std::memcpy(id128, id.internal, sizeof id.internal);
    });
}

int ljgpu_create_slab(const ljgpu_params* p, int rank, int world, const void* nccl_unique_id,
                      const double* x, const double* y, const double* z,
                      const double* vx, const double* vy, const double* vz, ljgpu_sim** out) {
    if (world < 2 || rank < 0 || rank >= world) return fail(LJGPU_EINVAL, "slab path needs world >= 2 and 0 <= rank < world");
    return create_common(p, rank, world, nccl_unique_id, x, y, z, vx, vy, vz, out);
}

int ljgpu_slab_info(ljgpu_sim* s, int32_t* rank, int32_t* world, uint32_t* first_layer, uint32_t* n_layers,
                    uint64_t* atoms_owned, uint64_t* atoms_halo) {
    return with_sim(s, [&] {
        if (rank) *rank = s->slab.rank;
        if (world) *world = s->slab.world;
        if (first_layer) *first_layer = s->slab.on ? s->slab.zlo : 0;
        if (n_layers) *n_layers = s->slab.on ? s->slab.nzl : s->nc;
        if (atoms_owned) *atoms_owned = s->n;
        if (atoms_halo) *atoms_halo = s->slab.on ? (uint64_t)s->slab.h_lo + s->slab.h_hi : 0;
    });
}

void ljgpu_destroy(ljgpu_sim* s) { delete s; }

int ljgpu_set_state(ljgpu_sim* s, const double* x, const double* y, const double* z,
                    const double* vx, const double* vy, const double* vz) {
    if (!x || !y || !z || !vx || !vy || !vz) return fail(LJGPU_EINVAL, "null argument");
    return with_sim(s, [&] { s->upload_state(x, y, z, vx, vy, vz); s->initial_forces(); }, true);
}

int ljgpu_update_params(ljgpu_sim* s, const ljgpu_params* p) {
    if (!p) return fail(LJGPU_EINVAL, "null argument");
    return with_sim(s, [&] {
        if (p->cell_length != s->p.cell_length || p->nr_particles != s->p.nr_particles ||
            p->rcut != s->p.rcut || p->shell != s->p.shell)
            throw LjError(LJGPU_EINVAL, "cell_length, nr_particles, rcut and shell cannot change on a live simulation");
        if (!(p->mass > 0) || !(p->sigma > 0) || !(p->tau != 0)) throw LjError(LJGPU_EINVAL, "bad parameter value");
        const bool force_changed = p->sigma != s->p.sigma || p->epsilon != s->p.epsilon;
        const int keep_kernel = s->p.force_kernel, keep_dev = s->p.device;
        s->p = *p; s->p.force_kernel = keep_kernel; s->p.device = keep_dev;
        const double dto = s->c.dt_over_tau;
        s->setup_constants();
        s->c.dt_over_tau = dto;
        s->pred_valid = false;                         // mass enters the kick factor
        s->drop_graphs();                              // constants are baked into the captured launches
        if (force_changed) {                           // forces of the current positions under the new potential
            if (s->slab.on) s->halo_exchange();
            s->launch_force(false);
            s->finalize_epot_only();
            s->sync();
        }
    });
}

int ljgpu_step(ljgpu_sim* s, uint32_t step, double dt) {
    return with_sim(s, [&] { s->run_steps(step, 1, dt); }, true);
}

int ljgpu_step_publish(ljgpu_sim* s, uint32_t step, double dt, double* x, double* y, double* z) {
    if (!x || !y || !z) return fail(LJGPU_EINVAL, "null argument");
    return with_sim(s, [&] { s->step_publish(step, dt, x, y, z); }, true);
}

int ljgpu_publish_wait(ljgpu_sim* s) {
    return with_sim(s, [&] { s->publish_wait(); });
}

int ljgpu_step_n(ljgpu_sim* s, uint32_t first, uint32_t n, double dt) {
    return with_sim(s, [&] { s->run_steps(first, n, dt); }, true);
}

int ljgpu_step_n_timed(ljgpu_sim* s, uint32_t first, uint32_t n, double dt, double* device_ms) {
    if (!device_ms) return fail(LJGPU_EINVAL, "null argument");
    return with_sim(s, [&] {
        struct Events {
            cudaEvent_t a = nullptr, b = nullptr;
            ~Events() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }      // also when run_steps throws
        } ev;
        CK(cudaEventCreate(&ev.a)); CK(cudaEventCreate(&ev.b));
        CK(cudaEventRecord(ev.a, s->st));
        s->run_steps(first, n, dt);
        CK(cudaEventRecord(ev.b, s->st));
        CK(cudaEventSynchronize(ev.b));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ev.a, ev.b));
        *device_ms = ms;
    }, true);
}

int ljgpu_get_positions(ljgpu_sim* s, double* x, double* y, double* z) {
    if (!x || !y || !z) return fail(LJGPU_EINVAL, "null argument");
    return with_sim(s, [&] { s->export3(0, x, y, z); });
}
int ljgpu_get_velocities(ljgpu_sim* s, double* x, double* y, double* z) {
    if (!x || !y || !z) return fail(LJGPU_EINVAL, "null argument");
    return with_sim(s, [&] { s->export3(1, x, y, z); });
}
int ljgpu_get_forces(ljgpu_sim* s, double* x, double* y, double* z) {
    if (!x || !y || !z) return fail(LJGPU_EINVAL, "null argument");
    return with_sim(s, [&] { s->export3(2, x, y, z); });
}

int ljgpu_get_energies(ljgpu_sim* s, double* ekin, double* epot, double* etot) {
    return with_sim(s, [&] {
        if (!s->energies_fresh) s->fetch_ctrl();
        if (ekin) *ekin = s->h_ctrl->ekin;
        if (epot) *epot = s->h_ctrl->epot;
        if (etot) *etot = s->h_ctrl->etot;
    });
}

int ljgpu_get_energy_history(ljgpu_sim* s, uint32_t max_entries, uint32_t* step_index,
                             double* ekin, double* epot, uint32_t* n_out) {
    if (!n_out) return fail(LJGPU_EINVAL, "null argument");
    return with_sim(s, [&] {
        s->fetch_ctrl();
        std::vector<HistEntry> h(kHistCap);
        CK(cudaMemcpyAsync(h.data(), s->hist, sizeof(HistEntry) * kHistCap, cudaMemcpyDeviceToHost, s->st));
        s->sync();
        const unsigned long long head = s->h_ctrl->hist_head;
        const unsigned long long avail = std::min<unsigned long long>(head, kHistCap);
        const unsigned long long take = std::min<unsigned long long>(avail, max_entries);
        for (unsigned long long k = 0; k < take; ++k) {
            const HistEntry& e = h[(head - take + k) % kHistCap];
            if (step_index) step_index[k] = e.step;
            if (ekin) ekin[k] = e.ekin;
            if (epot) epot[k] = e.epot;
        }
        *n_out = (uint32_t)take;
    });
}

int ljgpu_get_speed_histogram(ljgpu_sim* s, uint32_t nbins, double vmax, uint64_t* counts) {
    if (!counts) return fail(LJGPU_EINVAL, "null argument");
    if (nbins == 0 || nbins > (uint32_t)kMaxBins || !(vmax > 0)) return fail(LJGPU_EINVAL, "nbins must be 1..1024 and vmax positive");
    return with_sim(s, [&] {
        CK(cudaMemsetAsync(s->hist_counts, 0, nbins * sizeof(unsigned long long), s->st));
        const unsigned blocks = std::max(1u, std::min(cdiv(s->n, 256), 148u * 8u));
        speed_histogram_kernel<<<blocks, 256, 0, s->st>>>(s->n, s->v[0], s->v[1], s->v[2], nbins, vmax, s->hist_counts);
        s->check_launch("speed_histogram");
        if (s->slab.on)
            NCK(ljnccl::api().AllReduce(s->hist_counts, s->hist_counts, nbins, ljnccl::kUint64, ljnccl::kSum, s->slab.comm, s->st));
        static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "");
        CK(cudaMemcpyAsync(counts, s->hist_counts, nbins * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->st));
        s->sync();
    });
}

int ljgpu_set_sample_stride(ljgpu_sim* s, uint32_t stride) {
    return with_sim(s, [&] {
        s->sync();
        CK(cudaMemcpyAsync(&s->ctrl->sample_stride, &stride, sizeof(uint32_t), cudaMemcpyHostToDevice, s->st));
        CK(cudaMemsetAsync(&s->ctrl->sample_head, 0, sizeof(unsigned long long), s->st));     // a new series
        s->sync();
    });
}

int ljgpu_get_graph_stats(ljgpu_sim* s, int reference_truncation, ljgpu_graph_stats* out) {
    if (!out) return fail(LJGPU_EINVAL, "null argument");
    static_assert(sizeof(ljgpu_graph_stats) == sizeof(SeriesStatsOut), "same layout on both sides of the ABI");
    return with_sim(s, [&] {
        series_stats_kernel<<<1, 32, 0, s->st>>>(s->ctrl, reference_truncation, s->d_stats);
        s->check_launch("series_stats");
        CK(cudaMemcpyAsync(out, s->d_stats, sizeof(SeriesStatsOut), cudaMemcpyDeviceToHost, s->st));
        s->sync();
    });
}

int ljgpu_get_speed_statistics(ljgpu_sim* s, uint32_t nbins, double vmax, uint64_t* counts, double* expected,
                               double* chi2, uint32_t* bins_used) {
    if (!counts || !expected || !chi2 || !bins_used) return fail(LJGPU_EINVAL, "null argument");
    if (nbins == 0 || nbins > (uint32_t)kMaxBins || !(vmax > 0)) return fail(LJGPU_EINVAL, "nbins must be 1..1024 and vmax positive");
    return with_sim(s, [&] {
        CK(cudaMemsetAsync(s->hist_counts, 0, nbins * sizeof(unsigned long long), s->st));
        const unsigned blocks = std::max(1u, std::min(cdiv(s->n, 256), 148u * 8u));
        speed_histogram_kernel<<<blocks, 256, 0, s->st>>>(s->n, s->v[0], s->v[1], s->v[2], nbins, vmax, s->hist_counts);
        s->check_launch("speed_histogram");
        if (s->slab.on)
            NCK(ljnccl::api().AllReduce(s->hist_counts, s->hist_counts, nbins, ljnccl::kUint64, ljnccl::kSum, s->slab.comm, s->st));
        speed_mb_kernel<<<1, 256, 0, s->st>>>(nbins, vmax, s->p.kT, s->p.mass, (double)s->nglobal, s->hist_counts, s->d_mb + 2, s->d_mb);
        s->check_launch("speed_mb");
        double res[2] = {0.0, 0.0};
        CK(cudaMemcpyAsync(counts, s->hist_counts, nbins * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->st));
        CK(cudaMemcpyAsync(expected, s->d_mb + 2, nbins * sizeof(double), cudaMemcpyDeviceToHost, s->st));
        CK(cudaMemcpyAsync(res, s->d_mb, 2 * sizeof(double), cudaMemcpyDeviceToHost, s->st));
        s->sync();
        *chi2 = res[0]; *bins_used = (uint32_t)res[1];
    });
}

int ljgpu_get_cell_ids(ljgpu_sim* s, uint32_t* cell_of_atom, uint32_t dims[3]) {
    if (!cell_of_atom) return fail(LJGPU_EINVAL, "null argument");
    return with_sim(s, [&] {
        if (s->slab.on) CK(cudaMemsetAsync(s->out_u32, 0, (size_t)s->nglobal * 4, s->st));
        cell_id_probe_kernel<<<s->nb_atoms(), kBlock, 0, s->st>>>(s->c, s->pos, s->orig, s->out_u32);
        s->check_launch("cell_id_probe");
        if (s->slab.on)
            NCK(ljnccl::api().AllReduce(s->out_u32, s->out_u32, s->nglobal, ljnccl::kUint32, ljnccl::kSum, s->slab.comm, s->st));
        CK(cudaMemcpyAsync(cell_of_atom, s->out_u32, (size_t)s->nglobal * 4, cudaMemcpyDeviceToHost, s->st));
        s->sync();
        if (dims) dims[0] = dims[1] = dims[2] = s->nc;
    });
}

int ljgpu_get_neighbor_counts(ljgpu_sim* s, double cutoff, int inclusive, uint32_t* count_of_atom) {
    if (!count_of_atom) return fail(LJGPU_EINVAL, "null argument");
    return with_sim(s, [&] {
        if (!(cutoff > 0) || cutoff > s->p.rcut + s->p.shell)
            throw LjError(LJGPU_EINVAL, "cutoff must be in (0, rcut + shell]");
        s->probe_neighbor_counts(cutoff * cutoff, inclusive, count_of_atom);
    });
}

int ljgpu_get_list_counts(ljgpu_sim* s, int which, uint32_t* count_of_atom) {
    if (!count_of_atom) return fail(LJGPU_EINVAL, "null argument");
    return with_sim(s, [&] {
        if (s->mode != LJGPU_KERNEL_VERLET) throw LjError(LJGPU_ESTATE, "only the Verlet path keeps a neighbour list");
        if (which != 0 && which != 1) throw LjError(LJGPU_EINVAL, "which: 0 = the list of the last re-binning, 1 = the pruned inner list");
        const uint32_t* src = which == 0 ? s->nl_count : s->nl_in_count;
        if (which == 1) {
            s->fetch_ctrl();
            if (!s->nl_in || !s->h_ctrl->inner_valid) throw LjError(LJGPU_ESTATE, "no pruned inner list exists at the moment");
        }
        const size_t N = s->nglobal;
        if (s->slab.on) CK(cudaMemsetAsync(s->out_u32, 0, N * 4, s->st));
        if (s->n) {
            scatter_u32_kernel<<<s->nb_atoms(), kBlock, 0, s->st>>>(s->n, src, s->orig, s->out_u32);
            s->check_launch("scatter_u32");
        }
        if (s->slab.on)
            NCK(ljnccl::api().AllReduce(s->out_u32, s->out_u32, N, ljnccl::kUint32, ljnccl::kSum, s->slab.comm, s->st));
        CK(cudaMemcpyAsync(count_of_atom, s->out_u32, N * 4, cudaMemcpyDeviceToHost, s->st));
        s->sync();
    });
}

int ljgpu_write_movie_frame(ljgpu_sim* s, const char* path, int create) {
    if (!path) return fail(LJGPU_EINVAL, "null argument");
    return with_sim(s, [&] {
        const size_t N = s->nglobal;
        std::vector<double> buf(6 * N);
        s->export3(0, &buf[0], &buf[N], &buf[2 * N]);
        s->export3(1, &buf[3 * N], &buf[4 * N], &buf[5 * N]);
        if (s->slab.on && s->slab.rank != 0) return;             // every device holds the frame; rank 0 writes it
        FILE* fp = std::fopen(path, create ? "wb" : "ab");                 // .cpp:131-138
        if (!fp) throw LjError(LJGPU_EINVAL, std::string("cannot open ") + path);
        if (create) { const uint32_t n32 = (uint32_t)N; std::fwrite(&n32, sizeof n32, 1, fp); }
        const double dims[3] = {s->p.cell_length, s->p.cell_length, s->p.cell_length};
        std::fwrite(dims, sizeof(double), 3, fp);                           // .cpp:141-143
        std::vector<double> rec(7 * N);
        for (size_t i = 0; i < N; ++i) {                                    // .cpp:145-154
            double* r = &rec[7 * i];
            r[0] = buf[i]; r[1] = buf[N + i]; r[2] = buf[2 * N + i];
            r[3] = buf[3 * N + i]; r[4] = buf[4 * N + i]; r[5] = buf[5 * N + i];
            r[6] = std::sqrt(r[3] * r[3] + r[4] * r[4] + r[5] * r[5]);
        }
        std::fwrite(rec.data(), sizeof(double), rec.size(), fp);
        std::fclose(fp);
    });
}

int ljgpu_set_profiling(ljgpu_sim* s, int enabled) {
    // enabled == 1: every step is timed stage by stage (plain launches); enabled == k > 1: every k-th step is,
    // the others replay the CUDA graph, and the sampled stage times are scaled by k
    return with_sim(s, [&] { s->sync(); s->prof = enabled != 0; s->prof_every = enabled > 1 ? (unsigned)enabled : 1u; s->prof_tick = 0; });
}

int ljgpu_get_timers(ljgpu_sim* s, ljgpu_timers* out) {
    if (!out) return fail(LJGPU_EINVAL, "null argument");
    return with_sim(s, [&] {
        s->sync();
        out->ms_predict = s->ms[ST_PREDICT]; out->ms_kick_drift = s->ms[ST_KICK_DRIFT]; out->ms_force_prune = s->ms[ST_FORCE_PRUNE];
        out->ms_force = s->ms[ST_FORCE]; out->ms_kick2 = s->ms[ST_KICK2]; out->ms_rebuild = s->ms[ST_REBUILD]; out->ms_halo = s->ms[ST_HALO];
        out->n_predict = s->cnt[ST_PREDICT]; out->n_kick_drift = s->cnt[ST_KICK_DRIFT]; out->n_force_prune = s->cnt[ST_FORCE_PRUNE];
        out->n_force = s->cnt[ST_FORCE]; out->n_kick2 = s->cnt[ST_KICK2]; out->n_rebuild = s->cnt[ST_REBUILD]; out->n_halo = s->cnt[ST_HALO];
        out->steps = s->steps; out->rebuilds = s->rebuilds; out->kernel_launches = s->launches;
    });
}

int ljgpu_reset_timers(ljgpu_sim* s) {
    return with_sim(s, [&] {
        s->sync();
        for (int k = 0; k < ST_COUNT; ++k) { s->ms[k] = 0; s->cnt[k] = 0; }
        s->steps = 0; s->rebuilds = 0; s->launches = 0;
    });
}

int ljgpu_get_layout(ljgpu_sim* s, int32_t* force_kernel, uint32_t* cells_per_axis,
                     uint32_t* list_capacity, uint64_t* inner_list_steps) {
    return with_sim(s, [&] {
        if (force_kernel) *force_kernel = s->mode;
        if (cells_per_axis) *cells_per_axis = s->nc;
        if (list_capacity) *list_capacity = 8 * s->capq;
        if (inner_list_steps) { s->fetch_ctrl(); *inner_list_steps = s->h_ctrl->inner_used; }
    });
}

int ljgpu_get_list_stats(ljgpu_sim* s, double* mean_length, uint32_t* max_length) {
    return with_sim(s, [&] {
        double mean = 0.0; uint32_t mx = 0;
        if (s->mode == LJGPU_KERNEL_VERLET && s->nl_count) {
            std::vector<uint32_t> h(s->n);
            CK(cudaMemcpyAsync(h.data(), s->nl_count, (size_t)s->n * 4, cudaMemcpyDeviceToHost, s->st));
            s->sync();
            unsigned long long tot = 0;
            for (uint32_t c : h) { tot += c; mx = std::max(mx, c); }
            mean = (double)tot / (double)s->n;
        }
        if (mean_length) *mean_length = mean;
        if (max_length) *max_length = mx;
    });
}

int ljgpu_measure_fp64_peak(int device, double* tflops) {
    if (!tflops) return fail(LJGPU_EINVAL, "null argument");
    return guarded([&] {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); throw LjError(LJGPU_ECUDA, "no CUDA device available"); }
        if (device < 0) CK(cudaGetDevice(&device));
        if (device >= ndev) throw LjError(LJGPU_EINVAL, "device ordinal out of range");
        CK(cudaSetDevice(device));
        int sms = 0;
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        const int blocks = sms * 8, threads = 256, iters = 8192;       // 64 warps per SM, 8 chains each
        double* out = nullptr;
        CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        struct Guard { double*& o; cudaEvent_t &a, &b; ~Guard() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); cudaFree(o); } } guard{out, e0, e1};
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        float best = 1e30f;
        for (int r = 0; r < 6; ++r) {                                   // first pass warms up
            CK(cudaEventRecord(e0, nullptr));
            fp64_peak_kernel<<<blocks, threads, 0, nullptr>>>(out, iters, 1.0000001, 1e-9);
            CK(cudaGetLastError());
            CK(cudaEventRecord(e1, nullptr));
            CK(cudaEventSynchronize(e1));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r > 0 && ms < best) best = ms;
        }
        *tflops = 2.0 * 8.0 * (double)iters * threads * blocks / ((double)best * 1e-3) / 1e12;
    });
}

int ljgpu_register_host(void* p, size_t bytes) {
    if (!p) return fail(LJGPU_EINVAL, "null argument");
    return guarded([&] { CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable)); });
}

int ljgpu_unregister_host(void* p) {
    return guarded([&] { if (p) CK(cudaHostUnregister(p)); });
}

int ljgpu_alloc_host(size_t bytes, void** out) {
    if (!out) return fail(LJGPU_EINVAL, "null argument");
    return guarded([&] { CK(cudaMallocHost(out, bytes ? bytes : 1)); });
}

int ljgpu_free_host(void* p) {
    return guarded([&] { if (p) CK(cudaFreeHost(p)); });
}

} // extern "C"

#ifdef LJ_BRICK_TRACE
// developer build only: the brick kernel's timeline of its last TILE_FORCE_LIST launch, as raw TraceEvent records
extern "C" int ljgpu_debug_trace_dump(const char* path) {
    std::vector<ljgpu::TraceEvent> ev((size_t)160 * ljgpu::kTraceCap);
    unsigned n[160];
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(ev.data(), ljgpu::g_trace, ev.size() * sizeof(ljgpu::TraceEvent)) != cudaSuccess) return -2;
    if (cudaMemcpyFromSymbol(n, ljgpu::g_trace_n, sizeof n) != cudaSuccess) return -3;
    FILE* f = std::fopen(path, "wb");
    if (!f) return -4;
    std::fwrite(n, sizeof n, 1, f);
    std::fwrite(ev.data(), sizeof(ljgpu::TraceEvent), ev.size(), f);
    std::fclose(f);
    return 0;
}
#endif
