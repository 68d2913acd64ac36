/* include/ljgpu.h - C ABI of libljgpu.so, the B200 (sm_100a) Lennard-Jones integrator.
 *
 * This is the drop-in boundary for ONE path of ifilot/lennard-jones-live-simulator: the per-step
 * integrator behind class LennardJonesSimulation (src/simulator/lennardjonessimulation.h:113-155).
 * Each entry point names the reference interface it replaces (file:line under the reference's
 * src/simulator/).  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - every function returns an int status: LJGPU_OK (0) or a negative LJGPU_E* code; the message
 *     for the calling thread's last failure is ljgpu_last_error();
 *   - every array argument is a caller-owned HOST buffer of nr_particles elements, structure of
 *     arrays, in ORIGINAL ATOM ORDER (the order of the arrays given to ljgpu_create), whatever
 *     order the device keeps internally;
 *   - all arithmetic is fp64; indices are uint32_t;
 *   - one handle may be stepped by one thread while other threads call getters: calls on one
 *     handle are serialised internally, a getter sees the state after the last completed step;
 *   - there is no CPU fallback: without a usable CUDA device every call fails with LJGPU_ECUDA.
 */
#ifndef LJGPU_H
#define LJGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LJGPU_VERSION_MAJOR 0
#define LJGPU_VERSION_MINOR 1

enum {
    LJGPU_OK       =  0,
    LJGPU_EINVAL   = -1,   /* bad argument / unsupported parameter combination */
    LJGPU_ECUDA    = -2,   /* CUDA runtime failure (no device, launch error, out of memory ...) */
    LJGPU_ENCCL    = -3,   /* NCCL failure on the slab path */
    LJGPU_ESTATE   = -4    /* call not valid in the handle's current state */
};

/* Which force path integrates the system.  AUTO picks ALLPAIRS for nr_particles <= 8192 or when
 * the box holds fewer than 3 cells per axis, VERLET otherwise. */
enum {
    LJGPU_KERNEL_AUTO     = 0,
    LJGPU_KERNEL_ALLPAIRS = 1,  /* shared-memory tiled O(N^2), minimum image, atoms stay in caller order */
    LJGPU_KERNEL_CELL     = 2,  /* cell-sorted atoms, 27-cell stencil evaluated every step            */
    LJGPU_KERNEL_VERLET   = 3   /* cell-sorted atoms + device Verlet list (skin = shell), the fast path */
};

/* The ten keys of LennardJonesParameters (lennardjonesparameters.h:32-69) as MainWindow fills
 * them (src/gui/mainwindow.cpp:67-77), read ONCE here instead of by string on every stage
 * (lennardjonessimulation.cpp:259-266 etc.), plus the device-side choices. */
typedef struct ljgpu_params {
    double  cell_length;     /* cubic box edge L                                   */
    int32_t nr_particles;    /* N                                                  */
    int32_t force_kernel;    /* LJGPU_KERNEL_*                                     */
    double  kT;
    double  mass;
    double  sigma;
    double  rcut;
    double  epsilon;
    double  tau;             /* Berendsen coupling time (always on, .cpp:80,368-383) */
    double  shell;           /* Verlet skin                                        */
    double  stepsize;        /* informational: dt is passed to every step call     */
    int32_t device;          /* CUDA device ordinal; -1 = the calling thread's current device */
    int32_t reserved;
} ljgpu_params;

typedef struct ljgpu_sim ljgpu_sim;   /* opaque */

/* Wall/device time spent per stage since creation or the last ljgpu_reset_timers(); filled only
 * while profiling is enabled (ljgpu_set_profiling), measured with CUDA events on the
 * integrator's own stream. Milliseconds and launch counts. */
typedef struct ljgpu_timers {
    /* ms_force_prune / n_force_prune: force passes that also emitted the pruned inner list */
    double   ms_predict, ms_kick_drift, ms_force_prune, ms_force, ms_kick2, ms_rebuild, ms_halo;
    uint64_t n_predict, n_kick_drift, n_force_prune, n_force, n_kick2, n_rebuild, n_halo;
    uint64_t steps;            /* integrate steps completed                         */
    uint64_t rebuilds;         /* cell re-binnings (+ list builds) so far           */
    uint64_t kernel_launches;  /* every kernel this handle has launched             */
} ljgpu_timers;

/* ---- library -------------------------------------------------------------------------- */

/* Message of the calling thread's most recent failing call ("" if none). */
const char* ljgpu_last_error(void);

/* Number of CUDA devices visible (0 if none / driver missing). Never fails. */
int ljgpu_device_count(void);

/* ---- initial state (host) ---------------------------------------------------------------- */

/* LennardJonesSimulation::initialize() + get_gauss() (lennardjonessimulation.cpp:185-252):
 * simple-cubic start lattice and Gaussian velocities drawn from a process-wide, never-reseeded
 * std::default_random_engine + std::normal_distribution<double>(0,1), mean velocity removed.
 * Runs on the host with the same std:: calls as the reference, so a fresh process reproduces
 * the reference's start state bit for bit. */
int ljgpu_host_initialize(const ljgpu_params* p, double* x, double* y, double* z,
                          double* vx, double* vy, double* vz);

/* Reseed that process-wide generator to its default-constructed state (test convenience; the
 * reference has no such call). */
void ljgpu_host_rng_reset(void);

/* ---- lifecycle --------------------------------------------------------------------------- */

/* Constructor, lennardjonessimulation.cpp:31-63: uploads the state, bins the atoms, evaluates
 * forces and epot; ekin reads 0.0 until the first step, as in the reference. */
int ljgpu_create(const ljgpu_params* p, const double* x, const double* y, const double* z,
                 const double* vx, const double* vy, const double* vz, ljgpu_sim** out);

void ljgpu_destroy(ljgpu_sim* s);

/* Replace positions and velocities (original order) and redo what the constructor does. */
int ljgpu_set_state(ljgpu_sim* s, const double* x, const double* y, const double* z,
                    const double* vx, const double* vy, const double* vz);

/* The reference re-reads its parameter map on every stage, so a caller may change kT, tau, mass,
 * sigma or epsilon between steps (LennardJonesParameters::set_param, lennardjonesparameters.h:66).
 * This call applies such a change; cell_length, nr_particles, rcut and shell must be unchanged. */
int ljgpu_update_params(ljgpu_sim* s, const ljgpu_params* p);

/* ---- stepping ---------------------------------------------------------------------------- */

/* LennardJonesSimulation::integrate(step, dt), lennardjonessimulation.cpp:70-114.  Returns when
 * the step is complete and published. */
int ljgpu_step(ljgpu_sim* s, uint32_t step, double dt);

/* n consecutive integrate() calls (step indices first .. first+n-1) without returning to the
 * caller in between: the headless loop of ThreadIntegrate::run (threadintegrate.cpp:38-66)
 * minus its 2000 steps/s throttle. */
int ljgpu_step_n(ljgpu_sim* s, uint32_t first, uint32_t n, double dt);

/* ljgpu_step_n, additionally reporting the device time of the n steps (re-binning included) measured
 * with CUDA events recorded on the integrator's own stream around the whole batch. */
int ljgpu_step_n_timed(ljgpu_sim* s, uint32_t first, uint32_t n, double dt, double* device_ms);

/* integrate(step, dt) followed by publish_state() (lennardjonessimulation.cpp:70-98 and :574-576: every
 * reference step ends by making the new positions the renderer's buffer).  The step is complete and its
 * energies are final when the call returns; the positions of this step (original atom order, wrapped like
 * ljgpu_get_positions) are on their way into x, y, z - the device->host transfer runs on a second stream while
 * the caller's next step computes.  When the call returns, the delivery started by the PREVIOUS
 * ljgpu_step_publish is complete; the one started by this call is complete after the next ljgpu_step_publish
 * or ljgpu_publish_wait.  Alternate between two sets of buffers (pinned: ljgpu_alloc_host), exactly like the
 * reference's double buffer.
 * On slabs (peer-memory step loop) the call is collective and every device delivers only ITS slice of the frame:
 * original indices [rank * S, min(N, (rank + 1) * S)), S = ceil(N / world), of x, y and z - every device first scatters
 * the atoms it owns into the slices of their owners over NVLink, then each owner copies its slice out over its own
 * PCIe link.  x, y, z address the whole frame on every rank; pass memory shared by the ranks' processes (POSIX shared
 * memory pinned with ljgpu_register_host) and one consumer sees the whole system, moved once instead of `world` times. */
int ljgpu_step_publish(ljgpu_sim* s, uint32_t step, double dt, double* x, double* y, double* z);

/* Blocks until the delivery started by the last ljgpu_step_publish is complete (no-op if none is pending). */
int ljgpu_publish_wait(ljgpu_sim* s);

/* ---- queries (original atom order) --------------------------------------------------------- */

/* get_positions_copy / get_render_x,y,z (lennardjonessimulation.h:131-143, .cpp:159-171):
 * published positions, wrapped into [0, L] by the reference's own expression (.cpp:431-433). */
int ljgpu_get_positions(ljgpu_sim* s, double* x, double* y, double* z);
/* get_velocities_copy (.cpp:173-180). */
int ljgpu_get_velocities(ljgpu_sim* s, double* vx, double* vy, double* vz);
/* The reference's private fx,fy,fz (lennardjonessimulation.h:91) - parity probe. */
int ljgpu_get_forces(ljgpu_sim* s, double* fx, double* fy, double* fz);
/* get_ekin / get_epot / get_etot (lennardjonessimulation.h:146-148). */
int ljgpu_get_energies(ljgpu_sim* s, double* ekin, double* epot, double* etot);
/* Energies of the most recent steps, newest last: fills at most max_entries triples and returns
 * how many in *n_out.  step_index[] are the values passed as `step`. */
int ljgpu_get_energy_history(ljgpu_sim* s, uint32_t max_entries, uint32_t* step_index,
                             double* ekin, double* epot, uint32_t* n_out);

/* GraphWidget::set_particle_speed (graphwidget.cpp:205-233): counts[b] = number of atoms with
 * b == min((int)(|v| / vmax * nbins), nbins-1); the GUI uses nbins = 100, vmax = 10.0
 * (graphwidget.h:73-74). */
int ljgpu_get_speed_histogram(ljgpu_sim* s, uint32_t nbins, double vmax, uint64_t* counts);
/* GraphWidget::set_params (graphwidget.cpp:55-73): Maxwell-Boltzmann expectation per bin. Host. */
int ljgpu_maxwell_boltzmann(double kT, double mass, int32_t nr_particles, uint32_t nbins,
                            double vmax, double* expected);

/* GraphWidget statistics on the device (src/simulator/graphwidget.cpp:139-203, :235-248).  The integrator keeps the series
 * the widget plots: the energies of every `stride`-th step (ThreadIntegrate::run emits them every 1000 iterations,
 * threadintegrate.cpp:57-62; default stride 1000, ring of the last 8192 samples; setting the stride starts a new series).
 * ljgpu_get_graph_stats: minimum, maximum (the chart's axis range), mean and standard deviation (the average line and the
 * +-sigma band; variance over n - 1, 0 when not finite) of the kinetic [0], potential [1] and total [2] energy series.
 * reference_truncation != 0 reproduces the reference's std::accumulate(..., 0) - an int accumulator that truncates every
 * partial sum (:147, :169, :191, :236); 0 sums in double. */
typedef struct ljgpu_graph_stats {
    double   mean[3], stddev[3], min[3], max[3];
    uint32_t n_samples, reserved;
} ljgpu_graph_stats;
int ljgpu_set_sample_stride(ljgpu_sim* s, uint32_t stride);
int ljgpu_get_graph_stats(ljgpu_sim* s, int reference_truncation, ljgpu_graph_stats* out);
/* Speed histogram (as ljgpu_get_speed_histogram) together with the Maxwell-Boltzmann overlay the widget draws over it
 * (graphwidget.cpp:55-73: f(v) N binsize at the left bin edges, at the thermostat's kT) and their chi-square over the
 * bins_used bins that expect more than five atoms - all evaluated on the device. */
int ljgpu_get_speed_statistics(ljgpu_sim* s, uint32_t nbins, double vmax, uint64_t* counts, double* expected,
                               double* chi2, uint32_t* bins_used);

/* Parity probes of build_neighbor_list (lennardjonessimulation.cpp:441-551) on the published
 * positions, bit-exact with the reference's comparisons:
 *   cell_of_atom[i] = iz*nx*ny + iy*nx + ix with ix = min((unsigned)(x/dx), nx-1)   (:447-476)
 *   count_of_atom[i] = #{ j != i : min-image dist2 <= cutoff^2 }  when inclusive != 0  (:518-540)
 *                    = #{ j != i : min-image dist2 <  cutoff^2 }  when inclusive == 0  (:291-305)
 * cutoff must not exceed rcut + shell.  Both probes only read the simulation: the neighbour count bins a copy of the
 * published positions (it neither re-bins the integrator nor resets its displacement accumulators), so probing does
 * not change when the next list rebuild happens.  On slabs the calls are collective and return the whole system. */
int ljgpu_get_cell_ids(ljgpu_sim* s, uint32_t* cell_of_atom, uint32_t dims[3]);
int ljgpu_get_neighbor_counts(ljgpu_sim* s, double cutoff, int inclusive, uint32_t* count_of_atom);

/* Per-atom lengths of the lists the force kernel actually walks (Verlet path), original atom order:
 *   which = 0: the list built at the last re-binning, the reference's neighbor_list (lennardjonessimulation.cpp:518-550:
 *              j != i with min-image dist2 <= (rcut+shell)^2 at the positions of that re-binning); its lengths equal the
 *              reference's per-atom list lengths.  The acceptance test here contracts d2 with FMA, so a pair whose
 *              dist2 lies within an ulp of (rcut+shell)^2 may be classified differently - such a pair is 0.4 sigma
 *              outside the force cutoff and cannot change a force; ljgpu_get_neighbor_counts is the bit-exact probe;
 *   which = 1: the pruned inner list (entries within rcut + 0.12 sigma at the last pruning), LJGPU_ESTATE if none exists. */
int ljgpu_get_list_counts(ljgpu_sim* s, int which, uint32_t* count_of_atom);

/* write_to_movie_file (lennardjonessimulation.cpp:121-157): u32 N when create != 0, then one
 * frame = 3 f64 box edges + N x 7 f64 {x,y,z,vx,vy,vz,|v|}, native endianness. */
int ljgpu_write_movie_frame(ljgpu_sim* s, const char* path, int create);

/* ---- introspection --------------------------------------------------------------------------- */

/* enabled = 0 off; 1 = every step timed stage by stage; k > 1 = every k-th step timed (the others run as
 * CUDA-graph replays), stage times scaled by k so that ms_* / n_* stay per-launch averages. */
int ljgpu_set_profiling(ljgpu_sim* s, int enabled);
int ljgpu_get_timers(ljgpu_sim* s, ljgpu_timers* out);
int ljgpu_reset_timers(ljgpu_sim* s);
/* Force path actually in use (LJGPU_KERNEL_*), cells per axis, current list capacity per atom, and the number of
 * steps whose force pass was served by the pruned inner list. */
int ljgpu_get_layout(ljgpu_sim* s, int32_t* force_kernel, uint32_t* cells_per_axis,
                     uint32_t* list_capacity, uint64_t* inner_list_steps);

/* ---- several devices: z-slab decomposition ------------------------------------------------------
 * The box is cut into `world` slabs of whole cell layers along z (the slowest index of the reference's
 * cell id, lennardjonessimulation.cpp:476); device `rank` integrates the atoms of its layers and keeps
 * one halo layer from each ring neighbour.  Per step: boundary-layer positions to both neighbours
 * (NCCL send/recv over NVLink) and one 4-double all-reduce (kinetic sums for the thermostat,
 * potential energy, re-binning vote); at re-binning atoms that crossed a slab face migrate.
 * One process per device.  Every call on a slab handle is COLLECTIVE: all ranks make the same calls in
 * the same order.  Rank 0 obtains the id, the caller distributes it (e.g. torch.distributed, a file).
 * Every rank passes the WHOLE initial state; getters return the whole system on every rank. */
int ljgpu_slab_layers(uint32_t cells_per_axis, int world, int rank, uint32_t* first_layer, uint32_t* end_layer);
int ljgpu_slab_unique_id(void* id128);
int ljgpu_create_slab(const ljgpu_params* p, int rank, int world, const void* nccl_unique_id,
                      const double* x, const double* y, const double* z,
                      const double* vx, const double* vy, const double* vz, ljgpu_sim** out);
int ljgpu_slab_info(ljgpu_sim* s, int32_t* rank, int32_t* world, uint32_t* first_layer, uint32_t* n_layers,
                    uint64_t* atoms_owned, uint64_t* atoms_halo);

/* Mean and maximum Verlet-list length per atom of the current list (0 on the other paths). */
int ljgpu_get_list_stats(ljgpu_sim* s, double* mean_length, uint32_t* max_length);

/* Measured FP64 throughput of one device in TFLOP/s (eight independent DFMA chains per thread, 64 warps per SM,
 * best of five launches, CUDA events): the roofline denominator of the force kernels, measured on the spot instead
 * of quoted (nominal B200: 148 SMs x 64 DFMA/clk x 2 x 1.965 GHz = 37.2).  device = -1: the current device. */
int ljgpu_measure_fp64_peak(int device, double* tflops);

/* Page-locked host memory for query buffers (cudaMallocHost / cudaFreeHost): copies into such a
 * buffer are true DMA; any other host pointer works too, only slower. */
int ljgpu_alloc_host(size_t bytes, void** out);
int ljgpu_free_host(void* p);
/* Page-lock memory the caller already owns (cudaHostRegister / cudaHostUnregister), e.g. a POSIX shared-memory
 * segment that several ranks' processes publish into. */
int ljgpu_register_host(void* p, size_t bytes);
int ljgpu_unregister_host(void* p);

#ifdef __cplusplus
}
#endif
#endif /* LJGPU_H */
