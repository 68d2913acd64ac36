/* oracle/lj_oracle.h - CPU restatement of the reference Lennard-Jones integrator.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load liblj_oracle.so.  The product (libljgpu.so) never does.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors,
 * so the restatement is pinned against the reference's own sources compiled unmodified
 * (oracle/_ref/libljref.so, see oracle/Makefile) - bit-for-bit, see tests/test_oracle.py - and
 * against the fixtures under tests/golden/ that tests/golden/make_golden.py wrote from that
 * library.
 *
 * All file:line citations are into /root/reference/src/simulator/.
 */
#ifndef LJ_ORACLE_H
#define LJ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The ten keys of LennardJonesParameters as MainWindow fills them (src/gui/mainwindow.cpp:67-77). */
typedef struct ljo_params {
    double cell_length;
    int    nr_particles;
    double kT, mass, sigma, rcut, epsilon, tau, shell, stepsize;
} ljo_params;

typedef struct ljo_sim ljo_sim;

/* Restart the process-wide velocity generator (the reference's is a function-static that is
 * never reseeded, lennardjonessimulation.cpp:248-252; a fresh process starts here). */
void ljo_rng_reset(void);

/* Constructor semantics of lennardjonessimulation.cpp:31-63. */
ljo_sim* ljo_create(const ljo_params* p);
/* Same, but with caller-supplied positions/velocities instead of initialize(). */
ljo_sim* ljo_create_from_state(const ljo_params* p, const double* x, const double* y, const double* z,
                               const double* vx, const double* vy, const double* vz);
void ljo_destroy(ljo_sim* s);

void ljo_set_tau(ljo_sim* s, double tau);

/* integrate(step, dt), lennardjonessimulation.cpp:70-114 (the log line is omitted). */
void   ljo_integrate(ljo_sim* s, unsigned step, double dt);
/* Headless loop; returns wall seconds. */
double ljo_run(ljo_sim* s, unsigned first_step, unsigned nsteps, double dt);

int  ljo_n(const ljo_sim* s);
void ljo_get_positions(const ljo_sim* s, double* x, double* y, double* z);
void ljo_get_velocities(const ljo_sim* s, double* x, double* y, double* z);
void ljo_get_forces(const ljo_sim* s, double* x, double* y, double* z);
void ljo_get_dij(const ljo_sim* s, double* x, double* y, double* z);
void ljo_get_energies(const ljo_sim* s, double* ekin, double* epot, double* etot);

uint64_t ljo_list_size(const ljo_sim* s);          /* == reference neighbor_list.size() */
uint64_t ljo_rebuild_count(const ljo_sim* s);      /* list builds so far, constructor's included */
void ljo_neighbor_counts(const ljo_sim* s, unsigned* counts);
unsigned ljo_neighbors_of(const ljo_sim* s, unsigned i, unsigned* out, unsigned cap);
void ljo_incutoff_counts(const ljo_sim* s, unsigned* counts);
void ljo_cell_ids(const ljo_sim* s, unsigned* cell, unsigned* dims3);
/* Test yardstick: per-atom sum of |f_ij| over the accepted pairs. */
void ljo_force_abs_sums(const ljo_sim* s, double* out);
/* Test yardstick: epot of the current positions/list accumulated in long double. */
double ljo_epot_long_double(const ljo_sim* s);

/* Stateless helpers on caller arrays. */
void ljo_cell_ids_of(const ljo_params* p, int n, const double* x, const double* y, const double* z,
                     unsigned* cell, unsigned* dims3);
/* Brute-force O(N^2) per-atom count of j != i with min-image dist2 <= cutoff^2
 * (the acceptance test of lennardjonessimulation.cpp:518-540 without the cell pre-filter). */
void ljo_pair_counts_bruteforce(double cell_length, int n, const double* x, const double* y,
                                const double* z, double cutoff, unsigned* counts);

/* GraphWidget::set_particle_speed, graphwidget.cpp:205-233 (NUMBINS=100, MAXSPEED=10.0 are
 * graphwidget.h:73-74; here they are arguments). */
void ljo_speed_histogram(int n, const double* vx, const double* vy, const double* vz,
                         unsigned nbins, double maxspeed, uint64_t* counts);
/* GraphWidget::set_params, graphwidget.cpp:55-73: expected count per bin at the left bin edge. */
void ljo_maxwell_boltzmann(double kT, double m, int numparts, unsigned nbins, double maxspeed,
                           double* expected);

/* write_to_movie_file, lennardjonessimulation.cpp:121-157. Returns 0 on success. */
int ljo_write_movie(const ljo_sim* s, const char* path, int create);

#ifdef __cplusplus
}
#endif
#endif
