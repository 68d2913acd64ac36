/* oracle/lj_oracle.c - plain-C restatement of the reference Lennard-Jones integrator.
 *
 * TEST INFRASTRUCTURE ONLY (see lj_oracle.h).  Parity status: PINNED against the reference's own
 * sources compiled unmodified (oracle/_ref) and against tests/golden/.
 *
 * Every function cites the lines of /root/reference/src/simulator/lennardjonessimulation.cpp
 * (".cpp" below) whose arithmetic it restates.  The data structures are this file's own (one
 * position buffer instead of two, CSR cell table and CSR neighbour list instead of nested
 * vectors) but every floating-point expression is evaluated in the reference's order so that the
 * results agree to the last bit.  Build with -ffp-contract=off and no -march=native: the
 * reference's arithmetic is defined by an x86-64 baseline build, which has no fused multiply-add.
 */
#define _GNU_SOURCE 1
#include "lj_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

struct ljo_sim {
    ljo_params p;
    int     n;
    double  L;                 /* dims.x = dims.y = dims.z = cell_length (.cpp:36-37) */
    double *x, *y, *z;         /* published positions */
    double *vx, *vy, *vz;
    double *fx, *fy, *fz;
    double *dx, *dy, *dz;      /* accumulated displacement since the last list build */
    uint64_t *nl_off;          /* n+1 offsets into nl */
    unsigned *nl;              /* neighbour entries, per atom in the reference's order */
    uint64_t  nl_cap;
    uint64_t  builds;
    double ekin, epot, etot;
};

/* ------------------------------------------------------------------------------------------
 * Velocity generator: .cpp:248-252 uses a function-static std::default_random_engine and
 * std::normal_distribution<double>(0,1).  With libstdc++ (g++ 13) these are
 *   minstd_rand0:  x <- 16807 * x mod (2^31 - 1), seed 1, range [1, 2^31 - 2];
 *   generate_canonical<double,53>: two draws, (u1-1) + (u2-1)*R over R^2 with R = 2^31 - 2;
 *   normal_distribution: Marsaglia polar method, returns y*mult first and keeps x*mult.
 * Restated here; tests/test_oracle.py checks the stream bit-for-bit against oracle/_ref.
 * ------------------------------------------------------------------------------------------ */
static uint64_t g_lcg = 1u;
static int      g_have_saved = 0;
static double   g_saved = 0.0;

void ljo_rng_reset(void) { g_lcg = 1u; g_have_saved = 0; g_saved = 0.0; }

static uint32_t minstd0_next(void) {
    g_lcg = (g_lcg * 16807u) % 2147483647u;
    return (uint32_t)g_lcg;
}

static double canonical53(void) {
    const long double R = 2147483646.0L;          /* max - min + 1 */
    double sum = 0.0, tmp = 1.0;
    for (int k = 0; k < 2; ++k) {
        sum += (double)(minstd0_next() - 1u) * tmp;
        tmp = (double)((long double)tmp * R);
    }
    double ret = sum / tmp;
    if (ret >= 1.0) ret = nextafter(1.0, 0.0);
    return ret;
}

static double gauss01(void) {
    double ret;
    if (g_have_saved) {
        g_have_saved = 0;
        ret = g_saved;
    } else {
        double a, b, r2;
        do {
            a = 2.0 * canonical53() - 1.0;
            b = 2.0 * canonical53() - 1.0;
            r2 = a * a + b * b;
        } while (r2 > 1.0 || r2 == 0.0);
        const double mult = sqrt(-2 * log(r2) / r2);
        g_saved = a * mult;
        g_have_saved = 1;
        ret = b * mult;
    }
    return ret * 1.0 + 0.0;                       /* stddev 1, mean 0 */
}

/* ------------------------------------------------------------------------------------------ */

static double* dalloc(size_t n) {
    double* p = (double*)calloc(n ? n : 1, sizeof(double));
    if (!p) { fprintf(stderr, "lj_oracle: out of memory\n"); abort(); }
    return p;
}

/* One-shot minimum image of .cpp:295-302 and :523-530. */
static inline double min_image(double d, double L) {
    if (d >= L * 0.5) d -= L;
    else if (d < -L * 0.5) d += L;
    return d;
}

/* Cells per axis and edge, .cpp:447-457. */
static void cell_grid(double L, double rcut, double shell, unsigned* nc, double* edge) {
    const double csc = rcut + shell;
    unsigned n = (unsigned)(L / csc);
    if (n < 1u) n = 1u;
    *nc = n;
    *edge = L / (double)n;
}

/* Cell coordinate along one axis, .cpp:468-474. */
static inline unsigned cell_coord(double x, double edge, unsigned nc) {
    unsigned i = (unsigned)(x / edge);
    return i < nc - 1 ? i : nc - 1;
}

/* initialize(), .cpp:185-246: simple-cubic lattice, x outermost / z innermost, first n sites;
 * Gaussian velocities in the order vx,vy,vz per atom; mean velocity removed; dij zeroed. */
static void initialize(ljo_sim* s) {
    const size_t n = (size_t)s->n;
    const double L = s->L;
    const double volume = L * L * L;
    const double dl = pow(volume / (double)n, 1.0 / 3.0);
    const size_t nx = (size_t)ceil(L / dl), ny = nx, nz = nx;
    const double sp = L / (double)nx;

    size_t count = 0;
    for (size_t i = 0; i < nx; i++) {
        const double X = (i + 0.5) * sp;
        for (size_t j = 0; j < ny; j++) {
            const double Y = (j + 0.5) * sp;
            for (size_t k = 0; k < nz; k++) {
                const double Z = (k + 0.5) * sp;
                if (count >= n) break;
                s->x[count] = X; s->y[count] = Y; s->z[count] = Z;
                count++;
            }
        }
    }

    const double sqrtktm = sqrt(s->p.kT / s->p.mass);
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (size_t i = 0; i < n; i++) {
        s->vx[i] = sqrtktm * gauss01();
        s->vy[i] = sqrtktm * gauss01();
        s->vz[i] = sqrtktm * gauss01();
        sx += s->vx[i]; sy += s->vy[i]; sz += s->vz[i];
    }
    const double invn = 1.0 / (double)n;
    sx *= invn; sy *= invn; sz *= invn;
    for (size_t i = 0; i < n; i++) { s->vx[i] -= sx; s->vy[i] -= sy; s->vz[i] -= sz; }

    memset(s->dx, 0, n * sizeof(double));
    memset(s->dy, 0, n * sizeof(double));
    memset(s->dz, 0, n * sizeof(double));
}

/* build_neighbor_list(), .cpp:441-551.  Cells hold atoms in ascending index (the reference
 * appends in index order, :467-478); per atom the 27 stencil cells are visited cx outer, cy, cz
 * inner (:496-516), so the per-atom entry order equals the reference's.  With fewer than three
 * cells on an axis the stencil revisits cells and pairs are entered more than once, exactly as
 * the reference does. */
static void build_list(ljo_sim* s) {
    const unsigned n = (unsigned)s->n;
    const double L = s->L;
    const double csc = s->p.rcut + s->p.shell;
    const double cutsq = csc * csc;
    unsigned nc; double edge;
    cell_grid(L, s->p.rcut, s->p.shell, &nc, &edge);
    const unsigned ncell = nc * nc * nc;

    unsigned* cstart = (unsigned*)calloc((size_t)ncell + 1, sizeof(unsigned));
    unsigned* cfill  = (unsigned*)calloc((size_t)ncell, sizeof(unsigned));
    unsigned* catom  = (unsigned*)malloc((size_t)(n ? n : 1) * sizeof(unsigned));
    unsigned* cof    = (unsigned*)malloc((size_t)(n ? n : 1) * sizeof(unsigned));
    for (unsigned i = 0; i < n; i++) {
        const unsigned ix = cell_coord(s->x[i], edge, nc);
        const unsigned iy = cell_coord(s->y[i], edge, nc);
        const unsigned iz = cell_coord(s->z[i], edge, nc);
        cof[i] = iz * nc * nc + iy * nc + ix;
        cstart[cof[i] + 1]++;
    }
    for (unsigned c = 0; c < ncell; c++) cstart[c + 1] += cstart[c];
    for (unsigned i = 0; i < n; i++) catom[cstart[cof[i]] + cfill[cof[i]]++] = i;

    uint64_t used = 0;
    for (unsigned i = 0; i < n; i++) {
        s->dx[i] = 0.0; s->dy[i] = 0.0; s->dz[i] = 0.0;            /* :484-486 */
        s->nl_off[i] = used;

        const unsigned c = cof[i];
        const unsigned ix = c % nc, iy = (c / nc) % nc, iz = c / (nc * nc);
        const unsigned xx[3] = { ix == 0 ? nc - 1 : ix - 1, ix, ix == nc - 1 ? 0 : ix + 1 };
        const unsigned yy[3] = { iy == 0 ? nc - 1 : iy - 1, iy, iy == nc - 1 ? 0 : iy + 1 };
        const unsigned zz[3] = { iz == 0 ? nc - 1 : iz - 1, iz, iz == nc - 1 ? 0 : iz + 1 };

        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) for (int d = 0; d < 3; d++) {
            const unsigned cid = zz[d] * nc * nc + yy[b] * nc + xx[a];
            for (unsigned q = cstart[cid]; q < cstart[cid + 1]; q++) {
                const unsigned pid = catom[q];
                if (pid == i) continue;
                const double ddx = min_image(s->x[i] - s->x[pid], L);
                const double ddy = min_image(s->y[i] - s->y[pid], L);
                const double ddz = min_image(s->z[i] - s->z[pid], L);
                const double dist2 = ddx * ddx + ddy * ddy + ddz * ddz;
                if (dist2 <= cutsq) {
                    if (used == s->nl_cap) {
                        s->nl_cap = s->nl_cap ? s->nl_cap * 2 : (uint64_t)n * 64 + 1024;
                        s->nl = (unsigned*)realloc(s->nl, s->nl_cap * sizeof(unsigned));
                        if (!s->nl) { fprintf(stderr, "lj_oracle: out of memory\n"); abort(); }
                    }
                    s->nl[used++] = pid;
                }
            }
        }
    }
    s->nl_off[n] = used;
    s->builds++;
    free(cstart); free(cfill); free(catom); free(cof);
}

/* calculate_forces(), .cpp:258-330: full (two-sided) list, shifted potential, unshifted force,
 * sequential accumulation in list order, 24*eps applied afterwards. */
static void forces(ljo_sim* s) {
    const int n = s->n;
    const double L = s->L;
    const double rcutsq = s->p.rcut * s->p.rcut;
    const double sigmasq = s->p.sigma * s->p.sigma;
    const double prefctr = 24.0 * s->p.epsilon;
    const double sr2 = sigmasq / rcutsq;
    const double sr6 = sr2 * sr2 * sr2;
    const double sr12 = sr6 * sr6;
    const double epot_cutoff = sr12 - sr6;

    double epotsum = 0.0;
    for (int i = 0; i < n; i++) {
        double ax = 0.0, ay = 0.0, az = 0.0;
        for (uint64_t q = s->nl_off[i]; q < s->nl_off[i + 1]; q++) {
            const unsigned j = s->nl[q];
            const double ddx = min_image(s->x[i] - s->x[j], L);
            const double ddy = min_image(s->y[i] - s->y[j], L);
            const double ddz = min_image(s->z[i] - s->z[j], L);
            const double d2 = ddx * ddx + ddy * ddy + ddz * ddz;
            if (d2 >= rcutsq) continue;
            const double invd2 = 1.0 / d2;
            const double s2 = sigmasq * invd2;
            const double s6 = s2 * s2 * s2;
            const double s12 = s6 * s6;
            epotsum += (s12 - s6 - epot_cutoff);
            const double fr = (2.0 * s12 - s6) * invd2;
            ax += fr * ddx; ay += fr * ddy; az += fr * ddz;
        }
        s->fx[i] = ax; s->fy[i] = ay; s->fz[i] = az;
    }
    for (int i = 0; i < n; i++) { s->fx[i] *= prefctr; s->fy[i] *= prefctr; s->fz[i] *= prefctr; }

    double ep = epotsum * 4.0 * s->p.epsilon;
    ep *= 0.5;
    s->epot = ep;
}

/* update_velocities_half_dt(), .cpp:336-362 (OpenMP off: one sequential sum in index order). */
static void half_kick(ljo_sim* s, double dt) {
    const int n = s->n;
    const double m = s->p.mass;
    const double factor = 0.5 * dt / m;
    double sum = 0.0;
    for (int i = 0; i < n; i++) {
        s->vx[i] += factor * s->fx[i];
        s->vy[i] += factor * s->fy[i];
        s->vz[i] += factor * s->fz[i];
        sum += (s->vx[i] * s->vx[i] + s->vy[i] * s->vy[i] + s->vz[i] * s->vz[i]);
    }
    s->ekin = 0.5 * m * sum;
}

/* apply_berendsen_thermostat(), .cpp:368-383. */
static void berendsen(ljo_sim* s, double dt) {
    const size_t n = (size_t)s->n;
    const double ekin0 = 1.5 * (double)(n - 1) * s->p.kT;
    const double lambda = sqrt(1.0 + dt / s->p.tau * (ekin0 / s->ekin - 1.0));
    for (size_t i = 0; i < n; i++) { s->vx[i] *= lambda; s->vy[i] *= lambda; s->vz[i] *= lambda; }
}

/* update_positions(), .cpp:388-413 (the read and write buffers collapse into one here). */
static void drift(ljo_sim* s, double dt) {
    const int n = s->n;
    for (int i = 0; i < n; i++) {
        const double ix = s->vx[i] * dt, iy = s->vy[i] * dt, iz = s->vz[i] * dt;
        s->x[i] = s->x[i] + ix; s->y[i] = s->y[i] + iy; s->z[i] = s->z[i] + iz;
        s->dx[i] += ix; s->dy[i] += iy; s->dz[i] += iz;
    }
}

/* apply_boundary_conditions(), .cpp:419-435. */
static void wrap(ljo_sim* s) {
    const int n = s->n;
    const double L = s->L, invL = 1.0 / L;
    for (int i = 0; i < n; i++) {
        s->x[i] -= L * floor(s->x[i] * invL);
        s->y[i] -= L * floor(s->y[i] * invL);
        s->z[i] -= L * floor(s->z[i] * invL);
    }
}

/* check_update_neighbor_list(), .cpp:557-568. */
static int needs_rebuild(const ljo_sim* s) {
    const double thresh = s->p.shell * s->p.shell * 0.25;
    for (int i = 0; i < s->n; i++) {
        const double d2 = s->dx[i] * s->dx[i] + s->dy[i] * s->dy[i] + s->dz[i] * s->dz[i];
        if (d2 > thresh) return 1;
    }
    return 0;
}

static ljo_sim* alloc_sim(const ljo_params* p) {
    ljo_sim* s = (ljo_sim*)calloc(1, sizeof(ljo_sim));
    s->p = *p;
    s->n = p->nr_particles;
    s->L = p->cell_length;
    const size_t n = (size_t)s->n;
    s->x = dalloc(n); s->y = dalloc(n); s->z = dalloc(n);
    s->vx = dalloc(n); s->vy = dalloc(n); s->vz = dalloc(n);
    s->fx = dalloc(n); s->fy = dalloc(n); s->fz = dalloc(n);
    s->dx = dalloc(n); s->dy = dalloc(n); s->dz = dalloc(n);
    s->nl_off = (uint64_t*)calloc(n + 1, sizeof(uint64_t));
    return s;
}

/* Constructor, .cpp:31-63: initialize, list, forces; ekin stays 0 until the first step. */
ljo_sim* ljo_create(const ljo_params* p) {
    ljo_sim* s = alloc_sim(p);
    initialize(s);
    build_list(s);
    forces(s);
    return s;
}

ljo_sim* ljo_create_from_state(const ljo_params* p, const double* x, const double* y, const double* z,
                               const double* vx, const double* vy, const double* vz) {
    ljo_sim* s = alloc_sim(p);
    const size_t b = (size_t)s->n * sizeof(double);
    memcpy(s->x, x, b); memcpy(s->y, y, b); memcpy(s->z, z, b);
    memcpy(s->vx, vx, b); memcpy(s->vy, vy, b); memcpy(s->vz, vz, b);
    build_list(s);
    forces(s);
    return s;
}

void ljo_destroy(ljo_sim* s) {
    if (!s) return;
    free(s->x); free(s->y); free(s->z); free(s->vx); free(s->vy); free(s->vz);
    free(s->fx); free(s->fy); free(s->fz); free(s->dx); free(s->dy); free(s->dz);
    free(s->nl_off); free(s->nl); free(s);
}

void ljo_set_tau(ljo_sim* s, double tau) { s->p.tau = tau; }

/* integrate(), .cpp:70-98: kick, thermostat, drift, forces on the drifted (not yet wrapped)
 * positions, wrap, conditional rebuild, kick, etot. */
void ljo_integrate(ljo_sim* s, unsigned step, double dt) {
    (void)step;
    half_kick(s, dt);
    berendsen(s, dt);
    drift(s, dt);
    forces(s);
    wrap(s);
    if (needs_rebuild(s)) build_list(s);
    half_kick(s, dt);
    s->etot = s->ekin + s->epot;
}

double ljo_run(ljo_sim* s, unsigned first_step, unsigned nsteps, double dt) {
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (unsigned k = 0; k < nsteps; k++) ljo_integrate(s, first_step + k, dt);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

int ljo_n(const ljo_sim* s) { return s->n; }

static void out3(const ljo_sim* s, const double* a, const double* b, const double* c,
                 double* x, double* y, double* z) {
    const size_t bytes = (size_t)s->n * sizeof(double);
    memcpy(x, a, bytes); memcpy(y, b, bytes); memcpy(z, c, bytes);
}
void ljo_get_positions(const ljo_sim* s, double* x, double* y, double* z)  { out3(s, s->x, s->y, s->z, x, y, z); }
void ljo_get_velocities(const ljo_sim* s, double* x, double* y, double* z) { out3(s, s->vx, s->vy, s->vz, x, y, z); }
void ljo_get_forces(const ljo_sim* s, double* x, double* y, double* z)     { out3(s, s->fx, s->fy, s->fz, x, y, z); }
void ljo_get_dij(const ljo_sim* s, double* x, double* y, double* z)        { out3(s, s->dx, s->dy, s->dz, x, y, z); }
void ljo_get_energies(const ljo_sim* s, double* ekin, double* epot, double* etot) {
    *ekin = s->ekin; *epot = s->epot; *etot = s->etot;
}

/* The reference's flat vector is [n offsets] + per atom (entries + 1 sentinel), .cpp:543-550. */
uint64_t ljo_list_size(const ljo_sim* s) { return (uint64_t)s->n * 2u + s->nl_off[s->n]; }
uint64_t ljo_rebuild_count(const ljo_sim* s) { return s->builds; }

void ljo_neighbor_counts(const ljo_sim* s, unsigned* counts) {
    for (int i = 0; i < s->n; i++) counts[i] = (unsigned)(s->nl_off[i + 1] - s->nl_off[i]);
}

unsigned ljo_neighbors_of(const ljo_sim* s, unsigned i, unsigned* out, unsigned cap) {
    const uint64_t a = s->nl_off[i], b = s->nl_off[i + 1];
    for (uint64_t q = a; q < b && (q - a) < cap; q++) out[q - a] = s->nl[q];
    return (unsigned)(b - a);
}

/* List entries the force loop accepts, .cpp:291-305. */
void ljo_incutoff_counts(const ljo_sim* s, unsigned* counts) {
    const double L = s->L, rcutsq = s->p.rcut * s->p.rcut;
    for (int i = 0; i < s->n; i++) {
        unsigned c = 0;
        for (uint64_t q = s->nl_off[i]; q < s->nl_off[i + 1]; q++) {
            const unsigned j = s->nl[q];
            const double ddx = min_image(s->x[i] - s->x[j], L);
            const double ddy = min_image(s->y[i] - s->y[j], L);
            const double ddz = min_image(s->z[i] - s->z[j], L);
            const double d2 = ddx * ddx + ddy * ddy + ddz * ddz;
            if (d2 >= rcutsq) continue;
            c++;
        }
        counts[i] = c;
    }
}

/* Per-atom sum over accepted list entries of |f_ij| (same pair expressions as forces()): the
 * scale against which a rounding-level force difference has to be judged when the net force
 * nearly cancels (perfect start lattice). Not a reference quantity - a test yardstick. */
void ljo_force_abs_sums(const ljo_sim* s, double* out) {
    const double L = s->L, rcutsq = s->p.rcut * s->p.rcut, sigmasq = s->p.sigma * s->p.sigma;
    const double prefctr = 24.0 * s->p.epsilon;
    for (int i = 0; i < s->n; i++) {
        double acc = 0.0;
        for (uint64_t q = s->nl_off[i]; q < s->nl_off[i + 1]; q++) {
            const unsigned j = s->nl[q];
            const double ddx = min_image(s->x[i] - s->x[j], L);
            const double ddy = min_image(s->y[i] - s->y[j], L);
            const double ddz = min_image(s->z[i] - s->z[j], L);
            const double d2 = ddx * ddx + ddy * ddy + ddz * ddz;
            if (d2 >= rcutsq) continue;
            const double invd2 = 1.0 / d2;
            const double s2 = sigmasq * invd2;
            const double s6 = s2 * s2 * s2;
            const double s12 = s6 * s6;
            const double fr = (2.0 * s12 - s6) * invd2;
            acc += fabs(fr) * sqrt(d2);
        }
        out[i] = prefctr * acc;
    }
}

/* Test yardstick: the potential energy of the current positions over the current list with every
 * pair term formed as in forces() but accumulated in long double.  The reference adds ~N*55 terms
 * into ONE double in list order (.cpp:284-312); on the perfect start lattice identical terms
 * repeat and the rounding errors of that running sum do not average out, so its own deviation
 * from the exact sum can exceed 1e-12 relative.  This value measures that deviation. */
double ljo_epot_long_double(const ljo_sim* s) {
    const double L = s->L, rcutsq = s->p.rcut * s->p.rcut, sigmasq = s->p.sigma * s->p.sigma;
    const double sr2 = sigmasq / rcutsq, sr6 = sr2 * sr2 * sr2, sr12 = sr6 * sr6;
    const double epot_cutoff = sr12 - sr6;
    long double sum = 0.0L;
    for (int i = 0; i < s->n; i++) {
        for (uint64_t q = s->nl_off[i]; q < s->nl_off[i + 1]; q++) {
            const unsigned j = s->nl[q];
            const double ddx = min_image(s->x[i] - s->x[j], L);
            const double ddy = min_image(s->y[i] - s->y[j], L);
            const double ddz = min_image(s->z[i] - s->z[j], L);
            const double d2 = ddx * ddx + ddy * ddy + ddz * ddz;
            if (d2 >= rcutsq) continue;
            const long double invd2 = 1.0L / (long double)d2;
            const long double s2 = (long double)sigmasq * invd2;
            const long double s6 = s2 * s2 * s2;
            sum += s6 * s6 - s6 - (long double)epot_cutoff;
        }
    }
    return (double)(sum * 4.0L * (long double)s->p.epsilon * 0.5L);
}

void ljo_cell_ids_of(const ljo_params* p, int n, const double* x, const double* y, const double* z,
                     unsigned* cell, unsigned* dims3) {
    unsigned nc; double edge;
    cell_grid(p->cell_length, p->rcut, p->shell, &nc, &edge);
    for (int i = 0; i < n; i++) {
        const unsigned ix = cell_coord(x[i], edge, nc);
        const unsigned iy = cell_coord(y[i], edge, nc);
        const unsigned iz = cell_coord(z[i], edge, nc);
        cell[i] = iz * nc * nc + iy * nc + ix;                      /* .cpp:476 */
    }
    dims3[0] = dims3[1] = dims3[2] = nc;
}

void ljo_cell_ids(const ljo_sim* s, unsigned* cell, unsigned* dims3) {
    ljo_cell_ids_of(&s->p, s->n, s->x, s->y, s->z, cell, dims3);
}

void ljo_pair_counts_bruteforce(double L, int n, const double* x, const double* y, const double* z,
                                double cutoff, unsigned* counts) {
    const double cutsq = cutoff * cutoff;
    for (int i = 0; i < n; i++) {
        unsigned c = 0;
        for (int j = 0; j < n; j++) {
            if (j == i) continue;
            const double ddx = min_image(x[i] - x[j], L);
            const double ddy = min_image(y[i] - y[j], L);
            const double ddz = min_image(z[i] - z[j], L);
            const double dist2 = ddx * ddx + ddy * ddy + ddz * ddz;
            if (dist2 <= cutsq) c++;
        }
        counts[i] = c;
    }
}

/* graphwidget.cpp:225-230: speed = glm::length(v) = sqrt(x*x + y*y + z*z);
 * bin = min((int)(speed / MAXSPEED * NUMBINS), NUMBINS - 1). */
void ljo_speed_histogram(int n, const double* vx, const double* vy, const double* vz,
                         unsigned nbins, double maxspeed, uint64_t* counts) {
    for (unsigned b = 0; b < nbins; b++) counts[b] = 0;
    for (int i = 0; i < n; i++) {
        const double speed = sqrt(vx[i] * vx[i] + vy[i] * vy[i] + vz[i] * vz[i]);
        int bin = (int)(speed / maxspeed * nbins);
        if (bin > (int)nbins - 1) bin = (int)nbins - 1;
        counts[bin]++;
    }
}

/* graphwidget.cpp:60-68. */
void ljo_maxwell_boltzmann(double kT, double m, int numparts, unsigned nbins, double maxspeed,
                           double* expected) {
    const double binsize = maxspeed / (double)nbins;
    for (unsigned i = 0; i < nbins; i++) {
        const double v = binsize * i;
        const double f = 4.0 * M_PI * pow(m / (2.0 * M_PI * kT), 3. / 2.) * v * v * exp(-m * v * v / (2.0 * kT));
        expected[i] = f * numparts * binsize;
    }
}

/* write_to_movie_file(), .cpp:121-157: u32 n on create; per frame 3 f64 dims then per atom
 * 7 f64 {x,y,z,vx,vy,vz,|v|}; native endianness. */
int ljo_write_movie(const ljo_sim* s, const char* path, int create) {
    FILE* f = fopen(path, create ? "wb" : "ab");
    if (!f) return -1;
    if (create) { uint32_t n = (uint32_t)s->n; fwrite(&n, sizeof n, 1, f); }
    const double dims[3] = { s->L, s->L, s->L };
    fwrite(dims, sizeof(double), 3, f);
    for (int i = 0; i < s->n; i++) {
        double rec[7] = { s->x[i], s->y[i], s->z[i], s->vx[i], s->vy[i], s->vz[i], 0.0 };
        rec[6] = sqrt(rec[3] * rec[3] + rec[4] * rec[4] + rec[5] * rec[5]);
        fwrite(rec, sizeof(double), 7, f);
    }
    fclose(f);
    return 0;
}
