// oracle/ref_harness.cpp - C-callable window onto the UNMODIFIED reference integrator.
//
// TEST INFRASTRUCTURE ONLY.  This file is linked with the reference's own
//   /root/reference/src/simulator/lennardjonessimulation.cpp
//   /root/reference/src/simulator/lennardjonesparameters.cpp
// (compiled where they lie, by path, against the stand-in headers in oracle/shim/) into
// oracle/_ref/libljref.so.  Nothing on the product path may load that library: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
//
// The harness adds no arithmetic of its own to the stepped path: it constructs the reference
// class, calls its integrate(), and copies its private SoA members out.  The two read-only
// probes that the reference computes but does not store (per-atom cell id, in-cutoff pair
// count) repeat the reference's expressions and cite them.
#define private public
#include "lennardjonessimulation.h"
#undef private

#include <chrono>
#include <cstring>
#include <memory>

namespace {
struct Ref {
    std::shared_ptr<LennardJonesParameters> params;
    std::unique_ptr<LennardJonesSimulation> sim;
};
inline Ref* R(void* h) { return static_cast<Ref*>(h); }
inline void copy3(const std::vector<double>& a, const std::vector<double>& b,
                  const std::vector<double>& c, double* x, double* y, double* z) {
    std::memcpy(x, a.data(), a.size() * sizeof(double));
    std::memcpy(y, b.data(), b.size() * sizeof(double));
    std::memcpy(z, c.data(), c.size() * sizeof(double));
}
}

extern "C" {

// Parameter keys and types exactly as MainWindow sets them (src/gui/mainwindow.cpp:67-77).
void* ljref_create(double cell_length, int nr_particles, double kT, double mass, double sigma,
                   double rcut, double epsilon, double tau, double shell, double stepsize) {
    Ref* r = new Ref;
    r->params = std::make_shared<LennardJonesParameters>();
    r->params->set_param("cell_length", cell_length);
    r->params->set_param("nr_particles", nr_particles);
    r->params->set_param("kT", kT);
    r->params->set_param("mass", mass);
    r->params->set_param("sigma", sigma);
    r->params->set_param("rcut", rcut);
    r->params->set_param("epsilon", epsilon);
    r->params->set_param("tau", tau);
    r->params->set_param("shell", shell);
    r->params->set_param("stepsize", stepsize);
    r->sim.reset(new LennardJonesSimulation(r->params));   // ctor: .cpp:31-63
    return r;
}

void ljref_destroy(void* h) { delete R(h); }

void ljref_set_param(void* h, const char* key, double v) { R(h)->params->set_param(key, v); }

int ljref_n(void* h) { return (int)R(h)->sim->get_particle_count(); }

void ljref_integrate(void* h, unsigned step, double dt) { R(h)->sim->integrate(step, dt); }

// Headless loop over integrate(); returns wall seconds (constructor excluded).
double ljref_run(void* h, unsigned first_step, unsigned nsteps, double dt) {
    auto t0 = std::chrono::steady_clock::now();
    for (unsigned s = 0; s < nsteps; ++s) R(h)->sim->integrate(first_step + s, dt);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

void ljref_get_positions(void* h, double* x, double* y, double* z) {
    auto& s = *R(h)->sim;
    const auto& p = s.pos_buf[s.get_render_index()];
    copy3(p.x, p.y, p.z, x, y, z);
}
void ljref_get_velocities(void* h, double* x, double* y, double* z) {
    auto& s = *R(h)->sim; copy3(s.vx, s.vy, s.vz, x, y, z);
}
void ljref_get_forces(void* h, double* x, double* y, double* z) {
    auto& s = *R(h)->sim; copy3(s.fx, s.fy, s.fz, x, y, z);
}
void ljref_get_dij(void* h, double* x, double* y, double* z) {
    auto& s = *R(h)->sim; copy3(s.dijx, s.dijy, s.dijz, x, y, z);
}
void ljref_get_energies(void* h, double* ekin, double* epot, double* etot) {
    auto& s = *R(h)->sim;
    *ekin = s.get_ekin(); *epot = s.get_epot(); *etot = s.get_etot();
}

// Overwrite the state with caller data (original atom order), then do what the constructor
// does after initialize(): build_neighbor_list + calculate_forces on that buffer (.cpp:57-60).
void ljref_set_state(void* h, const double* x, const double* y, const double* z,
                     const double* vx, const double* vy, const double* vz) {
    auto& s = *R(h)->sim;
    const int idx = s.get_render_index();
    const size_t n = s.vx.size();
    std::memcpy(s.pos_buf[idx].x.data(), x, n * 8);
    std::memcpy(s.pos_buf[idx].y.data(), y, n * 8);
    std::memcpy(s.pos_buf[idx].z.data(), z, n * 8);
    std::memcpy(s.vx.data(), vx, n * 8);
    std::memcpy(s.vy.data(), vy, n * 8);
    std::memcpy(s.vz.data(), vz, n * 8);
    s.build_neighbor_list(idx);
    s.calculate_forces(idx);
}

// Individual stages, for stage-level parity and the CPU time split.
void ljref_stage_forces(void* h)  { auto& s = *R(h)->sim; s.calculate_forces(s.get_render_index()); }
void ljref_stage_build(void* h)   { auto& s = *R(h)->sim; s.build_neighbor_list(s.get_render_index()); }
void ljref_stage_wrap(void* h)    { auto& s = *R(h)->sim; s.apply_boundary_conditions(s.get_render_index()); }
void ljref_stage_kick(void* h, double dt)       { R(h)->sim->update_velocities_half_dt(dt); }
void ljref_stage_thermostat(void* h, double dt) { R(h)->sim->apply_berendsen_thermostat(dt); }
int  ljref_stage_check(void* h)   { return R(h)->sim->check_update_neighbor_list() ? 1 : 0; }

unsigned long long ljref_list_size(void* h) { return R(h)->sim->neighbor_list.size(); }

// Per-atom Verlet-list length: distance from the atom's offset to its sentinel
// (layout of .cpp:543-550: [N offsets][list_i..., sentinel i]...).
void ljref_neighbor_counts(void* h, unsigned* counts) {
    auto& s = *R(h)->sim;
    const auto& nl = s.neighbor_list;
    const unsigned n = (unsigned)s.vx.size();
    for (unsigned i = 0; i < n; ++i) {
        unsigned c = 0;
        for (size_t p = nl[i]; nl[p] != i; ++p) ++c;
        counts[i] = c;
    }
}

// Copy atom i's list (without the sentinel); returns its length.
unsigned ljref_neighbors_of(void* h, unsigned i, unsigned* out, unsigned cap) {
    auto& s = *R(h)->sim;
    const auto& nl = s.neighbor_list;
    unsigned c = 0;
    for (size_t p = nl[i]; nl[p] != i; ++p) { if (c < cap) out[c] = nl[p]; ++c; }
    return c;
}

// Per-atom count of list entries the force loop accepts (d2 < rcut^2), walking the reference's
// own list with the expressions of .cpp:291-305.
void ljref_incutoff_counts(void* h, unsigned* counts) {
    auto& s = *R(h)->sim;
    const auto& nl = s.neighbor_list;
    const auto& p = s.pos_buf[s.get_render_index()];
    const unsigned n = (unsigned)s.vx.size();
    const double rcut = s.params->get_param<double>("rcut");
    const double rcutsq = rcut * rcut;
    for (unsigned i = 0; i < n; ++i) {
        unsigned c = 0;
        for (size_t q = nl[i]; nl[q] != i; ++q) {
            const unsigned j = nl[q];
            double dx = p.x[i] - p.x[j], dy = p.y[i] - p.y[j], dz = p.z[i] - p.z[j];
            if (dx >= s.dims.x * 0.5) dx -= s.dims.x; else if (dx < -s.dims.x * 0.5) dx += s.dims.x;
            if (dy >= s.dims.y * 0.5) dy -= s.dims.y; else if (dy < -s.dims.y * 0.5) dy += s.dims.y;
            if (dz >= s.dims.z * 0.5) dz -= s.dims.z; else if (dz < -s.dims.z * 0.5) dz += s.dims.z;
            const double d2 = dx*dx + dy*dy + dz*dz;
            if (d2 >= rcutsq) continue;
            ++c;
        }
        counts[i] = c;
    }
}

// Per-atom cell id of the published positions, expressions of .cpp:447-476.
void ljref_cell_ids(void* h, unsigned* cell, unsigned* dims3) {
    auto& s = *R(h)->sim;
    const double csc = s.params->get_param<double>("rcut") + s.params->get_param<double>("shell");
    unsigned nx = std::max(1u, (unsigned)(s.dims.x / csc));
    unsigned ny = std::max(1u, (unsigned)(s.dims.y / csc));
    unsigned nz = std::max(1u, (unsigned)(s.dims.z / csc));
    const double dx = s.dims.x / (double)nx, dy = s.dims.y / (double)ny, dz = s.dims.z / (double)nz;
    const auto& p = s.pos_buf[s.get_render_index()];
    const unsigned n = (unsigned)s.vx.size();
    for (unsigned i = 0; i < n; ++i) {
        unsigned ix = std::min((unsigned)(p.x[i] / dx), nx - 1);
        unsigned iy = std::min((unsigned)(p.y[i] / dy), ny - 1);
        unsigned iz = std::min((unsigned)(p.z[i] / dz), nz - 1);
        cell[i] = iz * nx * ny + iy * nx + ix;
    }
    dims3[0] = nx; dims3[1] = ny; dims3[2] = nz;
}

void ljref_write_movie(void* h, const char* path, int create) {
    R(h)->sim->write_to_movie_file(path, create != 0);
}

} // extern "C"
