// lennardjonessimulation.cpp - see lennardjonessimulation.h.  Every method forwards to the C ABI.
#include "lennardjonessimulation.h"

#include <cstdio>
#include <cstdlib>
#include <stdexcept>

#include "../../include/ljgpu.h"

void LennardJonesSimulation::fail(const char* what) {
    throw std::runtime_error(std::string(what) + ": " + ljgpu_last_error());
}

namespace {
ljgpu_params to_struct(const LennardJonesParameters& p, int kernel, int device) {
    ljgpu_params s{};
    // same keys, same types as the reference reads them (.cpp:36,39,222-223,259-266,337-338,369-371,442-443,558)
    s.cell_length = p.get_param<double>("cell_length");
    s.nr_particles = p.get_param<int>("nr_particles");
    s.kT = p.get_param<double>("kT");
    s.mass = p.get_param<double>("mass");
    s.sigma = p.get_param<double>("sigma");
    s.rcut = p.get_param<double>("rcut");
    s.epsilon = p.get_param<double>("epsilon");
    s.tau = p.get_param<double>("tau");
    s.shell = p.get_param<double>("shell");
    s.stepsize = p.get_param<double>("stepsize");
    s.force_kernel = kernel;
    s.device = device;
    return s;
}
}

LennardJonesSimulation::LennardJonesSimulation(const std::shared_ptr<LennardJonesParameters>& _params,
                                               int force_kernel, int device)
    : params(_params), kernel_choice(force_kernel), device_choice(device) {
    this->ttime = std::chrono::system_clock::now();                       // .cpp:34
    const ljgpu_params s = to_struct(*params, force_kernel, device);
    this->dims = dvec3(s.cell_length, s.cell_length, s.cell_length);      // .cpp:36-37
    this->n = (size_t)s.nr_particles;
    if (s.nr_particles < 1) throw std::runtime_error("nr_particles must be positive");

    std::vector<double> st(6 * n);
    if (ljgpu_host_initialize(&s, &st[0], &st[n], &st[2 * n], &st[3 * n], &st[4 * n], &st[5 * n]) != LJGPU_OK)
        throw std::runtime_error("ljgpu_host_initialize: invalid parameters");
    if (ljgpu_create(&s, &st[0], &st[n], &st[2 * n], &st[3 * n], &st[4 * n], &st[5 * n], &sim) != LJGPU_OK)
        fail("ljgpu_create");
    for (int b = 0; b < kMirrors; ++b) {
        void* p = nullptr;
        if (ljgpu_alloc_host(3 * n * sizeof(double), &p) != LJGPU_OK) { ljgpu_destroy(sim); sim = nullptr; fail("ljgpu_alloc_host"); }
        mirror_buf[b] = static_cast<double*>(p);
    }
    mirror = n <= 131072 ? Mirror::Eager : Mirror::Lazy;
    refresh_mirror();                                                     // .cpp:62 publish_state
    refresh_energies();
}

LennardJonesSimulation::~LennardJonesSimulation() {
    if (sim) ljgpu_destroy(sim);
    for (int b = 0; b < kMirrors; ++b) if (mirror_buf[b]) ljgpu_free_host(mirror_buf[b]);
}

void LennardJonesSimulation::refresh_energies() {
    double k = 0, p = 0, t = 0;
    if (ljgpu_get_energies(sim, &k, &p, &t) != LJGPU_OK) fail("ljgpu_get_energies");
    ekin.store(k, std::memory_order_relaxed);
    epot.store(p, std::memory_order_relaxed);
    etot.store(t, std::memory_order_relaxed);
}

// Makes the newest positions the readers' buffer: completes a delivery that is in flight, or copies the
// published positions into a buffer the readers are NOT looking at; then flips the index with release
// semantics: the protocol of .h:76-84 / .cpp:574-576.
void LennardJonesSimulation::refresh_mirror() const {
    std::lock_guard<std::mutex> lk(mirror_mu);
    const uint64_t now = steps_done.load(std::memory_order_acquire);
    if (inflight_idx >= 0) {
        if (ljgpu_publish_wait(sim) != LJGPU_OK) fail("ljgpu_publish_wait");
        render_idx.store(inflight_idx, std::memory_order_release);
        mirrored_step.store(inflight_step, std::memory_order_relaxed);
        inflight_idx = -1;
    }
    if (mirrored_step.load(std::memory_order_relaxed) == now) return;
    const int w = (render_idx.load(std::memory_order_relaxed) + 1) % kMirrors;
    double* b = mirror_buf[w];
    if (ljgpu_get_positions(sim, b, b + n, b + 2 * n) != LJGPU_OK) fail("ljgpu_get_positions");
    render_idx.store(w, std::memory_order_release);
    mirrored_step.store(now, std::memory_order_relaxed);
}

void LennardJonesSimulation::integrate(unsigned int step, double dt) {
    if (mirror == Mirror::Eager) {
        // .cpp:98 publish_state: this step's positions start travelling into the free buffer and arrive while the
        // next step computes; the delivery of the previous step is complete when the call returns
        std::lock_guard<std::mutex> lk(mirror_mu);
        const int vis = render_idx.load(std::memory_order_relaxed), fly = inflight_idx;
        int tgt = 0;
        while (tgt == vis || tgt == fly) ++tgt;
        double* b = mirror_buf[tgt];
        if (ljgpu_step_publish(sim, step, dt, b, b + n, b + 2 * n) != LJGPU_OK) fail("ljgpu_step_publish");
        const uint64_t now = steps_done.fetch_add(1, std::memory_order_release) + 1;
        if (fly >= 0) {
            render_idx.store(fly, std::memory_order_release);
            mirrored_step.store(inflight_step, std::memory_order_relaxed);
        }
        inflight_idx = tgt; inflight_step = now;
    } else {
        if (ljgpu_step(sim, step, dt) != LJGPU_OK) fail("ljgpu_step");
        steps_done.fetch_add(1, std::memory_order_release);
    }
    refresh_energies();                                                   // .cpp:95-96
    if (step % 1000 == 0 && std::getenv("LJGPU_VERBOSE")) {               // .cpp:100-113
        const auto end = std::chrono::system_clock::now();
        const std::chrono::duration<double> el = end - ttime;
        std::fprintf(stderr, "Step %04u  Ekin = %.6f  Epot = %.6f  Etot = %.6f  Time = %.4f s\n",
                     step, get_ekin(), get_epot(), get_etot(), el.count());
        ttime = std::chrono::system_clock::now();
    }
}

void LennardJonesSimulation::integrate_n(unsigned int first_step, unsigned int count, double dt) {
    if (count == 0) return;
    if (ljgpu_step_n(sim, first_step, count, dt) != LJGPU_OK) fail("ljgpu_step_n");
    steps_done.fetch_add(count, std::memory_order_release);
    refresh_energies();
    if (mirror == Mirror::Eager) refresh_mirror();
}

int LennardJonesSimulation::get_render_index() const {
    refresh_mirror();          // lazy: copy on demand; eager: wait for the delivery in flight, so the newest step shows
    return render_idx.load(std::memory_order_acquire);
}
const double* LennardJonesSimulation::get_render_x() const { return mirror_buf[get_render_index()]; }
const double* LennardJonesSimulation::get_render_y() const { return mirror_buf[render_idx.load(std::memory_order_acquire)] + n; }
const double* LennardJonesSimulation::get_render_z() const { return mirror_buf[render_idx.load(std::memory_order_acquire)] + 2 * n; }

std::vector<LennardJonesSimulation::dvec3> LennardJonesSimulation::get_positions_copy() const {
    refresh_mirror();
    const double* b = mirror_buf[render_idx.load(std::memory_order_acquire)];
    std::vector<dvec3> out(n);
    for (size_t i = 0; i < n; ++i) out[i] = dvec3(b[i], b[n + i], b[2 * n + i]);     // .cpp:165-169
    return out;
}

std::vector<LennardJonesSimulation::dvec3> LennardJonesSimulation::get_velocities_copy() const {
    std::vector<double> v(3 * n);
    if (ljgpu_get_velocities(sim, &v[0], &v[n], &v[2 * n]) != LJGPU_OK) fail("ljgpu_get_velocities");
    std::vector<dvec3> out(n);
    for (size_t i = 0; i < n; ++i) out[i] = dvec3(v[i], v[n + i], v[2 * n + i]);     // .cpp:175-178
    return out;
}

void LennardJonesSimulation::write_to_movie_file(const std::string& moviefile, bool create) {
    if (ljgpu_write_movie_frame(sim, moviefile.c_str(), create ? 1 : 0) != LJGPU_OK) fail("ljgpu_write_movie_frame");
}

std::vector<uint64_t> LennardJonesSimulation::get_speed_histogram(unsigned nbins, double maxspeed) const {
    std::vector<uint64_t> h(nbins);
    if (ljgpu_get_speed_histogram(sim, nbins, maxspeed, h.data()) != LJGPU_OK) fail("ljgpu_get_speed_histogram");
    return h;
}

void LennardJonesSimulation::reload_params() {
    const ljgpu_params s = to_struct(*params, kernel_choice, device_choice);
    if (ljgpu_update_params(sim, &s) != LJGPU_OK) fail("ljgpu_update_params");
    refresh_energies();
}
