// lennardjonessimulation.h - drop-in host class with the public interface of the reference's
// LennardJonesSimulation (/root/reference/src/simulator/lennardjonessimulation.h:113-155), backed by
// libljgpu.so (include/ljgpu.h) instead of the CPU integrator.  MainWindow and ThreadIntegrate use only
// what is declared here (src/gui/mainwindow.cpp:211-245, src/simulator/threadintegrate.cpp:26-69).
//
// Errors surface as std::runtime_error, which the reference's LJSimApp::notify already catches and
// shows (src/ljsimapp.cpp:37-49).  There is no CPU fallback.
#pragma once
#include <atomic>
#include <chrono>
#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "dvec3.h"
#include "lennardjonesparameters.h"

struct ljgpu_sim;

class LennardJonesSimulation {
public:
    using dvec3 = ljgpu_host::dvec3;

    // When the host mirror of the published positions (get_render_x/y/z) is refreshed:
    //   Eager - every integrate() publishes (the reference's publish_state): the transfer of step k overlaps the
    //           computation of step k+1 (ljgpu_step_publish); a reader that asks earlier waits for it, so
    //           get_render_*() after integrate(k) always shows step k;
    //   Lazy  - on get_render_index() / get_positions_copy(), i.e. once per rendered frame;
    // The default is Eager up to 131072 atoms (GUI scale) and Lazy above.
    enum class Mirror { Eager, Lazy };

    // .h:113 - reads the ten keys (missing key -> std::runtime_error), builds the start state on the
    // host with the reference's own std:: calls, uploads it, bins, evaluates forces, publishes.
    explicit LennardJonesSimulation(const std::shared_ptr<LennardJonesParameters>& _params,
                                    int force_kernel = 0 /* LJGPU_KERNEL_AUTO */, int device = -1);
    ~LennardJonesSimulation();
    LennardJonesSimulation(const LennardJonesSimulation&) = delete;
    LennardJonesSimulation& operator=(const LennardJonesSimulation&) = delete;

    // .h:120 - one velocity-Verlet step; the log line every 1000 steps of .cpp:100-113 goes to stderr
    // when LJGPU_VERBOSE is set.
    void integrate(unsigned int step, double dt);
    // headless batch: n integrate() calls without returning in between
    void integrate_n(unsigned int first_step, unsigned int n, double dt);

    inline const auto& get_params() const { return this->params; }      // .h:123
    inline const auto& get_dims() const { return this->dims; }          // .h:126

    // ---------- rendering access (SoA), .h:131-143 ----------
    int get_render_index() const;
    const double* get_render_x() const;
    const double* get_render_y() const;
    const double* get_render_z() const;
    inline size_t get_particle_count() const { return n; }

    // .h:146-148
    inline double get_etot() const { return etot.load(std::memory_order_relaxed); }
    inline double get_ekin() const { return ekin.load(std::memory_order_relaxed); }
    inline double get_epot() const { return epot.load(std::memory_order_relaxed); }

    // .h:151-152
    std::vector<dvec3> get_positions_copy() const;
    std::vector<dvec3> get_velocities_copy() const;

    // .h:155
    void write_to_movie_file(const std::string& moviefile, bool create = false);

    // ---------- beyond the reference's surface ----------
    // GraphWidget::set_particle_speed (graphwidget.cpp:205-233) evaluated on the device.
    std::vector<uint64_t> get_speed_histogram(unsigned nbins = 100, double maxspeed = 10.0) const;
    // apply a change made through get_params()->set_param(...) to kT, tau, mass, sigma or epsilon
    // (the reference re-reads its map on every stage; this class reads it when told to)
    void reload_params();
    void set_mirror_mode(Mirror m) { mirror = m; }
    ljgpu_sim* handle() const { return sim; }

private:
    dvec3 dims;
    std::shared_ptr<LennardJonesParameters> params;
    size_t n = 0;
    ljgpu_sim* sim = nullptr;
    int kernel_choice = 0, device_choice = -1;

    // host mirror of the published positions: page-locked SoA buffers + atomic index (.h:76-84).  Three rather
    // than the reference's two: one is visible to the readers, one may be receiving the positions of the last
    // step (ljgpu_step_publish delivers while the next step computes), one is free for the next delivery.
    static constexpr int kMirrors = 3;
    mutable double* mirror_buf[kMirrors] = {nullptr, nullptr, nullptr};
    mutable std::atomic<int> render_idx{0};
    mutable int inflight_idx = -1;                     // buffer with a delivery in flight (guarded by mirror_mu)
    mutable uint64_t inflight_step = 0;                // steps_done value that delivery carries
    mutable std::atomic<uint64_t> mirrored_step{~0ull};
    mutable std::mutex mirror_mu;
    std::atomic<uint64_t> steps_done{0};
    Mirror mirror = Mirror::Eager;

    std::atomic<double> epot{0.0}, ekin{0.0}, etot{0.0};                   // .h:100-102
    std::chrono::time_point<std::chrono::system_clock> ttime;             // .h:105

    void refresh_mirror() const;
    void refresh_energies();
    [[noreturn]] static void fail(const char* what);
};
