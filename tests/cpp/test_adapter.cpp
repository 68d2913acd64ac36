// tests/cpp/test_adapter.cpp - exercises the C++ drop-in class the way the reference's callers do
// (MainWindow::build_simulation, mainwindow.cpp:227-245; ThreadIntegrate::run, threadintegrate.cpp:26-69)
// and checks it against the CPU oracle (oracle/lj_oracle.h - test infrastructure).
//   usage: test_adapter gpu        full run on cuda:0, compared with the oracle
//          test_adapter nogpu      must throw std::runtime_error("... no CUDA device ...")
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <vector>

#include "graphstats.h"
#include "lennardjonessimulation.h"
#include "threadintegrate.h"
#include "lj_oracle.h"
#include "ljgpu.h"

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); ++failures; } } while (0)

static double rel(double a, double b) { return std::fabs(a - b) / std::fmax(std::fabs(b), 1e-300); }

int main(int argc, char** argv) {
    const bool want_gpu = argc > 1 && std::strcmp(argv[1], "gpu") == 0;

    auto params = std::make_shared<LennardJonesParameters>();
    params->set_param("cell_length", 14.938);  params->set_param("nr_particles", 1000);
    params->set_param("kT", 1.0);  params->set_param("mass", 1.0);  params->set_param("sigma", 1.0);
    params->set_param("rcut", 2.5);  params->set_param("epsilon", 1.0);  params->set_param("tau", 0.5);
    params->set_param("shell", 0.4);  params->set_param("stepsize", "0.001");     // string, as the dialog stores it

    // missing key -> the reference's exception text (lennardjonesparameters.h:52-58)
    try { LennardJonesParameters().get_param<double>("kT"); CHECK(false); }
    catch (const std::runtime_error& e) { CHECK(std::strcmp(e.what(), "Could not find parameter: kT") == 0); }

    if (!want_gpu) {
        try { LennardJonesSimulation sim(params); CHECK(false); }
        catch (const std::runtime_error& e) { CHECK(std::strstr(e.what(), "no CUDA device") != nullptr); }
        std::printf(failures ? "FAILED\n" : "OK nogpu\n");
        return failures ? 1 : 0;
    }

    ljgpu_host_rng_reset();
    auto sim = std::make_shared<LennardJonesSimulation>(params);
    ljo_rng_reset();
    ljo_params op{14.938, 1000, 1.0, 1.0, 1.0, 2.5, 1.0, 0.5, 0.4, 0.001};
    ljo_sim* o = ljo_create(&op);
    const size_t n = sim->get_particle_count();
    CHECK(n == 1000);
    CHECK(sim->get_dims()[0] == 14.938);
    CHECK(sim->get_ekin() == 0.0);                                  // 0 before the first step
    double ek, ep, et;
    ljo_get_energies(o, &ek, &ep, &et);
    CHECK(rel(sim->get_epot(), ep) < 1e-12);

    // start positions are the reference's lattice, bit for bit, through every accessor
    std::vector<double> ox(n), oy(n), oz(n);
    ljo_get_positions(o, ox.data(), oy.data(), oz.data());
    auto pos = sim->get_positions_copy();
    const double* rx = sim->get_render_x();
    bool same = true;
    for (size_t i = 0; i < n; ++i) same = same && pos[i].x == ox[i] && pos[i].z == oz[i] && rx[i] == ox[i] && sim->get_render_y()[i] == oy[i];
    CHECK(same);

    const double dt = params->get_param<double>("stepsize");
    for (unsigned s = 1; s <= 10; ++s) {
        sim->integrate(s, dt);
        ljo_integrate(o, s, dt);
        ljo_get_energies(o, &ek, &ep, &et);
        CHECK(rel(sim->get_ekin(), ek) < 1e-12);
        CHECK(rel(sim->get_epot(), ep) < 1e-12);
        CHECK(std::fabs(sim->get_etot() - et) < 1e-12 * (std::fabs(ek) + std::fabs(ep)));
    }
    ljo_get_positions(o, ox.data(), oy.data(), oz.data());
    rx = sim->get_render_x();
    double maxd = 0;
    for (size_t i = 0; i < n; ++i) maxd = std::fmax(maxd, std::fabs(rx[i] - ox[i]));
    CHECK(maxd < 1e-12);

    // histogram: device kernel == GraphWidget arithmetic on the velocities copy
    auto vel = sim->get_velocities_copy();
    auto h_dev = sim->get_speed_histogram();
    auto h_host = ljgpu_host::speed_histogram(vel);
    bool hsame = true; uint64_t tot = 0;
    for (unsigned b = 0; b < 100; ++b) { hsame = hsame && h_dev[b] == h_host[b]; tot += h_dev[b]; }
    CHECK(hsame && tot == n);
    auto mb = ljgpu_host::maxwell_boltzmann(1.0, 1.0, (int)n);
    double mbsum = 0; for (double v : mb) mbsum += v;
    CHECK(std::fabs(mbsum - (double)n) < 1.0);

    // live parameter change, as set_param on the shared map does in the reference
    params->set_param("tau", 1e300);
    sim->reload_params();
    ljo_set_tau(o, 1e300);
    sim->integrate(11, dt); ljo_integrate(o, 11, dt);
    ljo_get_energies(o, &ek, &ep, &et);
    CHECK(rel(sim->get_ekin(), ek) < 1e-12);

    // publication: after every integrate() the render buffers show that very step, in both mirror modes
    // (eager: the delivery in flight is completed on demand and rotates through three buffers; lazy: copied on demand)
    for (int mode = 0; mode < 2; ++mode) {
        sim->set_mirror_mode(mode ? LennardJonesSimulation::Mirror::Lazy : LennardJonesSimulation::Mirror::Eager);
        for (unsigned s = 12 + 5 * mode; s < 17 + 5 * mode; ++s) {
            sim->integrate(s, dt); ljo_integrate(o, s, dt);
            ljo_get_positions(o, ox.data(), oy.data(), oz.data());
            const double* px = sim->get_render_x();
            const double* py = sim->get_render_y();
            const double* pz = sim->get_render_z();
            double md = 0;
            for (size_t i = 0; i < n; ++i)
                md = std::fmax(md, std::fmax(std::fabs(px[i] - ox[i]), std::fmax(std::fabs(py[i] - oy[i]), std::fabs(pz[i] - oz[i]))));
            CHECK(md < 1e-11);
        }
    }
    sim->set_mirror_mode(LennardJonesSimulation::Mirror::Eager);

    // movie file: header + one frame
    sim->write_to_movie_file("/tmp/ljgpu_adapter_movie.bin", true);
    FILE* fp = std::fopen("/tmp/ljgpu_adapter_movie.bin", "rb");
    CHECK(fp != nullptr);
    if (fp) { std::fseek(fp, 0, SEEK_END); CHECK(std::ftell(fp) == 4 + 24 + (long)n * 56); std::fclose(fp); }

    // driver thread: unthrottled, frame + per-1000 callbacks, pause/stop
    ThreadIntegrate th(sim);
    th.set_rate(0.0);
    std::atomic<int> frames{0}, kilo{0};
    th.signal_integration_step = [&] { frames++; };
    th.signal_ekin = [&](double, double e) { if (e > 0) kilo++; };
    th.start();
    // 2500 iterations cross the per-1000-step callback twice; LJGPU_ADAPTER_TEST_ITERS shortens the run where a step
    // is expensive (the CPU emulation of the library, tests/test_emulated_library.py)
    const char* it_env = std::getenv("LJGPU_ADAPTER_TEST_ITERS");
    const unsigned long want_iters = it_env ? std::strtoul(it_env, nullptr, 10) : 2500ul;
    while (th.iterations() < want_iters) std::this_thread::sleep_for(std::chrono::milliseconds(1));
    th.toggle_pause();
    CHECK(!th.is_running());
    th.stop(); th.wait();
    CHECK(th.iterations() >= want_iters && frames.load() >= 1);
    if (want_iters >= 2500) CHECK(kilo.load() >= 2);

    auto st = ljgpu_host::series_stats({1.0, 2.0, 3.0, 4.0});
    CHECK(std::fabs(st.mean - 2.5) < 1e-15 && std::fabs(st.stddev - std::sqrt(5.0 / 3.0)) < 1e-15);

    ljo_destroy(o);
    std::printf(failures ? "FAILED\n" : "OK gpu\n");
    return failures ? 1 : 0;
}
