// lj_host.cpp - host-side pieces of the C ABI that involve no device work.
//
// ljgpu_host_initialize restates LennardJonesSimulation::initialize() + get_gauss()
// (/root/reference/src/simulator/lennardjonessimulation.cpp:185-252) with the same std:: calls,
// so that a fresh process produces the reference's start state bit for bit (same libstdc++).
// Build without -march=native / with -ffp-contract=off: the reference's arithmetic has no FMA.
#include "../../include/ljgpu.h"

#include <cmath>
#include <mutex>
#include <random>
#include <string>

namespace {
std::mutex g_rng_mu;
// The reference keeps ONE generator + distribution for the whole process (function statics in
// get_gauss, .cpp:249-250): a second simulation continues the stream. Same here.
std::default_random_engine& generator() { static std::default_random_engine g; return g; }
std::normal_distribution<double>& distribution() { static std::normal_distribution<double> d(0.0, 1.0); return d; }
}

extern "C" {

void ljgpu_host_rng_reset(void) {
    std::lock_guard<std::mutex> lk(g_rng_mu);
    generator() = std::default_random_engine();
    distribution().reset();
}

int ljgpu_host_initialize(const ljgpu_params* p, double* x, double* y, double* z,
                          double* vx, double* vy, double* vz) {
    if (!p || !x || !y || !z || !vx || !vy || !vz || p->nr_particles < 1 || !(p->cell_length > 0) || !(p->mass > 0))
        return LJGPU_EINVAL;
    std::lock_guard<std::mutex> lk(g_rng_mu);
    const size_t n = (size_t)p->nr_particles;
    const double L = p->cell_length;

    // start lattice, .cpp:186-218: sites per axis from the mean spacing, x outermost, z innermost
    const double spacing_guess = std::pow(L * L * L / (double)n, 1.0 / 3.0);
    const size_t per_axis = (size_t)std::ceil(L / spacing_guess);
    const double h = L / (double)per_axis;
    size_t filled = 0;
    for (size_t a = 0; a < per_axis && filled < n; ++a)
        for (size_t b = 0; b < per_axis && filled < n; ++b)
            for (size_t c = 0; c < per_axis && filled < n; ++c) {
                x[filled] = (a + 0.5) * h;
                y[filled] = (b + 0.5) * h;
                z[filled] = (c + 0.5) * h;
                ++filled;
            }

    // velocities, .cpp:221-240: vx,vy,vz per atom in that draw order, then remove the mean
    const double width = std::sqrt(p->kT / p->mass);
    double mx = 0.0, my = 0.0, mz = 0.0;
    auto& gen = generator(); auto& dist = distribution();
    for (size_t i = 0; i < n; ++i) {
        vx[i] = width * dist(gen);
        vy[i] = width * dist(gen);
        vz[i] = width * dist(gen);
        mx += vx[i]; my += vy[i]; mz += vz[i];
    }
    const double inv = 1.0 / (double)n;
    mx *= inv; my *= inv; mz *= inv;
    for (size_t i = 0; i < n; ++i) { vx[i] -= mx; vy[i] -= my; vz[i] -= mz; }
    return LJGPU_OK;
}

// GraphWidget::set_params, /root/reference/src/simulator/graphwidget.cpp:55-73.
int ljgpu_maxwell_boltzmann(double kT, double mass, int32_t nr_particles, uint32_t nbins,
                            double vmax, double* expected) {
    if (!expected || nbins == 0 || !(vmax > 0) || !(kT > 0) || !(mass > 0)) return LJGPU_EINVAL;
    const double binsize = vmax / (double)nbins;
    for (uint32_t i = 0; i < nbins; ++i) {
        const double v = binsize * i;
        const double f = 4.0 * M_PI * std::pow(mass / (2.0 * M_PI * kT), 3. / 2.) * v * v * std::exp(-mass * v * v / (2.0 * kT));
        expected[i] = f * nr_particles * binsize;
    }
    return LJGPU_OK;
}

} // extern "C"
