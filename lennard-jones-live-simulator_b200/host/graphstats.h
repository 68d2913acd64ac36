// graphstats.h - the arithmetic of the reference's GraphWidget without the charts
// (/root/reference/src/simulator/graphwidget.cpp): speed histogram (:205-233), Maxwell-Boltzmann
// overlay (:55-73), and the running mean / standard-deviation bands of the energy series (:139-203,
// :235-248).  The reference seeds std::accumulate with the int 0, which truncates every partial sum
// (:147,169,191,236); `reference_truncation` reproduces that, the default does the sum in double.
#pragma once
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

#include "dvec3.h"

namespace ljgpu_host {

constexpr unsigned NUMBINS = 100;           // graphwidget.h:73
constexpr double MAXSPEED = 10.0;           // graphwidget.h:74

inline std::vector<unsigned> speed_histogram(const std::vector<dvec3>& velocities, unsigned nbins = NUMBINS, double maxspeed = MAXSPEED) {
    std::vector<unsigned> h(nbins, 0u);
    for (const auto& v : velocities) {
        const double speed = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
        int bin = (int)(speed / maxspeed * nbins);
        if (bin > (int)nbins - 1) bin = (int)nbins - 1;
        h[bin]++;
    }
    return h;
}

inline std::vector<double> maxwell_boltzmann(double temperature, double m, int numparts, unsigned nbins = NUMBINS, double maxspeed = MAXSPEED) {
    std::vector<double> out(nbins);
    const double binsize = maxspeed / (double)nbins;
    for (unsigned i = 0; i < nbins; i++) {
        const double v = binsize * i;
        const double f = 4.0 * M_PI * std::pow(m / (2.0 * M_PI * temperature), 3. / 2.) * v * v * std::exp(-m * v * v / (2.0 * temperature));
        out[i] = f * numparts * binsize;
    }
    return out;
}

struct SeriesStats { double mean, stddev; };

inline SeriesStats series_stats(const std::vector<double>& values, bool reference_truncation = false) {
    if (values.empty()) return {0.0, 0.0};
    double sum;
    if (reference_truncation) { int acc = 0; for (double v : values) acc = (int)(acc + v); sum = acc; }
    else sum = std::accumulate(values.begin(), values.end(), 0.0);
    const double avg = sum / (double)values.size();
    double var = 0.0;
    for (double v : values) var += (v - avg) * (v - avg);
    var /= (double)(values.size() - 1);                    // graphwidget.cpp:241
    if (!std::isfinite(var)) var = 0.0;                    // graphwidget.cpp:243-245
    return {avg, std::sqrt(var)};
}

} // namespace ljgpu_host
