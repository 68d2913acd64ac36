// TEST INFRASTRUCTURE: the adapter class compiled with -DLJGPU_WITH_QT against the REFERENCE's own
// lennardjonesparameters.h (Qt types from oracle/shim in this container, real Qt in the reference's build).
// Checks the contract the GUI relies on: values stored as int / double / QString are read through
// get_param<T>(QString), a missing key throws "Could not find parameter: <name>" from the constructor
// (lennardjonesparameters.h:52-58), and without a device the constructor fails loudly (no CPU fallback).
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>

#include "lennardjonessimulation.h"

int main() {
    auto p = std::make_shared<LennardJonesParameters>();          // the reference's class: QMap<QString, QVariant>
    p->set_param("cell_length", 14.938);                          // mainwindow.cpp:68-77 stores doubles and ints ...
    p->set_param("nr_particles", 1000);
    p->set_param("kT", QString("1.0"));                           // ... dialognewsimulation.cpp:75 stores strings
    for (const char* k : {"mass", "sigma", "epsilon"}) p->set_param(k, 1.0);
    p->set_param("rcut", 2.5); p->set_param("tau", 0.5); p->set_param("shell", 0.4);
    // "stepsize" is missing on purpose
    try {
        LennardJonesSimulation sim(p);
        std::puts("FAIL: constructor accepted a map without stepsize");
        return 1;
    } catch (const std::runtime_error& e) {
        if (!std::strstr(e.what(), "Could not find parameter: stepsize")) { std::printf("FAIL: %s\n", e.what()); return 1; }
    }
    p->set_param("stepsize", 0.001);
    if (p->get_param<int>("nr_particles") != 1000 || p->get_param<double>("kT") != 1.0) { std::puts("FAIL: conversions"); return 1; }
    try {
        LennardJonesSimulation sim(p);                            // needs a device from here on
        std::printf("QT SWITCH OK (device run: %zu atoms)\n", sim.get_particle_count());
    } catch (const std::runtime_error& e) {
        std::printf("QT SWITCH OK (no device: %s)\n", e.what());  // loud failure, not a CPU path
    }
    return 0;
}
