// threadintegrate.h - headless driver with the behaviour of the reference's QThread-based
// ThreadIntegrate (/root/reference/src/simulator/threadintegrate.{h,cpp}): starts paused or running,
// at most `ips` integrate() calls per second delivered in frames of 1/fps seconds, a frame callback per
// frame (signal_integration_step), and every 1000 iterations the energies and a velocities callback
// (signal_ekin/epot/etot, signal_velocities; threadintegrate.cpp:57-62).  Qt signals become
// std::function callbacks, QThread becomes std::thread.
#pragma once
#include <atomic>
#include <functional>
#include <memory>
#include <thread>

#include "lennardjonessimulation.h"

class ThreadIntegrate {
public:
    explicit ThreadIntegrate(const std::shared_ptr<LennardJonesSimulation>& _ljsim) : ljsim(_ljsim) {}
    ~ThreadIntegrate() { stop(); wait(); }

    // 0 removes the reference's 2000 steps/s throttle (threadintegrate.cpp:31-33)
    void set_rate(double integrations_per_second, double frames_per_second = 60.0) { ips = integrations_per_second; fps = frames_per_second; }

    std::function<void()> signal_integration_step;
    std::function<void(double, double)> signal_ekin, signal_epot, signal_etot;
    std::function<void()> signal_velocities;

    void start();                                   // QThread::start
    void stop();                                    // threadintegrate.cpp:74-77
    void toggle_pause();                            // threadintegrate.cpp:82-85
    void wait();
    bool is_running() const { return keeprunning.load(); }
    unsigned int iterations() const { return iterator.load(); }

private:
    void run();                                     // threadintegrate.cpp:26-69
    std::shared_ptr<LennardJonesSimulation> ljsim;
    std::thread worker;
    std::atomic<bool> keeprunning{true};
    std::atomic<bool> stop_requested{false};
    std::atomic<unsigned int> iterator{0};
    unsigned int local_iterator = 0;
    double ips = 2000.0, fps = 60.0;
};
