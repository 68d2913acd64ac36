// threadintegrate.cpp - see threadintegrate.h; loop structure of threadintegrate.cpp:26-69.
#include "threadintegrate.h"

#include <chrono>

void ThreadIntegrate::start() {
    stop_requested.store(false);
    worker = std::thread([this] { run(); });
}

void ThreadIntegrate::run() {
    const double dt = ljsim->get_params()->get_param<double>("stepsize");
    iterator.store(0);
    local_iterator = 0;
    const bool throttled = ips > 0.0;
    const double steps_per_frame = throttled ? ips / fps : 1e300;
    const auto frame_duration = std::chrono::duration<double, std::milli>(1000.0 / fps);
    auto ttime = std::chrono::system_clock::now();

    while (!stop_requested.load()) {
        if (keeprunning.load()) {
            if (local_iterator < steps_per_frame) {
                ljsim->integrate(iterator.load(), dt);
                local_iterator++;
                iterator.fetch_add(1);
            }
            const auto end = std::chrono::system_clock::now();
            const std::chrono::duration<double, std::milli> elapsed_ms = end - ttime;
            if (elapsed_ms >= frame_duration) {
                ttime = std::chrono::system_clock::now();
                local_iterator = 0;
                if (signal_integration_step) signal_integration_step();
            } else if (local_iterator >= steps_per_frame) {
                std::this_thread::sleep_for(std::chrono::milliseconds(1));
            }
            const unsigned int it = iterator.load();
            if (it % 1000 == 0) {
                if (signal_ekin) signal_ekin(it * dt, ljsim->get_ekin());
                if (signal_epot) signal_epot(it * dt, ljsim->get_epot());
                if (signal_etot) signal_etot(it * dt, ljsim->get_etot());
                if (signal_velocities) signal_velocities();
            }
        } else {
            std::this_thread::sleep_for(std::chrono::milliseconds(50));    // paused
        }
    }
}

void ThreadIntegrate::stop() { stop_requested.store(true); keeprunning.store(false); }
void ThreadIntegrate::toggle_pause() { keeprunning.store(!keeprunning.load()); }
void ThreadIntegrate::wait() { if (worker.joinable()) worker.join(); }
