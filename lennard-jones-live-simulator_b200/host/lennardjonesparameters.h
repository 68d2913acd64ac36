// lennardjonesparameters.h - Qt-free stand-in for the reference's parameter map, same interface:
//   /root/reference/src/simulator/lennardjonesparameters.h:32-69
//   get_param<T>(name) throws std::runtime_error("Could not find parameter: <name>") for a missing key
//   (:52-58); set_param(name, value) inserts or overwrites (:66-68).  Values may be int, double or a
//   numeric string, because the reference's New-simulation dialog stores strings
//   (dialognewsimulation.cpp:73-78) and QVariant::value<T>() converts on read.
//
// LJGPU_WITH_QT: building inside the reference's Qt tree.  Then this header only forwards to the reference's OWN
// lennardjonesparameters.h (QMap<QString, QVariant>, get_param<T>(const QString&), set_param(const QString&, const
// QVariant&)), named by the build through -DLJGPU_QT_PARAMETERS_HEADER="<path>", so MainWindow (mainwindow.cpp:67-77)
// and DialogNewSimulation (dialognewsimulation.cpp:73-78) keep filling a QVariant map without edits; the adapter class
// needs nothing but get_param<double|int>(name), which both classes provide.  (tests/test_abi.py compiles the adapter
// this way against the reference's header.)
#pragma once
#if defined(LJGPU_WITH_QT)
#ifndef LJGPU_QT_PARAMETERS_HEADER
#error "LJGPU_WITH_QT needs -DLJGPU_QT_PARAMETERS_HEADER=\"<reference>/src/simulator/lennardjonesparameters.h\""
#endif
#include LJGPU_QT_PARAMETERS_HEADER
#else
#include <cmath>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <variant>

class LennardJonesParameters {
public:
    using Value = std::variant<int, double, std::string>;

    LennardJonesParameters() = default;

    template <typename T>
    T get_param(const std::string& name) const {
        auto it = params.find(name);
        if (it == params.end()) throw std::runtime_error("Could not find parameter: " + name);
        return convert<T>(it->second);
    }

    inline void set_param(const std::string& name, const Value& value) { params[name] = value; }
    inline void set_param(const std::string& name, const char* value) { params[name] = std::string(value); }

    // the shipped default configuration, src/gui/mainwindow.cpp:67-77
    static LennardJonesParameters default_configuration() {
        LennardJonesParameters p;
        p.set_param("cell_length", 14.938); p.set_param("nr_particles", 1000);
        p.set_param("kT", 1.0); p.set_param("mass", 1.0); p.set_param("sigma", 1.0); p.set_param("rcut", 2.5);
        p.set_param("epsilon", 1.0); p.set_param("tau", 0.5); p.set_param("shell", 0.4); p.set_param("stepsize", 0.001);
        return p;
    }

private:
    std::map<std::string, Value> params;

    template <typename T>
    static T convert(const Value& v) {
        if (const int* i = std::get_if<int>(&v)) return static_cast<T>(*i);
        if (const double* d = std::get_if<double>(&v)) {
            if constexpr (std::is_integral<T>::value) return static_cast<T>(std::llround(*d));   // QVariant rounds
            else return static_cast<T>(*d);
        }
        const std::string& s = std::get<std::string>(v);
        const double d = std::strtod(s.c_str(), nullptr);
        if constexpr (std::is_integral<T>::value) return static_cast<T>(std::llround(d));
        else return static_cast<T>(d);
    }
};
#endif  // LJGPU_WITH_QT
