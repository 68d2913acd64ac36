// Minimal stand-in for glm's dvec3 (lennardjonessimulation.h:67, .cpp:37,141,168): three
// contiguous doubles with (x,y,z) ctor, .x/.y/.z and operator[]. TEST INFRASTRUCTURE ONLY.
#pragma once
namespace glm {
struct dvec3 {
    double x, y, z;
    dvec3() : x(0), y(0), z(0) {}
    dvec3(double a, double b, double c) : x(a), y(b), z(c) {}
    double& operator[](int i) { return (&x)[i]; }
    const double& operator[](int i) const { return (&x)[i]; }
};
}
