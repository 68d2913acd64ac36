// dvec3.h - the three-double vector the reference's getters return (glm::dvec3).  Define
// LJGPU_USE_GLM to use glm itself when building inside the reference's GUI tree.
#pragma once
#ifdef LJGPU_USE_GLM
#include <glm/glm.hpp>
namespace ljgpu_host { using dvec3 = glm::dvec3; }
#else
namespace ljgpu_host {
struct dvec3 {
    double x, y, z;
    dvec3() : x(0), y(0), z(0) {}
    dvec3(double a, double b, double c) : x(a), y(b), z(c) {}
    double& operator[](int i) { return (&x)[i]; }
    const double& operator[](int i) const { return (&x)[i]; }
};
}
#endif
