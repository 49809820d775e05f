// MaterialSpec of the host boundary (mirrors src/util/MaterialSpec.h:7-40: same fields, same
// factory names, same defaults) and its conversion to the C ABI's PtMaterial.
#pragma once

#include "Vec3.h"
#include "ptb200.h"

namespace ptb200 {

struct MaterialSpec {
  Vec3 emission;
  Vec3 diffuse;
  double indexOfRefraction{1.0};
  double reflectivity{-1};
  double reflectionConeAngleRadians{0.0};

  static double toRadians(double degrees) { return degrees / 360 * 2 * M_PI; }
  static MaterialSpec makeDiffuse(const Vec3 &colour) { return {Vec3(), colour}; }
  static MaterialSpec makeSpecular(const Vec3 &colour, double index) {
    return {Vec3(), colour, index};
  }
  static MaterialSpec makeLight(const Vec3 &colour) { return {colour, Vec3()}; }
  static MaterialSpec makeGlossy(const Vec3 &colour, double index, double coneDegrees) {
    return {Vec3(), colour, index, -1, toRadians(coneDegrees)};
  }
  static MaterialSpec makeReflective(const Vec3 &colour, double reflectivity,
                                     double coneDegrees) {
    return {Vec3(), colour, 1.0, reflectivity, toRadians(coneDegrees)};
  }
  bool operator==(const MaterialSpec &o) const {
    return emission == o.emission && diffuse == o.diffuse &&
           indexOfRefraction == o.indexOfRefraction && reflectivity == o.reflectivity &&
           reflectionConeAngleRadians == o.reflectionConeAngleRadians;
  }
  bool operator!=(const MaterialSpec &o) const { return !(*this == o); }

  [[nodiscard]] PtMaterial abi() const {
    return PtMaterial{{emission.x(), emission.y(), emission.z()},
                      {diffuse.x(), diffuse.y(), diffuse.z()},
                      indexOfRefraction,
                      reflectivity,
                      reflectionConeAngleRadians};
  }
};

} // namespace ptb200
