// Camera of the host boundary.  Builds the same 18 doubles the reference's Camera holds
// privately (src/math/Camera.h:11-18, constructor :40-46, setFocus :48-51) and exposes them
// as the C ABI's PtCamera; ray generation itself (Camera.h:20-37,54-60) is device code.
#pragma once

#include "Vec3.h"
#include "ptb200.h"

#include <cmath>

namespace ptb200 {

class Camera {
  PtCamera abi_{};

  static void put(double (&dst)[3], const Vec3 &v) {
    dst[0] = v.x();
    dst[1] = v.y();
    dst[2] = v.z();
  }

public:
  Camera(const Vec3 &eye, const Vec3 &lookAt, const Norm3 &up, int width, int height,
         double verticalFovDegrees) {
    // OrthoNormalBasis::fromZY(z, y): x = normalise(y cross z), y' = z cross x
    // (src/math/OrthoNormalBasis.cpp:34-38).
    const Norm3 zAxis = (lookAt - eye).normalised();
    const Norm3 xAxis = up.cross(zAxis).normalised();
    const Vec3 yAxis = zAxis.cross(xAxis);
    put(abi_.centre, eye);
    put(abi_.axisX, xAxis.toVec3());
    put(abi_.axisY, yAxis);
    put(abi_.axisZ, zAxis.toVec3());
    abi_.aspectRatio = static_cast<double>(width) / height;
    abi_.cameraPlaneDist = 1.0 / std::tan(verticalFovDegrees * M_PI / 360.0);
    abi_.reciprocalHeight = 1.0 / height;
    abi_.reciprocalWidth = 1.0 / width;
    abi_.apertureRadius = 0.0;
    abi_.focalDistance = 0.0;
  }

  void setFocus(const Vec3 &focalPoint, double apertureRadius) {
    const Vec3 centre(abi_.centre[0], abi_.centre[1], abi_.centre[2]);
    abi_.focalDistance = (focalPoint - centre).length();
    abi_.apertureRadius = apertureRadius;
  }

  [[nodiscard]] const PtCamera &abi() const noexcept { return abi_; }

  // A camera whose 18 doubles are already known (e.g. stored next to a scene fixture).
  static Camera fromAbi(const PtCamera &state) {
    Camera camera(Vec3(0, 0, 0), Vec3(0, 0, 1), Vec3(0, 1, 0).normalised(), 1, 1, 40.0);
    camera.abi_ = state;
    return camera;
  }
};

} // namespace ptb200
