// Host-side value types of the B200 backend.
//
// The reference keeps Vec3/Norm3 (src/math/Vec3.h:8-107, src/math/Norm3.h:7-49) as host C++
// and so do we: these are this repository's own small equivalents with the same member names,
// so scene recipes and the SceneBuilder concept read the same on either side.  Nothing here
// runs per ray; the per-ray arithmetic lives in csrc/ as device code.
#pragma once

#include <cmath>
#include <ostream>

namespace ptb200 {

class Norm3;

class Vec3 {
  double c_[3]{0.0, 0.0, 0.0};

public:
  constexpr Vec3() noexcept = default;
  constexpr Vec3(double x, double y, double z) noexcept : c_{x, y, z} {}

  [[nodiscard]] constexpr double x() const noexcept { return c_[0]; }
  [[nodiscard]] constexpr double y() const noexcept { return c_[1]; }
  [[nodiscard]] constexpr double z() const noexcept { return c_[2]; }
  [[nodiscard]] constexpr const double *data() const noexcept { return c_; }

  constexpr Vec3 operator+(const Vec3 &o) const noexcept {
    return {c_[0] + o.c_[0], c_[1] + o.c_[1], c_[2] + o.c_[2]};
  }
  constexpr Vec3 operator-(const Vec3 &o) const noexcept {
    return {c_[0] - o.c_[0], c_[1] - o.c_[1], c_[2] - o.c_[2]};
  }
  constexpr Vec3 operator-() const noexcept { return {-c_[0], -c_[1], -c_[2]}; }
  constexpr Vec3 operator*(double s) const noexcept { return {c_[0] * s, c_[1] * s, c_[2] * s}; }
  constexpr Vec3 operator*(const Vec3 &o) const noexcept {
    return {c_[0] * o.c_[0], c_[1] * o.c_[1], c_[2] * o.c_[2]};
  }
  // The reference divides by multiplying with the reciprocal (Vec3.h:51-54); keep that.
  constexpr Vec3 operator/(double s) const noexcept { return *this * (1.0 / s); }
  constexpr Vec3 &operator+=(const Vec3 &o) noexcept { return *this = *this + o; }
  constexpr Vec3 &operator-=(const Vec3 &o) noexcept { return *this = *this - o; }
  constexpr Vec3 &operator*=(double s) noexcept { return *this = *this * s; }
  friend constexpr Vec3 operator*(double s, const Vec3 &v) noexcept { return v * s; }

  constexpr bool operator==(const Vec3 &o) const noexcept {
    return c_[0] == o.c_[0] && c_[1] == o.c_[1] && c_[2] == o.c_[2];
  }
  constexpr bool operator!=(const Vec3 &o) const noexcept { return !(*this == o); }

  [[nodiscard]] constexpr double dot(const Vec3 &o) const noexcept {
    return c_[0] * o.c_[0] + c_[1] * o.c_[1] + c_[2] * o.c_[2];
  }
  [[nodiscard]] constexpr Vec3 cross(const Vec3 &o) const noexcept {
    return {c_[1] * o.c_[2] - c_[2] * o.c_[1], c_[2] * o.c_[0] - c_[0] * o.c_[2],
            c_[0] * o.c_[1] - c_[1] * o.c_[0]};
  }
  [[nodiscard]] constexpr double lengthSquared() const noexcept { return dot(*this); }
  [[nodiscard]] double length() const noexcept { return std::sqrt(lengthSquared()); }
  [[nodiscard]] inline Norm3 normalised() const noexcept;

  static constexpr Vec3 xAxis() noexcept { return {1, 0, 0}; }
  static constexpr Vec3 yAxis() noexcept { return {0, 1, 0}; }
  static constexpr Vec3 zAxis() noexcept { return {0, 0, 1}; }
};

// A Vec3 known to have unit length (src/math/Norm3.h:7-49).
class Norm3 {
  Vec3 v_{1, 0, 0};
  friend class Vec3;
  constexpr explicit Norm3(const Vec3 &v) noexcept : v_(v) {}

public:
  constexpr Norm3() noexcept = default;
  // For callers that already hold a unit vector (Norm3::fromNormal, Norm3.impl.h:31-34).
  static constexpr Norm3 fromNormal(const Vec3 &unit) noexcept { return Norm3(unit); }
  [[nodiscard]] constexpr const Vec3 &toVec3() const noexcept { return v_; }
  [[nodiscard]] constexpr double x() const noexcept { return v_.x(); }
  [[nodiscard]] constexpr double y() const noexcept { return v_.y(); }
  [[nodiscard]] constexpr double z() const noexcept { return v_.z(); }
  constexpr Norm3 operator-() const noexcept { return Norm3(-v_); }
  constexpr Vec3 operator*(double s) const noexcept { return v_ * s; }
  [[nodiscard]] constexpr double dot(const Norm3 &o) const noexcept { return v_.dot(o.v_); }
  [[nodiscard]] constexpr double dot(const Vec3 &o) const noexcept { return v_.dot(o); }
  [[nodiscard]] constexpr Vec3 cross(const Norm3 &o) const noexcept { return v_.cross(o.v_); }
  [[nodiscard]] constexpr Vec3 cross(const Vec3 &o) const noexcept { return v_.cross(o); }
  constexpr bool operator==(const Norm3 &o) const noexcept { return v_ == o.v_; }
  static constexpr Norm3 xAxis() noexcept { return Norm3(Vec3::xAxis()); }
  static constexpr Norm3 yAxis() noexcept { return Norm3(Vec3::yAxis()); }
  static constexpr Norm3 zAxis() noexcept { return Norm3(Vec3::zAxis()); }
};

inline Norm3 Vec3::normalised() const noexcept { return Norm3(*this / length()); }

inline std::ostream &operator<<(std::ostream &o, const Vec3 &v) {
  return o << '{' << v.x() << ", " << v.y() << ", " << v.z() << '}';
}
inline std::ostream &operator<<(std::ostream &o, const Norm3 &v) { return o << v.toVec3(); }

} // namespace ptb200
