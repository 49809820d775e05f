// ptb200::Scene — the drop-in for dod::Scene (src/dod/Scene.h:21-57).
//
// It satisfies the duck-typed SceneBuilder concept every createXScene<SB> recipe of the
// reference uses (addTriangle / addSphere / setEnvironmentColour, src/main/main.cpp:45-309)
// and offers the same render(camera, renderParams, updateFunc) -> ArrayOutput entry point
// (Scene.h:44-46), but owns nothing except flat host arrays: render() marshals them through
// the C ABI (include/ptb200.h) to the sm_100a kernels.  The visible-for-tests intersect
// functions (Scene.h:48-56) are forwarded the same way.  C ABI failures become
// std::runtime_error; there is no CPU fallback.
#pragma once

#include "ArrayOutput.h"
#include "Camera.h"
#include "MaterialSpec.h"
#include "RenderParams.h"
#include "Vec3.h"
#include "ptb200.h"

#include <cstring>
#include <functional>
#include <limits>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace ptb200 {

struct Ray { // src/math/Ray.h:5-28
  Vec3 origin;
  Norm3 direction;
  static Ray fromTwoPoints(const Vec3 &a, const Vec3 &b) { return {a, (b - a).normalised()}; }
};

struct Hit { // src/math/Hit.h:6-11
  double distance{};
  bool inside{};
  Vec3 position;
  Vec3 normal;
};

struct IntersectionRecord { // src/dod/IntersectionRecord.h:8-11 (material by value here)
  Hit hit;
  MaterialSpec material;
};

class Scene {
  std::vector<double> triangleVertices_;  // T x 9
  std::vector<uint32_t> triangleMaterial_;
  std::vector<double> sphereCentreRadius_; // S x 4
  std::vector<uint32_t> sphereMaterial_;
  std::vector<MaterialSpec> palette_;
  Vec3 environment_;
  PtRenderOptions options_{};
  PtStats lastStats_{};
  bool useAllDevices_{false};
  std::vector<int32_t> deviceList_; // explicit devices of a multi-GPU render (empty: options_.device)

  uint32_t intern(const MaterialSpec &material) {
    for (size_t i = palette_.size(); i-- > 0;) // recently used materials are at the back
      if (palette_[i] == material)
        return static_cast<uint32_t>(i);
    palette_.push_back(material);
    return static_cast<uint32_t>(palette_.size() - 1);
  }

  static void check(int code, const char *what) {
    if (code != PTB200_OK)
      throw std::runtime_error(std::string(what) + ": " + ptb200_last_error());
  }

  struct Marshalled {
    std::vector<PtMaterial> materials;
    PtScene scene{};
  };
  [[nodiscard]] Marshalled marshal() const {
    Marshalled m;
    for (const auto &mat : palette_)
      m.materials.push_back(mat.abi());
    m.scene.numTriangles = static_cast<uint32_t>(triangleMaterial_.size());
    m.scene.numSpheres = static_cast<uint32_t>(sphereMaterial_.size());
    m.scene.numMaterials = static_cast<uint32_t>(m.materials.size());
    m.scene.triangleVertices = triangleVertices_.data();
    m.scene.triangleMaterial = triangleMaterial_.data();
    m.scene.sphereCentreRadius = sphereCentreRadius_.data();
    m.scene.sphereMaterial = sphereMaterial_.data();
    m.scene.materials = m.materials.data();
    m.scene.environment[0] = environment_.x();
    m.scene.environment[1] = environment_.y();
    m.scene.environment[2] = environment_.z();
    return m;
  }

  [[nodiscard]] std::optional<IntersectionRecord> forwardIntersect(int which, const Ray &ray,
                                                                   double nearerThan) const {
    const auto m = marshal();
    const double packed[6] = {ray.origin.x(),    ray.origin.y(),    ray.origin.z(),
                              ray.direction.x(), ray.direction.y(), ray.direction.z()};
    PtHit hit{};
    check(ptb200_intersect(&m.scene, options_.device, which, nearerThan, 1, packed, &hit),
          "ptb200_intersect");
    if (!hit.hit)
      return std::nullopt;
    return IntersectionRecord{Hit{hit.distance, hit.inside != 0,
                                  Vec3(hit.position[0], hit.position[1], hit.position[2]),
                                  Vec3(hit.normal[0], hit.normal[1], hit.normal[2])},
                              palette_[static_cast<size_t>(hit.material)]};
  }

public:
  // --- SceneBuilder concept (src/dod/Scene.h:37-42) --------------------------------------
  void addTriangle(const Vec3 &v0, const Vec3 &v1, const Vec3 &v2, const MaterialSpec &material) {
    for (const Vec3 *v : {&v0, &v1, &v2}) {
      triangleVertices_.push_back(v->x());
      triangleVertices_.push_back(v->y());
      triangleVertices_.push_back(v->z());
    }
    triangleMaterial_.push_back(intern(material));
  }
  void addSphere(const Vec3 &centre, double radius, const MaterialSpec &material) {
    sphereCentreRadius_.insert(sphereCentreRadius_.end(),
                               {centre.x(), centre.y(), centre.z(), radius});
    sphereMaterial_.push_back(intern(material));
  }
  void setEnvironmentColour(const Vec3 &colour) { environment_ = colour; }

  // --- backend knobs (no reference equivalent) -------------------------------------------
  void setRngMode(int mode) { options_.rngMode = mode; }
  void setDevice(int device) { options_.device = device; }
  void setUseAllDevices(bool all) { useAllDevices_ = all; deviceList_.clear(); }
  // Render on exactly these CUDA ordinals (row-partitioned framebuffer, ptb200_render_multi).
  void setDevices(std::vector<int32_t> devices) { deviceList_ = std::move(devices); useAllDevices_ = false; }
  void setPassesPerBatch(int passes) { options_.passesPerBatch = passes; }
  // Exact-stream policies: lanes that share one pass (4, 8, 16, 32; 0 = chosen from the pass count).
  void setLanesPerPass(int lanes) { options_.lanesPerPass = lanes; }
  [[nodiscard]] const PtStats &lastStats() const noexcept { return lastStats_; }
  [[nodiscard]] size_t numTriangles() const noexcept { return triangleMaterial_.size(); }
  [[nodiscard]] size_t numSpheres() const noexcept { return sphereMaterial_.size(); }
  [[nodiscard]] const std::vector<MaterialSpec> &palette() const noexcept { return palette_; }
  [[nodiscard]] const std::vector<double> &triangleVertices() const noexcept {
    return triangleVertices_;
  }
  [[nodiscard]] const std::vector<uint32_t> &triangleMaterials() const noexcept {
    return triangleMaterial_;
  }
  [[nodiscard]] const std::vector<double> &sphereCentreRadius() const noexcept {
    return sphereCentreRadius_;
  }
  [[nodiscard]] const std::vector<uint32_t> &sphereMaterials() const noexcept {
    return sphereMaterial_;
  }
  [[nodiscard]] const Vec3 &environment() const noexcept { return environment_; }

  // --- dod::Scene::render (src/dod/Scene.h:44-46, Scene.cpp:198-254) ---------------------
  // updateFunc is invoked on the calling thread after each collected batch of passes, as the
  // reference does after each collected pass (Scene.cpp:242-245).
  ArrayOutput render(const Camera &camera, const RenderParams &renderParams,
                     const std::function<void(ArrayOutput &output)> &updateFunc) {
    const auto m = marshal();
    const PtRenderParams params = renderParams.abi();
    const size_t pixels = static_cast<size_t>(renderParams.width) * renderParams.height;
    std::vector<PtPixel> raw(pixels);

    struct Forward {
      const std::function<void(ArrayOutput &)> *fn;
      int width, height;
    } forward{&updateFunc, renderParams.width, renderParams.height};
    auto trampoline = [](void *user, const PtPixel *soFar, int32_t, int32_t) -> int {
      auto *f = static_cast<Forward *>(user);
      ArrayOutput partial(f->width, f->height);
      partial.addSamples(soFar);
      (*f->fn)(partial);
      return 0;
    };

    const bool multi = useAllDevices_ || deviceList_.size() > 1;
    if (multi) {
      check(ptb200_render_multi_progress(&m.scene, &camera.abi(), &params, &options_,
                                         deviceList_.empty() ? nullptr : deviceList_.data(),
                                         static_cast<int32_t>(deviceList_.size()), raw.data(),
                                         updateFunc ? +trampoline : nullptr, &forward, &lastStats_),
            "ptb200_render_multi_progress");
    } else {
      check(ptb200_render(&m.scene, &camera.abi(), &params, &options_, raw.data(),
                          updateFunc ? +trampoline : nullptr, &forward, &lastStats_),
            "ptb200_render");
    }
    ArrayOutput output(renderParams.width, renderParams.height);
    output.addSamples(raw.data());
    return output;
  }

  // --- visible for tests (src/dod/Scene.h:48-56) -----------------------------------------
  [[nodiscard]] std::optional<IntersectionRecord> intersectSpheres(const Ray &ray,
                                                                   double nearerThan) const {
    return forwardIntersect(1, ray, nearerThan);
  }
  [[nodiscard]] std::optional<IntersectionRecord> intersectTriangles(const Ray &ray,
                                                                     double nearerThan) const {
    return forwardIntersect(2, ray, nearerThan);
  }
  [[nodiscard]] std::optional<IntersectionRecord> intersect(const Ray &ray) const {
    return forwardIntersect(0, ray, std::numeric_limits<double>::infinity());
  }
};

} // namespace ptb200
