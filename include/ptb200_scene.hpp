// b200::Scene — the reference-side binding of libptb200.so: one more "way" for
// mattgodbolt/pt-three-ways, with the interface of dod::Scene (src/dod/Scene.h:33-46).
//
// This header is written against the REFERENCE's own types (Vec3, MaterialSpec, Camera,
// RenderParams, ArrayOutput) and is meant to be dropped into the reference tree as
// src/b200/Scene.h; it needs <reference>/src on the include path and is
// therefore not used by this repository's own host side (pt_three_ways_b200/host/Scene.h is the
// same adaptor over this repository's stand-in types).  oracle/b200_dropin.cpp compiles it
// against the reference tree where that is mounted, driven by the reference's own
// createXScene<SB> recipes (src/main/main.cpp:69-309) exactly as doRender does (:360-363).
#pragma once

#include "math/Camera.h"
#include "util/ArrayOutput.h"
#include "util/MaterialSpec.h"
#include "util/RenderParams.h"

#include <ptb200.h>

#include <cstdint>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <vector>

namespace b200 {

class Scene {
  std::vector<double> triangleVertices_;   // T x 9, as addTriangle receives them
  std::vector<uint32_t> triangleMaterial_;
  std::vector<double> sphereCentreRadius_; // S x 4
  std::vector<uint32_t> sphereMaterial_;
  std::vector<PtMaterial> palette_;        // the reference stores one MaterialSpec per primitive
  Vec3 environment_;
  PtRenderOptions options_{};              // keyed RNG, device 0 (include/ptb200.h)
  PtStats stats_{};

  uint32_t intern(const MaterialSpec &m) {
    const PtMaterial p{{m.emission.x(), m.emission.y(), m.emission.z()},
                       {m.diffuse.x(), m.diffuse.y(), m.diffuse.z()},
                       m.indexOfRefraction,
                       m.reflectivity,
                       m.reflectionConeAngleRadians};
    for (size_t i = palette_.size(); i-- > 0;)
      if (std::memcmp(&palette_[i], &p, sizeof p) == 0)
        return static_cast<uint32_t>(i);
    palette_.push_back(p);
    return static_cast<uint32_t>(palette_.size() - 1);
  }

  static void fill(ArrayOutput &output, const PtPixel *pixels) {
    for (int y = 0; y < output.height(); ++y)
      for (int x = 0; x < output.width(); ++x) {
        const PtPixel &q = pixels[x + static_cast<size_t>(y) * output.width()];
        output.addSamples(x, y, Vec3(q.sum[0], q.sum[1], q.sum[2]), static_cast<int>(q.numSamples));
      }
  }

public:
  // ---- the SceneBuilder concept (src/dod/Scene.h:37-42) ----
  void addTriangle(const Vec3 &v0, const Vec3 &v1, const Vec3 &v2, const MaterialSpec &material) {
    for (const Vec3 *v : {&v0, &v1, &v2})
      triangleVertices_.insert(triangleVertices_.end(), {v->x(), v->y(), v->z()});
    triangleMaterial_.push_back(intern(material));
  }
  void addSphere(const Vec3 &centre, double radius, const MaterialSpec &material) {
    sphereCentreRadius_.insert(sphereCentreRadius_.end(), {centre.x(), centre.y(), centre.z(), radius});
    sphereMaterial_.push_back(intern(material));
  }
  void setEnvironmentColour(const Vec3 &colour) { environment_ = colour; }

  // Backend-specific knobs (RNG policy, device, row partition); defaults need no call.
  void setOptions(const PtRenderOptions &options) { options_ = options; }
  [[nodiscard]] const PtStats &lastStats() const { return stats_; }

  [[nodiscard]] PtScene abi() const {
    PtScene s{};
    s.numTriangles = static_cast<uint32_t>(triangleMaterial_.size());
    s.numSpheres = static_cast<uint32_t>(sphereMaterial_.size());
    s.numMaterials = static_cast<uint32_t>(palette_.size());
    s.triangleVertices = triangleVertices_.data();
    s.triangleMaterial = triangleMaterial_.data();
    s.sphereCentreRadius = sphereCentreRadius_.data();
    s.sphereMaterial = sphereMaterial_.data();
    s.materials = palette_.data();
    s.environment[0] = environment_.x();
    s.environment[1] = environment_.y();
    s.environment[2] = environment_.z();
    return s;
  }

  // Camera keeps its 18 doubles private (src/math/Camera.h:11-18); the layout is plain.
  [[nodiscard]] static PtCamera abi(const Camera &camera) {
    static_assert(sizeof(Camera) == sizeof(PtCamera), "Camera layout changed");
    PtCamera c;
    std::memcpy(&c, &camera, sizeof c);
    return c;
  }

  // ---- dod::Scene::render (src/dod/Scene.h:44-46) ----
  [[nodiscard]] ArrayOutput render(const Camera &camera, const RenderParams &renderParams,
                                   const std::function<void(ArrayOutput &)> &updateFunc) {
    const PtScene s = abi();
    const PtCamera cam = abi(camera);
    const PtRenderParams p{renderParams.width,           renderParams.height,
                           renderParams.preview,         renderParams.samplesPerPixel,
                           renderParams.maxCpus,         renderParams.maxDepth,
                           renderParams.firstBounceUSamples, renderParams.firstBounceVSamples,
                           renderParams.seed};
    std::vector<PtPixel> pixels(static_cast<size_t>(renderParams.width) * renderParams.height);

    struct Trampoline {
      const std::function<void(ArrayOutput &)> *updateFunc;
      int width, height;
    } trampoline{&updateFunc, renderParams.width, renderParams.height};
    const PtProgressFn progress = [](void *user, const PtPixel *px, int32_t, int32_t) -> int {
      auto *t = static_cast<Trampoline *>(user);
      ArrayOutput partial(t->width, t->height);
      fill(partial, px);
      (*t->updateFunc)(partial); // on the calling thread, as Scene.cpp:245
      return 0;
    };
    if (ptb200_render(&s, &cam, &p, &options_, pixels.data(), updateFunc ? progress : nullptr,
                      &trampoline, &stats_) != PTB200_OK)
      throw std::runtime_error(ptb200_last_error()); // no exception crosses the C ABI

    ArrayOutput output(renderParams.width, renderParams.height);
    fill(output, pixels.data());
    return output;
  }
};

} // namespace b200
