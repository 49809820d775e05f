// The seven built-in scenes of the reference's CLI (src/main/main.cpp:40-309), restated as
// templates over an "Api" so the same recipe text can be instantiated with this repository's
// host types (ptb200::HostApi, below) or — in oracle/ref_tool.cpp, test infrastructure — with
// the reference's own Vec3/MaterialSpec/Camera, which is how the recipes are cross-checked.
//
// An Api provides: types Vec3, MaterialSpec, Camera; `Norm3-like unit(Vec3)` to normalise;
// and `void loadObj(const char *fileName, SceneBuilder &)` for the three OBJ-backed scenes.
#pragma once

#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>

namespace ptb200 {

template <typename Api>
struct SceneRecipes {
  using Vec3 = typename Api::Vec3;
  using MaterialSpec = typename Api::MaterialSpec;
  using Camera = typename Api::Camera;

  static Camera lookFrom(const Vec3 &eye, const Vec3 &at, const Vec3 &up, int width,
                         int height, double fovDegrees) {
    return Camera(eye, at, Api::unit(up), width, height, fovDegrees);
  }

  // hexColour (main.cpp:40-43): 0xRRGGBB with a 2.2 gamma.
  static Vec3 hexColour(uint32_t hex) {
    auto channel = [](unsigned v) { return std::pow((v & 0xffu) / 255.0, 2.2); };
    return Vec3(channel(hex >> 16u), channel(hex >> 8u), channel(hex));
  }

  // addCube (main.cpp:45-67): 12 triangles; a set bit selects the LOW coordinate.
  template <typename SB>
  static void addCube(SB &sb, const Vec3 &low, const Vec3 &high, const MaterialSpec &mat) {
    auto corner = [&](unsigned bits) {
      return Vec3((bits & 4u) ? low.x() : high.x(), (bits & 2u) ? low.y() : high.y(),
                  (bits & 1u) ? low.z() : high.z());
    };
    static constexpr unsigned faces[12][3] = {
        {0, 4, 6}, {0, 6, 2}, {1, 5, 7}, {1, 7, 3}, {0, 4, 5}, {0, 5, 1},
        {2, 6, 7}, {2, 7, 3}, {0, 2, 3}, {0, 3, 1}, {4, 6, 7}, {4, 7, 5}};
    for (const auto &f : faces)
      sb.addTriangle(corner(f[0]), corner(f[1]), corner(f[2]), mat);
  }

  template <typename SB>
  static Camera cornell(Api &api, SB &sb, int width, int height) { // main.cpp:69-86
    api.loadObj("CornellBox-Original.obj", sb);
    sb.addSphere(Vec3(-0.38, 0.281, 0.38), 0.28,
                 MaterialSpec::makeReflective(Vec3(0.999, 0.999, 0.999), 0.95, 5));
    sb.setEnvironmentColour(Vec3(0.725, 0.71, 0.68) * 0.1);
    Camera camera = lookFrom(Vec3(0, 1, 3), Vec3(0, 1, 0), Vec3(0, 1, 0), width, height, 50.0);
    camera.setFocus(Vec3(0, 0, 0), 0.01);
    return camera;
  }

  template <typename SB>
  static Camera suzanne(Api &api, SB &sb, int width, int height) { // main.cpp:88-114
    api.loadObj("suzanne.obj", sb);
    const auto light = MaterialSpec::makeLight(Vec3(4, 4, 4));
    sb.addSphere(Vec3(0.5, 1, 3), 1, light);
    sb.addSphere(Vec3(1, 1, 3), 1, light);
    const auto backdrop = MaterialSpec::makeDiffuse(Vec3(0.20, 0.30, 0.36));
    const Vec3 tl(-5, -5, -1), tr(5, -5, -1), bl(-5, 5, -1), br(5, 5, -1);
    sb.addTriangle(tl, tr, bl, backdrop);
    sb.addTriangle(tr, bl, br, backdrop);
    const Vec3 lookAt(1, -0.6, 0.4);
    Camera camera = lookFrom(Vec3(1, -0.45, 4), lookAt, Vec3(0, 1, 0), width, height, 40.0);
    camera.setFocus(lookAt, 0.01);
    return camera;
  }

  template <typename SB>
  static Camera ce(Api &api, SB &sb, int width, int height) { // main.cpp:116-137
    api.loadObj("ce.obj", sb);
    sb.addSphere(Vec3(0, 1.6, 0), 1.0, MaterialSpec::makeLight(Vec3(1, 1, 1) * 10));
    sb.addSphere(Vec3(-0.2, 5.9, -0.3), 5.0, MaterialSpec::makeLight(Vec3(2.27, 3, 2.97) * 0.25));
    sb.addSphere(Vec3(), 10, MaterialSpec::makeDiffuse(Vec3(0.2, 0.2, 0.2)));
    const Vec3 lookAt(0, 0, 0);
    Camera camera =
        lookFrom(Vec3(0.27, 1.15, 0.36), lookAt, Vec3(0, 0, -1), width, height, 40.0);
    camera.setFocus(lookAt, 0.01);
    return camera;
  }

  // Shared by the two sphere-only scenes (main.cpp:139-149,165-176): camera and key light.
  template <typename SB>
  static Camera sphereStage(SB &sb, int width, int height) {
    const Vec3 eye(0, 0, -3.2);
    Camera camera = lookFrom(eye, Vec3(0, 0, 0), Vec3(0, 1, 0), width, height, 40.0);
    const double lightRadius = 3.0;
    sb.addSphere(eye + Vec3(6, 6, 0) - Vec3(0, 0, lightRadius), lightRadius,
                 MaterialSpec::makeLight(Vec3(1, 1, 1) * 8));
    return camera;
  }

  template <typename SB>
  static Camera singleSphere(Api &, SB &sb, int width, int height) { // main.cpp:139-163
    Camera camera = sphereStage(sb, width, height);
    auto ball = MaterialSpec::makeDiffuse(Vec3(0.2, 0.2, 0.2));
    ball.indexOfRefraction = 1.3;
    ball.reflectionConeAngleRadians = 0.05;
    sb.addSphere(Vec3(), 1, ball);
    sb.addSphere(Vec3(), 10, MaterialSpec::makeDiffuse(Vec3(0.2, 0.2, 0.5)));
    return camera;
  }

  template <typename SB>
  static Camera multiSphere(Api &, SB &sb, int width, int height) { // main.cpp:165-199
    Camera camera = sphereStage(sb, width, height);
    const auto radius = 1.0 / 5.0;
    const auto gap = radius * 2.15;
    for (int y = -2; y <= 2; ++y) {
      for (int x = -4; x <= 4; ++x) {
        auto mat = MaterialSpec::makeDiffuse(Vec3(0.90, 0.91, 0.92));
        mat.reflectionConeAngleRadians = 0.075 * (x + 4);
        mat.indexOfRefraction = 1.0 + 0.15 * (y + 2);
        sb.addSphere(Vec3(x * gap, y * gap, 0), radius, mat);
      }
    }
    sb.addSphere(Vec3(), 10, MaterialSpec::makeDiffuse(Vec3(0.2, 0.2, 0.5)));
    return camera;
  }

  template <typename SB>
  static Camera example1(Api &, SB &sb, int width, int height) { // main.cpp:201-228
    sb.addSphere(Vec3(1.5, 1.25, 0), 1.25, MaterialSpec::makeSpecular(hexColour(0x004358), 1.3));
    sb.addSphere(Vec3(-1, 1, 2), 1.0, MaterialSpec::makeSpecular(hexColour(0xffe11a), 1.3));
    sb.addSphere(Vec3(-2.5, 0.75, 0), 0.75, MaterialSpec::makeSpecular(hexColour(0xfd7400), 1.3));
    sb.addSphere(Vec3(-0.75, 0.5, -1), 0.5, MaterialSpec::makeSpecular(hexColour(0), 1.3));
    addCube(sb, Vec3(-10, -1, -10), Vec3(10, 0, 10),
            MaterialSpec::makeGlossy(Vec3(1, 1, 1), 1.1, 10.0));
    sb.addSphere(Vec3(-1.5, 4, 0), 0.5, MaterialSpec::makeLight(Vec3(1, 1, 1) * 30));
    Camera camera = lookFrom(Vec3(0, 2, -5), Vec3(0, 0.25, 3), Vec3(0, 1, 0), width, height, 45.0);
    camera.setFocus(Vec3(-0.75, 1, -1), 0.1);
    return camera;
  }

  template <typename SB>
  static Camera bbcOwl(Api &, SB &sb, int width, int height) { // main.cpp:230-289
    // The BBC Micro owl logo: 21 rows of 17 cells, '*' = a small sphere.
    static const char *const rows[21] = {
        "* * * * * * * * *", " *     * *     * ", "*   *   *   *   *", "   * *     * *   ",
        "*   *       *   *", " *     * *     * ", "* *     *     * *", " * *         *   ",
        "* * * * * * *   *", " * * * *         ", "* * * * *       *", " * * * *         ",
        "  * * * *       *", "   * * * *       ", "    * * * *     *", "     * * * *     ",
        "      * * * *   *", "       * * * *   ", "      *   *   * *", " * * * * * *   * ",
        "                *"};
    constexpr int owlHeight = 21;
    constexpr size_t owlWidth = 17;
    const auto spacing = 0.1;
    const auto size = spacing * 0.7;
    auto y = owlHeight * spacing - spacing / 2;
    for (const char *row : rows) {
      auto x = owlWidth * spacing / 2;
      for (size_t i = 0; i < owlWidth; ++i) {
        if (row[i] == '*')
          sb.addSphere(Vec3(x, y, 0), size, MaterialSpec::makeSpecular(hexColour(0xfeffd5), 1.3));
        x -= spacing;
      }
      y -= spacing;
    }
    auto plane = MaterialSpec::makeReflective(Vec3(0.2, 0.2, 0.2), 0.75, 3.0);
    plane.indexOfRefraction = 1.5;
    addCube(sb, Vec3(-10, -1, -10), Vec3(10, 0, 10), plane);
    sb.addSphere(Vec3(-1.5, 4.0, -1), 0.75, MaterialSpec::makeLight(Vec3(1, 1, 1) * 30));
    sb.setEnvironmentColour(Vec3(0.2, 0.2, 0.5) * 0.05);
    Camera camera = lookFrom(Vec3(4, 2.0, -5), Vec3(0, 0.5, 0), Vec3(0, 1, 0), width, height, 33.0);
    camera.setFocus(Vec3(0, 0.5, 0), 0.1);
    return camera;
  }

  // createScene (main.cpp:291-309).
  template <typename SB>
  static Camera create(Api &api, SB &sb, const std::string &name, int width, int height) {
    if (name == "cornell")
      return cornell(api, sb, width, height);
    if (name == "suzanne")
      return suzanne(api, sb, width, height);
    if (name == "ce")
      return ce(api, sb, width, height);
    if (name == "single-sphere")
      return singleSphere(api, sb, width, height);
    if (name == "multi-sphere")
      return multiSphere(api, sb, width, height);
    if (name == "example1")
      return example1(api, sb, width, height);
    if (name == "bbc-owl")
      return bbcOwl(api, sb, width, height);
    throw std::runtime_error("Unknown scene " + name);
  }
};

} // namespace ptb200
