// The "Api" that instantiates SceneRecipes<> over this repository's host types.
#pragma once

#include "Camera.h"
#include "MaterialSpec.h"
#include "ObjLoader.h"
#include "Vec3.h"

#include <string>

namespace ptb200 {

struct HostApi {
  using Vec3 = ptb200::Vec3;
  using MaterialSpec = ptb200::MaterialSpec;
  using Camera = ptb200::Camera;

  std::string scenesDir;
  explicit HostApi(std::string dir = "scenes") : scenesDir(std::move(dir)) {}

  static Norm3 unit(const Vec3 &v) { return v.normalised(); }

  template <typename SB>
  void loadObj(const char *fileName, SB &sb) {
    DirRelativeOpener opener(scenesDir);
    auto in = opener.open(fileName);
    loadObjFile(*in, opener, sb);
  }
};

} // namespace ptb200
