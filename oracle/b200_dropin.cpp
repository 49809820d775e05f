// ============================================================================================
// TEST INFRASTRUCTURE — proof that the B200 backend drops into the reference ITSELF.
//
// Built only where /root/reference is mounted (oracle/Makefile target `_ref/b200_dropin`):
//   * include/ptb200_scene.hpp (b200::Scene, the adaptor INTEGRATION.md asks a maintainer to
//     add as src/b200/Scene.h) compiled against the reference's OWN Camera, ArrayOutput,
//     MaterialSpec, RenderParams and loadObjFile, from where they lie;
//   * driven by the reference's own createScene<SB> recipes (src/main/main.cpp:27-309, pulled in
//     textually by the Makefile as _ref/ref_recipes.inc) exactly as doRender does for dod::Scene
//     (main.cpp:360-363);
//   * linked against libptb200.so.
// The recipes open "scenes/<file>" relative to the working directory (main.cpp:71), so the
// Makefile copies the reference's scenes/ next to the binary (oracle/_ref/scenes/, git-ignored)
// and the tool chdir()s to its own directory.
//
//   b200_dropin render NAME W H SPP SEED OUT.raw [MODE]   createScene<b200::Scene> + render + save
//   b200_dropin arrays NAME W H OUT.bin                    the PtScene/PtCamera the adaptor marshals
// ============================================================================================
#include "ptb200_scene.hpp"

#include "math/Vec3.h"
#include "util/ObjLoader.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <libgen.h>
#include <string>
#include <unistd.h>

namespace refmain {
#include "ref_recipes.inc" // generated: src/main/main.cpp lines 27-309, verbatim
} // namespace refmain

namespace {

std::string absolutePath(const std::string &path) {
  if (!path.empty() && path[0] == '/')
    return path;
  char cwd[4096];
  if (!getcwd(cwd, sizeof cwd))
    return path;
  return std::string(cwd) + "/" + path;
}

void enterOwnDirectory() { // where the Makefile put scenes/
  char exe[4096];
  const ssize_t n = readlink("/proc/self/exe", exe, sizeof exe - 1);
  if (n <= 0)
    return;
  exe[n] = 0;
  if (chdir(dirname(exe)) != 0)
    std::perror("chdir");
}

} // namespace

int main(int argc, char **argv) {
  try {
    if (argc < 6) {
      std::cerr << "b200_dropin render NAME W H SPP SEED OUT.raw [rngMode]\n"
                   "b200_dropin arrays NAME W H OUT.bin\n";
      return 2;
    }
    const std::string command = argv[1], name = argv[2];
    RenderParams params;
    params.width = std::atoi(argv[3]);
    params.height = std::atoi(argv[4]);
    const std::string out = absolutePath(command == "render" ? (argc > 7 ? argv[7] : "") : argv[5]);
    enterOwnDirectory();

    b200::Scene scene;
    // main.cpp:360-363 with b200::Scene in the place of dod::Scene
    const Camera camera = refmain::createScene(scene, name, params);

    if (command == "arrays") {
      const PtScene s = scene.abi();
      const PtCamera c = b200::Scene::abi(camera);
      std::ofstream f(out, std::ios::binary);
      const uint32_t counts[4] = {s.numTriangles, s.numSpheres, s.numMaterials, 0};
      f.write(reinterpret_cast<const char *>(counts), sizeof counts);
      f.write(reinterpret_cast<const char *>(s.environment), sizeof s.environment);
      f.write(reinterpret_cast<const char *>(s.triangleVertices), size_t(s.numTriangles) * 72);
      f.write(reinterpret_cast<const char *>(s.triangleMaterial), size_t(s.numTriangles) * 4);
      f.write(reinterpret_cast<const char *>(s.sphereCentreRadius), size_t(s.numSpheres) * 32);
      f.write(reinterpret_cast<const char *>(s.sphereMaterial), size_t(s.numSpheres) * 4);
      f.write(reinterpret_cast<const char *>(s.materials), size_t(s.numMaterials) * sizeof(PtMaterial));
      f.write(reinterpret_cast<const char *>(&c), sizeof c);
      return f ? 0 : 1;
    }
    if (command != "render" || argc < 8)
      return 2;
    params.samplesPerPixel = std::atoi(argv[5]);
    params.seed = std::atoi(argv[6]);
    PtRenderOptions options{};
    options.rngMode = argc > 8 ? std::atoi(argv[8]) : PTB200_RNG_KEYED_PHILOX;
    scene.setOptions(options);
    int updates = 0;
    const auto start = std::chrono::steady_clock::now();
    const ArrayOutput output = scene.render(camera, params, [&](ArrayOutput &) { ++updates; });
    const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    output.save(out); // the reference's own raw writer (src/util/ArrayOutput.cpp:65-81)
    std::printf("{\"total_samples\": %zu, \"seconds\": %.6f, \"updates\": %d, \"casts\": %llu}\n",
                output.totalSamples(), seconds, updates,
                static_cast<unsigned long long>(scene.lastStats().casts));
    return 0;
  } catch (const std::exception &e) {
    std::cerr << "b200_dropin: " << e.what() << "\n";
    return 1;
  }
}
