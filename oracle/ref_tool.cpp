// ============================================================================================
// TEST INFRASTRUCTURE — driver around the REFERENCE's own, unmodified sources.
//
// Built only where /root/reference is mounted (oracle/Makefile target `_ref/ref_tool`), linked
// against the reference's src/dod/Scene.cpp, src/math/*.cpp and src/util/*.cpp compiled from
// where they lie.  The reference's scene recipes (src/main/main.cpp:27-309: DirRelativeOpener,
// hexColour, addCube, create*Scene, createScene) are pulled in textually by the Makefile into
// oracle/_ref/ref_recipes.inc (git-ignored, never committed) because main.cpp itself needs
// clara/libpng/range-v3, which do not exist here.
//
// Uses:  pin oracle/pt_oracle.cpp against the real thing (per-pass images, intersection
// records), generate tests/golden/ fixtures (tests/golden/make_golden.py), cross-check this
// repository's restated scene recipes (pt_three_ways_b200/host/SceneRecipes.h), and time the
// reference's own dod::Scene::render for bench.py --impl reference.
// ============================================================================================
#include "dod/Scene.h"
#include "fp/Render.h"
#include "fp/SceneBuilder.h"
#include "math/Camera.h"
#include "oo/Renderer.h"
#include "oo/SceneBuilder.h"
#include "math/Vec3.h"
#include "util/ArrayOutput.h"
#include "util/MaterialSpec.h"
#include "util/ObjLoader.h"
#include "util/RenderParams.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <random>
#include <string>
#include <string_view>
#include <unistd.h>
#include <vector>

namespace refmain {
#include "ref_recipes.inc" // generated: src/main/main.cpp lines 27-309, verbatim
} // namespace refmain

// This repository's restated recipes, instantiated over the REFERENCE's types.
#include "../pt_three_ways_b200/host/SceneRecipes.h"

namespace {

struct RecordingBuilder { // same idea as CaptureSceneBuilder, test/util/ObjLoaderTests.cpp:14-26
  std::vector<double> tri;
  std::vector<MaterialSpec> triMat;
  std::vector<double> sph;
  std::vector<MaterialSpec> sphMat;
  Vec3 env;
  void addTriangle(const Vec3 &a, const Vec3 &b, const Vec3 &c, const MaterialSpec &m) {
    for (const Vec3 *v : {&a, &b, &c}) {
      tri.push_back(v->x());
      tri.push_back(v->y());
      tri.push_back(v->z());
    }
    triMat.push_back(m);
  }
  void addSphere(const Vec3 &c, double r, const MaterialSpec &m) {
    sph.insert(sph.end(), {c.x(), c.y(), c.z(), r});
    sphMat.push_back(m);
  }
  void setEnvironmentColour(const Vec3 &c) { env = c; }
  bool operator==(const RecordingBuilder &o) const {
    return tri == o.tri && triMat == o.triMat && sph == o.sph && sphMat == o.sphMat && env == o.env;
  }
};

struct RefApi {
  using Vec3 = ::Vec3;
  using MaterialSpec = ::MaterialSpec;
  using Camera = ::Camera;
  static Norm3 unit(const ::Vec3 &v) { return v.normalised(); }
  template <typename SB>
  void loadObj(const char *name, SB &sb) {
    refmain::DirRelativeOpener opener("scenes");
    auto in = opener.open(name);
    loadObjFile(*in, opener, sb);
  }
};

static_assert(sizeof(Camera) == 18 * sizeof(double), "Camera layout (src/math/Camera.h:11-18)");

void materialDoubles(const MaterialSpec &m, double out[9]) {
  const double v[9] = {m.emission.x(), m.emission.y(), m.emission.z(),
                       m.diffuse.x(),  m.diffuse.y(),  m.diffuse.z(),
                       m.indexOfRefraction, m.reflectivity, m.reflectionConeAngleRadians};
  std::memcpy(out, v, sizeof v);
}

void writePtScene(const RecordingBuilder &rb, const Camera &camera, const std::string &path) {
  std::vector<MaterialSpec> palette;
  auto intern = [&](const MaterialSpec &m) {
    for (size_t i = 0; i < palette.size(); ++i)
      if (palette[i] == m)
        return static_cast<uint32_t>(i);
    palette.push_back(m);
    return static_cast<uint32_t>(palette.size() - 1);
  };
  std::vector<uint32_t> triIdx, sphIdx;
  for (auto &m : rb.triMat)
    triIdx.push_back(intern(m));
  for (auto &m : rb.sphMat)
    sphIdx.push_back(intern(m));
  std::ofstream out(path, std::ios::binary);
  const uint32_t counts[4] = {static_cast<uint32_t>(triIdx.size()),
                              static_cast<uint32_t>(sphIdx.size()),
                              static_cast<uint32_t>(palette.size()), 0};
  const double env[3] = {rb.env.x(), rb.env.y(), rb.env.z()};
  out.write("PTSCENE2", 8);
  out.write(reinterpret_cast<const char *>(counts), sizeof counts);
  out.write(reinterpret_cast<const char *>(env), sizeof env);
  out.write(reinterpret_cast<const char *>(rb.tri.data()), rb.tri.size() * 8);
  out.write(reinterpret_cast<const char *>(triIdx.data()), triIdx.size() * 4);
  out.write(reinterpret_cast<const char *>(rb.sph.data()), rb.sph.size() * 8);
  out.write(reinterpret_cast<const char *>(sphIdx.data()), sphIdx.size() * 4);
  for (auto &m : palette) {
    double v[9];
    materialDoubles(m, v);
    out.write(reinterpret_cast<const char *>(v), sizeof v);
  }
  // The recipe's camera (18 doubles, Camera.h:11-18).  Only aspectRatio_, reciprocalHeight_
  // and reciprocalWidth_ depend on the image size; readers patch those three.
  out.write(reinterpret_cast<const char *>(&camera), sizeof camera);
}

Camera resizedCamera(const Camera &recipeCamera, int width, int height) {
  double v[18];
  std::memcpy(v, &recipeCamera, sizeof v);
  v[12] = static_cast<double>(width) / height; // aspectRatio_      (Camera.h:43)
  v[14] = 1.0 / height;                        // reciprocalHeight_ (Camera.h:46)
  v[15] = 1.0 / width;                         // reciprocalWidth_  (Camera.h:46)
  Camera out = recipeCamera;
  std::memcpy(&out, v, sizeof v);
  return out;
}

// Loads a PTSCENE2 file into any SceneBuilder; returns the camera for the requested size.
template <typename SB>
Camera readPtScene(const std::string &path, SB &sb, int width, int height) {
  std::ifstream in(path, std::ios::binary);
  if (!in)
    throw std::runtime_error("Unable to open " + path);
  char magic[8];
  uint32_t counts[4];
  double env[3];
  in.read(magic, 8);
  in.read(reinterpret_cast<char *>(counts), sizeof counts);
  in.read(reinterpret_cast<char *>(env), sizeof env);
  if (std::memcmp(magic, "PTSCENE2", 8) != 0)
    throw std::runtime_error("bad magic in " + path);
  std::vector<double> tri(counts[0] * 9), sph(counts[1] * 4), mats(counts[2] * 9);
  std::vector<uint32_t> triIdx(counts[0]), sphIdx(counts[1]);
  in.read(reinterpret_cast<char *>(tri.data()), tri.size() * 8);
  in.read(reinterpret_cast<char *>(triIdx.data()), triIdx.size() * 4);
  in.read(reinterpret_cast<char *>(sph.data()), sph.size() * 8);
  in.read(reinterpret_cast<char *>(sphIdx.data()), sphIdx.size() * 4);
  in.read(reinterpret_cast<char *>(mats.data()), mats.size() * 8);
  auto material = [&](uint32_t i) {
    const double *m = &mats[9 * i];
    return MaterialSpec{Vec3(m[0], m[1], m[2]), Vec3(m[3], m[4], m[5]), m[6], m[7], m[8]};
  };
  for (uint32_t i = 0; i < counts[0]; ++i) {
    const double *t = &tri[9 * i];
    sb.addTriangle(Vec3(t[0], t[1], t[2]), Vec3(t[3], t[4], t[5]), Vec3(t[6], t[7], t[8]),
                   material(triIdx[i]));
  }
  for (uint32_t i = 0; i < counts[1]; ++i) {
    const double *s = &sph[4 * i];
    sb.addSphere(Vec3(s[0], s[1], s[2]), s[3], material(sphIdx[i]));
  }
  sb.setEnvironmentColour(Vec3(env[0], env[1], env[2]));
  double cam18[18];
  in.read(reinterpret_cast<char *>(cam18), sizeof cam18);
  if (!in)
    throw std::runtime_error("truncated " + path);
  Camera camera(Vec3(0, 0, 0), Vec3(0, 0, 1), Vec3(0, 1, 0).normalised(), 4, 3, 40.0);
  std::memcpy(&camera, cam18, sizeof cam18);
  return resizedCamera(camera, width, height);
}

bool endsWith(const std::string &s, const char *suffix) {
  const size_t n = std::strlen(suffix);
  return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

RenderParams sized(int w, int h) {
  RenderParams p;
  p.width = w;
  p.height = h;
  return p;
}

int usage() {
  std::cerr << "ref_tool scene NAME OUT.ptscene\n"
               "ref_tool camera NAME W H\n"
               "ref_tool check-recipes\n"
               "ref_tool pass NAME W H SEED PASS NU NV MAXDEPTH PREVIEW OUT.f64\n"
               "ref_tool render NAME W H SPP MAXCPUS SEED NU NV MAXDEPTH OUT.raw|-\n"
               "ref_tool passes NAME W H SPP THREADS SEED NU NV MAXDEPTH OUT.raw|-   (every pass kept)\n"
               "ref_tool intersect NAME|FILE.ptscene WHICH NEARER RAYS.f64 OUT.f64\n"
               "ref_tool fp-pass NAME W H SEED NU NV MAXDEPTH PREVIEW OUT.f64   (fp::render, one pass)\n"
               "ref_tool fp-render NAME W H SPP MAXCPUS SEED NU NV MAXDEPTH OUT.raw|-\n"
               "ref_tool oo-pass NAME W H SEED PASS NU NV MAXDEPTH PREVIEW OUT.f64   (oo::Renderer, one pass)\n"
               "ref_tool oo-render NAME W H SPP MAXCPUS SEED NU NV MAXDEPTH OUT.raw|-\n";
  return 2;
}

// NAME (reference recipe; needs the reference tree for the OBJ files) or FILE.ptscene
// (self-contained; works where /root/reference does not exist, e.g. on the GPU box).
std::string gStartDir;
std::string absolutePath(const std::string &p) {
  return (!p.empty() && p[0] == '/') ? p : gStartDir + "/" + p;
}
template <typename SB>
Camera makeScene(const std::string &which, SB &sb, int width, int height) {
  if (endsWith(which, ".ptscene"))
    return readPtScene(absolutePath(which), sb, width, height);
  const char *refRoot = std::getenv("PT_REFERENCE_ROOT");
  if (chdir(refRoot ? refRoot : "/root/reference") != 0) // the recipes open "scenes/<file>"
    throw std::runtime_error("reference tree not found; pass a .ptscene file instead");
  return refmain::createScene(sb, which, sized(width, height));
}

} // namespace

int main(int argc, char **argv) {
  if (argc < 2)
    return usage();
  const std::string cmd = argv[1];
  char cwdBuf[4096];
  gStartDir = getcwd(cwdBuf, sizeof cwdBuf) ? cwdBuf : ".";
  try {
    if (cmd == "scene" && argc == 4) {
      RecordingBuilder rb;
      Camera cam = makeScene(argv[2], rb, 64, 48);
      writePtScene(rb, cam, absolutePath(argv[3]));
      std::printf("{\"triangles\": %zu, \"spheres\": %zu}\n", rb.triMat.size(), rb.sphMat.size());
      return 0;
    }
    if (cmd == "camera" && argc == 5) {
      RecordingBuilder rb;
      Camera cam = makeScene(argv[2], rb, std::atoi(argv[3]), std::atoi(argv[4]));
      double v[18];
      std::memcpy(v, &cam, sizeof v);
      for (int i = 0; i < 18; ++i)
        std::printf("%a%c", v[i], i == 17 ? '\n' : ' ');
      return 0;
    }
    if (cmd == "check-recipes") {
      const char *names[] = {"cornell", "suzanne", "ce", "single-sphere", "multi-sphere",
                             "example1", "bbc-owl"};
      const int sizes[][2] = {{64, 48}, {640, 480}, {1280, 720}, {1920, 1080}, {256, 256}, {16, 16}};
      int bad = 0;
      for (const char *name : names) {
        for (auto &wh : sizes) {
          RecordingBuilder real, restated;
          Camera a = makeScene(name, real, wh[0], wh[1]);
          RefApi api;
          Camera b = ptb200::SceneRecipes<RefApi>::create(api, restated, name, wh[0], wh[1]);
          // resizedCamera must reproduce what the constructor computes for another size.
          RecordingBuilder scratch;
          Camera c = resizedCamera(makeScene(name, scratch, 64, 48), wh[0], wh[1]);
          // The restated recipe goes through one more function call than the reference's
          // text; under -funsafe-math-optimizations GCC may contract/fold differently, so
          // the camera is compared to 4 ulp, everything else exactly.
          double av[18], bv[18];
          std::memcpy(av, &a, sizeof av);
          std::memcpy(bv, &b, sizeof bv);
          bool cameraClose = true;
          for (int i = 0; i < 18; ++i)
            cameraClose = cameraClose && std::fabs(av[i] - bv[i]) <= 4 * 2.3e-16 * std::fabs(av[i]);
          const bool same = real == restated && cameraClose && std::memcmp(&a, &c, sizeof a) == 0;
          if (!same) {
            ++bad;
            std::printf("MISMATCH %s %dx%d scene=%d restatedCam=%d resizedCam=%d\n", name, wh[0], wh[1],
                        int(real == restated), int(std::memcmp(&a, &b, sizeof a) == 0),
                        int(std::memcmp(&a, &c, sizeof a) == 0));
          }
        }
        std::printf("%s ok\n", name);
      }
      return bad ? 1 : 0;
    }
    if (cmd == "pass" && argc == 12) {
      RenderParams p = sized(std::atoi(argv[3]), std::atoi(argv[4]));
      p.seed = std::atoi(argv[5]);
      const int pass = std::atoi(argv[6]);
      p.firstBounceUSamples = std::atoi(argv[7]);
      p.firstBounceVSamples = std::atoi(argv[8]);
      p.maxDepth = std::atoi(argv[9]);
      p.preview = std::atoi(argv[10]) != 0;
      dod::Scene scene;
      Camera cam = makeScene(argv[2], scene, p.width, p.height);
      // The body of the per-pass lambda of dod::Scene::render (src/dod/Scene.cpp:210-217),
      // driving the reference's own radiance() and Camera::randomRay().
      std::vector<double> colours(static_cast<size_t>(p.width) * p.height * 3);
      std::mt19937 rng(p.seed + pass);
      for (int y = 0; y < p.height; ++y) {
        for (int x = 0; x < p.width; ++x) {
          auto ray = cam.randomRay(x, y, rng);
          const Vec3 c = scene.radiance(rng, ray, 0, p);
          double *dst = &colours[3 * (static_cast<size_t>(x) + static_cast<size_t>(y) * p.width)];
          dst[0] = c.x();
          dst[1] = c.y();
          dst[2] = c.z();
        }
      }
      std::ofstream out(absolutePath(argv[11]), std::ios::binary);
      out.write(reinterpret_cast<const char *>(colours.data()), colours.size() * 8);
      return 0;
    }
    if (cmd == "render" && argc == 12) {
      RenderParams p = sized(std::atoi(argv[3]), std::atoi(argv[4]));
      p.samplesPerPixel = std::atoi(argv[5]);
      p.maxCpus = std::atoi(argv[6]);
      p.seed = std::atoi(argv[7]);
      p.firstBounceUSamples = std::atoi(argv[8]);
      p.firstBounceVSamples = std::atoi(argv[9]);
      p.maxDepth = std::atoi(argv[10]);
      dod::Scene scene;
      Camera cam = makeScene(argv[2], scene, p.width, p.height);
      std::streambuf *saved = std::cout.rdbuf(std::cerr.rdbuf()); // Progressifier prints to cout
      const auto t0 = std::chrono::steady_clock::now();
      ArrayOutput output = scene.render(cam, p, [](ArrayOutput &) {}); // the unmodified entry point
      const auto t1 = std::chrono::steady_clock::now();
      std::cout.rdbuf(saved);
      const std::string outPath = argv[11];
      if (outPath != "-")
        output.save(absolutePath(outPath));
      std::printf("{\"seconds\": %.6f, \"total_samples\": %zu, \"pixels\": %d}\n",
                  std::chrono::duration<double>(t1 - t0).count(), output.totalSamples(),
                  p.width * p.height);
      return 0;
    }
    if (cmd == "passes" && argc == 12) {
      // The reference's own per-pass work (the lambda of dod::Scene::render, Scene.cpp:210-217:
      // its radiance() and Camera::randomRay(), compiled from its sources) with the passes handed
      // to THREADS workers one at a time and EVERY pass kept — Scene::render itself abandons the
      // passes still in flight when the last one is launched (Scene.cpp:251).  This is the
      // reference's arithmetic at the rate a fair scheduler gives it: bench.py's CPU baseline.
      RenderParams p = sized(std::atoi(argv[3]), std::atoi(argv[4]));
      p.samplesPerPixel = std::atoi(argv[5]);
      const int threads = std::max(1, std::atoi(argv[6]));
      p.seed = std::atoi(argv[7]);
      p.firstBounceUSamples = std::atoi(argv[8]);
      p.firstBounceVSamples = std::atoi(argv[9]);
      p.maxDepth = std::atoi(argv[10]);
      dod::Scene scene;
      const Camera cam = makeScene(argv[2], scene, p.width, p.height);
      std::atomic<int> nextPass{0};
      std::vector<ArrayOutput> partial(threads, ArrayOutput(p.width, p.height));
      const auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> workers;
      for (int t = 0; t < threads; ++t)
        workers.emplace_back([&, t] {
          for (int pass = nextPass++; pass < p.samplesPerPixel; pass = nextPass++) {
            std::mt19937 rng(p.seed + pass);
            for (int y = 0; y < p.height; ++y)
              for (int x = 0; x < p.width; ++x) {
                auto ray = cam.randomRay(x, y, rng);
                partial[t].addSamples(x, y, scene.radiance(rng, ray, 0, p), 1);
              }
          }
        });
      for (auto &w : workers)
        w.join();
      ArrayOutput output(p.width, p.height);
      for (auto &part : partial)
        output += part;
      const auto t1 = std::chrono::steady_clock::now();
      const std::string outPath = argv[11];
      if (outPath != "-")
        output.save(absolutePath(outPath));
      std::printf("{\"seconds\": %.6f, \"total_samples\": %zu, \"pixels\": %d}\n",
                  std::chrono::duration<double>(t1 - t0).count(), output.totalSamples(),
                  p.width * p.height);
      return 0;
    }
    if (cmd == "fp-pass" && argc == 11) {
      // One whole-screen pass of the `fp` way (src/fp/Render.cpp:120-135) through its public
      // entry point: samplesPerPixel = 1, maxCpus = 1 runs renderWholeScreen(seed) exactly once.
      RenderParams p = sized(std::atoi(argv[3]), std::atoi(argv[4]));
      p.seed = std::atoi(argv[5]);
      p.firstBounceUSamples = std::atoi(argv[6]);
      p.firstBounceVSamples = std::atoi(argv[7]);
      p.maxDepth = std::atoi(argv[8]);
      p.preview = std::atoi(argv[9]) != 0;
      p.samplesPerPixel = 1;
      p.maxCpus = 1;
      fp::SceneBuilder builder;
      Camera cam = makeScene(argv[2], builder, p.width, p.height);
      std::streambuf *saved = std::cout.rdbuf(std::cerr.rdbuf());
      ArrayOutput output = fp::render(cam, builder.scene(), p, [](const ArrayOutput &) {});
      std::cout.rdbuf(saved);
      std::vector<double> colours(static_cast<size_t>(p.width) * p.height * 3);
      for (int y = 0; y < p.height; ++y)
        for (int x = 0; x < p.width; ++x) {
          const Vec3 c = output.rawPixelAt(x, y); // sum * (1/1)
          double *dst = &colours[3 * (static_cast<size_t>(x) + static_cast<size_t>(y) * p.width)];
          dst[0] = c.x();
          dst[1] = c.y();
          dst[2] = c.z();
        }
      std::ofstream out(absolutePath(argv[10]), std::ios::binary);
      out.write(reinterpret_cast<const char *>(colours.data()), colours.size() * 8);
      return 0;
    }
    if (cmd == "fp-render" && argc == 12) {
      RenderParams p = sized(std::atoi(argv[3]), std::atoi(argv[4]));
      p.samplesPerPixel = std::atoi(argv[5]);
      p.maxCpus = std::atoi(argv[6]);
      p.seed = std::atoi(argv[7]);
      p.firstBounceUSamples = std::atoi(argv[8]);
      p.firstBounceVSamples = std::atoi(argv[9]);
      p.maxDepth = std::atoi(argv[10]);
      fp::SceneBuilder builder;
      Camera cam = makeScene(argv[2], builder, p.width, p.height);
      std::streambuf *saved = std::cout.rdbuf(std::cerr.rdbuf());
      const auto t0 = std::chrono::steady_clock::now();
      ArrayOutput output = fp::render(cam, builder.scene(), p, [](const ArrayOutput &) {});
      const auto t1 = std::chrono::steady_clock::now();
      std::cout.rdbuf(saved);
      const std::string outPath = argv[11];
      if (outPath != "-")
        output.save(absolutePath(outPath));
      std::printf("{\"seconds\": %.6f, \"total_samples\": %zu, \"pixels\": %d}\n",
                  std::chrono::duration<double>(t1 - t0).count(), output.totalSamples(),
                  p.width * p.height);
      return 0;
    }
    if (cmd == "oo-pass" && argc == 12) {
      // The body of the per-pass lambda of oo::Renderer::render (src/oo/Renderer.cpp:97-107),
      // driving the reference's own oo::Renderer::radiance() ("visible for testing",
      // src/oo/Renderer.h:43) and Camera::randomRay().
      RenderParams p = sized(std::atoi(argv[3]), std::atoi(argv[4]));
      p.seed = std::atoi(argv[5]);
      const int pass = std::atoi(argv[6]);
      p.firstBounceUSamples = std::atoi(argv[7]);
      p.firstBounceVSamples = std::atoi(argv[8]);
      p.maxDepth = std::atoi(argv[9]);
      p.preview = std::atoi(argv[10]) != 0;
      oo::SceneBuilder builder;
      Camera cam = makeScene(argv[2], builder, p.width, p.height);
      oo::Renderer renderer(builder.scene(), cam, p);
      std::vector<double> colours(static_cast<size_t>(p.width) * p.height * 3);
      std::mt19937 rng(p.seed + pass);
      for (int y = 0; y < p.height; ++y) {
        for (int x = 0; x < p.width; ++x) {
          auto ray = cam.randomRay(x, y, rng);
          const Vec3 c = renderer.radiance(rng, ray, 0);
          double *dst = &colours[3 * (static_cast<size_t>(x) + static_cast<size_t>(y) * p.width)];
          dst[0] = c.x();
          dst[1] = c.y();
          dst[2] = c.z();
        }
      }
      std::ofstream out(absolutePath(argv[11]), std::ios::binary);
      out.write(reinterpret_cast<const char *>(colours.data()), colours.size() * 8);
      return 0;
    }
    if (cmd == "oo-render" && argc == 12) {
      RenderParams p = sized(std::atoi(argv[3]), std::atoi(argv[4]));
      p.samplesPerPixel = std::atoi(argv[5]);
      p.maxCpus = std::atoi(argv[6]);
      p.seed = std::atoi(argv[7]);
      p.firstBounceUSamples = std::atoi(argv[8]);
      p.firstBounceVSamples = std::atoi(argv[9]);
      p.maxDepth = std::atoi(argv[10]);
      oo::SceneBuilder builder;
      Camera cam = makeScene(argv[2], builder, p.width, p.height);
      oo::Renderer renderer(builder.scene(), cam, p);
      std::streambuf *saved = std::cout.rdbuf(std::cerr.rdbuf());
      const auto t0 = std::chrono::steady_clock::now();
      ArrayOutput output = renderer.render([](const ArrayOutput &) {}); // the unmodified entry point
      const auto t1 = std::chrono::steady_clock::now();
      std::cout.rdbuf(saved);
      const std::string outPath = argv[11];
      if (outPath != "-")
        output.save(absolutePath(outPath));
      std::printf("{\"seconds\": %.6f, \"total_samples\": %zu, \"pixels\": %d}\n",
                  std::chrono::duration<double>(t1 - t0).count(), output.totalSamples(),
                  p.width * p.height);
      return 0;
    }
    if (cmd == "intersect" && argc == 7) {
      dod::Scene scene;
      makeScene(argv[2], scene, 64, 48);
      const int mode = std::atoi(argv[3]);
      const double nearer = std::string(argv[4]) == "inf" ? std::numeric_limits<double>::infinity()
                                                           : std::atof(argv[4]);
      std::ifstream in(absolutePath(argv[5]), std::ios::binary | std::ios::ate);
      const size_t bytes = static_cast<size_t>(in.tellg());
      in.seekg(0);
      std::vector<double> rays(bytes / 8, 0.0);
      in.read(reinterpret_cast<char *>(rays.data()), static_cast<std::streamsize>(bytes));
      const size_t n = rays.size() / 6;
      std::vector<double> out(n * 18, 0.0);
      for (size_t i = 0; i < n; ++i) {
        const double *r = &rays[6 * i];
        // Directions arrive already normalised; Norm3 has no public constructor from raw
        // components, so go through fromNormal (asserts unit length in debug builds only).
        const Ray ray(Vec3(r[0], r[1], r[2]), Norm3::fromNormal(Vec3(r[3], r[4], r[5])));
        auto rec = mode == 1   ? scene.intersectSpheres(ray, nearer)
                   : mode == 2 ? scene.intersectTriangles(ray, nearer)
                               : scene.intersect(ray);
        double *o = &out[18 * i];
        if (rec) {
          o[0] = 1;
          o[1] = rec->hit.distance;
          o[2] = rec->hit.inside ? 1 : 0;
          o[3] = rec->hit.position.x(); o[4] = rec->hit.position.y(); o[5] = rec->hit.position.z();
          o[6] = rec->hit.normal.x(); o[7] = rec->hit.normal.y(); o[8] = rec->hit.normal.z();
          materialDoubles(rec->material, o + 9);
        }
      }
      std::ofstream outFile(absolutePath(argv[6]), std::ios::binary);
      outFile.write(reinterpret_cast<const char *>(out.data()), out.size() * 8);
      return 0;
    }
  } catch (const std::exception &e) {
    std::cerr << "error: " << e.what() << "\n";
    return 1;
  }
  return usage();
}
