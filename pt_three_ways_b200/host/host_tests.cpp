// Host-side tests of the drop-in boundary (no GPU, no Catch2): the reference's own host-level
// test cases restated against this repository's host types —
//   test/util/ObjLoaderTests.cpp:36-97  (tokenizer edge cases, parse errors, a triangle, MTL)
//   test/util/ArrayOutputTests.cpp:9-39 (construction, raw-file round trip)
//   test/math/*                          (Vec3/Norm3 algebra used by Camera)
// plus checks the reference has no test for: scene recipes + loader against the PTSCENE2
// fixtures the reference's own loader produced (needs --scenes DIR with the OBJ files), the
// SceneBuilder adaptor's flat arrays, and that render() fails loudly without a CUDA device.
#include "ArrayOutput.h"
#include "HostApi.h"
#include "ObjLoader.h"
#include "PngWriter.h"
#include "Scene.h"
#include "SceneRecipes.h"
#include "../csrc/pt_mt19937.cuh" // the per-lane engine of the fp way, host-compiled

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <random>
#include <sstream>
#include <unistd.h>

using namespace ptb200;

static int failures = 0;
static int checks = 0;
#define CHECK(cond)                                                                              \
  do {                                                                                           \
    ++checks;                                                                                    \
    if (!(cond)) {                                                                               \
      ++failures;                                                                                \
      std::cerr << __FILE__ << ":" << __LINE__ << ": CHECK failed: " #cond "\n";                 \
    }                                                                                            \
  } while (0)

struct ThrowingOpener : ObjLoaderOpener {
  std::unique_ptr<std::istream> open(const std::string &) override {
    throw std::runtime_error("Unexpected");
  }
};
struct CaptureSceneBuilder {
  struct Triangle {
    Vec3 v0, v1, v2;
    MaterialSpec material;
  };
  std::vector<Triangle> triangles;
  void addTriangle(const Vec3 &a, const Vec3 &b, const Vec3 &c, const MaterialSpec &m) {
    triangles.push_back({a, b, c, m});
  }
};
static CaptureSceneBuilder L(const char *text) {
  ThrowingOpener opener;
  std::istringstream in(text);
  CaptureSceneBuilder csb;
  loadObjFile(in, opener, csb);
  return csb;
}
static std::string thrown(const char *text) {
  try {
    L(text);
  } catch (const std::exception &e) {
    return e.what();
  }
  return "";
}

static void objLoaderTests() {
  for (const char *blank : {"", "\n", "  \n", "  \n  ", "\r", "  \r", "  \r  ", "\r\n", "  \r\n",
                            "  \r\n  ", "# comment", "  # comment", "  # comment\n#another\n"})
    CHECK(L(blank).triangles.empty());
  CHECK(thrown("nope") == "Unknown directive 'nope' on line 1");
  CHECK(thrown("\nblargh") == "Unknown directive 'blargh' on line 2");
  auto res = L("\nv 0 0 0\nv 0 0 1\nv 0 1 0\nf -3 -2 -1\n");
  CHECK(res.triangles.size() == 1);
  if (res.triangles.size() == 1) {
    CHECK(res.triangles[0].v0 == Vec3(0, 0, 0));
    CHECK(res.triangles[0].v1 == Vec3(0, 0, 1));
    CHECK(res.triangles[0].v2 == Vec3(0, 1, 0));
  }
  // fan triangulation, 1-based indices, trailing CR and a/b/c tokens
  auto quad = L("v 0 0 0\r\nv 1 0 0\r\nv 1 1 0\r\nv 0 1 0\r\nf 1/1/1 2/2/2 3/3/3 4/4/4 \r\n");
  CHECK(quad.triangles.size() == 2);
  if (quad.triangles.size() == 2) {
    CHECK(quad.triangles[1].v0 == Vec3(0, 0, 0));
    CHECK(quad.triangles[1].v1 == Vec3(1, 1, 0));
    CHECK(quad.triangles[1].v2 == Vec3(0, 1, 0));
  }
  std::istringstream mtl(R"(
newmtl leftWall
  Ns 10.0000
  Ni 1.5000
  illum 2
  Ka 0.63 0.065 0.05 # Red
  Kd 0.63 0.065 0.05
  Ks 0 0 0
  Ke 0 0 0


newmtl light
  Ns 10.0000
  Ni 1.0000
  illum 2
  Ka 0.78 0.78 0.78 # White
  Kd 0.78 0.78 0.78
  Ks 0 0 0
  Ke 17 12 4
newmtl mirror
  illum 3
  Ka 0.1 0.2 0.2
  Ns 250
)");
  auto mats = loadMaterials(mtl);
  CHECK(mats.size() == 3);
  CHECK(mats.at("leftWall").diffuse == Vec3(0.63, 0.065, 0.05));
  CHECK(mats.at("leftWall").emission == Vec3());
  CHECK(mats.at("leftWall").indexOfRefraction == 1.5);
  CHECK(mats.at("leftWall").reflectionConeAngleRadians == M_PI * 0.9);
  CHECK(mats.at("leftWall").reflectivity == -1);
  CHECK(mats.at("light").diffuse == Vec3(0.78, 0.78, 0.78));
  CHECK(mats.at("light").emission == Vec3(17, 12, 4));
  CHECK(mats.at("mirror").reflectivity == Vec3(0.1, 0.2, 0.2).length()); // illum 3 -> |Ka|
  CHECK(mats.at("mirror").reflectionConeAngleRadians == 0.0);            // clamp(1 - 2.5, 0, 1)
}

static void arrayOutputTests() {
  ArrayOutput fresh(10, 20);
  CHECK(fresh.width() == 10);
  CHECK(fresh.height() == 20);
  CHECK((fresh.pixelAt(0, 0) == ArrayOutput::Pixel{0, 0, 0}));
  CHECK(fresh.rawPixelAt(0, 0) == Vec3());
  char path[] = "/tmp/ptb200arrayoutputXXXXXX";
  const int fd = mkstemp(path);
  CHECK(fd >= 0);
  ArrayOutput ao(7, 5);
  ao.addSamples(0, 0, Vec3(0.2, 0.3, 0.4), 12);
  ao.addSamples(1, 0, Vec3(0.4, 0.6, 0.7), 1);
  ao.addSamples(0, 3, Vec3(0.1, 0.2, 0.3), 2);
  ao.save(path);
  ArrayOutput loaded = ArrayOutput::load(path);
  close(fd);
  unlink(path);
  CHECK(loaded.width() == 7 && loaded.height() == 5);
  for (int y = 0; y < 5; ++y)
    for (int x = 0; x < 7; ++x) {
      CHECK(loaded.rawPixelAt(x, y) == ao.rawPixelAt(x, y));
      CHECK(loaded.pixelAt(x, y) == ao.pixelAt(x, y));
    }
  CHECK(loaded.totalSamples() == 15);
  ArrayOutput sum(7, 5);
  sum += ao;
  sum += loaded;
  CHECK(sum.totalSamples() == 30);
  bool threw = false;
  try {
    sum += fresh;
  } catch (const std::logic_error &) {
    threw = true;
  }
  CHECK(threw);
  CHECK(ao.rawPixelAt(0, 0) == Vec3(0.2, 0.3, 0.4) * (1.0 / 12));
}

static void mathTests() {
  const Vec3 a(1, 2, 3), b(-2, 0.5, 4);
  CHECK(a.dot(b) == 1 * -2 + 2 * 0.5 + 3 * 4);
  CHECK(a.cross(b) == Vec3(2 * 4 - 3 * 0.5, 3 * -2 - 1 * 4, 1 * 0.5 - 2 * -2));
  CHECK(std::fabs(a.normalised().toVec3().length() - 1.0) < 1e-15);
  CHECK((a / 2.0) == a * 0.5);
  const Camera cam(Vec3(0, 1, 3), Vec3(0, 1, 0), Vec3(0, 1, 0).normalised(), 640, 480, 50.0);
  const PtCamera &abi = cam.abi();
  CHECK(abi.centre[2] == 3 && abi.axisZ[2] == -1 && abi.axisX[0] == -1 && abi.axisY[1] == 1);
  CHECK(abi.aspectRatio == 640.0 / 480 && abi.reciprocalWidth == 1.0 / 640);
  CHECK(std::fabs(abi.cameraPlaneDist - 1.0 / std::tan(50.0 * M_PI / 360.0)) < 1e-15);
}

static void sceneAdaptorTests() {
  Scene scene;
  const auto red = MaterialSpec::makeDiffuse(Vec3(1, 0, 0));
  const auto light = MaterialSpec::makeLight(Vec3(5, 5, 5));
  scene.addTriangle(Vec3(0, 0, 3), Vec3(0, 1, 3), Vec3(1, 1, 3), red);
  scene.addTriangle(Vec3(0, 0, 4), Vec3(0, 1, 4), Vec3(1, 1, 4), light);
  scene.addTriangle(Vec3(0, 0, 5), Vec3(0, 1, 5), Vec3(1, 1, 5), red);
  scene.addSphere(Vec3(0, 0, 30), 10, light);
  scene.setEnvironmentColour(Vec3(0.1, 0.2, 0.3));
  CHECK(scene.numTriangles() == 3 && scene.numSpheres() == 1);
  CHECK(scene.palette().size() == 2);
  CHECK(scene.triangleMaterials()[0] == 0 && scene.triangleMaterials()[1] == 1 &&
        scene.triangleMaterials()[2] == 0 && scene.sphereMaterials()[0] == 1);
  CHECK(scene.triangleVertices().size() == 27 && scene.triangleVertices()[2] == 3.0);
  CHECK(scene.sphereCentreRadius()[3] == 10.0);
  int32_t devices = 0;
  ptb200_device_count(&devices);
  if (devices == 0) { // there is no CPU fallback: the adaptor must throw, not render
    bool threw = false;
    std::string what;
    try {
      RenderParams p;
      p.width = 8;
      p.height = 6;
      p.samplesPerPixel = 1;
      Camera cam(Vec3(0, 0, 0), Vec3(0, 0, 1), Vec3(0, 1, 0).normalised(), 8, 6, 40.0);
      scene.render(cam, p, nullptr);
    } catch (const std::runtime_error &e) {
      threw = true;
      what = e.what();
    }
    CHECK(threw);
    CHECK(what.find("no CPU fallback") != std::string::npos);
  }
}

// The reference's own dod tests (test/dod/SphereTests.cpp:15-52, SceneTests.cpp:16-79,
// TriangleTests.cpp:15-42), run through the drop-in ptb200::Scene on the GPU.
static bool near(double a, double b, double tol = 1e-4) { return std::fabs(a - b) <= tol; }
static bool nearVec(const Vec3 &v, double x, double y, double z) {
  return near(v.x(), x) && near(v.y(), y) && near(v.z(), z); // ApproxVec3, 1e-4
}
static void gpuReferenceKats() {
  const double inf = std::numeric_limits<double>::infinity();
  MaterialSpec mat;
  {
    Scene scene;
    scene.addSphere(Vec3(10, 20, 30), 15, mat);
    CHECK(!scene.intersectSpheres(Ray::fromTwoPoints(Vec3(0, 0, 0), Vec3(0, 1, 0)), inf));
    CHECK(!scene.intersectSpheres(Ray::fromTwoPoints(Vec3(0, 0, 0), Vec3(-10, -20, -30)), inf));
    auto ir = scene.intersectSpheres(Ray::fromTwoPoints(Vec3(0, 0, 0), Vec3(10, 20, 30)), inf);
    CHECK(ir.has_value());
    if (ir) {
      CHECK(near(ir->hit.distance, 22.416738, 22.416738 * 1.2e-5)); // Catch Approx: 100 float epsilons, relative
      CHECK(nearVec(ir->hit.position, 5.99108, 11.9822, 17.9732));
      CHECK(nearVec(ir->hit.normal, -0.267261, -0.534522, -0.801784));
      CHECK(!ir->hit.inside);
    }
    CHECK(!scene.intersectSpheres(Ray::fromTwoPoints(Vec3(0, 0, 0), Vec3(10, 20, 30)), 22.0));
    CHECK(!scene.intersect(Ray::fromTwoPoints(Vec3(0, 0, 0), Vec3(0, 1, 0))));
  }
  {
    Scene scene;
    scene.addSphere(Vec3(0, 0, 30), 10, mat);
    auto ir = scene.intersectSpheres(Ray::fromTwoPoints(Vec3(0, 0, 0), Vec3(0, 0, 2)), inf);
    CHECK(ir && ir->hit.distance == 20 && !ir->hit.inside);
    if (ir) {
      CHECK(nearVec(ir->hit.position, 0, 0, 20));
      CHECK(nearVec(ir->hit.normal, 0, 0, -1));
    }
    ir = scene.intersectSpheres(Ray::fromTwoPoints(Vec3(0, 0, 30), Vec3(0, 0, 2)), inf);
    CHECK(ir && ir->hit.distance == 10 && ir->hit.inside);
    if (ir)
      CHECK(nearVec(ir->hit.normal, 0, 0, 1));
  }
  for (int order = 0; order < 2; ++order) { // picks the nearer of two spheres, either insertion order
    Scene scene;
    const auto m1 = MaterialSpec::makeDiffuse(Vec3(1, 1, 1)), m2 = MaterialSpec::makeDiffuse(Vec3(1, 0, 0));
    scene.addSphere(Vec3(0, 0, order ? 90 : 30), 10, m1);
    scene.addSphere(Vec3(0, 0, order ? 30 : 90), 10, m2);
    auto ir = scene.intersect(Ray::fromTwoPoints(Vec3(0, 0, 0), Vec3(0, 0, 2)));
    CHECK(ir && ir->hit.distance == 20);
    if (ir)
      CHECK(ir->material == (order ? m2 : m1));
  }
  for (int winding = 0; winding < 2; ++winding) {
    Scene scene;
    if (winding == 0)
      scene.addTriangle(Vec3(0, 0, 3), Vec3(0, 1, 3), Vec3(1, 1, 3), mat);
    else
      scene.addTriangle(Vec3(0, 0, 3), Vec3(1, 1, 3), Vec3(0, 1, 3), mat);
    CHECK(!scene.intersectTriangles(Ray::fromTwoPoints(Vec3(0, 0, 0), Vec3(0, 1, 0)), inf));
    CHECK(!scene.intersectTriangles(Ray::fromTwoPoints(Vec3(0, 0, 0), Vec3(0, 0, -1)), inf));
    auto ir = scene.intersectTriangles(Ray::fromTwoPoints(Vec3(0, 0, 0), Vec3(0, 0, 1)), inf);
    CHECK(ir.has_value());
    if (ir) {
      CHECK(near(ir->hit.distance, 3.0, 1e-9));
      CHECK(nearVec(ir->hit.position, 0, 0, 3));
      CHECK(nearVec(ir->hit.normal, 0, 0, -1));
    }
    if (winding == 0)
      CHECK(!scene.intersectTriangles(Ray::fromTwoPoints(Vec3(0, 0, 0), Vec3(0, 0, 1)), 2.999));
  }
  // render() through the adaptor: sample counts, determinism (test/seed_tests.sh), updateFunc
  Scene scene;
  scene.addSphere(Vec3(0, 0, 5), 1, MaterialSpec::makeLight(Vec3(2, 2, 2)));
  scene.addTriangle(Vec3(-5, -1, 0), Vec3(5, -1, 0), Vec3(0, -1, 10), MaterialSpec::makeDiffuse(Vec3(0.5, 0.5, 0.5)));
  scene.setEnvironmentColour(Vec3(0.1, 0.1, 0.1));
  RenderParams p;
  p.width = 16;
  p.height = 16;
  p.samplesPerPixel = 16;
  p.seed = 1;
  Camera cam(Vec3(0, 0, 0), Vec3(0, 0, 5), Vec3(0, 1, 0).normalised(), 16, 16, 50.0);
  int updates = 0;
  const ArrayOutput a = scene.render(cam, p, [&](ArrayOutput &partial) {
    ++updates;
    CHECK(partial.width() == 16 && partial.totalSamples() % 256 == 0);
  });
  CHECK(updates >= 1);
  CHECK(a.totalSamples() == 16u * 16u * 16u);
  const ArrayOutput b = scene.render(cam, p, nullptr);
  p.seed = 2;
  const ArrayOutput c = scene.render(cam, p, nullptr);
  bool same = true, differs = false;
  for (int y = 0; y < 16; ++y)
    for (int x = 0; x < 16; ++x) {
      same = same && a.rawPixelAt(x, y) == b.rawPixelAt(x, y);
      differs = differs || a.rawPixelAt(x, y) != c.rawPixelAt(x, y);
    }
  CHECK(same);
  CHECK(differs);
  CHECK(a.rawPixelAt(8, 8).x() > 1.0); // looks straight at the light
}

static void pngWriterTests() {
  char path[] = "/tmp/ptb200pngXXXXXX";
  const int fd = mkstemp(path);
  {
    PngWriter pw(path, 3, 2);
    CHECK(pw.ok());
    const std::uint8_t row[9] = {255, 0, 0, 0, 255, 0, 0, 0, 255};
    pw.addRow(row);
    pw.addRow(row);
  }
  std::ifstream in(path, std::ios::binary);
  std::vector<unsigned char> bytes((std::istreambuf_iterator<char>(in)), {});
  close(fd);
  unlink(path);
  CHECK(bytes.size() > 57);
  CHECK(bytes.size() >= 8 && std::memcmp(bytes.data(), "\x89PNG\r\n\x1a\n", 8) == 0);
  CHECK(bytes.size() >= 24 && bytes[19] == 3 && bytes[23] == 2); // IHDR width/height
}

// Reads a PTSCENE2 fixture (see pt_three_ways_b200/scenefile.py for the layout).
struct Fixture {
  std::vector<double> tri, sph, mats;
  std::vector<uint32_t> triMat, sphMat;
  double env[3];
  double camera[18];
};
static bool readFixture(const std::string &path, Fixture &f) {
  std::ifstream in(path, std::ios::binary);
  if (!in)
    return false;
  char magic[8];
  uint32_t counts[4];
  in.read(magic, 8);
  in.read(reinterpret_cast<char *>(counts), 16);
  in.read(reinterpret_cast<char *>(f.env), 24);
  f.tri.resize(counts[0] * 9);
  f.triMat.resize(counts[0]);
  f.sph.resize(counts[1] * 4);
  f.sphMat.resize(counts[1]);
  f.mats.resize(counts[2] * 9);
  in.read(reinterpret_cast<char *>(f.tri.data()), f.tri.size() * 8);
  in.read(reinterpret_cast<char *>(f.triMat.data()), f.triMat.size() * 4);
  in.read(reinterpret_cast<char *>(f.sph.data()), f.sph.size() * 8);
  in.read(reinterpret_cast<char *>(f.sphMat.data()), f.sphMat.size() * 4);
  in.read(reinterpret_cast<char *>(f.mats.data()), f.mats.size() * 8);
  in.read(reinterpret_cast<char *>(f.camera), 144);
  return static_cast<bool>(in) && std::memcmp(magic, "PTSCENE2", 8) == 0;
}

// Our loader + recipes must hand the SceneBuilder exactly what the reference's did.
static void recipeTests(const std::string &scenesDir, const std::string &fixtureDir, bool haveObj) {
  for (const char *name : {"cornell", "suzanne", "ce", "single-sphere", "multi-sphere", "example1",
                           "bbc-owl"}) {
    const bool needsObj = std::string(name) == "cornell" || std::string(name) == "suzanne" ||
                          std::string(name) == "ce";
    if (needsObj && !haveObj)
      continue;
    Fixture f;
    if (!readFixture(fixtureDir + "/" + name + ".ptscene", f)) {
      CHECK(false && "fixture unreadable");
      continue;
    }
    HostApi api(scenesDir);
    Scene scene;
    Camera cam = SceneRecipes<HostApi>::create(api, scene, name, 64, 48);
    CHECK(scene.triangleVertices() == f.tri);
    CHECK(scene.sphereCentreRadius() == f.sph);
    CHECK(scene.environment() == Vec3(f.env[0], f.env[1], f.env[2]));
    CHECK(scene.numTriangles() == f.triMat.size() && scene.numSpheres() == f.sphMat.size());
    // material per primitive, by value (palette order may differ)
    bool materialsEqual = scene.numTriangles() == f.triMat.size() && scene.numSpheres() == f.sphMat.size();
    auto sameMaterial = [&](const MaterialSpec &m, uint32_t fixtureIndex) {
      // The fixture comes from the reference built with -funsafe-math-optimizations, which
      // may reassociate MaterialSpec::toRadians (angle / 360 * 2 * pi); allow 4 ulp.
      const PtMaterial a = m.abi();
      double mine[9];
      std::memcpy(mine, &a, 72);
      for (int k = 0; k < 9; ++k) {
        const double want = f.mats[9 * fixtureIndex + static_cast<size_t>(k)];
        if (std::fabs(mine[k] - want) > 4 * 2.3e-16 * std::fabs(want))
          return false;
      }
      return true;
    };
    for (size_t i = 0; materialsEqual && i < f.triMat.size(); ++i)
      materialsEqual = sameMaterial(scene.palette()[scene.triangleMaterials()[i]], f.triMat[i]);
    for (size_t i = 0; materialsEqual && i < f.sphMat.size(); ++i)
      materialsEqual = sameMaterial(scene.palette()[scene.sphereMaterials()[i]], f.sphMat[i]);
    CHECK(materialsEqual);
    if (!materialsEqual)
      std::cerr << "  scene " << name << "\n";
    // camera: our constructor vs the reference's, 4 ulp (the reference build reassociates)
    double ours[18];
    std::memcpy(ours, &cam.abi(), 144);
    bool cameraClose = true;
    for (int i = 0; i < 18; ++i)
      cameraClose = cameraClose && std::fabs(ours[i] - f.camera[i]) <= 4 * 2.3e-16 * std::fabs(f.camera[i]);
    CHECK(cameraClose);
  }
}

// The lazily seeded per-lane mt19937 of the `fp` way (csrc/pt_mt19937.cuh) against
// std::mt19937: first generation from the two running seed words, the prefetch hand-over of
// the camera draws, and the in-place later generations.
static void laneMt19937Tests() {
  const uint32_t noLimit = mtStoreLimit(100000);
  CHECK(noLimit == 0xffffffffu);
  for (uint32_t seed : {0u, 1u, 5489u, 0xffffffffu, 123456789u, 640u * 480u * 7u + 17u}) {
    for (uint32_t cameraWords : {4u, 8u}) {
      std::mt19937 reference(seed);
      LaneMt19937 rng;
      uint32_t history[kMtHistoryWords];
      for (uint32_t &word : history)
        word = 0xdeadbeefu;
      rng.seed(seed);
      int bad = 0;
      for (uint32_t i = 0; i < cameraWords; ++i)
        bad += rng.word<true>(history, noLimit) != reference();
      for (uint32_t i = 0; i < kMtPrefetchWords; ++i) // what the megakernel does when the sample starts
        history[i] = history[kMtWords + i];
      for (int i = 0; i < 3000; ++i)
        bad += rng.word<false>(history, noLimit) != reference();
      CHECK(bad == 0);
    }
    // A sample that draws at most maxWords <= 624 words skips the stores nobody reads back.
    for (uint32_t maxWords : {9u, 100u, 227u, 228u, 235u, 488u, 623u, 624u}) {
      std::mt19937 reference(seed);
      LaneMt19937 rng;
      uint32_t history[kMtHistoryWords];
      for (uint32_t &word : history)
        word = 0xdeadbeefu;
      rng.seed(seed);
      const uint32_t limit = mtStoreLimit(maxWords);
      int bad = 0;
      for (uint32_t i = 0; i < 8; ++i)
        bad += rng.word<true>(history, limit) != reference();
      for (uint32_t i = 0; i < kMtPrefetchWords; ++i)
        history[i] = history[kMtWords + i];
      for (uint32_t i = 8; i < maxWords; ++i)
        bad += rng.word<false>(history, limit) != reference();
      CHECK(bad == 0);
      CHECK(limit >= 624 || history[limit] == 0xdeadbeefu); // really skipped
    }
  }
  // Groups of six (one (u, v, p) triple) through every regime: before/after word 227, across the
  // end of the first generation and into the in-place generations, with and without skipped stores.
  for (uint32_t seed : {7u, 0x9e3779b9u}) {
    for (uint32_t cameraWords : {4u, 8u}) {
      for (uint32_t groups : {80u, 103u, 400u}) {
        std::mt19937 reference(seed);
        LaneMt19937 rng;
        uint32_t history[kMtHistoryWords];
        for (uint32_t &word : history)
          word = 0xdeadbeefu;
        rng.seed(seed);
        const uint32_t limit = mtStoreLimit(cameraWords + 6ull * groups);
        int bad = 0;
        for (uint32_t i = 0; i < cameraWords; ++i)
          bad += rng.word<true>(history, 0u) != reference();
        for (uint32_t i = 0; i < kMtPrefetchWords; ++i)
          history[i] = history[kMtWords + i];
        for (uint32_t g = 0; g < groups; ++g) {
          uint32_t out[6] = {0, 0, 0, 0, 0, 0};
          rng.six(history, limit, out);
          for (uint32_t value : out)
            bad += value != reference();
        }
        CHECK(bad == 0);
      }
    }
  }
  std::mt19937 defaultSeeded; // the standard's own known answer: 10000th output = 4123659995
  LaneMt19937 rng;
  uint32_t history[kMtHistoryWords];
  rng.seed(5489u);
  uint32_t last = 0;
  for (int i = 0; i < 10000; ++i)
    last = rng.word<false>(history, 0xffffffffu);
  CHECK(last == 4123659995u);
}

// The exact-stream policies' engine (csrc/pt_mt19937.cuh: GroupMt19937) against std::mt19937: the
// lazy twist of sub-warp groups under every interleaving of advance(), draws and skips; the host build
// plays all lanes of a group (loads of a batch before its stores).
template <int kGroup>
static void groupMt19937Case() {
  std::mt19937 pattern(static_cast<uint32_t>(kGroup));
  for (uint32_t seed : {0u, 1u, 5489u, 123456789u}) {
    uint32_t state[624];
    GroupMt19937<kGroup> rng{state, 0, 0, 0xffffffffu, 0u};
    rng.seed(seed);
    std::mt19937 reference(seed);
    int bad = 0;
    for (int iteration = 0; iteration < 60000; ++iteration) {
      if (pattern() % 3)
        rng.advance();
      if (pattern() % 50 == 0) { // radiance() at the deepest level discards whole triples
        const uint32_t count = pattern() % 2000;
        rng.skip(count);
        reference.discard(count);
      }
      for (uint32_t draws = pattern() % 14; draws; --draws)
        bad += rng.next() != reference();
      bad += rng.regen > rng.index || (kGroup < 32 && rng.regen % kGroup != 0);
    }
    CHECK(bad == 0);
  }
}
static void groupMt19937Tests() {
  groupMt19937Case<4>();
  groupMt19937Case<8>();
  groupMt19937Case<16>();
  groupMt19937Case<32>();
}

int main(int argc, char **argv) {
  std::string scenesDir = "/root/reference/scenes", fixtureDir = "tests/golden/scenes";
  for (int i = 1; i + 1 < argc; i += 2) {
    if (std::string(argv[i]) == "--scenes")
      scenesDir = argv[i + 1];
    if (std::string(argv[i]) == "--fixtures")
      fixtureDir = argv[i + 1];
  }
  objLoaderTests();
  arrayOutputTests();
  mathTests();
  sceneAdaptorTests();
  pngWriterTests();
  laneMt19937Tests();
  groupMt19937Tests();
  int32_t deviceCount = 0;
  ptb200_device_count(&deviceCount);
  if (deviceCount > 0)
    gpuReferenceKats();
  const bool haveObj = static_cast<bool>(std::ifstream(scenesDir + "/CornellBox-Original.obj"));
  recipeTests(scenesDir, fixtureDir, haveObj);
  std::cout << checks << " checks, " << failures << " failures"
            << (deviceCount > 0 ? " (GPU KATs ran)" : " (no GPU: KATs through the adaptor skipped)")
            << (haveObj ? "" : " (OBJ-backed recipes skipped: no scenes dir)") << "\n";
  return failures ? 1 : 0;
}
