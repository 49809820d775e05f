// pt_b200 — command-line driver with the reference CLI's flags (src/main/main.cpp:382-404):
//   -w/--width -h/--height --max-cpus --spp --first-bounce-u --first-bounce-v --max-depth
//   --seed --preview --save-every --way --scene --raw <output>
// (--way dod|fp|oo) plus backend flags:  --rng keyed|exact   --gpus N (0 = all)   --scenes DIR   --device K
//   --ptscene FILE (a PTSCENE2 fixture instead of --scene: the reference loader's own output,
//   usable where the OBJ files are not available).
// Scene building, OBJ/MTL loading, the framebuffer and the PNG/raw writers are host C++ here
// as they are in the reference; render() goes to the GPU through the C ABI.
#include "ArrayOutput.h"
#include "HostApi.h"
#include "PngWriter.h"
#include "Progressifier.h"
#include "Scene.h"
#include "SceneRecipes.h"

#include <chrono>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <iostream>
#include <random>
#include <string>
#include <vector>

using namespace ptb200;

namespace {

void savePng(const ArrayOutput &output, const std::string &name) {
  PngWriter pw(name.c_str(), output.width(), output.height());
  if (!pw.ok()) {
    std::cerr << "Unable to save PNG\n";
    return;
  }
  std::vector<std::uint8_t> row(static_cast<size_t>(output.width()) * 3);
  for (int y = 0; y < output.height(); ++y) {
    for (int x = 0; x < output.width(); ++x) {
      const auto colour = output.pixelAt(x, y);
      for (int c = 0; c < 3; ++c)
        row[static_cast<size_t>(x) * 3 + c] = colour[static_cast<size_t>(c)];
    }
    pw.addRow(row.data());
  }
}

// Reads a PTSCENE2 fixture (layout: pt_three_ways_b200/scenefile.py) through the SceneBuilder
// calls; the stored camera is the recipe's for 64x48 and only three fields depend on the size.
Camera loadPtScene(const std::string &path, Scene &scene, int width, int height) {
  std::ifstream in(path, std::ios::binary);
  if (!in)
    throw std::runtime_error("Unable to open " + path);
  char magic[8];
  uint32_t counts[4];
  double env[3];
  in.read(magic, 8);
  in.read(reinterpret_cast<char *>(counts), sizeof counts);
  in.read(reinterpret_cast<char *>(env), sizeof env);
  if (!in || std::string(magic, 8) != "PTSCENE2")
    throw std::runtime_error("Bad file " + path + " : not a PTSCENE2 scene");
  std::vector<double> tri(static_cast<size_t>(counts[0]) * 9), sph(static_cast<size_t>(counts[1]) * 4),
      mats(static_cast<size_t>(counts[2]) * 9);
  std::vector<uint32_t> triMat(counts[0]), sphMat(counts[1]);
  in.read(reinterpret_cast<char *>(tri.data()), static_cast<std::streamsize>(tri.size() * 8));
  in.read(reinterpret_cast<char *>(triMat.data()), static_cast<std::streamsize>(triMat.size() * 4));
  in.read(reinterpret_cast<char *>(sph.data()), static_cast<std::streamsize>(sph.size() * 8));
  in.read(reinterpret_cast<char *>(sphMat.data()), static_cast<std::streamsize>(sphMat.size() * 4));
  in.read(reinterpret_cast<char *>(mats.data()), static_cast<std::streamsize>(mats.size() * 8));
  PtCamera cam{};
  in.read(reinterpret_cast<char *>(&cam), sizeof cam);
  if (!in)
    throw std::runtime_error("Bad file " + path + " : truncated");
  auto material = [&](uint32_t i) {
    const double *m = &mats[9 * static_cast<size_t>(i)];
    return MaterialSpec{Vec3(m[0], m[1], m[2]), Vec3(m[3], m[4], m[5]), m[6], m[7], m[8]};
  };
  for (uint32_t i = 0; i < counts[0]; ++i) {
    const double *t = &tri[9 * static_cast<size_t>(i)];
    scene.addTriangle(Vec3(t[0], t[1], t[2]), Vec3(t[3], t[4], t[5]), Vec3(t[6], t[7], t[8]), material(triMat[i]));
  }
  for (uint32_t i = 0; i < counts[1]; ++i) {
    const double *s = &sph[4 * static_cast<size_t>(i)];
    scene.addSphere(Vec3(s[0], s[1], s[2]), s[3], material(sphMat[i]));
  }
  scene.setEnvironmentColour(Vec3(env[0], env[1], env[2]));
  cam.aspectRatio = static_cast<double>(width) / height; // Camera.h:43,46
  cam.reciprocalHeight = 1.0 / height;
  cam.reciprocalWidth = 1.0 / width;
  return Camera::fromAbi(cam);
}

int usage(const char *argv0) {
  std::cerr << "usage: " << argv0
            << " [-w W] [-h H] [--spp N] [--max-cpus N] [--first-bounce-u N] [--first-bounce-v N]\n"
               "       [--max-depth N] [--seed N] [--preview] [--save-every SECS] [--way dod|fp|oo]\n"
               "       [--scene NAME] [--raw] [--rng keyed|exact] [--gpus N] [--device K]\n"
               "       [--lanes-per-pass 0|4|8|16|32] [--scenes DIR] [--ptscene FILE] <output>\n"
               "       --merge <a.raw> <b.raw>... [--raw] <output>   (sum framebuffers instead of rendering)\n"
               "  --way dod (the default here; the reference's default is oo) renders the dod estimator with\n"
               "  --rng keyed (default) or exact; --gpus N uses devices [--device, --device + N), 0 = all;\n"
               "  --lanes-per-pass: how many lanes share a pass of the exact stream (0 = chosen from --spp).\n";
  return 1;
}

} // namespace

int main(int argc, const char *argv[]) {
  RenderParams renderParams;
  bool raw = false;
  int saveEvery = 30;
  int gpus = 1;
  int device = 0;
  int lanesPerPass = 0;
  std::string way = "dod";
  std::string sceneName = "cornell";
  std::string scenesDir = "scenes";
  std::string rng = "keyed";
  std::string ptsceneFile;
  std::string outputName;
  std::vector<std::string> mergeInputs; // --merge: sum raw framebuffers instead of rendering

  for (int i = 1; i < argc; ++i) {
    const std::string arg = argv[i];
    auto value = [&]() -> std::string {
      if (i + 1 >= argc) {
        std::cerr << "Error in command line: missing value for " << arg << '\n';
        std::exit(1);
      }
      return argv[++i];
    };
    if (arg == "-w" || arg == "--width") renderParams.width = std::stoi(value());
    else if (arg == "-h" || arg == "--height") renderParams.height = std::stoi(value());
    else if (arg == "--max-cpus") renderParams.maxCpus = std::stoi(value());
    else if (arg == "--spp") renderParams.samplesPerPixel = std::stoi(value());
    else if (arg == "--first-bounce-u") renderParams.firstBounceUSamples = std::stoi(value());
    else if (arg == "--first-bounce-v") renderParams.firstBounceVSamples = std::stoi(value());
    else if (arg == "--max-depth") renderParams.maxDepth = std::stoi(value());
    else if (arg == "--seed") renderParams.seed = std::stoi(value());
    else if (arg == "--preview") renderParams.preview = true;
    else if (arg == "--save-every") saveEvery = std::stoi(value());
    else if (arg == "--way") way = value();
    else if (arg == "--scene") sceneName = value();
    else if (arg == "--raw") raw = true;
    else if (arg == "--rng") rng = value();
    else if (arg == "--gpus") gpus = std::stoi(value());
    else if (arg == "--device") device = std::stoi(value());
    else if (arg == "--lanes-per-pass") lanesPerPass = std::stoi(value());
    else if (arg == "--scenes") scenesDir = value();
    else if (arg == "--ptscene") ptsceneFile = value();
    else if (arg == "--merge") {
      while (i + 1 < argc && std::string(argv[i + 1]).rfind("--", 0) != 0 && i + 2 < argc)
        mergeInputs.push_back(argv[++i]); // every following name but the last one (the output)
    }
    else if (arg == "--help" || arg == "-?") return usage(argv[0]);
    else if (!arg.empty() && arg[0] == '-') {
      std::cerr << "Error in command line: unknown option " << arg << '\n';
      return 1;
    } else outputName = arg;
  }
  if (outputName.empty()) {
    std::cerr << "Missing output filename.\n";
    return usage(argv[0]);
  }
  // --way dod (default): the dod estimator with --rng keyed|exact.  --way fp: the reference's
  // fp way (src/fp/Render.cpp), one mt19937 per pass and pixel — exact and parallel.  --way oo:
  // the reference's oo way (src/oo/Renderer.cpp), always its exact per-pass stream.
  if (way != "dod" && way != "b200" && way != "fp" && way != "oo") {
    std::cerr << "Unknown way " << way << "\n"; // main.cpp:365
    return 1;
  }
  if (rng != "keyed" && rng != "exact") {
    std::cerr << "Unknown rng " << rng << " (keyed: counter-based Philox, parallel over pixels; exact: the "
                 "reference's mt19937 stream, parallel over passes)\n";
    return 1;
  }
  if (gpus < 0) {
    std::cerr << "Error in command line: --gpus must be 0 (all devices) or a positive count\n";
    return 1;
  }
  if (renderParams.seed == 0) { // main.cpp:426-429
    std::random_device rd;
    renderParams.seed = static_cast<int>(rd());
  }

  std::function<void(const ArrayOutput &)> save;
  if (raw)
    save = [outputName](const ArrayOutput &output) { output.save(outputName); };
  else
    save = [outputName](const ArrayOutput &output) { savePng(output, outputName); };

  if (!mergeInputs.empty()) {
    // Partial renders (other seeds, GPUs, machines; the reference's own --raw files included)
    // are summed pixel by pixel — what the reference's raw_to_png tool does with
    // ArrayOutput::load and operator+= (src/main/raw_to_png.cpp:39-59) — and written as PNG or,
    // with --raw, as one raw file again.
    try {
      ArrayOutput merged = ArrayOutput::load(mergeInputs.front());
      for (size_t k = 1; k < mergeInputs.size(); ++k)
        merged += ArrayOutput::load(mergeInputs[k]); // throws std::logic_error on a size mismatch
      const double perPixel = static_cast<double>(merged.totalSamples()) /
                              (static_cast<double>(merged.width()) * static_cast<double>(merged.height()));
      std::cout << "Merged " << mergeInputs.size() << " framebuffers of " << merged.width() << "x" << merged.height()
                << ": " << merged.totalSamples() << " samples, " << perPixel << " per pixel\n";
      save(merged);
    } catch (const std::exception &e) {
      std::cerr << "Error: " << e.what() << '\n';
      return 1;
    }
    return 0;
  }

  try {
    using namespace std::chrono_literals;
    const auto every = std::chrono::seconds(saveEvery);
    auto nextSave = std::chrono::system_clock::now() + every;
    // The reference prints progress from inside render() (Scene.cpp:233,244); the GPU backend
    // reports after each collected batch of passes, through the same updateFunc.
    Progressifier progressifier(static_cast<size_t>(renderParams.samplesPerPixel));
    const size_t pixelCount = static_cast<size_t>(renderParams.width) * static_cast<size_t>(renderParams.height);
    auto throttledSave = [&](ArrayOutput &output) { // main.cpp:331-343
      progressifier.update(output.totalSamples() / pixelCount);
      if (every == 0s)
        return;
      const auto now = std::chrono::system_clock::now();
      if (now > nextSave) {
        save(output);
        nextSave = now + every;
      }
    };

    HostApi api(scenesDir);
    Scene scene;
    Camera camera = ptsceneFile.empty()
                        ? SceneRecipes<HostApi>::create(api, scene, sceneName, renderParams.width,
                                                        renderParams.height)
                        : loadPtScene(ptsceneFile, scene, renderParams.width, renderParams.height);
    std::cout << "Scene contains " << scene.numTriangles() << " triangles and "
              << scene.numSpheres() << " spheres.\n"; // main.cpp:320-323
    scene.setRngMode(way == "fp"   ? PTB200_RNG_MT19937_PER_PIXEL
                     : way == "oo" ? PTB200_RNG_MT19937_SEQUENTIAL_OO
                     : rng == "exact" ? PTB200_RNG_MT19937_SEQUENTIAL : PTB200_RNG_KEYED_PHILOX);
    scene.setDevice(device);
    scene.setLanesPerPass(lanesPerPass);
    if (gpus == 0) { // every visible device
      scene.setUseAllDevices(true);
    } else if (gpus > 1) { // devices [device, device + gpus)
      std::vector<int32_t> devices;
      for (int g = 0; g < gpus; ++g)
        devices.push_back(device + g);
      scene.setDevices(devices);
    }

    const auto start = std::chrono::system_clock::now();
    ArrayOutput output = scene.render(camera, renderParams, throttledSave);
    const auto end = std::chrono::system_clock::now();
    save(output);

    const auto taken = end - start;
    const auto totalSamples = output.totalSamples();
    const auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(taken).count();
    std::cout << "Took " << std::chrono::duration_cast<std::chrono::seconds>(taken).count() << "s\n";
    std::cout << "Total samples: " << totalSamples << "\n";
    std::cout << "Samples/ms: " << static_cast<double>(totalSamples) / static_cast<double>(ms ? ms : 1)
              << "\n"; // main.cpp:464-473
    const PtStats &stats = scene.lastStats();
    std::cout << "GPU: " << stats.casts << " ray casts, " << stats.kernelMs << " ms in kernels\n";
  } catch (const std::exception &e) {
    std::cerr << "Error: " << e.what() << '\n';
    return 1;
  }
  return 0;
}
