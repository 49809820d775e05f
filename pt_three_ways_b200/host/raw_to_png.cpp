// raw_to_png_b200 — merges raw framebuffers and writes a PNG, like the reference's raw_to_png
// (src/main/raw_to_png.cpp:9-81): `raw_to_png_b200 out.png a.raw b.raw ...` loads each file with
// ArrayOutput::load, sums them with operator+= (partial renders from different seeds, GPUs or
// machines merge this way) and saves the 8-bit gamma-2.2 image.  The raw format is the
// reference's own (src/util/ArrayOutput.cpp:65-110), so files are interchangeable.
#include "ArrayOutput.h"
#include "PngWriter.h"

#include <iomanip>
#include <iostream>
#include <optional>
#include <string>
#include <vector>

using namespace ptb200;

int main(int argc, const char *argv[]) {
  if (argc < 2) {
    std::cerr << "Missing output filename.\nusage: " << argv[0] << " <output.png> <input.raw>...\n";
    return 1;
  }
  if (argc < 3) {
    std::cerr << "Missing inputs.\nusage: " << argv[0] << " <output.png> <input.raw>...\n";
    return 1;
  }
  const std::string outputName = argv[1];
  try {
    std::optional<ArrayOutput> accumulator;
    size_t totalSamples = 0;
    for (int i = 2; i < argc; ++i) {
      std::cout << "Loading " << argv[i] << "...\n";
      const ArrayOutput input = ArrayOutput::load(argv[i]);
      if (!accumulator) {
        std::cout << "  width: " << input.width() << " height: " << input.height() << '\n';
        accumulator.emplace(input.width(), input.height());
      }
      const size_t samples = input.totalSamples();
      totalSamples += samples;
      std::cout << "  samples: " << samples << '\n';
      if (accumulator->width() != input.width() || accumulator->height() != input.height()) {
        std::cerr << "Mismatch in size, width " << input.width() << " height " << input.height() << '\n';
        return 1;
      }
      *accumulator += input;
    }
    const double averageSpp =
        static_cast<double>(totalSamples) / (static_cast<double>(accumulator->width()) * accumulator->height());
    std::cout << "Saving " << outputName << " with " << totalSamples << " samples (" << std::fixed
              << std::setprecision(1) << averageSpp << " per pixel)...\n";
    PngWriter pw(outputName.c_str(), accumulator->width(), accumulator->height());
    if (!pw.ok()) {
      std::cerr << "Unable to save PNG\n";
      return 1;
    }
    std::vector<std::uint8_t> row(static_cast<size_t>(accumulator->width()) * 3);
    for (int y = 0; y < accumulator->height(); ++y) {
      for (int x = 0; x < accumulator->width(); ++x) {
        const auto colour = accumulator->pixelAt(x, y);
        for (int c = 0; c < 3; ++c)
          row[static_cast<size_t>(x) * 3 + c] = colour[static_cast<size_t>(c)];
      }
      pw.addRow(row.data());
    }
  } catch (const std::exception &e) {
    std::cerr << "Error: " << e.what() << '\n';
    return 1;
  }
  return 0;
}
