// Framebuffer of the host boundary: per pixel, the sum of sample colours and the sample count.
// Mirrors the behaviour of src/util/ArrayOutput.{h,cpp} and src/util/SampledPixel.{h,cpp}
// (same method names) including the raw interchange format of ArrayOutput::save/load
// (ArrayOutput.cpp:21-28,65-110): 16-byte header {u32 signature=1, u32 version=1, u32 height,
// u32 width} then per pixel 3 x f64 sum + u32 count, 28 bytes, no padding.  Files written
// here can be merged by the reference's raw_to_png and vice versa.
#pragma once

#include "Vec3.h"
#include "ptb200.h"

#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace ptb200 {

class SampledPixel {
  Vec3 colour_;
  size_t numSamples_{};

public:
  void accumulate(const Vec3 &sample, int num) noexcept {
    colour_ += sample;
    numSamples_ += static_cast<size_t>(num);
  }
  void accumulate(const SampledPixel &other) noexcept {
    colour_ += other.colour_;
    numSamples_ += other.numSamples_;
  }
  [[nodiscard]] Vec3 result() const noexcept {
    return numSamples_ == 0 ? colour_ : colour_ * (1.0 / static_cast<double>(numSamples_));
  }
  [[nodiscard]] Vec3 rawResult() const noexcept { return colour_; }
  [[nodiscard]] size_t numSamples() const noexcept { return numSamples_; }
};

class ArrayOutput {
  int width_;
  int height_;
  std::vector<SampledPixel> output_;

  [[nodiscard]] size_t indexOf(int x, int y) const noexcept {
    return static_cast<size_t>(x) + static_cast<size_t>(y) * static_cast<size_t>(width_);
  }

public:
  using Pixel = std::array<std::uint8_t, 3>;

  ArrayOutput(int width, int height)
      : width_(width), height_(height),
        output_(static_cast<size_t>(width) * static_cast<size_t>(height)) {}

  [[nodiscard]] int width() const noexcept { return width_; }
  [[nodiscard]] int height() const noexcept { return height_; }

  void addSamples(int x, int y, const Vec3 &colour, int numSamples) noexcept {
    output_[indexOf(x, y)].accumulate(colour, numSamples);
  }
  // Bulk version of addSamples for what the device returns.
  void addSamples(const PtPixel *pixels) noexcept {
    for (size_t i = 0; i < output_.size(); ++i)
      output_[i].accumulate(Vec3(pixels[i].sum[0], pixels[i].sum[1], pixels[i].sum[2]),
                            static_cast<int>(pixels[i].numSamples));
  }
  [[nodiscard]] Vec3 rawPixelAt(int x, int y) const noexcept {
    return output_[indexOf(x, y)].result();
  }
  [[nodiscard]] const SampledPixel &sampledPixelAt(int x, int y) const noexcept {
    return output_[indexOf(x, y)];
  }
  // 8-bit sRGB-ish: clamp, gamma 2.2, round (ArrayOutput.cpp:9-12,32-37).
  [[nodiscard]] Pixel pixelAt(int x, int y) const noexcept {
    const Vec3 raw = rawPixelAt(x, y);
    auto toByte = [](double c) {
      return static_cast<std::uint8_t>(
          std::lround(std::pow(std::clamp(c, 0.0, 1.0), 1.0 / 2.2) * 255));
    };
    return Pixel{toByte(raw.x()), toByte(raw.y()), toByte(raw.z())};
  }

  ArrayOutput &operator+=(const ArrayOutput &rhs) {
    if (rhs.width_ != width_ || rhs.height_ != height_)
      throw std::logic_error("Two differently-sized arrays were attempted to be combined");
    for (size_t i = 0; i < output_.size(); ++i)
      output_[i].accumulate(rhs.output_[i]);
    return *this;
  }

  [[nodiscard]] size_t totalSamples() const noexcept {
    size_t total = 0;
    for (const auto &pixel : output_)
      total += pixel.numSamples();
    return total;
  }

  void save(const std::string &filename) const {
    std::unique_ptr<FILE, int (*)(FILE *)> out(std::fopen(filename.c_str(), "wb"), &std::fclose);
    if (!out)
      throw std::runtime_error("Unable to open " + filename);
    std::vector<unsigned char> bytes(16 + output_.size() * 28);
    const std::uint32_t header[4] = {1u, 1u, static_cast<std::uint32_t>(height_),
                                     static_cast<std::uint32_t>(width_)};
    std::memcpy(bytes.data(), header, 16);
    unsigned char *cursor = bytes.data() + 16;
    for (const auto &pixel : output_) {
      const Vec3 sum = pixel.rawResult();
      const double rgb[3] = {sum.x(), sum.y(), sum.z()};
      const auto count = static_cast<std::uint32_t>(pixel.numSamples());
      std::memcpy(cursor, rgb, 24);
      std::memcpy(cursor + 24, &count, 4);
      cursor += 28;
    }
    if (std::fwrite(bytes.data(), 1, bytes.size(), out.get()) != bytes.size())
      throw std::runtime_error("Unable to write to " + filename);
  }

  [[nodiscard]] static ArrayOutput load(const std::string &filename) {
    std::unique_ptr<FILE, int (*)(FILE *)> in(std::fopen(filename.c_str(), "rb"), &std::fclose);
    if (!in)
      throw std::runtime_error("Unable to open " + filename);
    std::uint32_t header[4];
    if (std::fread(header, 1, 16, in.get()) != 16)
      throw std::runtime_error("Unable to read from " + filename);
    if (header[0] != 1u)
      throw std::runtime_error("Bad file " + filename + " : bad signature");
    if (header[1] != 1u)
      throw std::runtime_error("Bad file " + filename + " : bad version");
    ArrayOutput result(static_cast<int>(header[3]), static_cast<int>(header[2]));
    std::vector<unsigned char> bytes(result.output_.size() * 28);
    if (std::fread(bytes.data(), 1, bytes.size(), in.get()) != bytes.size())
      throw std::runtime_error("Unable to read from " + filename);
    const unsigned char *cursor = bytes.data();
    for (auto &pixel : result.output_) {
      double rgb[3];
      std::uint32_t count;
      std::memcpy(rgb, cursor, 24);
      std::memcpy(&count, cursor + 24, 4);
      pixel.accumulate(Vec3(rgb[0], rgb[1], rgb[2]), static_cast<int>(count));
      cursor += 28;
    }
    return result;
  }
};

} // namespace ptb200
