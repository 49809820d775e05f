// Minimal dependency-free PNG writer (8-bit RGB) for the CLI driver: the reference writes PNGs
// through libpng (src/main/PngWriter.{h,cpp}), which this image does not have.  Rows are stored
// with filter 0 inside "stored" (uncompressed) deflate blocks; CRC-32 and Adler-32 are computed
// here.  Same usage shape as the reference's writer: construct, ok(), addRow() per row.
#pragma once

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace ptb200 {

class PngWriter {
  FILE *file_{nullptr};
  int width_;
  int height_;
  std::vector<std::uint8_t> raw_; // filter byte + RGB per row

  static std::uint32_t crc32(const std::uint8_t *data, size_t n, std::uint32_t crc = 0) {
    static std::uint32_t table[256];
    static bool ready = false;
    if (!ready) {
      for (std::uint32_t i = 0; i < 256; ++i) {
        std::uint32_t c = i;
        for (int k = 0; k < 8; ++k)
          c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        table[i] = c;
      }
      ready = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i)
      crc = table[(crc ^ data[i]) & 0xffu] ^ (crc >> 8);
    return ~crc;
  }
  static void be32(std::vector<std::uint8_t> &out, std::uint32_t v) {
    for (int shift = 24; shift >= 0; shift -= 8)
      out.push_back(static_cast<std::uint8_t>(v >> shift));
  }
  void chunk(const char type[4], const std::vector<std::uint8_t> &payload) {
    std::vector<std::uint8_t> buf;
    be32(buf, static_cast<std::uint32_t>(payload.size()));
    buf.insert(buf.end(), type, type + 4);
    buf.insert(buf.end(), payload.begin(), payload.end());
    be32(buf, crc32(buf.data() + 4, buf.size() - 4));
    std::fwrite(buf.data(), 1, buf.size(), file_);
  }

public:
  PngWriter(const char *filename, int width, int height)
      : file_(std::fopen(filename, "wb")), width_(width), height_(height) {
    raw_.reserve(static_cast<size_t>(height) * (static_cast<size_t>(width) * 3 + 1));
  }
  PngWriter(const PngWriter &) = delete;
  PngWriter &operator=(const PngWriter &) = delete;
  [[nodiscard]] bool ok() const noexcept { return file_ != nullptr; }

  void addRow(const std::uint8_t *rgb) {
    raw_.push_back(0); // filter: none
    raw_.insert(raw_.end(), rgb, rgb + static_cast<size_t>(width_) * 3);
  }

  ~PngWriter() {
    if (!file_)
      return;
    static const std::uint8_t signature[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::fwrite(signature, 1, 8, file_);
    std::vector<std::uint8_t> header;
    be32(header, static_cast<std::uint32_t>(width_));
    be32(header, static_cast<std::uint32_t>(height_));
    header.insert(header.end(), {8, 2, 0, 0, 0}); // 8 bits, RGB, deflate, no filter, no interlace
    chunk("IHDR", header);
    std::vector<std::uint8_t> z = {0x78, 0x01};
    std::uint32_t a = 1, b = 0;
    for (std::uint8_t byte : raw_) {
      a = (a + byte) % 65521u;
      b = (b + a) % 65521u;
    }
    for (size_t pos = 0; pos < raw_.size() || pos == 0;) {
      const size_t n = std::min<size_t>(65535, raw_.size() - pos);
      const bool last = pos + n >= raw_.size();
      z.push_back(last ? 1 : 0);
      z.push_back(static_cast<std::uint8_t>(n & 0xff));
      z.push_back(static_cast<std::uint8_t>(n >> 8));
      z.push_back(static_cast<std::uint8_t>(~n & 0xff));
      z.push_back(static_cast<std::uint8_t>((~n >> 8) & 0xff));
      z.insert(z.end(), raw_.begin() + static_cast<long>(pos), raw_.begin() + static_cast<long>(pos + n));
      pos += n;
      if (last)
        break;
    }
    be32(z, (b << 16) | a);
    chunk("IDAT", z);
    chunk("IEND", {});
    std::fclose(file_);
  }
};

} // namespace ptb200
