// Console progress like the reference's Progressifier (src/util/Progressifier.{h,cpp}): prints
// "<UTC time> : xx.xx% (done / total)" each time progress advances by at least 5 %.
#pragma once

#include <chrono>
#include <cstddef>
#include <cstdio>
#include <ctime>
#include <iomanip>
#include <iostream>

namespace ptb200 {

class Progressifier {
  size_t numWork_{};
  double minProgress_{5.0};
  double lastProgress_{};

public:
  explicit Progressifier(size_t numWork) noexcept : numWork_(numWork) {}

  void update(size_t numDone) noexcept {
    const double progress = numWork_ ? static_cast<double>(numDone) / static_cast<double>(numWork_) * 100 : 100.0;
    if (progress >= lastProgress_ + minProgress_) {
      const std::time_t now = std::chrono::system_clock::to_time_t(std::chrono::system_clock::now());
      std::tm utc{};
      gmtime_r(&now, &utc);
      char stamp[32];
      std::strftime(stamp, sizeof stamp, "%Y-%m-%d %H:%M:%S", &utc);
      std::cout << stamp << " : " << std::fixed << std::setprecision(2) << progress << "% (" << numDone
                << " / " << numWork_ << ")\n"
                << std::flush;
      lastProgress_ = progress;
    }
  }
  void numLeft(size_t numLeft) noexcept { update(numWork_ - numLeft); }
};

} // namespace ptb200
