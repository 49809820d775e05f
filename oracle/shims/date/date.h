// TEST INFRASTRUCTURE — build shim, not product code.
//
// Stand-in for Howard Hinnant's date library (date/date.h, 2.4.1 per the reference's
// conanfile.txt), absent from /root/reference and from this image.  The reference needs
// exactly one thing from it: streaming a system_clock::time_point in
// src/util/Progressifier.cpp:3,15-16.
#pragma once

#include <chrono>
#include <ctime>
#include <ostream>

namespace date {

inline std::ostream &operator<<(std::ostream &out,
                                const std::chrono::system_clock::time_point &when) {
  const std::time_t seconds = std::chrono::system_clock::to_time_t(when);
  std::tm broken{};
  gmtime_r(&seconds, &broken);
  char text[32];
  std::strftime(text, sizeof text, "%Y-%m-%d %H:%M:%S", &broken);
  return out << text;
}

} // namespace date
