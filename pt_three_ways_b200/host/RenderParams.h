// RenderParams (mirrors src/util/RenderParams.h:3-13, same names and defaults).
#pragma once

#include "ptb200.h"

namespace ptb200 {

struct RenderParams {
  int width{1920};
  int height{1080};
  bool preview{false};
  int samplesPerPixel{40};
  int maxCpus{1};
  int maxDepth{5};
  int firstBounceUSamples{4};
  int firstBounceVSamples{4};
  int seed{0};

  [[nodiscard]] PtRenderParams abi() const {
    return PtRenderParams{width,   height,   preview ? 1 : 0,     samplesPerPixel,
                          maxCpus, maxDepth, firstBounceUSamples, firstBounceVSamples,
                          seed};
  }
};

} // namespace ptb200
