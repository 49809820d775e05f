// Two std::mt19937 engines: LaneMt19937 (one per lane, the `fp` way) and, at the end of the file,
// GroupMt19937 (one per pass, shared by the lanes that walk the pass: the exact-stream policies).
//
// std::mt19937 for ONE LANE, for the reference's `fp` way, which seeds a fresh engine per
// (pass, pixel) (src/fp/Render.cpp:125-126) and then draws at most a few hundred words from it.
//
// A textbook engine needs 624 words of state per lane, written once by the seeding recurrence and
// twisted in place.  A sample of the default configuration draws <= 488 words, all of the first
// generation, and word k of that generation depends on three SEED words only:
//     new[k] = (k < 227 ? seed[k+397] : new[k-227]) ^ twist(seed[k], seed[k+1])
// The seed words come from a first-order recurrence (x[i] = 1812433253*(x[i-1]^(x[i-1]>>30)) + i),
// so the engine keeps two running copies of it in registers — `a` = seed[k] and `b` = seed[k+397]
// (397 steps ahead, paid once per sample) — and only the words it has GENERATED go to the lane's
// history array (read back 227 draws later, and by later generations).  From word 624 on it is
// the textbook in-place algorithm on that array.
//
// The history array is a per-thread CONTIGUOUS slice of a global scratch buffer, not thread-local
// memory: local memory interleaves the lanes of a warp word by word, so lanes at different draw
// counts dirty one 32-byte sector per 4-byte word (measured: 31 KB of DRAM traffic per sample,
// 2.4 TB/s, long-scoreboard bound); a contiguous slice fills a sector with eight consecutive
// draws of one lane and the touched prefix of all resident lanes (~1 KB each) stays in L2.
// When a sample can draw at most `maxWords` <= 624 words, word k is read back only if
// k + 227 < maxWords, so later words are not stored at all (`storeLimit`).
//
// Compiled by nvcc for the kernels and by g++ for tests/host unit test (PT_HD empty).
#pragma once

#include <cstdint>

#ifdef __CUDACC__
#define PT_HD __host__ __device__ __forceinline__
#else
#define PT_HD inline
#endif

namespace ptb200 {

constexpr uint32_t kMtWords = 624;
constexpr uint32_t kMtShift = 397;
// The history array has 8 extra words: the first words of the PREFETCHED next sample (its
// camera-ray draws) wait there while the current sample still owns [0, 624).
constexpr uint32_t kMtPrefetchWords = 8;
constexpr uint32_t kMtHistoryWords = kMtWords + kMtPrefetchWords;
constexpr uint32_t kMtHistoryStride = 640; // words per thread in the scratch buffer (128-byte multiple)

// Words [storeLimit, 624) of a sample's first generation are never read back.
inline uint32_t mtStoreLimit(uint64_t maxWordsPerSample) {
  if (maxWordsPerSample > kMtWords)
    return 0xffffffffu;
  const uint32_t readBack = maxWordsPerSample > kMtWords - kMtShift
                                ? static_cast<uint32_t>(maxWordsPerSample) - (kMtWords - kMtShift)
                                : 0u;
  return readBack > kMtPrefetchWords ? readBack : kMtPrefetchWords;
}

struct LaneMt19937 {
  uint32_t a; // seed[k]      (meaningful while k < 624)
  uint32_t b; // seed[k+397]  (meaningful while k < 227)
  uint32_t k; // words drawn so far

  static PT_HD uint32_t lcg(uint32_t x, uint32_t i) { return 1812433253u * (x ^ (x >> 30)) + i; }
  static PT_HD uint32_t twist(uint32_t upper, uint32_t lower) {
    const uint32_t y = (upper & 0x80000000u) | (lower & 0x7fffffffu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
  }
  static PT_HD uint32_t temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }

  PT_HD void seed(uint32_t value) { // mersenne_twister_engine::seed(value), lazily
    a = value;
    uint32_t x = value;
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (uint32_t i = 1; i <= kMtShift; ++i)
      x = lcg(x, i);
    b = x;
    k = 0;
  }

  // Next 32-bit output.  `store` is where generated word k is kept: history + k, except for the
  // prefetched camera draws (history + 624 + k, moved to the front when the sample starts).
  template <bool kPrefetch>
  PT_HD uint32_t word(uint32_t *history, uint32_t storeLimit) {
    const uint32_t j = k % kMtWords;
    uint32_t w;
    if (k < kMtWords) {
      const uint32_t next = k + 1 < kMtWords ? lcg(a, k + 1) : history[0]; // word 623 pairs with NEW word 0
      uint32_t source;
      if (k < kMtWords - kMtShift) {
        source = b;
        b = lcg(b, k + kMtShift + 1); // runs past seed[623] at k == 226; never read after that
      } else {
        source = history[k - (kMtWords - kMtShift)];
      }
      w = source ^ twist(a, next);
      a = next;
    } else { // later generations: in place, as the serial algorithm does it
      w = history[(j + kMtShift) % kMtWords] ^ twist(history[j], history[(j + 1) % kMtWords]);
    }
    if (kPrefetch)
      history[kMtWords + j] = w;
    else if (k < storeLimit)
      history[j] = w;
    ++k;
    return temper(w);
  }

  // The next six outputs (three doubles: one (u, v, p) triple).  Straight-line and free of
  // divergence while the six words lie inside the first generation (k + 6 <= 623, every group
  // of the default configuration): lanes before and after word 227 differ by a predicated load.
  PT_HD void six(uint32_t *history, uint32_t storeLimit, uint32_t (&out)[6]) {
    if (k + 6 <= kMtWords - 1) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
      for (int i = 0; i < 6; ++i) {
        const uint32_t next = lcg(a, k + 1);
        const uint32_t source = k < kMtWords - kMtShift ? b : history[k - (kMtWords - kMtShift)];
        b = lcg(b, k + kMtShift + 1); // unused once k >= 227
        const uint32_t w = source ^ twist(a, next);
        if (k < storeLimit)
          history[k] = w;
        a = next;
        ++k;
        out[i] = temper(w);
      }
    } else {
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
      for (int i = 0; i < 6; ++i) { // shift the outputs through: no dynamic register indexing
        const uint32_t drawn = word<false>(history, storeLimit);
        out[0] = out[1];
        out[1] = out[2];
        out[2] = out[3];
        out[3] = out[4];
        out[4] = out[5];
        out[5] = drawn;
      }
    }
  }
};

// std::mt19937 with its state in shared memory, one instance per pass; `index` and `regen` are uniform
// over the group's lanes.
//
// Sub-warp groups twist LAZILY.  The textbook engine regenerates all 624 words when the last one has
// been consumed — 78 batches on 8 lanes, while the other groups of the warp wait.  Word j of the next
// generation only needs old[j], old[j+1] and (j < 227 ? old : new)[(j+397) mod 624], so a consumed word
// can be regenerated at once, in index order: words [0, regen) already belong to the next generation,
// [regen, index) are consumed, [index, 624) are still to be drawn.  advance() runs where the warp's
// groups are together (the top of the state machine's loop) and regenerates a full batch of kGroup
// words behind `index`: one 19-instruction batch per iteration serves every group of the warp at once.
// wrap() completes the generation when `index` reaches 624.  A whole warp (kGroup == 32: 624 is not a
// multiple of the batch) keeps the bulk twist.
template <int kGroup>
struct GroupMt19937 {
  uint32_t *state; // 624 words
  int index;       // next word to draw
  int regen;       // next word to regenerate (kGroup < 32)
  unsigned mask;   // the lanes of this group
  unsigned glane;  // lane within the group

  PT_HD void seed(uint32_t value) {
    if (glane == 0) {
      uint32_t x = value;
      state[0] = x;
      for (int i = 1; i < 624; ++i) {
        x = 1812433253u * (x ^ (x >> 30)) + static_cast<uint32_t>(i);
        state[i] = x;
      }
    }
    index = 624; // every word "consumed": the first draw twists the seed words
    regen = 0;
    syncGroup();
  }
  // Twists words [first, first + kGroup) in place; loads precede stores within the batch, so word i
  // sees old[i], old[i+1] and (i < 227 ? old : new)[i+397 mod 624] as the serial algorithm does (a
  // batch is at most 32 < 227 words).  Only a whole warp's last batch runs past 624.
  static PT_HD uint32_t twistedWord(const uint32_t *state, int i) {
    const int next = i + 1 == 624 ? 0 : i + 1;
    const int far = i < 227 ? i + 397 : i - 227;
    const uint32_t y = (state[i] & 0x80000000u) | (state[next] & 0x7fffffffu);
    return state[far] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
  }
  PT_HD void syncGroup() {
#ifdef __CUDA_ARCH__
    __syncwarp(mask);
#endif
  }
  PT_HD void twistBatch(int first) {
#ifdef __CUDA_ARCH__
    const int i = first + static_cast<int>(glane);
    uint32_t value = 0;
    if (kGroup < 32 || i < 624)
      value = twistedWord(state, i);
    __syncwarp(mask);
    if (kGroup < 32 || i < 624)
      state[i] = value;
    __syncwarp(mask);
#else // the host unit test plays every lane of the group: all loads, then all stores
    uint32_t values[kGroup];
    for (int lane = 0; lane < kGroup; ++lane)
      values[lane] = first + lane < 624 ? twistedWord(state, first + lane) : 0u;
    for (int lane = 0; lane < kGroup; ++lane)
      if (first + lane < 624)
        state[first + lane] = values[lane];
#endif
  }
  // Regenerates a full batch of consumed words if there is one (two for four-lane groups: a path
  // iteration draws six words or so).  EVERY lane of the warp must call it together: the batch is
  // predicated, not branched, so that the barrier between its loads and its stores is the plain
  // full-warp one (a __syncwarp on a group's mask is emulated by a loop over the warp's masks).
  // What it leaves behind, wrap() completes.
  PT_HD void advance() {
    if (kGroup < 32) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
      for (int batch = 0; batch < (kGroup >= 8 ? 1 : 8 / kGroup); ++batch) {
        const bool due = index - regen >= kGroup;
#ifdef __CUDA_ARCH__
        const int i = regen + static_cast<int>(glane);
        const uint32_t value = due ? twistedWord(state, i) : 0u;
        __syncwarp();
        if (due)
          state[i] = value;
        __syncwarp();
#else
        if (due)
          twistBatch(regen);
#endif
        regen += due ? kGroup : 0;
      }
    }
  }
  // index == 624: completes the generation and starts drawing from it.
#ifdef __CUDACC__
  __host__ __device__ __noinline__
#endif
  void wrap() {
    for (int first = kGroup < 32 ? regen : 0; first < 624; first += kGroup)
      twistBatch(first);
    index = 0;
    regen = 0;
  }
  static PT_HD uint32_t temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
  PT_HD uint32_t next() {
    if (index >= 624)
      wrap();
    return temper(state[index++]);
  }
  // Discards `count` outputs.
  PT_HD void skip(uint32_t count) {
    while (count) {
      if (index >= 624)
        wrap();
      const uint32_t room = static_cast<uint32_t>(624 - index);
      const uint32_t step = count < room ? count : room;
      index += static_cast<int>(step);
      count -= step;
    }
  }
};

} // namespace ptb200
