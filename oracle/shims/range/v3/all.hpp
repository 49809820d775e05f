// TEST INFRASTRUCTURE — stand-in for range-v3 0.9.1 (conanfile.txt of the reference), which is
// not installed here.  Provides exactly what src/fp/Render.cpp uses, with range-v3's evaluation
// order: views::ints(lo, hi), views::cartesian_product(a, b) (last range varies fastest, yields
// a tuple), `| views::transform(f)` (lazy; f runs once per dereference) and accumulate(range,
// init) (init = init + *it, left to right).
#pragma once

#include <tuple>
#include <type_traits>
#include <utility>

namespace ranges {

namespace shim {

struct IntsView {
  int lo, hi;
  struct iterator {
    int value;
    int operator*() const { return value; }
    iterator &operator++() {
      ++value;
      return *this;
    }
    bool operator!=(const iterator &o) const { return value != o.value; }
  };
  iterator begin() const { return {lo}; }
  iterator end() const { return {hi < lo ? lo : hi}; }
  bool empty() const { return hi <= lo; }
};

struct ProductView {
  IntsView outer, inner;
  struct iterator {
    int a, b, bLo, bHi;
    std::tuple<int, int> operator*() const { return {a, b}; }
    iterator &operator++() {
      if (++b >= bHi) {
        b = bLo;
        ++a;
      }
      return *this;
    }
    bool operator!=(const iterator &o) const { return a != o.a || b != o.b; }
  };
  iterator end() const { return {outer.hi < outer.lo ? outer.lo : outer.hi, inner.lo, inner.lo, inner.hi}; }
  iterator begin() const {
    if (outer.empty() || inner.empty())
      return end();
    return {outer.lo, inner.lo, inner.lo, inner.hi};
  }
};

template <typename Base, typename F>
struct TransformView {
  Base base;
  F f;
  struct iterator {
    decltype(std::declval<const Base &>().begin()) it;
    const F *f;
    decltype(auto) operator*() const { return (*f)(*it); }
    iterator &operator++() {
      ++it;
      return *this;
    }
    bool operator!=(const iterator &o) const { return it != o.it; }
  };
  iterator begin() const { return {base.begin(), &f}; }
  iterator end() const { return {base.end(), &f}; }
};

template <typename F>
struct TransformClosure {
  F f;
};

template <typename Range, typename F>
auto operator|(Range &&range, TransformClosure<F> closure) {
  return TransformView<std::decay_t<Range>, F>{std::forward<Range>(range), std::move(closure.f)};
}

} // namespace shim

namespace views {
inline shim::IntsView ints(int lo, int hi) { return {lo, hi}; }
inline shim::ProductView cartesian_product(shim::IntsView outer, shim::IntsView inner) { return {outer, inner}; }
template <typename F>
shim::TransformClosure<std::decay_t<F>> transform(F &&f) {
  return {std::forward<F>(f)};
}
} // namespace views

template <typename Range, typename T>
T accumulate(Range &&range, T init) {
  for (auto it = range.begin(), last = range.end(); it != last; ++it)
    init = std::move(init) + *it;
  return init;
}

} // namespace ranges
