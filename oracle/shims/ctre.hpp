// TEST INFRASTRUCTURE — build shim, not product code.
//
// Stand-in for the third-party CTRE header (compile-time regular expressions, v2.6.4 per
// the reference's conanfile.txt:2-8), which is not vendored in /root/reference and is not
// installed here.  The reference uses CTRE in exactly one place, the OBJ/MTL tokenizer
// (src/util/ObjLoaderImpl.h:20-21,32-37), with exactly one pattern:
//
//        \s*((#.*)|[^ \t\n\r#]+)
//
// This shim provides just enough surface for that translation unit to compile unmodified:
//   ctll::fixed_string{"..."}         (the pattern literal is accepted and ignored)
//   ctre::range<pattern>(string_view) (iterable of successive *search* matches)
//   match: operator bool, get<1>().to_view()
// The matcher below is hard-wired to the pattern above and is written from the regex
// semantics, not from CTRE's sources.
#pragma once

#include <cstddef>
#include <string_view>

namespace ctll {
template <std::size_t N>
struct fixed_string {
  char text[N]{};
  constexpr fixed_string(const char (&s)[N]) noexcept {
    for (std::size_t i = 0; i < N; ++i)
      text[i] = s[i];
  }
};
template <std::size_t N>
fixed_string(const char (&)[N]) -> fixed_string<N>;
} // namespace ctll

namespace ctre {

struct shim_capture {
  std::string_view view;
  [[nodiscard]] std::string_view to_view() const noexcept { return view; }
};

struct shim_match {
  bool ok{false};
  std::string_view group1;
  std::size_t endOffset{0};
  explicit operator bool() const noexcept { return ok; }
  template <int I>
  [[nodiscard]] shim_capture get() const noexcept {
    static_assert(I == 1, "only capture group 1 is used by the reference");
    return shim_capture{group1};
  }
};

namespace detail {
// Regex \s : space, \t, \n, \v, \f, \r.
constexpr bool isRegexSpace(char c) noexcept {
  return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r';
}
// The negated class [^ \t\n\r#].
constexpr bool isTokenChar(char c) noexcept {
  return !(c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '#');
}

// Search for the pattern starting at or after `from`.
inline shim_match searchFrom(std::string_view text, std::size_t from) noexcept {
  for (std::size_t start = from; start <= text.size(); ++start) {
    std::size_t pos = start;
    while (pos < text.size() && isRegexSpace(text[pos]))
      ++pos; // \s* is greedy; backtracking can never help because neither alternative
             // can begin with a whitespace character that \s would have consumed,
             // except \v and \f which are token characters: handle by trying shorter
             // whitespace runs below.
    for (;;) {
      if (pos < text.size()) {
        if (text[pos] == '#') { // (#.*) : '.' does not match \n
          std::size_t end = pos;
          while (end < text.size() && text[end] != '\n')
            ++end;
          return shim_match{true, text.substr(pos, end - pos), end};
        }
        if (isTokenChar(text[pos])) {
          std::size_t end = pos;
          while (end < text.size() && isTokenChar(text[end]))
            ++end;
          return shim_match{true, text.substr(pos, end - pos), end};
        }
      }
      if (pos == start)
        break;
      --pos; // backtrack the greedy \s*
    }
  }
  return shim_match{};
}
} // namespace detail

class shim_range {
  std::string_view text_;

public:
  explicit shim_range(std::string_view text) noexcept : text_(text) {}

  struct sentinel {};
  class iterator {
    std::string_view text_;
    shim_match current_;

  public:
    explicit iterator(std::string_view text) noexcept
        : text_(text), current_(detail::searchFrom(text, 0)) {}
    const shim_match &operator*() const noexcept { return current_; }
    iterator &operator++() noexcept {
      current_ = detail::searchFrom(text_, current_.endOffset);
      return *this;
    }
    bool operator!=(sentinel) const noexcept { return current_.ok; }
  };
  [[nodiscard]] iterator begin() const noexcept { return iterator(text_); }
  [[nodiscard]] sentinel end() const noexcept { return {}; }
};

template <auto &Pattern>
shim_range range(std::string_view text) noexcept {
  (void)Pattern;
  return shim_range(text);
}

} // namespace ctre
