// Wavefront OBJ/MTL loading for the host boundary, with the semantics of the reference's
// loader (src/util/ObjLoaderImpl.h:20-103, src/util/ObjLoader.cpp:15-108) so that the SoA
// arrays handed to the GPU are the ones dod::Scene would hold:
//   * tokens are maximal runs of characters outside {space, tab, LF, CR, '#'}; '#' starts a
//     comment to end of line; an unknown directive is an error carrying the line number;
//   * `v` appends a vertex; `f` fan-triangulates (i0, ik, ik+1) with 1-based or negative
//     (relative) indices, and an `a/b/c` token parses as `a`; `g`, `o`, `s` are ignored;
//   * `usemtl` copies the named material into every following triangle, `mtllib` loads
//     materials through the opener;
//   * MTL: Ke -> emission, Kd -> diffuse, Ni -> index of refraction, Ns -> cone angle
//     pi * clamp(1 - Ns/100, 0, 1), `illum 3` -> reflectivity = |Ka| when the block is
//     flushed; `illum` and `Ka` persist across newmtl blocks; Ks and d are ignored.
// Written from those rules with a hand-rolled scanner (the reference uses the CTRE regex
// library, which this image does not have).
#pragma once

#include "MaterialSpec.h"
#include "Vec3.h"

#include <algorithm>
#include <fstream>
#include <istream>
#include <memory>
#include <stdexcept>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

namespace ptb200 {

struct ObjLoaderOpener { // same role as src/util/ObjLoader.h:9-12
  virtual ~ObjLoaderOpener() = default;
  virtual std::unique_ptr<std::istream> open(const std::string &filename) = 0;
};

struct DirRelativeOpener : ObjLoaderOpener { // src/main/main.cpp:27-38
  std::string dir;
  explicit DirRelativeOpener(std::string d) : dir(std::move(d)) {}
  std::unique_ptr<std::istream> open(const std::string &filename) override {
    const std::string full = dir + "/" + filename;
    auto stream = std::make_unique<std::ifstream>(full);
    if (!*stream)
      throw std::runtime_error("Unable to open " + full);
    return stream;
  }
};

namespace objimpl {

inline bool isSeparator(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }

inline std::vector<std::string_view> splitFields(std::string_view line) {
  std::vector<std::string_view> fields;
  size_t pos = 0;
  while (pos < line.size()) {
    if (isSeparator(line[pos])) {
      ++pos;
      continue;
    }
    if (line[pos] == '#')
      break;
    size_t end = pos;
    while (end < line.size() && !isSeparator(line[end]) && line[end] != '#')
      ++end;
    fields.push_back(line.substr(pos, end - pos));
    pos = end;
  }
  return fields;
}

inline double asDouble(std::string_view sv) { return std::stod(std::string(sv)); }
inline int asInt(std::string_view sv) { return std::stoi(std::string(sv)); }
inline size_t asIndex(std::string_view sv, size_t count) {
  const long value = std::stol(std::string(sv));
  return value < 0 ? static_cast<size_t>(value + static_cast<long>(count))
                   : static_cast<size_t>(value - 1);
}

template <typename Handler>
void forEachDirective(std::istream &in, Handler &&handler) {
  std::string line;
  int lineNumber = 0;
  while (std::getline(in, line)) {
    ++lineNumber;
    auto fields = splitFields(line);
    if (fields.empty())
      continue;
    const std::string_view command = fields.front();
    fields.erase(fields.begin());
    if (!handler(command, fields))
      throw std::runtime_error("Unknown directive '" + std::string(command) + "' on line " +
                               std::to_string(lineNumber));
  }
}

inline Vec3 threeDoubles(const std::vector<std::string_view> &p, const char *what) {
  if (p.size() != 3)
    throw std::runtime_error(std::string("Wrong number of params for ") + what);
  return Vec3(asDouble(p[0]), asDouble(p[1]), asDouble(p[2]));
}

} // namespace objimpl

inline std::unordered_map<std::string, MaterialSpec> loadMaterials(std::istream &in) {
  using namespace objimpl;
  if (!in)
    throw std::runtime_error("Bad input stream");
  in.exceptions(std::ios_base::badbit);
  std::unordered_map<std::string, MaterialSpec> result;
  MaterialSpec *current = nullptr;
  int illum = 2;
  Vec3 ambient;
  auto flush = [&] {
    if (current && illum == 3)
      current->reflectivity = ambient.length();
    current = nullptr;
  };
  auto need = [&](const char *what) -> MaterialSpec & {
    if (!current)
      throw std::runtime_error(std::string("Unexpected ") + what);
    return *current;
  };
  forEachDirective(in, [&](std::string_view cmd, const std::vector<std::string_view> &p) {
    if (cmd == "newmtl") {
      flush();
      if (p.size() != 1)
        throw std::runtime_error("Wrong number of params for newmtl");
      current = &result.emplace(std::string(p[0]), MaterialSpec{}).first->second;
    } else if (cmd == "Ke") {
      need("Ke").emission = threeDoubles(p, "Ke");
    } else if (cmd == "Kd") {
      need("Kd").diffuse = threeDoubles(p, "Kd");
    } else if (cmd == "Ka") {
      need("Ka");
      ambient = threeDoubles(p, "Ka");
    } else if (cmd == "Ni") {
      auto &mat = need("Ni");
      if (p.size() != 1)
        throw std::runtime_error("Wrong number of params for Ni");
      mat.indexOfRefraction = asDouble(p[0]);
    } else if (cmd == "Ns") {
      auto &mat = need("Ns");
      if (p.size() != 1)
        throw std::runtime_error("Wrong number of params for Ns");
      mat.reflectionConeAngleRadians = M_PI * std::clamp(1 - asDouble(p[0]) / 100, 0.0, 1.0);
    } else if (cmd == "illum") {
      need("illum");
      if (p.size() != 1)
        throw std::runtime_error("Wrong number of params for illum");
      illum = asInt(p[0]);
    } else if (cmd == "Ks" || cmd == "d") {
      // ignored, as in the reference
    } else {
      return false;
    }
    return true;
  });
  flush();
  return result;
}

template <typename SceneBuilder>
void loadObjFile(std::istream &in, ObjLoaderOpener &opener, SceneBuilder &sb) {
  using namespace objimpl;
  if (!in)
    throw std::runtime_error("Bad input stream");
  in.exceptions(std::ios_base::badbit);
  std::vector<Vec3> vertices;
  std::unordered_map<std::string, MaterialSpec> materials;
  MaterialSpec currentMaterial;
  forEachDirective(in, [&](std::string_view cmd, const std::vector<std::string_view> &p) {
    if (cmd == "v") {
      vertices.push_back(threeDoubles(p, "v"));
    } else if (cmd == "f") {
      std::vector<size_t> indices;
      indices.reserve(p.size());
      for (auto field : p)
        indices.push_back(asIndex(field, vertices.size()));
      for (size_t k = 1; k + 1 < p.size(); ++k)
        sb.addTriangle(vertices.at(indices[0]), vertices.at(indices[k]),
                       vertices.at(indices[k + 1]), currentMaterial);
    } else if (cmd == "g" || cmd == "o" || cmd == "s") {
      // groups, object names, smoothing: ignored
    } else if (cmd == "usemtl") {
      const std::string name(p.at(0));
      const auto found = materials.find(name);
      if (found == materials.end())
        throw std::runtime_error("Can't find material " + name);
      currentMaterial = found->second;
    } else if (cmd == "mtllib") {
      const auto file = opener.open(std::string(p.at(0)));
      materials = loadMaterials(*file);
    } else {
      return false;
    }
    return true;
  });
}

} // namespace ptb200
