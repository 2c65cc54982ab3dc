// util/util.hxx -- small host helpers the drivers use
// (reference: include/sparsify.me/util/util.hxx:20-61): get_random, ceil_div, mat_sz, read_shapes.
#pragma once
#include <fstream>
#include <iostream>
#include <random>
#include <string>
#include <tuple>
#include <vector>

#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <sparsify.me/util/launch.hxx>
#include <sparsify.me/util/timer.hxx>

namespace sparsifyme {
namespace util {

// U[begin, end).  One engine per host thread, seeded from the OS once (the reference builds a
// new random_device + mt19937 for every single value, util.hxx:21-26, which dominates the
// drivers' set-up time); the distribution is the same.
template <typename type_t = float>
type_t get_random(type_t begin = 0.0f, type_t end = 1.0f) {
  static thread_local std::mt19937 engine{std::random_device{}()};
  std::uniform_real_distribution<double> dist(static_cast<double>(begin), static_cast<double>(end));
  return static_cast<type_t>(dist(engine));
}

template <typename type_t>
constexpr type_t ceil_div(type_t x, type_t y) {
  return (x + y - 1) / y;
}

// one row of datasets/*.csv: m, n, k, b
typedef std::tuple<int, int, int, int> mat_sz;

// Reads a shape table: first line is the header, every other non-empty line holds four
// comma-separated integers.  CRLF endings (the per-model CSVs) are accepted.  Throws a
// `const char*` when the file cannot be opened, like the reference (util.hxx:41).
inline std::vector<mat_sz> read_shapes(std::string filename) {
  std::ifstream in(filename);
  if (!in.is_open()) throw "Unable to open shape CSV file.";
  std::vector<mat_sz> shapes;
  std::string line;
  std::getline(in, line);  // header
  while (std::getline(in, line)) {
    int v[4] = {0, 0, 0, 0};
    std::size_t pos = 0;
    int got = 0;
    while (got < 4 && pos < line.size()) {
      std::size_t comma = line.find(',', pos);
      if (comma == std::string::npos) comma = line.size();
      const std::string field = line.substr(pos, comma - pos);
      if (field.find_first_of("0123456789") == std::string::npos) break;
      v[got++] = std::stoi(field);
      pos = comma + 1;
    }
    if (got == 4) shapes.emplace_back(v[0], v[1], v[2], v[3]);
  }
  return shapes;
}

}  // namespace util
}  // namespace sparsifyme
