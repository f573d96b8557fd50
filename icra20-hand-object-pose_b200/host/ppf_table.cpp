#include "ppf_table.h"

#include <algorithm>
#include <array>
#include <cstring>
#include <fstream>

#include "hop_c_api.h"

std::vector<int32_t> buildPPFTable(const Cloud &m) {
  std::vector<std::array<int32_t, 4>> keys;
  const size_t n = m.size();
  keys.reserve(n * (n - 1) / 2);
  for (size_t i = 0; i < n; ++i)
    for (size_t j = i + 1; j < n; ++j) {
      std::array<int32_t, 4> k;
      hop_compute_ppf(&m.xyz[3 * i], &m.nrm[3 * i], &m.xyz[3 * j], &m.nrm[3 * j], k.data());
      keys.push_back(k);
    }
  std::sort(keys.begin(), keys.end());
  keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
  std::vector<int32_t> flat;
  flat.reserve(4 * keys.size());
  for (const auto &k : keys) flat.insert(flat.end(), k.begin(), k.end());
  return flat;
}

bool savePPFTable(const std::string &path, const std::vector<int32_t> &keys) {
  std::ofstream f(path, std::ios::binary);
  if (!f) return false;
  f.write("HOPPPF1\n", 8);
  const int32_t n = (int32_t)(keys.size() / 4);
  f.write((const char *)&n, 4);
  f.write((const char *)keys.data(), sizeof(int32_t) * keys.size());
  return (bool)f;
}

bool loadPPFTable(const std::string &path, std::vector<int32_t> &keys) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  char magic[8];
  int32_t n = 0;
  if (!f.read(magic, 8) || std::memcmp(magic, "HOPPPF1\n", 8) != 0 || !f.read((char *)&n, 4) || n < 0) return false;
  keys.resize(4 * (size_t)n);
  return (bool)f.read((char *)keys.data(), sizeof(int32_t) * keys.size());
}
