// ppf_table.h -- the model's PPF table (src/perception/src/app/computePPF.cpp:86-107): the 4-int key of every point pair
// of the 5 mm model.  The reference serialises a std::map<vector<int>, vector<pair<int,int>>> with Boost (a platform-
// specific binary, files not shipped); only key MEMBERSHIP is ever queried (matchBase.hpp:134,159,201), so the table here
// is the sorted set of distinct keys in a small portable file: "HOPPPF1\n", int32 n, n x 4 int32.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "cloud.h"

std::vector<int32_t> buildPPFTable(const Cloud &model_5mm);   // n x 4, distinct, sorted; keys by libhop's hop_compute_ppf
bool savePPFTable(const std::string &path, const std::vector<int32_t> &keys);
bool loadPPFTable(const std::string &path, std::vector<int32_t> &keys);
