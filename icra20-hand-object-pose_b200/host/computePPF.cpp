// computePPF -- the reference's offline PPF-table builder (src/perception/src/app/computePPF.cpp:40-108):
//     computePPF <config.yaml> <out_dir>
// writes <out_dir>/model.ply (the 1 mm model) and <out_dir>/ppf (the key table of the 5 mm model, libhop's portable format).
// The model's own normals are used (the reference re-estimates them with MLS; models here ship with normals).
#include <cstdio>
#include <iostream>

#include "ConfigParser.h"
#include "cloud.h"
#include "ppf_table.h"

int main(int argc, char **argv) {
  if (argc < 3) { std::cout << "Arguments:\narg1: config_dir, arg2: out_dir" << std::endl; return 1; }
  ConfigParser cfg(argv[1]);
  const std::string out_dir = argv[2];
  Cloud cloud, m5;
  std::string err;
  if (!loadPLYFile(cfg.object_model_path, cloud, &err)) { printf("%s\n", err.c_str()); return 1; }
  downsamplePointCloud(cloud, cloud, 0.001f);
  savePLYFile(out_dir + "/model.ply", cloud);
  downsamplePointCloud(cloud, m5, 0.005f);
  const std::vector<int32_t> keys = buildPPFTable(m5);
  if (!savePPFTable(out_dir + "/ppf", keys)) { printf("cannot write %s/ppf\n", out_dir.c_str()); return 1; }
  printf("%d points, %d distinct PPF keys\n", (int)m5.size(), (int)(keys.size() / 4));
  return 0;
}
