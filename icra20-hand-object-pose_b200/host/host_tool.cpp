// host_tool -- small command-line probes of the host-side pieces, used by tests/test_host_cpp.py (no GPU needed):
//   host_tool yaml <file> <key[.key...]>        prints a scalar or a space-separated sequence
//   host_tool config <file>                     prints the matrices ConfigParser derives
//   host_tool png <file>                        width height sum min max of a 16-bit PNG
//   host_tool voxel <in.ply> <leaf> <out.ply>   pcl::VoxelGrid restatement
//   host_tool depthcloud <png> fx fy cx cy <out.ply>
//   host_tool normals <in.ply> <radius> <out.ply>
//   host_tool frame <png> fx fy cx cy <cam_in_handbase.txt> <out.bin>   the whole front end; out.bin = int32 n, then n x 7 float32
//                                                                       (xyz, normal, confidence), then the 16 floats of
//                                                                       cam_in_handbase.inverse() (column-major)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>

#include "ConfigParser.h"
#include "cloud.h"
#include "urdf.h"

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  const std::string cmd = argv[1];
  try {
    if (cmd == "yaml" && argc >= 4) {
      miniyaml::Node n = miniyaml::LoadFile(argv[2]);
      const miniyaml::Node *cur = &n;
      std::stringstream ss(argv[3]);
      std::string part;
      while (std::getline(ss, part, '.')) cur = &(*cur)[part];
      if (cur->kind == miniyaml::Node::Scalar) std::cout << cur->scalar << "\n";
      else if (cur->kind == miniyaml::Node::Sequence) { for (size_t i = 0; i < cur->size(); ++i) std::cout << (i ? " " : "") << (*cur)[i].scalar; std::cout << "\n"; }
      else if (cur->kind == miniyaml::Node::Map) { for (auto &k : cur->keys()) std::cout << k << " "; std::cout << "\n"; }
      else { std::cout << "<undefined>\n"; return 3; }
      return 0;
    }
    if (cmd == "config") {
      ConfigParser cfg(argv[2]);
      std::cout.precision(9);
      std::cout << "cam1_in_leftarm\n" << cfg.cam1_in_leftarm << "\nhandbase_in_palm\n" << cfg.handbase_in_palm << "\npalm_in_baselink\n" << cfg.palm_in_baselink
                << "\nleftarm_in_base\n" << cfg.leftarm_in_base << "\n";
      const Mat4f hic = cfg.cam1_in_leftarm.inverse() * (cfg.leftarm_in_base.inverse() * cfg.palm_in_baselink * cfg.handbase_in_palm);
      std::cout << "handbase_in_cam\n" << hic << "\n";
      std::cout << "K " << cfg.cam_intrinsic(0, 0) << " " << cfg.cam_intrinsic(1, 1) << " " << cfg.cam_intrinsic(0, 2) << " " << cfg.cam_intrinsic(1, 2) << "\n";
      return 0;
    }
    if (cmd == "png") {
      std::vector<uint16_t> pix; int w, h; std::string err;
      if (!readPNG16(argv[2], pix, w, h, &err)) { std::cout << err << "\n"; return 3; }
      unsigned long long sum = 0; uint16_t mn = 65535, mx = 0;
      for (uint16_t v : pix) { sum += v; mn = std::min(mn, v); mx = std::max(mx, v); }
      std::cout << w << " " << h << " " << sum << " " << mn << " " << mx << "\n";
      return 0;
    }
    if (cmd == "urdf") {   // parseUrdfLinks: name parent | scale | tf_init rows | tf_in_parent rows, one link per line
      std::vector<UrdfLink> links; std::string err;
      if (!parseUrdfLinks(argv[2], links, &err)) { std::cout << err << "\n"; return 3; }
      std::cout.precision(9);
      for (const UrdfLink &L : links) {
        std::cout << L.name << " " << (L.parent.empty() ? "-" : L.parent) << " " << L.scale[0] << " " << L.scale[1] << " " << L.scale[2];
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) std::cout << " " << L.tf_init(r, c);
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) std::cout << " " << L.tf_in_parent(r, c);
        std::cout << "\n";
      }
      return 0;
    }
    if (cmd == "obj") {   // loadOBJMesh: vertex / face counts, index range, signed volume (orientation) and the first face
      std::vector<float> V; std::vector<int32_t> F; std::string err;
      if (!loadOBJMesh(argv[2], V, F, &err)) { std::cout << err << "\n"; return 3; }
      double vol = 0;
      for (size_t f = 0; f + 2 < F.size(); f += 3) {
        const float *a = &V[3 * F[f]], *b = &V[3 * F[f + 1]], *c = &V[3 * F[f + 2]];
        vol += (double)a[0] * ((double)b[1] * c[2] - (double)b[2] * c[1]) - (double)a[1] * ((double)b[0] * c[2] - (double)b[2] * c[0]) +
               (double)a[2] * ((double)b[0] * c[1] - (double)b[1] * c[0]);
      }
      std::cout.precision(9);
      std::cout << V.size() / 3 << " " << F.size() / 3 << " " << vol / 6 << " " << F[0] << " " << F[1] << " " << F[2] << "\n";
      return 0;
    }
    if (cmd == "voxel" && argc >= 5) {
      Cloud c, o; std::string err;
      if (!loadPLYFile(argv[2], c, &err)) { std::cout << err << "\n"; return 3; }
      downsamplePointCloud(c, o, (float)atof(argv[3]));
      return savePLYFile(argv[4], o) ? 0 : 3;
    }
    if (cmd == "depthcloud" && argc >= 8) {
      std::vector<float> d; int w, h;
      readDepthImage(d, w, h, argv[2]);
      Mat3f K; for (int i = 0; i < 9; ++i) K.m[i] = 0; K(0, 0) = atof(argv[3]); K(1, 1) = atof(argv[4]); K(0, 2) = atof(argv[5]); K(1, 2) = atof(argv[6]); K(2, 2) = 1;
      Cloud c; convert3dOrganized(d, w, h, K, c);
      passThrough(c, c, 2, 0.1f, 2.0f);
      return savePLYFile(argv[7], c) ? 0 : 3;
    }
    if (cmd == "normals" && argc >= 5) {
      Cloud c; std::string err;
      if (!loadPLYFile(argv[2], c, &err)) { std::cout << err << "\n"; return 3; }
      const float o[3] = {0, 0, 0};
      estimateNormals(c, (float)atof(argv[3]), o);
      return savePLYFile(argv[4], c) ? 0 : 3;
    }
    if (cmd == "frame" && argc >= 9) {
      std::vector<float> d; int w, h;
      readDepthImage(d, w, h, argv[2]);
      if (d.empty()) return 3;
      Mat3f K; for (int i = 0; i < 9; ++i) K.m[i] = 0; K(0, 0) = atof(argv[3]); K(1, 1) = atof(argv[4]); K(0, 2) = atof(argv[5]); K(1, 2) = atof(argv[6]); K(2, 2) = 1;
      std::vector<float> t;
      if (!parsePoseTxt(argv[7], t) || t.size() < 16) return 3;
      Mat4f T; for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T(r, c) = t[4 * r + c];
      Cloud seg;
      const auto t0 = std::chrono::steady_clock::now();
      frameToObjectSegment(d, w, h, K, T, seg);
      std::cout << "frame_ms " << std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() << "\n";
      FILE *f = fopen(argv[8], "wb");
      if (!f) return 3;
      const int32_t n = (int32_t)seg.size();
      fwrite(&n, 4, 1, f);
      for (int32_t i = 0; i < n; ++i) { fwrite(&seg.xyz[3 * i], 4, 3, f); fwrite(&seg.nrm[3 * i], 4, 3, f); fwrite(&seg.conf[i], 4, 1, f); }
      const Mat4f Ti = T.inverse();
      fwrite(Ti.data(), 4, 16, f);
      fclose(f);
      return 0;
    }
  } catch (const std::exception &e) { std::cout << "error: " << e.what() << "\n"; return 4; }
  return 2;
}
