// cloud.h -- host-side point clouds and the pre-processing the reference does with PCL / OpenCV before the hot path
// (main_realdata_auto.cpp:54-96,144-181; Utils.cpp:36-115,333-340).  Plain C++, no PCL.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "mat.h"

struct Cloud {               // pcl::PointCloud<pcl::PointSurfel>: xyz + normal + confidence
  std::vector<float> xyz, nrm, conf;
  size_t size() const { return xyz.size() / 3; }
  bool has_normals() const { return nrm.size() == xyz.size() && !xyz.empty(); }
  void push(const float *p, const float *n, float c) {
    xyz.insert(xyz.end(), p, p + 3);
    if (n) nrm.insert(nrm.end(), n, n + 3);
    conf.push_back(c);
  }
  void clear() { xyz.clear(); nrm.clear(); conf.clear(); }
};

// ---- io ----
bool loadPLYFile(const std::string &path, Cloud &out, std::string *err = nullptr);   // ascii / binary_little_endian vertices
bool savePLYFile(const std::string &path, const Cloud &c);                          // ascii, x y z nx ny nz confidence
// igl::readOBJ as SDFchecker::registerMesh uses it (SDFchecker.cpp:36-49): "v x y z" vertices, "f a b c" faces (a/b/c forms and
// negative indices accepted, polygons fanned into triangles); V = nv x 3, F = nf x 3 zero-based
bool loadOBJMesh(const std::string &path, std::vector<float> &V, std::vector<int32_t> &F, std::string *err = nullptr);
bool saveOBJMesh(const std::string &path, const std::vector<float> &V, const std::vector<int32_t> &F);   // "v" + "f" lines (best.obj)
bool saveOBJVertices(const std::string &path, const Cloud &c);                      // "v x y z" lines (best.obj stand-in when no mesh)
bool readPNG16(const std::string &path, std::vector<uint16_t> &pix, int &width, int &height, std::string *err = nullptr);
bool parsePoseTxt(const std::string &path, std::vector<float> &data);               // Utils.cpp:516-543: whitespace separated floats
bool savePoseTxt(const std::string &path, const Mat4f &T);

// ---- Utils:: ----
void readDepthImage(std::vector<float> &depth_m, int &w, int &h, const std::string &path);         // Utils.cpp:36-55 (0 outside 0.1..2 m)
void convert3dOrganized(const std::vector<float> &depth_m, int w, int h, const Mat3f &K, Cloud &out);  // Utils.cpp:78-115 (invalid = 0,0,0)
void downsamplePointCloud(const Cloud &in, Cloud &out, float leaf);                               // Utils.cpp:333-340 = pcl::VoxelGrid
void passThrough(const Cloud &in, Cloud &out, int axis, float lo, float hi);                      // pcl::PassThrough (keeps lo <= v <= hi)
void transformPointCloudWithNormals(const Cloud &in, Cloud &out, const Mat4f &T);                 // PCL 1.9 order
void estimateNormals(Cloud &c, float radius, const float *viewpoint);                             // PCA over a radius, flipped to the viewpoint
void removeAllNaNFromPointCloud(Cloud &c);
void getMinMax3D(const Cloud &c, float *mn, float *mx);

// The per-frame front end of main_realdata_auto.cpp:54-96,144-181 as one function (what hop_frame_to_scene does on the device):
// depth [m] -> organized cloud -> z pass-through -> VoxelGrid(1 mm) -> hand-base crop box -> normals(3 mm) -> VoxelGrid(3 mm)
// -> NaN removal -> normals towards the camera -> confidence 1.  cropped (may be null) receives the cloud before the normals.
void frameToObjectSegment(const std::vector<float> &depth_m, int w, int h, const Mat3f &K, const Mat4f &cam_in_handbase, Cloud &object_segment,
                          Cloud *cropped = nullptr);
