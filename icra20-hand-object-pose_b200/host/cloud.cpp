#include "cloud.h"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>
#include <unordered_map>

// --------------------------------------------------------------------------------------------------------------- PLY
namespace {
struct PlyProp { std::string type, name; int bytes; };
int ply_type_bytes(const std::string &t) {
  if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
  if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
  if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
  if (t == "double" || t == "float64") return 8;
  return 0;
}
double ply_read_bin(const unsigned char *p, const std::string &t) {
  if (t == "float" || t == "float32") { float v; std::memcpy(&v, p, 4); return v; }
  if (t == "double" || t == "float64") { double v; std::memcpy(&v, p, 8); return v; }
  if (t == "uchar" || t == "uint8") return *p;
  if (t == "char" || t == "int8") return *(const signed char *)p;
  if (t == "short" || t == "int16") { int16_t v; std::memcpy(&v, p, 2); return v; }
  if (t == "ushort" || t == "uint16") { uint16_t v; std::memcpy(&v, p, 2); return v; }
  if (t == "int" || t == "int32") { int32_t v; std::memcpy(&v, p, 4); return v; }
  if (t == "uint" || t == "uint32") { uint32_t v; std::memcpy(&v, p, 4); return v; }
  return 0;
}
}  // namespace

bool loadPLYFile(const std::string &path, Cloud &out, std::string *err) {
  out.clear();
  std::ifstream f(path, std::ios::binary);
  auto fail = [&](const std::string &m) { if (err) *err = m; return false; };
  if (!f) return fail("cannot open " + path);
  std::string line, fmt;
  std::getline(f, line);
  if (line.substr(0, 3) != "ply") return fail("not a PLY file");
  size_t n_vertex = 0;
  std::vector<PlyProp> props;
  bool in_vertex = false;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    std::istringstream is(line);
    std::string tok;
    is >> tok;
    if (tok == "format") is >> fmt;
    else if (tok == "element") { std::string name; size_t n; is >> name >> n; in_vertex = name == "vertex"; if (in_vertex) n_vertex = n; }
    else if (tok == "property" && in_vertex) {
      PlyProp p; is >> p.type;
      if (p.type == "list") return fail("list property inside the vertex element");
      is >> p.name; p.bytes = ply_type_bytes(p.type);
      if (!p.bytes) return fail("unknown PLY type " + p.type);
      props.push_back(p);
    } else if (tok == "end_header") break;
  }
  int ix = -1, iy = -1, iz = -1, inx = -1, iny = -1, inz = -1, ic = -1;
  for (size_t k = 0; k < props.size(); ++k) {
    const std::string &n = props[k].name;
    if (n == "x") ix = k; else if (n == "y") iy = k; else if (n == "z") iz = k;
    else if (n == "nx" || n == "normal_x") inx = k; else if (n == "ny" || n == "normal_y") iny = k; else if (n == "nz" || n == "normal_z") inz = k;
    else if (n == "confidence") ic = k;
  }
  if (ix < 0 || iy < 0 || iz < 0) return fail("PLY has no x/y/z");
  const bool has_n = inx >= 0 && iny >= 0 && inz >= 0;
  std::vector<double> v(props.size());
  size_t stride = 0;
  for (const PlyProp &p : props) stride += p.bytes;
  std::vector<unsigned char> rec(stride);
  for (size_t i = 0; i < n_vertex; ++i) {
    if (fmt == "ascii") {
      for (size_t k = 0; k < props.size(); ++k) if (!(f >> v[k])) return fail("truncated ascii PLY");
    } else if (fmt == "binary_little_endian") {
      if (!f.read((char *)rec.data(), stride)) return fail("truncated binary PLY");
      size_t off = 0;
      for (size_t k = 0; k < props.size(); ++k) { v[k] = ply_read_bin(rec.data() + off, props[k].type); off += props[k].bytes; }
    } else return fail("unsupported PLY format " + fmt);
    const float p[3] = {(float)v[ix], (float)v[iy], (float)v[iz]};
    const float n[3] = {has_n ? (float)v[inx] : 0.f, has_n ? (float)v[iny] : 0.f, has_n ? (float)v[inz] : 0.f};
    out.push(p, has_n ? n : nullptr, ic >= 0 ? (float)v[ic] : 1.f);
  }
  return true;
}

bool savePLYFile(const std::string &path, const Cloud &c) {
  std::ofstream f(path);
  if (!f) return false;
  const bool n = c.has_normals();
  f << "ply\nformat ascii 1.0\nelement vertex " << c.size() << "\nproperty float x\nproperty float y\nproperty float z\n";
  if (n) f << "property float nx\nproperty float ny\nproperty float nz\n";
  f << "property float confidence\nend_header\n";
  f.precision(9);
  for (size_t i = 0; i < c.size(); ++i) {
    f << c.xyz[3 * i] << " " << c.xyz[3 * i + 1] << " " << c.xyz[3 * i + 2];
    if (n) f << " " << c.nrm[3 * i] << " " << c.nrm[3 * i + 1] << " " << c.nrm[3 * i + 2];
    f << " " << (i < c.conf.size() ? c.conf[i] : 1.f) << "\n";
  }
  return true;
}

bool loadOBJMesh(const std::string &path, std::vector<float> &V, std::vector<int32_t> &F, std::string *err) {
  std::ifstream f(path);
  if (!f) { if (err) *err = "cannot open " + path; return false; }
  V.clear(); F.clear();
  std::string line;
  while (std::getline(f, line)) {
    std::istringstream ss(line);
    std::string tag;
    if (!(ss >> tag)) continue;
    if (tag == "v") {
      float x, y, z;
      if (!(ss >> x >> y >> z)) { if (err) *err = "bad vertex line in " + path; return false; }
      V.push_back(x); V.push_back(y); V.push_back(z);
    } else if (tag == "f") {
      std::vector<int32_t> idx;
      std::string tok;
      while (ss >> tok) {
        const long v = std::strtol(tok.c_str(), nullptr, 10);   // "a", "a/b", "a/b/c", "a//c": the vertex index comes first
        if (v == 0) { if (err) *err = "bad face line in " + path; return false; }
        idx.push_back(v > 0 ? (int32_t)(v - 1) : (int32_t)(V.size() / 3 + v));
      }
      for (size_t k = 2; k < idx.size(); ++k) { F.push_back(idx[0]); F.push_back(idx[k - 1]); F.push_back(idx[k]); }
    }
  }
  const int32_t nv = (int32_t)(V.size() / 3);
  for (int32_t i : F) if (i < 0 || i >= nv) { if (err) *err = "face index out of range in " + path; return false; }
  if (V.empty() || F.empty()) { if (err) *err = "no mesh in " + path; return false; }
  return true;
}

bool saveOBJMesh(const std::string &path, const std::vector<float> &V, const std::vector<int32_t> &F) {
  std::ofstream f(path);
  if (!f) return false;
  f.precision(9);
  for (size_t i = 0; i + 2 < V.size(); i += 3) f << "v " << V[i] << " " << V[i + 1] << " " << V[i + 2] << "\n";
  for (size_t i = 0; i + 2 < F.size(); i += 3) f << "f " << F[i] + 1 << " " << F[i + 1] + 1 << " " << F[i + 2] + 1 << "\n";
  return true;
}

bool saveOBJVertices(const std::string &path, const Cloud &c) {
  std::ofstream f(path);
  if (!f) return false;
  f.precision(9);
  for (size_t i = 0; i < c.size(); ++i) f << "v " << c.xyz[3 * i] << " " << c.xyz[3 * i + 1] << " " << c.xyz[3 * i + 2] << "\n";
  return true;
}

// --------------------------------------------------------------------------------------------------------------- PNG
// 16-bit (or 8-bit) grayscale, non-interlaced PNG through zlib: what cv::imread(path, CV_16UC1) reads for the depth frames.
bool readPNG16(const std::string &path, std::vector<uint16_t> &pix, int &width, int &height, std::string *err) {
  auto fail = [&](const std::string &m) { if (err) *err = m; return false; };
  std::ifstream f(path, std::ios::binary);
  if (!f) return fail("cannot open " + path);
  std::vector<unsigned char> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (buf.size() < 8 || std::memcmp(buf.data(), sig, 8) != 0) return fail("not a PNG file");
  auto be32 = [&](size_t o) { return (uint32_t)buf[o] << 24 | (uint32_t)buf[o + 1] << 16 | (uint32_t)buf[o + 2] << 8 | buf[o + 3]; };
  size_t o = 8;
  int depth = 0, ctype = -1, interlace = 0;
  std::vector<unsigned char> idat;
  while (o + 12 <= buf.size()) {
    const uint32_t len = be32(o);
    const std::string type((const char *)&buf[o + 4], 4);
    if (o + 12 + len > buf.size()) return fail("truncated PNG chunk");
    const unsigned char *d = &buf[o + 8];
    if (type == "IHDR") {
      if (len != 13) return fail("bad PNG IHDR chunk");
      const uint32_t w32 = be32(o + 8), h32 = be32(o + 12);
      if (w32 == 0 || h32 == 0 || w32 > 16384 || h32 > 16384) return fail("PNG dimensions out of range");
      width = (int)w32; height = (int)h32; depth = d[8]; ctype = d[9]; interlace = d[12];
    }
    else if (type == "IDAT") idat.insert(idat.end(), d, d + len);
    else if (type == "IEND") break;
    o += 12 + len;
  }
  if (width <= 0 || height <= 0) return fail("PNG without an IHDR chunk");
  if (ctype != 0 || (depth != 16 && depth != 8) || interlace != 0) return fail("only non-interlaced 8/16-bit grayscale PNGs are supported");
  const int bpp = depth / 8;
  const size_t stride = (size_t)width * bpp;
  std::vector<unsigned char> raw((stride + 1) * height);
  uLongf raw_len = raw.size();
  if (uncompress(raw.data(), &raw_len, idat.data(), idat.size()) != Z_OK || raw_len != raw.size()) return fail("PNG inflate failed");
  std::vector<unsigned char> img(stride * height);
  for (int y = 0; y < height; ++y) {
    const unsigned char *in = &raw[(stride + 1) * y];
    const int ft = in[0];
    unsigned char *cur = &img[stride * y];
    const unsigned char *up = y ? &img[stride * (y - 1)] : nullptr;
    for (size_t x = 0; x < stride; ++x) {
      const int a = x >= (size_t)bpp ? cur[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= (size_t)bpp) ? up[x - bpp] : 0;
      int pred = 0;
      switch (ft) {
        case 0: pred = 0; break;
        case 1: pred = a; break;
        case 2: pred = b; break;
        case 3: pred = (a + b) >> 1; break;
        case 4: { const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
        default: return fail("bad PNG filter type");
      }
      cur[x] = (unsigned char)(in[1 + x] + pred);
    }
  }
  pix.resize((size_t)width * height);
  for (size_t i = 0; i < pix.size(); ++i) pix[i] = bpp == 2 ? (uint16_t)(img[2 * i] << 8 | img[2 * i + 1]) : img[i];
  return true;
}

bool parsePoseTxt(const std::string &path, std::vector<float> &data) {
  data.clear();
  std::ifstream f(path);
  if (!f) { printf("opening failed: \n%s\n", path.c_str()); return false; }
  float v;
  while (f >> v) data.push_back(v);
  return true;
}

bool savePoseTxt(const std::string &path, const Mat4f &T) {
  std::ofstream f(path);
  if (!f) return false;
  f << T << std::endl;
  return true;
}

// --------------------------------------------------------------------------------------------------------------- Utils
void readDepthImage(std::vector<float> &depth_m, int &w, int &h, const std::string &path) {
  std::vector<uint16_t> raw;
  std::string err;
  if (!readPNG16(path, raw, w, h, &err)) { printf("readDepthImage: %s\n", err.c_str()); depth_m.clear(); w = h = 0; return; }
  const double SR300_DEPTH_UNIT = 0.001;  // Utils.h:107: a double literal -- the product is taken in double and rounded to float once
  depth_m.resize(raw.size());
  for (size_t i = 0; i < raw.size(); ++i) {
    const float d = (float)((double)(float)raw[i] * SR300_DEPTH_UNIT);
    depth_m[i] = (d > 2.0 || d < 0.1) ? 0.f : d;
  }
}

void convert3dOrganized(const std::vector<float> &depth_m, int w, int h, const Mat3f &K, Cloud &out) {
  out.clear();
  out.xyz.resize((size_t)3 * w * h, 0.f);
  out.conf.assign((size_t)w * h, 1.f);
  for (int u = 0; u < h; ++u)
    for (int v = 0; v < w; ++v) {
      const float depth = depth_m[(size_t)u * w + v];
      float *p = &out.xyz[3 * ((size_t)u * w + v)];
      if (depth > 0.1 && depth < 2.0) {
        p[0] = (float)((v - K(0, 2)) * depth / K(0, 0));
        p[1] = (float)((u - K(1, 2)) * depth / K(1, 1));
        p[2] = depth;
      }
    }
}

void getMinMax3D(const Cloud &c, float *mn, float *mx) {
  for (int k = 0; k < 3; ++k) { mn[k] = std::numeric_limits<float>::max(); mx[k] = -std::numeric_limits<float>::max(); }
  for (size_t i = 0; i < c.size(); ++i)
    for (int k = 0; k < 3; ++k) {
      const float v = c.xyz[3 * i + k];
      if (!std::isfinite(v)) continue;
      mn[k] = std::min(mn[k], v); mx[k] = std::max(mx[k], v);
    }
}

// pcl::VoxelGrid (downsample_all_data): leaf index from floor(coord / leaf), points sorted by linear leaf index (x fastest),
// one output point per leaf = centroid of positions and confidence, normalised mean normal.
void downsamplePointCloud(const Cloud &in, Cloud &out, float leaf) {
  Cloud res;
  const size_t n = in.size();
  if (n == 0) { out = res; return; }
  float mn[3], mx[3];
  getMinMax3D(in, mn, mx);
  const float inv = 1.0f / leaf;
  long long minb[3], divb[3];
  for (int k = 0; k < 3; ++k) { minb[k] = (long long)std::floor(mn[k] * inv); divb[k] = (long long)std::floor(mx[k] * inv) - minb[k] + 1; }
  std::vector<std::pair<long long, uint32_t>> idx;
  idx.reserve(n);
  for (size_t i = 0; i < n; ++i) {
    const float *p = &in.xyz[3 * i];
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    const long long ijk0 = (long long)std::floor(p[0] * inv) - minb[0], ijk1 = (long long)std::floor(p[1] * inv) - minb[1],
                    ijk2 = (long long)std::floor(p[2] * inv) - minb[2];
    idx.emplace_back(ijk0 + ijk1 * divb[0] + ijk2 * divb[0] * divb[1], (uint32_t)i);
  }
  std::sort(idx.begin(), idx.end());
  const bool hn = in.has_normals();
  for (size_t a = 0; a < idx.size();) {
    size_t b = a;
    float s[3] = {0, 0, 0}, sn[3] = {0, 0, 0}, sc = 0.f;
    while (b < idx.size() && idx[b].first == idx[a].first) {
      const uint32_t i = idx[b].second;
      for (int k = 0; k < 3; ++k) { s[k] += in.xyz[3 * i + k]; if (hn) sn[k] += in.nrm[3 * i + k]; }
      sc += i < in.conf.size() ? in.conf[i] : 1.f;
      ++b;
    }
    const float cnt = (float)(b - a);
    const float p[3] = {s[0] / cnt, s[1] / cnt, s[2] / cnt};
    float nn[3] = {sn[0], sn[1], sn[2]};
    const float len = std::sqrt(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
    if (len > 0) { nn[0] /= len; nn[1] /= len; nn[2] /= len; }
    res.push(p, hn ? nn : nullptr, sc / cnt);
    a = b;
  }
  out = res;
}

void passThrough(const Cloud &in, Cloud &out, int axis, float lo, float hi) {
  Cloud res;
  const bool hn = in.has_normals();
  for (size_t i = 0; i < in.size(); ++i) {
    const float *p = &in.xyz[3 * i];
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    if (p[axis] < lo || p[axis] > hi) continue;
    res.push(p, hn ? &in.nrm[3 * i] : nullptr, i < in.conf.size() ? in.conf[i] : 1.f);
  }
  out = res;
}

void transformPointCloudWithNormals(const Cloud &in, Cloud &out, const Mat4f &T) {
  Cloud res = in;
  const bool hn = in.has_normals();
  for (size_t i = 0; i < in.size(); ++i) {
    const float x = in.xyz[3 * i], y = in.xyz[3 * i + 1], z = in.xyz[3 * i + 2];
    for (int r = 0; r < 3; ++r) res.xyz[3 * i + r] = T(r, 0) * x + T(r, 1) * y + T(r, 2) * z + T(r, 3);
    if (hn) {
      const float a = in.nrm[3 * i], b = in.nrm[3 * i + 1], c = in.nrm[3 * i + 2];
      for (int r = 0; r < 3; ++r) res.nrm[3 * i + r] = T(r, 0) * a + T(r, 1) * b + T(r, 2) * c;
    }
  }
  out = res;
}

void removeAllNaNFromPointCloud(Cloud &c) {
  Cloud res;
  const bool hn = c.has_normals();
  for (size_t i = 0; i < c.size(); ++i) {
    bool ok = true;
    for (int k = 0; k < 3; ++k) ok = ok && std::isfinite(c.xyz[3 * i + k]) && (!hn || std::isfinite(c.nrm[3 * i + k]));
    if (ok) res.push(&c.xyz[3 * i], hn ? &c.nrm[3 * i] : nullptr, i < c.conf.size() ? c.conf[i] : 1.f);
  }
  c = res;
}

// Surface normals by PCA over the neighbours within `radius` (hash grid of cell = radius), flipped towards the viewpoint
// (pcl::flipNormalTowardsViewpoint).  Stand-in for the reference's integral-image / MLS normals, which are PCL library
// algorithms outside the hot path; points with fewer than 3 neighbours get a NaN normal (dropped by the caller).
void estimateNormals(Cloud &c, float radius, const float *vp) {
  const size_t n = c.size();
  c.nrm.assign(3 * n, std::numeric_limits<float>::quiet_NaN());
  if (n == 0) return;
  const float inv = 1.f / radius;
  auto key = [&](long long i, long long j, long long k) { return (uint64_t)((i & 0x1fffff) | ((j & 0x1fffff) << 21) | ((k & 0x1fffff) << 42)); };
  std::unordered_map<uint64_t, std::vector<uint32_t>> grid;
  grid.reserve(n);
  for (size_t i = 0; i < n; ++i) {
    const float *p = &c.xyz[3 * i];
    grid[key((long long)std::floor(p[0] * inv), (long long)std::floor(p[1] * inv), (long long)std::floor(p[2] * inv))].push_back((uint32_t)i);
  }
  const float r2 = radius * radius;
  for (size_t i = 0; i < n; ++i) {
    const float *p = &c.xyz[3 * i];
    const long long ci = (long long)std::floor(p[0] * inv), cj = (long long)std::floor(p[1] * inv), ck = (long long)std::floor(p[2] * inv);
    double s[3] = {0, 0, 0}, ss[6] = {0, 0, 0, 0, 0, 0};
    int cnt = 0;
    for (long long a = -1; a <= 1; ++a) for (long long b = -1; b <= 1; ++b) for (long long d = -1; d <= 1; ++d) {
      auto it = grid.find(key(ci + a, cj + b, ck + d));
      if (it == grid.end()) continue;
      for (uint32_t j : it->second) {
        const float *q = &c.xyz[3 * j];
        const float dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
        if (dx * dx + dy * dy + dz * dz > r2) continue;
        s[0] += dx; s[1] += dy; s[2] += dz;
        ss[0] += dx * dx; ss[1] += dx * dy; ss[2] += dx * dz; ss[3] += dy * dy; ss[4] += dy * dz; ss[5] += dz * dz;
        ++cnt;
      }
    }
    if (cnt < 3) continue;
    const double m[3] = {s[0] / cnt, s[1] / cnt, s[2] / cnt};
    double C[3][3] = {{ss[0] / cnt - m[0] * m[0], ss[1] / cnt - m[0] * m[1], ss[2] / cnt - m[0] * m[2]},
                      {0, ss[3] / cnt - m[1] * m[1], ss[4] / cnt - m[1] * m[2]}, {0, 0, ss[5] / cnt - m[2] * m[2]}};
    C[1][0] = C[0][1]; C[2][0] = C[0][2]; C[2][1] = C[1][2];
    // smallest eigenvector by Jacobi rotations
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 12; ++sweep)
      for (int a = 0; a < 2; ++a) for (int b = a + 1; b < 3; ++b) {
        if (std::fabs(C[a][b]) < 1e-30) continue;
        const double th = 0.5 * std::atan2(2 * C[a][b], C[b][b] - C[a][a]), cs = std::cos(th), sn = std::sin(th);
        for (int k = 0; k < 3; ++k) { const double x = C[k][a], y = C[k][b]; C[k][a] = cs * x - sn * y; C[k][b] = sn * x + cs * y; }
        for (int k = 0; k < 3; ++k) { const double x = C[a][k], y = C[b][k]; C[a][k] = cs * x - sn * y; C[b][k] = sn * x + cs * y; }
        for (int k = 0; k < 3; ++k) { const double x = V[k][a], y = V[k][b]; V[k][a] = cs * x - sn * y; V[k][b] = sn * x + cs * y; }
      }
    int mi = 0;
    if (C[1][1] < C[mi][mi]) mi = 1;
    if (C[2][2] < C[mi][mi]) mi = 2;
    float nx = (float)V[0][mi], ny = (float)V[1][mi], nz = (float)V[2][mi];
    if ((vp[0] - p[0]) * nx + (vp[1] - p[1]) * ny + (vp[2] - p[2]) * nz < 0) { nx = -nx; ny = -ny; nz = -nz; }
    c.nrm[3 * i] = nx; c.nrm[3 * i + 1] = ny; c.nrm[3 * i + 2] = nz;
  }
}

void frameToObjectSegment(const std::vector<float> &depth_m, int w, int h, const Mat3f &K, const Mat4f &cam_in_handbase, Cloud &object_segment,
                          Cloud *cropped) {
  Cloud scene;
  convert3dOrganized(depth_m, w, h, K, scene);
  passThrough(scene, scene, 2, 0.1f, 2.0f);
  downsamplePointCloud(scene, scene, 0.001f);
  transformPointCloudWithNormals(scene, scene, cam_in_handbase);
  passThrough(scene, scene, 2, -0.12f, 0.05f);
  passThrough(scene, scene, 0, -0.25f, -0.07f);
  passThrough(scene, scene, 1, -0.2f, 0.2f);
  transformPointCloudWithNormals(scene, scene, cam_in_handbase.inverse());
  if (cropped) *cropped = scene;
  // object segment: normals over 3 mm, 3 mm voxels, normals towards the camera (main_realdata_auto.cpp:154-177)
  object_segment = scene;
  if (object_segment.size() == 0) return;
  const float origin[3] = {0, 0, 0};
  estimateNormals(object_segment, 0.003f, origin);
  downsamplePointCloud(object_segment, object_segment, 0.003f);
  removeAllNaNFromPointCloud(object_segment);
  for (size_t i = 0; i < object_segment.size(); ++i) {   // pcl::flipNormalTowardsViewpoint
    float *n = &object_segment.nrm[3 * i];
    const float *p = &object_segment.xyz[3 * i];
    if (-p[0] * n[0] - p[1] * n[1] - p[2] * n[2] < 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
  }
  std::fill(object_segment.conf.begin(), object_segment.conf.end(), 1.f);
}
