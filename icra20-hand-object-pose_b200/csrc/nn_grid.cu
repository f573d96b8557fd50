// nn_grid.cu -- exact nearest-neighbour voxel grid (replaces pcl::KdTreeFLANN / gr::KdTree on the hot path).
//
// The reference rebuilds a kd-tree of the TRANSFORMED model for every hypothesis (PoseEstimator.cpp:263 ->
// Utils.cpp:214 setInputTarget; Utils.cpp:376-379).  Here the model stays fixed, the scene is moved by the
// inverse pose, and ONE structure per (cloud, radius) is shared by every hypothesis of every frame.
//
// Structure: dense voxel grid (edge e, half diagonal h) over the bounding box inflated by the query radius R.
// For a voxel with centre c let d_c = distance from c to its nearest point.  Any query q inside the voxel has its
// nearest neighbour m* within |c - m*| <= d_c + 2h, and if only neighbours within R matter, within R + h.  The
// voxel's list therefore holds { m : |m - c| <= min(d_c + 2h, R + h) } -- a superset of every possible answer, so
// scanning the list is EXACT.  Built by scatter from the points (three passes: min distance, count, fill).
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>

#include "hop_common.cuh"

namespace {

struct GridGeom {
  float ox, oy, oz, e;
  int nx, ny, nz;
  int K;        // scatter half-width in voxels
  float rc2;    // (R + h)^2
  float h2x;    // 2h
  float rc;     // R + h
};

__device__ __forceinline__ float vox_center_d2(const GridGeom &g, int vx, int vy, int vz, float4 p) {
  float cx = g.ox + (vx + 0.5f) * g.e, cy = g.oy + (vy + 0.5f) * g.e, cz = g.oz + (vz + 0.5f) * g.e;
  float dx = p.x - cx, dy = p.y - cy, dz = p.z - cz;
  return dx * dx + dy * dy + dz * dz;
}

// PASS 0: dnn2[v] = min squared distance (as ordered uint)   PASS 1: cnt[v] += 1   PASS 2: fill cand
template <int PASS>
__global__ void grid_scatter_kernel(GridGeom g, const float4 *__restrict__ pw, int n, unsigned int *__restrict__ dnn2,
                                    unsigned int *__restrict__ cnt, const unsigned int *__restrict__ off,
                                    float4 *__restrict__ cand) {
  const int S = 2 * g.K + 1;
  const int S3 = S * S * S;
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    float4 p = pw[i];
    if (!(fabsf(p.x) < 1e20f && fabsf(p.y) < 1e20f && fabsf(p.z) < 1e20f)) continue;  // sentinel / non-finite
    int bx = __float2int_rd((p.x - g.ox) / g.e) - g.K;
    int by = __float2int_rd((p.y - g.oy) / g.e) - g.K;
    int bz = __float2int_rd((p.z - g.oz) / g.e) - g.K;
    for (int k = threadIdx.x; k < S3; k += blockDim.x) {
      int vz = bz + k / (S * S), rem = k % (S * S);
      int vy = by + rem / S, vx = bx + rem % S;
      if ((unsigned)vx >= (unsigned)g.nx || (unsigned)vy >= (unsigned)g.ny || (unsigned)vz >= (unsigned)g.nz) continue;
      float d2 = vox_center_d2(g, vx, vy, vz, p);
      if (d2 > g.rc2) continue;
      size_t v = ((size_t)vz * g.ny + vy) * g.nx + vx;
      if (PASS == 0) {
        atomicMin(&dnn2[v], __float_as_uint(d2));
      } else {
        float lim = fminf(sqrtf(__uint_as_float(dnn2[v])) + g.h2x, g.rc);
        if (d2 <= lim * lim) {
          unsigned int slot = atomicAdd(&cnt[v], 1u);
          if (PASS == 2) cand[(size_t)off[v] + slot] = make_float4(p.x, p.y, p.z, __int_as_float(i));
        }
      }
    }
  }
}

// writes cell[v] = (offset,count) and orders each list by point index (deterministic ties)
__global__ void grid_finalize_kernel(size_t n_vox, const unsigned int *__restrict__ off, const unsigned int *__restrict__ cnt,
                                     uint2 *__restrict__ cell, float4 *__restrict__ cand, unsigned int *__restrict__ max_list) {
  size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int c = 0;
  if (v < n_vox) {
    c = cnt[v];
    unsigned int o = off[v];
    cell[v] = make_uint2(o, c);
    float4 *l = cand + o;
    for (unsigned int a = 1; a < c; ++a) {
      float4 key = l[a];
      int ki = __float_as_int(key.w);
      int b = (int)a - 1;
      while (b >= 0 && __float_as_int(l[b].w) > ki) { l[b + 1] = l[b]; --b; }
      l[b + 1] = key;
    }
  }
  // block max -> one atomic
  for (int o2 = 16; o2 > 0; o2 >>= 1) c = max(c, __shfl_xor_sync(0xffffffffu, c, o2));
  if ((threadIdx.x & 31) == 0 && c) atomicMax(max_list, c);
}

__global__ void fill_u32_kernel(unsigned int *p, size_t n, unsigned int v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

__global__ void nn_query_kernel(NNGridDev g, const float *__restrict__ q, int nq, int32_t *__restrict__ idx, float *__restrict__ d2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  float bd; float4 bp;
  int j = nn_query(g, q[3 * i], q[3 * i + 1], q[3 * i + 2], bd, bp);
  if (j >= 0 && bd > g.radius * g.radius) j = -1;
  idx[i] = j;
  d2[i] = j >= 0 ? bd : 3.0e38f;
}

}  // namespace

void hop_free_nn_grid(NNGridHost *g) {
  if (!g) return;
  cudaFree(g->d_cell);
  cudaFree(g->d_cand);
  delete g;
}

static float auto_voxel(const hop_cloud *c, float radius) {
  // point spacing estimate from the bounding-box surface (clouds on this path are surface samples)
  float dx = std::max(c->bbox_max[0] - c->bbox_min[0], 1e-6f), dy = std::max(c->bbox_max[1] - c->bbox_min[1], 1e-6f),
        dz = std::max(c->bbox_max[2] - c->bbox_min[2], 1e-6f);
  float area = dx * dy + dy * dz + dz * dx;  // ~ half the box surface
  float spacing = std::sqrt(area / std::max(c->n, 1));
  return std::min(std::max(spacing, radius / 12.f), radius * 0.5f);
}

int hop_build_nn_grid(hop_ctx *ctx, hop_cloud *cloud, float radius, float voxel, NNGridHost **out) {
  if (!cloud || cloud->n <= 0 || !(radius > 0.f)) { ctx->err = "hop_build_nn_grid: empty cloud or bad radius"; return HOP_EINVAL; }
  NNGridHost *G = *out ? *out : new NNGridHost();
  float e = voxel > 0.f ? voxel : auto_voxel(cloud, radius);
  const int64_t kMaxVox = 48ll << 20;
  GridGeom g;
  for (;;) {
    float pad = radius + 2.f * e;
    g.e = e;
    g.ox = cloud->bbox_min[0] - pad; g.oy = cloud->bbox_min[1] - pad; g.oz = cloud->bbox_min[2] - pad;
    g.nx = (int)std::ceil((cloud->bbox_max[0] - cloud->bbox_min[0] + 2.f * pad) / e) + 1;
    g.ny = (int)std::ceil((cloud->bbox_max[1] - cloud->bbox_min[1] + 2.f * pad) / e) + 1;
    g.nz = (int)std::ceil((cloud->bbox_max[2] - cloud->bbox_min[2] + 2.f * pad) / e) + 1;
    if ((int64_t)g.nx * g.ny * g.nz <= kMaxVox) break;
    e *= 1.26f;
  }
  const float h = 0.5f * std::sqrt(3.f) * e * 1.002f + 1e-7f;  // half diagonal, with slack for rounding at voxel faces
  const float R = radius * 1.0005f + 1e-7f;
  g.rc = R + h; g.rc2 = g.rc * g.rc; g.h2x = 2.f * h;
  g.K = (int)std::ceil(g.rc / e) + 1;
  const int64_t n_vox = (int64_t)g.nx * g.ny * g.nz;

  // aux arrays in scratch: dnn2 | cnt | off | total | maxlist | cub temp
  size_t cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (unsigned int *)nullptr, (unsigned int *)nullptr, (int)n_vox, ctx->stream);
  size_t aux = sizeof(unsigned int) * (size_t)n_vox;
  size_t need = 3 * aux + 256 + cub_bytes + 256;
  char *base = (char *)ctx->ensure_scratch(need);
  if (!base) { ctx->err = "hop_build_nn_grid: scratch allocation failed"; if (!*out) delete G; return HOP_ENOMEM; }
  unsigned int *d_dnn2 = (unsigned int *)base, *d_cnt = (unsigned int *)(base + aux), *d_off = (unsigned int *)(base + 2 * aux);
  unsigned int *d_small = (unsigned int *)(base + 3 * aux);  // [0] = max list
  void *d_cub = base + 3 * aux + 256;

  if (G->cap_vox < n_vox) {
    cudaFree(G->d_cell); G->d_cell = nullptr;
    HOP_CUDA(ctx, cudaMalloc(&G->d_cell, sizeof(uint2) * (size_t)n_vox));
    G->cap_vox = n_vox;
  }
  const int fill_blocks = (int)std::min<int64_t>((n_vox + 255) / 256, 148 * 16);
  fill_u32_kernel<<<fill_blocks, 256, 0, ctx->stream>>>(d_dnn2, (size_t)n_vox, 0x7f800000u);
  HOP_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, aux, ctx->stream));
  HOP_CUDA(ctx, cudaMemsetAsync(d_small, 0, 256, ctx->stream));
  const int S3 = (2 * g.K + 1) * (2 * g.K + 1) * (2 * g.K + 1);
  const int threads = S3 >= 1024 ? 256 : (S3 >= 256 ? 128 : 64);
  const int blocks = cloud->n;
  grid_scatter_kernel<0><<<blocks, threads, 0, ctx->stream>>>(g, cloud->d_pw, cloud->n, d_dnn2, d_cnt, d_off, nullptr);
  grid_scatter_kernel<1><<<blocks, threads, 0, ctx->stream>>>(g, cloud->d_pw, cloud->n, d_dnn2, d_cnt, d_off, nullptr);
  cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_cnt, d_off, (int)n_vox, ctx->stream);
  ctx->launches += 5;
  // total = off[last] + cnt[last]
  unsigned int tail[2];
  HOP_CUDA(ctx, cudaMemcpyAsync(&tail[0], d_off + (n_vox - 1), sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
  HOP_CUDA(ctx, cudaMemcpyAsync(&tail[1], d_cnt + (n_vox - 1), sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
  HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const int64_t total = (int64_t)tail[0] + tail[1];
  if (G->cap_cand < total || !G->d_cand) {
    cudaFree(G->d_cand); G->d_cand = nullptr;
    int64_t cap = std::max<int64_t>(total + total / 4, 1024);
    HOP_CUDA(ctx, cudaMalloc(&G->d_cand, sizeof(float4) * (size_t)cap));
    G->cap_cand = cap;
  }
  HOP_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, aux, ctx->stream));
  grid_scatter_kernel<2><<<blocks, threads, 0, ctx->stream>>>(g, cloud->d_pw, cloud->n, d_dnn2, d_cnt, d_off, G->d_cand);
  grid_finalize_kernel<<<(unsigned)((n_vox + 127) / 128), 128, 0, ctx->stream>>>((size_t)n_vox, d_off, d_cnt, G->d_cell, G->d_cand, d_small);
  ctx->launches += 2;
  unsigned int max_list = 0;
  HOP_CUDA(ctx, cudaMemcpyAsync(&max_list, d_small, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
  HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  HOP_CUDA(ctx, cudaGetLastError());

  G->radius = radius; G->voxel = e; G->n_vox = n_vox; G->n_cand = total; G->max_list = (int)max_list;
  G->dev.ox = g.ox; G->dev.oy = g.oy; G->dev.oz = g.oz; G->dev.inv_e = 1.f / e;
  G->dev.nx = g.nx; G->dev.ny = g.ny; G->dev.nz = g.nz; G->dev.radius = radius;
  G->dev.cell = G->d_cell; G->dev.cand = G->d_cand;
  *out = G;
  return HOP_OK;
}

// cache lookup: a grid built for radius r serves any query radius <= r; rebuilt when the cloud changed
int hop_get_nn_grid(hop_ctx *ctx, hop_cloud *cloud, float radius, float voxel, NNGridHost **out) {
  for (size_t i = 0; i < cloud->grids.size(); ++i) {
    NNGridHost *G = cloud->grids[i];
    if (std::fabs(G->radius - radius) <= 1e-9f + 1e-6f * radius && (voxel <= 0.f || std::fabs(G->voxel - voxel) < 1e-9f)) {
      if (cloud->grid_version[i] != cloud->version) {
        int rc = hop_build_nn_grid(ctx, cloud, radius, voxel, &cloud->grids[i]);
        if (rc != HOP_OK) return rc;
        cloud->grid_version[i] = cloud->version;
      }
      *out = cloud->grids[i];
      return HOP_OK;
    }
  }
  NNGridHost *G = nullptr;
  int rc = hop_build_nn_grid(ctx, cloud, radius, voxel, &G);
  if (rc != HOP_OK) return rc;
  cloud->grids.push_back(G);
  cloud->grid_version.push_back(cloud->version);
  *out = G;
  return HOP_OK;
}

extern "C" int hop_cloud_prepare_nn(hop_ctx *ctx, hop_cloud *cloud, float radius, float voxel, int64_t *stats) {
  if (!ctx || !cloud) return HOP_EINVAL;
  NNGridHost *G = nullptr;
  int rc = hop_get_nn_grid(ctx, cloud, radius, voxel, &G);
  if (rc != HOP_OK) return rc;
  if (stats) {
    stats[0] = G->n_vox; stats[1] = G->n_cand; stats[2] = G->max_list;
    stats[3] = G->n_vox * (int64_t)sizeof(uint2) + G->n_cand * (int64_t)sizeof(float4);
  }
  return HOP_OK;
}

extern "C" int hop_cloud_drop_nn(hop_ctx *ctx, hop_cloud *cloud) {
  if (!ctx || !cloud) return HOP_EINVAL;
  cloud->version++;
  return HOP_OK;
}

extern "C" int hop_cloud_nn_query(hop_ctx *ctx, hop_cloud *cloud, float radius, const float *queries, int nq, int32_t *idx, float *d2) {
  if (!ctx || !cloud || !queries || !idx || !d2 || nq < 0) return HOP_EINVAL;
  if (nq == 0) return HOP_OK;
  NNGridHost *G = nullptr;
  int rc = hop_get_nn_grid(ctx, cloud, radius, 0.f, &G);
  if (rc != HOP_OK) return rc;
  float *d_q = nullptr; int32_t *d_idx = nullptr; float *d_d2 = nullptr;
  HOP_CUDA(ctx, cudaMalloc(&d_q, sizeof(float) * 3 * (size_t)nq));
  HOP_CUDA(ctx, cudaMalloc(&d_idx, sizeof(int32_t) * (size_t)nq));
  HOP_CUDA(ctx, cudaMalloc(&d_d2, sizeof(float) * (size_t)nq));
  HOP_CUDA(ctx, cudaMemcpyAsync(d_q, queries, sizeof(float) * 3 * (size_t)nq, cudaMemcpyHostToDevice, ctx->stream));
  nn_query_kernel<<<(nq + 127) / 128, 128, 0, ctx->stream>>>(G->dev, d_q, nq, d_idx, d_d2);
  ctx->launches += 1;
  HOP_CUDA(ctx, cudaMemcpyAsync(idx, d_idx, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, ctx->stream));
  HOP_CUDA(ctx, cudaMemcpyAsync(d2, d_d2, sizeof(float) * (size_t)nq, cudaMemcpyDeviceToHost, ctx->stream));
  HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(d_q); cudaFree(d_idx); cudaFree(d_d2);
  return HOP_OK;
}
