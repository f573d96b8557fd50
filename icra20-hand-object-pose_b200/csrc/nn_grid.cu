// nn_grid.cu -- exact nearest-neighbour voxel grid (replaces pcl::KdTreeFLANN / gr::KdTree on the hot path).
//
// The reference rebuilds a kd-tree of the TRANSFORMED model for every hypothesis (PoseEstimator.cpp:263 ->
// Utils.cpp:214 setInputTarget; Utils.cpp:376-379).  Here the model stays fixed, the scene is moved by the
// inverse pose, and ONE structure per (cloud, radius) is shared by every hypothesis of every frame.
//
// Structure: dense voxel grid (edge e, half diagonal h) over the bounding box inflated by the query radius R.
// For a voxel V with centre c let m* be the point nearest to c and d_c = |c - m*|.  The voxel's list holds every
// point m that passes ALL of these necessary conditions for being the nearest neighbour (within R) of some q in V:
//   (1) |m - c| <= R + h                       only neighbours within R matter;
//   (2) |m - c| <= d_c + 2h                    triangle inequality through the centre;
//   (3) |m - c|^2 - d_c^2 <= e * |m - m*|_1    m is not DOMINATED by m* on V: |q-m|^2 - |q-m*|^2 is linear in q, its
//                                               minimum over the cube is (|c-m|^2 - d_c^2) - e*|m - m*|_1; when that is
//                                               positive m* is strictly nearer than m everywhere in V.
// The list is therefore a superset of every possible answer and scanning it is EXACT; (3) keeps lists at a handful
// of entries even for voxels a centimetre away from the surface, where (2) alone admits hundreds.
// Built by scatter from the points (three passes: nearest-to-centre, count, fill), with no host synchronisation:
// the candidate buffer is sized from a geometric upper bound and the fill pass is bounds-checked.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "hop_common.cuh"

namespace {

struct GridGeom {
  float ox, oy, oz, e;
  int nx, ny, nz;
  int K;        // scatter half-width in voxels
  float rc2;    // (R + h)^2
  float h2x;    // 2h
  float rc;     // R + h
  float e_dom;  // voxel edge with head room, for the dominance test
};

__device__ __forceinline__ float vox_center_d2(const GridGeom &g, int vx, int vy, int vz, float4 p) {
  float cx = g.ox + (vx + 0.5f) * g.e, cy = g.oy + (vy + 0.5f) * g.e, cz = g.oz + (vz + 0.5f) * g.e;
  float dx = p.x - cx, dy = p.y - cy, dz = p.z - cz;
  return dx * dx + dy * dy + dz * dz;
}

// PASS 0: near[v] = min over points of (d^2 bits << 32 | index)   PASS 1: cnt[v] += 1   PASS 2: fill cand
template <int PASS>
__global__ void grid_scatter_kernel(GridGeom g, const float4 *__restrict__ pw, int n, unsigned long long *__restrict__ near,
                                    unsigned int *__restrict__ cnt, const unsigned int *__restrict__ off,
                                    float4 *__restrict__ cand, unsigned int cap, unsigned int *__restrict__ overflow) {
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
      if (d2 > g.rc2) continue;                                                        // (1)
      size_t v = nn_vox_index(vx, vy, vz, g.nx, g.ny);
      if (PASS == 0) {
        atomicMin(&near[v], ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned int)i);
      } else {
        const unsigned long long nr = near[v];
        const float dc2 = __uint_as_float((unsigned int)(nr >> 32));
        const float lim = sqrtf(dc2) + g.h2x;
        if (d2 > lim * lim) continue;                                                  // (2)
        const float4 ms = pw[(unsigned int)nr];
        const float l1 = fabsf(p.x - ms.x) + fabsf(p.y - ms.y) + fabsf(p.z - ms.z);
        // (3), with head room for the rounding of d2/dc2 and of the query's voxel assignment at the faces
        if (d2 - dc2 > g.e_dom * l1 + 8e-6f * d2 + 1e-12f) continue;
        unsigned int slot = atomicAdd(&cnt[v], 1u);
        if (PASS == 2) {
          const unsigned int at = off[v] + slot;
          if (at < cap) cand[at] = make_float4(p.x, p.y, p.z, __int_as_float(i));
          else *overflow = 1u;
        }
      }
    }
  }
}

// per voxel: order the list by point index (deterministic ties), drop every entry that another entry of the list
// DOMINATES on the voxel's cube (condition (3) between any two list members, not only against m*), write
// cell[v] = (offset, count).  "m' dominates m" is a strict partial order, so every dropped point is dominated by a
// kept one and can never be the nearest neighbour of a query inside the voxel: the list stays exact.
__global__ void grid_finalize_kernel(GridGeom g, size_t n_vox, const unsigned int *__restrict__ off, const unsigned int *__restrict__ cnt,
                                     uint2 *__restrict__ cell, float4 *__restrict__ cand, unsigned int cap,
                                     unsigned int *__restrict__ max_list) {
  size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int c = 0;
  if (v < n_vox) {
    c = cnt[v];
    unsigned int o = off[v];
    if (o >= cap) c = 0; else if (o + c > cap) c = cap - o;  // (overflow is flagged by the fill pass)
    float4 *l = cand + o;
    for (unsigned int a = 1; a < c; ++a) {
      float4 key = l[a];
      int ki = __float_as_int(key.w);
      int b = (int)a - 1;
      while (b >= 0 && __float_as_int(l[b].w) > ki) { l[b + 1] = l[b]; --b; }
      l[b + 1] = key;
    }
    if (c > 1) {
      int vx, vy, vz;
      nn_vox_coords(v, g.nx, g.ny, vx, vy, vz);
      unsigned int kept = 0;
      for (unsigned int a = 0; a < c; ++a) {
        const float4 m = l[a];
        const float d2 = vox_center_d2(g, vx, vy, vz, m);
        bool dominated = false;
        // members before `kept` were compacted to the front; members from `a` on are still in place: together they
        // are the whole original list minus entries already dropped (dropping is transitive-safe, see above)
        for (unsigned int b = 0; b < c && !dominated; ++b) {
          if (b >= kept && b <= a) { if (b < a) b = a; continue; }
          const float4 q = l[b];
          const float e2 = vox_center_d2(g, vx, vy, vz, q);
          const float l1 = fabsf(m.x - q.x) + fabsf(m.y - q.y) + fabsf(m.z - q.z);
          dominated = d2 - e2 > g.e_dom * l1 + 8e-6f * d2 + 1e-12f;
        }
        if (!dominated) l[kept++] = m;
      }
      c = kept;
    }
    cell[v] = make_uint2(o, c);
    if (v == n_vox - 1) { max_list[1] = off[v] + cnt[v]; }   // entries allocated (before the pairwise pruning)
  }
  // block max -> one atomic; entries kept -> one atomic per warp
  unsigned int mx = c, sum = c;
  for (int o2 = 16; o2 > 0; o2 >>= 1) { mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o2)); sum += __shfl_xor_sync(0xffffffffu, sum, o2); }
  if ((threadIdx.x & 31) == 0 && mx) { atomicMax(max_list, mx); atomicAdd(max_list + 3, sum); }
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

// ---- Morton (Z-curve) query order ------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int spread10(unsigned int v) {  // 10 bits -> every third bit
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__global__ void morton_keys_kernel(const float4 *__restrict__ pw, int n, int n_padded, float ox, float oy, float oz, float scale,
                                   unsigned int *__restrict__ keys, int *__restrict__ idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_padded) return;
  unsigned int k = 0xffffffffu;  // padding and non-finite points go last
  if (i < n) {
    const float4 p = pw[i];
    if (fabsf(p.x) < 1e20f && fabsf(p.y) < 1e20f && fabsf(p.z) < 1e20f) {
      const unsigned int x = (unsigned int)fminf(fmaxf((p.x - ox) * scale, 0.f), 1023.f), y = (unsigned int)fminf(fmaxf((p.y - oy) * scale, 0.f), 1023.f),
                         z = (unsigned int)fminf(fmaxf((p.z - oz) * scale, 0.f), 1023.f);
      k = spread10(x) | (spread10(y) << 1) | (spread10(z) << 2);
    } else k = 0xfffffffeu;
  }
  keys[i] = k;
  idx[i] = i;
}

__global__ void gather_cloud_kernel(const float4 *__restrict__ pw, const float4 *__restrict__ nv, const int *__restrict__ idx, int n_padded,
                                    float4 *__restrict__ pw_q, float4 *__restrict__ nv_q) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_padded) return;
  const int j = idx[i];
  pw_q[i] = pw[j];
  nv_q[i] = nv[j];
}

}  // namespace

int hop_cloud_query_order(hop_ctx *ctx, hop_cloud *c) {
  if (c->q_version == c->version && c->d_pw_q) return HOP_OK;
  if (c->n_padded <= 0) return HOP_OK;
  if (c->q_capacity < c->n_padded) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(c->d_pw_q); cudaFree(c->d_nv_q);
    c->d_pw_q = c->d_nv_q = nullptr;
    HOP_CUDA(ctx, cudaMalloc(&c->d_pw_q, sizeof(float4) * (size_t)c->n_padded));
    HOP_CUDA(ctx, cudaMalloc(&c->d_nv_q, sizeof(float4) * (size_t)c->n_padded));
    c->q_capacity = c->n_padded;
  }
  const int np = c->n_padded;
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (unsigned int *)nullptr, (unsigned int *)nullptr, (int *)nullptr, (int *)nullptr, np, 0, 32, ctx->stream);
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t arr = up(sizeof(int) * (size_t)np);
  char *base = (char *)ctx->ensure_scratch(4 * arr + sort_bytes + 256);
  if (!base) { ctx->err = "hop_cloud_query_order: scratch allocation failed"; return HOP_ENOMEM; }
  unsigned int *k_in = (unsigned int *)base, *k_out = (unsigned int *)(base + arr);
  int *i_in = (int *)(base + 2 * arr), *i_out = (int *)(base + 3 * arr);
  const float ext = std::max(std::max(c->bbox_max[0] - c->bbox_min[0], c->bbox_max[1] - c->bbox_min[1]), std::max(c->bbox_max[2] - c->bbox_min[2], 1e-9f));
  morton_keys_kernel<<<(np + 255) / 256, 256, 0, ctx->stream>>>(c->d_pw, c->n, np, c->bbox_min[0], c->bbox_min[1], c->bbox_min[2], 1023.f / ext, k_in, i_in);
  cub::DeviceRadixSort::SortPairs(base + 4 * arr, sort_bytes, k_in, k_out, i_in, i_out, np, 0, 32, ctx->stream);
  gather_cloud_kernel<<<(np + 255) / 256, 256, 0, ctx->stream>>>(c->d_pw, c->d_nv, i_out, np, c->d_pw_q, c->d_nv_q);
  ctx->launches += 3;
  HOP_CUDA(ctx, cudaGetLastError());
  c->q_version = c->version;
  return HOP_OK;
}

void hop_free_nn_grid(NNGridHost *g) {
  if (!g) return;
  cudaFree(g->d_cell);
  cudaFree(g->d_cand);
  cudaFree(g->d_info);
  if (g->ready) cudaEventDestroy(g->ready);
  delete g;
}

// The voxel edge when the caller names none: `scale` x the point spacing estimated from the bounding-box surface (clouds on this path are
// surface samples), not below radius / 12 x scale, not above radius x max_frac.  Both hot kernels are bound by the L1TEX data pipe and
// most of its wavefronts are candidate-list entries: shorter lists beat fewer cells.  Measured at the headline size
// (10 k-point model, gpurun_out/r03i, r03j): scale 1.0 -> ICP 16.0 ms + LCP 4.6; 0.85 -> 15.5 + 4.2; 0.7 -> 14.9 + 3.9; 0.5 -> 15.4 + 3.5 (and
// 4x the memory); 1.25 -> 16.9 + 5.2.  0.7: 52 MB instead of 25 MB for the model's ICP grid, +8 % hypotheses/s there, +14 % at 50 k x 50 k,
// neutral on the small batches and on the frame path.
static float auto_voxel(const hop_cloud *c, float radius, float max_frac, float scale) {
  float dx = std::max(c->bbox_max[0] - c->bbox_min[0], 1e-6f), dy = std::max(c->bbox_max[1] - c->bbox_min[1], 1e-6f),
        dz = std::max(c->bbox_max[2] - c->bbox_min[2], 1e-6f);
  float area = dx * dy + dy * dz + dz * dx;  // ~ half the box surface
  float spacing = std::sqrt(area / std::max(c->n, 1));
  return std::min(std::max(spacing, radius / 12.f) * scale, radius * max_frac);
}

// the automatic voxel edge with the cap at `max_frac` x radius (callers that know more than the cloud: hop_lcp_scene_voxel, api.cu)
float hop_auto_voxel(const hop_ctx *ctx, const hop_cloud *cloud, float radius, float max_frac) {
  return auto_voxel(cloud, radius, std::min(ctx->tune.voxel_max_frac, max_frac), ctx->tune.voxel_scale);
}

int hop_build_nn_grid(hop_ctx *ctx, hop_cloud *cloud, float radius, float voxel, NNGridHost **out) {
  if (!cloud || cloud->n <= 0 || !(radius > 0.f)) { ctx->err = "hop_build_nn_grid: empty cloud or bad radius"; return HOP_EINVAL; }
  ProfScope ps(ctx, HOP_PROF_NN_BUILD);
  NNGridHost *G = *out ? *out : new NNGridHost();
  // a grid with a small radius (LCP: 1 mm) is capped at a fraction of the radius; for a cloud that stays (hop_cloud_hint_static: the model)
  // half of it -- the lists of lcp_score_kernel's first query get shorter (its 4.3 ms at the headline size go to 3.5-3.9, gpurun_out/r03l) --
  // while a per-frame scene grid keeps the whole radius: its build is (radius / edge)^3 scatter work on every frame
  const float max_frac = cloud->is_static ? std::min(ctx->tune.voxel_max_frac, 0.5f) : ctx->tune.voxel_max_frac;
  float e = voxel > 0.f ? voxel : auto_voxel(cloud, radius, max_frac, ctx->tune.voxel_scale);
  const int64_t kMaxVox = 48ll << 20;
  GridGeom g;
  for (;;) {
    float pad = radius + 2.f * e;
    g.e = e;
    g.ox = cloud->bbox_min[0] - pad; g.oy = cloud->bbox_min[1] - pad; g.oz = cloud->bbox_min[2] - pad;
    auto tiles = [](float extent, float edge) { return (((int)std::ceil(extent / edge) + 1) + 3) / 4 * 4; };  // whole 4x4x4 tiles
    g.nx = tiles(cloud->bbox_max[0] - cloud->bbox_min[0] + 2.f * pad, e);
    g.ny = tiles(cloud->bbox_max[1] - cloud->bbox_min[1] + 2.f * pad, e);
    g.nz = tiles(cloud->bbox_max[2] - cloud->bbox_min[2] + 2.f * pad, e);
    if ((int64_t)g.nx * g.ny * g.nz <= kMaxVox) break;
    e *= 1.26f;
  }
  const float h = 0.5f * std::sqrt(3.f) * e * 1.002f + 1e-7f;  // half diagonal, with slack for rounding at voxel faces
  const float R = radius * 1.0005f + 1e-7f;
  g.rc = R + h; g.rc2 = g.rc * g.rc; g.h2x = 2.f * h;
  g.e_dom = e * 1.004f + 2e-7f;
  g.K = (int)std::ceil(g.rc / e) + 1;
  const int64_t n_vox = (int64_t)g.nx * g.ny * g.nz;

  // aux arrays in scratch: near (u64) | cnt | off | cub temp
  size_t cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (unsigned int *)nullptr, (unsigned int *)nullptr, (int)n_vox, ctx->stream);
  size_t aux = (sizeof(unsigned int) * (size_t)n_vox + 255) / 256 * 256;
  size_t need = 4 * aux + cub_bytes + 256;
  char *base = (char *)ctx->ensure_scratch(need);
  if (!base) { ctx->err = "hop_build_nn_grid: scratch allocation failed"; if (!*out) delete G; return HOP_ENOMEM; }
  unsigned long long *d_near = (unsigned long long *)base;
  unsigned int *d_cnt = (unsigned int *)(base + 2 * aux), *d_off = (unsigned int *)(base + 3 * aux);
  void *d_cub = base + 4 * aux;

  if (!G->d_info) HOP_CUDA(ctx, cudaMalloc(&G->d_info, 4 * sizeof(unsigned int)));  // [0] max list [1] entries allocated [2] overflow [3] entries kept
  if (G->cap_vox < n_vox) {
    cudaFree(G->d_cell); G->d_cell = nullptr;
    HOP_CUDA(ctx, cudaMalloc(&G->d_cell, sizeof(uint2) * (size_t)n_vox));
    G->cap_vox = n_vox;
  }
  HOP_CUDA(ctx, cudaMemsetAsync(d_near, 0xff, sizeof(unsigned long long) * (size_t)n_vox, ctx->stream));
  HOP_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, aux, ctx->stream));
  HOP_CUDA(ctx, cudaMemsetAsync(G->d_info, 0, 4 * sizeof(unsigned int), ctx->stream));
  const int S3 = (2 * g.K + 1) * (2 * g.K + 1) * (2 * g.K + 1);
  const int threads = S3 >= 1024 ? 256 : (S3 >= 256 ? 128 : 64);
  const int blocks = cloud->n;
  grid_scatter_kernel<0><<<blocks, threads, 0, ctx->stream>>>(g, cloud->d_pw, cloud->n, d_near, d_cnt, d_off, nullptr, 0u, G->d_info + 2);
  grid_scatter_kernel<1><<<blocks, threads, 0, ctx->stream>>>(g, cloud->d_pw, cloud->n, d_near, d_cnt, d_off, nullptr, 0u, G->d_info + 2);
  cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_cnt, d_off, (int)n_vox, ctx->stream);
  ctx->launches += 4;

  // Candidate capacity.  A voxel lists a point only when its centre lies within rc of it, and the cubes of such
  // voxels fit in a ball of radius rc + h: at most (4/3) pi (rc + h)^3 / e^3 entries per point.  When that bound is
  // affordable the buffer is sized from it and the build never synchronises with the host (per-frame scene grids);
  // otherwise (large static model grids, built once) the exact total is read back.
  const double per_point = 4.18879 * std::pow((double)(g.rc + h) / e, 3.0);
  const double bound = per_point * cloud->n + 1024.0;
  int64_t need_cap;
  if (bound * sizeof(float4) <= 512.0 * 1024 * 1024) {
    need_cap = (int64_t)bound;
  } else {
    unsigned int tail[2];
    HOP_CUDA(ctx, cudaMemcpyAsync(&tail[0], d_off + (n_vox - 1), sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    HOP_CUDA(ctx, cudaMemcpyAsync(&tail[1], d_cnt + (n_vox - 1), sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    need_cap = (int64_t)tail[0] + tail[1] + 16;
  }
  if (G->cap_cand < need_cap || !G->d_cand) {
    if (G->d_cand) { HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(G->d_cand); G->d_cand = nullptr; }
    HOP_CUDA(ctx, cudaMalloc(&G->d_cand, sizeof(float4) * (size_t)need_cap));
    G->cap_cand = need_cap;
  }
  const unsigned int cap = (unsigned int)std::min<int64_t>(G->cap_cand, 0xffffffffll);
  HOP_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, aux, ctx->stream));
  grid_scatter_kernel<2><<<blocks, threads, 0, ctx->stream>>>(g, cloud->d_pw, cloud->n, d_near, d_cnt, d_off, G->d_cand, cap, G->d_info + 2);
  grid_finalize_kernel<<<(unsigned)((n_vox + 127) / 128), 128, 0, ctx->stream>>>(g, (size_t)n_vox, d_off, d_cnt, G->d_cell, G->d_cand, cap, G->d_info);
  ctx->launches += 2;
  HOP_CUDA(ctx, cudaGetLastError());

  G->radius = radius; G->voxel = e; G->n_vox = n_vox; G->n_cand = -1; G->max_list = -1;  // statistics are read lazily
  G->dev.ox = g.ox; G->dev.oy = g.oy; G->dev.oz = g.oz; G->dev.inv_e = 1.f / e;
  G->dev.nx = g.nx; G->dev.ny = g.ny; G->dev.nz = g.nz; G->dev.radius = radius;
  G->dev.cell = G->d_cell; G->dev.cand = G->d_cand;
  *out = G;
  return HOP_OK;
}

// device-side statistics of the last build (synchronises)
static int grid_fetch_stats(hop_ctx *ctx, NNGridHost *G) {
  unsigned int info[4] = {0, 0, 0, 0};
  HOP_CUDA(ctx, cudaMemcpyAsync(info, G->d_info, sizeof(info), cudaMemcpyDeviceToHost, ctx->stream));
  HOP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  G->max_list = (int)info[0]; G->n_cand = info[3];  // entries kept after the pairwise pruning ([1] = allocated)
  if (info[2]) { ctx->err = "nn grid: candidate buffer overflow (bound violated)"; return HOP_ENOMEM; }
  return HOP_OK;
}

// cache lookup: a grid built for radius r serves any query radius <= r; rebuilt when the cloud changed
static int get_nn_grid_on_current_stream(hop_ctx *ctx, hop_cloud *cloud, float radius, float voxel, NNGridHost **out);

int hop_get_nn_grid(hop_ctx *ctx, hop_cloud *cloud, float radius, float voxel, NNGridHost **out) {
  // a rebuild on this stream must come after a build of the same grid still running on the side stream
  if (ctx->side && ctx->stream != ctx->side) hop_cloud_join_pending(ctx, cloud);
  return get_nn_grid_on_current_stream(ctx, cloud, radius, voxel, out);
}

// an up-to-date grid of this radius as it is (whatever its voxel edge); built with `voxel_if_built` (0 = automatic) when there is none
int hop_get_nn_grid_any(hop_ctx *ctx, hop_cloud *cloud, float radius, float voxel_if_built, NNGridHost **out) {
  if (ctx->side && ctx->stream != ctx->side) hop_cloud_join_pending(ctx, cloud);
  for (size_t i = 0; i < cloud->grids.size(); ++i) {
    NNGridHost *G = cloud->grids[i];
    if (std::fabs(G->radius - radius) <= 1e-9f + 1e-6f * radius && cloud->grid_version[i] == cloud->version) { *out = G; return HOP_OK; }
  }
  return get_nn_grid_on_current_stream(ctx, cloud, radius, voxel_if_built, out);
}

static int get_nn_grid_on_current_stream(hop_ctx *ctx, hop_cloud *cloud, float radius, float voxel, NNGridHost **out) {
  for (size_t i = 0; i < cloud->grids.size(); ++i) {
    NNGridHost *G = cloud->grids[i];
    // one grid per (cloud, radius): a request that names another voxel edge rebuilds it in place (an edge derived from the cloud's
    // extent changes a little with every frame: a grid per edge would pile up)
    if (std::fabs(G->radius - radius) <= 1e-9f + 1e-6f * radius) {
      const bool other_voxel = voxel > 0.f && std::fabs(G->voxel - voxel) > 1e-6f * voxel;
      if (cloud->grid_version[i] != cloud->version || other_voxel) {
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
  HOP_ENTER(ctx);
  if (!ctx || !cloud) return HOP_EINVAL;
  NNGridHost *G = nullptr;
  int rc = hop_get_nn_grid(ctx, cloud, radius, voxel, &G);
  if (rc != HOP_OK) return rc;
  if (stats) {
    rc = grid_fetch_stats(ctx, G);
    if (rc != HOP_OK) return rc;
    stats[0] = G->n_vox; stats[1] = G->n_cand; stats[2] = G->max_list;
    stats[3] = G->n_vox * (int64_t)sizeof(uint2) + G->n_cand * (int64_t)sizeof(float4);
  }
  return HOP_OK;
}

// Builds (or refreshes) the cloud's grid for `radius` on the context's side stream: ordered after everything enqueued on the main
// stream so far (the cloud's upload, the last consumers of the grid's previous contents), concurrent with whatever the main
// stream is given next.  The first main-stream call that looks the grid up waits for it (hop_get_nn_grid).
extern "C" int hop_cloud_prepare_nn_async(hop_ctx *ctx, hop_cloud *cloud, float radius, float voxel) {
  HOP_ENTER(ctx);
  if (!ctx || !cloud) return HOP_EINVAL;
  if (!ctx->side) {
    HOP_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking));
    HOP_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  }
  hop_cloud_join_pending(ctx, cloud);
  HOP_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
  HOP_CUDA(ctx, cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
  cudaStream_t main_stream = ctx->stream;
  ctx->stream = ctx->side;
  std::swap(ctx->d_scratch, ctx->d_side_scratch); std::swap(ctx->scratch_bytes, ctx->side_scratch_bytes);
  NNGridHost *G = nullptr;
  int rc = get_nn_grid_on_current_stream(ctx, cloud, radius, voxel, &G);
  cudaError_t e = cudaSuccess;
  if (rc == HOP_OK) {
    if (!G->ready) e = cudaEventCreateWithFlags(&G->ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(G->ready, ctx->side);
    G->pending = (e == cudaSuccess);
  }
  std::swap(ctx->d_scratch, ctx->d_side_scratch); std::swap(ctx->scratch_bytes, ctx->side_scratch_bytes);
  ctx->stream = main_stream;
  if (rc != HOP_OK) return rc;
  HOP_CUDA(ctx, e);
  return HOP_OK;
}

extern "C" int hop_cloud_hint_static(hop_ctx *ctx, hop_cloud *cloud, int is_static) {
  HOP_ENTER(ctx);
  if (!ctx || !cloud) return HOP_EINVAL;
  cloud->is_static = is_static != 0;   // (applies to grids built from now on; results do not depend on it)
  return HOP_OK;
}

extern "C" int hop_cloud_drop_nn(hop_ctx *ctx, hop_cloud *cloud) {
  HOP_ENTER(ctx);
  if (!ctx || !cloud) return HOP_EINVAL;
  cloud->version++;
  return HOP_OK;
}

extern "C" int hop_cloud_nn_query(hop_ctx *ctx, hop_cloud *cloud, float radius, const float *queries, int nq, int32_t *idx, float *d2) {
  HOP_ENTER(ctx);
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
