// segment.cu — the stage in front of the path: what turns a Kinect frame into the cluster estimateFinalPose is
// handed (D&L/src/rosinterface.cpp:212-250 -> ProcessingPcd::getPassThrough, ObjectSegmentationPlane::getSegmentedObjectsOnPlane).
//   ope_pass_through        pcl::PassThrough on z, y, x in sequence (D&L/src/processingpcd.cpp:8-41): one ordered compaction
//   ope_euclidean_clusters  pcl::EuclideanClusterExtraction (D&L/src/objectsegmentationplane.cpp:74-90: tolerance 0.05,
//                           300 <= size <= 1e5): connected components of the "closer than the tolerance" graph by a lock-free
//                           union-find over the cloud's Morton grid; the host only renumbers the components by size
//   ope_plane_ransac        pcl::SACSegmentation, SACMODEL_PLANE / SAC_RANSAC, threshold 0.01, refined coefficients
//                           (D&L/src/objectsegmentationplane.cpp:36-55): all candidate planes of the replayed sample table scored
//                           in one launch (a block per hypothesis), the adaptive stopping rule replayed on the host
//   ope_prism_select        the polygonal-prism crop over the plane's padded bounding rectangle (:174-218)
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <vector>

#include "ope_host.cuh"
#include "ope_device.cuh"

namespace ope {

static constexpr int kSegThreads = 256;

// ---- pass-through: keep[i] = finite && lo <= field <= hi for all three fields; order-preserving compaction ----
__global__ void pass_flag_kernel(const float4* __restrict__ pts, int n, float x0, float x1, float y0, float y1, float z0, float z1,
                                 int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = __ldg(pts + i);
  // three PassThrough filters in sequence (z, y, x): a point survives iff it is finite and inside every closed interval
  const bool keep = finite3(p.x, p.y, p.z) && !(p.z < z0 || p.z > z1) && !(p.y < y0 || p.y > y1) && !(p.x < x0 || p.x > x1);
  flags[i] = keep ? 1 : 0;
}
__global__ void pass_compact_kernel(const float4* __restrict__ pts, int n, const int* __restrict__ pos, float4* __restrict__ out,
                                    int* __restrict__ out_idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (pos[i + 1] > pos[i]) { out[pos[i]] = __ldg(pts + i); if (out_idx) out_idx[pos[i]] = i; }
}

// ---- Euclidean clustering: union-find, the smaller index is the root ----
__device__ __forceinline__ int uf_find(int* parent, int i) {
  int p = parent[i];
  while (p != i) {
    const int gp = parent[p];
    if (gp != p) parent[i] = gp;   // path halving (benign race: any ancestor is a valid parent)
    i = p; p = gp;
  }
  return i;
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  for (;;) {
    a = uf_find(parent, a); b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }   // hook the larger root under the smaller
    const int old = atomicCAS(parent + a, a, b);
    if (old == a) return;
  }
}
// one warp per point: every indexed point closer than the tolerance (squared distance < tol^2, the radiusSearch of the
// reference) and with a smaller index is united with it
__global__ void __launch_bounds__(kSegThreads) cluster_union_kernel(GridView g, const float4* __restrict__ pts, int n, float r2,
                                                                    int* __restrict__ parent) {
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = wid; i < n; i += n_warps) {
    const float4 q = __ldg(pts + i);
    if (!finite3(q.x, q.y, q.z)) continue;
    grid_radius_ranges(g, q.x, q.y, q.z, r2, 64, [&](int b, int e) {
      for (int s = b + lane; s < e; s += 32) {
        const float4 c = __ldg(g.pts + s);
        const int j = __float_as_int(c.w);
        if (j < i && dist2(q.x, q.y, q.z, c.x, c.y, c.z) < r2) uf_union(parent, i, j);
      }
    });
  }
}
__global__ void cluster_init_kernel(const float4* __restrict__ pts, int n, int* __restrict__ parent) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) parent[i] = i;
}
__global__ void cluster_flatten_kernel(const float4* __restrict__ pts, int n, int* __restrict__ parent, int* __restrict__ root) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 q = __ldg(pts + i);
  root[i] = finite3(q.x, q.y, q.z) ? uf_find(parent, i) : -1;
}

}  // namespace ope

using namespace ope;

extern "C" {

// limits: x_min, x_max, y_min, y_max, z_min, z_max (ProcessingPcd::getPassThrough's argument order). out_idx (n entries) may be NULL.
int ope_pass_through(ope_ctx* ctx, const ope_cloud* cloud, const float limits[6], ope_cloud** out, int32_t* out_idx, size_t* out_n) {
  OPE_ENTER(ctx);
  if (!ctx || !cloud || !limits || !out) return OPE_ERR_INVALID;
  *out = nullptr;
  if (out_n) *out_n = 0;
  const size_t n = cloud->n;
  if (n == 0) return cloud_alloc(ctx, 0, false, out);
  if (n > 0x7ffffffeull) return fail(ctx, OPE_ERR_INVALID, "cloud too large");
  Scratch<int> flags(ctx), d_idx(ctx);
  OPE_TRY(flags.alloc(n + 1));
  OPE_CUDA_TRY(ctx, cudaMemsetAsync(flags.p + n, 0, sizeof(int), ctx->stream));
  pass_flag_kernel<<<div_up(n, kSegThreads), kSegThreads, 0, ctx->stream>>>(cloud->pts, (int)n, limits[0], limits[1], limits[2], limits[3],
                                                                           limits[4], limits[5], flags.p);
  OPE_TRY(check_launch(ctx, "pass_flag_kernel"));
  OPE_TRY(exclusive_scan_i32(ctx, flags.p, n + 1));
  void* h;
  OPE_TRY(read_back(ctx, flags.p + n, sizeof(int), &h));
  const size_t m = (size_t) * (const int*)h;
  ope_cloud* o = nullptr;
  OPE_TRY(cloud_alloc(ctx, m, false, &o));
  if (out_idx) { int rc = d_idx.alloc(m); if (rc != OPE_OK) { ope_cloud_free(ctx, o); return rc; } }
  pass_compact_kernel<<<div_up(n, kSegThreads), kSegThreads, 0, ctx->stream>>>(cloud->pts, (int)n, flags.p, o->pts, out_idx ? d_idx.p : nullptr);
  int rc = check_launch(ctx, "pass_compact_kernel");
  if (rc == OPE_OK && out_idx && m > 0) {
    cudaError_t e = cudaMemcpyAsync(out_idx, d_idx.p, m * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = stream_sync(ctx);
    if (e != cudaSuccess) rc = fail(ctx, OPE_ERR_CUDA, "index download failed: %s", cudaGetErrorString(e));
  }
  if (rc != OPE_OK) { ope_cloud_free(ctx, o); return rc; }
  if (out_n) *out_n = m;
  *out = o;
  return OPE_OK;
}

// labels: n entries — cluster number (0 = largest; equal sizes by the smaller first index) or -1 (not finite / component outside
// [min_size, max_size]). Returns the number of clusters kept in *n_clusters.
int ope_euclidean_clusters(ope_ctx* ctx, const ope_cloud* cloud, float tolerance, int min_size, int max_size, int32_t* labels, int* n_clusters) {
  OPE_ENTER(ctx);
  if (!ctx || !cloud || !labels || !n_clusters || !(tolerance > 0)) return OPE_ERR_INVALID;
  *n_clusters = 0;
  const size_t n = cloud->n;
  if (n == 0) return OPE_OK;
  if (n > 0x7ffffffeull) return fail(ctx, OPE_ERR_INVALID, "cloud too large");
  OPE_TRY(cloud_bbox(ctx, const_cast<ope_cloud*>(cloud)));
  std::vector<int> root(n, -1);
  if (cloud->n_finite > 0) {
    GridView g;
    OPE_TRY(cloud_grid(ctx, cloud, tolerance, &g));
    Scratch<int> parent(ctx), d_root(ctx);
    OPE_TRY(parent.alloc(n)); OPE_TRY(d_root.alloc(n));
    cluster_init_kernel<<<div_up(n, kSegThreads), kSegThreads, 0, ctx->stream>>>(cloud->pts, (int)n, parent.p);
    OPE_TRY(check_launch(ctx, "cluster_init_kernel"));
    const unsigned blocks = (unsigned)std::min<size_t>(div_up(n * 32, kSegThreads), (size_t)ctx->sm_count * 16);
    cluster_union_kernel<<<blocks, kSegThreads, 0, ctx->stream>>>(g, cloud->pts, (int)n, tolerance * tolerance, parent.p);
    OPE_TRY(check_launch(ctx, "cluster_union_kernel"));
    cluster_flatten_kernel<<<div_up(n, kSegThreads), kSegThreads, 0, ctx->stream>>>(cloud->pts, (int)n, parent.p, d_root.p);
    OPE_TRY(check_launch(ctx, "cluster_flatten_kernel"));
    OPE_CUDA_TRY(ctx, cudaMemcpyAsync(root.data(), d_root.p, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    OPE_CUDA_TRY(ctx, stream_sync(ctx));
  }
  // bookkeeping on the host: component sizes, the size window, the reference's ordering (largest first)
  std::vector<int> size(n, 0);
  for (size_t i = 0; i < n; ++i) if (root[i] >= 0) size[(size_t)root[i]]++;
  std::vector<int> roots;
  for (size_t i = 0; i < n; ++i) if (size[i] >= min_size && size[i] <= max_size && size[i] > 0) roots.push_back((int)i);   // a root is its component's smallest index
  std::sort(roots.begin(), roots.end(), [&](int a, int b) { return size[(size_t)a] != size[(size_t)b] ? size[(size_t)a] > size[(size_t)b] : a < b; });
  std::vector<int> number(n, -1);
  for (size_t c = 0; c < roots.size(); ++c) number[(size_t)roots[c]] = (int)c;
  for (size_t i = 0; i < n; ++i) labels[i] = root[i] >= 0 ? number[(size_t)root[i]] : -1;
  *n_clusters = (int)roots.size();
  return OPE_OK;
}

}  // extern "C"
