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
#include <climits>
#include <numeric>
#include <random>
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


// ---- plane RANSAC (pcl::SACSegmentation, SACMODEL_PLANE) -------------------------------------------------------------------
// Eigen's 4-float reduction order of the whole path: (p0 + p1) + (p2 + p3)
OPE_HD float redux4(float p0, float p1, float p2, float p3) { return (p0 + p1) + (p2 + p3); }
OPE_HD float plane_distance(const float c[4], float x, float y, float z) { return redux4(c[0] * x, c[1] * y, c[2] * z, c[3] * 1.0f); }

// countWithinDistance for every candidate plane at once: block h counts the points with |c_h . p| < threshold
__global__ void __launch_bounds__(kSegThreads) plane_count_kernel(const float4* __restrict__ pts, int n, const float* __restrict__ coeffs,
                                                                  double threshold, int* __restrict__ counts) {
  __shared__ int total;
  const float c[4] = {coeffs[4 * blockIdx.x], coeffs[4 * blockIdx.x + 1], coeffs[4 * blockIdx.x + 2], coeffs[4 * blockIdx.x + 3]};
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  int cnt = 0;
  for (int i = threadIdx.x; i < n; i += kSegThreads) {
    const float4 p = __ldg(pts + i);
    cnt += (double)fabsf(plane_distance(c, p.x, p.y, p.z)) < threshold ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&total, cnt);
  __syncthreads();
  if (threadIdx.x == 0) counts[blockIdx.x] = total;
}
// selectWithinDistance: flags for the ordered compaction
__global__ void plane_flag_kernel(const float4* __restrict__ pts, int n, const float* __restrict__ coeff, double threshold, int invert,
                                  int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float c[4] = {coeff[0], coeff[1], coeff[2], coeff[3]};
  const float4 p = __ldg(pts + i);
  const bool in = (double)fabsf(plane_distance(c, p.x, p.y, p.z)) < threshold;
  flags[i] = (in != (invert != 0)) ? 1 : 0;
}
__global__ void index_compact_kernel(int n, const int* __restrict__ pos, int* __restrict__ out_idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && pos[i + 1] > pos[i]) out_idx[pos[i]] = i;
}
__global__ void gather_points_kernel(const float4* __restrict__ pts, const int* __restrict__ idx, int m, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) out[i] = __ldg(pts + idx[i]);
}
// optimizeModelCoefficients: computeMeanAndCovarianceMatrix accumulates nine FLOAT sums over the inliers IN INDEX ORDER, so the
// sums are order-dependent and each is one serial chain of additions. The chain is all that stays serial: warps 1.. gather the
// next chunk of inliers and write its nine term arrays (x*x, x*y, ... z) to shared memory while lanes 0..8 of warp 0 — one per
// accumulator, the same instruction stream for all nine — add up the current chunk. Thread 0 finishes with eigen33 and
// d = -n . centroid. One block.
static constexpr int kRefitChunk = 512;
__global__ void __launch_bounds__(kSegThreads) plane_refit_kernel(const float4* __restrict__ pts, const int* __restrict__ inliers, int m,
                                                                  const float* __restrict__ in, float* __restrict__ out) {
  __shared__ float terms[2][9][kRefitChunk + 1];   // + 1: the nine lanes of a step read nine different banks
  __shared__ float accu[9];
  const int tid = threadIdx.x;
  auto fill = [&](int chunk) {   // by the threads of warps 1..: terms of chunk `chunk` into its buffer
    const int base = chunk * kRefitChunk;
    const int cnt = min(kRefitChunk, m - base);
    float (*t)[kRefitChunk + 1] = terms[chunk & 1];
    for (int j = tid - 32; j < cnt; j += kSegThreads - 32) {
      const float4 p = __ldg(pts + inliers[base + j]);
      t[0][j] = p.x * p.x; t[1][j] = p.x * p.y; t[2][j] = p.x * p.z;
      t[3][j] = p.y * p.y; t[4][j] = p.y * p.z; t[5][j] = p.z * p.z;
      t[6][j] = p.x; t[7][j] = p.y; t[8][j] = p.z;
    }
  };
  const int n_chunks = (m + kRefitChunk - 1) / kRefitChunk;
  float acc = 0.0f;
  if (tid >= 32 && n_chunks > 0) fill(0);
  __syncthreads();
  for (int c = 0; c < n_chunks; ++c) {
    if (tid >= 32) {
      if (c + 1 < n_chunks) fill(c + 1);
    } else if (tid < 9) {
      const int cnt = min(kRefitChunk, m - c * kRefitChunk);
      const float* t = terms[c & 1][tid];
#pragma unroll 8
      for (int j = 0; j < cnt; ++j) acc += t[j];
    }
    __syncthreads();
  }
  if (tid < 9) accu[tid] = acc / (float)m;
  __syncthreads();
  if (tid == 0) {
    if (m < 4) { for (int i = 0; i < 4; ++i) out[i] = in[i]; return; }
    float cov[9];
    cov[0] = accu[0] - accu[6] * accu[6]; cov[1] = accu[1] - accu[6] * accu[7]; cov[2] = accu[2] - accu[6] * accu[8];
    cov[4] = accu[3] - accu[7] * accu[7]; cov[5] = accu[4] - accu[7] * accu[8]; cov[8] = accu[5] - accu[8] * accu[8];
    cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
    float ev, nv[3];
    eigen33(cov, ev, nv);
    out[0] = nv[0]; out[1] = nv[1]; out[2] = nv[2];
    out[3] = -1.0f * redux4(nv[0] * accu[6], nv[1] * accu[7], nv[2] * accu[8], 0.0f * 1.0f);
  }
}
// SampleConsensusModelPlane::projectPoints
OPE_HD void plane_project(const float c[4], float x, float y, float z, float q[3]) {
  float mc[4] = {c[0], c[1], c[2], 0.0f};
  const float nrm = sqrtf(redux4(mc[0] * mc[0], mc[1] * mc[1], mc[2] * mc[2], mc[3] * mc[3]));
  for (int i = 0; i < 4; ++i) mc[i] /= nrm;
  const float dist = redux4(mc[0] * x, mc[1] * y, mc[2] * z, c[3] * 1.0f);
  q[0] = x - mc[0] * dist; q[1] = y - mc[1] * dist; q[2] = z - mc[2] * dist;
}
// bounding box (x, y) of the projected inliers: out = {min x, min y, max x, max y}, float bits compared through atomics on ints
__device__ __forceinline__ void atomic_min_f(float* a, float v) { v += 0.0f; if (v >= 0) atomicMin((int*)a, __float_as_int(v)); else atomicMax((unsigned*)a, __float_as_uint(v)); }
__device__ __forceinline__ void atomic_max_f(float* a, float v) { v += 0.0f; if (v >= 0) atomicMax((int*)a, __float_as_int(v)); else atomicMin((unsigned*)a, __float_as_uint(v)); }
__global__ void projected_bbox_kernel(const float4* __restrict__ pts, const int* __restrict__ inliers, int m, const float* __restrict__ coeff,
                                      float* __restrict__ out) {
  const float c[4] = {coeff[0], coeff[1], coeff[2], coeff[3]};
  float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    const float4 p = __ldg(pts + inliers[i]);
    float q[3];
    plane_project(c, p.x, p.y, p.z, q);
    mnx = fminf(mnx, q[0]); mny = fminf(mny, q[1]); mxx = fmaxf(mxx, q[0]); mxy = fmaxf(mxy, q[1]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
    mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
  }
  if ((threadIdx.x & 31) == 0 && mnx <= mxx) { atomic_min_f(out, mnx); atomic_min_f(out + 1, mny); atomic_max_f(out + 2, mxx); atomic_max_f(out + 3, mxy); }
}
// pcl::ExtractPolygonalPrismData over the 4-corner rectangle: signed height in [0, FLT_MAX] and the projected point inside the polygon
struct PrismArgs { float mc[4]; float poly[4][2]; int k1, k2; };
__global__ void prism_flag_kernel(const float4* __restrict__ pts, int n, PrismArgs a, int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = __ldg(pts + i);
  const double distance = (double)(a.mc[0] * p.x + a.mc[1] * p.y + a.mc[2] * p.z + a.mc[3]);
  bool keep = !(distance < 0.0 || distance > (double)FLT_MAX);
  if (keep) {
    float q[3];
    plane_project(a.mc, p.x, p.y, p.z, q);
    const double px = q[a.k1], py = q[a.k2];
    bool in = false;
    double xold = a.poly[3][0], yold = a.poly[3][1];
    for (int e = 0; e < 4; ++e) {
      const double xnew = a.poly[e][0], ynew = a.poly[e][1];
      double x1, x2, y1, y2;
      if (xnew > xold) { x1 = xold; x2 = xnew; y1 = yold; y2 = ynew; } else { x1 = xnew; x2 = xold; y1 = ynew; y2 = yold; }
      if ((xnew < px) == (px <= xold) && (py - y1) * (x2 - x1) < (y2 - y1) * (px - x1)) in = !in;
      xold = xnew; yold = ynew;
    }
    keep = in;
  }
  flags[i] = keep ? 1 : 0;
}

}  // namespace ope

using namespace ope;

namespace {

// ordered compaction of the points whose flag is set: device index list (ascending) + count
int compact_flags(ope_ctx* ctx, int* d_flags /* n + 1 */, size_t n, int** d_idx, size_t* m) {
  *d_idx = nullptr; *m = 0;
  OPE_CUDA_TRY(ctx, cudaMemsetAsync(d_flags + n, 0, sizeof(int), ctx->stream));
  OPE_TRY(exclusive_scan_i32(ctx, d_flags, n + 1));
  void* h;
  OPE_TRY(read_back(ctx, d_flags + n, sizeof(int), &h));
  *m = (size_t) * (const int*)h;
  OPE_TRY(dalloc(ctx, d_idx, std::max<size_t>(*m, 1)));
  index_compact_kernel<<<div_up(n, kSegThreads), kSegThreads, 0, ctx->stream>>>((int)n, d_flags, *d_idx);
  return check_launch(ctx, "index_compact_kernel");
}

struct PlaneSampler {   // SampleConsensusModel::drawIndexSample: boost::mt19937(12345) through uniform_int<>(0, INT_MAX) = mt() >> 1
  std::mt19937 rng{12345u};
  std::vector<int> shuffled;
  explicit PlaneSampler(size_t n) : shuffled(n) { for (size_t i = 0; i < n; ++i) shuffled[i] = (int)i; }
  void draw(int out[3]) {
    const size_t index_size = shuffled.size();
    for (unsigned i = 0; i < 3; ++i) std::swap(shuffled[i], shuffled[i + ((unsigned)(rng() >> 1) % (index_size - i))]);
    out[0] = shuffled[0]; out[1] = shuffled[1]; out[2] = shuffled[2];
  }
};
bool sample_good(const float4& p0, const float4& p1, const float4& p2) {
  const float d0 = (p1.x - p0.x) / (p2.x - p0.x), d1 = (p1.y - p0.y) / (p2.y - p0.y), d2 = (p1.z - p0.z) / (p2.z - p0.z);
  return (d0 != d1) || (d2 != d1);
}
bool plane_from_sample(const float4& p0, const float4& p1, const float4& p2, float c[4]) {
  const float a[3] = {p1.x - p0.x, p1.y - p0.y, p1.z - p0.z}, b[3] = {p2.x - p0.x, p2.y - p0.y, p2.z - p0.z};
  const float d0 = a[0] / b[0], d1 = a[1] / b[1], d2 = a[2] / b[2];
  if ((d0 == d1) && (d2 == d1)) return false;
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
  c[3] = 0.0f;
  const float nrm = std::sqrt(redux4(c[0] * c[0], c[1] * c[1], c[2] * c[2], c[3] * c[3]));
  for (int i = 0; i < 4; ++i) c[i] /= nrm;
  c[3] = -1.0f * redux4(c[0] * p0.x, c[1] * p0.y, c[2] * p0.z, c[3] * 1.0f);
  return true;
}

// pcl::SACSegmentation::segment (SACMODEL_PLANE, SAC_RANSAC, optimized coefficients) on n device points without NaN.
// The samples are drawn on the host exactly as PCL draws them; their points are fetched in one gather; every candidate plane is
// scored in one launch; the adaptive stopping rule is replayed over the counts. found = 0: no model.
int plane_segment_device(ope_ctx* ctx, const float4* pts, size_t n, const ope_segment_params& P, float coeff[4], int** d_inliers, size_t* n_inliers,
                         int* iterations, int* found) {
  *d_inliers = nullptr; *n_inliers = 0; *iterations = 0; *found = 0;
  if (n < 3) return OPE_OK;
  const int M = P.max_iterations + 1;   // the loop runs while iterations <= max_iterations
  PlaneSampler sampler(n);
  std::vector<int> h_idx((size_t)3 * M);
  for (int h = 0; h < M; ++h) sampler.draw(&h_idx[(size_t)3 * h]);
  Scratch<int> d_idx(ctx), d_counts(ctx), flags(ctx);
  Scratch<float4> d_spts(ctx);
  Scratch<float> d_coeffs(ctx);
  OPE_TRY(d_idx.alloc((size_t)3 * M)); OPE_TRY(d_spts.alloc((size_t)3 * M)); OPE_TRY(d_coeffs.alloc((size_t)4 * (M + 2))); OPE_TRY(d_counts.alloc(M));
  OPE_CUDA_TRY(ctx, cudaMemcpyAsync(d_idx.p, h_idx.data(), (size_t)3 * M * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  gather_points_kernel<<<div_up((size_t)3 * M, kSegThreads), kSegThreads, 0, ctx->stream>>>(pts, d_idx.p, 3 * M, d_spts.p);
  OPE_TRY(check_launch(ctx, "gather_points_kernel"));
  std::vector<float4> sp((size_t)3 * M);
  OPE_CUDA_TRY(ctx, cudaMemcpyAsync(sp.data(), d_spts.p, sp.size() * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
  OPE_CUDA_TRY(ctx, stream_sync(ctx));
  std::vector<float> h_coeffs((size_t)4 * M);
  for (int h = 0; h < M; ++h) {
    const float4 &p0 = sp[(size_t)3 * h], &p1 = sp[(size_t)3 * h + 1], &p2 = sp[(size_t)3 * h + 2];
    // a degenerate draw makes PCL draw again / skip the iteration, which shifts the whole random sequence: practically never
    // (three exactly collinear float points), and then this one-gather scheme does not apply
    if (!sample_good(p0, p1, p2) || !plane_from_sample(p0, p1, p2, &h_coeffs[(size_t)4 * h]))
      return fail(ctx, OPE_ERR_UNSUPPORTED, "plane RANSAC drew a degenerate sample (hypothesis %d)", h);
  }
  OPE_CUDA_TRY(ctx, cudaMemcpyAsync(d_coeffs.p, h_coeffs.data(), h_coeffs.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  plane_count_kernel<<<M, kSegThreads, 0, ctx->stream>>>(pts, (int)n, d_coeffs.p, P.distance_threshold, d_counts.p);
  OPE_TRY(check_launch(ctx, "plane_count_kernel"));
  std::vector<int> counts((size_t)M);
  OPE_CUDA_TRY(ctx, cudaMemcpyAsync(counts.data(), d_counts.p, (size_t)M * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  OPE_CUDA_TRY(ctx, stream_sync(ctx));
  // RandomSampleConsensus::computeModel [UPSTREAM ransac.hpp], replayed
  int it = 0, n_best = -INT_MAX, best = -1;
  double k = 1.0;
  const double log_probability = std::log(1.0 - P.probability), one_over_indices = 1.0 / (double)n;
  while (it < k && it < M) {
    if (counts[(size_t)it] > n_best) {
      n_best = counts[(size_t)it]; best = it;
      const double w = (double)n_best * one_over_indices;
      double p_no_outliers = 1.0 - std::pow(w, 3.0);
      p_no_outliers = std::max(std::numeric_limits<double>::epsilon(), p_no_outliers);
      p_no_outliers = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no_outliers);
      k = log_probability / std::log(p_no_outliers);
    }
    ++it;
    if (it > P.max_iterations) break;
  }
  *iterations = it;
  if (best < 0) return OPE_OK;
  // inliers of the best model, least-squares refit over them, inliers of the refined model
  float* d_best = d_coeffs.p + (size_t)4 * best;
  float* d_refined = d_coeffs.p + (size_t)4 * M;
  OPE_TRY(flags.alloc(n + 1));
  plane_flag_kernel<<<div_up(n, kSegThreads), kSegThreads, 0, ctx->stream>>>(pts, (int)n, d_best, P.distance_threshold, 0, flags.p);
  OPE_TRY(check_launch(ctx, "plane_flag_kernel"));
  int* d_in = nullptr;
  size_t m = 0;
  OPE_TRY(compact_flags(ctx, flags.p, n, &d_in, &m));
  plane_refit_kernel<<<1, kSegThreads, 0, ctx->stream>>>(pts, d_in, (int)m, d_best, d_refined);
  int rc = check_launch(ctx, "plane_refit_kernel");
  dfree(ctx, d_in);
  OPE_TRY(rc);
  plane_flag_kernel<<<div_up(n, kSegThreads), kSegThreads, 0, ctx->stream>>>(pts, (int)n, d_refined, P.distance_threshold, 0, flags.p);
  OPE_TRY(check_launch(ctx, "plane_flag_kernel"));
  OPE_TRY(compact_flags(ctx, flags.p, n, d_inliers, n_inliers));
  void* h;
  rc = read_back(ctx, d_refined, 4 * sizeof(float), &h);
  if (rc != OPE_OK) { dfree(ctx, *d_inliers); *d_inliers = nullptr; return rc; }
  std::memcpy(coeff, h, 4 * sizeof(float));
  *found = 1;
  return OPE_OK;
}

}  // namespace

extern "C" {

void ope_segment_params_default(ope_segment_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->distance_threshold = 0.01; p->max_iterations = 50; p->probability = 0.99; p->hull_margin = 0.1;
  p->cluster_tolerance = 0.05f; p->min_cluster_size = 300; p->max_cluster_size = 100000;
}

// pcl::SACSegmentation on a device cloud without NaN points: refined plane coefficients, the inlier indices (ascending; out_idx has
// room for ope_cloud_size entries, may be NULL), the RANSAC iteration count. *found = 0 when no plane exists (fewer than 3 points).
int ope_plane_ransac(ope_ctx* ctx, const ope_cloud* cloud, const ope_segment_params* prm, float coeff[4], int32_t* out_idx, size_t* out_n,
                     int32_t* iterations, int32_t* found) {
  OPE_ENTER(ctx);
  if (!ctx || !cloud || !coeff || !found) return OPE_ERR_INVALID;
  ope_segment_params P;
  if (prm) P = *prm; else ope_segment_params_default(&P);
  int* d_in = nullptr;
  size_t m = 0;
  int it = 0, ok = 0;
  OPE_TRY(plane_segment_device(ctx, cloud->pts, cloud->n, P, coeff, &d_in, &m, &it, &ok));
  int rc = OPE_OK;
  if (ok && out_idx && m > 0) {
    cudaError_t e = cudaMemcpyAsync(out_idx, d_in, m * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = stream_sync(ctx);
    if (e != cudaSuccess) rc = fail(ctx, OPE_ERR_CUDA, "index download failed: %s", cudaGetErrorString(e));
  }
  dfree(ctx, d_in);
  if (out_n) *out_n = ok ? m : 0;
  if (iterations) *iterations = it;
  *found = ok;
  return rc;
}

// ObjectSegmentationPlane::getSegmentedObjectsOnPlane (D&L/src/objectsegmentationplane.cpp:122-282) on a pass-through-filtered
// device cloud: plane, prism over the padded hull rectangle, second plane on the prism's points, Euclidean clusters of the rest.
// labels (n entries): OPE_SEG_OUTSIDE_PRISM / OPE_SEG_PLANE / OPE_SEG_NO_CLUSTER / cluster number (0 = largest).
// *n_clusters = -1 when no plane was found (the reference then passes the whole cloud on).
int ope_segment_objects_on_plane(ope_ctx* ctx, const ope_cloud* cloud, const ope_segment_params* prm, int32_t* labels, float plane1[4],
                                 float plane2[4], int32_t iters[2], int32_t* n_clusters) {
  OPE_ENTER(ctx);
  if (!ctx || !cloud || !labels || !n_clusters) return OPE_ERR_INVALID;
  ope_segment_params P;
  if (prm) P = *prm; else ope_segment_params_default(&P);
  const size_t n = cloud->n;
  *n_clusters = -1;
  for (size_t i = 0; i < n; ++i) labels[i] = OPE_SEG_OUTSIDE_PRISM;
  if (n > 0x7ffffffeull) return fail(ctx, OPE_ERR_INVALID, "cloud too large");
  float c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0};
  int it1 = 0, it2 = 0, ok = 0;
  int* d_in1 = nullptr;
  size_t m1 = 0;
  OPE_TRY(plane_segment_device(ctx, cloud->pts, n, P, c1, &d_in1, &m1, &it1, &ok));
  struct Free { ope_ctx* c; int*& p; ~Free() { dfree(c, p); } } f1{ctx, d_in1};
  if (iters) { iters[0] = it1; iters[1] = 0; }
  if (!ok) return OPE_OK;
  // bounding rectangle of the projected inliers (= of their convex hull), padded, z from the plane equation (:174-196)
  Scratch<float> d_c(ctx), d_box(ctx);
  OPE_TRY(d_c.alloc(4)); OPE_TRY(d_box.alloc(4));
  const float init[4] = {INFINITY, INFINITY, -INFINITY, -INFINITY};
  OPE_CUDA_TRY(ctx, cudaMemcpyAsync(d_c.p, c1, 16, cudaMemcpyHostToDevice, ctx->stream));
  OPE_CUDA_TRY(ctx, cudaMemcpyAsync(d_box.p, init, 16, cudaMemcpyHostToDevice, ctx->stream));
  OPE_CUDA_TRY(ctx, stream_sync(ctx));   // `init` / `c1` are stack memory
  projected_bbox_kernel<<<(unsigned)std::min<size_t>(div_up(std::max<size_t>(m1, 1), kSegThreads), (size_t)ctx->sm_count * 4), kSegThreads, 0, ctx->stream>>>(
      cloud->pts, d_in1, (int)m1, d_c.p, d_box.p);
  OPE_TRY(check_launch(ctx, "projected_bbox_kernel"));
  void* h;
  OPE_TRY(read_back(ctx, d_box.p, 16, &h));
  float box[4];
  std::memcpy(box, h, 16);
  const float vx[4] = {(float)(box[0] - P.hull_margin), (float)(box[0] - P.hull_margin), (float)(box[2] + P.hull_margin), (float)(box[2] + P.hull_margin)};
  const float vy[4] = {(float)(box[1] - P.hull_margin), (float)(box[3] + P.hull_margin), (float)(box[3] + P.hull_margin), (float)(box[1] - P.hull_margin)};
  float rect[4][3];
  for (int i = 0; i < 4; ++i) { rect[i][0] = vx[i]; rect[i][1] = vy[i]; rect[i][2] = -((c1[0] * vx[i]) + (c1[1] * vy[i]) + c1[3]) / c1[2]; }
  // ExtractPolygonalPrismData: plane of the four corners (mean + covariance + eigen33), flipped towards the viewpoint (0, 0, 0)
  PrismArgs pa;
  {
    float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
      const float* p = rect[i];
      accu[0] += p[0] * p[0]; accu[1] += p[0] * p[1]; accu[2] += p[0] * p[2];
      accu[3] += p[1] * p[1]; accu[4] += p[1] * p[2]; accu[5] += p[2] * p[2];
      accu[6] += p[0]; accu[7] += p[1]; accu[8] += p[2];
    }
    for (int i = 0; i < 9; ++i) accu[i] /= 4.0f;
    float cov[9];
    cov[0] = accu[0] - accu[6] * accu[6]; cov[1] = accu[1] - accu[6] * accu[7]; cov[2] = accu[2] - accu[6] * accu[8];
    cov[4] = accu[3] - accu[7] * accu[7]; cov[5] = accu[4] - accu[7] * accu[8]; cov[8] = accu[5] - accu[8] * accu[8];
    cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
    float ev, nv[3];
    eigen33(cov, ev, nv);
    float* mc = pa.mc;
    mc[0] = nv[0]; mc[1] = nv[1]; mc[2] = nv[2]; mc[3] = 0.0f;
    mc[3] = -1.0f * redux4(mc[0] * accu[6], mc[1] * accu[7], mc[2] * accu[8], mc[3] * 1.0f);
    const float vp[3] = {0.0f - rect[0][0], 0.0f - rect[0][1], 0.0f - rect[0][2]};
    const float cos_theta = redux4(vp[0] * mc[0], vp[1] * mc[1], vp[2] * mc[2], 0.0f * mc[3]);
    if (cos_theta < 0) {
      for (int i = 0; i < 4; ++i) mc[i] *= -1.0f;
      mc[3] = 0.0f;
      mc[3] = -1.0f * redux4(mc[0] * rect[0][0], mc[1] * rect[0][1], mc[2] * rect[0][2], mc[3] * 1.0f);
    }
    int k0 = (std::fabs(mc[0]) > std::fabs(mc[1])) ? 0 : 1;
    k0 = (std::fabs(mc[k0]) > std::fabs(mc[2])) ? k0 : 2;
    pa.k1 = (k0 + 1) % 3; pa.k2 = (k0 + 2) % 3;
    for (int i = 0; i < 4; ++i) { pa.poly[i][0] = rect[i][pa.k1]; pa.poly[i][1] = rect[i][pa.k2]; }
  }
  Scratch<int> flags(ctx);
  OPE_TRY(flags.alloc(n + 1));
  prism_flag_kernel<<<div_up(std::max<size_t>(n, 1), kSegThreads), kSegThreads, 0, ctx->stream>>>(cloud->pts, (int)n, pa, flags.p);
  OPE_TRY(check_launch(ctx, "prism_flag_kernel"));
  int* d_prism = nullptr;
  size_t mp = 0;
  OPE_TRY(compact_flags(ctx, flags.p, n, &d_prism, &mp));
  Free f2{ctx, d_prism};
  // cloudObjWithPlane (:206-213) and the second plane (:217)
  Scratch<float4> sub(ctx);
  OPE_TRY(sub.alloc(std::max<size_t>(mp, 1)));
  if (mp) { gather_points_kernel<<<div_up(mp, kSegThreads), kSegThreads, 0, ctx->stream>>>(cloud->pts, d_prism, (int)mp, sub.p); OPE_TRY(check_launch(ctx, "gather_points_kernel")); }
  int* d_in2 = nullptr;
  size_t m2 = 0;
  OPE_TRY(plane_segment_device(ctx, sub.p, mp, P, c2, &d_in2, &m2, &it2, &ok));
  Free f3{ctx, d_in2};
  if (iters) iters[1] = it2;
  if (!ok) return OPE_OK;
  // the rest: prism points that are not on the second plane (:229), clustered (:233-235)
  Scratch<float> d_c2(ctx);
  OPE_TRY(d_c2.alloc(4));
  OPE_CUDA_TRY(ctx, cudaMemcpyAsync(d_c2.p, c2, 16, cudaMemcpyHostToDevice, ctx->stream));
  OPE_CUDA_TRY(ctx, stream_sync(ctx));
  Scratch<int> flags2(ctx);
  OPE_TRY(flags2.alloc(mp + 1));
  plane_flag_kernel<<<div_up(std::max<size_t>(mp, 1), kSegThreads), kSegThreads, 0, ctx->stream>>>(sub.p, (int)mp, d_c2.p, P.distance_threshold, 1, flags2.p);
  OPE_TRY(check_launch(ctx, "plane_flag_kernel"));
  int* d_rest = nullptr;
  size_t mr = 0;
  OPE_TRY(compact_flags(ctx, flags2.p, mp, &d_rest, &mr));
  Free f4{ctx, d_rest};
  ope_cloud* rest = nullptr;
  OPE_TRY(cloud_alloc(ctx, mr, false, &rest));
  struct FreeCloud { ope_ctx* c; ope_cloud* p; ~FreeCloud() { if (p) ope_cloud_free(c, p); } } fc{ctx, rest};
  if (mr) { gather_points_kernel<<<div_up(mr, kSegThreads), kSegThreads, 0, ctx->stream>>>(sub.p, d_rest, (int)mr, rest->pts); OPE_TRY(check_launch(ctx, "gather_points_kernel")); }
  std::vector<int> h_prism(mp), h_in2(m2), h_rest(mr), cl(std::max<size_t>(mr, 1));
  if (mp) OPE_CUDA_TRY(ctx, cudaMemcpyAsync(h_prism.data(), d_prism, mp * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (m2) OPE_CUDA_TRY(ctx, cudaMemcpyAsync(h_in2.data(), d_in2, m2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (mr) OPE_CUDA_TRY(ctx, cudaMemcpyAsync(h_rest.data(), d_rest, mr * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  OPE_CUDA_TRY(ctx, stream_sync(ctx));
  int k = 0;
  OPE_TRY(ope_euclidean_clusters(ctx, rest, P.cluster_tolerance, P.min_cluster_size, P.max_cluster_size, cl.data(), &k));
  for (size_t j = 0; j < mp; ++j) labels[(size_t)h_prism[j]] = OPE_SEG_NO_CLUSTER;
  for (size_t j = 0; j < m2; ++j) labels[(size_t)h_prism[(size_t)h_in2[j]]] = OPE_SEG_PLANE;
  for (size_t j = 0; j < mr; ++j) labels[(size_t)h_prism[(size_t)h_rest[j]]] = cl[j] >= 0 ? cl[j] : OPE_SEG_NO_CLUSTER;
  if (plane1) std::memcpy(plane1, c1, 16);
  if (plane2) std::memcpy(plane2, c2, 16);
  *n_clusters = k;
  return OPE_OK;
}

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
