// batch.cu — ope_pose_batch with frame-spanning launches (BASELINE.json configs[4]): every stage of the first-frame
// path of PoseEstimator::estimateFinalPose (D&L/src/poseestimator.cpp:383-448) runs as ONE launch for a whole chunk of frames —
// blockIdx.y (or a ticket) selects the frame, sizes are read from device memory — instead of ~40 launches and ~25 host round
// trips per frame from a pool of worker threads. Per chunk the host waits a handful of times (to size the next stage's grid).
//
//   stage                                  launch(es)                                        per-frame data
//   target UniformSampling 1 cm + 8 mm     bsample_kernel<false>      grid (frames, 2)        cluster -> tp1, tp2 (voxel-key order)
//   normals k = 30 of tp1, tp2             normals_smem_batch_kernel  grid (<=8, 2 frames)
//   SPFH + FPFH of tp1                     spfh_/fpfh_batch_kernel    grid (n/8, frames)
//   5 nearest target features per model    feature_knn_batch_kernel (+ merge)                 the model's descriptors are shared
//   SAC-IA 400 x 5                         sacia_smem_batch_kernel    grid (400, frames)      + select: coarse pose per frame
//   model under the coarse pose, 8 mm      bsample_kernel<true>       grid (frames)           157 825 points, never materialised
//   normals k = 30 of sp2                  normals_smem_batch_kernel
//   removeNaNNormalsFromPointCloud x 2     compact_normals_batch_kernel
//   ICP-with-normals <= 100 iterations     icp_small_batch_kernel     ticketed blocks of all frames
//   getFitnessScore                        fitness_smem_batch_kernel  grid (frames)
//
// The kernels are the single-frame kernels' bodies (features.cu, registration.cu) behind a frame index, so a frame's result is
// what ope_pose_estimate_final returns for it — bit for bit, tests/test_gpu_parity.py. Frames outside the fast path's envelope
// (empty or tiny clusters, clouds beyond the shared-memory capacities, voxel tables that overflow) are finished by the per-frame
// path in pipeline.cu. The frame-invariant model side (1 cm sample, normals, FPFH; dense Umeyama of the model onto itself) is
// computed once per call.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <chrono>
#include <cstdio>
#include "ope_host.cuh"
#include "ope_device.cuh"

namespace ope {

static constexpr int kBCap = 4096;        // capacity of a sampled cloud (normals / ICP shared-memory paths)
static constexpr int kBFeatCap = 2048;    // capacity of the 1 cm target sample (FPFH shared-memory path)
static constexpr int kBTable = 8192;      // voxel hash slots per block (load factor <= 0.5 at kBCap voxels)
static constexpr int kBThreads = 1024;
static constexpr unsigned long long kEmpty = ~0ull;

struct BSampleArgs {
  const float4* pts;                // XFORM: the shared model; else the concatenated clusters
  const int* offsets;               // !XFORM: first point of frame f
  const int* counts;                // !XFORM: points of frame f
  int n_model;                      // XFORM
  const ope_reg_result* xforms;     // XFORM: frame f's coarse pose
  const int* active;                // frames to process (null: all)
  float inv_leaf[2];
  int cap[2];                       // output capacity per leaf
  float4* out[2];                   // leaf l, frame f: out[l] + f * kBCap
  int* out_counts[2];               // leaf l: counts[f]
  int* flags;                       // per frame: |= 1 when a sample overflowed its capacity / the voxel frame is too large
};

struct SampleSmem {
  unsigned long long keys[kBTable];
  unsigned long long vals[kBTable];
  unsigned long long list_key[kBCap];
  unsigned long long list_val[kBCap];
  float red[6][kBThreads / 32];
  int scan[kBThreads / 32];
  int n_finite, occupied, overflow, total;
  float bbox[6];
};

__device__ __forceinline__ unsigned hash_cell(long long c) {
  unsigned long long x = (unsigned long long)c * 0x9E3779B97F4A7C15ull;
  return (unsigned)(x >> 40);
}

// pcl::UniformSampling of one cloud per block: bounding box -> voxel frame -> per voxel the point with the smallest
// (||p4 - ijk4||^2, index) -> selected points in ascending voxel index. The voxel table is a shared-memory hash (linear probing,
// 64-bit keys); the result equals uniform_sample_device + gather_cloud (grid.cu) point for point.
template <bool XFORM>
__global__ void __launch_bounds__(kBThreads) bsample_kernel(BSampleArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SampleSmem* sm = reinterpret_cast<SampleSmem*>(smem_raw);
  const int f = blockIdx.x, leaf = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (a.active && !a.active[f]) return;
  const float4* pts = XFORM ? a.pts : a.pts + a.offsets[f];
  const int n = XFORM ? a.n_model : a.counts[f];
  Mat4 T;
  if (XFORM) for (int i = 0; i < 16; ++i) T.m[i] = a.xforms[f].T[i];
  auto load = [&](int i) {
    float4 p = __ldg(pts + i);
    if (XFORM) { float x, y, z; xform_point(T, p.x, p.y, p.z, x, y, z); p.x = x; p.y = y; p.z = z; }
    return p;
  };
  // ---- bounding box of the finite points ----
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  int fin = 0;
  for (int i = tid; i < n; i += kBThreads) {
    const float4 p = load(i);
    if (!finite3(p.x, p.y, p.z)) continue;
    ++fin;
    lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
    hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
  }
  for (int o = 16; o > 0; o >>= 1) {
    for (int d = 0; d < 3; ++d) { lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o)); hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o)); }
    fin += __shfl_xor_sync(0xffffffffu, fin, o);
  }
  if (tid == 0) { sm->n_finite = 0; sm->occupied = 0; sm->overflow = 0; }
  if (lane == 0) for (int d = 0; d < 3; ++d) { sm->red[d][warp] = lo[d]; sm->red[3 + d][warp] = hi[d]; }
  for (int s = tid; s < kBTable; s += kBThreads) { sm->keys[s] = kEmpty; sm->vals[s] = kEmpty; }
  __syncthreads();
  if (lane == 0) atomicAdd(&sm->n_finite, fin);
  if (tid < 6) {
    float v = sm->red[tid][0];
    for (int w = 1; w < kBThreads / 32; ++w) v = tid < 3 ? fminf(v, sm->red[tid][w]) : fmaxf(v, sm->red[tid][w]);
    sm->bbox[tid] = v;
  }
  __syncthreads();
  int* out_count = a.out_counts[leaf] + f;
  if (sm->n_finite == 0) { if (tid == 0) *out_count = 0; return; }
  // ---- the voxel frame of PCL (pcl_voxel_frame, grid.cu): min_b = floor(min * inv), dim = max_b - min_b + 1 ----
  const float inv = a.inv_leaf[leaf];
  int min_b[3];
  long long dim[3];
  bool too_large = false;
  for (int d = 0; d < 3; ++d) {
    const float l = sm->bbox[d] * inv, h = sm->bbox[3 + d] * inv;
    min_b[d] = (int)floorf(l);
    dim[d] = (long long)(int)floorf(h) - min_b[d] + 1;
    if (dim[d] > 0x7fffffffll) too_large = true;
  }
  if (!too_large && (double)dim[0] * (double)dim[1] * (double)dim[2] > (double)(1ll << 28)) too_large = true;
  if (too_large) { if (tid == 0) { *out_count = 0; atomicOr(a.flags + f, 1); } return; }
  // ---- per voxel: argmin over (||p4 - ijk4||^2, index) ----
  for (int i = tid; i < n; i += kBThreads) {
    const float4 p = load(i);
    if (!finite3(p.x, p.y, p.z)) continue;
    const int ix = (int)floorf(p.x * inv), iy = (int)floorf(p.y * inv), iz = (int)floorf(p.z * inv);
    const long long cell = ((long long)(iz - min_b[2]) * dim[1] + (iy - min_b[1])) * dim[0] + (ix - min_b[0]);
    const float da = p.x - (float)ix, db = p.y - (float)iy, dc = p.z - (float)iz;
    float d = da * da;
    d = d + db * db;
    d = d + dc * dc;
    d = d + 1.0f;  // 4th component of (p4 - ijk4): (1 - 0)^2
    const unsigned long long packed = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)i;
    unsigned slot = hash_cell(cell) & (kBTable - 1);
    bool placed = false;
    for (int probe = 0; probe < kBTable; ++probe) {
      unsigned long long k = sm->keys[slot];
      if (k == kEmpty) {
        if (sm->occupied >= kBTable / 2) break;                 // table full: give up (the frame is flagged below)
        k = atomicCAS(&sm->keys[slot], kEmpty, (unsigned long long)cell);
        if (k == kEmpty) { atomicAdd(&sm->occupied, 1); k = (unsigned long long)cell; }
      }
      if (k == (unsigned long long)cell) { placed = true; break; }
      slot = (slot + 1) & (kBTable - 1);
    }
    if (!placed) { sm->overflow = 1; continue; }
    if (sm->vals[slot] > packed) atomicMin(&sm->vals[slot], packed);
  }
  __syncthreads();
  const int cap = a.cap[leaf];
  if (sm->overflow || sm->occupied > cap) { if (tid == 0) { *out_count = 0; atomicOr(a.flags + f, 1); } return; }
  // ---- compact the occupied slots, sort by voxel index, emit the points ----
  constexpr int kPer = kBTable / kBThreads;
  int mine = 0;
  for (int u = 0; u < kPer; ++u) mine += sm->keys[tid * kPer + u] != kEmpty ? 1 : 0;
  int incl = mine;
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) sm->scan[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = sm->scan[lane];
    int inc2 = v;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc2, o); if (lane >= o) inc2 += t; }
    sm->scan[lane] = inc2 - v;
    if (lane == 31) sm->total = inc2;
  }
  __syncthreads();
  int pos = sm->scan[warp] + incl - mine;
  for (int u = 0; u < kPer; ++u) {
    const int s = tid * kPer + u;
    if (sm->keys[s] != kEmpty) { sm->list_key[pos] = sm->keys[s]; sm->list_val[pos] = sm->vals[s]; ++pos; }
  }
  const int m = sm->total;
  int m2 = 1;
  while (m2 < m) m2 <<= 1;
  for (int j = m + tid; j < m2; j += kBThreads) { sm->list_key[j] = kEmpty; sm->list_val[j] = kEmpty; }
  __syncthreads();
  for (int k2 = 2; k2 <= m2; k2 <<= 1)
    for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
      for (int i = tid; i < m2; i += kBThreads) {
        const int l = i ^ j2;
        if (l > i) {
          const bool up = (i & k2) == 0;
          const unsigned long long ki = sm->list_key[i], kl = sm->list_key[l];
          if ((ki > kl) == up) {
            sm->list_key[i] = kl; sm->list_key[l] = ki;
            const unsigned long long vi = sm->list_val[i];
            sm->list_val[i] = sm->list_val[l]; sm->list_val[l] = vi;
          }
        }
      }
      __syncthreads();
    }
  float4* out = a.out[leaf] + (size_t)f * kBCap;
  for (int j = tid; j < m; j += kBThreads) {
    const int idx = (int)(unsigned)(sm->list_val[j] & 0xffffffffull);
    float4 p = load(idx);
    out[j] = p;
  }
  if (tid == 0) *out_count = m;
}

// pcl::removeNaNNormalsFromPointCloud (D&L/src/poseestimator.cpp:215-216) for one cloud per block: points whose normal is finite,
// in order; the compacted points carry their new index in .w (the ICP kernel's work order).
__global__ void __launch_bounds__(kBThreads) compact_normals_batch_kernel(const float4* __restrict__ pts, const float4* __restrict__ nrm,
                                                                          const int* __restrict__ counts, const int* __restrict__ active,
                                                                          int frames, float4* __restrict__ opts, float4* __restrict__ onrm,
                                                                          int* __restrict__ ocounts) {
  __shared__ int scan[kBThreads / 32];
  __shared__ int total;
  const int c = blockIdx.x;   // cloud: [which][frame]
  const int f = c % frames;
  if (active && !active[f]) return;
  const int n = counts[c];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kPer = kBCap / kBThreads;
  const size_t base = (size_t)c * kBCap;
  bool keep[kPer];
  int mine = 0;
  for (int u = 0; u < kPer; ++u) {
    const int i = tid * kPer + u;
    keep[u] = false;
    if (i < n) { const float4 v = __ldg(nrm + base + i); keep[u] = finite3(v.x, v.y, v.z); }
    mine += keep[u] ? 1 : 0;
  }
  int incl = mine;
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) scan[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = scan[lane];
    int inc2 = v;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc2, o); if (lane >= o) inc2 += t; }
    scan[lane] = inc2 - v;
    if (lane == 31) total = inc2;
  }
  __syncthreads();
  int pos = scan[warp] + incl - mine;
  for (int u = 0; u < kPer; ++u) {
    const int i = tid * kPer + u;
    if (keep[u]) {
      float4 p = __ldg(pts + base + i);
      p.w = __int_as_float(pos);
      opts[base + pos] = p;
      onrm[base + pos] = __ldg(nrm + base + i);
      ++pos;
    }
  }
  if (tid == 0) ocounts[c] = total;
}

// Work order of the ICP kernel: both fine clouds re-ordered along the Morton curve of their own bounding box, so that the 32 points
// of a warp and the 8 points of a target group are spatial neighbours (the voxel-key order of the sampling is a z-major raster:
// thin rows). One block per cloud: bounding box, 30-bit Morton key | position, bitonic sort in shared memory, gather. The clouds
// stay self-consistent (points and normals move together, .w = new position), so nothing but the order of equal-distance ties
// and of the double-precision moment sums changes.
__device__ __forceinline__ unsigned spread10(unsigned v) {
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__global__ void __launch_bounds__(kBThreads) morton_sort_batch_kernel(const float4* __restrict__ pts, const float4* __restrict__ nrm,
                                                                     const int* __restrict__ counts, const int* __restrict__ active, int frames,
                                                                     float4* __restrict__ opts, float4* __restrict__ onrm) {
  __shared__ unsigned long long keys[kBCap];
  __shared__ float red[6][kBThreads / 32];
  __shared__ float box[6];
  const int c = blockIdx.x, f = c % frames;
  if (active && !active[f]) return;
  const int n = counts[c];
  if (n <= 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t base = (size_t)c * kBCap;
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = tid; i < n; i += kBThreads) {
    const float4 p = __ldg(pts + base + i);
    lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
    hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
  }
  for (int o = 16; o > 0; o >>= 1)
    for (int d = 0; d < 3; ++d) { lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o)); hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o)); }
  if (lane == 0) for (int d = 0; d < 3; ++d) { red[d][warp] = lo[d]; red[3 + d][warp] = hi[d]; }
  __syncthreads();
  if (tid < 6) {
    float v = red[tid][0];
    for (int w = 1; w < kBThreads / 32; ++w) v = tid < 3 ? fminf(v, red[tid][w]) : fmaxf(v, red[tid][w]);
    box[tid] = v;
  }
  __syncthreads();
  const float ext = fmaxf(fmaxf(box[3] - box[0], box[4] - box[1]), fmaxf(box[5] - box[2], 1e-9f));
  const float scale = 1023.0f / ext;
  int m2 = 1;
  while (m2 < n) m2 <<= 1;
  for (int i = tid; i < m2; i += kBThreads) {
    unsigned long long k = ~0ull;
    if (i < n) {
      const float4 p = __ldg(pts + base + i);
      const unsigned x = (unsigned)fminf(fmaxf((p.x - box[0]) * scale, 0.0f), 1023.0f), y = (unsigned)fminf(fmaxf((p.y - box[1]) * scale, 0.0f), 1023.0f),
                     z = (unsigned)fminf(fmaxf((p.z - box[2]) * scale, 0.0f), 1023.0f);
      k = ((unsigned long long)(spread10(x) | (spread10(y) << 1) | (spread10(z) << 2)) << 32) | (unsigned)i;
    }
    keys[i] = k;
  }
  __syncthreads();
  for (int k2 = 2; k2 <= m2; k2 <<= 1)
    for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
      for (int i = tid; i < m2; i += kBThreads) {
        const int l = i ^ j2;
        if (l > i) {
          const bool up = (i & k2) == 0;
          const unsigned long long ki = keys[i], kl = keys[l];
          if ((ki > kl) == up) { keys[i] = kl; keys[l] = ki; }
        }
      }
      __syncthreads();
    }
  for (int j = tid; j < n; j += kBThreads) {
    const int src = (int)(unsigned)(keys[j] & 0xffffffffull);
    float4 p = __ldg(pts + base + src);
    p.w = __int_as_float(j);
    opts[base + j] = p;
    onrm[base + j] = __ldg(nrm + base + src);
  }
}

namespace {
// per-stage device time of the chunk (CUDA events on the context's stream), accumulated into ctx->batch_stage_ms
struct ChunkTimer {
  ope_ctx* ctx;
  cudaEvent_t ev[16];
  std::chrono::steady_clock::time_point host[16];
  int n = 0;
  bool on, trace;
  explicit ChunkTimer(ope_ctx* c) : ctx(c), on(c->batch_timing), trace(c->batch_timing && std::getenv("OPE_BATCH_TRACE") != nullptr) {
    if (on) for (auto& e : ev) cudaEventCreate(&e);
    mark();
  }
  void mark() { if (on && n < 16) { host[n] = std::chrono::steady_clock::now(); cudaEventRecord(ev[n++], ctx->stream); } }
  ~ChunkTimer() {
    if (!on) return;
    cudaEventSynchronize(ev[n - 1]);
    for (int i = 1; i < n; ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      ctx->batch_stage_ms[i - 1] += ms;
      // OPE_BATCH_TRACE: the host's wall clock between the same marks (the device time includes what the device waited for the host)
      if (trace) fprintf(stderr, "[ope batch] stage %2d: device %8.3f ms, host %8.3f ms\n", i - 1, ms,
                         std::chrono::duration<double, std::milli>(host[i] - host[i - 1]).count());
    }
    for (auto& e : ev) cudaEventDestroy(e);
  }
};
struct DevGuardF { ope_ctx* ctx; float* p = nullptr; explicit DevGuardF(ope_ctx* c) : ctx(c) {} ~DevGuardF() { dfree(ctx, p); } };
template <typename T>
int download(ope_ctx* ctx, const T* d, size_t n, std::vector<T>& h) {
  h.resize(n);
  if (n == 0) return OPE_OK;
  OPE_CUDA_TRY(ctx, cudaMemcpyAsync(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
  OPE_CUDA_TRY(ctx, stream_sync(ctx));
  return OPE_OK;
}
}  // namespace

// One chunk of frames through the frame-spanning launches. Returns in `done[i]` whether frame i was finished here (else the
// caller runs it through the per-frame path). d_model: the full-resolution model; sp / d_fs: its coarse sample (with normals)
// and descriptors; rigid: the dense Umeyama of the model onto itself (what every first frame computes, :425-436).
int pose_batch_chunk(ope_ctx* ctx, const ope_pose_params& P, const ope_cloud* d_model, const ope_cloud* sp, const float* d_fs, const Mat4& rigid,
                     const ope_frame_input* frames, size_t n_frames, const ope_rng_table* tables, ope_pose_result* results, char* done,
                     const std::atomic<size_t>* tables_ready, size_t tables_needed) {
  const int B = (int)n_frames;
  const int H = P.sacia.max_iterations, S = P.sacia.nr_samples, K = P.sacia.k_correspondences;
  for (int i = 0; i < B; ++i) done[i] = 0;
  if (B == 0) return OPE_OK;
  if (!icp_small_batch_applicable(P.icp, 1, 1) || P.icp.n_rejectors < 0 || (int)sp->n < S || sp->n == 0) return OPE_OK;   // not this path's configuration
  // ---- stage the clusters: one concatenated float4 array ----
  std::vector<int> h_off((size_t)B), h_cnt((size_t)B), h_active((size_t)B, 1);
  size_t total = 0;
  for (int i = 0; i < B; ++i) {
    const ope_frame_input& in = frames[i];
    const size_t n = in.points ? in.n : (in.cloud ? ((const ope_cloud*)in.cloud)->n : 0);
    if (n == 0 || n > 0x3fffffff) h_active[(size_t)i] = 0;
    h_off[(size_t)i] = (int)total; h_cnt[(size_t)i] = h_active[(size_t)i] ? (int)n : 0;
    total += (size_t)h_cnt[(size_t)i];
  }
  if (total == 0 || total > 0x7fffffffull) return OPE_OK;
  Scratch<float4> clusters(ctx), sampled(ctx), normals(ctx), compacted(ctx), cnormals(ctx);
  Scratch<int> d_off(ctx), d_cnt(ctx), d_active(ctx), d_counts(ctx), d_ccounts(ctx), d_flags(ctx), d_samples(ctx), d_picks(ctx), d_knn(ctx);
  Scratch<float> spfh(ctx), fpfh(ctx), errors(ctx), transforms(ctx);
  Scratch<ope_reg_result> d_coarse(ctx), d_fine(ctx);
  Scratch<double> d_fit(ctx);
  OPE_TRY(clusters.alloc(total));
  {
    void* stage = nullptr;
    OPE_TRY(stage_reserve(ctx, total * sizeof(float4), &stage));
    float4* h = (float4*)stage;
    for (int i = 0; i < B; ++i) {
      if (!h_active[(size_t)i]) continue;
      const ope_frame_input& in = frames[i];
      float4* dst = h + h_off[(size_t)i];
      if (in.points) {
        const unsigned char* base = (const unsigned char*)in.points + in.offset;
        for (size_t j = 0; j < in.n; ++j) { float v[3]; std::memcpy(v, base + j * in.stride, 12); dst[j] = make_float4(v[0], v[1], v[2], 1.0f); }
      }
    }
    OPE_CUDA_TRY(ctx, cudaMemcpyAsync(clusters.p, stage, total * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    for (int i = 0; i < B; ++i) {   // device-resident clusters: device-to-device into their slot
      const ope_frame_input& in = frames[i];
      if (h_active[(size_t)i] && !in.points)
        OPE_CUDA_TRY(ctx, cudaMemcpyAsync(clusters.p + h_off[(size_t)i], ((const ope_cloud*)in.cloud)->pts, (size_t)h_cnt[(size_t)i] * sizeof(float4),
                                          cudaMemcpyDeviceToDevice, ctx->stream));
    }
  }
  // sampled clouds: [0] tp1 (1 cm target), [1] tp2 (8 mm target), [2] sp2 (8 mm source), each B x kBCap
  const size_t slab = (size_t)B * kBCap;
  OPE_TRY(sampled.alloc(3 * slab)); OPE_TRY(normals.alloc(3 * slab)); OPE_TRY(compacted.alloc(3 * slab)); OPE_TRY(cnormals.alloc(3 * slab));
  OPE_TRY(d_off.alloc(B)); OPE_TRY(d_cnt.alloc(B)); OPE_TRY(d_active.alloc(B)); OPE_TRY(d_counts.alloc(3 * (size_t)B)); OPE_TRY(d_ccounts.alloc(3 * (size_t)B));
  OPE_TRY(d_flags.alloc(B));
  OPE_CUDA_TRY(ctx, cudaMemcpyAsync(d_off.p, h_off.data(), B * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  OPE_CUDA_TRY(ctx, cudaMemcpyAsync(d_cnt.p, h_cnt.data(), B * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  OPE_CUDA_TRY(ctx, cudaMemcpyAsync(d_active.p, h_active.data(), B * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  OPE_CUDA_TRY(ctx, cudaMemsetAsync(d_flags.p, 0, B * sizeof(int), ctx->stream));
  OPE_CUDA_TRY(ctx, cudaMemsetAsync(d_counts.p, 0, 3 * (size_t)B * sizeof(int), ctx->stream));
  OPE_CUDA_TRY(ctx, cudaMemsetAsync(d_ccounts.p, 0, 3 * (size_t)B * sizeof(int), ctx->stream));
  ChunkTimer tm(ctx);
  OPE_TRY(dyn_smem(ctx, (const void*)bsample_kernel<false>, sizeof(SampleSmem)));
  OPE_TRY(dyn_smem(ctx, (const void*)bsample_kernel<true>, sizeof(SampleSmem)));
  tm.mark();   // [0] staging + H2D of the clusters
  // ---- target UniformSampling at the coarse and the fine leaf (subSampleAndCalculateNormals, :131-158) ----
  BSampleArgs sa;
  std::memset(&sa, 0, sizeof(sa));
  sa.pts = clusters.p; sa.offsets = d_off.p; sa.counts = d_cnt.p; sa.active = d_active.p; sa.flags = d_flags.p;
  sa.inv_leaf[0] = 1.0f / P.coarse_leaf; sa.inv_leaf[1] = 1.0f / P.fine_leaf;
  sa.cap[0] = kBFeatCap; sa.cap[1] = kBCap;
  sa.out[0] = sampled.p; sa.out[1] = sampled.p + slab;
  sa.out_counts[0] = d_counts.p; sa.out_counts[1] = d_counts.p + B;
  bsample_kernel<false><<<dim3(B, 2), kBThreads, sizeof(SampleSmem), ctx->stream>>>(sa);
  OPE_TRY(check_launch(ctx, "bsample_kernel<target>"));
  std::vector<int> h_counts, h_flags;
  OPE_TRY(download(ctx, d_counts.p, 2 * (size_t)B, h_counts));
  OPE_TRY(download(ctx, d_flags.p, (size_t)B, h_flags));
  int max_tp = 0, max_tp1 = 0, n_act = 0;
  for (int i = 0; i < B; ++i) {
    if (!h_active[(size_t)i]) continue;
    const int n1 = h_counts[(size_t)i], n2 = h_counts[(size_t)B + i];
    // the envelope of the fast path: coarse + fine stage both run, every cloud fits its shared-memory kernel
    if (h_flags[(size_t)i] || n1 < std::max(P.min_target_features, 1) || n1 > kBFeatCap || n2 < std::max(P.min_target_points, 1) || n2 > kBCap)
      h_active[(size_t)i] = 0;
    else { max_tp = std::max(max_tp, std::max(n1, n2)); max_tp1 = std::max(max_tp1, n1); ++n_act; }
  }
  if (n_act == 0) return OPE_OK;
  OPE_CUDA_TRY(ctx, cudaMemcpyAsync(d_active.p, h_active.data(), B * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  for (int i = 0; i < B; ++i) if (!h_active[(size_t)i]) { h_counts[(size_t)i] = 0; h_counts[(size_t)B + i] = 0; }
  OPE_CUDA_TRY(ctx, cudaMemcpyAsync(d_counts.p, h_counts.data(), 2 * (size_t)B * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  tm.mark();   // [1] target sampling
  // ---- normals of tp1 and tp2 in one launch, FPFH of tp1 (estimateCoarsePose, :16-48) ----
  const float vp[3] = {0, 0, 0};
  OPE_TRY(normals_smem_batch(ctx, sampled.p, d_counts.p, kBCap, 2 * B, max_tp, P.normal_k, vp, normals.p));
  tm.mark();   // [2] target normals
  OPE_TRY(spfh.alloc((size_t)B * kBCap * 33)); OPE_TRY(fpfh.alloc((size_t)B * kBCap * 33));
  OPE_TRY(fpfh_smem_batch(ctx, sampled.p, normals.p, d_counts.p, kBCap, B, max_tp1, P.fpfh_radius, spfh.p, fpfh.p));
  tm.mark();   // [3] FPFH
  // ---- findSimilarFeatures for every model point at once, SAC-IA pool, first strictly-lower error wins (:50-64) ----
  const int ns = (int)sp->n;
  OPE_TRY(d_knn.alloc((size_t)B * ns * K));
  OPE_TRY(feature_knn_batch(ctx, fpfh.p, d_counts.p, kBCap, B, max_tp1, d_fs, ns, 33, K, d_knn.p));
  tm.mark();   // [4] feature k-NN
  OPE_TRY(d_samples.alloc((size_t)B * H * S)); OPE_TRY(d_picks.alloc((size_t)B * H * S));
  {
    void* stage = nullptr;
    const size_t tb = (size_t)B * H * S;
    OPE_CUDA_TRY(ctx, stream_sync(ctx));   // the cluster staging buffer is reused
    OPE_TRY(stage_reserve(ctx, 2 * tb * sizeof(int), &stage));
    int* hs = (int*)stage;
    int* hp = hs + tb;
    // the tables of this chunk may still be on their way (drawn by a helper thread while the device worked)
    if (tables_ready) while (tables_ready->load(std::memory_order_acquire) < tables_needed) sched_yield();
    for (int i = 0; i < B; ++i) {
      if (h_active[(size_t)i] && (tables[i].n_hypotheses != H || tables[i].nr_samples != S || !tables[i].samples || !tables[i].picks))
        return fail(ctx, OPE_ERR_INVALID, "decision table of frame %d does not match the SAC-IA parameters", i);
    }
    for (int i = 0; i < B; ++i) {
      int* s = hs + (size_t)i * H * S;
      int* p = hp + (size_t)i * H * S;
      if (!h_active[(size_t)i]) { std::memset(s, 0, (size_t)H * S * sizeof(int)); std::memset(p, 0, (size_t)H * S * sizeof(int)); continue; }
      std::memcpy(s, tables[i].samples, (size_t)H * S * sizeof(int));
      std::memcpy(p, tables[i].picks, (size_t)H * S * sizeof(int));
      for (size_t e = 0; e < (size_t)H * S; ++e)
        if (s[e] < 0 || s[e] >= ns || p[e] < 0 || p[e] >= K) return fail(ctx, OPE_ERR_INVALID, "rng table entry out of range");
    }
    OPE_CUDA_TRY(ctx, cudaMemcpyAsync(d_samples.p, hs, tb * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    OPE_CUDA_TRY(ctx, cudaMemcpyAsync(d_picks.p, hp, tb * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  }
  OPE_TRY(errors.alloc((size_t)B * H)); OPE_TRY(transforms.alloc((size_t)B * H * 16)); OPE_TRY(d_coarse.alloc(B)); OPE_TRY(d_fine.alloc(B));
  SaciaBatch sb;
  std::memset(&sb, 0, sizeof(sb));
  sb.src = sp->pts; sb.ns = ns; sb.tgt = sampled.p; sb.counts = d_counts.p; sb.stride = kBCap;
  sb.nr_samples = S; sb.k_corr = K; sb.H = H; sb.samples = d_samples.p; sb.picks = d_picks.p; sb.knn_idx = d_knn.p; sb.active = d_active.p;
  sb.threshold = (float)P.sacia.max_correspondence_distance; sb.errors = errors.p; sb.transforms = transforms.p;
  {
    const char* me = std::getenv("OPE_BATCH_MORTON");
    if (!(me && std::atoi(me) == 0)) {   // the distance scan reads a Morton-ordered copy of the target (slot [0] of the compaction arrays is free)
      morton_sort_batch_kernel<<<B, kBThreads, 0, ctx->stream>>>(sampled.p, normals.p, d_counts.p, d_active.p, B, compacted.p, cnormals.p);
      OPE_TRY(check_launch(ctx, "morton_sort_batch_kernel"));
      sb.tgt_scan = compacted.p;
    }
  }
  Scratch<unsigned> d_best(ctx);
  OPE_TRY(d_best.alloc(B));
  {
    std::vector<unsigned> inf((size_t)B, 0x7f800000u);
    OPE_CUDA_TRY(ctx, cudaMemcpyAsync(d_best.p, inf.data(), (size_t)B * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
    OPE_CUDA_TRY(ctx, stream_sync(ctx));
  }
  sb.best_bits = d_best.p;
  { const char* e = std::getenv("OPE_SACIA_EARLY_EXIT"); sb.early_exit = !(e && std::atoi(e) == 0); }
  OPE_TRY(sacia_batch_device(ctx, sb, B, max_tp1, d_coarse.p));
  tm.mark();   // [5] SAC-IA
  // ---- estimateFinePose (:161-379): the model under the coarse pose sampled at the fine leaf, never materialised ----
  std::memset(&sa, 0, sizeof(sa));
  sa.pts = d_model->pts; sa.n_model = (int)d_model->n; sa.xforms = d_coarse.p; sa.active = d_active.p; sa.flags = d_flags.p;
  sa.inv_leaf[0] = 1.0f / P.fine_leaf; sa.cap[0] = kBCap; sa.out[0] = sampled.p + 2 * slab; sa.out_counts[0] = d_counts.p + 2 * (size_t)B;
  bsample_kernel<true><<<dim3(B, 1), kBThreads, sizeof(SampleSmem), ctx->stream>>>(sa);
  OPE_TRY(check_launch(ctx, "bsample_kernel<model>"));
  std::vector<int> h_sp2;
  OPE_TRY(download(ctx, d_counts.p + 2 * (size_t)B, (size_t)B, h_sp2));
  OPE_TRY(download(ctx, d_flags.p, (size_t)B, h_flags));
  int max_sp2 = 0;
  for (int i = 0; i < B; ++i) {
    if (!h_active[(size_t)i]) continue;
    if (h_flags[(size_t)i] || h_sp2[(size_t)i] <= 0) { h_active[(size_t)i] = 0; continue; }
    max_sp2 = std::max(max_sp2, h_sp2[(size_t)i]);
  }
  OPE_CUDA_TRY(ctx, cudaMemcpyAsync(d_active.p, h_active.data(), B * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  tm.mark();   // [6] model sampling
  OPE_TRY(normals_smem_batch(ctx, sampled.p + 2 * slab, d_counts.p + 2 * (size_t)B, kBCap, B, max_sp2, P.normal_k, vp, normals.p + 2 * slab));
  tm.mark();   // [7] source normals
  // removeNaNNormalsFromPointCloud on both fine clouds (:215-216): clouds [1] (tp2) and [2] (sp2)
  compact_normals_batch_kernel<<<2 * B, kBThreads, 0, ctx->stream>>>(sampled.p + slab, normals.p + slab, d_counts.p + B, d_active.p, B,
                                                                     compacted.p + slab, cnormals.p + slab, d_ccounts.p + B);
  OPE_TRY(check_launch(ctx, "compact_normals_batch_kernel"));
  std::vector<int> h_cc;
  OPE_TRY(download(ctx, d_ccounts.p + B, 2 * (size_t)B, h_cc));   // [0..B) tp2, [B..2B) sp2
  // the ICP's work order: both fine clouds along their Morton curve (the raw samples' arrays are free by now)
  const char* me = std::getenv("OPE_BATCH_MORTON");
  const bool morton = !(me && std::atoi(me) == 0);
  const float4* fine_pts = compacted.p;
  const float4* fine_nrm = cnormals.p;
  if (morton) {
    morton_sort_batch_kernel<<<2 * B, kBThreads, 0, ctx->stream>>>(compacted.p + slab, cnormals.p + slab, d_ccounts.p + B, d_active.p, B,
                                                                   sampled.p + slab, normals.p + slab);
    OPE_TRY(check_launch(ctx, "morton_sort_batch_kernel"));
    fine_pts = sampled.p; fine_nrm = normals.p;
  }
  std::vector<IcpBatchFrame> icp_frames;
  std::vector<int> icp_of;
  int max_tgt = 0;
  for (int i = 0; i < B; ++i) {
    if (!h_active[(size_t)i]) continue;
    const int nt = h_cc[(size_t)i], nsrc = h_cc[(size_t)B + i];
    if (nt < P.min_target_points || !icp_small_batch_applicable(P.icp, (size_t)nsrc, (size_t)nt)) { h_active[(size_t)i] = 0; continue; }
    IcpBatchFrame fr;
    fr.tgt_pts = fine_pts + slab + (size_t)i * kBCap; fr.tgt_nrm = fine_nrm + slab + (size_t)i * kBCap; fr.n_tgt = nt;
    fr.src_pts = fine_pts + 2 * slab + (size_t)i * kBCap; fr.src_nrm = fine_nrm + 2 * slab + (size_t)i * kBCap; fr.n_src = nsrc;
    icp_frames.push_back(fr);
    icp_of.push_back(i);
    max_tgt = std::max(max_tgt, nt);
  }
  if (icp_frames.empty()) return OPE_OK;
  Scratch<ope_reg_result> d_icp(ctx);
  Scratch<int> d_icp_active(ctx), d_tc(ctx), d_sc(ctx);
  const int NI = (int)icp_frames.size();
  OPE_TRY(d_icp.alloc(NI));
  tm.mark();   // [8] NaN-normal compaction
  OPE_TRY(icp_small_batch_device(ctx, P.icp, icp_frames.data(), NI, d_icp.p));
  tm.mark();   // [9] ICP
  // ---- getFitnessScore (:354) of every aligned pair; the ICP results are indexed by their position in icp_frames ----
  OPE_TRY(d_fit.alloc(B));
  {
    // scatter the compact ICP results back to frame order so that the fitness kernel can index by frame
    std::vector<ope_reg_result> h_icp;
    OPE_TRY(download(ctx, d_icp.p, (size_t)NI, h_icp));
    std::vector<ope_reg_result> by_frame((size_t)B);
    std::memset(by_frame.data(), 0, by_frame.size() * sizeof(ope_reg_result));
    for (int j = 0; j < NI; ++j) by_frame[(size_t)icp_of[(size_t)j]] = h_icp[(size_t)j];
    OPE_CUDA_TRY(ctx, cudaMemcpyAsync(d_fine.p, by_frame.data(), (size_t)B * sizeof(ope_reg_result), cudaMemcpyHostToDevice, ctx->stream));
    OPE_CUDA_TRY(ctx, cudaMemcpyAsync(d_active.p, h_active.data(), B * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    OPE_TRY(fitness_batch_device(ctx, fine_pts + slab, d_ccounts.p + B, fine_pts + 2 * slab, d_ccounts.p + 2 * (size_t)B, kBCap, B, max_tgt,
                                 d_fine.p, d_active.p, d_fit.p));
    tm.mark();   // [10] fitness
    std::vector<ope_reg_result> h_coarse;
    std::vector<double> h_fit;
    OPE_TRY(download(ctx, d_coarse.p, (size_t)B, h_coarse));
    OPE_TRY(download(ctx, d_fit.p, (size_t)B, h_fit));   // also orders the stack vectors above after their copies
    // ---- estimateFinalPose (:383-448): finalPose = rigidmodelPose * (coarse * fine) ----
    for (int i = 0; i < B; ++i) {
      if (!h_active[(size_t)i]) continue;
      ope_pose_result& r = results[i];
      std::memset(&r, 0, sizeof(r));
      Mat4 coarse, fine;
      std::memcpy(coarse.m, h_coarse[(size_t)i].T, 64);
      std::memcpy(fine.m, by_frame[(size_t)i].T, 64);
      const Mat4 final_pose = mat4_mul(rigid, mat4_mul(coarse, fine));
      std::memcpy(r.final_pose, final_pose.m, 64); std::memcpy(r.coarse_pose, coarse.m, 64);
      std::memcpy(r.fine_pose, fine.m, 64); std::memcpy(r.rigid_model_pose, rigid.m, 64);
      r.fitness = h_fit[(size_t)i];
      const int nsrc = h_cc[(size_t)B + i], nt = h_cc[(size_t)i];
      r.align_strength = (double)by_frame[(size_t)i].n_correspondences / (double)((long)nsrc + (long)nt);
      r.ran_coarse = 1;
      r.icp_iterations = by_frame[(size_t)i].iterations; r.icp_converged = by_frame[(size_t)i].converged; r.icp_state = by_frame[(size_t)i].state;
      r.n_src_coarse = ns; r.n_tgt_coarse = h_counts[(size_t)i]; r.n_src_fine = nsrc; r.n_tgt_fine = nt;
      r.sacia_best_iteration = h_coarse[(size_t)i].best_iteration; r.sacia_best_error = h_coarse[(size_t)i].best_error;
      done[i] = 1;
    }
  }
  return OPE_OK;
}

}  // namespace ope
