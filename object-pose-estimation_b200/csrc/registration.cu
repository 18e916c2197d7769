// registration.cu — rigid transform estimation (K10/K11), fitness (K9), the fused ICP loop (K8) and the SAC-IA
// hypothesis pool (K7).
//
// ICP is ONE persistent cooperative kernel per align(): every iteration fuses correspondence search on the
// target grid, the rejector chain, the centroid / 3x3 cross-covariance reduction (double, fixed-shape tree:
// thread -> warp shuffle -> block -> per-block partials in HBM), ONE grid barrier, then in every block
// redundantly the reduction of the per-block partials in a fixed order, the 3x3 Jacobi SVD (Umeyama), the
// convergence test and the in-place rigid transform of the block's own source points. Nothing but the 4x4,
// the flags and the iteration count returns to the host.
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "ope_host.cuh"
#include "ope_octet.cuh"

namespace cg = cooperative_groups;

namespace ope {

// ===================================================================================== small kernels ===
__global__ void transform_kernel(const float4* __restrict__ in, const float4* __restrict__ in_n, int n, Mat4 T,
                                 float4* __restrict__ out, float4* __restrict__ out_n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = __ldg(in + i);
  float4 o = p;
  xform_point(T, p.x, p.y, p.z, o.x, o.y, o.z);
  out[i] = o;
  if (in_n && out_n) {
    float4 v = __ldg(in_n + i);
    float4 w = v;
    xform_normal(T, v.x, v.y, v.z, w.x, w.y, w.z);
    out_n[i] = w;
  }
}

static constexpr int kRedThreads = 256;
static constexpr int kMomentAcc = 16;  // n, Ss[3], St[3], Sts[9]

__device__ __forceinline__ double warp_sum_d(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-level sum of NACC doubles per thread -> out[NACC] valid in ALL threads of warp 0 / written to dst by thread 0.
template <int NACC>
__device__ __forceinline__ void block_reduce_store(double* acc, double* smem /* (blockDim/32)*NACC */, double* dst) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int a = 0; a < NACC; ++a) acc[a] = warp_sum_d(acc[a]);
  if (lane == 0)
    for (int a = 0; a < NACC; ++a) smem[warp * NACC + a] = acc[a];
  __syncthreads();
  if (threadIdx.x < NACC) {
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += smem[w * NACC + threadIdx.x];
    dst[threadIdx.x] = s;
  }
  __syncthreads();
}

// raw moments of n pairs (src[is[i]], tgt[it[i]]) -> partials[block][16]
__global__ void moments_kernel(const float4* __restrict__ src, const float4* __restrict__ tgt, const int* __restrict__ is,
                               const int* __restrict__ it, int n, double* __restrict__ partials) {
  __shared__ double smem[(kRedThreads / 32) * kMomentAcc];
  double acc[kMomentAcc];
  for (int a = 0; a < kMomentAcc; ++a) acc[a] = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 s = __ldg(src + (is ? is[i] : i));
    const float4 t = __ldg(tgt + (it ? it[i] : i));
    acc[0] += 1.0;
    acc[1] += s.x; acc[2] += s.y; acc[3] += s.z;
    acc[4] += t.x; acc[5] += t.y; acc[6] += t.z;
    const double sv[3] = {s.x, s.y, s.z}, tv[3] = {t.x, t.y, t.z};
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r) acc[7 + c * 3 + r] += tv[r] * sv[c];
  }
  block_reduce_store<kMomentAcc>(acc, smem, partials + (size_t)blockIdx.x * kMomentAcc);
}
__global__ void umeyama_final_kernel(const double* __restrict__ partials, int nblocks, float* __restrict__ out16) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double acc[kMomentAcc];
  for (int a = 0; a < kMomentAcc; ++a) acc[a] = 0.0;
  for (int b = 0; b < nblocks; ++b)
    for (int a = 0; a < kMomentAcc; ++a) acc[a] += partials[(size_t)b * kMomentAcc + a];
  Mat4 T;
  umeyama_from_moments(acc, T);
  for (int i = 0; i < 16; ++i) out16[i] = T.m[i];
}

// fitness: sum of NN-1 squared distances <= max_range of the transformed source (double), plus the count.
// One octet (8 lanes) per source point.
__global__ void __launch_bounds__(kRedThreads) fitness_kernel(GridView g, const float4* __restrict__ src, int n, Mat4 T,
                                                              float max_range_f, double* __restrict__ partials) {
  __shared__ double smem[(kRedThreads / 32) * 2];
  __shared__ OctStack stacks[kRedThreads / 8];
  const Octet o = octet_self();
  OctStack* st = &stacks[threadIdx.x >> 3];
  const int oct_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, n_oct = (gridDim.x * blockDim.x) >> 3;
  double acc[2] = {0.0, 0.0};
  for (int i = oct_id; i < n; i += n_oct) {
    const float4 p = __ldg(src + i);
    float x, y, z;
    xform_point(T, p.x, p.y, p.z, x, y, z);
    const bool ok = finite3(x, y, z);
    float d2;
    const int idx = octet_nn1(g, st, o, ok, x, y, z, FLT_MAX, d2);
    if (o.sub == 0u && ok && idx >= 0 && d2 <= max_range_f) { acc[0] += (double)d2; acc[1] += 1.0; }
  }
  block_reduce_store<2>(acc, smem, partials + (size_t)blockIdx.x * 2);
}
__global__ void sum_partials_kernel(const double* __restrict__ partials, int nblocks, int nacc, double* __restrict__ out) {
  if (blockIdx.x != 0 || threadIdx.x >= nacc) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += partials[(size_t)b * nacc + threadIdx.x];
  out[threadIdx.x] = s;
}

// =============================================================================================== ICP ====
struct IcpDev {
  GridView grid;
  const float4* tgt_pts;      // original order (match index -> point)
  const float4* tgt_nrm;      // original order or null
  const float4* src0_pts;     // untouched input (variant MODCORR reads stale normals/points from here)
  const float4* src0_nrm;
  float4* cur_pts;            // working copy, transformed in place every iteration
  float4* cur_nrm;
  int n_src;
  // parameters
  int max_iterations, min_corr, estimator, k_search, n_rej, transformation, variant, force_all;
  int rej_kind[OPE_MAX_REJECTORS];
  double rej_thr[OPE_MAX_REJECTORS];
  double max_corr_dist;       // as given (unsquared)
  float max_d2_f;             // squared, clamped to FLT_MAX (nearest estimator search bound)
  double rot_thr, trans_thr, rel_mse_thr, abs_mse_thr;
  int max_similar, fail_after_max;
  // outputs
  int* corr_match;            // n_src, -1 = no correspondence this iteration
  float* corr_d2;             // n_src
  double* partials;           // 2 * gridDim * kIcpAcc (double buffered)
  ope_reg_result* result;     // device copy
  long long* phase_cycles;    // optional (OPE_PROFILE=1): block 0's cycles in [search+reduce, grid barrier, partial sums + SVD, transform]
  Mat4 guess;
};

static constexpr int kIcpThreads = 256;
static constexpr int kIcpAcc = 17;  // moments[16] + sum of correspondence distances

// One correspondence for source point i in its CURRENT position p, computed by the octet that owns the point
// (8 lanes, same arguments). Returns the match index or -1 in every lane.
__device__ __forceinline__ int icp_correspond(const IcpDev& a, OctStack* st, OctKnnList* L, const Octet& o, int i, const float4 p,
                                              float& d2_out) {
  const bool ok = finite3(p.x, p.y, p.z);
  const bool stale = a.variant == OPE_ICP_VARIANT_MODCORR;
  int match = -1;
  float d2 = 0.0f;
  if (a.estimator == OPE_EST_NEAREST) {
    match = octet_nn1(a.grid, st, o, ok, p.x, p.y, p.z, a.max_d2_f, d2);
    if (match >= 0 && (double)d2 > a.max_corr_dist * a.max_corr_dist) match = -1;
  } else {
    const int cnt = octet_knn(a.grid, st, L, o, ok, p.x, p.y, p.z, a.k_search);
    if (cnt > 0) {
      // CorrespondenceEstimationNormalShooting: among the k nearest, the one with the smallest squared distance to the
      // line through p along the source normal (double); the first minimum in list order wins.
      const float4 nr = stale ? __ldg(a.src0_nrm + i) : a.cur_nrm[i];
      const double N[3] = {nr.x, nr.y, nr.z};
      double min_dist = DBL_MAX;
      int min_index = 0x7fffffff;
      for (int j = (int)o.sub; j < cnt; j += 8) {
        const float4 t = __ldg(a.tgt_pts + L->i[j]);
        const float px = t.x - p.x, py = t.y - p.y, pz = t.z - p.z;
        const double V[3] = {px, py, pz};
        const double C0 = N[1] * V[2] - N[2] * V[1], C1 = N[2] * V[0] - N[0] * V[2], C2 = N[0] * V[1] - N[1] * V[0];
        const double dist = C0 * C0 + C1 * C1 + C2 * C2;
        if (dist < min_dist) { min_dist = dist; min_index = j; }
      }
#pragma unroll
      for (int s = 1; s < 8; s <<= 1) {
        const double od = __shfl_xor_sync(o.mask, min_dist, s);
        const int oi = __shfl_xor_sync(o.mask, min_index, s);
        if (od < min_dist || (od == min_dist && oi < min_index)) { min_dist = od; min_index = oi; }
      }
      if (min_index != 0x7fffffff && !(min_dist > a.max_corr_dist)) {  // sic: squared cross norm vs unsquared threshold
        match = L->i[min_index];
        d2 = L->d[min_index];
      }
    }
    __syncwarp(o.mask);  // the list is reused by the octet's next point
  }
  if (match >= 0) {
    // rejector chain, VP/impl/icp_mod.hpp:194-208
    for (int r = 0; r < a.n_rej; ++r) {
      const float4 sn = stale ? __ldg(a.src0_nrm + i) : a.cur_nrm[i];
      double score;
      if (a.rej_kind[r] == OPE_REJ_SURFACE_NORMAL) {
        const float4 tn = __ldg(a.tgt_nrm + match);
        score = (double)((sn.x * tn.x) + (sn.y * tn.y) + (sn.z * tn.z));
      } else {
        const float4 sp = stale ? __ldg(a.src0_pts + i) : p;
        const double s = (double)sqrtf(sp.x * sp.x + sp.y * sp.y + sp.z * sp.z);
        score = (double)((sn.x * (-sp.x / s)) + (sn.y * (-sp.y / s)) + (sn.z * (-sp.z / s)));
      }
      if (!(score > a.rej_thr[r])) { match = -1; break; }
    }
  }
  d2_out = d2;
  return match;
}

__global__ void __launch_bounds__(kIcpThreads) icp_kernel(IcpDev a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double smem[(kIcpThreads / 32) * kIcpAcc];
  __shared__ double totals[kIcpAcc];
  __shared__ OctStack stacks[kIcpThreads / 8];
  __shared__ OctKnnList lists[kIcpThreads / 8];
  const Octet o = octet_self();
  OctStack* st = &stacks[threadIdx.x >> 3];
  OctKnnList* L = &lists[threadIdx.x >> 3];
  const int oct_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, n_oct = (gridDim.x * blockDim.x) >> 3;
  __shared__ Mat4 T_inc;
  __shared__ int s_stop;  // 0 continue, 1 stop, 2 stop without a transform (not enough correspondences)

  // input_transformed = guess applied to input (VP/impl/icp_mod.hpp:132-139)
  const bool have_guess = !mat4_is_identity(a.guess);
  for (int i = oct_id; i < a.n_src; i += n_oct) {
    if (o.sub != 0u) continue;
    float4 p = __ldg(a.src0_pts + i);
    float4 v = a.src0_nrm ? __ldg(a.src0_nrm + i) : make_float4(0, 0, 0, 0);
    if (have_guess && finite3(p.x, p.y, p.z)) {
      float x, y, z;
      xform_point(a.guess, p.x, p.y, p.z, x, y, z);
      p.x = x; p.y = y; p.z = z;
      if (a.src0_nrm && finite3(v.x, v.y, v.z)) {
        xform_normal(a.guess, v.x, v.y, v.z, x, y, z);
        v.x = x; v.y = y; v.z = z;
      }
    }
    a.cur_pts[i] = p;
    if (a.cur_nrm) a.cur_nrm[i] = v;
  }
  // per-block replicated state (identical in every block: same inputs, same order of operations)
  Mat4 final_t = a.guess;
  double prev_mse = DBL_MAX, cur_mse = DBL_MAX;
  int similar = 0, iterations = 0, state = OPE_CONV_NOT_CONVERGED, converged = 0, n_corr = 0;
  int pass = 0;  // uniform across all threads: selects the partials buffer

  long long t_phase[4] = {0, 0, 0, 0};
  const bool prof = a.phase_cycles != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  for (;; ++pass) {
    long long t0 = prof ? clock64() : 0;
    // ---- phase 1: correspondences + moments ----
    double acc[kIcpAcc];
#pragma unroll
    for (int k = 0; k < kIcpAcc; ++k) acc[k] = 0.0;
    for (int i = oct_id; i < a.n_src; i += n_oct) {  // one octet (8 lanes) per source point
      const float4 p = a.cur_pts[i];
      float d2 = 0.0f;
      const int m = icp_correspond(a, st, L, o, i, p, d2);
      if (o.sub != 0u) continue;  // lane 0 of the octet records and accumulates
      a.corr_match[i] = m;
      a.corr_d2[i] = d2;
      if (m >= 0) {
        const float4 t = __ldg(a.tgt_pts + m);
        acc[0] += 1.0;
        acc[1] += p.x; acc[2] += p.y; acc[3] += p.z;
        acc[4] += t.x; acc[5] += t.y; acc[6] += t.z;
        const double sv[3] = {p.x, p.y, p.z}, tv[3] = {t.x, t.y, t.z};
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int r = 0; r < 3; ++r) acc[7 + c * 3 + r] += tv[r] * sv[c];
        acc[16] += (double)d2;
      }
    }
    double* my_partials = a.partials + ((size_t)(pass & 1) * gridDim.x + blockIdx.x) * kIcpAcc;
    block_reduce_store<kIcpAcc>(acc, smem, my_partials);
    __threadfence();
    if (prof) { const long long t1 = clock64(); t_phase[0] += t1 - t0; t0 = t1; }
    grid.sync();
    if (prof) { const long long t1 = clock64(); t_phase[1] += t1 - t0; t0 = t1; }
    // ---- phase 2: every block reduces all partials in the same order ----
    if (threadIdx.x < kIcpAcc) {
      const double* base = a.partials + (size_t)(pass & 1) * gridDim.x * kIcpAcc;
      double s = 0.0;
      for (int b = 0; b < (int)gridDim.x; ++b) s += __ldcg(base + (size_t)b * kIcpAcc + threadIdx.x);
      totals[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int stop = 0;
      n_corr = (int)totals[0];
      if (n_corr < a.min_corr) {
        state = OPE_CONV_NO_CORRESPONDENCES; converged = 0; stop = 2;  // VP/impl/icp_mod.hpp:232-240
        T_inc = mat4_identity();
      } else {
        Mat4 T;
        umeyama_from_moments(totals, T);
        T_inc = T;
        final_t = mat4_mul(T, final_t);
        ++iterations;
        // DefaultConvergenceCriteria::hasConverged
        state = OPE_CONV_NOT_CONVERGED;
        int conv = 0;
        if (iterations >= a.max_iterations) {
          if (!a.fail_after_max) { state = OPE_CONV_ITERATIONS; conv = 1; }
          else { conv = 0; stop = 1; }
        } else {
          const double cos_angle = 0.5 * (double)(T(0, 0) + T(1, 1) + T(2, 2) - 1);
          const double translation_sqr = (double)(T(0, 3) * T(0, 3) + T(1, 3) * T(1, 3) + T(2, 3) * T(2, 3));
          int hit = 0, hit_state = 0;
          if (cos_angle >= a.rot_thr && translation_sqr <= a.trans_thr) { hit = 1; hit_state = OPE_CONV_TRANSFORM; }
          else {
            cur_mse = totals[16] / (double)n_corr;
            if (fabs(cur_mse - prev_mse) < a.abs_mse_thr) { hit = 1; hit_state = OPE_CONV_ABS_MSE; }
            else if (fabs(cur_mse - prev_mse) / prev_mse < a.rel_mse_thr) { hit = 1; hit_state = OPE_CONV_REL_MSE; }
            else prev_mse = cur_mse;
          }
          if (hit) {
            if (similar < a.max_similar) ++similar;
            else { similar = 0; state = hit_state; conv = 1; }
          }
        }
        converged = conv;
        if (a.force_all && iterations < a.max_iterations) conv = 0;
        if (conv) stop = 1;
      }
      s_stop = stop;
    }
    __syncthreads();
    if (prof) { const long long t1 = clock64(); t_phase[2] += t1 - t0; t0 = t1; }
    // ---- phase 3: transformCloud(input_transformed, transformation_), own points only ----
    const int stop = s_stop;
    if (stop != 2) {
      const Mat4 T = T_inc;
      // the octet that searches point i is also the one that moves it (lane 0): no cross-block hazard on cur_pts
      for (int i = oct_id; i < a.n_src; i += n_oct) {
        if (o.sub != 0u) continue;
        float4 p = a.cur_pts[i];
        if (!finite3(p.x, p.y, p.z)) continue;
        float x, y, z;
        xform_point(T, p.x, p.y, p.z, x, y, z);
        p.x = x; p.y = y; p.z = z;
        a.cur_pts[i] = p;
        if (a.cur_nrm) {
          float4 v = a.cur_nrm[i];
          if (finite3(v.x, v.y, v.z)) {
            xform_normal(T, v.x, v.y, v.z, x, y, z);
            v.x = x; v.y = y; v.z = z;
            a.cur_nrm[i] = v;
          }
        }
      }
    }
    if (prof) { const long long t1 = clock64(); t_phase[3] += t1 - t0; }
    if (stop) break;
  }
  if (prof) for (int i = 0; i < 4; ++i) a.phase_cycles[i] = t_phase[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ope_reg_result r;
    for (int i = 0; i < 16; ++i) r.T[i] = final_t.m[i];
    r.converged = converged; r.state = state; r.iterations = iterations; r.n_correspondences = n_corr;
    r.last_mse = cur_mse; r.best_error = 0.0; r.best_iteration = 0; r.reserved = 0;
    *a.result = r;
  }
}

// one estimation + rejection pass (no loop), one octet per source point
__global__ void __launch_bounds__(kIcpThreads) correspond_once_kernel(IcpDev a) {
  __shared__ OctStack stacks[kIcpThreads / 8];
  __shared__ OctKnnList lists[kIcpThreads / 8];
  const Octet o = octet_self();
  OctStack* st = &stacks[threadIdx.x >> 3];
  OctKnnList* L = &lists[threadIdx.x >> 3];
  const int oct_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, n_oct = (gridDim.x * blockDim.x) >> 3;
  for (int i = oct_id; i < a.n_src; i += n_oct) {
    const float4 p = a.cur_pts[i];
    float d2 = 0.0f;
    const int m = icp_correspond(a, st, L, o, i, p, d2);
    if (o.sub == 0u) { a.corr_match[i] = m; a.corr_d2[i] = d2; }
  }
}

// ============================================================================================ SAC-IA ====
struct SaciaDev {
  GridView grid;
  const float4* src;      // ns
  const float4* tgt;      // original order
  int ns;
  int nr_samples, k_corr;
  const int* samples;     // H * nr_samples
  const int* picks;       // H * nr_samples
  const int* knn_idx;     // ns * k_corr: feature-space neighbours of every source point
  int h_begin;
  float threshold;        // TruncatedError threshold (applied to squared distances)
  float* terms;           // scratch: gridDim * ns floats
  float* errors;          // H
  float* transforms;      // H * 16
};

static constexpr int kSaciaThreads = 256;
static constexpr int kSaciaMaxSamples = 16;

// one block per hypothesis: 5-point Umeyama (double moments, as everywhere), then every source point's
// truncated NN error in parallel, then the float sum in point order by one thread (bit-exact with the reference's
// serial `error += ...`, so the first-lowest-error hypothesis is the same one).
__global__ void __launch_bounds__(kSaciaThreads) sacia_kernel(SaciaDev a) {
  __shared__ Mat4 T;
  __shared__ OctStack stacks[kSaciaThreads / 8];
  const int h = a.h_begin + blockIdx.x;
  if (threadIdx.x == 0) {
    double acc[16];
    for (int i = 0; i < 16; ++i) acc[i] = 0.0;
    for (int j = 0; j < a.nr_samples; ++j) {
      const int si = a.samples[(size_t)h * a.nr_samples + j];
      const int pick = a.picks[(size_t)h * a.nr_samples + j];
      int ti = a.knn_idx[(size_t)si * a.k_corr + pick];
      if (ti < 0) ti = a.knn_idx[(size_t)si * a.k_corr];  // fewer than k target features
      const float4 sp = __ldg(a.src + si), tp = __ldg(a.tgt + ti);
      const double sv[3] = {sp.x, sp.y, sp.z}, tv[3] = {tp.x, tp.y, tp.z};
      acc[0] += 1.0;
      for (int k = 0; k < 3; ++k) { acc[1 + k] += sv[k]; acc[4 + k] += tv[k]; }
      for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) acc[7 + c * 3 + r] += tv[r] * sv[c];
    }
    Mat4 M;
    umeyama_from_moments(acc, M);  // double moments + double SVD, rounded to float once (same as every other Umeyama here)
    T = M;
    for (int i = 0; i < 16; ++i) a.transforms[(size_t)h * 16 + i] = M.m[i];
  }
  __syncthreads();
  const Mat4 M = T;
  float* terms = a.terms + (size_t)blockIdx.x * a.ns;
  const Octet o = octet_self();
  OctStack* st = &stacks[threadIdx.x >> 3];
  for (int i = threadIdx.x >> 3; i < a.ns; i += kSaciaThreads / 8) {  // one octet per source point
    const float4 p = __ldg(a.src + i);
    float x, y, z;
    xform_point(M, p.x, p.y, p.z, x, y, z);
    const bool ok = finite3(x, y, z);
    float d2;
    const int idx = octet_nn1(a.grid, st, o, ok, x, y, z, a.threshold, d2);
    if (o.sub == 0u) terms[i] = (ok && idx >= 0 && d2 <= a.threshold) ? d2 / a.threshold : 1.0f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float error = 0.0f;
    for (int i = 0; i < a.ns; ++i) error += terms[i];
    a.errors[h] = error;
  }
}
// first strictly-lower error wins, in hypothesis order
__global__ void sacia_select_kernel(const float* __restrict__ errors, const float* __restrict__ transforms, int h_begin,
                                    int h_end, ope_reg_result* __restrict__ out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int best = -1;
  float lowest = 0.0f;
  for (int h = h_begin; h < h_end; ++h) {
    const float e = errors[h];
    if (best < 0 || e < lowest) { lowest = e; best = h; }
  }
  ope_reg_result r;
  for (int i = 0; i < 16; ++i) r.T[i] = (best >= 0) ? transforms[(size_t)best * 16 + i] : ((i % 5 == 0) ? 1.0f : 0.0f);
  r.converged = best >= 0; r.state = 0; r.iterations = 0; r.n_correspondences = 0; r.last_mse = 0.0;
  r.best_error = lowest; r.best_iteration = best; r.reserved = 0;
  *out = r;
}

// ================================================================================================ host ==
int transform_device(ope_ctx* ctx, const ope_cloud* in, const Mat4& T, ope_cloud* out) {
  if (in->n == 0) return OPE_OK;
  transform_kernel<<<div_up(in->n, 256), 256, 0, ctx->stream>>>(in->pts, in->normals, (int)in->n, T, out->pts, out->normals);
  return check_launch(ctx, "transform_kernel");
}

int umeyama_device(ope_ctx* ctx, const float4* src, const float4* tgt, const int* d_isrc, const int* d_itgt, size_t n,
                   float T[16]) {
  const int nb = (int)std::min<size_t>(std::max<size_t>(1, (n + kRedThreads - 1) / kRedThreads), (size_t)ctx->sm_count * 4);
  Scratch<double> partials(ctx);
  Scratch<float> out(ctx);
  OPE_TRY(partials.alloc((size_t)nb * kMomentAcc));
  OPE_TRY(out.alloc(16));
  moments_kernel<<<nb, kRedThreads, 0, ctx->stream>>>(src, tgt, d_isrc, d_itgt, (int)n, partials.p);
  OPE_TRY(check_launch(ctx, "moments_kernel"));
  umeyama_final_kernel<<<1, 32, 0, ctx->stream>>>(partials.p, nb, out.p);
  OPE_TRY(check_launch(ctx, "umeyama_final_kernel"));
  void* h;
  OPE_TRY(read_back(ctx, out.p, 16 * sizeof(float), &h));
  std::memcpy(T, h, 16 * sizeof(float));
  return OPE_OK;
}

int fitness_device(ope_ctx* ctx, const ope_cloud* src, const ope_cloud* tgt, const Mat4& T, double max_range, double* out) {
  *out = DBL_MAX;
  if (src->n == 0) return OPE_OK;
  OPE_TRY(cloud_bbox(ctx, const_cast<ope_cloud*>(tgt)));
  GridView g;
  OPE_TRY(cloud_grid(ctx, tgt, knn_cell_size(tgt, 1), &g));
  const int nb = (int)std::min<size_t>(std::max<size_t>(1, (src->n * 8 + kRedThreads - 1) / kRedThreads), (size_t)ctx->sm_count * 4);
  Scratch<double> partials(ctx), fin(ctx);
  OPE_TRY(partials.alloc((size_t)nb * 2));
  OPE_TRY(fin.alloc(2));
  const float mr = max_range >= (double)FLT_MAX ? FLT_MAX : (float)max_range;
  fitness_kernel<<<nb, kRedThreads, 0, ctx->stream>>>(g, src->pts, (int)src->n, T, mr, partials.p);
  OPE_TRY(check_launch(ctx, "fitness_kernel"));
  sum_partials_kernel<<<1, 32, 0, ctx->stream>>>(partials.p, nb, 2, fin.p);
  OPE_TRY(check_launch(ctx, "sum_partials_kernel"));
  void* h;
  OPE_TRY(read_back(ctx, fin.p, 2 * sizeof(double), &h));
  const double* r = (const double*)h;
  *out = r[1] > 0 ? r[0] / r[1] : DBL_MAX;
  return OPE_OK;
}

static int icp_fill(ope_ctx* ctx, const ope_cloud* src, const ope_cloud* tgt, const ope_icp_params& prm, IcpDev* a) {
  const bool need_src_normals = prm.estimator == OPE_EST_NORMAL_SHOOTING || prm.n_rejectors > 0;
  bool need_tgt_normals = false;
  for (int r = 0; r < prm.n_rejectors; ++r) {
    if (prm.rejector_kind[r] != OPE_REJ_SURFACE_NORMAL && prm.rejector_kind[r] != OPE_REJ_SELF_OCCLUDED_NORMAL)
      return fail(ctx, OPE_ERR_UNSUPPORTED, "unknown correspondence rejector kind %d", prm.rejector_kind[r]);
    if (prm.rejector_kind[r] == OPE_REJ_SURFACE_NORMAL) need_tgt_normals = true;
  }
  if (prm.n_rejectors < 0 || prm.n_rejectors > OPE_MAX_REJECTORS) return fail(ctx, OPE_ERR_INVALID, "bad rejector count");
  if (need_src_normals && !src->normals) return fail(ctx, OPE_ERR_INVALID, "estimator/rejector needs source normals");
  if (need_tgt_normals && !tgt->normals) return fail(ctx, OPE_ERR_INVALID, "rejector needs target normals");
  if (prm.use_reciprocal) return fail(ctx, OPE_ERR_UNSUPPORTED, "reciprocal correspondences are not implemented");
  if (prm.transformation != OPE_TE_SVD) return fail(ctx, OPE_ERR_UNSUPPORTED, "only TransformationEstimationSVD is implemented on the device");
  if (prm.estimator == OPE_EST_NORMAL_SHOOTING && (prm.k_search < 1 || prm.k_search > 32))
    return fail(ctx, OPE_ERR_INVALID, "normal shooting k must be in [1, 32]");
  OPE_TRY(cloud_bbox(ctx, const_cast<ope_cloud*>(tgt)));
  const int kk = prm.estimator == OPE_EST_NORMAL_SHOOTING ? prm.k_search : 1;
  OPE_TRY(cloud_grid(ctx, tgt, knn_cell_size(tgt, kk), &a->grid));
  a->tgt_pts = tgt->pts; a->tgt_nrm = tgt->normals;
  a->src0_pts = src->pts; a->src0_nrm = src->normals;
  a->n_src = (int)src->n;
  a->max_iterations = prm.max_iterations; a->min_corr = prm.min_number_correspondences;
  a->estimator = prm.estimator; a->k_search = prm.k_search; a->n_rej = prm.n_rejectors;
  a->transformation = prm.transformation; a->variant = prm.variant; a->force_all = prm.force_all_iterations;
  for (int r = 0; r < OPE_MAX_REJECTORS; ++r) { a->rej_kind[r] = prm.rejector_kind[r]; a->rej_thr[r] = prm.rejector_threshold[r]; }
  a->max_corr_dist = prm.max_correspondence_distance;
  const double m2 = prm.max_correspondence_distance * prm.max_correspondence_distance;
  a->max_d2_f = (m2 >= (double)FLT_MAX || !(m2 == m2)) ? FLT_MAX : (float)m2;
  a->rot_thr = 1.0 - prm.transformation_epsilon;
  a->trans_thr = prm.transformation_epsilon;
  a->rel_mse_thr = prm.euclidean_fitness_epsilon;
  a->abs_mse_thr = prm.mse_threshold_absolute;
  a->max_similar = prm.max_iterations_similar_transforms;
  a->fail_after_max = prm.failure_after_max_iterations;
  return OPE_OK;
}

static void corr_to_host(const std::vector<int>& match, const std::vector<float>& d2, ope_correspondence* out, size_t* n) {
  size_t w = 0;
  for (size_t i = 0; i < match.size(); ++i)
    if (match[i] >= 0) out[w++] = ope_correspondence{(int32_t)i, match[i], d2[i]};
  if (n) *n = w;
}

int icp_device(ope_ctx* ctx, const ope_cloud* src, const ope_cloud* tgt, const ope_icp_params& prm, const Mat4& guess,
               ope_reg_result* res, ope_correspondence* out_corr_host, ope_cloud** out_aligned) {
  IcpDev a;
  std::memset(&a, 0, sizeof(a));
  OPE_TRY(icp_fill(ctx, src, tgt, prm, &a));
  a.guess = guess;
  const size_t n = src->n;
  ope_cloud* work = nullptr;
  OPE_TRY(cloud_alloc(ctx, n, src->normals != nullptr, &work));
  a.cur_pts = work->pts; a.cur_nrm = work->normals;
  Scratch<int> match(ctx);
  Scratch<float> d2(ctx);
  Scratch<double> partials(ctx);
  Scratch<ope_reg_result> dres(ctx);
  int rc = match.alloc(n);
  if (rc == OPE_OK) rc = d2.alloc(n);
  if (rc == OPE_OK) rc = dres.alloc(1);
  // cooperative grid: enough threads for one point each, capped by co-residency
  int per_sm = 0;
  if (rc == OPE_OK && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, icp_kernel, kIcpThreads, 0) != cudaSuccess)
    rc = fail(ctx, OPE_ERR_CUDA, "occupancy query failed");
  int blocks = 1;
  if (rc == OPE_OK) {
    if (per_sm < 1) rc = fail(ctx, OPE_ERR_CUDA, "icp_kernel cannot be resident");
    const int max_blocks = per_sm * ctx->sm_count;
    blocks = (int)std::min<size_t>(std::max<size_t>(1, (n * 8 + kIcpThreads - 1) / kIcpThreads), (size_t)max_blocks);
  }
  if (rc == OPE_OK) rc = partials.alloc((size_t)2 * blocks * kIcpAcc);
  Scratch<long long> phases(ctx);
  const bool profile = std::getenv("OPE_PROFILE") != nullptr;
  if (rc == OPE_OK && profile) { rc = phases.alloc(4); a.phase_cycles = phases.p; }
  a.corr_match = match.p; a.corr_d2 = d2.p; a.partials = partials.p; a.result = dres.p;
  if (rc == OPE_OK) {
    void* args[] = {(void*)&a};
    cudaEventRecord(ctx->kev[0][0], ctx->stream);
    cudaError_t e = cudaLaunchCooperativeKernel((void*)icp_kernel, dim3(blocks), dim3(kIcpThreads), args, 0, ctx->stream);
    cudaEventRecord(ctx->kev[0][1], ctx->stream);
    ctx->kev_valid[0] = true;
    if (e != cudaSuccess) rc = fail(ctx, OPE_ERR_CUDA, "cooperative launch of icp_kernel failed: %s", cudaGetErrorString(e));
    else rc = check_launch(ctx, "icp_kernel");
  }
  if (rc == OPE_OK) {
    void* h;
    rc = read_back(ctx, dres.p, sizeof(ope_reg_result), &h);
    if (rc == OPE_OK) std::memcpy(res, h, sizeof(ope_reg_result));
  }
  if (rc == OPE_OK && profile) {
    void* h;
    if (read_back(ctx, phases.p, 4 * sizeof(long long), &h) == OPE_OK) {
      const long long* c = (const long long*)h;
      fprintf(stderr, "[ope profile] icp_kernel blocks=%d block0 cycles: search+reduce %lld | grid barrier %lld | partials+svd %lld | transform %lld (per iteration: %.0f %.0f %.0f %.0f)\n",
              blocks, c[0], c[1], c[2], c[3], (double)c[0] / std::max(res->iterations, 1), (double)c[1] / std::max(res->iterations, 1),
              (double)c[2] / std::max(res->iterations, 1), (double)c[3] / std::max(res->iterations, 1));
    }
  }
  if (rc == OPE_OK && out_corr_host && n > 0) {
    std::vector<int> hm(n);
    std::vector<float> hd(n);
    cudaError_t e = cudaMemcpyAsync(hm.data(), match.p, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hd.data(), d2.p, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = fail(ctx, OPE_ERR_CUDA, "correspondence download failed: %s", cudaGetErrorString(e));
    else corr_to_host(hm, hd, out_corr_host, nullptr);
  }
  if (rc == OPE_OK && out_aligned) {
    // output = *input_ transformed by the final transformation (VP/impl/icp_mod.hpp:269-271)
    Mat4 F;
    std::memcpy(F.m, res->T, sizeof(F.m));
    rc = transform_device(ctx, src, F, work);
    if (rc == OPE_OK) { *out_aligned = work; work = nullptr; }
  }
  if (work) ope_cloud_free(ctx, work);
  return rc;
}

static inline int libc_random_index(int n) { return (int)(n * (rand() / (RAND_MAX + 1.0))); }

// selectSamples [UPSTREAM ia_ransac.hpp] on the host: it consumes libc rand() serially (SURVEY hard part 3)
static int draw_samples(const float* src, size_t ns, size_t stride_f, int nr_samples, float& min_sample_distance, int32_t* out) {
  if (nr_samples > (int)ns) return OPE_ERR_INVALID;
  int without = 0;
  const int max_without = (int)(3 * ns);
  int cnt = 0;
  while (cnt < nr_samples) {
    const int si = libc_random_index((int)ns);
    bool valid = true;
    for (int i = 0; i < cnt; ++i) {
      const float* a = src + (size_t)si * stride_f;
      const float* b = src + (size_t)out[i] * stride_f;
      const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
      volatile float s = dx * dx;
      s = s + dy * dy;
      s = s + dz * dz;
      const float dist = std::sqrt((float)s);
      if (si == out[i] || dist < min_sample_distance) { valid = false; break; }
    }
    if (valid) { out[cnt++] = si; without = 0; }
    else ++without;
    if (without >= max_without) { min_sample_distance *= 0.5f; without = 0; }
  }
  return OPE_OK;
}

int sacia_device(ope_ctx* ctx, const ope_cloud* src, const float* d_fsrc, const ope_cloud* tgt, const float* d_ftgt,
                 const ope_sacia_params& prm, const ope_rng_table* table, const float* host_src_xyz3, ope_reg_result* res,
                 float* out_errors_host) {
  const int H = prm.max_iterations, S = prm.nr_samples, K = prm.k_correspondences;
  if (H < 1 || S < 1 || S > kSaciaMaxSamples || K < 1 || K > 16) return fail(ctx, OPE_ERR_INVALID, "bad SAC-IA parameters");
  if (src->n == 0 || tgt->n == 0) return fail(ctx, OPE_ERR_EMPTY, "empty cloud");
  if ((size_t)S > src->n) return fail(ctx, OPE_ERR_INVALID, "The number of samples must not be greater than the number of points");
  int h0 = 0, h1 = H;
  if (prm.hypothesis_end > prm.hypothesis_begin) { h0 = std::max(0, prm.hypothesis_begin); h1 = std::min(H, prm.hypothesis_end); }
  // ---- decision table: replayed or drawn from libc rand() ----
  std::vector<int32_t> hs, hp;
  const int32_t *samples = nullptr, *picks = nullptr;
  if (table) {
    if (table->n_hypotheses < H || table->nr_samples != S) return fail(ctx, OPE_ERR_INVALID, "rng table shape mismatch");
    samples = table->samples; picks = table->picks;
  } else {
    std::vector<float> xyz;
    const float* hx = host_src_xyz3;
    if (!hx) {
      xyz.resize(3 * src->n);
      OPE_TRY(ope_cloud_download(ctx, src, xyz.data(), nullptr));
      hx = xyz.data();
    }
    hs.resize((size_t)H * S); hp.resize((size_t)H * S);
    float msd = prm.min_sample_distance;
    for (int it = 0; it < H; ++it) {
      int rc = draw_samples(hx, src->n, 3, S, msd, hs.data() + (size_t)it * S);
      if (rc != OPE_OK) return fail(ctx, rc, "selectSamples failed");
      for (int s = 0; s < S; ++s) hp[(size_t)it * S + s] = libc_random_index(K);
    }
    samples = hs.data(); picks = hp.data();
  }
  for (size_t i = 0; i < (size_t)H * S; ++i)
    if (samples[i] < 0 || (size_t)samples[i] >= src->n || picks[i] < 0 || picks[i] >= K)
      return fail(ctx, OPE_ERR_INVALID, "rng table entry out of range");
  // ---- device side ----
  const size_t ns = src->n;
  Scratch<int> d_samples(ctx), d_picks(ctx), d_knn(ctx);
  Scratch<float> d_terms(ctx), d_errors(ctx), d_T(ctx), d_knn_d2(ctx);
  Scratch<ope_reg_result> d_res(ctx);
  OPE_TRY(d_samples.alloc((size_t)H * S)); OPE_TRY(d_picks.alloc((size_t)H * S));
  OPE_TRY(d_knn.alloc(ns * K)); OPE_TRY(d_knn_d2.alloc(ns * K));
  const int nh = h1 - h0;
  OPE_TRY(d_terms.alloc((size_t)std::max(nh, 1) * ns));
  OPE_TRY(d_errors.alloc(H)); OPE_TRY(d_T.alloc((size_t)H * 16)); OPE_TRY(d_res.alloc(1));
  OPE_CUDA_TRY(ctx, cudaMemcpyAsync(d_samples.p, samples, (size_t)H * S * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  OPE_CUDA_TRY(ctx, cudaMemcpyAsync(d_picks.p, picks, (size_t)H * S * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  OPE_CUDA_TRY(ctx, cudaMemsetAsync(d_errors.p, 0xff, (size_t)H * sizeof(float), ctx->stream));  // NaN = not evaluated
  // findSimilarFeatures for every source point at once (K6)
  OPE_TRY(feature_knn_device(ctx, d_ftgt, tgt->n, d_fsrc, ns, 33, K, d_knn.p, d_knn_d2.p));
  SaciaDev a;
  OPE_TRY(cloud_bbox(ctx, const_cast<ope_cloud*>(tgt)));
  OPE_TRY(cloud_grid(ctx, tgt, knn_cell_size(tgt, 1), &a.grid));
  a.src = src->pts; a.tgt = tgt->pts; a.ns = (int)ns; a.nr_samples = S; a.k_corr = K;
  a.samples = d_samples.p; a.picks = d_picks.p; a.knn_idx = d_knn.p; a.h_begin = h0;
  a.threshold = (float)prm.max_correspondence_distance;
  a.terms = d_terms.p; a.errors = d_errors.p; a.transforms = d_T.p;
  if (nh > 0) {
    cudaEventRecord(ctx->kev[1][0], ctx->stream);
    sacia_kernel<<<nh, kSaciaThreads, 0, ctx->stream>>>(a);
    cudaEventRecord(ctx->kev[1][1], ctx->stream);
    ctx->kev_valid[1] = true;
    OPE_TRY(check_launch(ctx, "sacia_kernel"));
  }
  sacia_select_kernel<<<1, 32, 0, ctx->stream>>>(d_errors.p, d_T.p, h0, h1, d_res.p);
  OPE_TRY(check_launch(ctx, "sacia_select_kernel"));
  void* h;
  OPE_TRY(read_back(ctx, d_res.p, sizeof(ope_reg_result), &h));
  std::memcpy(res, h, sizeof(ope_reg_result));
  res->iterations = H;
  if (out_errors_host) {
    OPE_CUDA_TRY(ctx, cudaMemcpyAsync(out_errors_host, d_errors.p, (size_t)H * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    OPE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return OPE_OK;
}

}  // namespace ope

// =========================================================================================== C ABI =====
using namespace ope;

extern "C" {

void ope_icp_params_default(ope_icp_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->max_iterations = 10;                                // VP/registration_mod.h:102-130
  p->transformation_epsilon = 0.0;
  p->euclidean_fitness_epsilon = -DBL_MAX;
  p->max_correspondence_distance = std::sqrt(DBL_MAX);
  p->min_number_correspondences = 3;
  p->estimator = OPE_EST_NEAREST;
  p->k_search = 10;
  p->transformation = OPE_TE_SVD;
  p->variant = OPE_ICP_VARIANT_MOD;
  p->mse_threshold_absolute = 1e-12;                     // VP/default_convergence_criteria_mod.h:94-110
}
void ope_sacia_params_default(ope_sacia_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->max_iterations = 10; p->nr_samples = 3; p->k_correspondences = 10; p->min_sample_distance = 0.0f;
  p->max_correspondence_distance = std::sqrt(DBL_MAX);
}
void ope_pose_params_default(ope_pose_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->coarse_leaf = 0.01f; p->fine_leaf = 0.008f; p->normal_k = 30; p->fpfh_radius = 0.03f;
  ope_sacia_params_default(&p->sacia);
  p->sacia.max_iterations = 400; p->sacia.nr_samples = 5; p->sacia.k_correspondences = 5;
  p->sacia.min_sample_distance = 0.01f; p->sacia.max_correspondence_distance = 0.05;
  p->min_target_features = 10; p->min_target_points = 100;
  ope_icp_params_default(&p->icp);
  p->icp.max_iterations = 100; p->icp.transformation_epsilon = 1e-8; p->icp.euclidean_fitness_epsilon = 1e-8;
  p->icp.estimator = OPE_EST_NORMAL_SHOOTING; p->icp.k_search = 20;
  p->icp.n_rejectors = 2;
  p->icp.rejector_kind[0] = OPE_REJ_SURFACE_NORMAL; p->icp.rejector_threshold[0] = 0.7;
  p->icp.rejector_kind[1] = OPE_REJ_SELF_OCCLUDED_NORMAL; p->icp.rejector_threshold[1] = 0.6;
  p->icp.transformation = OPE_TE_SVD; p->icp.with_normals = 1;
  p->coarse_refit_threshold = 1e-4;
}

int ope_cloud_transform(ope_ctx* ctx, const ope_cloud* cloud, const float T[16], ope_cloud** out) {
  if (!ctx || !cloud || !T || !out) return OPE_ERR_INVALID;
  ope_cloud* o = nullptr;
  OPE_TRY(cloud_alloc(ctx, cloud->n, cloud->normals != nullptr, &o));
  Mat4 M;
  std::memcpy(M.m, T, sizeof(M.m));
  int rc = transform_device(ctx, cloud, M, o);
  if (rc != OPE_OK) { ope_cloud_free(ctx, o); return rc; }
  OPE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  *out = o;
  return OPE_OK;
}

int ope_umeyama(ope_ctx* ctx, const ope_cloud* src, const ope_cloud* tgt, const int32_t* isrc, const int32_t* itgt, size_t n,
                float T[16]) {
  if (!ctx || !src || !tgt || !T || n == 0) return OPE_ERR_INVALID;
  if ((!isrc && n > src->n) || (!itgt && n > tgt->n)) return fail(ctx, OPE_ERR_INVALID, "n exceeds cloud size");
  Scratch<int> ds(ctx), dt(ctx);
  if (isrc) {
    for (size_t i = 0; i < n; ++i) if (isrc[i] < 0 || (size_t)isrc[i] >= src->n) return fail(ctx, OPE_ERR_INVALID, "source index out of range");
    OPE_TRY(ds.alloc(n));
    OPE_CUDA_TRY(ctx, cudaMemcpyAsync(ds.p, isrc, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  }
  if (itgt) {
    for (size_t i = 0; i < n; ++i) if (itgt[i] < 0 || (size_t)itgt[i] >= tgt->n) return fail(ctx, OPE_ERR_INVALID, "target index out of range");
    OPE_TRY(dt.alloc(n));
    OPE_CUDA_TRY(ctx, cudaMemcpyAsync(dt.p, itgt, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  }
  return umeyama_device(ctx, src->pts, tgt->pts, isrc ? ds.p : nullptr, itgt ? dt.p : nullptr, n, T);
}

int ope_fitness(ope_ctx* ctx, const ope_cloud* src, const ope_cloud* tgt, const float T[16], double max_range, double* out) {
  if (!ctx || !src || !tgt || !T || !out) return OPE_ERR_INVALID;
  if (tgt->n == 0) return fail(ctx, OPE_ERR_EMPTY, "No input target dataset was given!");
  Mat4 M;
  std::memcpy(M.m, T, sizeof(M.m));
  return fitness_device(ctx, src, tgt, M, max_range, out);
}

int ope_correspondences(ope_ctx* ctx, const ope_cloud* src, const ope_cloud* tgt, const ope_icp_params* prm,
                        ope_correspondence* out, size_t* out_n) {
  if (!ctx || !src || !tgt || !prm || !out || !out_n) return OPE_ERR_INVALID;
  *out_n = 0;
  if (tgt->n == 0) return fail(ctx, OPE_ERR_EMPTY, "No input target dataset was given!");
  if (src->n == 0) return OPE_OK;
  IcpDev a;
  std::memset(&a, 0, sizeof(a));
  OPE_TRY(icp_fill(ctx, src, tgt, *prm, &a));
  a.variant = OPE_ICP_VARIANT_MOD;
  a.cur_pts = src->pts; a.cur_nrm = src->normals;
  const size_t n = src->n;
  Scratch<int> match(ctx);
  Scratch<float> d2(ctx);
  OPE_TRY(match.alloc(n)); OPE_TRY(d2.alloc(n));
  a.corr_match = match.p; a.corr_d2 = d2.p;
  correspond_once_kernel<<<(unsigned)std::min<size_t>(div_up(n * 8, kIcpThreads), (size_t)ctx->sm_count * 8), kIcpThreads, 0, ctx->stream>>>(a);
  OPE_TRY(check_launch(ctx, "correspond_once_kernel"));
  std::vector<int> hm(n);
  std::vector<float> hd(n);
  OPE_CUDA_TRY(ctx, cudaMemcpyAsync(hm.data(), match.p, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  OPE_CUDA_TRY(ctx, cudaMemcpyAsync(hd.data(), d2.p, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  OPE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  corr_to_host(hm, hd, out, out_n);
  return OPE_OK;
}

int ope_icp_align(ope_ctx* ctx, const ope_cloud* src, const ope_cloud* tgt, const ope_icp_params* prm, const float guess[16],
                  ope_reg_result* res, ope_correspondence* out_corr, ope_cloud** out_aligned) {
  if (!ctx || !src || !prm || !res) return OPE_ERR_INVALID;
  Mat4 I = mat4_identity();
  std::memset(res, 0, sizeof(*res));
  std::memcpy(res->T, I.m, sizeof(I.m));
  if (out_aligned) *out_aligned = nullptr;
  // Registration::initCompute: no target -> PCL_ERROR + return, transforms stay identity (VP/impl/registration_mod.hpp:73-77)
  if (!tgt || tgt->n == 0) return fail(ctx, OPE_ERR_EMPTY, "No input target dataset was given!");
  Mat4 G = I;
  if (guess) std::memcpy(G.m, guess, sizeof(G.m));
  return icp_device(ctx, src, tgt, *prm, G, res, out_corr, out_aligned);
}

int ope_sacia_draw(const float* src_xyz, size_t ns, size_t stride_bytes, int iterations, int nr_samples, int k_correspondences,
                   float* min_sample_distance, int32_t* samples, int32_t* picks) {
  if (!src_xyz || !samples || !picks || !min_sample_distance || stride_bytes % 4 != 0 || stride_bytes < 12) return OPE_ERR_INVALID;
  for (int it = 0; it < iterations; ++it) {
    int rc = draw_samples(src_xyz, ns, stride_bytes / 4, nr_samples, *min_sample_distance, samples + (size_t)it * nr_samples);
    if (rc != OPE_OK) return rc;
    for (int s = 0; s < nr_samples; ++s) picks[(size_t)it * nr_samples + s] = libc_random_index(k_correspondences);
  }
  return OPE_OK;
}

int ope_sacia_align(ope_ctx* ctx, const ope_cloud* src, const float* fsrc, const ope_cloud* tgt, const float* ftgt,
                    const ope_sacia_params* prm, const ope_rng_table* table, ope_reg_result* res, float* out_errors) {
  if (!ctx || !src || !fsrc || !tgt || !ftgt || !prm || !res) return OPE_ERR_INVALID;
  Mat4 I = mat4_identity();
  std::memset(res, 0, sizeof(*res));
  std::memcpy(res->T, I.m, sizeof(I.m));
  Scratch<float> dfs(ctx), dft(ctx);
  OPE_TRY(dfs.alloc(src->n * 33)); OPE_TRY(dft.alloc(tgt->n * 33));
  if (src->n) OPE_CUDA_TRY(ctx, cudaMemcpyAsync(dfs.p, fsrc, src->n * 33 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  if (tgt->n) OPE_CUDA_TRY(ctx, cudaMemcpyAsync(dft.p, ftgt, tgt->n * 33 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  return sacia_device(ctx, src, dfs.p, tgt, dft.p, *prm, table, nullptr, res, out_errors);
}

}  // extern "C"
