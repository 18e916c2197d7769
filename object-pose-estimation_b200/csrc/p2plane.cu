// p2plane.cu — point-to-plane transformation estimation on the device:
//   TransformationEstimationPointToPlaneLLS  [UPSTREAM transformation_estimation_point_to_plane_lls.hpp]
//       the default estimator of IterativeClosestPointWithNormals (VP/icp_mod.h:352-357): 6x6 normal equations, one pass;
//   TransformationEstimationPointToPlane     [UPSTREAM transformation_estimation_point_to_plane.h, ..._lm.hpp]
//       what BuildModel plugs in (BM/src/regmeshpcd.cpp:162,193): Eigen's Levenberg-Marquardt (MINPACK lmdif) over the
//       6-parameter rigid warp (WarpPointRigid6D) with a forward-difference Jacobian, residual (warp(s) - t) . n_t.
//
// Device side: ONE fused pass per Jacobian — every correspondence evaluates its residual under the 7 warps (x and x + h e_j)
// in float exactly as the reference's functor would, forms its Jacobian row (f(x + h e_j) - f(x)) / h in float, and the 21
// entries of J^T J, the 6 of J^T f and f^T f are accumulated in double (warp shuffles -> block -> one final warp per
// accumulator). The m x 6 Jacobian is never stored: 28 doubles come back to the host. A trial point costs one pass that
// returns f^T f. Host side: the lmdif driver works from those normal equations — the column-pivoted R of J = Q R is the pivoted
// Cholesky factor of J^T J and Q^T f = R^-T P^T J^T f — which is algebraically what Eigen's ColPivHouseholderQR hands to
// lmpar/qrsolv; with double accumulation the squared condition number is harmless for a 6x6 system.
// HBM-bound in principle (48 B per correspondence and pass), latency/launch bound at the sizes the apps use (1e3..1e5 pairs).
#include <cfloat>
#include <cmath>
#include <vector>

#include "ope_host.cuh"

namespace ope {

namespace {

constexpr int kP2pThreads = 256;
constexpr int kAccLls = 29;   // ATA upper triangle (21), ATb (6), pairs, sum of correspondence distances
constexpr int kAccJac = 30;   // J^T J upper triangle (21), J^T f (6), f^T f, pairs, sum of correspondence distances
constexpr int kAccF = 3;      // f^T f, pairs, sum of correspondence distances

struct P2pArgs {
  const float4* src;     // source points (by source index)
  const float4* tgt;     // target points
  const float4* tgt_n;   // target normals
  const int* is;         // n source indices or null (identity)
  const int* it;         // n target indices; < 0: no correspondence for this entry
  const float* d2;       // n correspondence distances or null
  int n;
  Mat4 W[7];             // warps: [0] at x, [1 + j] at x + h[j] e_j
  float h[6];
};

__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NACC>
__device__ __forceinline__ void block_sum_store(double* acc, double* dst) {
  __shared__ double sm[(kP2pThreads / 32) * NACC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < NACC; ++a) acc[a] = warp_sum(acc[a]);
  if (lane == 0)
    for (int a = 0; a < NACC; ++a) sm[warp * NACC + a] = acc[a];
  __syncthreads();
  if (threadIdx.x < NACC) {
    double s = 0.0;
    for (int w = 0; w < kP2pThreads / 32; ++w) s += sm[w * NACC + threadIdx.x];
    dst[threadIdx.x] = s;
  }
}

// (warp(s) - t) . n as the reference's functor evaluates it: Vector4f with w = 0, Eigen's packet reduction with SSE3+
// horizontal adds, (p0 + p1) + (p2 + p3) — the one Eigen model of the whole path
__device__ __forceinline__ float plane_residual(const Mat4& W, const float4 s, const float4 t, const float4 n) {
  float wx, wy, wz;
  xform_point(W, s.x, s.y, s.z, wx, wy, wz);
  const float p0 = (wx - t.x) * n.x, p1 = (wy - t.y) * n.y, p2 = (wz - t.z) * n.z;
  return (p0 + p1) + (p2 + 0.0f);
}

// MODE 0: LLS normal equations; 1: LM residual norm at W[0]; 2: LM Jacobian normal equations at W[0..6]
template <int MODE>
__global__ void __launch_bounds__(kP2pThreads) p2p_accum_kernel(const P2pArgs a, double* __restrict__ partials) {
  constexpr int NACC = MODE == 0 ? kAccLls : (MODE == 1 ? kAccF : kAccJac);
  double acc[NACC];
#pragma unroll
  for (int e = 0; e < NACC; ++e) acc[e] = 0.0;
  for (int i = blockIdx.x * kP2pThreads + threadIdx.x; i < a.n; i += gridDim.x * kP2pThreads) {
    const int m = __ldg(a.it + i);
    if (m < 0) continue;
    const int si = a.is ? __ldg(a.is + i) : i;
    const float4 s = __ldg(a.src + si), t = __ldg(a.tgt + m), n = __ldg(a.tgt_n + m);
    acc[NACC - 2] += 1.0;
    if (a.d2) acc[NACC - 1] += (double)__ldg(a.d2 + i);
    if (MODE == 0) {
      if (!finite3(s.x, s.y, s.z) || !finite3(t.x, t.y, t.z) || !finite3(n.x, n.y, n.z)) continue;
      double v[6];
      v[0] = (double)(n.z * s.y - n.y * s.z);
      v[1] = (double)(n.x * s.z - n.z * s.x);
      v[2] = (double)(n.y * s.x - n.x * s.y);
      v[3] = (double)n.x; v[4] = (double)n.y; v[5] = (double)n.z;
      const double d = (double)(n.x * t.x + n.y * t.y + n.z * t.z - n.x * s.x - n.y * s.y - n.z * s.z);
      int e = 0;
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = r; c < 6; ++c) acc[e++] += v[r] * v[c];
#pragma unroll
      for (int r = 0; r < 6; ++r) acc[21 + r] += v[r] * d;
    } else if (MODE == 1) {
      const float f = plane_residual(a.W[0], s, t, n);
      acc[0] += (double)f * (double)f;
    } else {
      const float f = plane_residual(a.W[0], s, t, n);
      double J[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) J[j] = (double)((plane_residual(a.W[1 + j], s, t, n) - f) / a.h[j]);
      int e = 0;
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = r; c < 6; ++c) acc[e++] += J[r] * J[c];
#pragma unroll
      for (int r = 0; r < 6; ++r) acc[21 + r] += J[r] * (double)f;
      acc[27] += (double)f * (double)f;
    }
  }
  block_sum_store<NACC>(acc, partials + (size_t)blockIdx.x * NACC);
}

// one warp per accumulator, lanes stride over the blocks, fixed shuffle tree
__global__ void p2p_final_kernel(const double* __restrict__ partials, int nblocks, int nacc, double* __restrict__ out) {
  const int a = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (a >= nacc) return;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += partials[(size_t)b * nacc + a];
  s = warp_sum(s);
  if (lane == 0) out[a] = s;
}

// WarpPointRigid6D::setParam [UPSTREAM warp_point_rigid_6d.h], float: translation, then q = (sqrt(1 - |v|^2), v) normalised
// and Eigen's Quaternion::toRotationMatrix
Mat4 warp_rigid_6d(const float x[6]) {
  Mat4 M;
  for (int i = 0; i < 16; ++i) M.m[i] = 0.0f;
  M(0, 3) = x[0]; M(1, 3) = x[1]; M(2, 3) = x[2]; M(3, 3) = 1.0f;
  volatile float qq = 0.0f * 0.0f + x[3] * x[3];
  qq = qq + x[4] * x[4];
  qq = qq + x[5] * x[5];
  float qw = std::sqrt(1.0f - qq);
  volatile float nn = x[3] * x[3] + x[4] * x[4];
  nn = nn + x[5] * x[5];
  nn = nn + qw * qw;
  const float nrm = std::sqrt((float)nn);
  const float qx = x[3] / nrm, qy = x[4] / nrm, qz = x[5] / nrm;
  qw = qw / nrm;
  const float tx = 2.0f * qx, ty = 2.0f * qy, tz = 2.0f * qz;
  const float twx = tx * qw, twy = ty * qw, twz = tz * qw;
  const float txx = tx * qx, txy = ty * qx, txz = tz * qx;
  const float tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  M(0, 0) = 1.0f - (tyy + tzz); M(0, 1) = txy - twz; M(0, 2) = txz + twy;
  M(1, 0) = txy + twz; M(1, 1) = 1.0f - (txx + tzz); M(1, 2) = tyz - twx;
  M(2, 0) = txz - twy; M(2, 1) = tyz + twx; M(2, 2) = 1.0f - (txx + tyy);
  return M;
}

struct Accum {
  ope_ctx* ctx;
  P2pArgs a;
  Scratch<double> partials, out;
  int blocks = 1;
  explicit Accum(ope_ctx* c) : ctx(c), partials(c), out(c) {}
  int init() {
    blocks = (int)std::max<unsigned>(1u, std::min<unsigned>(div_up((size_t)std::max(a.n, 1), kP2pThreads), (unsigned)ctx->sm_count * 4u));
    OPE_TRY(partials.alloc((size_t)blocks * kAccJac));
    return out.alloc(kAccJac);
  }
  template <int MODE>
  int run(int nacc, const double** host) {
    p2p_accum_kernel<MODE><<<blocks, kP2pThreads, 0, ctx->stream>>>(a, partials.p);
    OPE_TRY(check_launch(ctx, "p2p_accum_kernel"));
    p2p_final_kernel<<<1, 32 * nacc, 0, ctx->stream>>>(partials.p, blocks, nacc, out.p);
    OPE_TRY(check_launch(ctx, "p2p_final_kernel"));
    void* h;
    OPE_TRY(read_back(ctx, out.p, (size_t)nacc * sizeof(double), &h));
    *host = (const double*)h;
    return OPE_OK;
  }
};

// ---- 6x6 linear algebra of the LM driver (double) ---------------------------------------------------------
constexpr int N6 = 6;
inline double norm6(const double* v) { double s = 0; for (int i = 0; i < N6; ++i) s += v[i] * v[i]; return std::sqrt(s); }

// P^T (J^T J) P = R^T R with the pivot rule of a column-pivoted QR (largest remaining column norm first, first on ties);
// qtf = R^-T P^T (J^T f). Returns the numerical rank by Eigen's ColPivHouseholderQR::rank() threshold (eps_float * 6 * |R00|).
int pivoted_factor(const double A[36], const double g[6], double R[36], int ipvt[6], double qtf[6]) {
  double W[36];
  for (int i = 0; i < 36; ++i) { W[i] = A[i]; R[i] = 0.0; }
  for (int j = 0; j < N6; ++j) ipvt[j] = j;
  for (int k = 0; k < N6; ++k) {
    int best = k;
    for (int j = k + 1; j < N6; ++j) if (W[j * N6 + j] > W[best * N6 + best]) best = j;
    if (best != k) {
      for (int i = 0; i < N6; ++i) std::swap(W[i * N6 + k], W[i * N6 + best]);
      for (int i = 0; i < N6; ++i) std::swap(W[k * N6 + i], W[best * N6 + i]);
      for (int i = 0; i < k; ++i) std::swap(R[i * N6 + k], R[i * N6 + best]);
      std::swap(ipvt[k], ipvt[best]);
    }
    const double piv = W[k * N6 + k];
    if (!(piv > 0.0)) break;   // the remaining columns are (numerically) in the span: their R rows stay zero
    const double rkk = std::sqrt(piv);
    R[k * N6 + k] = rkk;
    for (int j = k + 1; j < N6; ++j) R[k * N6 + j] = W[k * N6 + j] / rkk;
    for (int i = k + 1; i < N6; ++i)
      for (int j = k + 1; j < N6; ++j) W[i * N6 + j] -= R[k * N6 + i] * R[k * N6 + j];
  }
  int rank = 0;
  const double thr = std::fabs(R[0]) * (double)FLT_EPSILON * N6;
  for (int j = 0; j < N6; ++j) if (std::fabs(R[j * N6 + j]) > thr) ++rank;
  for (int j = 0; j < N6; ++j) {   // R^T y = P^T g
    if (R[j * N6 + j] == 0.0) { qtf[j] = 0.0; continue; }
    double s = g[ipvt[j]];
    for (int i = 0; i < j; ++i) s -= R[i * N6 + j] * qtf[i];
    qtf[j] = s / R[j * N6 + j];
  }
  return rank;
}

// MINPACK qrsolv: min ||R P^T x - qtb||^2 + ||D x||^2 by Givens rotations; S is left in the strict lower part of r and sdiag
void qrsolv6(double* r, const int* ipvt, const double* diag, const double* qtb, double* x, double* sdiag) {
  double wa[N6];
  for (int j = 0; j < N6; ++j) {
    for (int i = j; i < N6; ++i) r[i * N6 + j] = r[j * N6 + i];
    x[j] = r[j * N6 + j];
    wa[j] = qtb[j];
  }
  for (int j = 0; j < N6; ++j) {
    const int l = ipvt[j];
    if (diag[l] != 0.0) {
      for (int k = j; k < N6; ++k) sdiag[k] = 0.0;
      sdiag[j] = diag[l];
      double qtbpj = 0.0;
      for (int k = j; k < N6; ++k) {
        if (sdiag[k] == 0.0) continue;
        double c, s;
        const double rkk = r[k * N6 + k];
        if (std::fabs(rkk) < std::fabs(sdiag[k])) {
          const double ct = rkk / sdiag[k];
          s = 0.5 / std::sqrt(0.25 + 0.25 * ct * ct);
          c = s * ct;
        } else {
          const double tg = sdiag[k] / rkk;
          c = 0.5 / std::sqrt(0.25 + 0.25 * tg * tg);
          s = c * tg;
        }
        r[k * N6 + k] = c * rkk + s * sdiag[k];
        const double t = c * wa[k] + s * qtbpj;
        qtbpj = -s * wa[k] + c * qtbpj;
        wa[k] = t;
        for (int i = k + 1; i < N6; ++i) {
          const double u = c * r[i * N6 + k] + s * sdiag[i];
          sdiag[i] = -s * r[i * N6 + k] + c * sdiag[i];
          r[i * N6 + k] = u;
        }
      }
    }
    sdiag[j] = r[j * N6 + j];
    r[j * N6 + j] = x[j];
  }
  int nsing = N6;
  for (int j = 0; j < N6; ++j) {
    if (sdiag[j] == 0.0 && nsing == N6) nsing = j;
    if (nsing < N6) wa[j] = 0.0;
  }
  for (int j = nsing - 1; j >= 0; --j) {
    double sum = 0.0;
    for (int i = j + 1; i < nsing; ++i) sum += r[i * N6 + j] * wa[i];
    wa[j] = (wa[j] - sum) / sdiag[j];
  }
  for (int j = 0; j < N6; ++j) x[ipvt[j]] = wa[j];
}

// Eigen lmpar2 / MINPACK lmpar: the damping parameter for the trust region ||D p|| <= delta and the step p
void lmpar6(const double* R, int rank, const int* ipvt, const double* diag, const double* qtb, double delta, double& par, double* x) {
  const double dwarf = (double)FLT_MIN;
  double wa1[N6], wa2[N6], s[36], sdiag[N6];
  for (int j = 0; j < N6; ++j) wa1[j] = j < rank ? qtb[j] : 0.0;
  for (int j = rank - 1; j >= 0; --j) {
    double t = wa1[j];
    for (int i = j + 1; i < rank; ++i) t -= R[j * N6 + i] * wa1[i];
    wa1[j] = t / R[j * N6 + j];
  }
  for (int j = 0; j < N6; ++j) x[ipvt[j]] = wa1[j];
  int iter = 0;
  for (int j = 0; j < N6; ++j) wa2[j] = diag[j] * x[j];
  double dxnorm = norm6(wa2);
  double fp = dxnorm - delta;
  if (fp <= 0.1 * delta) { par = 0.0; return; }
  double parl = 0.0;
  if (rank == N6) {
    for (int j = 0; j < N6; ++j) { const int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
    for (int j = 0; j < N6; ++j) {
      double sum = 0.0;
      for (int i = 0; i < j; ++i) sum += R[i * N6 + j] * wa1[i];
      wa1[j] = (wa1[j] - sum) / R[j * N6 + j];
    }
    const double t = norm6(wa1);
    parl = fp / delta / t / t;
  }
  for (int j = 0; j < N6; ++j) {
    double sum = 0.0;
    for (int i = 0; i <= j; ++i) sum += R[i * N6 + j] * qtb[i];
    wa1[j] = sum / diag[ipvt[j]];
  }
  const double gnorm = norm6(wa1);
  double paru = gnorm / delta;
  if (paru == 0.0) paru = dwarf / std::min(delta, 0.1);
  par = std::min(std::max(par, parl), paru);
  if (par == 0.0) par = gnorm / dxnorm;
  for (;;) {
    ++iter;
    if (par == 0.0) par = std::max(dwarf, 0.001 * paru);
    const double sq = std::sqrt(par);
    for (int j = 0; j < N6; ++j) wa1[j] = sq * diag[j];
    for (int i = 0; i < 36; ++i) s[i] = R[i];
    qrsolv6(s, ipvt, wa1, qtb, x, sdiag);
    for (int j = 0; j < N6; ++j) wa2[j] = diag[j] * x[j];
    dxnorm = norm6(wa2);
    const double prev = fp;
    fp = dxnorm - delta;
    if (std::fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= prev && prev < 0.0) || iter == 10) break;
    for (int j = 0; j < N6; ++j) { const int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
    for (int j = 0; j < N6; ++j) {
      wa1[j] /= sdiag[j];
      const double t = wa1[j];
      for (int i = j + 1; i < N6; ++i) wa1[i] -= s[i * N6 + j] * t;
    }
    const double t = norm6(wa1);
    const double parc = fp / delta / t / t;
    if (fp > 0.0) parl = std::max(parl, par);
    if (fp < 0.0) paru = std::min(paru, par);
    par = std::max(parl, par + parc);
  }
}

// Eigen::LevenbergMarquardt::minimize with NumericalDiff<Forward> (defaults: factor 100, maxfev 400, ftol = xtol =
// sqrt(eps_float), gtol 0), the residuals living on the device
int lm_minimize(Accum& acc, float x[6], int32_t info[3]) {
  const double factor = 100.0, ftol = (double)std::sqrt(FLT_EPSILON), xtol = ftol, gtol = 0.0, eps_mach = (double)FLT_EPSILON;
  const int maxfev = 400;
  const float eps = std::sqrt(FLT_EPSILON);
  int nfev = 1, iter = 1, status = -1;
  double fnorm = 0.0, par = 0.0, delta = 0.0, xnorm = 0.0, gnorm = 0.0;
  double diag[N6], colnorm[N6], R[36], qtf[N6], wa1[N6], wa3[N6], A[36], g[N6];
  int ipvt[N6];
  float xt[N6];
  bool have_fnorm = false;
  for (;;) {
    // Jacobian by forward differences: f(x) again plus one evaluation per parameter, fused into one pass
    acc.a.W[0] = warp_rigid_6d(x);
    for (int j = 0; j < N6; ++j) {
      for (int k = 0; k < N6; ++k) xt[k] = x[k];
      float h = eps * std::fabs(x[j]);
      if (h == 0.0f) h = eps;
      volatile float xj = x[j] + h;
      xt[j] = xj;
      acc.a.h[j] = h;
      acc.a.W[1 + j] = warp_rigid_6d(xt);
    }
    const double* r;
    OPE_TRY(acc.run<2>(kAccJac, &r));
    nfev += N6 + 1;
    {
      int e = 0;
      for (int i = 0; i < N6; ++i) for (int j = i; j < N6; ++j) { A[i * N6 + j] = A[j * N6 + i] = r[e++]; }
      for (int i = 0; i < N6; ++i) g[i] = r[21 + i];
      if (!have_fnorm) { fnorm = std::sqrt(r[27]); have_fnorm = true; }   // minimizeInit's evaluation of f(x0)
    }
    for (int j = 0; j < N6; ++j) colnorm[j] = std::sqrt(A[j * N6 + j]);
    const int rank = pivoted_factor(A, g, R, ipvt, qtf);
    if (iter == 1) {
      for (int j = 0; j < N6; ++j) diag[j] = colnorm[j] == 0.0 ? 1.0 : colnorm[j];
      for (int j = 0; j < N6; ++j) wa3[j] = diag[j] * (double)x[j];
      xnorm = norm6(wa3);
      delta = factor * xnorm;
      if (delta == 0.0) delta = factor;
    }
    gnorm = 0.0;
    if (fnorm != 0.0)
      for (int j = 0; j < N6; ++j)
        if (colnorm[ipvt[j]] != 0.0) {
          double s = 0.0;
          for (int i = 0; i <= j; ++i) s += R[i * N6 + j] * (qtf[i] / fnorm);
          gnorm = std::max(gnorm, std::fabs(s / colnorm[ipvt[j]]));
        }
    if (gnorm <= gtol) { status = 4; break; }
    for (int j = 0; j < N6; ++j) diag[j] = std::max(diag[j], colnorm[j]);
    double ratio = 0.0;
    do {
      lmpar6(R, rank, ipvt, diag, qtf, delta, par, wa1);
      for (int j = 0; j < N6; ++j) { wa1[j] = -wa1[j]; xt[j] = (float)((double)x[j] + wa1[j]); wa3[j] = diag[j] * wa1[j]; }
      const double pnorm = norm6(wa3);
      if (iter == 1) delta = std::min(delta, pnorm);
      acc.a.W[0] = warp_rigid_6d(xt);
      OPE_TRY(acc.run<1>(kAccF, &r));
      ++nfev;
      const double fnorm1 = std::sqrt(r[0]);
      double actred = -1.0;
      if (0.1 * fnorm1 < fnorm) actred = 1.0 - (fnorm1 / fnorm) * (fnorm1 / fnorm);
      for (int i = 0; i < N6; ++i) { double s = 0.0; for (int j = i; j < N6; ++j) s += R[i * N6 + j] * wa1[ipvt[j]]; wa3[i] = s; }
      const double t1 = norm6(wa3) / fnorm, t2 = std::sqrt(par) * pnorm / fnorm;
      const double temp1 = t1 * t1, temp2 = t2 * t2;
      const double prered = temp1 + temp2 / 0.5, dirder = -(temp1 + temp2);
      ratio = prered != 0.0 ? actred / prered : 0.0;
      if (ratio <= 0.25) {
        double temp = 0.5;
        if (actred < 0.0) temp = 0.5 * dirder / (dirder + 0.5 * actred);
        if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
        delta = temp * std::min(delta, pnorm / 0.1);
        par /= temp;
      } else if (!(par != 0.0 && ratio < 0.75)) {
        delta = pnorm / 0.5;
        par = 0.5 * par;
      }
      if (ratio >= 1e-4) {
        for (int j = 0; j < N6; ++j) { x[j] = xt[j]; wa3[j] = diag[j] * (double)x[j]; }
        xnorm = norm6(wa3);
        fnorm = fnorm1;
        ++iter;
      }
      const bool small_red = std::fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0;
      if (small_red && delta <= xtol * xnorm) { status = 3; break; }
      if (small_red) { status = 1; break; }
      if (delta <= xtol * xnorm) { status = 2; break; }
      if (nfev >= maxfev) { status = 5; break; }
      if (std::fabs(actred) <= eps_mach && prered <= eps_mach && 0.5 * ratio <= 1.0) { status = 6; break; }
      if (delta <= eps_mach * xnorm) { status = 7; break; }
      if (gnorm <= eps_mach) { status = 8; break; }
    } while (ratio < 1e-4);
    if (status != -1) break;
  }
  if (info) { info[0] = status; info[1] = nfev; info[2] = iter; }
  return OPE_OK;
}

bool solve6(double A[36], double b[6], double x[6]) {
  for (int c = 0; c < 6; ++c) {
    int best = c;
    double bv = std::fabs(A[c * 6 + c]);
    for (int r = c + 1; r < 6; ++r) if (std::fabs(A[r * 6 + c]) > bv) { bv = std::fabs(A[r * 6 + c]); best = r; }
    if (bv == 0) return false;
    if (best != c) { for (int k = 0; k < 6; ++k) std::swap(A[c * 6 + k], A[best * 6 + k]); std::swap(b[c], b[best]); }
    for (int r = c + 1; r < 6; ++r) {
      const double f = A[r * 6 + c] / A[c * 6 + c];
      for (int k = c; k < 6; ++k) A[r * 6 + k] -= f * A[c * 6 + k];
      b[r] -= f * b[c];
    }
  }
  for (int r = 5; r >= 0; --r) {
    double s = b[r];
    for (int k = r + 1; k < 6; ++k) s -= A[r * 6 + k] * x[k];
    x[r] = s / A[r * 6 + r];
  }
  return true;
}

}  // namespace

// Estimate the rigid transform of the pairs (src[is[i]], tgt[it[i]]), it[i] >= 0, with the point-to-plane metric.
// kind = OPE_TE_POINT_TO_PLANE_LLS | OPE_TE_POINT_TO_PLANE. n_pairs / sum_d2 (sum of d2 over the pairs, for the ICP loop's MSE)
// and lm_info (status, nfev, iterations) may be null.
int point_to_plane_device(ope_ctx* ctx, const float4* src, const float4* tgt, const float4* tgt_n, const int* d_is, const int* d_it,
                          const float* d_d2, size_t n, int kind, Mat4* T, int* n_pairs, double* sum_d2, int32_t* lm_info) {
  *T = mat4_identity();
  if (n_pairs) *n_pairs = 0;
  if (sum_d2) *sum_d2 = 0.0;
  if (lm_info) lm_info[0] = lm_info[1] = lm_info[2] = 0;
  if (kind != OPE_TE_POINT_TO_PLANE_LLS && kind != OPE_TE_POINT_TO_PLANE) return fail(ctx, OPE_ERR_UNSUPPORTED, "unknown point-to-plane estimator %d", kind);
  if (!tgt_n) return fail(ctx, OPE_ERR_INVALID, "point-to-plane estimation needs target normals");
  if (n == 0) return OPE_OK;
  if (n > 0x7fffffffull) return fail(ctx, OPE_ERR_INVALID, "too many correspondences");
  Accum acc(ctx);
  acc.a.src = src; acc.a.tgt = tgt; acc.a.tgt_n = tgt_n; acc.a.is = d_is; acc.a.it = d_it; acc.a.d2 = d_d2; acc.a.n = (int)n;
  for (int j = 0; j < 7; ++j) acc.a.W[j] = mat4_identity();
  for (int j = 0; j < 6; ++j) acc.a.h[j] = 1.0f;
  OPE_TRY(acc.init());
  const double* r;
  if (kind == OPE_TE_POINT_TO_PLANE_LLS) {
    OPE_TRY(acc.run<0>(kAccLls, &r));
    if (n_pairs) *n_pairs = (int)r[27];
    if (sum_d2) *sum_d2 = r[28];
    double A[36], b[6], x[6] = {0, 0, 0, 0, 0, 0};
    int e = 0;
    for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { A[i * 6 + j] = A[j * 6 + i] = r[e++]; }
    for (int i = 0; i < 6; ++i) b[i] = r[21 + i];
    solve6(A, b, x);
    // constructTransformationMatrix(alpha, beta, gamma, tx, ty, tz) [UPSTREAM], double trigonometry rounded once
    const double al = x[0], be = x[1], ga = x[2];
    Mat4 M = mat4_identity();
    M(0, 0) = (float)(cos(ga) * cos(be));
    M(0, 1) = (float)(-sin(ga) * cos(al) + cos(ga) * sin(be) * sin(al));
    M(0, 2) = (float)(sin(ga) * sin(al) + cos(ga) * sin(be) * cos(al));
    M(1, 0) = (float)(sin(ga) * cos(be));
    M(1, 1) = (float)(cos(ga) * cos(al) + sin(ga) * sin(be) * sin(al));
    M(1, 2) = (float)(-cos(ga) * sin(al) + sin(ga) * sin(be) * cos(al));
    M(2, 0) = (float)(-sin(be));
    M(2, 1) = (float)(cos(be) * sin(al));
    M(2, 2) = (float)(cos(be) * cos(al));
    M(0, 3) = (float)x[3]; M(1, 3) = (float)x[4]; M(2, 3) = (float)x[5];
    *T = M;
    return OPE_OK;
  }
  // LM: the pair count first (fewer than 4 pairs: PCL_ERROR + identity, transformation_estimation_lm.hpp)
  OPE_TRY(acc.run<1>(kAccF, &r));
  const int pairs = (int)r[1];
  if (n_pairs) *n_pairs = pairs;
  if (sum_d2) *sum_d2 = r[2];
  if (pairs < 4) return OPE_OK;
  float x[6] = {0, 0, 0, 0, 0, 0};
  OPE_TRY(lm_minimize(acc, x, lm_info));
  *T = warp_rigid_6d(x);
  return OPE_OK;
}

}  // namespace ope
