// orc_linalg.h — TEST INFRASTRUCTURE (CPU oracle). Not part of the product path.
//
// Small dense algebra the reference gets from Eigen / pcl::common, restated in float32 with a fixed,
// sequential operation order (PCL/Eigen are absent from this container -> parity unpinned):
//   * eigen33 / computeRoots        [UPSTREAM pcl/common/impl/eigen.hpp]   
//   * 3x3 SVD (one-sided Jacobi) standing in for Eigen::JacobiSVD          
//   * umeyama(src, dst, with_scaling=false) as called by TransformationEstimationSVD
//     (call sites D&L/src/poseestimator.cpp:306,435; SAC-IA default estimator)
//   * 4x4 float helpers with the arithmetic of pcl::transformPointCloud [UPSTREAM common/impl/transforms.hpp]
// All matrices are COLUMN-MAJOR like Eigen: M(r,c) = m[c*rows + r].
#pragma once
#include <algorithm>
#include "orc_kdtree.h"   // orc::finite3
#include <cfloat>
#include <cmath>
#include <cstring>

namespace orc {

struct Mat4 {
  float m[16];
  float& operator()(int r, int c) { return m[c * 4 + r]; }
  float operator()(int r, int c) const { return m[c * 4 + r]; }
  static Mat4 identity() {
    Mat4 I; std::memset(I.m, 0, sizeof(I.m));
    I.m[0] = I.m[5] = I.m[10] = I.m[15] = 1.0f; return I;
  }
  bool isIdentity() const { Mat4 I = identity(); return std::memcmp(m, I.m, sizeof(m)) == 0; }
};

// C = A * B, each entry summed left to right in float.
inline Mat4 mul(const Mat4& A, const Mat4& B) {
  Mat4 C;
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      float s = A(r, 0) * B(0, c);
      s = s + A(r, 1) * B(1, c);
      s = s + A(r, 2) * B(2, c);
      s = s + A(r, 3) * B(3, c);
      C(r, c) = s;
    }
  return C;
}

// pcl::transformPointCloud arithmetic: ((t00*x + t01*y) + t02*z) + t03.
inline void xformPoint(const Mat4& T, const float* p, float* o) {
  float x = p[0], y = p[1], z = p[2];
  o[0] = T(0, 0) * x + T(0, 1) * y + T(0, 2) * z + T(0, 3);
  o[1] = T(1, 0) * x + T(1, 1) * y + T(1, 2) * z + T(1, 3);
  o[2] = T(2, 0) * x + T(2, 1) * y + T(2, 2) * z + T(2, 3);
}
// pcl::transformPointCloudWithNormals: rotation part only.
inline void xformNormal(const Mat4& T, const float* n, float* o) {
  float x = n[0], y = n[1], z = n[2];
  o[0] = T(0, 0) * x + T(0, 1) * y + T(0, 2) * z;
  o[1] = T(1, 0) * x + T(1, 1) * y + T(1, 2) * z;
  o[2] = T(2, 0) * x + T(2, 1) * y + T(2, 2) * z;
}

// ---------------------------------------------------------------------------------------------
// The ONE statement of the Eigen build the restatement assumes (two different orders were assumed before).
// Eigen evaluates a 4-float reduction (Vector4f dot / squaredNorm / norm) as one packet product followed by a horizontal add:
//   level 3  SSE3 and later, `haddps` twice:          (p0 + p1) + (p2 + p3)      <- ASSUMED (default)
//   level 2  plain SSE2, movehl + add + shuffle:      (p0 + p2) + (p1 + p3)
//   level 0  not vectorised (EIGEN_DONT_VECTORIZE):   ((p0 + p1) + p2) + p3
// Assumed platform: PCL 1.7.1 / 1.7.2 (the only releases the vendored headers compile against) with Eigen 3.2.x,
// built with PCL's own SSE flags (-msse4.2 -mfpmath=sse), which reach the reference's translation units through
// PCL_DEFINITIONS (D&L/CMakeLists.txt:60 add_definitions(${PCL_DEFINITIONS})): level 3. For the w = 0 vectors of the pair
// features and the point-to-plane residual, level 3 equals the plain left-to-right 3-term sum. Expressions Eigen 3.2 cannot
// vectorise — UniformSampling's `(pt - ijk.cast<float>()).squaredNorm()`: int->float casts have no packet form before
// Eigen 3.3 — are summed left to right whatever the level (g_cast_vectorized = false); with Eigen >= 3.3 they would follow
// the level too. orc_set_eigen_model() switches both for the sensitivity test (tests/test_oracle.py); the device implements
// the default only. This choice is [UPSTREAM]-unpinned: no Eigen exists in this image to check it.
inline int& eigen_redux_level() { static int level = 3; return level; }
inline bool& eigen_cast_vectorized() { static bool v = false; return v; }
inline float redux4(float p0, float p1, float p2, float p3) {
  const int level = eigen_redux_level();
  if (level == 3) return (p0 + p1) + (p2 + p3);
  if (level == 2) return (p0 + p2) + (p1 + p3);
  return ((p0 + p1) + p2) + p3;
}
inline float dot3w0(const float* a, const float* b) { return redux4(a[0] * b[0], a[1] * b[1], a[2] * b[2], 0.0f * 0.0f); }

// ---------------------------------------------------------------------------------------------
// Float transcendentals as CORRECTLY ROUNDED functions: evaluated in double and rounded once. The reference calls
// libm's float functions (acos/atan2f/cos/sin on float arguments); their last bit depends on the libm build (glibc < 2.41
// does not promise correct rounding), so the restatement pins them to the ideal value — independent of the host's libm,
// and what the device computes as well.
inline float atan2_cr(float y, float x) { return (float)std::atan2((double)y, (double)x); }
inline float acos_cr(float x) { return (float)std::acos((double)x); }
inline float cos_cr(float x) { return (float)std::cos((double)x); }
inline float sin_cr(float x) { return (float)std::sin((double)x); }

// ---------------------------------------------------------------------------------------------
// pcl::computeRoots2 / computeRoots / eigen33 (smallest eigenvalue + eigenvector), float.
// m is symmetric 3x3, column-major (indexing symmetric so order is irrelevant).
inline void computeRoots2(float b, float c, float roots[3]) {
  roots[0] = 0.0f;
  float d = (float)(b * b - 4.0 * c);  // upstream: Scalar(b*b - 4.0*c) — double intermediate
  if (d < 0.0f) d = 0.0f;
  float sd = std::sqrt(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}

inline void computeRoots(const float m[9], float roots[3]) {
  auto M = [&](int r, int c) { return m[c * 3 + r]; };
  float c0 = M(0, 0) * M(1, 1) * M(2, 2) + 2.0f * M(0, 1) * M(0, 2) * M(1, 2) - M(0, 0) * M(1, 2) * M(1, 2) -
             M(1, 1) * M(0, 2) * M(0, 2) - M(2, 2) * M(0, 1) * M(0, 1);
  float c1 = M(0, 0) * M(1, 1) - M(0, 1) * M(0, 1) + M(0, 0) * M(2, 2) - M(0, 2) * M(0, 2) + M(1, 1) * M(2, 2) -
             M(1, 2) * M(1, 2);
  float c2 = M(0, 0) + M(1, 1) + M(2, 2);
  if (std::fabs(c0) < FLT_EPSILON) {
    computeRoots2(c2, c1, roots);
    return;
  }
  const float s_inv3 = 1.0f / 3.0f;
  const float s_sqrt3 = std::sqrt(3.0f);
  float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.0f) a_over_3 = 0.0f;
  float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.0f) q = 0.0f;
  float rho = std::sqrt(-a_over_3);
  float theta = atan2_cr(std::sqrt(-q), half_b) * s_inv3;
  float cos_theta = cos_cr(theta);
  float sin_theta = sin_cr(theta);
  roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  if (roots[1] >= roots[2]) {
    std::swap(roots[1], roots[2]);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  }
  if (roots[0] <= 0.0f) computeRoots2(c2, c1, roots);
}

inline void cross3(const float a[3], const float b[3], float o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

inline void eigen33(const float mat[9], float& eigenvalue, float evec[3]) {
  float scale = 0.0f;
  for (int i = 0; i < 9; ++i) scale = std::max(scale, std::fabs(mat[i]));
  if (scale <= FLT_MIN) scale = 1.0f;
  float s[9];
  for (int i = 0; i < 9; ++i) s[i] = mat[i] / scale;
  float roots[3];
  computeRoots(s, roots);
  eigenvalue = roots[0] * scale;
  s[0] -= roots[0]; s[4] -= roots[0]; s[8] -= roots[0];
  // rows of the (symmetric) matrix
  float r0[3] = {s[0], s[3], s[6]}, r1[3] = {s[1], s[4], s[7]}, r2[3] = {s[2], s[5], s[8]};
  float v1[3], v2[3], v3[3];
  cross3(r0, r1, v1); cross3(r0, r2, v2); cross3(r1, r2, v3);
  float l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
  float l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
  float l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
  const float* v; float l;
  if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; }
  else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; }
  else { v = v3; l = l3; }
  float n = std::sqrt(l);
  evec[0] = v[0] / n; evec[1] = v[1] / n; evec[2] = v[2] / n;
}

// ---------------------------------------------------------------------------------------------
// One-sided Jacobi SVD of a 3x3 matrix A (column-major): A = U diag(s) V^T, s descending,
// U and V orthogonal (columns completed for rank-deficient input).
template <typename T>
inline void svd3(const T Ain[9], T U[9], T S[3], T V[9]) {
  T A[9];
  for (int i = 0; i < 9; ++i) { A[i] = Ain[i]; V[i] = 0; }
  V[0] = V[4] = V[8] = 1;
  const T eps = std::is_same<T, float>::value ? (T)4.76837158e-7 /* 2^-21 */ : (T)1e-15;
  const T eps2 = eps * eps;
  for (int sweep = 0; sweep < 15; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        T* ap = A + 3 * p; T* aq = A + 3 * q;
        T alpha = ap[0] * ap[0] + ap[1] * ap[1] + ap[2] * ap[2];
        T beta = aq[0] * aq[0] + aq[1] * aq[1] + aq[2] * aq[2];
        T gamma = ap[0] * aq[0] + ap[1] * aq[1] + ap[2] * aq[2];
        if (gamma == 0 || gamma * gamma <= eps2 * (alpha * beta)) continue;  // sqrt-free relative test
        rotated = true;
        T zeta = (beta - alpha) / (2 * gamma);
        T t = (zeta >= 0 ? (T)1 : (T)-1) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
        T c = 1 / std::sqrt(1 + t * t);
        T s = c * t;
        for (int i = 0; i < 3; ++i) {
          T x = ap[i], y = aq[i];
          ap[i] = c * x - s * y; aq[i] = s * x + c * y;
          T vx = V[3 * p + i], vy = V[3 * q + i];
          V[3 * p + i] = c * vx - s * vy; V[3 * q + i] = s * vx + c * vy;
        }
      }
    if (!rotated) break;
  }
  T nrm[3];
  for (int j = 0; j < 3; ++j) nrm[j] = std::sqrt(A[3 * j] * A[3 * j] + A[3 * j + 1] * A[3 * j + 1] + A[3 * j + 2] * A[3 * j + 2]);
  int ord[3] = {0, 1, 2};
  for (int i = 0; i < 2; ++i)  // stable selection sort, descending
    for (int j = i + 1; j < 3; ++j)
      if (nrm[ord[j]] > nrm[ord[i]]) std::swap(ord[i], ord[j]);
  T Vs[9];
  for (int j = 0; j < 3; ++j) {
    int o = ord[j];
    S[j] = nrm[o];
    for (int i = 0; i < 3; ++i) { Vs[3 * j + i] = V[3 * o + i]; U[3 * j + i] = A[3 * o + i]; }
  }
  for (int i = 0; i < 9; ++i) V[i] = Vs[i];
  const T tiny = S[0] * (std::is_same<T, float>::value ? (T)1e-6 : (T)1e-13);
  // normalise / complete U
  if (S[0] > 0) { for (int i = 0; i < 3; ++i) U[i] /= S[0]; }
  else { U[0] = 1; U[1] = 0; U[2] = 0; }
  if (S[1] > tiny) { for (int i = 0; i < 3; ++i) U[3 + i] /= S[1]; }
  else {
    // any unit vector orthogonal to u0
    T* u0 = U; int k = 0;
    if (std::fabs(u0[1]) < std::fabs(u0[k])) k = 1;
    if (std::fabs(u0[2]) < std::fabs(u0[k])) k = 2;
    T e[3] = {0, 0, 0}; e[k] = 1;
    T d = u0[k];
    T w[3] = {e[0] - d * u0[0], e[1] - d * u0[1], e[2] - d * u0[2]};
    T n = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    for (int i = 0; i < 3; ++i) U[3 + i] = w[i] / n;
  }
  if (S[2] > tiny) { for (int i = 0; i < 3; ++i) U[6 + i] /= S[2]; }
  else {
    T* a = U; T* b = U + 3;
    U[6] = a[1] * b[2] - a[2] * b[1];
    U[7] = a[2] * b[0] - a[0] * b[2];
    U[8] = a[0] * b[1] - a[1] * b[0];
  }
}

template <typename T>
inline T det3(const T m[9]) {
  return m[0] * (m[4] * m[8] - m[7] * m[5]) - m[3] * (m[1] * m[8] - m[7] * m[2]) + m[6] * (m[1] * m[5] - m[4] * m[2]);
}

// Rotation/translation from the Umeyama quantities: sigma (3x3, col-major, = 1/n * dst_demean * src_demean^T),
// the two means. Follows Eigen::umeyama with_scaling=false including the rank-2 branch.
template <typename T>
inline void umeyamaFromSigma(const T sigma[9], const T src_mean[3], const T dst_mean[3], float Tout[16]) {
  T U[9], S[3], V[9];
  svd3<T>(sigma, U, S, V);
  T Sd[3] = {1, 1, 1};
  if (det3<T>(sigma) < 0) Sd[2] = -1;
  int rank = 0;
  const T prec = std::is_same<T, float>::value ? (T)1e-5 : (T)1e-12;  // NumTraits::dummy_precision
  for (int i = 0; i < 3; ++i)
    if (!(std::fabs(S[i]) <= std::fabs(S[0]) * prec)) ++rank;
  if (rank == 2) {
    if (det3<T>(U) * det3<T>(V) > 0) { Sd[2] = 1; }
    else { Sd[2] = -1; }
  }
  T R[9];  // R = U * diag(Sd) * V^T
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) {
      T s = U[0 * 3 + r] * Sd[0] * V[0 * 3 + c];
      s = s + U[1 * 3 + r] * Sd[1] * V[1 * 3 + c];
      s = s + U[2 * 3 + r] * Sd[2] * V[2 * 3 + c];
      R[c * 3 + r] = s;
    }
  for (int i = 0; i < 16; ++i) Tout[i] = 0.0f;
  Tout[15] = 1.0f;
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) Tout[c * 4 + r] = (float)R[c * 3 + r];
  for (int r = 0; r < 3; ++r) {
    T rs = R[0 * 3 + r] * src_mean[0];
    rs = rs + R[1 * 3 + r] * src_mean[1];
    rs = rs + R[2 * 3 + r] * src_mean[2];
    Tout[12 + r] = (float)(dst_mean[r] - rs);
  }
}

// Eigen::umeyama(src, dst, false) over n pairs given by callbacks returning pointers to xyz.
//
// Eigen evaluates the means and the 3x3 cross-covariance with vectorised float reductions whose summation order
// cannot be reproduced without Eigen, and ICP's stop test (cos >= 1 - 1e-8 on a float trace) turns 1e-7 differences
// into a different stopping iteration. The oracle therefore accumulates the raw moments in DOUBLE (exact products of
// float inputs, summation error ~1e-16 relative, i.e. independent of the order to far below float resolution) and
// rounds the 3x3 cross-covariance to float; any float summation order Eigen may use lies within the float rounding
// bound of that matrix. From there on it is float like the reference: Jacobi SVD of the float matrix, R = U S V^T in
// float; the translation mu_dst - R mu_src is closed in double and rounded once. A parallel reduction on the device
// reproduces this bit for bit (barring ~1e-9-probability double-rounding ties).
//
// `origin` (3 floats or null = 0): the moments are accumulated relative to it. The cross-covariance does not depend on the origin,
// but sum d s^T / n - mu_d mu_s^T cancels ~3 digits for a 10 cm object at 1 m, which multiplies the summation-order noise of the
// double sums by 10^3 before the rounding to float (observed: one flipped rounding per ~10^5 alignment iterations between two
// summation orders); about a point of the cloud itself it cancels ~1 digit. Callers pass the first point of the TARGET cloud.
template <typename SrcAt, typename DstAt>
inline void umeyama(size_t n, SrcAt srcAt, DstAt dstAt, float Tout[16], const float* origin = nullptr) {
  double o[3] = {0.0, 0.0, 0.0};
  if (origin && finite3(origin)) { o[0] = origin[0]; o[1] = origin[1]; o[2] = origin[2]; }
  double acc[16];
  for (int i = 0; i < 16; ++i) acc[i] = 0.0;
  for (size_t i = 0; i < n; ++i) {
    const float* sp = srcAt(i); const float* dp = dstAt(i);
    const double s[3] = {(double)sp[0] - o[0], (double)sp[1] - o[1], (double)sp[2] - o[2]};
    const double d[3] = {(double)dp[0] - o[0], (double)dp[1] - o[1], (double)dp[2] - o[2]};
    acc[0] += 1.0;
    for (int k = 0; k < 3; ++k) { acc[1 + k] += s[k]; acc[4 + k] += d[k]; }
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r) acc[7 + c * 3 + r] += d[r] * s[c];
  }
  const double nn = acc[0];
  double ms[3], mt[3];
  float sigma[9];
  for (int k = 0; k < 3; ++k) { ms[k] = acc[1 + k] / nn; mt[k] = acc[4 + k] / nn; }
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) sigma[c * 3 + r] = (float)(acc[7 + c * 3 + r] / nn - mt[r] * ms[c]);
  for (int k = 0; k < 3; ++k) { ms[k] += o[k]; mt[k] += o[k]; }
  float U[9], S[3], V[9];
  svd3<float>(sigma, U, S, V);
  float Sd[3] = {1, 1, 1};
  if (det3<float>(sigma) < 0) Sd[2] = -1;
  int rank = 0;
  for (int i = 0; i < 3; ++i)
    if (!(std::fabs(S[i]) <= std::fabs(S[0]) * 1e-5f)) ++rank;  // NumTraits<float>::dummy_precision
  if (rank == 2) Sd[2] = (det3<float>(U) * det3<float>(V) > 0) ? 1.0f : -1.0f;
  float R[9];
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) {
      float s = U[0 * 3 + r] * Sd[0] * V[0 * 3 + c];
      s = s + U[1 * 3 + r] * Sd[1] * V[1 * 3 + c];
      s = s + U[2 * 3 + r] * Sd[2] * V[2 * 3 + c];
      R[c * 3 + r] = s;
    }
  for (int i = 0; i < 16; ++i) Tout[i] = 0.0f;
  Tout[15] = 1.0f;
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) Tout[c * 4 + r] = R[c * 3 + r];
  for (int r = 0; r < 3; ++r) {
    double rs = (double)R[0 * 3 + r] * ms[0];
    rs = rs + (double)R[1 * 3 + r] * ms[1];
    rs = rs + (double)R[2 * 3 + r] * ms[2];
    Tout[12 + r] = (float)(mt[r] - rs);
  }
}

}  // namespace orc
