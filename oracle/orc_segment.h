// orc_segment.h — TEST INFRASTRUCTURE (CPU oracle). ObjectSegmentationPlane::getSegmentedObjectsOnPlane
// (D&L/src/objectsegmentationplane.cpp:122-282) restated: plane by pcl::SACSegmentation (SACMODEL_PLANE, SAC_RANSAC, threshold
// 0.01, setOptimizeCoefficients(true); :36-55 — setAxis / setEpsAngle are ignored by SACMODEL_PLANE), inliers projected onto the
// plane, the bounding rectangle of their convex hull padded by 0.1 m (:174-196), pcl::ExtractPolygonalPrismData over that
// rectangle (:199-203), a second plane fit on the prism's points, the non-plane rest (:223-231) and Euclidean clustering.
// Everything below the call sites is [UPSTREAM] PCL 1.7.2 (sample_consensus/impl/{ransac,sac_model_plane}.hpp, sac_model.h,
// segmentation/impl/{sac_segmentation,extract_polygonal_prism_data}.hpp, filters/impl/project_inliers.hpp), restated from the
// published algorithms: PARITY UNPINNED. Random sampling: SampleConsensusModel seeds boost::mt19937 with 12345 for every
// segment() call (random = false) and draws through boost::uniform_int<>(0, INT_MAX), i.e. mt19937() >> 1.
#pragma once
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <random>
#include <vector>

#include "orc_kdtree.h"
#include "orc_linalg.h"

namespace orc {

struct PlaneSampler {   // SampleConsensusModel::drawIndexSample over shuffled_indices_ (sac_model.h)
  std::mt19937 rng{12345u};
  std::vector<int> shuffled;
  explicit PlaneSampler(size_t n) : shuffled(n) { for (size_t i = 0; i < n; ++i) shuffled[i] = (int)i; }
  unsigned rnd() { return (unsigned)(rng() >> 1); }
  void draw(int out[3]) {
    const size_t index_size = shuffled.size();
    for (unsigned i = 0; i < 3; ++i) std::swap(shuffled[i], shuffled[i + (rnd() % (index_size - i))]);
    out[0] = shuffled[0]; out[1] = shuffled[1]; out[2] = shuffled[2];
  }
};

// SampleConsensusModelPlane::isSampleGood: (p1 - p0) / (p2 - p0) componentwise, not all equal
inline bool planeSampleGood(const float* p0, const float* p1, const float* p2) {
  const float d0 = (p1[0] - p0[0]) / (p2[0] - p0[0]), d1 = (p1[1] - p0[1]) / (p2[1] - p0[1]), d2 = (p1[2] - p0[2]) / (p2[2] - p0[2]);
  return (d0 != d1) || (d2 != d1);
}
// SampleConsensusModelPlane::computeModelCoefficients: normalised cross product, d = -n . p0
inline bool planeFromSample(const float* p0, const float* p1, const float* p2, float c[4]) {
  const float a[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, b[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
  const float d0 = a[0] / b[0], d1 = a[1] / b[1], d2 = a[2] / b[2];
  if ((d0 == d1) && (d2 == d1)) return false;   // collinear
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
  c[3] = 0.0f;
  const float nrm = std::sqrt(redux4(c[0] * c[0], c[1] * c[1], c[2] * c[2], c[3] * c[3]));
  for (int i = 0; i < 4; ++i) c[i] /= nrm;
  c[3] = -1.0f * redux4(c[0] * p0[0], c[1] * p0[1], c[2] * p0[2], c[3] * 1.0f);
  return true;
}
inline float planeDistance(const float c[4], const float* p) { return redux4(c[0] * p[0], c[1] * p[1], c[2] * p[2], c[3] * 1.0f); }

// least-squares refit over the inliers (optimizeModelCoefficients): single-pass float mean + covariance in index order, eigen33
inline void planeRefit(const float* pts, size_t stride, const std::vector<int32_t>& inliers, const float in[4], float out[4]) {
  if (inliers.size() < 4) { for (int i = 0; i < 4; ++i) out[i] = in[i]; return; }
  float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int32_t idx : inliers) {
    const float* p = pts + (size_t)idx * stride;
    accu[0] += p[0] * p[0]; accu[1] += p[0] * p[1]; accu[2] += p[0] * p[2];
    accu[3] += p[1] * p[1]; accu[4] += p[1] * p[2]; accu[5] += p[2] * p[2];
    accu[6] += p[0]; accu[7] += p[1]; accu[8] += p[2];
  }
  const float fc = (float)inliers.size();
  for (int i = 0; i < 9; ++i) accu[i] /= fc;
  float cov[9];
  cov[0] = accu[0] - accu[6] * accu[6]; cov[1] = accu[1] - accu[6] * accu[7]; cov[2] = accu[2] - accu[6] * accu[8];
  cov[4] = accu[3] - accu[7] * accu[7]; cov[5] = accu[4] - accu[7] * accu[8]; cov[8] = accu[5] - accu[8] * accu[8];
  cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
  float ev, n[3];
  eigen33(cov, ev, n);
  out[0] = n[0]; out[1] = n[1]; out[2] = n[2]; out[3] = 0.0f;
  out[3] = -1.0f * redux4(out[0] * accu[6], out[1] * accu[7], out[2] * accu[8], out[3] * 1.0f);   // xyz_centroid[3] = 1
}

// pcl::SACSegmentation::segment for SACMODEL_PLANE / SAC_RANSAC with refined coefficients. Returns false when no model was found.
inline bool planeSegment(const float* pts, size_t n, size_t stride, double threshold, int max_iterations, double probability, float coeff[4],
                         std::vector<int32_t>& inliers, int* iterations_out = nullptr) {
  inliers.clear();
  if (n < 3) return false;
  PlaneSampler sampler(n);
  int iterations = 0, n_best = -INT_MAX;
  double k = 1.0;
  const double log_probability = std::log(1.0 - probability), one_over_indices = 1.0 / (double)n;
  unsigned skipped = 0;
  const unsigned max_skip = (unsigned)max_iterations * 10;
  float best[4] = {0, 0, 0, 0};
  bool have = false;
  while (iterations < k && skipped < max_skip) {
    int s[3];
    bool good = false;
    for (unsigned it = 0; it < 1000 && !good; ++it) {   // getSamples: max_sample_checks_
      sampler.draw(s);
      good = planeSampleGood(pts + (size_t)s[0] * stride, pts + (size_t)s[1] * stride, pts + (size_t)s[2] * stride);
    }
    if (!good) break;
    float c[4];
    if (!planeFromSample(pts + (size_t)s[0] * stride, pts + (size_t)s[1] * stride, pts + (size_t)s[2] * stride, c)) { ++skipped; continue; }
    int cnt = 0;
    for (size_t i = 0; i < n; ++i) if (std::fabs(planeDistance(c, pts + i * stride)) < threshold) ++cnt;
    if (cnt > n_best) {
      n_best = cnt; have = true;
      for (int i = 0; i < 4; ++i) best[i] = c[i];
      const double w = (double)n_best * one_over_indices;
      double p_no_outliers = 1.0 - std::pow(w, 3.0);
      p_no_outliers = std::max(std::numeric_limits<double>::epsilon(), p_no_outliers);
      p_no_outliers = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no_outliers);
      k = log_probability / std::log(p_no_outliers);
    }
    ++iterations;
    if (iterations > max_iterations) break;
  }
  if (iterations_out) *iterations_out = iterations;
  if (!have) return false;
  for (size_t i = 0; i < n; ++i) if (std::fabs(planeDistance(best, pts + i * stride)) < threshold) inliers.push_back((int32_t)i);
  float refined[4];
  planeRefit(pts, stride, inliers, best, refined);
  for (int i = 0; i < 4; ++i) coeff[i] = refined[i];
  inliers.clear();
  for (size_t i = 0; i < n; ++i) if (std::fabs(planeDistance(refined, pts + i * stride)) < threshold) inliers.push_back((int32_t)i);
  return true;
}

// SampleConsensusModelPlane::projectPoints: p - n_unit * (coeff' . p) with coeff' = (n_unit, d)
inline void planeProject(const float c[4], const float* p, float out[3]) {
  float mc[4] = {c[0], c[1], c[2], 0.0f};
  const float nrm = std::sqrt(redux4(mc[0] * mc[0], mc[1] * mc[1], mc[2] * mc[2], mc[3] * mc[3]));
  for (int i = 0; i < 4; ++i) mc[i] /= nrm;
  const float dist = redux4(mc[0] * p[0], mc[1] * p[1], mc[2] * p[2], c[3] * 1.0f);
  for (int i = 0; i < 3; ++i) out[i] = p[i] - mc[i] * dist;
}

// pcl::isXYPointIn2DXYPolygon (crossing number, doubles)
inline bool pointInPolygon(double px, double py, const float poly[][2], int n) {
  bool in = false;
  double xold = poly[n - 1][0], yold = poly[n - 1][1];
  for (int i = 0; i < n; ++i) {
    const double xnew = poly[i][0], ynew = poly[i][1];
    double x1, x2, y1, y2;
    if (xnew > xold) { x1 = xold; x2 = xnew; y1 = yold; y2 = ynew; } else { x1 = xnew; x2 = xold; y1 = ynew; y2 = yold; }
    if ((xnew < px) == (px <= xold) && (py - y1) * (x2 - x1) < (y2 - y1) * (px - x1)) in = !in;
    xold = xnew; yold = ynew;
  }
  return in;
}

// the padded bounding rectangle of the projected plane inliers (their convex hull has the same bounding box), with z from the plane
inline void hullRectangle(const float* pts, size_t stride, const std::vector<int32_t>& inliers, const float c[4], double margin, float rect[4][3]) {
  float mnx = FLT_MAX, mny = FLT_MAX, mxx = -FLT_MAX, mxy = -FLT_MAX;   // getMinMax3D of the hull cloud
  for (int32_t idx : inliers) {
    float q[3];
    planeProject(c, pts + (size_t)idx * stride, q);
    mnx = std::min(mnx, q[0]); mny = std::min(mny, q[1]); mxx = std::max(mxx, q[0]); mxy = std::max(mxy, q[1]);
  }
  // vectorX.push_back(minPt.x - 0.1): float minus double literal, stored as float
  const float vx[4] = {(float)(mnx - margin), (float)(mnx - margin), (float)(mxx + margin), (float)(mxx + margin)};
  const float vy[4] = {(float)(mny - margin), (float)(mxy + margin), (float)(mxy + margin), (float)(mny - margin)};
  for (int i = 0; i < 4; ++i) {
    rect[i][0] = vx[i]; rect[i][1] = vy[i];
    rect[i][2] = -((c[0] * vx[i]) + (c[1] * vy[i]) + c[3]) / c[2];
  }
}

// pcl::ExtractPolygonalPrismData::segment over a 4-point hull, height limits [0, FLT_MAX], viewpoint at the origin
inline void prismSelect(const float* pts, size_t n, size_t stride, const float rect[4][3], std::vector<int32_t>& out) {
  out.clear();
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
  float mc[4] = {nv[0], nv[1], nv[2], 0.0f};
  mc[3] = -1.0f * redux4(mc[0] * accu[6], mc[1] * accu[7], mc[2] * accu[8], mc[3] * 1.0f);
  // flip the plane normal towards the viewpoint (0, 0, 0): vp - hull[0]
  const float vp[3] = {0.0f - rect[0][0], 0.0f - rect[0][1], 0.0f - rect[0][2]};
  const float cos_theta = redux4(vp[0] * mc[0], vp[1] * mc[1], vp[2] * mc[2], 0.0f * mc[3]);
  if (cos_theta < 0) {
    for (int i = 0; i < 4; ++i) mc[i] *= -1.0f;
    mc[3] = 0.0f;
    mc[3] = -1.0f * redux4(mc[0] * rect[0][0], mc[1] * rect[0][1], mc[2] * rect[0][2], mc[3] * 1.0f);
  }
  int k0 = (std::fabs(mc[0]) > std::fabs(mc[1])) ? 0 : 1;
  k0 = (std::fabs(mc[k0]) > std::fabs(mc[2])) ? k0 : 2;
  const int k1 = (k0 + 1) % 3, k2 = (k0 + 2) % 3;
  float poly[4][2];
  for (int i = 0; i < 4; ++i) { poly[i][0] = rect[i][k1]; poly[i][1] = rect[i][k2]; }
  for (size_t i = 0; i < n; ++i) {
    const float* p = pts + i * stride;
    const double distance = (double)(mc[0] * p[0] + mc[1] * p[1] + mc[2] * p[2] + mc[3]);   // pointToPlaneDistanceSigned
    if (distance < 0.0 || distance > (double)FLT_MAX) continue;
    float q[3];
    planeProject(mc, p, q);
    if (!pointInPolygon(q[k1], q[k2], poly, 4)) continue;
    out.push_back((int32_t)i);
  }
}

}  // namespace orc
