// ope_pcl/filters.h — down-sampling classes and the free functions the reference's hot path calls.
//   UniformSampling  D&L/src/poseestimator.cpp:141-144 (setInputCloud, setRadiusSearch, compute(PointCloud<int>&))
//   VoxelGrid        D&L/src/processingpcd.cpp:51-56  (setInputCloud, setLeafSize, filter)
//   copyPointCloud, removeNaNFromPointCloud, removeNaNNormalsFromPointCloud, transformPointCloud, compute3DCentroid
//                    D&L/src/poseestimator.cpp:26-28,68,145,192-216,358; D&L/src/rosinterface.cpp:269-273
// The classes run on the device through include/ope_cuda.h; the free functions are plain host loops over the
// caller's clouds (a copy, a compaction, one 3x4 multiply per point) with the same float arithmetic as the device.
#pragma once
#include "common.h"

namespace OPE_PCL_NAMESPACE {

// ---- pcl::PCLBase: setInputCloud / setIndices ---------------------------------------------------------------------
template <typename PointT>
class PCLBase {
 public:
  typedef PointCloud<PointT> PointCloudT;
  typedef typename PointCloudT::Ptr PointCloudPtr;
  typedef typename PointCloudT::ConstPtr PointCloudConstPtr;
  virtual ~PCLBase() {}
  virtual void setInputCloud(const PointCloudConstPtr& cloud) { input_ = cloud; }
  PointCloudConstPtr const getInputCloud() const { return input_; }

 protected:
  PointCloudConstPtr input_;
};

// ---- pcl::UniformSampling (PCL 1.7: a Keypoint that outputs indices) ------------------------------------------------
template <typename PointT>
class UniformSampling : public PCLBase<PointT> {
 public:
  using PCLBase<PointT>::input_;
  void setRadiusSearch(double radius) { leaf_ = radius; }
  // indices of the point closest to each voxel's "centre" (SURVEY A.1), ascending voxel key
  void compute(PointCloud<int>& output) {
    output.points.clear(); output.width = 0; output.height = 1; output.is_dense = true;
    if (!input_ || input_->points.empty()) { detail::pcl_error("pcl::UniformSampling::compute", "input cloud is empty"); return; }
    ope_ctx* ctx = detail::context();
    detail::DeviceCloud dc;
    if (!ctx || !detail::upload(*input_, dc, "pcl::UniformSampling::compute")) return;
    std::vector<int32_t> idx(input_->points.size());
    size_t m = 0;
    if (!detail::check(ope_uniform_sample(ctx, dc.get(), (float)leaf_, idx.data(), &m), "pcl::UniformSampling::compute")) return;
    output.points.assign(idx.begin(), idx.begin() + m);
    output.width = (std::uint32_t)m;
  }

 private:
  double leaf_ = 0.0;
};

// ---- pcl::VoxelGrid ---------------------------------------------------------------------------------------------------
template <typename PointT>
class VoxelGrid : public PCLBase<PointT> {
 public:
  using PCLBase<PointT>::input_;
  void setLeafSize(float lx, float ly, float lz) { leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz; }
  void filter(PointCloud<PointT>& output) {
    if (!input_) { detail::pcl_error("pcl::VoxelGrid::filter", "No input dataset given!"); output.clear(); return; }
    ope_ctx* ctx = detail::context();
    detail::DeviceCloud dc;
    if (!ctx || !detail::upload(*input_, dc, "pcl::VoxelGrid::filter")) { output.clear(); return; }
    const size_t n = input_->points.size();
    std::vector<float> rgb, oxyz(3 * n + 3), orgb(n + 1);
    if (detail::has_rgb<PointT>::value) { rgb.resize(n); for (size_t i = 0; i < n; ++i) rgb[i] = get_rgb(input_->points[i], detail::has_rgb<PointT>()); }
    size_t m = 0;
    const int rc = ope_voxel_grid(ctx, dc.get(), rgb.empty() ? nullptr : rgb.data(), leaf_[0], leaf_[1], leaf_[2], oxyz.data(),
                                  orgb.data(), &m);
    if (rc == OPE_ERR_GRID_TOO_LARGE) {  // PCL: warn and return the input unchanged
      detail::pcl_error("pcl::VoxelGrid::applyFilter", "Leaf size is too small for the input dataset. Integer indices would overflow.");
      output = *input_;
      return;
    }
    if (!detail::check(rc, "pcl::VoxelGrid::filter")) { output.clear(); return; }
    output.points.assign(m, PointT());
    for (size_t i = 0; i < m; ++i) {
      output.points[i].x = oxyz[3 * i]; output.points[i].y = oxyz[3 * i + 1]; output.points[i].z = oxyz[3 * i + 2];
      if (!rgb.empty()) set_rgb(output.points[i], orgb[i], detail::has_rgb<PointT>());
    }
    output.width = (std::uint32_t)m; output.height = 1; output.is_dense = true;
  }

 private:
  template <typename P> static float get_rgb(const P& p, std::true_type) { return p.rgb; }
  template <typename P> static float get_rgb(const P&, std::false_type) { return 0.0f; }
  template <typename P> static void set_rgb(P& p, float v, std::true_type) { p.rgb = v; }
  template <typename P> static void set_rgb(P&, float, std::false_type) {}
  float leaf_[3] = {0, 0, 0};
};

// ---- free functions -----------------------------------------------------------------------------------------------------
namespace detail {
template <typename A, typename B> inline void copy_xyz(const A& a, B& b, std::true_type) { b.x = a.x; b.y = a.y; b.z = a.z; }
template <typename A, typename B> inline void copy_xyz(const A&, B&, std::false_type) {}
template <typename A, typename B> inline void copy_nrm(const A& a, B& b, std::true_type) { b.normal_x = a.normal_x; b.normal_y = a.normal_y; b.normal_z = a.normal_z; }
template <typename A, typename B> inline void copy_nrm(const A&, B&, std::false_type) {}
template <typename A, typename B> inline void copy_rgb(const A& a, B& b, std::true_type) { b.rgb = a.rgb; }
template <typename A, typename B> inline void copy_rgb(const A&, B&, std::false_type) {}
template <typename A, typename B> inline void copy_cur(const A& a, B& b, std::true_type) { b.curvature = a.curvature; }
template <typename A, typename B> inline void copy_cur(const A&, B&, std::false_type) {}
// copy the fields both point types have (pcl::copyPointCloud between different point types)
template <typename A, typename B>
inline void copy_point(const A& a, B& b) {
  copy_xyz(a, b, std::integral_constant<bool, has_xyz<A>::value && has_xyz<B>::value>());
  copy_nrm(a, b, std::integral_constant<bool, has_normal<A>::value && has_normal<B>::value>());
  copy_rgb(a, b, std::integral_constant<bool, has_rgb<A>::value && has_rgb<B>::value>());
  copy_cur(a, b, std::integral_constant<bool, has_curvature<A>::value && has_curvature<B>::value>());
}
}  // namespace detail

template <typename PointInT, typename PointOutT>
inline void copyPointCloud(const PointCloud<PointInT>& in, PointCloud<PointOutT>& out) {
  if ((const void*)&in == (const void*)&out) return;  // D&L/src/poseestimator.cpp:220 copies a cloud onto itself
  out.points.resize(in.points.size());
  out.width = in.width; out.height = in.height; out.is_dense = in.is_dense; out.sensor_origin_ = in.sensor_origin_;
  for (size_t i = 0; i < in.points.size(); ++i) detail::copy_point(in.points[i], out.points[i]);
}
template <typename PointT>
inline void copyPointCloud(const PointCloud<PointT>& in, const std::vector<int>& indices, PointCloud<PointT>& out) {
  std::vector<PointT> tmp(indices.size());
  for (size_t i = 0; i < indices.size(); ++i) tmp[i] = in.points[indices[i]];
  out.points.swap(tmp);
  out.width = (std::uint32_t)indices.size(); out.height = 1; out.is_dense = in.is_dense; out.sensor_origin_ = in.sensor_origin_;
}
template <typename PointT>
inline void copyPointCloud(const PointCloud<PointT>& in, const PointCloud<int>& indices, PointCloud<PointT>& out) {
  copyPointCloud(in, indices.points, out);
}

template <typename PointT>
inline void removeNaNFromPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, std::vector<int>& index) {
  std::vector<PointT> tmp;
  tmp.reserve(in.points.size());
  index.clear();
  for (size_t i = 0; i < in.points.size(); ++i) {
    const PointT& p = in.points[i];
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    tmp.push_back(p);
    index.push_back((int)i);
  }
  out.sensor_origin_ = in.sensor_origin_;
  out.points.swap(tmp);
  out.width = (std::uint32_t)out.points.size(); out.height = 1; out.is_dense = true;
}
template <typename PointT>
inline void removeNaNNormalsFromPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, std::vector<int>& index) {
  std::vector<PointT> tmp;
  tmp.reserve(in.points.size());
  index.clear();
  for (size_t i = 0; i < in.points.size(); ++i) {
    const PointT& p = in.points[i];
    if (!std::isfinite(p.normal_x) || !std::isfinite(p.normal_y) || !std::isfinite(p.normal_z)) continue;
    tmp.push_back(p);
    index.push_back((int)i);
  }
  out.sensor_origin_ = in.sensor_origin_;
  out.points.swap(tmp);
  out.width = (std::uint32_t)out.points.size(); out.height = 1;
}

namespace detail {
template <typename P> inline void xform_normal_of(const Eigen::Matrix4f& T, const P& in, P& out, std::true_type) {
  out.normal_x = T(0, 0) * in.normal_x + T(0, 1) * in.normal_y + T(0, 2) * in.normal_z;
  out.normal_y = T(1, 0) * in.normal_x + T(1, 1) * in.normal_y + T(1, 2) * in.normal_z;
  out.normal_z = T(2, 0) * in.normal_x + T(2, 1) * in.normal_y + T(2, 2) * in.normal_z;
}
template <typename P> inline void xform_normal_of(const Eigen::Matrix4f&, const P&, P&, std::false_type) {}
}  // namespace detail

// pcl::transformPointCloud: xyz only (normals are left as copied), [UPSTREAM common/impl/transforms.hpp]
template <typename PointT>
inline void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Matrix4f& T) {
  if (&in != &out) { out.points = in.points; out.width = in.width; out.height = in.height; out.is_dense = in.is_dense; out.sensor_origin_ = in.sensor_origin_; }
  for (size_t i = 0; i < out.points.size(); ++i) {
    PointT& p = out.points[i];
    if (!in.is_dense && (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z))) continue;
    const float x = p.x, y = p.y, z = p.z;
    p.x = T(0, 0) * x + T(0, 1) * y + T(0, 2) * z + T(0, 3);
    p.y = T(1, 0) * x + T(1, 1) * y + T(1, 2) * z + T(1, 3);
    p.z = T(2, 0) * x + T(2, 1) * y + T(2, 2) * z + T(2, 3);
  }
}
template <typename PointT>
inline void transformPointCloudWithNormals(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Matrix4f& T) {
  const PointCloud<PointT> src = in;  // `in` may alias `out`
  transformPointCloud(src, out, T);
  for (size_t i = 0; i < out.points.size(); ++i) detail::xform_normal_of(T, src.points[i], out.points[i], detail::has_normal<PointT>());
}

// pcl::compute3DCentroid: mean of the finite points, w = 1; returns the number of points used
template <typename PointT>
inline unsigned compute3DCentroid(const PointCloud<PointT>& cloud, Eigen::Vector4f& centroid) {
  centroid = Eigen::Vector4f();
  unsigned n = 0;
  for (const PointT& p : cloud.points) {
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    centroid[0] += p.x; centroid[1] += p.y; centroid[2] += p.z;
    ++n;
  }
  if (n) { centroid[0] /= (float)n; centroid[1] /= (float)n; centroid[2] /= (float)n; }
  centroid[3] = 1.0f;
  return n;
}

}  // namespace OPE_PCL_NAMESPACE
