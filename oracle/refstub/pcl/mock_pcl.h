// mock_pcl.h — TEST INFRASTRUCTURE (oracle/_ref build only; the product never sees this).
//
// The minimum of the un-vendored Point Cloud Library API that the reference's VENDORED registration sources
// (/root/reference/DetectAndLocalize/include/pcl/registration/*.h, impl/*.hpp, impl/*.cpp — "VP") need in order to compile
// UNMODIFIED in this container, which has no PCL, FLANN, Eigen or Boost. Everything here is our own code and stands in for
// [UPSTREAM] PCL: containers, point types, the blob conversions, a kd-tree adaptor over the oracle's exact search
// (orc_kdtree.h), the transformation estimators (the oracle's Umeyama / point-to-plane solvers), transformPointCloud and
// DefaultConvergenceCriteria::hasConverged (PCL 1.7.2, restated). What the build validates is therefore the VP-held half of
// the path: the ICP loop (impl/icp_mod.hpp:118-272, impl/icp_modCorr.hpp), Registration::align / getFitnessScore
// (impl/registration_mod.hpp:131-219), CorrespondenceEstimation incl. reciprocal and fixed correspondences
// (impl/correspondence_estimation_mod.hpp:127-303), the normal-shooting loop (impl/correspondence_estimation_normal_shooting_
// weighted.hpp:104-145), the rejector scores (correspondence_rejection_mod.h:352-391), the self-occluded-normal rejector
// (impl/correspondence_rejection_self_occluded_normal.cpp:43-64), getAlignStrength (icp_mod.h:249-260) and the wiring of the
// convergence criteria (default_convergence_criteria_mod.h). The [UPSTREAM] pieces stay unpinned.
#ifndef OPE_REFSTUB_MOCK_PCL_H_
#define OPE_REFSTUB_MOCK_PCL_H_

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <type_traits>
#include <vector>

#include <Eigen/Core>
#include <boost/shim.hpp>

#include "../../orc_kdtree.h"
#include "../../orc_linalg.h"
#include "../../../include/ope_types.h"

extern "C" int orc_point_to_plane(const float* src, size_t ns, size_t sstride, const float* tgt, size_t nt, size_t tstride,
                                  const float* tgt_normals4, const int32_t* is, const int32_t* it, size_t n, int kind, float T[16],
                                  int32_t* lm_info);

#define PCL_EXPORTS
#define PCL_DEPRECATED(msg)
#define pcl_isfinite(x) std::isfinite(x)
#define PCL_ERROR(...) std::fprintf(stderr, __VA_ARGS__)
#define PCL_WARN(...) do { } while (0)
#define PCL_DEBUG(...) do { } while (0)
#define PCL_INFO(...) do { } while (0)

namespace pcl {

using std::uint8_t;
using std::uint32_t;

// ---- point types (SURVEY A.9 layouts) ------------------------------------------------------------------------------
struct PCLPointField {
  std::string name;
  uint32_t offset = 0;
  uint8_t datatype = 7;  // FLOAT32
  uint32_t count = 1;
};

struct EIGEN_ALIGN16 PointXYZ {
  union { float data[4]; struct { float x, y, z; }; };
  PointXYZ() : data{0.f, 0.f, 0.f, 1.f} {}
  Eigen::Vector3f getVector3fMap() const { return Eigen::Vector3f(x, y, z); }
  static void fields(std::vector<PCLPointField>& f) { f = {{"x", 0, 7, 1}, {"y", 4, 7, 1}, {"z", 8, 7, 1}}; }
};
struct EIGEN_ALIGN16 Normal {
  union { float data_n[4]; float normal[3]; struct { float normal_x, normal_y, normal_z; }; };
  union { struct { float curvature; }; float data_c[4]; };
  Normal() : data_n{0.f, 0.f, 0.f, 0.f}, data_c{0.f, 0.f, 0.f, 0.f} {}
  static void fields(std::vector<PCLPointField>& f) {
    f = {{"normal_x", 0, 7, 1}, {"normal_y", 4, 7, 1}, {"normal_z", 8, 7, 1}, {"curvature", 16, 7, 1}};
  }
};
struct EIGEN_ALIGN16 PointNormal {
  union { float data[4]; struct { float x, y, z; }; };
  union { float data_n[4]; float normal[3]; struct { float normal_x, normal_y, normal_z; }; };
  union { struct { float curvature; }; float data_c[4]; };
  PointNormal() : data{0.f, 0.f, 0.f, 1.f}, data_n{0.f, 0.f, 0.f, 0.f}, data_c{0.f, 0.f, 0.f, 0.f} {}
  Eigen::Vector3f getVector3fMap() const { return Eigen::Vector3f(x, y, z); }
  static void fields(std::vector<PCLPointField>& f) {
    f = {{"x", 0, 7, 1}, {"y", 4, 7, 1}, {"z", 8, 7, 1}, {"normal_x", 16, 7, 1}, {"normal_y", 20, 7, 1}, {"normal_z", 24, 7, 1},
         {"curvature", 32, 7, 1}};
  }
};
static_assert(sizeof(PointXYZ) == 16 && sizeof(Normal) == 32 && sizeof(PointNormal) == 48, "PCL point layouts");

struct PCLHeader { uint32_t seq = 0; uint64_t stamp = 0; std::string frame_id; };

template <typename PointT>
class PointCloud {
 public:
  typedef boost::shared_ptr<PointCloud<PointT> > Ptr;
  typedef boost::shared_ptr<const PointCloud<PointT> > ConstPtr;
  PCLHeader header;
  std::vector<PointT> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void resize(size_t n) { points.resize(n); width = (uint32_t)n; height = 1; }
  PointT& operator[](size_t i) { return points[i]; }
  const PointT& operator[](size_t i) const { return points[i]; }
  void push_back(const PointT& p) { points.push_back(p); width = (uint32_t)points.size(); height = 1; }
  Ptr makeShared() const { return Ptr(new PointCloud<PointT>(*this)); }
};

template <typename PointT> void getFields(const PointCloud<PointT>&, std::vector<PCLPointField>& f) { PointT::fields(f); }

struct PCLPointCloud2 {
  typedef boost::shared_ptr<PCLPointCloud2> Ptr;
  typedef boost::shared_ptr<const PCLPointCloud2> ConstPtr;
  PCLHeader header;
  uint32_t height = 0, width = 0;
  std::vector<PCLPointField> fields;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  bool is_dense = true;
};
template <typename PointT> void toPCLPointCloud2(const PointCloud<PointT>& c, PCLPointCloud2& b) {
  b.header = c.header; b.width = (uint32_t)c.points.size(); b.height = 1; b.is_dense = c.is_dense;
  PointT::fields(b.fields);
  b.point_step = sizeof(PointT); b.row_step = b.point_step * b.width;
  b.data.resize((size_t)b.row_step);
  if (!c.points.empty()) std::memcpy(b.data.data(), c.points.data(), b.data.size());
}
// field-name mapping, like pcl::fromPCLPointCloud2 + createMapping: fields the blob lacks keep their default value
template <typename PointT> void fromPCLPointCloud2(const PCLPointCloud2& b, PointCloud<PointT>& c) {
  std::vector<PCLPointField> want;
  PointT::fields(want);
  const size_t n = (size_t)b.width * b.height;
  c.header = b.header; c.points.assign(n, PointT()); c.width = b.width; c.height = b.height; c.is_dense = b.is_dense;
  for (const PCLPointField& w : want)
    for (const PCLPointField& h : b.fields)
      if (h.name == w.name)
        for (size_t i = 0; i < n; ++i)
          std::memcpy(reinterpret_cast<uint8_t*>(&c.points[i]) + w.offset, b.data.data() + i * b.point_step + h.offset, sizeof(float));
}

typedef boost::shared_ptr<std::vector<int> > IndicesPtr;
typedef boost::shared_ptr<const std::vector<int> > IndicesConstPtr;
struct PointIndices {
  typedef boost::shared_ptr<PointIndices> Ptr;
  typedef boost::shared_ptr<const PointIndices> ConstPtr;
  PCLHeader header;
  std::vector<int> indices;
};

// ---- PCLBase [UPSTREAM pcl_base.hpp]: input cloud + (fake) index list ------------------------------------------------
template <typename PointT>
class PCLBase {
 public:
  typedef pcl::PointCloud<PointT> PointCloud;
  typedef typename PointCloud::Ptr PointCloudPtr;
  typedef typename PointCloud::ConstPtr PointCloudConstPtr;
  PCLBase() : input_(), indices_(), use_indices_(false), fake_indices_(false) {}
  virtual ~PCLBase() {}
  virtual void setInputCloud(const PointCloudConstPtr& cloud) { input_ = cloud; }
  PointCloudConstPtr const getInputCloud() const { return input_; }
  virtual void setIndices(const IndicesPtr& indices) { indices_ = indices; fake_indices_ = false; use_indices_ = true; }
  IndicesPtr const getIndices() { return indices_; }
 protected:
  PointCloudConstPtr input_;
  IndicesPtr indices_;
  bool use_indices_, fake_indices_;
  bool initCompute() {
    if (!input_) return false;
    if (!indices_) { fake_indices_ = true; indices_.reset(new std::vector<int>); }
    if (fake_indices_ && indices_->size() != input_->points.size()) {
      indices_->resize(input_->points.size());
      for (size_t i = 0; i < indices_->size(); ++i) (*indices_)[i] = (int)i;
    }
    return true;
  }
  bool deinitCompute() { return true; }
};

// ---- search::KdTree adaptor over the oracle's exact kd-tree (KdTreeFLANN semantics) ---------------------------
template <typename PointT> class PointRepresentation { public: virtual ~PointRepresentation() {} };
namespace search {
template <typename PointT>
class KdTree {
 public:
  typedef boost::shared_ptr<KdTree<PointT> > Ptr;
  typedef boost::shared_ptr<const KdTree<PointT> > ConstPtr;
  typedef typename pcl::PointCloud<PointT>::ConstPtr PointCloudConstPtr;
  typedef boost::shared_ptr<const pcl::PointRepresentation<PointT> > PointRepresentationConstPtr;
  KdTree(bool = true) {}
  void setPointRepresentation(const PointRepresentationConstPtr&) {}
  void setInputCloud(const PointCloudConstPtr& cloud, const IndicesConstPtr& = IndicesConstPtr()) {
    input_ = cloud;
    tree_.build(cloud->points.empty() ? nullptr : &cloud->points[0].x, cloud->points.size(), sizeof(PointT) / sizeof(float));
  }
  PointCloudConstPtr getInputCloud() const { return input_; }
  int nearestKSearch(const PointT& p, int k, std::vector<int>& idx, std::vector<float>& d2) const {
    std::vector<orc::Neighbor> nn((size_t)std::max(k, 1));
    const int cnt = tree_.knn(&p.x, k, nn.data());
    idx.resize((size_t)cnt); d2.resize((size_t)cnt);          // FLANN shrinks the result when fewer than k points exist
    for (int i = 0; i < cnt; ++i) { idx[(size_t)i] = nn[(size_t)i].idx; d2[(size_t)i] = nn[(size_t)i].d2; }
    return cnt;
  }
  int nearestKSearch(const pcl::PointCloud<PointT>& cloud, int index, int k, std::vector<int>& idx, std::vector<float>& d2) const {
    return nearestKSearch(cloud.points[(size_t)index], k, idx, d2);
  }
 private:
  PointCloudConstPtr input_;
  orc::KdTree tree_;
};
}  // namespace search

// ---- correspondences -----------------------------------------------------------------------------------------------------
struct Correspondence {
  int index_query;
  int index_match;
  union { float distance; float weight; };
  Correspondence() : index_query(0), index_match(-1), distance(std::numeric_limits<float>::max()) {}
  Correspondence(int q, int m, float d) : index_query(q), index_match(m), distance(d) {}
};
typedef std::vector<Correspondence> Correspondences;
typedef boost::shared_ptr<Correspondences> CorrespondencesPtr;
typedef boost::shared_ptr<const Correspondences> CorrespondencesConstPtr;
inline void getRejectedQueryIndices(const Correspondences& before, const Correspondences& after, std::vector<int>& indices,
                                    bool = true) {
  indices.clear();
  for (const Correspondence& b : before) {
    bool kept = false;
    for (const Correspondence& a : after) if (a.index_query == b.index_query) { kept = true; break; }
    if (!kept) indices.push_back(b.index_query);
  }
}

// ---- point-type metaprogramming used by the (never taken) PointSource != PointTarget branches ----------------------------
namespace traits { template <typename T> struct fieldList { typedef T type; }; }
template <typename A, typename B> struct intersect { typedef A type; };
template <typename S, typename T> struct NdConcatenateFunctor { NdConcatenateFunctor(const S&, T&) {} };
template <typename List, typename F> void for_each_type(F) {}
template <typename S, typename T> bool isSamePointType() { return std::is_same<S, T>::value; }
template <typename S, typename T> void copyPoint(const S&, T&) {}
template <typename T> void copyPoint(const T& a, T& b) { b = a; }

// ---- pcl::transformPointCloud / transformPointCloudWithNormals [UPSTREAM common/impl/transforms.hpp, 1.7.2: explicit
// row-times-point sums, the rotation applied to the normals without Transform::rotation()] ---------------------------------
template <typename PointT, typename Scalar>
void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Matrix<Scalar, 4, 4>& T) {
  orc::Mat4 M;
  for (int i = 0; i < 16; ++i) M.m[i] = (float)T.d[i];
  if (&in != &out) out = in;
  for (size_t i = 0; i < out.points.size(); ++i) {
    PointT& p = out.points[i];
    if (!in.is_dense && !orc::finite3(&p.x)) continue;
    float o[3];
    orc::xformPoint(M, &p.x, o);
    p.x = o[0]; p.y = o[1]; p.z = o[2];
  }
}
template <typename PointT, typename Scalar>
void transformPointCloudWithNormals(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Matrix<Scalar, 4, 4>& T) {
  orc::Mat4 M;
  for (int i = 0; i < 16; ++i) M.m[i] = (float)T.d[i];
  if (&in != &out) out = in;
  for (size_t i = 0; i < out.points.size(); ++i) {
    PointT& p = out.points[i];
    if (!orc::finite3(&p.x)) continue;          // the oracle's (and the device's) convention for non-finite points
    float o[3];
    orc::xformPoint(M, &p.x, o);
    p.x = o[0]; p.y = o[1]; p.z = o[2];
    if (!orc::finite3(p.normal)) continue;
    orc::xformNormal(M, p.normal, o);
    p.normal_x = o[0]; p.normal_y = o[1]; p.normal_z = o[2];
  }
}

namespace registration {

// ---- TransformationEstimation and the three estimators the apps plug in [UPSTREAM] ---------------------------------------
template <typename PointSource, typename PointTarget, typename Scalar = float>
class TransformationEstimation {
 public:
  typedef Eigen::Matrix<Scalar, 4, 4> Matrix4;
  typedef boost::shared_ptr<TransformationEstimation<PointSource, PointTarget, Scalar> > Ptr;
  typedef boost::shared_ptr<const TransformationEstimation<PointSource, PointTarget, Scalar> > ConstPtr;
  virtual ~TransformationEstimation() {}
  virtual void estimateRigidTransformation(const pcl::PointCloud<PointSource>& src, const pcl::PointCloud<PointTarget>& tgt,
                                           const pcl::Correspondences& corr, Matrix4& T) const = 0;
};
// TransformationEstimationSVD -> pcl::umeyama: the oracle's restatement, moments relative to the first target point
template <typename PointSource, typename PointTarget, typename Scalar = float>
class TransformationEstimationSVD : public TransformationEstimation<PointSource, PointTarget, Scalar> {
 public:
  typedef typename TransformationEstimation<PointSource, PointTarget, Scalar>::Matrix4 Matrix4;
  void estimateRigidTransformation(const pcl::PointCloud<PointSource>& src, const pcl::PointCloud<PointTarget>& tgt,
                                   const pcl::Correspondences& corr, Matrix4& T) const {
    float M[16];
    orc::umeyama(corr.size(), [&](size_t i) { return &src.points[(size_t)corr[i].index_query].x; },
                 [&](size_t i) { return &tgt.points[(size_t)corr[i].index_match].x; }, M, tgt.points.empty() ? nullptr : &tgt.points[0].x);
    for (int i = 0; i < 16; ++i) T.d[i] = (Scalar)M[i];
  }
};
template <typename PointSource, typename PointTarget, typename Scalar, int Kind>
class PointToPlaneViaOracle : public TransformationEstimation<PointSource, PointTarget, Scalar> {
 public:
  typedef typename TransformationEstimation<PointSource, PointTarget, Scalar>::Matrix4 Matrix4;
  void estimateRigidTransformation(const pcl::PointCloud<PointSource>& src, const pcl::PointCloud<PointTarget>& tgt,
                                   const pcl::Correspondences& corr, Matrix4& T) const {
    std::vector<int32_t> is(corr.size()), it(corr.size());
    for (size_t i = 0; i < corr.size(); ++i) { is[i] = corr[i].index_query; it[i] = corr[i].index_match; }
    std::vector<float> tn(4 * tgt.points.size());
    for (size_t i = 0; i < tgt.points.size(); ++i) {
      tn[4 * i] = tgt.points[i].normal_x; tn[4 * i + 1] = tgt.points[i].normal_y; tn[4 * i + 2] = tgt.points[i].normal_z;
      tn[4 * i + 3] = tgt.points[i].curvature;
    }
    float M[16];
    orc_point_to_plane(&src.points[0].x, src.points.size(), sizeof(PointSource) / 4, &tgt.points[0].x, tgt.points.size(),
                       sizeof(PointTarget) / 4, tn.data(), is.data(), it.data(), corr.size(), Kind, M, nullptr);
    for (int i = 0; i < 16; ++i) T.d[i] = (Scalar)M[i];
  }
};
template <typename S, typename T, typename Scalar, int Kind, bool HasNormals> struct PointToPlaneSelect {
  typedef PointToPlaneViaOracle<S, T, Scalar, Kind> type;
};
template <typename S, typename T, typename Scalar, int Kind> struct PointToPlaneSelect<S, T, Scalar, Kind, false> {
  typedef TransformationEstimationSVD<S, T, Scalar> type;   // never used: a cloud without normals cannot feed point-to-plane
};
template <typename T> struct has_normals : std::false_type {};
template <> struct has_normals<pcl::PointNormal> : std::true_type {};
template <typename S, typename T, typename Scalar = float>
class TransformationEstimationPointToPlaneLLS : public PointToPlaneSelect<S, T, Scalar, OPE_TE_POINT_TO_PLANE_LLS, has_normals<T>::value>::type {};
template <typename S, typename T, typename Scalar = float>
class TransformationEstimationPointToPlane : public PointToPlaneSelect<S, T, Scalar, OPE_TE_POINT_TO_PLANE, has_normals<T>::value>::type {};

// ---- ConvergenceCriteria [UPSTREAM convergence_criteria.h] ---------------------------------------------------------------
class ConvergenceCriteria {
 public:
  typedef boost::shared_ptr<ConvergenceCriteria> Ptr;
  typedef boost::shared_ptr<const ConvergenceCriteria> ConstPtr;
  ConvergenceCriteria() {}
  virtual ~ConvergenceCriteria() {}
  virtual bool hasConverged() = 0;
  operator bool() { return hasConverged(); }
};

}  // namespace registration
}  // namespace pcl
#endif
