// ope_pcl/common.h — the types that cross the PCL-style boundary, with PCL's names, layouts and conventions, and the
// plumbing every shim class shares (one CUDA context per host thread, device-cloud handles, error reporting).
//
// The reference instantiates header-only PCL templates in its own translation units; these headers are
// the same kind of thing — header-only classes with PCL's public signatures — whose bodies marshal into the C ABI of
// libope_cuda.so (include/ope_cuda.h). They need neither PCL nor Eigen nor Boost. The namespace is OPE_PCL_NAMESPACE
// (default `ope_pcl`); the forwarding headers under include/ope_pcl_compat/pcl/ set it to `pcl`, so a translation unit
// written against <pcl/...> compiles unchanged with -Iinclude/ope_pcl_compat.
//
// Layouts ([UPSTREAM] pcl/impl/point_types.hpp): PointXYZ 16 B, PointXYZRGB 32 B, Normal 32 B,
// PointNormal / PointXYZRGBNormal 48 B, FPFHSignature33 132 B, Correspondence 12 B {index_query, index_match, distance};
// Eigen::Matrix4f is 16 floats column-major.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <ostream>
#include <type_traits>
#include <vector>

#include "../ope_cuda.h"

#ifndef OPE_PCL_NAMESPACE
#define OPE_PCL_NAMESPACE ope_pcl
#endif

// ---- Eigen stand-ins (only what the reference's call sites use: D&L/src/poseestimator.cpp:64,357,421,439) ----------
#ifdef OPE_PCL_USE_EIGEN
#include <Eigen/Core>
#else
#ifndef OPE_PCL_MINI_EIGEN
#define OPE_PCL_MINI_EIGEN
namespace Eigen {
class Matrix4f {
 public:
  float m[16];  // column-major
  Matrix4f() { std::memset(m, 0, sizeof(m)); }
  static Matrix4f Identity() { Matrix4f I; I.m[0] = I.m[5] = I.m[10] = I.m[15] = 1.0f; return I; }
  static Matrix4f Zero() { return Matrix4f(); }
  void setIdentity() { *this = Identity(); }
  float& operator()(int r, int c) { return m[c * 4 + r]; }
  float operator()(int r, int c) const { return m[c * 4 + r]; }
  float* data() { return m; }
  const float* data() const { return m; }
  // each entry summed left to right in float, like the device's mat4_mul
  Matrix4f operator*(const Matrix4f& B) const {
    Matrix4f C;
    for (int c = 0; c < 4; ++c)
      for (int r = 0; r < 4; ++r) {
        float s = (*this)(r, 0) * B(0, c);
        s = s + (*this)(r, 1) * B(1, c);
        s = s + (*this)(r, 2) * B(2, c);
        s = s + (*this)(r, 3) * B(3, c);
        C(r, c) = s;
      }
    return C;
  }
  bool operator==(const Matrix4f& o) const { return std::memcmp(m, o.m, sizeof(m)) == 0; }
  bool operator!=(const Matrix4f& o) const { return !(*this == o); }
  bool isIdentity() const { return *this == Identity(); }
};
inline std::ostream& operator<<(std::ostream& os, const Matrix4f& M) {
  for (int r = 0; r < 4; ++r) { for (int c = 0; c < 4; ++c) os << M(r, c) << (c < 3 ? " " : ""); if (r < 3) os << "\n"; }
  return os;
}
class Vector4f {
 public:
  float v[4];
  Vector4f() { v[0] = v[1] = v[2] = v[3] = 0.0f; }
  float& operator[](int i) { return v[i]; }
  float operator[](int i) const { return v[i]; }
  float& operator()(int i) { return v[i]; }
  float operator()(int i) const { return v[i]; }
};
}  // namespace Eigen
#endif
#endif

namespace OPE_PCL_NAMESPACE {

// ---- point types ---------------------------------------------------------------------------------------------------
struct alignas(16) PointXYZ { float x, y, z, data_w = 1.0f; };
struct alignas(16) PointXYZRGB {
  float x = 0, y = 0, z = 0, data_w = 1.0f;
  union { float rgb; std::uint32_t rgba; struct { std::uint8_t b, g, r, a; }; };
  float pad_[3];
  PointXYZRGB() : rgba(0) { pad_[0] = pad_[1] = pad_[2] = 0; }
};
struct alignas(16) Normal {
  float normal_x = 0, normal_y = 0, normal_z = 0, data_n_w = 0;
  float curvature = 0;
  float pad_[3] = {0, 0, 0};
};
struct alignas(16) PointNormal {
  float x = 0, y = 0, z = 0, data_w = 1.0f;
  float normal_x = 0, normal_y = 0, normal_z = 0, data_n_w = 0;
  float curvature = 0;
  float pad_[3] = {0, 0, 0};
};
struct alignas(16) PointXYZRGBNormal {
  float x = 0, y = 0, z = 0, data_w = 1.0f;
  float normal_x = 0, normal_y = 0, normal_z = 0, data_n_w = 0;
  union { float rgb; std::uint32_t rgba; };
  float curvature = 0;
  float pad_[2] = {0, 0};
  PointXYZRGBNormal() : rgba(0) {}
};
struct FPFHSignature33 { float histogram[33]; };
static_assert(sizeof(PointXYZ) == 16 && sizeof(PointXYZRGB) == 32 && sizeof(Normal) == 32 && sizeof(PointNormal) == 48 &&
              sizeof(PointXYZRGBNormal) == 48 && sizeof(FPFHSignature33) == 132, "PCL point layouts (SURVEY A.9)");

struct Correspondence {
  int index_query = 0;
  int index_match = -1;
  float distance = std::numeric_limits<float>::max();
  Correspondence() {}
  Correspondence(int q, int m, float d) : index_query(q), index_match(m), distance(d) {}
};
static_assert(sizeof(Correspondence) == sizeof(ope_correspondence), "pcl::Correspondence is 12 bytes");
typedef std::vector<Correspondence> Correspondences;
typedef std::shared_ptr<Correspondences> CorrespondencesPtr;

// ---- pcl::PointCloud --------------------------------------------------------------------------------------------------
template <typename PointT>
class PointCloud {
 public:
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  typedef PointT PointType;
  std::vector<PointT> points;
  std::uint32_t width = 0, height = 0;
  bool is_dense = true;
  Eigen::Vector4f sensor_origin_;

  PointCloud() {}
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); width = height = 0; }
  void resize(std::size_t n) { points.resize(n); width = (std::uint32_t)n; height = 1; }
  void push_back(const PointT& p) { points.push_back(p); width = (std::uint32_t)points.size(); height = 1; }
  PointT& operator[](std::size_t i) { return points[i]; }
  const PointT& operator[](std::size_t i) const { return points[i]; }
  typename std::vector<PointT>::iterator begin() { return points.begin(); }
  typename std::vector<PointT>::iterator end() { return points.end(); }
  typename std::vector<PointT>::const_iterator begin() const { return points.begin(); }
  typename std::vector<PointT>::const_iterator end() const { return points.end(); }
  Ptr makeShared() const { return Ptr(new PointCloud<PointT>(*this)); }
  // *aligned += *target (BM/src/regmeshpcd.cpp:254)
  PointCloud& operator+=(const PointCloud& rhs) {
    points.insert(points.end(), rhs.points.begin(), rhs.points.end());
    width = (std::uint32_t)points.size(); height = 1;
    is_dense = is_dense && rhs.is_dense;
    return *this;
  }
};

namespace detail {

// field detection (PCL uses its field-list traits; the shim only needs "has normals", "has rgb", "has xyz")
template <typename T, typename = void> struct has_xyz : std::false_type {};
template <typename T> struct has_xyz<T, decltype((void)std::declval<T&>().x, void())> : std::true_type {};
template <typename T, typename = void> struct has_normal : std::false_type {};
template <typename T> struct has_normal<T, decltype((void)std::declval<T&>().normal_x, void())> : std::true_type {};
template <typename T, typename = void> struct has_rgb : std::false_type {};
template <typename T> struct has_rgb<T, decltype((void)std::declval<T&>().rgb, void())> : std::true_type {};
template <typename T, typename = void> struct has_curvature : std::false_type {};
template <typename T> struct has_curvature<T, decltype((void)std::declval<T&>().curvature, void())> : std::true_type {};

// PCL_ERROR: message to stderr, no exception, the caller returns with its outputs untouched
inline void pcl_error(const char* where, const char* what) { std::fprintf(stderr, "[%s] %s\n", where, what); }

// one CUDA context per host thread, created on first use; nullptr (after one PCL_ERROR-style line) when there is no
// usable device — there is no CPU fallback behind these classes.
inline ope_ctx* context() {
  struct Holder {
    ope_ctx* ctx = nullptr;
    bool tried = false;
    ~Holder() { if (ctx) ope_ctx_destroy(ctx); }
  };
  static thread_local Holder h;
  if (!h.tried) {
    h.tried = true;
    int dev = 0;
    if (const char* e = std::getenv("OPE_DEVICE")) dev = std::atoi(e);
    const int rc = ope_ctx_create(dev, nullptr, &h.ctx);
    if (rc != OPE_OK) {
      h.ctx = nullptr;
      pcl_error("ope_pcl", rc == OPE_ERR_NO_DEVICE ? "no usable CUDA device: libope_cuda has no CPU fallback, calls will do nothing"
                                                    : "ope_ctx_create failed");
    }
  }
  return h.ctx;
}
inline bool check(int rc, const char* where) {
  if (rc == OPE_OK) return true;
  ope_ctx* c = context();
  pcl_error(where, c ? ope_last_error(c) : "no CUDA context");
  return false;
}

// owning handle of a device cloud
class DeviceCloud {
 public:
  DeviceCloud() {}
  ~DeviceCloud() { reset(); }
  DeviceCloud(const DeviceCloud&) = delete;
  DeviceCloud& operator=(const DeviceCloud&) = delete;
  void reset(ope_cloud* c = nullptr) {
    if (h_ && context()) ope_cloud_free(context(), h_);
    h_ = c;
  }
  ope_cloud* get() const { return h_; }
  ope_cloud** out() { reset(); return &h_; }
  explicit operator bool() const { return h_ != nullptr; }

 private:
  ope_cloud* h_ = nullptr;
};

template <typename PointT>
constexpr std::size_t normal_offset_impl(std::true_type) { return offsetof(PointT, normal_x); }
template <typename PointT>
constexpr std::size_t normal_offset_impl(std::false_type) { return 0; }
template <typename PointT>
constexpr std::size_t normal_offset() { return normal_offset_impl<PointT>(has_normal<PointT>()); }

// upload the points (and, when the point type carries them, the normals) of the whole cloud
template <typename PointT>
inline bool upload(const PointCloud<PointT>& cloud, DeviceCloud& out, const char* where) {
  ope_ctx* ctx = context();
  if (!ctx) return false;
  const void* base = cloud.points.empty() ? nullptr : (const void*)cloud.points.data();
  const void* nbase = nullptr;
  std::size_t noff = 0;
  if (has_normal<PointT>::value && base) { nbase = base; noff = normal_offset<PointT>(); }
  return check(ope_cloud_upload(ctx, base, cloud.points.size(), sizeof(PointT), 0, nbase, sizeof(PointT), noff, out.out()), where);
}

// PCL keeps the caller's shared_ptr and re-reads *input_ / *target_ on every align(); the shim keeps a device copy instead, so it
// must notice a cloud that was modified IN PLACE between two calls (the reference does `*p_sourceCloud = *alignedCloud` and
// re-aligns on the same object): storage address, size and a 64-bit hash of the bytes.
struct CloudFingerprint {
  const void* data = nullptr;
  std::size_t size = 0;
  std::uint64_t hash = 0;
  bool operator==(const CloudFingerprint& o) const { return data == o.data && size == o.size && hash == o.hash; }
  bool operator!=(const CloudFingerprint& o) const { return !(*this == o); }
};
template <typename PointT>
inline CloudFingerprint fingerprint(const PointCloud<PointT>& cloud) {
  CloudFingerprint f;
  f.data = cloud.points.empty() ? nullptr : (const void*)cloud.points.data();
  f.size = cloud.points.size();
  const std::size_t words = f.size * sizeof(PointT) / 8;
  const unsigned char* p = (const unsigned char*)f.data;
  std::uint64_t h0 = 0x9e3779b97f4a7c15ull, h1 = 0xc2b2ae3d27d4eb4full, h2 = 0x165667b19e3779f9ull, h3 = 0x27d4eb2f165667c5ull;
  std::size_t i = 0;
  for (; i + 4 <= words; i += 4) {   // four independent multiply-rotate lanes: memory-bound on any host
    std::uint64_t w[4];
    std::memcpy(w, p + 8 * i, 32);
    h0 = (h0 ^ w[0]) * 0x9fb21c651e98df25ull; h0 = (h0 << 29) | (h0 >> 35);
    h1 = (h1 ^ w[1]) * 0x9fb21c651e98df25ull; h1 = (h1 << 29) | (h1 >> 35);
    h2 = (h2 ^ w[2]) * 0x9fb21c651e98df25ull; h2 = (h2 << 29) | (h2 >> 35);
    h3 = (h3 ^ w[3]) * 0x9fb21c651e98df25ull; h3 = (h3 << 29) | (h3 >> 35);
  }
  for (; i < words; ++i) { std::uint64_t w; std::memcpy(&w, p + 8 * i, 8); h0 = (h0 ^ w) * 0x9fb21c651e98df25ull; h0 = (h0 << 29) | (h0 >> 35); }
  f.hash = h0 ^ (h1 * 3) ^ (h2 * 5) ^ (h3 * 7) ^ (std::uint64_t)f.size;
  return f;
}

// points from one cloud, normals (normal_x.. + curvature, 16 bytes) from another
template <typename PointT, typename NormalT>
inline bool upload_with_normals(const PointCloud<PointT>& cloud, const PointCloud<NormalT>& normals, DeviceCloud& out,
                                const char* where) {
  ope_ctx* ctx = context();
  if (!ctx) return false;
  if (normals.points.size() != cloud.points.size()) { pcl_error(where, "The number of points in the input dataset differs from the number of points in the dataset containing the normals!"); return false; }
  const void* base = cloud.points.empty() ? nullptr : (const void*)cloud.points.data();
  const void* nbase = normals.points.empty() ? nullptr : (const void*)normals.points.data();
  return check(ope_cloud_upload(ctx, base, cloud.points.size(), sizeof(PointT), 0, nbase, sizeof(NormalT),
                                normal_offset<NormalT>(), out.out()), where);
}

inline void to_c(const Eigen::Matrix4f& M, float T[16]) { std::memcpy(T, M.data(), 16 * sizeof(float)); }
inline Eigen::Matrix4f from_c(const float T[16]) { Eigen::Matrix4f M; std::memcpy(M.data(), T, 16 * sizeof(float)); return M; }

}  // namespace detail

namespace search {
// pcl::search::KdTree / pcl::KdTreeFLANN: the application creates one and hands it to setSearchMethod
// (D&L/src/poseestimator.cpp:151-152). The spatial index lives on the device next to the cloud it indexes (a
// Morton-ordered grid), so this is a tag object that only keeps the API shape.
template <typename PointT>
class KdTree {
 public:
  typedef std::shared_ptr<KdTree<PointT>> Ptr;
  typedef std::shared_ptr<const KdTree<PointT>> ConstPtr;
  explicit KdTree(bool sorted = true) : sorted_(sorted) {}
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr& cloud) { input_ = cloud; }
  typename PointCloud<PointT>::ConstPtr getInputCloud() const { return input_; }

 private:
  bool sorted_;
  typename PointCloud<PointT>::ConstPtr input_;
};
}  // namespace search
template <typename PointT>
using KdTreeFLANN = search::KdTree<PointT>;

}  // namespace OPE_PCL_NAMESPACE
