// orc_kdtree.h — TEST INFRASTRUCTURE (CPU oracle). Not part of the product path.
//
// Exact k-NN / radius search with the semantics of pcl::KdTreeFLANN as the reference uses it
// (call sites VP/impl/registration_mod.hpp:149, VP/impl/correspondence_estimation_mod.hpp:170,
// VP/impl/correspondence_estimation_normal_shooting_weighted.hpp:109):
//   * metric flann::L2_Simple: ((dx*dx) + dy*dy) + dz*dz accumulated left to right in float32, no FMA;
//     returned distances are SQUARED;
//   * exact search (unlimited checks, eps = 0), leaf size 15 (KDTreeSingleIndexParams(15));
//   * k-NN results ascending by distance; non-finite points are not indexed;
//   * radius search keeps d2 < r*r (strict).
// Ties are traversal-order dependent in FLANN; the oracle's canonical order is ascending (d2, index)
// for k-NN and ascending index for radius results (SURVEY hard part 2).
// PCL/FLANN are not available in this container -> parity unpinned at this boundary.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace orc {

inline float dist2(const float* a, const float* b) {
  float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  float r = dx * dx;
  r = r + dy * dy;
  r = r + dz * dz;
  return r;
}

inline bool finite3(const float* p) { return std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]); }

struct Neighbor {
  float d2;
  int32_t idx;
  bool operator<(const Neighbor& o) const { return d2 < o.d2 || (d2 == o.d2 && idx < o.idx); }
};

class KdTree {
 public:
  // xyz: first point; stride in floats between consecutive points.
  void build(const float* xyz, size_t n, size_t stride) {
    base_ = xyz; stride_ = stride; n_ = n;
    perm_.clear(); nodes_.clear();
    perm_.reserve(n);
    for (size_t i = 0; i < n; ++i)
      if (finite3(p(i))) perm_.push_back((int32_t)i);
    if (perm_.empty()) return;
    nodes_.reserve(2 * perm_.size() / kLeaf + 8);
    buildRec(0, (int)perm_.size());
  }
  size_t size() const { return perm_.size(); }

  // k nearest of q, ascending (d2, idx). Returns number found (min(k, size)).
  int knn(const float* q, int k, Neighbor* out) const {
    if (perm_.empty() || k <= 0) return 0;
    Heap h{out, 0, k};
    float off[3] = {0, 0, 0};
    searchRec(0, q, h, 0.0f, off);
    std::sort(out, out + h.n);
    return h.n;
  }

  // all neighbours with d2 < r2 (strict), ascending index.
  void radius(const float* q, float r2, std::vector<Neighbor>& out) const {
    out.clear();
    if (perm_.empty()) return;
    float off[3] = {0, 0, 0};
    radiusRec(0, q, r2, out, 0.0f, off);
    std::sort(out.begin(), out.end(), [](const Neighbor& a, const Neighbor& b) { return a.idx < b.idx; });
  }

 private:
  static constexpr int kLeaf = 15;
  struct Node {
    int32_t left, right;  // children (internal) ; leaf: left = begin, right = end, dim = -1
    int32_t dim;
    float lo, hi;         // split: left subtree coordinate <= lo ... right subtree >= hi
  };
  struct Heap {  // bounded max-heap on (d2, idx)
    Neighbor* a; int n; int k;
    float worst() const { return n < k ? INFINITY : a[0].d2; }
    void push(Neighbor v) {
      if (n < k) { a[n++] = v; std::push_heap(a, a + n); }
      else if (v < a[0]) { std::pop_heap(a, a + n); a[n - 1] = v; std::push_heap(a, a + n); }
    }
  };
  const float* p(size_t i) const { return base_ + i * stride_; }

  int buildRec(int b, int e) {
    int id = (int)nodes_.size();
    nodes_.push_back(Node{});
    if (e - b <= kLeaf) { nodes_[id] = Node{b, e, -1, 0, 0}; return id; }
    float mn[3], mx[3];
    for (int d = 0; d < 3; ++d) { mn[d] = INFINITY; mx[d] = -INFINITY; }
    for (int i = b; i < e; ++i) {
      const float* q = p(perm_[i]);
      for (int d = 0; d < 3; ++d) { mn[d] = std::min(mn[d], q[d]); mx[d] = std::max(mx[d], q[d]); }
    }
    int dim = 0;
    for (int d = 1; d < 3; ++d) if (mx[d] - mn[d] > mx[dim] - mn[dim]) dim = d;
    int mid = (b + e) / 2;
    std::nth_element(perm_.begin() + b, perm_.begin() + mid, perm_.begin() + e,
                     [&](int32_t x, int32_t y) { return p(x)[dim] < p(y)[dim]; });
    float lo = -INFINITY, hi = INFINITY;
    for (int i = b; i < mid; ++i) lo = std::max(lo, p(perm_[i])[dim]);
    for (int i = mid; i < e; ++i) hi = std::min(hi, p(perm_[i])[dim]);
    int l = buildRec(b, mid);
    int r = buildRec(mid, e);
    nodes_[id] = Node{l, r, dim, lo, hi};
    return id;
  }

  // mind2: lower bound (in exact arithmetic) of the squared distance from q to this subtree's box;
  // pruning applies a 1e-6 relative safety factor because the float L2_Simple sum rounds differently.
  void searchRec(int id, const float* q, Heap& h, float mind2, float* off) const {
    const Node& nd = nodes_[id];
    if (nd.dim < 0) {
      for (int i = nd.left; i < nd.right; ++i) {
        int32_t j = perm_[i];
        h.push(Neighbor{dist2(q, p(j)), j});
      }
      return;
    }
    int d = nd.dim;
    float v = q[d];
    float dl = v - nd.lo, dh = nd.hi - v;  // >0 => q is beyond that side's extreme
    int nearc, farc; float cut;
    if (dl + (v - nd.hi) < 0) { nearc = nd.left; farc = nd.right; cut = dh > 0 ? dh : 0; }
    else { nearc = nd.right; farc = nd.left; cut = dl > 0 ? dl : 0; }
    searchRec(nearc, q, h, mind2, off);
    float old = off[d];
    float far2 = mind2 - old * old + cut * cut;
    if (far2 * 0.999999f <= h.worst()) {
      off[d] = cut;
      searchRec(farc, q, h, far2 > 0 ? far2 : 0, off);
      off[d] = old;
    }
  }

  void radiusRec(int id, const float* q, float r2, std::vector<Neighbor>& out, float mind2, float* off) const {
    const Node& nd = nodes_[id];
    if (nd.dim < 0) {
      for (int i = nd.left; i < nd.right; ++i) {
        int32_t j = perm_[i];
        float d2 = dist2(q, p(j));
        if (d2 < r2) out.push_back(Neighbor{d2, j});
      }
      return;
    }
    int d = nd.dim;
    float v = q[d];
    float dl = v - nd.lo, dh = nd.hi - v;
    int nearc, farc; float cut;
    if (dl + (v - nd.hi) < 0) { nearc = nd.left; farc = nd.right; cut = dh > 0 ? dh : 0; }
    else { nearc = nd.right; farc = nd.left; cut = dl > 0 ? dl : 0; }
    radiusRec(nearc, q, r2, out, mind2, off);
    float old = off[d];
    float far2 = mind2 - old * old + cut * cut;
    if (far2 * 0.999999f <= r2) {
      off[d] = cut;
      radiusRec(farc, q, r2, out, far2 > 0 ? far2 : 0, off);
      off[d] = old;
    }
  }

  const float* base_ = nullptr;
  size_t stride_ = 3, n_ = 0;
  std::vector<int32_t> perm_;
  std::vector<Node> nodes_;
};

}  // namespace orc
