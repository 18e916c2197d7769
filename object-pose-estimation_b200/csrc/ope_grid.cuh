// ope_grid.cuh — voxel-grid spatial index replacing pcl::KdTreeFLANN.
//
// Layout in HBM: the indexed cloud is binned into a 2^bits x 2^bits x 2^bits grid of cubic cells of edge h and
// counting-sorted by the MORTON code of its cell into one float4 array {x, y, z, original_index_bits}; points of one
// cell are ordered by original index, so every traversal is deterministic. `start` is the exclusive prefix sum of
// the per-code counts (2^(3*bits) + 1 int32). Because the order is Morton, every aligned 2^l-cube of cells — every
// node of the implicit octree over the grid — is ONE contiguous range of float4:
//     node (level l, code m)  ->  points [ start[m << 3l], start[(m+1) << 3l] )
// so the tree needs no node storage, empty space is skipped at the coarsest level that is empty, and a leaf is a
// run of 16-byte loads.
//
// Search = pruned depth-first traversal, nearest child first. A node is skipped when a conservative lower bound of
// its distance to the query exceeds the current bound (k-th best distance, radius, or the caller's range limit).
// The bound is evaluated in "cell units" u = (q - o) * inv_h with the same float expression that binned the points,
// which is monotone, so a point stored in cell c has u_p in [c, c+1) by construction; OPE_CELL_SLACK absorbs the
// rounding of the subtraction. Distances of points use ope::dist2 (FLANN L2_Simple order, no FMA) and ties break on
// the smaller original index — the oracle's canonical order — so indices are bit-exact against the CPU restatement
// for near and far queries alike.
#pragma once
#include "ope_device.cuh"

namespace ope {

struct GridView {
  float ox, oy, oz;   // origin = min corner of the indexed points
  float h, inv_h;     // cell edge
  int bits;           // octree depth: 2^bits cells per axis
  int n;              // number of indexed (finite) points
  const int* start;   // 2^(3*bits) + 1, indexed by Morton code
  const float4* pts;  // n, sorted by Morton code then original index; .w = original index bits
};

#if defined(__CUDACC__) && defined(__CUDA_ARCH__)
#define OPE_LDG(p) __ldg(p)
#else
#define OPE_LDG(p) (*(p))
#endif

#define OPE_CELL_SLACK 4e-4f      // cells: rounding of (p - o) * inv_h on both sides, grids are at most 512 cells wide
#define OPE_MAX_BITS 9
#ifndef OPE_NN1_LEAF
#define OPE_NN1_LEAF 8
#endif
#ifndef OPE_COUNT
#define OPE_COUNT(what)   // tests/hostemu defines this to count node visits / point tests
#endif

OPE_HD int imin(int a, int b) { return a < b ? a : b; }
OPE_HD int imax(int a, int b) { return a > b ? a : b; }

// spread the low 10 bits of v so that there are two zero bits between each
OPE_HD unsigned part1by2(unsigned v) {
  v &= 0x000003ffu;
  v = (v ^ (v << 16)) & 0xff0000ffu;
  v = (v ^ (v << 8)) & 0x0300f00fu;
  v = (v ^ (v << 4)) & 0x030c30c3u;
  v = (v ^ (v << 2)) & 0x09249249u;
  return v;
}
OPE_HD unsigned compact1by2(unsigned v) {
  v &= 0x09249249u;
  v = (v ^ (v >> 2)) & 0x030c30c3u;
  v = (v ^ (v >> 4)) & 0x0300f00fu;
  v = (v ^ (v >> 8)) & 0xff0000ffu;
  v = (v ^ (v >> 16)) & 0x000003ffu;
  return v;
}
OPE_HD unsigned morton3(unsigned x, unsigned y, unsigned z) { return part1by2(x) | (part1by2(y) << 1) | (part1by2(z) << 2); }

// cell coordinate of a point along one axis (the binning function; clamped to the grid)
OPE_HD int grid_cell_coord(float v, float o, float inv_h, int bits) {
  int c = (int)floorf((v - o) * inv_h);
  return imin(imax(c, 0), (1 << bits) - 1);
}
OPE_HD unsigned grid_cell_code(const GridView& g, float x, float y, float z) {
  return morton3((unsigned)grid_cell_coord(x, g.ox, g.inv_h, g.bits), (unsigned)grid_cell_coord(y, g.oy, g.inv_h, g.bits),
                 (unsigned)grid_cell_coord(z, g.oz, g.inv_h, g.bits));
}

// conservative squared distance (metric units) from the query (cell units ux,uy,uz) to the box of cells
// [cx*s, (cx+1)*s) x ... ; never larger than the distance to any point stored in that box.
OPE_HD float oct_box_d2(float ux, float uy, float uz, int cx, int cy, int cz, int s, float h) {
  const float fs = (float)s;
  const float lx = (float)cx * fs, ly = (float)cy * fs, lz = (float)cz * fs;
  float dx = fmaxf(fmaxf(lx - ux, ux - (lx + fs)), 0.0f);
  float dy = fmaxf(fmaxf(ly - uy, uy - (ly + fs)), 0.0f);
  float dz = fmaxf(fmaxf(lz - uz, uz - (lz + fs)), 0.0f);
  dx = fmaxf(dx - OPE_CELL_SLACK, 0.0f);
  dy = fmaxf(dy - OPE_CELL_SLACK, 0.0f);
  dz = fmaxf(dz - OPE_CELL_SLACK, 0.0f);
  const float d = (dx * dx + dy * dy + dz * dz) * (h * h);
  return d * 0.99999f;
}

// ---- implicit octree nodes --------------------------------------------------------------------------------
OPE_HD void oct_node_range(const GridView& g, int level, unsigned code, int& b, int& e) {
  const int shift = 3 * level;
  b = OPE_LDG(g.start + ((size_t)code << shift));
  e = OPE_LDG(g.start + ((size_t)(code + 1u) << shift));
}
OPE_HD float oct_node_d2(const GridView& g, float ux, float uy, float uz, int level, unsigned code) {
  return oct_box_d2(ux, uy, uz, (int)compact1by2(code), (int)compact1by2(code >> 1), (int)compact1by2(code >> 2), 1 << level,
                    g.h);
}
// Stackless pruned depth-first traversal of the subtree rooted at (root_level, root_code), children in Morton order.
// bound(): current squared-distance bound — nodes whose conservative lower bound EXCEEDS it are skipped (it may shrink
// while ranges are consumed). range(b, e): consume the points g.pts[b..e) of a node that is a single cell or holds at
// most `leaf` points. The node (skip_level, skip_code) is not entered (it was consumed by the seed step). The only
// state is (level, code): no stack, no local memory. Control flow depends only on (g, u, bound()), so a warp that
// shares one query runs it uniformly.
template <typename BoundF, typename RangeF>
OPE_HD void oct_traverse(const GridView& g, float ux, float uy, float uz, int root_level, unsigned root_code, int skip_level,
                         unsigned skip_code, int leaf, BoundF&& bound, RangeF&& range) {
  int level = root_level;
  unsigned code = root_code;
  for (;;) {
    bool descend = false;
    OPE_COUNT(0);
    if (!(level == skip_level && code == skip_code)) {
      if (oct_node_d2(g, ux, uy, uz, level, code) <= bound()) {
        int b, e;
        oct_node_range(g, level, code, b, e);
        OPE_COUNT(1);
        if (e > b) {
          if (level == 0 || e - b <= leaf) { OPE_COUNT(2); range(b, e); }
          else descend = true;
        }
      }
    }
    if (descend) { --level; code <<= 3; continue; }
    for (;;) {  // next node in depth-first order
      if (level == root_level) return;
      if ((code & 7u) != 7u) { ++code; break; }
      code >>= 3; ++level;
    }
  }
}

// Greedy seed: walk from the root toward the query, at each level into the nearest child that still holds at least
// `need` points, and stop at the first node that is a single cell, holds at most `leaf` points, or has no such child.
// need = 1, leaf >= 1 for nearest-neighbour; need = k, leaf < k for k-NN (then no ancestor of the seed can be a leaf).
OPE_HD void oct_seed(const GridView& g, float ux, float uy, float uz, int need, int leaf, int& level, unsigned& code, int& b, int& e) {
  level = g.bits; code = 0u;
  oct_node_range(g, level, code, b, e);
  while (level > 0 && e - b > leaf) {
    const int cl = level - 1;
    const int shift = 3 * cl;
    float best_d2 = FLT_MAX;
    int best_j = -1, bb = 0, be = 0;
    int prev = OPE_LDG(g.start + ((size_t)(code << 3) << shift));
    for (unsigned j = 0; j < 8u; ++j) {
      const unsigned cc = (code << 3) | j;
      const int nxt = OPE_LDG(g.start + ((size_t)(cc + 1u) << shift));
      if (nxt - prev >= need) {
        const float d2 = oct_node_d2(g, ux, uy, uz, cl, cc);
        if (d2 < best_d2) { best_d2 = d2; best_j = (int)j; bb = prev; be = nxt; }
      }
      prev = nxt;
    }
    if (best_j < 0) break;
    level = cl; code = (code << 3) | (unsigned)best_j; b = bb; e = be;
  }
}

// ---- ball -> nodes --------------------------------------------------------------------------------------------
// The cells that can hold a point within squared metric distance r2 of the query, expressed as at most 2x2x2 nodes of
// the SMALLEST level L at which the ball's cell range spans at most two nodes per axis. Unlike one enclosing ancestor
// (which degenerates to the root whenever the ball straddles a high-level boundary) this start set is always local:
// near queries get <= 8 single cells, far ones <= 8 coarse nodes that the pruned traversal then descends.
struct BallNodes {
  int L;        // node level
  int nlo[3];   // node coordinates (level L) of the low corner
  int span[3];  // 0 or 1 extra node along each axis
  bool hit;     // false: the ball misses the grid entirely
};
OPE_HD BallNodes ball_nodes(const GridView& g, float ux, float uy, float uz, float r2) {
  BallNodes B;
  B.hit = true;
  if (!(r2 < FLT_MAX)) {
    B.L = g.bits;
    for (int d = 0; d < 3; ++d) { B.nlo[d] = 0; B.span[d] = 0; }
    return B;
  }
  const float fm = (float)((1 << g.bits) - 1);
  const float rho = sqrtf(r2) * g.inv_h * 1.0001f + 2.0f * OPE_CELL_SLACK;
  const float u[3] = {ux, uy, uz};
  int lo[3], hi[3];
  for (int d = 0; d < 3; ++d) {
    const float a = floorf(u[d] - rho), b = floorf(u[d] + rho);
    if (b < 0.0f || a > fm) B.hit = false;
    lo[d] = (int)fminf(fmaxf(a, 0.0f), fm);
    hi[d] = (int)fminf(fmaxf(b, 0.0f), fm);
  }
  int L = 0;
  while (L < g.bits && (((hi[0] >> L) - (lo[0] >> L)) > 1 || ((hi[1] >> L) - (lo[1] >> L)) > 1 || ((hi[2] >> L) - (lo[2] >> L)) > 1)) ++L;
  B.L = L;
  for (int d = 0; d < 3; ++d) { B.nlo[d] = lo[d] >> L; B.span[d] = (hi[d] >> L) - (lo[d] >> L); }
  return B;
}
// j-th node (j in 0..7: bit 0 = x, bit 1 = y, bit 2 = z) of the ball's start set; false when it does not exist
OPE_HD bool ball_node(const BallNodes& B, int j, unsigned& code) {
  const int dx = j & 1, dy = (j >> 1) & 1, dz = j >> 2;
  if (!B.hit || dx > B.span[0] || dy > B.span[1] || dz > B.span[2]) return false;
  code = morton3((unsigned)(B.nlo[0] + dx), (unsigned)(B.nlo[1] + dy), (unsigned)(B.nlo[2] + dz));
  return true;
}
// conservative squared distance to the j-th start node, from its coordinates (no Morton decode)
OPE_HD float ball_node_d2(const GridView& g, const BallNodes& B, int j, float ux, float uy, float uz) {
  return oct_box_d2(ux, uy, uz, B.nlo[0] + (j & 1), B.nlo[1] + ((j >> 1) & 1), B.nlo[2] + (j >> 2), 1 << B.L, g.h);
}

// Seed for an unseeded nearest-neighbour query: climb from the query's own (clamped) cell to the first non-empty
// ancestor, descend greedily toward the query while the node is large, scan what is left. Cheap for queries near the
// indexed surface (one or two levels), bounded by 2*bits node visits for far ones.
#ifndef OPE_PROBE_LEAF
#define OPE_PROBE_LEAF 8
#endif
template <typename ScanF>
OPE_HD void nn1_probe(const GridView& g, float qx, float qy, float qz, float ux, float uy, float uz, ScanF&& scan) {
  unsigned code = grid_cell_code(g, qx, qy, qz);
  int level = 0, b, e;
  for (;;) {
    oct_node_range(g, level, code, b, e);
    OPE_COUNT(1);
    if (e > b || level == g.bits) break;
    code >>= 3; ++level;
  }
  while (level > 0 && e - b > OPE_PROBE_LEAF) {
    const int cl = level - 1;
    const int shift = 3 * cl;
    float best_d2 = FLT_MAX;
    int best_j = -1, bb = 0, be = 0;
    int prev = OPE_LDG(g.start + ((size_t)(code << 3) << shift));
    for (unsigned j = 0; j < 8u; ++j) {
      const unsigned cc = (code << 3) | j;
      const int nxt = OPE_LDG(g.start + ((size_t)(cc + 1u) << shift));
      if (nxt > prev) {
        const float d2 = oct_node_d2(g, ux, uy, uz, cl, cc);
        if (d2 < best_d2) { best_d2 = d2; best_j = (int)j; bb = prev; be = nxt; }
      }
      prev = nxt;
    }
    OPE_COUNT(1);
    level = cl; code = (code << 3) | (unsigned)best_j; b = bb; e = be;
  }
  scan(b, imin(e, b + 4 * OPE_PROBE_LEAF));  // a seed only: any point is a valid upper bound
}

// ---- nearest neighbour with a certificate --------------------------------------------------------------------
// State of one query: the best candidate (d1, i1) in the canonical (d2, index) order and s2, the smallest squared
// distance seen among the OTHER points. A search that examines every point within the scan radius
//     r_s = min(sqrt(d1), sqrt(max_d2)) + gap
// ends with the exact nearest neighbour and with the certificate radius R^2 = min(s2, r_s^2): every indexed point other
// than i1 is at least R away. A caller that later moves the query by delta may keep i1 without searching as long as
// d(q', x_i1) + delta < R - delta (triangle inequality) — this is how the ICP loop skips most searches once the
// increments become small, while staying exact. gap = 0 reproduces the plain search.
struct Nn1State {
  float d1;  // best squared distance (FLT_MAX: none)
  int i1;    // its original index (0x7fffffff: none)
  float s2;  // smallest squared distance among the other points examined (FLT_MAX: none)
};
OPE_HD void nn1_offer(Nn1State& st, float d2, int idx) {
  if (idx == st.i1) return;  // the seed is met again during the scan
  if (nb_less(d2, idx, st.d1, st.i1)) { st.s2 = st.d1; st.d1 = d2; st.i1 = idx; }
  else st.s2 = fminf(st.s2, d2);
}
// squared scan radius for the current best (slightly inflated so that rounding can only enlarge the scanned set)
OPE_HD float nn1_scan_r2(float d1, float max_d2, float gap) {
  const float m = fminf(d1, max_d2);
  if (!(m < FLT_MAX)) return FLT_MAX;
  if (gap <= 0.0f) return m;
  const float r = sqrtf(m) + gap;
  return r * r * 1.000001f;
}

// Fast path: the whole candidate set of the scan radius, when it is at most OPE_NN1_FAST_MAX points in <= 8 nodes, is
// scanned directly (16 independent `start` loads, then runs of 16-byte point loads). Returns false — nothing scanned —
// when the set is larger; the caller then runs a pruned traversal (thread-level below, a warp per query on the device).
// On success *r2_out is the squared radius within which every point has been examined.
#ifndef OPE_NN1_FAST_MAX
#define OPE_NN1_FAST_MAX 96
#endif
OPE_HD bool nn1_fast(const GridView& g, float qx, float qy, float qz, float ux, float uy, float uz, float max_d2, float gap,
                     Nn1State& st, float* r2_out, int* total_out = nullptr, unsigned* node_mask_out = nullptr) {
  const float bound0 = nn1_scan_r2(st.d1, max_d2, gap);
  if (r2_out) *r2_out = bound0;
  if (total_out) *total_out = 0;
  if (node_mask_out) *node_mask_out = 0u;
  const BallNodes B = ball_nodes(g, ux, uy, uz, bound0);
  if (!B.hit) return true;
  const int sh = 3 * B.L;
  int nb[8], ne[8];
  unsigned nc[8];
  int total = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    nb[j] = ne[j] = 0; nc[j] = 0u;
    if (ball_node(B, j, nc[j])) {
      nb[j] = OPE_LDG(g.start + ((size_t)nc[j] << sh));
      ne[j] = OPE_LDG(g.start + ((size_t)(nc[j] + 1u) << sh));
      OPE_COUNT(1);
    }
    total += ne[j] - nb[j];
  }
  if (total_out) *total_out = total;
  if (node_mask_out) {
    unsigned m = 0u;
#pragma unroll
    for (int j = 0; j < 8; ++j) m |= (ne[j] > nb[j]) ? (1u << j) : 0u;
    *node_mask_out = m;
  }
  if (total > OPE_NN1_FAST_MAX) return false;
  // First two points of every node: 16 predicated, mutually independent 16-byte loads in straight-line code (one L2
  // round trip for the common case of a handful of points per cell), pruned against the radius the query arrived with.
  float4 pa[8], pb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (ne[j] > nb[j] && ball_node_d2(g, B, j, ux, uy, uz) > bound0) ne[j] = nb[j];  // node outside the ball
    pa[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    pb[j] = pa[j];
    if (ne[j] - nb[j] > 0) pa[j] = OPE_LDG(g.pts + nb[j]);
    if (ne[j] - nb[j] > 1) pb[j] = OPE_LDG(g.pts + nb[j] + 1);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (ne[j] - nb[j] > 0) { OPE_COUNT(2); nn1_offer(st, dist2(qx, qy, qz, pa[j].x, pa[j].y, pa[j].z), f2i(pa[j].w)); }
    if (ne[j] - nb[j] > 1) nn1_offer(st, dist2(qx, qy, qz, pb[j].x, pb[j].y, pb[j].z), f2i(pb[j].w));
  }
  // the rest (cells with more than two points), OPE_NN1_REST_WIDTH loads in flight (the offers stay in index order; measured on C2:
  // 2 -> 5.71 ms, 4 -> 5.53, 8 -> 5.55; deferring more of these queries to the warp queue, OPE_NN1_FAST_MAX 32, -> 6.87)
#ifndef OPE_NN1_REST_WIDTH
#define OPE_NN1_REST_WIDTH 4
#endif
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    for (int i = nb[j] + 2; i < ne[j]; i += OPE_NN1_REST_WIDTH) {
      float4 pw[OPE_NN1_REST_WIDTH];
#pragma unroll
      for (int u = 0; u < OPE_NN1_REST_WIDTH; ++u) pw[u] = OPE_LDG(g.pts + (i + u < ne[j] ? i + u : i));
#pragma unroll
      for (int u = 0; u < OPE_NN1_REST_WIDTH; ++u)
        if (i + u < ne[j]) nn1_offer(st, dist2(qx, qy, qz, pw[u].x, pw[u].y, pw[u].z), f2i(pw[u].w));
    }
  }
  return true;
}
// Medium queries: pruned, stackless depth-first traversal of the ball's start nodes by ONE thread (every thread of a warp
// runs its own; the work is a few dozen node visits and a hundred point tests). On return st is exact and *r2_out is the
// squared radius the certificate may rely on.
#ifndef OPE_NN1_DFS_LEAF
#define OPE_NN1_DFS_LEAF 12
#endif
OPE_HD void nn1_thread_dfs(const GridView& g, float qx, float qy, float qz, float ux, float uy, float uz, float max_d2, float gap,
                           Nn1State& st, float* r2_out) {
  const BallNodes B = ball_nodes(g, ux, uy, uz, nn1_scan_r2(st.d1, max_d2, gap));
  for (int j = 0; j < 8; ++j) {
    unsigned code;
    if (!ball_node(B, j, code)) continue;
    oct_traverse(g, ux, uy, uz, B.L, code, -1, 0u, OPE_NN1_DFS_LEAF, [&]() { return nn1_scan_r2(st.d1, max_d2, gap); },
                 [&](int b, int e) {
                   for (int i = b; i < e; ++i) {
                     const float4 p = OPE_LDG(g.pts + i);
                     nn1_offer(st, dist2(qx, qy, qz, p.x, p.y, p.z), f2i(p.w));
                   }
                 });
  }
  if (r2_out) *r2_out = nn1_scan_r2(st.d1, max_d2, gap);
}

// plain form (no certificate)
OPE_HD bool nn1_fast(const GridView& g, float qx, float qy, float qz, float ux, float uy, float uz, float max_d2, float& best_d2,
                     int& best_i) {
  Nn1State st;
  st.d1 = best_d2; st.i1 = best_i; st.s2 = FLT_MAX;
  const bool done = nn1_fast(g, qx, qy, qz, ux, uy, uz, max_d2, 0.0f, st, nullptr);
  best_d2 = st.d1; best_i = st.i1;
  return done;
}

// Exact nearest neighbour (k = 1). Only neighbours with d2 <= max_d2 matter to the caller (FLT_MAX: unbounded).
// seed_idx >= 0 names an indexed point (coordinates seed_pts[seed_idx], original order) used as the initial bound — e.g.
// the previous ICP iteration's match; the result is the same exact nearest neighbour either way.
// Returns the original index or -1 (nothing indexed); the caller rejects best_d2 > max_d2.
OPE_HD int grid_nn1(const GridView& g, float qx, float qy, float qz, float max_d2, float& best_d2, int seed_idx = -1,
                    const float4* seed_pts = nullptr, float gap = 0.0f, float* cert_r2 = nullptr) {
  best_d2 = FLT_MAX;
  if (cert_r2) *cert_r2 = 0.0f;
  Nn1State st;
  st.d1 = FLT_MAX; st.i1 = 0x7fffffff; st.s2 = FLT_MAX;
  if (g.n <= 0) return -1;
  const float ux = (qx - g.ox) * g.inv_h, uy = (qy - g.oy) * g.inv_h, uz = (qz - g.oz) * g.inv_h;
  auto scan = [&](int b, int e) {
    for (int i = b; i < e; ++i) {
      const float4 p = OPE_LDG(g.pts + i);
      nn1_offer(st, dist2(qx, qy, qz, p.x, p.y, p.z), f2i(p.w));
    }
  };
  if (seed_idx >= 0 && seed_pts) {
    const float4 s = OPE_LDG(seed_pts + seed_idx);
    if (finite3(s.x, s.y, s.z)) { st.d1 = dist2(qx, qy, qz, s.x, s.y, s.z); st.i1 = seed_idx; }
  }
  if (st.i1 == 0x7fffffff) nn1_probe(g, qx, qy, qz, ux, uy, uz, scan);
  st.s2 = FLT_MAX;  // the probe is only a seed: the search below establishes the second-best distance
  float r2 = 0.0f;
  if (!nn1_fast(g, qx, qy, qz, ux, uy, uz, max_d2, gap, st, &r2)) nn1_thread_dfs(g, qx, qy, qz, ux, uy, uz, max_d2, gap, st, &r2);
  best_d2 = st.d1;
  if (cert_r2) *cert_r2 = fminf(st.s2, r2);  // every indexed point other than the result is at least sqrt(this) away
  return st.i1 == 0x7fffffff ? -1 : st.i1;
}

// Exact k nearest (k <= KMAX), ascending (d2, index) into bd/bi. Returns the count found (min(k, n)).
template <int KMAX>
OPE_HD int grid_knn(const GridView& g, float qx, float qy, float qz, int k, float* bd, int* bi) {
  if (g.n <= 0 || k <= 0) return 0;
  if (k > g.n) k = g.n;
  int cnt = 0;
  const float ux = (qx - g.ox) * g.inv_h, uy = (qy - g.oy) * g.inv_h, uz = (qz - g.oz) * g.inv_h;
  auto scan = [&](int b, int e) {
    for (int i = b; i < e; ++i) {
      const float4 p = OPE_LDG(g.pts + i);
      const float d2 = dist2(qx, qy, qz, p.x, p.y, p.z);
      const int idx = f2i(p.w);
      if (cnt == k && !nb_less(d2, idx, bd[k - 1], bi[k - 1])) continue;
      int j = cnt < k ? cnt : k - 1;
      while (j > 0 && nb_less(d2, idx, bd[j - 1], bi[j - 1])) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; --j; }
      bd[j] = d2; bi[j] = idx;
      if (cnt < k) ++cnt;
    }
  };
  const int leaf = imin(k - 1, 16);  // < k: the seed node (>= k points) has no ancestor that is consumed whole
  int sl, b, e;
  unsigned sc;
  oct_seed(g, ux, uy, uz, k, leaf, sl, sc, b, e);
  scan(b, e);
  const BallNodes B = ball_nodes(g, ux, uy, uz, cnt == k ? bd[k - 1] : FLT_MAX);
  for (int j = 0; j < 8; ++j) {
    unsigned code;
    if (!ball_node(B, j, code)) continue;
    if (B.L <= sl && (code >> (3 * (sl - B.L))) == sc) continue;  // inside the seed node: already consumed
    oct_traverse(g, ux, uy, uz, B.L, code, sl, sc, leaf, [&]() { return cnt == k ? bd[k - 1] : FLT_MAX; }, scan);
  }
  return cnt;
}

// Radius traversal: range(b, e) is called for every leaf range that may hold points with d2 <= r2; the caller tests
// d2 < r2 per point. Starts from the ball's <= 8 start nodes (ball_nodes), each traversed with pruning.
template <typename RangeF>
OPE_HD void grid_radius_ranges(const GridView& g, float qx, float qy, float qz, float r2, int leaf, RangeF&& range) {
  if (g.n <= 0) return;
  const float ux = (qx - g.ox) * g.inv_h, uy = (qy - g.oy) * g.inv_h, uz = (qz - g.oz) * g.inv_h;
  const BallNodes B = ball_nodes(g, ux, uy, uz, r2);
  for (int j = 0; j < 8; ++j) {
    unsigned code;
    if (!ball_node(B, j, code)) continue;
    oct_traverse(g, ux, uy, uz, B.L, code, -1, 0u, leaf, [&]() { return r2; }, range);
  }
}

// Every indexed point with d2 < r2 (strict): f(px, py, pz, original_index, d2), in traversal order.
template <typename F>
OPE_HD void grid_radius_visit(const GridView& g, float qx, float qy, float qz, float r2, F&& f) {
  grid_radius_ranges(g, qx, qy, qz, r2, 32, [&](int b, int e) {
    for (int i = b; i < e; ++i) {
      const float4 p = OPE_LDG(g.pts + i);
      const float d2 = dist2(qx, qy, qz, p.x, p.y, p.z);
      if (d2 < r2) f(p.x, p.y, p.z, f2i(p.w), d2);
    }
  });
}

}  // namespace ope
