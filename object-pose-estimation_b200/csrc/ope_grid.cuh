// ope_grid.cuh — uniform-grid spatial index replacing pcl::KdTreeFLANN.
//
// Layout in HBM: points of the indexed cloud are counting-sorted by linear cell id (x fastest) into one
// float4 array {x, y, z, original_index_bits}; within a cell they are ordered by original index, so every
// traversal below is deterministic. cell_start[c] .. cell_start[c+1] delimits cell c, hence a run of cells
// x0..x1 of one (y,z) row is ONE contiguous range of float4 — the search loops read rows, not cells:
// 16-byte coalescable loads, two cell_start look-ups per row.
//
// Exactness: a query scans the cube of cells with Chebyshev radius r around its own cell; every point
// outside that cube is farther than r*h*(1 - slack) from the query (the binning function is monotone), so
// the search stops as soon as the current k-th best distance is inside that bound, and doubles r otherwise.
// Distances use ope::dist2 (FLANN L2_Simple order, no FMA) and ties break on the smaller original index —
// the oracle's canonical order — so indices are bit-exact against the CPU restatement.
#pragma once
#include "ope_device.cuh"

namespace ope {

struct GridView {
  float ox, oy, oz;   // origin = min corner of the indexed points
  float h, inv_h;     // cell edge
  int nx, ny, nz;
  int n;              // number of indexed (finite) points
  const int* cell_start;   // nx*ny*nz + 1
  const float4* pts;       // n, sorted by cell then original index; .w = original index bits
};

#if defined(__CUDACC__) && defined(__CUDA_ARCH__)
#define OPE_LDG(p) __ldg(p)
#else
#define OPE_LDG(p) (*(p))
#endif

OPE_HD int grid_cell_coord(float v, float o, float inv_h) { return (int)floorf((v - o) * inv_h); }

// relative slack of the "outside the cube" bound: binning rounding (<= 2.4e-7 cells per cell of offset,
// dims are capped at 4096) plus margin.
#define OPE_GRID_SLACK 4e-3f

OPE_HD int imin(int a, int b) { return a < b ? a : b; }
OPE_HD int imax(int a, int b) { return a > b ? a : b; }
OPE_HD int iabs(int a) { return a < 0 ? -a : a; }

// Visit every indexed point in the cube [c-r, c+r]^3 (clamped to the grid) that is NOT in the cube of
// radius r_prev (r_prev < 0: nothing excluded). f(px, py, pz, original_index) is called per point.
template <typename F>
OPE_HD void grid_visit_shell(const GridView& g, int cx, int cy, int cz, int r, int r_prev, F&& f) {
  const int z0 = imax(cz - r, 0), z1 = imin(cz + r, g.nz - 1);
  const int y0 = imax(cy - r, 0), y1 = imin(cy + r, g.ny - 1);
  const int x0 = imax(cx - r, 0), x1 = imin(cx + r, g.nx - 1);
  if (x0 > x1) return;
  for (int z = z0; z <= z1; ++z)
    for (int y = y0; y <= y1; ++y) {
      const int row = (z * g.ny + y) * g.nx;
      const bool inner = r_prev >= 0 && iabs(z - cz) <= r_prev && iabs(y - cy) <= r_prev;
      // up to two x segments: [x0, min(x1, cx-r_prev-1)] and [max(x0, cx+r_prev+1), x1]; or the full row
      int segs[4];
      int nseg = 0;
      if (!inner) { segs[0] = x0; segs[1] = x1; nseg = 1; }
      else {
        int a1 = imin(x1, cx - r_prev - 1);
        if (x0 <= a1) { segs[0] = x0; segs[1] = a1; nseg = 1; }
        int b0 = imax(x0, cx + r_prev + 1);
        if (b0 <= x1) { segs[2 * nseg] = b0; segs[2 * nseg + 1] = x1; ++nseg; }
      }
      for (int s = 0; s < nseg; ++s) {
        const int b = OPE_LDG(g.cell_start + row + segs[2 * s]);
        const int e = OPE_LDG(g.cell_start + row + segs[2 * s + 1] + 1);
        for (int i = b; i < e; ++i) {
          const float4 p = OPE_LDG(g.pts + i);
          f(p.x, p.y, p.z, f2i(p.w));
        }
      }
    }
}

// Chebyshev distance (in cells) from the query's (unclamped) cell to the grid box; 0 when inside.
OPE_HD int grid_outside_cells(const GridView& g, int cx, int cy, int cz) {
  int d = 0;
  d = imax(d, imax(-cx, cx - (g.nx - 1)));
  d = imax(d, imax(-cy, cy - (g.ny - 1)));
  d = imax(d, imax(-cz, cz - (g.nz - 1)));
  return d;
}
// smallest r whose cube covers the whole grid
OPE_HD int grid_cover_radius(const GridView& g, int cx, int cy, int cz) {
  int r = imax(iabs(cx), iabs(g.nx - 1 - cx));
  r = imax(r, imax(iabs(cy), iabs(g.ny - 1 - cy)));
  r = imax(r, imax(iabs(cz), iabs(g.nz - 1 - cz)));
  return r;
}

// Exact nearest neighbour (k = 1). Only neighbours with d2 <= max_d2 matter to the caller: the search may
// stop once everything unvisited is farther than that (pass FLT_MAX for an unbounded search).
// Returns original index or -1 (nothing indexed / nothing within the bound visited).
OPE_HD int grid_nn1(const GridView& g, float qx, float qy, float qz, float max_d2, float& best_d2) {
  best_d2 = FLT_MAX;
  int best_i = -1;
  if (g.n <= 0) return -1;
  const int cx = grid_cell_coord(qx, g.ox, g.inv_h), cy = grid_cell_coord(qy, g.oy, g.inv_h),
            cz = grid_cell_coord(qz, g.oz, g.inv_h);
  const int cover = grid_cover_radius(g, cx, cy, cz);
  int r = imax(1, grid_outside_cells(g, cx, cy, cz));
  int r_prev = -1;
  for (;;) {
    grid_visit_shell(g, cx, cy, cz, r, r_prev, [&](float px, float py, float pz, int idx) {
      float d2 = dist2(qx, qy, qz, px, py, pz);
      if (nb_less(d2, idx, best_d2, best_i < 0 ? 0x7fffffff : best_i)) { best_d2 = d2; best_i = idx; }
    });
    if (r >= cover) break;
    float bound = (float)r * g.h * (1.0f - OPE_GRID_SLACK);
    float bound2 = bound * bound;
    if (best_i >= 0 && best_d2 <= bound2) break;
    if (bound2 > max_d2) break;  // anything unvisited is beyond the caller's range
    r_prev = r;
    r = imin(r * 2, cover);
  }
  return best_i;
}

// Exact k nearest (k <= KMAX), ascending (d2, index) into bd/bi. Returns the count found (min(k, n)).
template <int KMAX>
OPE_HD int grid_knn(const GridView& g, float qx, float qy, float qz, int k, float* bd, int* bi) {
  if (g.n <= 0 || k <= 0) return 0;
  if (k > g.n) k = g.n;
  int cnt = 0;
  const int cx = grid_cell_coord(qx, g.ox, g.inv_h), cy = grid_cell_coord(qy, g.oy, g.inv_h),
            cz = grid_cell_coord(qz, g.oz, g.inv_h);
  const int cover = grid_cover_radius(g, cx, cy, cz);
  int r = imax(1, grid_outside_cells(g, cx, cy, cz));
  int r_prev = -1;
  for (;;) {
    grid_visit_shell(g, cx, cy, cz, r, r_prev, [&](float px, float py, float pz, int idx) {
      float d2 = dist2(qx, qy, qz, px, py, pz);
      if (cnt == k && !nb_less(d2, idx, bd[k - 1], bi[k - 1])) return;
      int j = cnt < k ? cnt : k - 1;  // insertion position search from the tail
      while (j > 0 && nb_less(d2, idx, bd[j - 1], bi[j - 1])) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; --j; }
      bd[j] = d2; bi[j] = idx;
      if (cnt < k) ++cnt;
    });
    if (r >= cover) break;
    float bound = (float)r * g.h * (1.0f - OPE_GRID_SLACK);
    if (cnt == k && bd[k - 1] <= bound * bound) break;
    r_prev = r;
    r = imin(r * 2, cover);
  }
  return cnt;
}

// Number of rings that certainly contain every point with distance < radius.
OPE_HD int grid_radius_rings(const GridView& g, float radius) {
  return (int)floorf(radius * g.inv_h * (1.0f + OPE_GRID_SLACK)) + 1;
}

}  // namespace ope
