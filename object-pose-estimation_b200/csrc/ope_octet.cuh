// ope_octet.cuh — cooperative nearest-neighbour search: EIGHT LANES PER QUERY over the implicit octree of ope_grid.cuh.
//
// Why octets. One query per thread leaves a B200 mostly empty at the reference's cloud sizes (50 000 queries = 16 % of
// the resident-thread capacity) and walks the tree with data-dependent loops that serialise inside a warp (ncu on the
// thread-per-query version: ~1/13 SIMT efficiency, barrier stalls behind the slowest lane). An octet matches the
// octree's arity: at an internal node lane j tests child j (one 4-byte `start` load + one box distance per lane, the
// sibling's load gives the range end by shuffle), at a leaf the lanes test eight points per step (16-byte loads of
// consecutive float4). There is no divergence inside an octet, the dependent-load chain per query shrinks from ~100
// node visits to ~20 steps, and there are 8x more threads to hide L2 latency with.
//
// Traversal is depth-first with the nearest child on top of a per-octet stack in shared memory (64 entries x 16 B:
// node, lower bound, point range), pruned against the octet-wide current bound; entries are re-checked when popped.
// All warp primitives are scoped to the octet's own 8-lane mask, so the four octets of a warp are independent.
// Results are bit-identical to ope::grid_nn1 / ope::grid_knn (same exact distances, same (d2, index) tie order, the
// same conservative bounds), which the host-compiled tests check against the CPU oracle.
#pragma once
#include "ope_grid.cuh"

namespace ope {

static constexpr int kOctStack = 64;   // >= 7 * OPE_MAX_BITS + 1
static constexpr int kOctLeaf = 32;    // nodes with at most this many points are scanned, 8 points per step

struct __align__(16) OctEntry {
  unsigned node;  // level << 27 | code
  float d2;       // conservative lower bound of the node's distance to the query
  int b, e;       // point range
};
struct OctStack { OctEntry s[kOctStack]; };

// top-k list of one octet (k <= 32), ascending (d2, index), in shared memory
struct OctKnnList { float d[32]; int i[32]; };

struct Octet {
  unsigned sub;    // lane within the octet, 0..7
  unsigned obase;  // first lane of the octet within the warp: 0, 8, 16, 24
  unsigned mask;   // the octet's lanes
};
__device__ __forceinline__ Octet octet_self() {
  Octet o;
  const unsigned lane = threadIdx.x & 31u;
  o.sub = lane & 7u; o.obase = lane & 24u; o.mask = 0xffu << o.obase;
  return o;
}
__device__ __forceinline__ float octet_min(const Octet& o, float v) {
  v = fminf(v, __shfl_xor_sync(o.mask, v, 1));
  v = fminf(v, __shfl_xor_sync(o.mask, v, 2));
  v = fminf(v, __shfl_xor_sync(o.mask, v, 4));
  return v;
}
__device__ __forceinline__ int octet_sum(const Octet& o, int v) {
  v += __shfl_xor_sync(o.mask, v, 1);
  v += __shfl_xor_sync(o.mask, v, 2);
  v += __shfl_xor_sync(o.mask, v, 4);
  return v;
}

__device__ __forceinline__ void octet_push_root(const GridView& g, OctStack* st, const Octet& o, float ux, float uy, float uz) {
  if (o.sub == 0u) {
    OctEntry en;
    en.node = (unsigned)g.bits << 27; en.d2 = oct_box_d2(ux, uy, uz, 0, 0, 0, 1 << g.bits, g.h); en.b = 0; en.e = g.n;
    st->s[0] = en;
  }
  __syncwarp(o.mask);
}

// Expand an internal node: lane `sub` owns child `sub`; children that are non-empty and within `bound` are pushed,
// the nearest one on top. `split` is octet-uniform. Returns the number pushed (octet-uniform).
__device__ __forceinline__ int octet_expand(const GridView& g, OctStack* st, const Octet& o, int sp, bool split, unsigned node, int e,
                                            float ux, float uy, float uz, float bound) {
  const int level = (int)(node >> 27);
  const unsigned code = node & 0x07ffffffu;
  const int cl = level - 1;
  const unsigned cc = (code << 3) | o.sub;
  int cb = 0;
  if (split) cb = __ldg(g.start + ((size_t)cc << (3 * cl)));
  const int nb = __shfl_down_sync(o.mask, cb, 1, 8);
  const int ce = (o.sub == 7u) ? e : nb;
  float cd2 = FLT_MAX;
  bool pass = false;
  if (split && ce > cb) {
    cd2 = oct_node_d2(g, ux, uy, uz, cl, cc);
    pass = cd2 <= bound;
  }
  const unsigned om = (__ballot_sync(o.mask, pass) >> o.obase) & 0xffu;
  const int npass = __popc(om);
  // nearest passing child (ties: lower child index) goes on top of the stack
  float md = pass ? cd2 : FLT_MAX;
  unsigned mi = o.sub;
#pragma unroll
  for (int s = 1; s < 8; s <<= 1) {
    const float od = __shfl_xor_sync(o.mask, md, s);
    const unsigned oi = __shfl_xor_sync(o.mask, mi, s);
    if (od < md || (od == md && oi < mi)) { md = od; mi = oi; }
  }
  if (pass) {
    const unsigned others = om & ~(1u << mi);
    const int pos = (o.sub == mi) ? sp + npass - 1 : sp + __popc(others & ((1u << o.sub) - 1u));
    OctEntry en;
    en.node = ((unsigned)cl << 27) | cc; en.d2 = cd2; en.b = cb; en.e = ce;
    st->s[pos] = en;
  }
  __syncwarp(o.mask);
  return npass;
}

// Exact nearest neighbour for the octet's query. The 8 lanes of the octet must call this together with the same
// arguments; `active` false = no query. Returns the original index or -1, and its squared distance, in every lane.
__device__ __forceinline__ int octet_nn1(const GridView& g, OctStack* st, const Octet& o, bool active, float qx, float qy, float qz,
                                         float max_d2, float& out_d2) {
  const float ux = (qx - g.ox) * g.inv_h, uy = (qy - g.oy) * g.inv_h, uz = (qz - g.oz) * g.inv_h;
  float best_d2 = FLT_MAX;
  int best_i = 0x7fffffff;
  int sp = 0;
  if (active && g.n > 0) { octet_push_root(g, st, o, ux, uy, uz); sp = 1; }
  while (sp > 0) {
    --sp;
    const OctEntry en = st->s[sp];
    __syncwarp(o.mask);  // every lane holds the popped entry before its slot can be overwritten
    const float bound = fminf(octet_min(o, best_d2), max_d2);
    const bool live = en.d2 <= bound;
    const int level = (int)(en.node >> 27);
    const bool leaf = level == 0 || en.e - en.b <= kOctLeaf;
    if (live && leaf) {
      for (int i = en.b + (int)o.sub; i < en.e; i += 8) {
        const float4 p = __ldg(g.pts + i);
        const float d2 = dist2(qx, qy, qz, p.x, p.y, p.z);
        const int idx = __float_as_int(p.w);
        if (nb_less(d2, idx, best_d2, best_i)) { best_d2 = d2; best_i = idx; }
      }
    }
    sp += octet_expand(g, st, o, sp, live && !leaf, en.node, en.e, ux, uy, uz, bound);
  }
#pragma unroll
  for (int s = 1; s < 8; s <<= 1) {
    const float od = __shfl_xor_sync(o.mask, best_d2, s);
    const int oi = __shfl_xor_sync(o.mask, best_i, s);
    if (nb_less(od, oi, best_d2, best_i)) { best_d2 = od; best_i = oi; }
  }
  out_d2 = best_d2;
  return best_i == 0x7fffffff ? -1 : best_i;
}

// Insert (d2, idx) into the octet's sorted list (cnt entries, capacity k <= 32) when `doit`; the eight lanes move four
// slots each. All arguments octet-uniform; executed by all 8 lanes; returns the new count.
__device__ __forceinline__ int octet_list_insert(OctKnnList* L, const Octet& o, int cnt, int k, float d2, int idx, bool doit) {
  int less = 0;
  float od[4];
  int oi[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int s = (int)o.sub + 8 * r;
    od[r] = FLT_MAX; oi[r] = 0x7fffffff;
    if (doit && s < cnt) { od[r] = L->d[s]; oi[r] = L->i[s]; if (nb_less(od[r], oi[r], d2, idx)) ++less; }
  }
  const int p = octet_sum(o, less);  // entries that sort before the new one
  const int ncnt = !doit ? cnt : (cnt < k ? cnt + 1 : k);
  __syncwarp(o.mask);  // all old values are in registers
  if (doit) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int s = (int)o.sub + 8 * r;
      if (s < cnt && s >= p && s + 1 < ncnt) { L->d[s + 1] = od[r]; L->i[s + 1] = oi[r]; }
    }
    if (o.sub == 0u && p < ncnt) { L->d[p] = d2; L->i[p] = idx; }
  }
  __syncwarp(o.mask);
  return ncnt;
}

// Exact k nearest (k <= 32) for the octet's query into the octet's shared list, ascending (d2, index).
// Returns the count (min(k, n)), identical in every lane of the octet.
__device__ __forceinline__ int octet_knn(const GridView& g, OctStack* st, OctKnnList* L, const Octet& o, bool active, float qx,
                                         float qy, float qz, int k) {
  const float ux = (qx - g.ox) * g.inv_h, uy = (qy - g.oy) * g.inv_h, uz = (qz - g.oz) * g.inv_h;
  if (k > g.n) k = g.n;
  int cnt = 0;
  int sp = 0;
  if (active && g.n > 0 && k > 0) { octet_push_root(g, st, o, ux, uy, uz); sp = 1; }
  while (sp > 0) {
    --sp;
    const OctEntry en = st->s[sp];
    __syncwarp(o.mask);
    float bound = (cnt == k) ? L->d[k - 1] : FLT_MAX;
    const bool live = en.d2 <= bound;
    const int level = (int)(en.node >> 27);
    const bool leaf = level == 0 || en.e - en.b <= kOctLeaf;
    if (live && leaf) {
      // eight candidates per step; the ones that beat the current k-th are inserted one at a time, in lane order
      for (int base = en.b; base < en.e; base += 8) {
        const int i = base + (int)o.sub;
        float d2 = FLT_MAX;
        int idx = 0x7fffffff;
        if (i < en.e) {
          const float4 p = __ldg(g.pts + i);
          d2 = dist2(qx, qy, qz, p.x, p.y, p.z);
          idx = __float_as_int(p.w);
        }
        const bool cand = idx != 0x7fffffff && (cnt < k || nb_less(d2, idx, L->d[k - 1], L->i[k - 1]));
        unsigned cm = (__ballot_sync(o.mask, cand) >> o.obase) & 0xffu;
        while (cm != 0u) {
          const int src = __ffs(cm) - 1;
          cm &= cm - 1u;
          const float cd = __shfl_sync(o.mask, d2, src, 8);
          const int ci = __shfl_sync(o.mask, idx, src, 8);
          const bool still = cnt < k || nb_less(cd, ci, L->d[k - 1], L->i[k - 1]);
          cnt = octet_list_insert(L, o, cnt, k, cd, ci, still);
        }
      }
      bound = (cnt == k) ? L->d[k - 1] : FLT_MAX;
    }
    sp += octet_expand(g, st, o, sp, live && !leaf, en.node, en.e, ux, uy, uz, bound);
  }
  return cnt;
}

}  // namespace ope
