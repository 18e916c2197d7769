// featgemm.cu — feature-space k-NN (K6: KdTreeFLANN<FPFHSignature33>::nearestKSearch for every SAC-IA source point) as a
// tcgen05 / TMA distance GEMM with a fused candidate selection, followed by an exact FP32 re-rank.
//
//   ||q - t||^2 = ||q||^2 + ||t||^2 - 2 q.t          the contraction q.t over all (query, target) pairs is a GEMM
//
// Indices must equal the exact float32 L2_Simple ranking bit for bit (north-star), so the tensor cores only NOMINATE:
//   1. prep kernels write bf16 operands with K padded to 128: a float x is split as hi = bf16(x), lo = bf16(x - hi);
//        A row (query)  = [ q_hi | q_hi | q_lo | 1 1 1 | 0.. ]
//        B row (target) = [ t_hi | t_lo | t_hi | p0 p1 p2 | 0.. ]      p = 3-way bf16 split of -||t||^2 / 2
//      so that one GEMM yields  s = q.t - ||t||^2/2  with ~2^-16 relative error per product (the dropped q_lo.t_lo term
//      is 2^-18), and  ||q - t||^2 ~= ||q||^2 - 2 s.  Padding rows of B carry p0 = -1e30 and are never nominated.
//   2. featgemm_kernel: one CTA per 128 queries x a range of 256-target tiles. Warp 0 streams B tiles with TMA
//      (cp.async.bulk.tensor, 128-byte swizzle) through a 2-stage shared-memory ring, one elected thread of warp 1 issues
//      tcgen05.mma (M = 128, N = 256, K = 16 x 8) into one of two 256-column TMEM accumulators, warps 2-5 read the
//      finished accumulator with tcgen05.ld (16 epilogue warps: thread = query row x one of four 64-column groups) and keep
//      the 8 best scores per (query, group) in registers while the next tile is being multiplied. Nothing but 32 (index,
//      score) pairs per query and split leaves the SM.
//   3. featgemm_rerank_kernel recomputes the exact left-to-right float32 distance of every nominated target, ranks by
//      (distance, index), and PROVES completeness: every target that was not nominated in a candidate list has an approximate
//      distance >= that list's worst entry, hence an exact distance >= it - eps; if that does not exceed the k-th exact
//      distance the query is flagged and answered by the exact kernel instead (features.cu). eps (derived at its use) bounds the bf16-split
//      and accumulation error: 1e-4 * ||q|| * max||t|| + 1e-2.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "ope_host.cuh"

namespace ope {

static constexpr int FG_M = 128;        // queries per CTA (UMMA M)
static constexpr int FG_N = 256;        // targets per tile (UMMA N)
static constexpr int FG_K = 128;        // padded contraction length in bf16 elements
static constexpr int FG_KB = 64;        // elements per 128-byte swizzle row = one TMA box / k-block
static constexpr int FG_STAGES = 2;     // shared-memory stages of B, and TMEM accumulator stages
static constexpr int FG_GROUPS = 4;     // column groups of a tile: each TMEM lane quadrant is read by FG_GROUPS epilogue warps
static constexpr int FG_GCOLS = FG_N / FG_GROUPS;   // 64 columns per group
static constexpr int FG_CAND = 8;       // candidates kept per query per (split, column group): 32 per split
static constexpr int FG_EPI_WARPS = 4 * FG_GROUPS;
static constexpr int FG_THREADS = 64 + 32 * FG_EPI_WARPS;  // warp 0 TMA producer, warp 1 MMA issuer (+ TMEM allocation), 16 epilogue warps
static constexpr int FG_A_BYTES = FG_M * FG_K * 2;            // 32 KB
static constexpr int FG_B_STAGE_BYTES = FG_N * FG_K * 2;      // 64 KB
static constexpr int FG_SMEM_BYTES = 1024 + FG_A_BYTES + FG_STAGES * FG_B_STAGE_BYTES + 256;
static constexpr int FG_MAX_DIM = (FG_K - 3) / 3;             // 41

// ------------------------------------------------------------------------------------------------ PTX helpers ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded spin: a protocol bug must end as a launch failure, never as a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && spins > (1u << 26)) asm volatile("trap;");
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
// shared-memory matrix descriptor of a K-major, 128-byte-swizzled operand tile (rows of 128 bytes, 8-row groups 1024 bytes
// apart): start address >> 4 | LBO (ignored for swizzled K-major, 1) << 16 | SBO (1024 >> 4) << 32 | version 1 << 46 |
// SWIZZLE_128B (2) << 61
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = F32 (1 << 4), A = B = BF16 (1 << 7, 1 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
static constexpr uint32_t FG_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FG_N >> 3) << 17) | ((uint32_t)(FG_M >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(FG_IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
        "=r"(v[31])
      : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ prep kernels ----
// one thread per row: bf16 hi/lo split of the features into the padded operand row; A rows also produce ||q||^2
__global__ void featgemm_prep_kernel(const float* __restrict__ f, int n, int n_pad, int dim, int is_target, __nv_bfloat16* __restrict__ out,
                                     float* __restrict__ norm2, unsigned* __restrict__ max_norm2_bits) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_pad) return;
  __nv_bfloat16* o = out + (size_t)r * FG_K;
  const __nv_bfloat16 zero = __float2bfloat16(0.0f);
  for (int c = 0; c < FG_K; ++c) o[c] = zero;
  if (r >= n) {
    if (is_target) o[3 * dim] = __float2bfloat16(-1e30f);  // padding target: never nominated
    else if (norm2) norm2[r] = 0.0f;
    return;
  }
  float s = 0.0f;
  bool finite_row = true;
  for (int c = 0; c < dim; ++c) finite_row = finite_row && isfinite(__ldg(f + (size_t)r * dim + c));
  if (is_target && !finite_row) {   // a non-finite target row has no finite distance to anything: like padding, never nominated
    o[3 * dim] = __float2bfloat16(-1e30f);
    return;
  }
  for (int c = 0; c < dim; ++c) {
    const float x = __ldg(f + (size_t)r * dim + c);
    const __nv_bfloat16 hi = __float2bfloat16(x);
    const __nv_bfloat16 lo = __float2bfloat16(x - __bfloat162float(hi));
    s += x * x;
    if (is_target) { o[c] = hi; o[dim + c] = lo; o[2 * dim + c] = hi; }
    else { o[c] = hi; o[dim + c] = hi; o[2 * dim + c] = lo; }
  }
  if (is_target) {
    const float p = -0.5f * s;
    const __nv_bfloat16 p0 = __float2bfloat16(p);
    const float r1 = p - __bfloat162float(p0);
    const __nv_bfloat16 p1 = __float2bfloat16(r1);
    const __nv_bfloat16 p2 = __float2bfloat16(r1 - __bfloat162float(p1));
    o[3 * dim] = p0; o[3 * dim + 1] = p1; o[3 * dim + 2] = p2;
    if (s == s && s < 3.0e38f) atomicMax(max_norm2_bits, __float_as_uint(s));  // s >= 0: unsigned order == float order
  } else {
    const __nv_bfloat16 one = __float2bfloat16(1.0f);
    o[3 * dim] = one; o[3 * dim + 1] = one; o[3 * dim + 2] = one;
    norm2[r] = s;
  }
}

// ------------------------------------------------------------------------------------------------- GEMM kernel ----
__global__ void __launch_bounds__(FG_THREADS, 1)
    featgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int nq, int total_tiles,
                    int tiles_per_split, int* __restrict__ cand_idx, float* __restrict__ cand_score) {
  extern __shared__ uint8_t fg_smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)fg_smem_raw + 1023) & ~(uintptr_t)1023);  // swizzle-128B tiles need 1024-byte alignment
  uint8_t* sA = base;
  uint8_t* sB = base + FG_A_BYTES;
  uint64_t* bars = (uint64_t*)(sB + FG_STAGES * FG_B_STAGE_BYTES);
  // bars[0] a_full | [1..2] b_full | [3..4] b_empty | [5..6] t_full | [7..8] t_empty ; then the TMEM base address
  uint32_t* tmem_slot = (uint32_t*)(bars + 9);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x, split = blockIdx.y, n_splits = gridDim.y;
  const int tile0 = split * tiles_per_split;
  const int ntiles = max(0, min(tiles_per_split, total_tiles - tile0));
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  if (threadIdx.x == 0) {
    mbar_init(bar(0), 1);
    for (int s = 0; s < FG_STAGES; ++s) { mbar_init(bar(1 + s), 1); mbar_init(bar(3 + s), 1); mbar_init(bar(5 + s), 1); mbar_init(bar(7 + s), FG_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: 2 accumulators x 256 columns = the whole 512-column TMEM of this SM (one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0 && ntiles > 0) {
      // ===== TMA producer: the query tile once, then the target tiles through the ring =====
      mbar_expect_tx(bar(0), FG_A_BYTES);
      tma_load_2d(&tmA, bar(0), smem_u32(sA), 0, m_tile * FG_M);
      tma_load_2d(&tmA, bar(0), smem_u32(sA + FG_M * 128), FG_KB, m_tile * FG_M);
      for (int t = 0; t < ntiles; ++t) {
        const int s = t % FG_STAGES;
        const uint32_t ph = (uint32_t)(t / FG_STAGES) & 1u;
        mbar_wait(bar(3 + s), ph ^ 1u);  // slot free (passes at once the first time round)
        mbar_expect_tx(bar(1 + s), FG_B_STAGE_BYTES);
        uint8_t* dst = sB + (size_t)s * FG_B_STAGE_BYTES;
        tma_load_2d(&tmB, bar(1 + s), smem_u32(dst), 0, (tile0 + t) * FG_N);
        tma_load_2d(&tmB, bar(1 + s), smem_u32(dst + FG_N * 128), FG_KB, (tile0 + t) * FG_N);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && ntiles > 0) {
      // ===== MMA issuer: one thread =====
      mbar_wait(bar(0), 0);
      for (int t = 0; t < ntiles; ++t) {
        const int s = t % FG_STAGES;
        const uint32_t ph = (uint32_t)(t / FG_STAGES) & 1u;
        mbar_wait(bar(7 + s), ph ^ 1u);  // accumulator drained by the epilogue
        mbar_wait(bar(1 + s), ph);       // operands landed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB + (size_t)s * FG_B_STAGE_BYTES);
        const uint32_t d = tmem_base + (uint32_t)s * FG_N;
#pragma unroll
        for (int kb = 0; kb < FG_K / FG_KB; ++kb)
#pragma unroll
          for (int kk = 0; kk < FG_KB / 16; ++kk)
            umma_bf16(d, umma_desc(a0 + kb * (FG_M * 128) + kk * 32), umma_desc(b0 + kb * (FG_N * 128) + kk * 32), (kb | kk) != 0);
        umma_commit(bar(3 + s));  // frees the shared-memory slot when the MMAs have read it
        umma_commit(bar(5 + s));  // accumulator complete
      }
    }
  } else {
    // ===== epilogue: thread = (query row, column group); TMEM lanes 32*(warp % 4) .. +31 belong to this warp =====
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int qi = m_tile * FG_M + row;
    float bs[FG_CAND];
    int bi[FG_CAND];
#pragma unroll
    for (int c = 0; c < FG_CAND; ++c) { bs[c] = -INFINITY; bi[c] = -1; }
    for (int t = 0; t < ntiles; ++t) {
      const int s = t % FG_STAGES;
      const uint32_t ph = (uint32_t)(t / FG_STAGES) & 1u;
      mbar_wait(bar(5 + s), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)s * FG_N + (uint32_t)(grp * FG_GCOLS);
      const int col0 = (tile0 + t) * FG_N + grp * FG_GCOLS;
      for (int c0 = 0; c0 < FG_GCOLS; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);
        // Hot path: 31 max operations and one vote per 32 columns. Only when some row of the warp sees a score above its
        // current 16th best does the warp walk the columns; the votes are warp-uniform, so the insertion stays a real,
        // rarely taken branch instead of 256 predicated copies of it.
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = fmaxf(fmaxf(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), fmaxf(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
        const float mx = fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));   // a tree, not a 31-deep chain
        if (__any_sync(0xffffffffu, mx > bs[FG_CAND - 1])) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float sc = __uint_as_float(v[j]);
            const bool c = sc > bs[FG_CAND - 1];
            if (__any_sync(0xffffffffu, c)) {
              if (c) {  // replace the worst, bubble up (static indexing)
                bs[FG_CAND - 1] = sc; bi[FG_CAND - 1] = col0 + c0 + j;
#pragma unroll
                for (int p = FG_CAND - 1; p > 0; --p)
                  if (bs[p] > bs[p - 1]) {
                    const float ts = bs[p]; bs[p] = bs[p - 1]; bs[p - 1] = ts;
                    const int ti = bi[p]; bi[p] = bi[p - 1]; bi[p - 1] = ti;
                  }
              }
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(7 + s));
    }
    if (qi < nq) {
      const size_t o = (((size_t)qi * n_splits + split) * FG_GROUPS + grp) * FG_CAND;
#pragma unroll
      for (int c = 0; c < FG_CAND; ++c) { cand_idx[o + c] = bi[c]; cand_score[o + c] = bs[c]; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// ----------------------------------------------------------------------------------------------- exact re-rank ----
static constexpr int kRrMaxK = 16;
__global__ void featgemm_rerank_kernel(const float* __restrict__ ftgt, int nt, const float* __restrict__ fqry, int nq, int dim, int k,
                                       int n_splits, int tiles_per_split, const int* __restrict__ cand_idx,
                                       const float* __restrict__ cand_score, const float* __restrict__ qnorm2,
                                       const unsigned* __restrict__ max_tnorm2_bits, int* __restrict__ out_idx, float* __restrict__ out_d2,
                                       int* __restrict__ flags) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  float bd[kRrMaxK];
  int bi[kRrMaxK];
  int cnt = 0;
  const float* q = fqry + (size_t)qi * dim;
  const int n_lists = n_splits * FG_GROUPS;
  for (int s = 0; s < n_lists; ++s)
    for (int c = 0; c < FG_CAND; ++c) {
      const int idx = cand_idx[((size_t)qi * n_lists + s) * FG_CAND + c];
      if (idx < 0 || idx >= nt) continue;
      const float* f = ftgt + (size_t)idx * dim;
      float d = 0.0f;
      for (int e = 0; e < dim; ++e) { const float df = q[e] - f[e]; d += df * df; }  // FLANN L2_Simple order (-fmad=false)
      if (!isfinite(d)) continue;
      if (cnt == k && !nb_less(d, idx, bd[k - 1], bi[k - 1])) continue;
      int j = cnt < k ? cnt : k - 1;
      while (j > 0 && nb_less(d, idx, bd[j - 1], bi[j - 1])) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; --j; }
      bd[j] = d; bi[j] = idx;
      if (cnt < k) ++cnt;
    }
  // completeness proof: in a split that had more real targets than candidate slots, everything that was not nominated has an
  // approximate distance >= the split's worst nominated one
  const float qn2 = qnorm2[qi];
  // Error budget of the approximate score s = q.t - |t|^2/2, hence of d = |q|^2 - 2s (Cauchy-Schwarz over the dim products):
  //   operands: x = hi + lo with |x - hi - lo| <= 2^-17 |x| (two bf16 roundings), and the product q_lo * t_lo is dropped:
  //             |delta s| <= (2 * 2^-17 + 2^-16) |q||t| = 2^-15 |q||t| = 3.1e-5 |q||t|
  //   tensor-core accumulation of K = 128 float terms, possibly truncating: <= 128 * 2^-23 * 3 |q||t| = 4.6e-5 |q||t|
  //   norm columns: 3-way bf16 split of -|t|^2/2, residual 2^-25 relative: negligible
  // => |delta d| <= 2 * 7.7e-5 |q||t| = 1.5e-4 |q||t|. The proof uses 4e-4 |q| max|t| (2.6x that) plus an absolute 1e-2 for
  // tiny norms; a query whose margin is thinner falls back to the exact kernel, which costs time, never correctness.
  const float eps = 4e-4f * sqrtf(qn2) * sqrtf(__uint_as_float(*max_tnorm2_bits)) + 1e-2f;
  bool unsafe = false;
  for (int s = 0; s < n_lists; ++s) {
    // a list whose worst entry is not a real, finite candidate was never full: everything finite in its column group was
    // nominated and there is nothing to prove
    const float worst = cand_score[((size_t)qi * n_lists + s) * FG_CAND + FG_CAND - 1];
    const int widx = cand_idx[((size_t)qi * n_lists + s) * FG_CAND + FG_CAND - 1];
    if (widx < 0 || widx >= nt || !(worst > -1e29f)) continue;
    const float d_out = qn2 - 2.0f * worst;       // approximate distance of the worst nominated target of this list
    if (cnt < k || !(d_out - eps > bd[k - 1])) unsafe = true;
  }
  // a NaN query has no finite distance to anything: exact semantics = no neighbours, nothing to prove
  flags[qi] = (unsafe && qn2 == qn2) ? 1 : 0;
  for (int j = 0; j < k; ++j) {
    out_idx[(size_t)qi * k + j] = j < cnt ? bi[j] : -1;
    if (out_d2) out_d2[(size_t)qi * k + j] = j < cnt ? bd[j] : INFINITY;
  }
}
__global__ void flagged_list_kernel(const int* __restrict__ flags, const int* __restrict__ pos, int n, int* __restrict__ list) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flags[i]) list[pos[i]] = i;
}

// ======================================================================================================== host ==
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static const EncodeTiledFn fn = []() -> EncodeTiledFn {   // initialised once, thread-safe (batch workers call this concurrently)
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      return (EncodeTiledFn)p;
    return nullptr;
  }();
  return fn;
}
// [rows, 128] bf16 row-major operand, box = 64 elements (128 bytes) x box_rows, 128-byte swizzle
static int make_operand_map(ope_ctx* ctx, const void* base, size_t rows, int box_rows, CUtensorMap* out) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(ctx, OPE_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t gdim[2] = {(cuuint64_t)FG_K, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)FG_K * 2};
  const cuuint32_t box[2] = {(cuuint32_t)FG_KB, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, OPE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return OPE_OK;
}

int feature_knn_exact_device(ope_ctx* ctx, const float* d_ftgt, size_t nt, const float* d_fqry, size_t nq, int dim, int k, const int* d_qlist,
                             size_t n_list, int* d_idx, float* d_d2);

bool feature_knn_gemm_applicable(size_t nt, size_t nq, int dim, int k) {
  if (dim > FG_MAX_DIM || k > kRrMaxK || k > FG_CAND) return false;
  const char* mode = std::getenv("OPE_FEATURE_KNN");
  if (mode && std::strcmp(mode, "exact") == 0) return false;
  if (mode && std::strcmp(mode, "gemm") == 0) return nt > 0 && nq > 0;
  return nt >= 1024 && nq >= 256 && (double)nt * (double)nq >= 1048576.0;  // below this the launch latency of either path dominates
}

// d_idx / d_d2: nq * k. *n_fallback (host) receives the number of queries answered by the exact kernel.
int feature_knn_gemm_device(ope_ctx* ctx, const float* d_ftgt, size_t nt, const float* d_fqry, size_t nq, int dim, int k, int* d_idx,
                            float* d_d2, int* n_fallback) {
  if (n_fallback) *n_fallback = 0;
  const size_t nq_pad = (nq + FG_M - 1) / FG_M * FG_M, nt_pad = (nt + FG_N - 1) / FG_N * FG_N;
  const int total_tiles = (int)(nt_pad / FG_N), m_tiles = (int)(nq_pad / FG_M);
  // splits of the target range: fill the SMs with ONE wave of CTAs (m_tiles * splits <= SM count), at least 4 tiles each
  int splits = std::max(1, std::min(ctx->sm_count / std::max(m_tiles, 1), (total_tiles + 3) / 4));
  splits = std::min(splits, 64);
  const int tiles_per_split = (total_tiles + splits - 1) / splits;
  splits = (total_tiles + tiles_per_split - 1) / tiles_per_split;
  Scratch<__nv_bfloat16> A(ctx), B(ctx);
  Scratch<float> qn2(ctx), cscore(ctx);
  Scratch<unsigned> tmax(ctx);
  Scratch<int> cidx(ctx), flags(ctx), list(ctx);
  OPE_TRY(A.alloc(nq_pad * FG_K)); OPE_TRY(B.alloc(nt_pad * FG_K)); OPE_TRY(qn2.alloc(nq_pad)); OPE_TRY(tmax.alloc(1));
  OPE_TRY(cidx.alloc(nq * splits * FG_GROUPS * FG_CAND)); OPE_TRY(cscore.alloc(nq * splits * FG_GROUPS * FG_CAND)); OPE_TRY(flags.alloc(nq + 1));
  OPE_CUDA_TRY(ctx, cudaMemsetAsync(tmax.p, 0, sizeof(unsigned), ctx->stream));
  featgemm_prep_kernel<<<div_up(nq_pad, 128), 128, 0, ctx->stream>>>(d_fqry, (int)nq, (int)nq_pad, dim, 0, A.p, qn2.p, nullptr);
  OPE_TRY(check_launch(ctx, "featgemm_prep_kernel"));
  featgemm_prep_kernel<<<div_up(nt_pad, 128), 128, 0, ctx->stream>>>(d_ftgt, (int)nt, (int)nt_pad, dim, 1, B.p, nullptr, tmax.p);
  OPE_TRY(check_launch(ctx, "featgemm_prep_kernel"));
  CUtensorMap tmA, tmB;
  OPE_TRY(make_operand_map(ctx, A.p, nq_pad, FG_M, &tmA));
  OPE_TRY(make_operand_map(ctx, B.p, nt_pad, FG_N, &tmB));
  OPE_TRY(dyn_smem(ctx, (const void*)featgemm_kernel, FG_SMEM_BYTES));
  cudaEventRecord(ctx->kev[2][0], ctx->stream);
  featgemm_kernel<<<dim3(m_tiles, splits), FG_THREADS, FG_SMEM_BYTES, ctx->stream>>>(tmA, tmB, (int)nq, total_tiles, tiles_per_split, cidx.p,
                                                                                      cscore.p);
  cudaEventRecord(ctx->kev[2][1], ctx->stream);
  ctx->kev_valid[2] = true;
  OPE_TRY(check_launch(ctx, "featgemm_kernel"));
  OPE_CUDA_TRY(ctx, cudaMemsetAsync(flags.p + nq, 0, sizeof(int), ctx->stream));
  featgemm_rerank_kernel<<<div_up(nq, 128), 128, 0, ctx->stream>>>(d_ftgt, (int)nt, d_fqry, (int)nq, dim, k, splits, tiles_per_split, cidx.p,
                                                                    cscore.p, qn2.p, tmax.p, d_idx, d_d2, flags.p);
  OPE_TRY(check_launch(ctx, "featgemm_rerank_kernel"));
  // queries whose candidate set could not be proven complete: exact kernel
  Scratch<int> pos(ctx);
  OPE_TRY(pos.alloc(nq + 1));
  OPE_CUDA_TRY(ctx, cudaMemcpyAsync(pos.p, flags.p, (nq + 1) * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
  OPE_TRY(exclusive_scan_i32(ctx, pos.p, nq + 1));
  void* h;
  OPE_TRY(read_back(ctx, pos.p + nq, sizeof(int), &h));
  const int n_flagged = *(const int*)h;
  if (n_fallback) *n_fallback = n_flagged;
  ctx->feature_knn_fallbacks += n_flagged;
  ctx->feature_knn_gemm_queries += (int64_t)nq;
  if (n_flagged > 0) {
    OPE_TRY(list.alloc((size_t)n_flagged));
    flagged_list_kernel<<<div_up(nq, 256), 256, 0, ctx->stream>>>(flags.p, pos.p, (int)nq, list.p);
    OPE_TRY(check_launch(ctx, "flagged_list_kernel"));
    OPE_TRY(feature_knn_exact_device(ctx, d_ftgt, nt, d_fqry, nq, dim, k, list.p, (size_t)n_flagged, d_idx, d_d2));
  }
  return OPE_OK;
}

}  // namespace ope
