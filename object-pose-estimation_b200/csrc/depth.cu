// depth.cu — depth image -> point cloud on the device (the stage right before the registration path).
//
// Restates DataGrabber::rgbd2Pcl / depthToMeter (D&L/src/datagrabber.cpp:9-62,121-174) including its quirks: the pixel is
// visited columns-outer / rows-inner, the ROW index is fed as "x" and the COLUMN index as "y"
// (cloud.y = (row - cx) * Z / fx, cloud.x = (col - cy) * Z / fy), pixels with depth 0 or Z > z_max are dropped, and the
// output is the COMPACTED list in that traversal order (the organised width/height the reference sets are meaningless).
//
// HBM-bound byte shuffle: 2 bytes in per pixel, 16 bytes out per kept pixel. Two passes so that the output order is
// deterministic without atomics: (1) per-column valid counts, (2) a scan over the columns of every frame, (3) the write pass
// (kernels below).
// Frames are batched along blockIdx.y so that 1 024 frames (629 MB in, up to 5 GB out) are one launch.
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "ope_host.cuh"

namespace ope {

struct DepthParams {
  int rows, cols;
  float fx, fy, cx, cy, scale, z_max;
  float r_fx, r_fy, r_scale;   // correctly rounded reciprocals (host, IEEE division)
  int raw_lo, raw_hi;          // a pixel is kept iff raw_lo <= raw <= raw_hi (the reference's depth > 0, Z != 0, Z <= z_max tests,
                               // evaluated on the host for all 65 536 raw values with the reference's float arithmetic)
};

__device__ __forceinline__ bool depth_keep(const DepthParams& P, unsigned short raw) { return (int)raw >= P.raw_lo && (int)raw <= P.raw_hi; }

__device__ __forceinline__ bool depth_point(const DepthParams& P, int i, int j, unsigned short raw, float4& out) {
  if (!depth_keep(P, raw)) return false;
  const float Z = div_by_const((float)raw, P.scale, P.r_scale);
  const float X = div_by_const(((float)i - P.cx) * Z, P.fx, P.r_fx);  // p_FeatX = row (sic)
  const float Y = div_by_const(((float)j - P.cy) * Z, P.fy, P.r_fy);  // p_FeatY = column (sic)
  out = make_float4(Y, X, Z, 1.0f);
  return true;
}

// host: complete the parameter block (reciprocals, kept range of raw values)
static DepthParams depth_params(int rows, int cols, float fx, float fy, float cx, float cy, float scale, float z_max) {
  DepthParams P{rows, cols, fx, fy, cx, cy, scale, z_max, 0, 0, 0, 0, -1};
  P.r_fx = 1.0f / fx; P.r_fy = 1.0f / fy; P.r_scale = 1.0f / scale;
  // Z = (float)raw / scale is monotone in raw, so the kept set is an interval (remembered per thread: a camera keeps its scale)
  thread_local float c_scale = 0.0f, c_zmax = 0.0f;
  thread_local int c_lo = 0, c_hi = -1;
  if (c_scale == scale && c_zmax == z_max) { P.raw_lo = c_lo; P.raw_hi = c_hi; return P; }
  int lo = 65536, hi = -1;
  for (int raw = 1; raw <= 65535; ++raw) {
    volatile float Z = (float)raw / scale;
    if (Z == 0.0f || Z > z_max) continue;
    lo = std::min(lo, raw); hi = std::max(hi, raw);
  }
  P.raw_lo = lo; P.raw_hi = hi;
  c_scale = scale; c_zmax = z_max; c_lo = lo; c_hi = hi;
  return P;
}

// Pass 1: col_count[frame * cols + j] = number of kept pixels of column j. No transposition needed: lane = column, warp w
// walks rows w, w + 32, ... with all its loads independent (up to 16 in flight per thread); the 32 per-warp partial counts of
// a column are summed through shared memory.
__global__ void __launch_bounds__(1024) depth_count_kernel(const unsigned short* __restrict__ depth, DepthParams P, int* __restrict__ col_count) {
  __shared__ int part[32][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int frame = blockIdx.y;
  const int j = blockIdx.x * 32 + lane;
  const unsigned short* d = depth + (size_t)frame * P.rows * P.cols;
  int cnt = 0;
  if (j < P.cols) {
    for (int r0 = w; r0 < P.rows; r0 += 32 * 16) {
      unsigned short v[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) { const int i = r0 + 32 * t; v[t] = i < P.rows ? __ldg(d + (size_t)i * P.cols + j) : (unsigned short)0; }
#pragma unroll
      for (int t = 0; t < 16; ++t) cnt += (r0 + 32 * t < P.rows && depth_keep(P, v[t])) ? 1 : 0;
    }
  }
  part[w][lane] = cnt;
  __syncthreads();
  if (w == 0 && j < P.cols) {
    int s = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) s += part[k][lane];
    col_count[(size_t)frame * P.cols + j] = s;
  }
}

// Pass 2: col_start (exclusive scan of col_count over the whole batch) gives every column its output offset. A block owns 32
// adjacent columns of one frame and walks down the rows kChunk at a time: the chunk is loaded row-wise (64-byte segments, 8
// independent loads per thread), transposed through shared memory, and warp w compacts column w with ballots, writing 32
// consecutive float4 (512 B) per step.
static constexpr int kChunk = 256;
__global__ void __launch_bounds__(1024) depth_write_kernel(const unsigned short* __restrict__ depth, DepthParams P,
                                                           const int* __restrict__ col_start, float4* __restrict__ out) {
  __shared__ unsigned short tile[kChunk][34];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int frame = blockIdx.y;
  const int j0 = blockIdx.x * 32;
  const unsigned short* d = depth + (size_t)frame * P.rows * P.cols;
  const int j = j0 + w;  // this warp's column
  int run = 0;           // kept pixels of column j so far
  const int base = j < P.cols ? col_start[(size_t)frame * P.cols + j] : 0;
  for (int r0 = 0; r0 < P.rows; r0 += kChunk) {
    unsigned short v[kChunk / 32];
    const int cj = j0 + lane;
#pragma unroll
    for (int t = 0; t < kChunk / 32; ++t) {
      const int ri = r0 + w + 32 * t;
      v[t] = (ri < P.rows && cj < P.cols) ? __ldg(d + (size_t)ri * P.cols + cj) : (unsigned short)0;
    }
#pragma unroll
    for (int t = 0; t < kChunk / 32; ++t) tile[w + 32 * t][lane] = v[t];
    __syncthreads();
#pragma unroll
    for (int t = 0; t < kChunk / 32; ++t) {
      const int i = r0 + 32 * t + lane;  // transposed read: lane = row within the group, warp = column
      float4 pt = make_float4(0, 0, 0, 0);
      const bool keep = (i < P.rows && j < P.cols) && depth_point(P, i, j, tile[32 * t + lane][w], pt);
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) __stcs(out + (size_t)base + run + __popc(m & ((1u << lane) - 1u)), pt);   // streamed: never re-read by this kernel
      run += __popc(m);
    }
    __syncthreads();
  }
}

// Single pass (used whenever 32 columns of one frame fit in shared memory): the depth image is read ONCE. A block takes a ticket
// (tickets are handed out in output order: frame-major, then column block), loads its 32 columns x all rows into shared memory,
// counts the kept pixels per column, publishes its total, and obtains the number of points before it by decoupled look-back over the
// earlier tickets (they hold earlier tickets, so they are running or done: no deadlock); then it compacts its columns out of
// shared memory. Traffic = 2 B per pixel in + 16 B per kept pixel out + 4 B per column.
static constexpr unsigned long long kFlagAggregate = 1ull << 62, kFlagInclusive = 2ull << 62, kFlagMask = 3ull << 62;

__global__ void __launch_bounds__(1024) depth_fused_kernel(const unsigned short* __restrict__ depth, DepthParams P, int col_blocks,
                                                           unsigned* __restrict__ ticket, unsigned long long* __restrict__ status,
                                                           int* __restrict__ col_start, float4* __restrict__ out, int n_tickets) {
  extern __shared__ unsigned short tile[];   // [rows][34]
  __shared__ int s_ticket, s_col[33], s_base;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_ticket = (int)atomicAdd(ticket, 1u);
  __syncthreads();
  const int t = s_ticket;
  const int frame = t / col_blocks, j0 = (t % col_blocks) * 32;
  const unsigned short* d = depth + (size_t)frame * P.rows * P.cols;
  {   // row-wise load: warp w reads rows w, w + 32, ...; 16 independent 64-byte segments in flight per warp
    const int cj = j0 + lane;
    for (int r0 = w; r0 < P.rows; r0 += 32 * 16) {
      unsigned short v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) { const int i = r0 + 32 * u; v[u] = (i < P.rows && cj < P.cols) ? __ldg(d + (size_t)i * P.cols + cj) : (unsigned short)0; }
#pragma unroll
      for (int u = 0; u < 16; ++u) { const int i = r0 + 32 * u; if (i < P.rows) tile[i * 34 + lane] = v[u]; }
    }
  }
  __syncthreads();
  const int j = j0 + w;   // this warp's column
  int cnt = 0;
  for (int r0 = 0; r0 < P.rows; r0 += 32) {
    const int i = r0 + lane;
    const bool keep = (i < P.rows && j < P.cols) && depth_keep(P, tile[i * 34 + w]);
    cnt += __popc(__ballot_sync(0xffffffffu, keep));
  }
  if (lane == 0) s_col[w] = cnt;
  __syncthreads();
  if (w == 0) {
    // exclusive scan of the 32 column counts, block total, publish, look back
    const int mine = s_col[lane];
    int incl = mine;
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    long long prefix = 0;
    if (t == 0) {
      if (lane == 0) atomicExch(status + t, kFlagInclusive | (unsigned long long)total);
    } else {
      if (lane == 0) atomicExch(status + t, kFlagAggregate | (unsigned long long)total);
      int look = t - 1;
      for (;;) {   // 32 predecessors per step, newest in lane 0
        const int q = look - lane;
        unsigned long long sv = kFlagInclusive;   // before ticket 0: an inclusive prefix of 0
        if (q >= 0) { do { sv = *((volatile unsigned long long*)(status + q)); } while ((sv & kFlagMask) == 0ull); }
        const unsigned inc_mask = __ballot_sync(0xffffffffu, (sv & kFlagMask) == kFlagInclusive);
        const int first_inc = inc_mask ? __ffs(inc_mask) - 1 : 32;   // nearest predecessor that already knows its inclusive prefix
        long long contrib = lane <= first_inc ? (long long)(sv & ~kFlagMask) : 0ll;
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        prefix += contrib;
        if (inc_mask) break;
        look -= 32;
      }
      if (lane == 0) atomicExch(status + t, kFlagInclusive | (unsigned long long)(prefix + total));
    }
    s_col[lane] = (int)prefix + incl - mine;
    if (lane == 0) { s_base = (int)prefix; if (t == n_tickets - 1) col_start[(size_t)n_tickets / col_blocks * P.cols] = (int)prefix + total; }
    __threadfence();
  }
  __syncthreads();
  if (j < P.cols) {
    const int base = s_col[w];
    if (lane == 0) col_start[(size_t)frame * P.cols + j] = base;
    int run = 0;
    for (int r0 = 0; r0 < P.rows; r0 += 32) {
      const int i = r0 + lane;
      float4 pt = make_float4(0, 0, 0, 0);
      const bool keep = i < P.rows && depth_point(P, i, j, tile[i * 34 + w], pt);
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) __stcs(out + (size_t)base + run + __popc(m & ((1u << lane) - 1u)), pt);
      run += __popc(m);
    }
  }
}

int depth_to_cloud_device(ope_ctx* ctx, const unsigned short* d_depth, int frames, const DepthParams& P, float4* d_out, int* d_col_start,
                          bool counts_only_then_scan) {
  (void)counts_only_then_scan;
  const size_t ncol = (size_t)frames * P.cols;
  const size_t tile_bytes = (size_t)P.rows * 34 * sizeof(unsigned short);
  if (tile_bytes <= 200 * 1024 && !std::getenv("OPE_DEPTH_TWO_PASS")) {
    const int col_blocks = (int)div_up((size_t)P.cols, 32);
    const int n_tickets = col_blocks * frames;
    Scratch<unsigned long long> status(ctx);
    OPE_TRY(status.alloc((size_t)n_tickets + 1));
    OPE_CUDA_TRY(ctx, cudaMemsetAsync(status.p, 0, ((size_t)n_tickets + 1) * sizeof(unsigned long long), ctx->stream));
    unsigned* ticket = reinterpret_cast<unsigned*>(status.p + n_tickets);
    OPE_TRY(dyn_smem(ctx, (const void*)depth_fused_kernel, tile_bytes));
    depth_fused_kernel<<<n_tickets, 1024, tile_bytes, ctx->stream>>>(d_depth, P, col_blocks, ticket, status.p, d_col_start, d_out, n_tickets);
    return check_launch(ctx, "depth_fused_kernel");
  }
  dim3 grid(div_up((size_t)P.cols, 32), frames);
  depth_count_kernel<<<grid, 1024, 0, ctx->stream>>>(d_depth, P, d_col_start);
  OPE_TRY(check_launch(ctx, "depth_count_kernel"));
  OPE_TRY(exclusive_scan_i32(ctx, d_col_start, ncol + 1));
  depth_write_kernel<<<grid, 1024, 0, ctx->stream>>>(d_depth, P, d_col_start, d_out);
  return check_launch(ctx, "depth_write_kernel");
}

}  // namespace ope

using namespace ope;

extern "C" {

int ope_depth_to_cloud(ope_ctx* ctx, const uint16_t* depth, int rows, int cols, float fx, float fy, float cx, float cy, float scale,
                       float z_max, ope_cloud** out) {
  OPE_ENTER(ctx);
  if (!ctx || !depth || !out || rows <= 0 || cols <= 0 || !(scale > 0)) return OPE_ERR_INVALID;
  *out = nullptr;
  const size_t npx = (size_t)rows * cols;
  const DepthParams P = depth_params(rows, cols, fx, fy, cx, cy, scale, z_max);
  Scratch<unsigned short> dd(ctx);
  Scratch<int> cs(ctx);
  Scratch<float4> tmp(ctx);
  OPE_TRY(dd.alloc(npx)); OPE_TRY(cs.alloc((size_t)cols + 1)); OPE_TRY(tmp.alloc(npx));
  void* stage = nullptr;
  OPE_TRY(stage_reserve(ctx, npx * 2, &stage));
  std::memcpy(stage, depth, npx * 2);
  OPE_CUDA_TRY(ctx, cudaMemcpyAsync(dd.p, stage, npx * 2, cudaMemcpyHostToDevice, ctx->stream));
  OPE_CUDA_TRY(ctx, cudaMemsetAsync(cs.p + cols, 0, sizeof(int), ctx->stream));
  OPE_TRY(depth_to_cloud_device(ctx, dd.p, 1, P, tmp.p, cs.p, true));
  void* h;
  OPE_TRY(read_back(ctx, cs.p + cols, sizeof(int), &h));
  const size_t n = (size_t) * (const int*)h;
  ope_cloud* c = nullptr;
  OPE_TRY(cloud_alloc(ctx, n, false, &c));
  if (n) {
    cudaError_t e = cudaMemcpyAsync(c->pts, tmp.p, n * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess) e = ope::stream_sync(ctx);
    if (e != cudaSuccess) { ope_cloud_free(ctx, c); return fail(ctx, OPE_ERR_CUDA, "depth cloud copy failed: %s", cudaGetErrorString(e)); }
  }
  *out = c;
  return OPE_OK;
}

/* Batched, device-resident form (C5: a batch of frames in one launch). d_depth: frames*rows*cols uint16 on the device;
 * d_out: room for frames*rows*cols float4; d_col_start: frames*cols + 1 int32 (receives the output offset of every column;
 * the last entry is the total number of points; frame f occupies [d_col_start[f*cols], d_col_start[(f+1)*cols])). */
int ope_depth_to_cloud_batch(ope_ctx* ctx, const uint16_t* d_depth, int frames, int rows, int cols, float fx, float fy, float cx, float cy,
                             float scale, float z_max, void* d_out, int32_t* d_col_start) {
  OPE_ENTER(ctx);
  if (!ctx || !d_depth || !d_out || !d_col_start || frames <= 0 || rows <= 0 || cols <= 0 || !(scale > 0)) return OPE_ERR_INVALID;
  if ((size_t)frames * rows * cols > 0x7fffffffull) return fail(ctx, OPE_ERR_INVALID, "batch too large for 32-bit point offsets");
  const DepthParams P = depth_params(rows, cols, fx, fy, cx, cy, scale, z_max);
  OPE_CUDA_TRY(ctx, cudaMemsetAsync(d_col_start + (size_t)frames * cols, 0, sizeof(int), ctx->stream));
  return depth_to_cloud_device(ctx, d_depth, frames, P, (float4*)d_out, d_col_start, true);
}

}  // extern "C"
