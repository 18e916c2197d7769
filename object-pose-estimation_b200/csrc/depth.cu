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
#include "ope_host.cuh"

namespace ope {

struct DepthParams {
  int rows, cols;
  float fx, fy, cx, cy, scale, z_max;
};

__device__ __forceinline__ bool depth_point(const DepthParams& P, int i, int j, unsigned short raw, float4& out) {
  const float r = (float)raw;
  if (r <= 0.0f) return false;
  const float Z = r / P.scale;
  if (Z == 0.0f || Z > P.z_max) return false;
  const float X = ((float)i - P.cx) * Z / P.fx;  // p_FeatX = row (sic)
  const float Y = ((float)j - P.cy) * Z / P.fy;  // p_FeatY = column (sic)
  out = make_float4(Y, X, Z, 1.0f);
  return true;
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
      for (int t = 0; t < 16; ++t) { float4 pt; cnt += depth_point(P, r0 + 32 * t, j, v[t], pt) ? 1 : 0; }
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

int depth_to_cloud_device(ope_ctx* ctx, const unsigned short* d_depth, int frames, const DepthParams& P, float4* d_out, int* d_col_start,
                          bool counts_only_then_scan) {
  (void)counts_only_then_scan;
  const size_t ncol = (size_t)frames * P.cols;
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
  if (!ctx || !depth || !out || rows <= 0 || cols <= 0 || !(scale > 0)) return OPE_ERR_INVALID;
  *out = nullptr;
  const size_t npx = (size_t)rows * cols;
  DepthParams P{rows, cols, fx, fy, cx, cy, scale, z_max};
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
  if (!ctx || !d_depth || !d_out || !d_col_start || frames <= 0 || rows <= 0 || cols <= 0 || !(scale > 0)) return OPE_ERR_INVALID;
  if ((size_t)frames * rows * cols > 0x7fffffffull) return fail(ctx, OPE_ERR_INVALID, "batch too large for 32-bit point offsets");
  DepthParams P{rows, cols, fx, fy, cx, cy, scale, z_max};
  OPE_CUDA_TRY(ctx, cudaMemsetAsync(d_col_start + (size_t)frames * cols, 0, sizeof(int), ctx->stream));
  return depth_to_cloud_device(ctx, d_depth, frames, P, (float4*)d_out, d_col_start, true);
}

}  // extern "C"
