// pipeline.cu — PoseEstimator::estimateFinalPose (D&L/src/poseestimator.cpp:383-448) sequenced on the device:
// UniformSampling -> NormalEstimation -> FPFH -> SAC-IA -> ICP-with-normals -> fitness -> dense Umeyama.
// Every stage is a call into the kernels of grid.cu / features.cu / registration.cu; only sizes, the 4x4s and the
// scores cross back to the host between stages.
#include <algorithm>
#include <cmath>

#include "ope_host.cuh"

using namespace ope;

struct ope_pose_tracker {
  ope_ctx* ctx = nullptr;
  ope_pose_params prm;
  int firstTimePose = 0;           // D&L/include/poseestimator.h:50-53
  double fitnessScoreFine = 10;
  double alignedStrength = 0.0;
  ope_cloud* alignedSource = nullptr;
  ope_cloud* cloudModel = nullptr;
  double stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evt0 = nullptr, evt1 = nullptr;
};

namespace {

int clone_cloud(ope_ctx* ctx, const ope_cloud* in, ope_cloud** out) {
  ope_cloud* o = nullptr;
  OPE_TRY(cloud_alloc(ctx, in->n, in->normals != nullptr, &o));
  cudaError_t e = cudaSuccess;
  if (in->n) {
    e = cudaMemcpyAsync(o->pts, in->pts, in->n * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess && in->normals)
      e = cudaMemcpyAsync(o->normals, in->normals, in->n * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream);
  }
  if (e != cudaSuccess) { ope_cloud_free(ctx, o); return fail(ctx, OPE_ERR_CUDA, "device copy failed: %s", cudaGetErrorString(e)); }
  *out = o;
  return OPE_OK;
}

struct StageTimer {
  ope_pose_tracker* t;
  explicit StageTimer(ope_pose_tracker* tr) : t(tr) { cudaEventRecord(t->ev0, t->ctx->stream); }
  // closes the current interval into slot `s` and starts the next one
  void lap(int s) {
    cudaEventRecord(t->ev1, t->ctx->stream);
    cudaEventSynchronize(t->ev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, t->ev0, t->ev1);
    t->stage_ms[s] += ms;
    std::swap(t->ev0, t->ev1);
  }
};

// subSampleAndCalculateNormals, D&L/src/poseestimator.cpp:131-158
int sub_sample_and_normals(ope_pose_tracker* t, StageTimer& tm, ope_cloud* in, float leaf, ope_cloud** out) {
  ope_ctx* ctx = t->ctx;
  int* d_idx = nullptr;
  size_t m = 0;
  OPE_TRY(uniform_sample_device(ctx, in, leaf, &d_idx, &m));
  ope_cloud* ds = nullptr;
  int rc = gather_cloud(ctx, in, d_idx, m, &ds);
  dfree(ctx, d_idx);
  if (rc != OPE_OK) return rc;
  dfree(ctx, ds->normals);  // the sub-sampled PointT cloud carries no normals
  ds->normals = nullptr;
  tm.lap(0);
  const float vp[3] = {0, 0, 0};
  rc = normals_device(ctx, ds, t->prm.normal_k, vp);
  if (rc != OPE_OK) { ope_cloud_free(ctx, ds); return rc; }
  tm.lap(1);
  *out = ds;
  return OPE_OK;
}

struct CloudGuard {
  ope_ctx* ctx;
  ope_cloud* c = nullptr;
  explicit CloudGuard(ope_ctx* x) : ctx(x) {}
  ~CloudGuard() { if (c) ope_cloud_free(ctx, c); }
};
struct DevGuard {
  ope_ctx* ctx;
  float* p = nullptr;
  explicit DevGuard(ope_ctx* x) : ctx(x) {}
  ~DevGuard() { dfree(ctx, p); }
};

}  // namespace

extern "C" {

int ope_pose_tracker_create(ope_ctx* ctx, const ope_pose_params* prm, ope_pose_tracker** out) {
  if (!ctx || !out) return OPE_ERR_INVALID;
  ope_pose_tracker* t = new ope_pose_tracker();
  t->ctx = ctx;
  if (prm) t->prm = *prm; else ope_pose_params_default(&t->prm);
  if (cudaEventCreate(&t->ev0) != cudaSuccess || cudaEventCreate(&t->ev1) != cudaSuccess ||
      cudaEventCreate(&t->evt0) != cudaSuccess || cudaEventCreate(&t->evt1) != cudaSuccess) {
    delete t;
    return fail(ctx, OPE_ERR_CUDA, "event creation failed");
  }
  *out = t;
  return OPE_OK;
}

void ope_pose_tracker_destroy(ope_pose_tracker* t) {
  if (!t) return;
  if (t->alignedSource) ope_cloud_free(t->ctx, t->alignedSource);
  if (t->cloudModel) ope_cloud_free(t->ctx, t->cloudModel);
  cudaEventDestroy(t->ev0); cudaEventDestroy(t->ev1); cudaEventDestroy(t->evt0); cudaEventDestroy(t->evt1);
  delete t;
}

int ope_pose_stage_ms(const ope_pose_tracker* t, double out[8]) {
  if (!t || !out) return OPE_ERR_INVALID;
  std::memcpy(out, t->stage_ms, sizeof(t->stage_ms));
  return OPE_OK;
}

int ope_pose_estimate_final_device(ope_pose_tracker* t, ope_cloud** source, const ope_cloud* target,
                                   const ope_rng_table* table, ope_pose_result* res) {
  if (!t || !source || !*source || !res) return OPE_ERR_INVALID;
  ope_ctx* ctx = t->ctx;
  const ope_pose_params& P = t->prm;
  std::memset(res, 0, sizeof(*res));
  std::memset(t->stage_ms, 0, sizeof(t->stage_ms));
  cudaEventRecord(t->evt0, ctx->stream);
  StageTimer tm(t);
  ope_cloud* p_source = *source;
  const size_t nt = target ? target->n : 0;
  if (t->firstTimePose == 0) {  // :385-388
    if (t->cloudModel) { ope_cloud_free(ctx, t->cloudModel); t->cloudModel = nullptr; }
    OPE_TRY(clone_cloud(ctx, p_source, &t->cloudModel));
  }
  t->firstTimePose++;
  Mat4 coarse = mat4_identity(), fine = mat4_identity();

  // ---- COARSE: estimateCoarsePose, :16-73 ----
  if (nt != 0 && t->fitnessScoreFine > P.coarse_refit_threshold) {
    res->ran_coarse = 1;
    CloudGuard sp(ctx), tp(ctx);
    OPE_TRY(sub_sample_and_normals(t, tm, p_source, P.coarse_leaf, &sp.c));
    OPE_TRY(sub_sample_and_normals(t, tm, const_cast<ope_cloud*>(target), P.coarse_leaf, &tp.c));
    DevGuard fs(ctx), ft(ctx);
    OPE_TRY(fpfh_device(ctx, sp.c, P.fpfh_radius, &fs.p, nullptr));
    OPE_TRY(fpfh_device(ctx, tp.c, P.fpfh_radius, &ft.p, nullptr));
    tm.lap(2);
    res->n_src_coarse = (int32_t)sp.c->n; res->n_tgt_coarse = (int32_t)tp.c->n;
    if (t->alignedSource) { ope_cloud_free(ctx, t->alignedSource); t->alignedSource = nullptr; }
    if ((int)tp.c->n < P.min_target_features) {
      OPE_TRY(clone_cloud(ctx, p_source, &t->alignedSource));  // :41
    } else {
      ope_reg_result rr;
      OPE_TRY(sacia_device(ctx, sp.c, fs.p, tp.c, ft.p, P.sacia, table, nullptr, &rr, nullptr));
      tm.lap(3);
      std::memcpy(coarse.m, rr.T, sizeof(coarse.m));
      res->sacia_best_iteration = rr.best_iteration; res->sacia_best_error = rr.best_error;
      OPE_TRY(cloud_alloc(ctx, p_source->n, false, &t->alignedSource));
      ope_cloud view = *p_source;  // transform points only
      view.normals = nullptr; view.grids.clear();
      OPE_TRY(transform_device(ctx, &view, coarse, t->alignedSource));  // :67-70
      tm.lap(6);
    }
  }
  // ---- FINE: estimateFinePose, :161-379 ----
  if (nt != 0 && t->alignedSource) {
    CloudGuard sp(ctx), tp(ctx);
    OPE_TRY(sub_sample_and_normals(t, tm, t->alignedSource, P.fine_leaf, &sp.c));
    OPE_TRY(sub_sample_and_normals(t, tm, const_cast<ope_cloud*>(target), P.fine_leaf, &tp.c));
    OPE_TRY(remove_nan_normals_device(ctx, &sp.c));  // :215-216
    OPE_TRY(remove_nan_normals_device(ctx, &tp.c));
    tm.lap(1);
    res->n_src_fine = (int32_t)sp.c->n; res->n_tgt_fine = (int32_t)tp.c->n;
    if ((int)tp.c->n >= P.min_target_points) {  // :218-223
      ope_reg_result rr;
      OPE_TRY(icp_device(ctx, sp.c, tp.c, P.icp, mat4_identity(), &rr, nullptr, nullptr));
      tm.lap(4);
      std::memcpy(fine.m, rr.T, sizeof(fine.m));
      OPE_TRY(fitness_device(ctx, sp.c, tp.c, fine, DBL_MAX, &t->fitnessScoreFine));  // :354
      tm.lap(5);
      ope_cloud* moved = nullptr;
      OPE_TRY(cloud_alloc(ctx, t->alignedSource->n, false, &moved));
      int rc = transform_device(ctx, t->alignedSource, fine, moved);  // :358-360
      if (rc != OPE_OK) { ope_cloud_free(ctx, moved); return rc; }
      ope_cloud_free(ctx, t->alignedSource);
      t->alignedSource = moved;
      t->alignedStrength = (double)rr.n_correspondences / (double)((long)sp.c->n + (long)tp.c->n);  // VP/icp_mod.h:249-260
      res->icp_iterations = rr.iterations; res->icp_converged = rr.converged; res->icp_state = rr.state;
      tm.lap(6);
    }
  }
  const Mat4 pose = mat4_mul(coarse, fine);  // :421 (sic: coarse * fine)
  Mat4 rigid = mat4_identity();
  if (t->cloudModel && t->cloudModel->n > 0 && t->cloudModel->n <= p_source->n) {  // :425-436
    OPE_TRY(umeyama_device(ctx, t->cloudModel->pts, p_source->pts, nullptr, nullptr, t->cloudModel->n, rigid.m));
  }
  const Mat4 final_pose = mat4_mul(rigid, pose);  // :439
  if (t->alignedSource) {  // *p_sourceCloud = *alignedSource, :441
    ope_cloud* copy = nullptr;
    OPE_TRY(clone_cloud(ctx, t->alignedSource, &copy));
    ope_cloud_free(ctx, p_source);
    *source = copy;
  }
  tm.lap(6);
  std::memcpy(res->final_pose, final_pose.m, 64); std::memcpy(res->coarse_pose, coarse.m, 64);
  std::memcpy(res->fine_pose, fine.m, 64); std::memcpy(res->rigid_model_pose, rigid.m, 64);
  res->fitness = t->fitnessScoreFine; res->align_strength = t->alignedStrength;
  cudaEventRecord(t->evt1, ctx->stream);
  cudaEventSynchronize(t->evt1);
  float ms = 0;
  cudaEventElapsedTime(&ms, t->evt0, t->evt1);
  t->stage_ms[7] = ms;
  return OPE_OK;
}

int ope_pose_estimate_final(ope_pose_tracker* t, float* source_xyz, size_t ns, const void* target, size_t nt, size_t tstride,
                            size_t toffset, const ope_rng_table* table, ope_pose_result* res) {
  if (!t || !source_xyz || !res) return OPE_ERR_INVALID;
  ope_ctx* ctx = t->ctx;
  ope_cloud* src = nullptr;
  OPE_TRY(ope_cloud_upload(ctx, source_xyz, ns, 12, 0, nullptr, 0, 0, &src));
  CloudGuard tg(ctx);
  if (nt > 0) {
    int rc = ope_cloud_upload(ctx, target, nt, tstride, toffset, nullptr, 0, 0, &tg.c);
    if (rc != OPE_OK) { ope_cloud_free(ctx, src); return rc; }
  }
  int rc = ope_pose_estimate_final_device(t, &src, tg.c, table, res);
  if (rc == OPE_OK && src->n == ns) rc = ope_cloud_download(ctx, src, source_xyz, nullptr);
  ope_cloud_free(ctx, src);
  return rc;
}

}  // extern "C"
