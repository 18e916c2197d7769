/* ope_oracle.h — TEST INFRASTRUCTURE: C ABI of the CPU oracle (libope_oracle.so).
 *
 * A single-threaded CPU restatement of the reference's registration hot path (SURVEY section 8c,
 * Appendix A). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it, and only as the checker / timed baseline. The product (libope_cuda.so) never links,
 * loads or calls it.
 *
 * PARITY UNPINNED: the arithmetic of the reference lives in PCL/FLANN/Eigen, none of which exist in
 * this container, and the reference ships no golden vectors; each function below cites the reference
 * call site or vendored source it follows and SURVEY Appendix A for the [UPSTREAM] algorithm.
 *
 * Conventions: points are float triples at `stride` floats apart (stride >= 3); normals are float
 * quadruples nx,ny,nz,curvature at stride 4; matrices are 4x4 column-major (Eigen::Matrix4f).
 */
#ifndef OPE_ORACLE_H_
#define OPE_ORACLE_H_

#include "../include/ope_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- search (KdTreeFLANN semantics) ---- */
/* k-NN of each query in tgt; out_idx/out_d2 are nq*k, padded with -1 / +inf when fewer than k exist.
 * brute != 0 uses an O(N*M) scan instead of the kd-tree (they must agree). */
int orc_knn(const float* tgt, size_t nt, size_t tstride, const float* qry, size_t nq, size_t qstride,
            int k, int brute, int32_t* out_idx, float* out_d2);
/* radius search, d2 < r*r strict, ascending index. offsets is nq+1; idx/d2 hold up to `capacity`
 * entries. Returns the total count (which may exceed capacity: call again with a larger buffer). */
int64_t orc_radius(const float* tgt, size_t nt, size_t tstride, const float* qry, size_t nq, size_t qstride,
                   float radius, int64_t capacity, int64_t* offsets, int32_t* out_idx, float* out_d2);
/* feature-space k-NN, L2_Simple over `dim` floats summed left to right (K6). Brute force. */
int orc_feature_knn(const float* ftgt, size_t nt, const float* fqry, size_t nq, int dim, int k,
                    int32_t* out_idx, float* out_d2);

/* ---- down-sampling ---- */
/* pcl::UniformSampling::compute (call site D&L/src/poseestimator.cpp:141-145). Output indices in
 * ascending voxel-key order. out_idx must hold n entries. */
int orc_uniform_sample(const float* pts, size_t n, size_t stride, float leaf, int32_t* out_idx, size_t* out_n);
/* pcl::VoxelGrid::filter (call site D&L/src/processingpcd.cpp:45-59). rgb may be NULL; otherwise
 * n packed 0x00RRGGBB floats averaged per channel. Outputs hold up to n entries. */
int orc_voxel_grid(const float* pts, size_t n, size_t stride, const float* rgb, float lx, float ly, float lz,
                   float* out_xyz /* M*3 */, float* out_rgb /* M or NULL */, size_t* out_n);

/* ---- features ---- */
/* pcl::NormalEstimation::compute with setKSearch(k) (call sites D&L/src/poseestimator.cpp:151-156,
 * BM/src/regmeshpcd.cpp:74-90). out is n*4: nx, ny, nz, curvature. */
int orc_normals_knn(const float* pts, size_t n, size_t stride, int k, const float vp[3], float* out);
/* pcl::FPFHEstimation::compute with setRadiusSearch(radius) (call site D&L/src/poseestimator.cpp:121-125).
 * normals n*4; out n*33. */
int orc_fpfh(const float* pts, size_t n, size_t stride, const float* normals, float radius, float* out);
/* SPFH pass only (n*33), exposed for stage-wise parity tests. */
int orc_spfh(const float* pts, size_t n, size_t stride, const float* normals, float radius, float* out);

/* ---- rigid transforms ---- */
/* TransformationEstimationSVD::estimateRigidTransformation with index lists (NULL = identity correspondences,
 * D&L/src/poseestimator.cpp:429-435). */
int orc_umeyama(const float* src, size_t sstride, const float* tgt, size_t tstride, const int32_t* is,
                const int32_t* it, size_t n, float T[16]);
/* pcl::transformPointCloud / transformPointCloudWithNormals arithmetic. normals may be NULL. */
int orc_transform(const float* pts, size_t n, size_t stride, const float* normals, const float T[16],
                  float* out_pts /* n*3 */, float* out_normals /* n*4 or NULL */);

/* TransformationEstimationPointToPlaneLLS / TransformationEstimationPointToPlane (Levenberg-Marquardt) [UPSTREAM], the
 * estimators of IterativeClosestPointWithNormals (VP/icp_mod.h:352-357) and BuildModel (BM/src/regmeshpcd.cpp:162,193).
 * kind = OPE_TE_POINT_TO_PLANE_LLS | OPE_TE_POINT_TO_PLANE; lm_info (3 ints: status, nfev, iterations) may be NULL. */
int orc_point_to_plane(const float* src, size_t ns, size_t sstride, const float* tgt, size_t nt, size_t tstride,
                       const float* tgt_normals4, const int32_t* is, const int32_t* it, size_t n, int kind, float T[16],
                       int32_t* lm_info);

/* route of the LM linear algebra: 0 = normal equations accumulated in double (canonical), 1 = Householder QR of the full
 * m x 6 Jacobian as Eigen does (cross-check; oracle/orc_lm.h) */
void orc_lm_set_route(int householder);

/* ---- SURVEY 8f-2 groundwork: oracle only, the product has no segmentation stage yet ---- */
/* pcl::PassThrough (unorganised, lo <= field <= hi, NaN points removed), D&L/src/processingpcd.cpp:8-41. field 0/1/2 = x/y/z.
 * out_idx: room for n entries; returns the number kept (ascending original indices) or -1. */
int64_t orc_pass_through(const float* pts, size_t n, size_t stride, int field, float lo, float hi, int32_t* out_idx);
/* pcl::EuclideanClusterExtraction (D&L/src/objectsegmentationplane.cpp:74-90): labels[i] = cluster number (0 = largest; ties by
 * smaller first index) or -1 for points in no kept cluster; returns the number of clusters or -1. */
int orc_euclidean_clusters(const float* pts, size_t n, size_t stride, float tolerance, int min_size, int max_size, int32_t* labels);

/* ObjectSegmentationPlane::getSegmentedObjectsOnPlane (D&L/src/objectsegmentationplane.cpp:122-282) on an already pass-through-
 * filtered cloud without NaN points: plane RANSAC (+ refit), prism over the padded hull rectangle, second plane, clusters of the
 * rest. labels[i]: OPE_SEG_OUTSIDE_PRISM / OPE_SEG_PLANE (second plane's inlier) / OPE_SEG_NO_CLUSTER / cluster number (0 = largest).
 * plane1 / plane2: the two refined plane equations; iters: RANSAC iterations of the two fits. Returns the number of clusters,
 * -1 when no plane was found (the reference then hands the whole cloud on, :149-152,224-227). */
int orc_segment_objects_on_plane(const float* pts, size_t n, size_t stride, const ope_segment_params* prm, int32_t* labels, float plane1[4],
                                 float plane2[4], int32_t iters[2]);
void orc_segment_params_default(ope_segment_params* p);

/* ---- depth image -> cloud ---- */
/* DataGrabber::rgbd2Pcl + depthToMeter, D&L/src/datagrabber.cpp:9-62,121-174: columns outer, rows inner; Z = depth / scale;
 * the reference passes (row, col) as (x, y): y_out = (row - cx) * Z / fx, x_out = (col - cy) * Z / fy (sic); points with
 * depth == 0, Z == 0 or Z > z_max are dropped; the output is the compacted list in that traversal order.
 * out_xyz holds rows*cols*3 floats; returns the number of points. */
int64_t orc_depth_to_cloud(const uint16_t* depth, int rows, int cols, float fx, float fy, float cx, float cy, float scale, float z_max,
                           float* out_xyz);

/* ---- registration ---- */
/* Registration::getFitnessScore(max_range), VP/impl/registration_mod.hpp:131-165. */
int orc_fitness(const float* src, size_t ns, size_t sstride, const float* tgt, size_t nt, size_t tstride,
                const float T[16], double max_range, double* out);
/* One correspondence-estimation + rejector pass (VP/impl/correspondence_estimation_mod.hpp:127-213,
 * normal shooting per VP/impl/correspondence_estimation_normal_shooting_weighted.hpp:104-145, rejector chain
 * VP/impl/icp_mod.hpp:194-208) on clouds as given. out holds ns entries. */
int orc_correspondences(const float* src, size_t ns, size_t sstride, const float* src_normals,
                        const float* tgt, size_t nt, size_t tstride, const float* tgt_normals,
                        const ope_icp_params* prm, ope_correspondence* out, size_t* out_n);
/* IterativeClosestPoint[WithNormals]::align, VP/impl/registration_mod.hpp:176-219 + VP/impl/icp_mod.hpp:118-272.
 * Normals may be NULL when the configuration does not need them. out_corr (ns entries) may be NULL. */
int orc_icp(const float* src, size_t ns, size_t sstride, const float* src_normals,
            const float* tgt, size_t nt, size_t tstride, const float* tgt_normals,
            const ope_icp_params* prm, const float guess[16], ope_reg_result* res,
            ope_correspondence* out_corr);
/* The same with the reference's own extension, fixed correspondences (IterativeClosestPoint::setFixedCorrespondences,
 * VP/icp_mod.h:267-281; VP/impl/icp_mod.hpp:150-151,209-225; VP/impl/correspondence_estimation_mod.hpp:134-161): `fixed` is
 * in/out — every iteration rewrites its distances (squared distance * 1e10). out_corr must hold ns + 2*n_fixed entries.
 * prm->use_reciprocal selects determineReciprocalCorrespondences (VP/impl/correspondence_estimation_mod.hpp:216-303), which
 * ignores the fixed list. */
int orc_icp_fixed(const float* src, size_t ns, size_t sstride, const float* src_normals,
                  const float* tgt, size_t nt, size_t tstride, const float* tgt_normals,
                  const ope_icp_params* prm, const float guess[16], ope_correspondence* fixed, size_t n_fixed,
                  ope_reg_result* res, ope_correspondence* out_corr);
/* One determineCorrespondences / determineReciprocalCorrespondences pass of the nearest-neighbour estimator with a fixed list
 * (no rejectors). out holds ns + n_fixed entries. */
int orc_correspondences_fixed(const float* src, size_t ns, size_t sstride, const float* tgt, size_t nt, size_t tstride,
                              const ope_icp_params* prm, ope_correspondence* fixed, size_t n_fixed, ope_correspondence* out,
                              size_t* out_n);
/* Draw the libc rand() decisions of a SAC-IA run (selectSamples + the pick in findSimilarFeatures)
 * WITHOUT reseeding: consumes the process-wide rand() stream exactly as PCL would. samples/picks hold
 * iterations*nr_samples entries. min_sample_distance is in/out (it is halved after 3*N failed draws). */
int orc_sacia_draw(const float* src, size_t ns, size_t sstride, int iterations, int nr_samples,
                   int k_correspondences, float* min_sample_distance, int32_t* samples, int32_t* picks);
/* SampleConsensusInitialAlignment::align [UPSTREAM ia_ransac.hpp] (call site D&L/src/poseestimator.cpp:50-64).
 * table == NULL -> decisions are drawn from libc rand() on the fly (same stream as orc_sacia_draw). */
int orc_sacia(const float* src, size_t ns, size_t sstride, const float* fsrc,
              const float* tgt, size_t nt, size_t tstride, const float* ftgt,
              const ope_sacia_params* prm, const ope_rng_table* table, ope_reg_result* res,
              float* out_errors /* max_iterations floats or NULL */);
void orc_srand(unsigned seed);
/* Which Eigen build the restatement mimics for 4-float packet reductions (oracle/orc_linalg.h: 3 = SSE3+ hadd [default],
 * 2 = SSE2, 0 = scalar) and whether int->float casts vectorise (Eigen >= 3.3; default 0). Sensitivity tests only. */
void orc_set_eigen_model(int redux_level, int cast_vectorized);

/* ---- PoseEstimator (D&L/src/poseestimator.cpp:16-448) ---- */
typedef struct orc_pose_estimator orc_pose_estimator;
orc_pose_estimator* orc_pose_create(const ope_pose_params* prm);
void orc_pose_destroy(orc_pose_estimator*);
/* estimateFinalPose(p_sourceCloud, p_targetCloud, fitness, strength), :383-448. `source` is the caller's
 * in/out cloud (n*3 floats, overwritten with alignedSource like *p_sourceCloud = *alignedSource, :441).
 * The table, when given, replays the SAC-IA decisions. */
int orc_pose_estimate_final(orc_pose_estimator*, float* source, size_t ns, const float* target, size_t nt,
                            size_t tstride, const ope_rng_table* table, ope_pose_result* res);
/* per-stage wall-clock of the last estimate_final call, seconds:
 * [0] down-sample [1] normals [2] fpfh [3] sac-ia ("Initial Alignment") [4] icp ("Final Alignment")
 * [5] fitness [6] dense umeyama + transforms [7] total */
void orc_pose_stage_seconds(const orc_pose_estimator*, double out[8]);

/* default-initialisers (reference defaults, include/ope_types.h) */
void orc_icp_params_default(ope_icp_params*);
void orc_sacia_params_default(ope_sacia_params*);
void orc_pose_params_default(ope_pose_params*);

#ifdef __cplusplus
}
#endif
#endif
