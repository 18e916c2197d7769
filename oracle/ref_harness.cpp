// ref_harness.cpp — TEST INFRASTRUCTURE. Drives the reference's OWN vendored registration sources, compiled unmodified from
// /root/reference/DetectAndLocalize/include/pcl/registration (VP) against the mock PCL in oracle/refstub (this container has no
// PCL / Eigen / FLANN / Boost), and exposes them through a small C ABI so that tests/test_ref.py can assert
// oracle == reference on the ICP configurations. Built by `make -C oracle ref` into oracle/_ref/ (git-ignored; never shipped
// with, linked into or loaded by the product). Two builds of this file exist because the two vendored ICP variants define the
// same class names: default = VP/icp_mod.h (+ impl/icp_mod.hpp), -DREF_MODCORR = VP/icp_modCorr.h (+ impl/icp_modCorr.hpp).
//
// What runs here is reference code: Registration::align / getFitnessScore (VP/impl/registration_mod.hpp:131-219),
// IterativeClosestPoint[WithNormals]::computeTransformation / transformCloud / getAlignStrength (VP/impl/icp_mod.hpp:48-318,
// VP/icp_mod.h:249-260), CorrespondenceEstimation::determine[Reciprocal]Correspondences with fixed correspondences
// (VP/impl/correspondence_estimation_mod.hpp:127-303), CorrespondenceEstimationNormalShootingWeighted::determineCorrespondences
// (VP/impl/correspondence_estimation_normal_shooting_weighted.hpp:60-196), DataContainer's scores
// (VP/correspondence_rejection_mod.h:352-391), CorrespondenceRejectorSelfOccludedNormal
// (VP/impl/correspondence_rejection_self_occluded_normal.cpp:43-64), DefaultConvergenceCriteria's state
// (VP/default_convergence_criteria_mod.h). What stands in for [UPSTREAM] PCL is listed in refstub/pcl/mock_pcl.h.
#ifdef REF_MODCORR
#include <pcl/registration/icp_modCorr.h>
#else
#include <pcl/registration/icp_mod.h>
#endif
#include <pcl/registration/correspondence_estimation_normal_shooting_weighted.h>
#include <pcl/registration/correspondence_rejection_self_occluded_normal.h>
#include <pcl/registration/correspondence_rejection_surface_normal.h>
// the one non-template translation unit of the vendored tree on this path
#include <pcl/registration/impl/correspondence_rejection_self_occluded_normal.cpp>

namespace {

using pcl::PointNormal;
using pcl::PointXYZ;
typedef pcl::PointCloud<PointXYZ> CloudXYZ;
typedef pcl::PointCloud<PointNormal> CloudPN;

CloudXYZ::Ptr makeXYZ(const float* xyz, size_t n) {
  CloudXYZ::Ptr c(new CloudXYZ);
  c->resize(n);
  for (size_t i = 0; i < n; ++i) { c->points[i].x = xyz[3 * i]; c->points[i].y = xyz[3 * i + 1]; c->points[i].z = xyz[3 * i + 2]; }
  return c;
}
CloudPN::Ptr makePN(const float* xyz, const float* nrm4, size_t n) {
  CloudPN::Ptr c(new CloudPN);
  c->resize(n);
  for (size_t i = 0; i < n; ++i) {
    PointNormal& p = c->points[i];
    p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2];
    if (nrm4) { p.normal_x = nrm4[4 * i]; p.normal_y = nrm4[4 * i + 1]; p.normal_z = nrm4[4 * i + 2]; p.curvature = nrm4[4 * i + 3]; }
  }
  return c;
}

template <typename ICP>
void configure(ICP& icp, const ope_icp_params& prm) {
  icp.setMaximumIterations(prm.max_iterations);
  icp.setTransformationEpsilon(prm.transformation_epsilon);
  icp.setEuclideanFitnessEpsilon(prm.euclidean_fitness_epsilon);
  icp.setMaxCorrespondenceDistance(prm.max_correspondence_distance);
  icp.setUseReciprocalCorrespondences(prm.use_reciprocal != 0);
  icp.getConvergeCriteria()->setAbsoluteMSE(prm.mse_threshold_absolute);
  icp.getConvergeCriteria()->setMaximumIterationsSimilarTransforms(prm.max_iterations_similar_transforms);
  icp.getConvergeCriteria()->setFailureAfterMaximumIterations(prm.failure_after_max_iterations != 0);
}

template <typename ICP, typename CloudT>
void finish(ICP& icp, CloudT& output, ope_reg_result* res, const pcl::Correspondences& corr, ope_correspondence* out_corr,
            float* out_aligned_xyz, double* fitness, double* align_strength) {
  const Eigen::Matrix4f T = icp.getFinalTransformation();
  for (int i = 0; i < 16; ++i) res->T[i] = T.d[i];
  res->converged = icp.hasConverged() ? 1 : 0;
  res->state = (int)icp.getConvergeCriteria()->getConvergenceState();
  res->n_correspondences = (int32_t)corr.size();
  res->last_mse = 0; res->best_error = 0; res->best_iteration = 0; res->reserved = 0;
  if (out_corr)
    for (size_t i = 0; i < corr.size(); ++i) out_corr[i] = ope_correspondence{corr[i].index_query, corr[i].index_match, corr[i].distance};
  if (out_aligned_xyz)
    for (size_t i = 0; i < output.points.size(); ++i) {
      out_aligned_xyz[3 * i] = output.points[i].x; out_aligned_xyz[3 * i + 1] = output.points[i].y; out_aligned_xyz[3 * i + 2] = output.points[i].z;
    }
  if (fitness) *fitness = icp.getFitnessScore();
  if (align_strength) *align_strength = icp.getAlignStrength();
}

// exposes the protected members the harness reports (iteration count, last correspondence set)
template <typename Base>
struct Open : public Base {
  int iterations() const { return this->nr_iterations_; }
  const pcl::Correspondences& corr() const { return *this->correspondences_; }
};

}  // namespace

extern "C" {

__attribute__((visibility("default"))) const char* ref_variant(void) {
#ifdef REF_MODCORR
  return "icp_modCorr";
#else
  return "icp_mod";
#endif
}

// IterativeClosestPoint / IterativeClosestPointWithNormals ::align through the vendored loop.
//   src/tgt: n*3 floats; normals: n*4 floats (nx, ny, nz, curvature) or NULL;
//   prm->estimator / rejectors / transformation / with_normals select the plug-ins exactly as the apps do
//   (D&L/src/poseestimator.cpp:242-341, BM/src/regmeshpcd.cpp:8-44,140-199); fixed: the reference's own extension
//   (setFixedCorrespondences, VP/icp_mod.h:267-281), n_fixed = 0 for none; guess may be NULL.
__attribute__((visibility("default"))) int ref_icp(const float* src, size_t ns, const float* src_normals, const float* tgt, size_t nt,
                                                   const float* tgt_normals, const ope_icp_params* prm, const float* guess,
                                                   ope_correspondence* fixed, size_t n_fixed, ope_reg_result* res,
                                                   ope_correspondence* out_corr, float* out_aligned_xyz, double* fitness,
                                                   double* align_strength) {
  Eigen::Matrix4f G = Eigen::Matrix4f::Identity();
  if (guess) for (int i = 0; i < 16; ++i) G.d[i] = guess[i];
  pcl::Correspondences fixed_list;
  for (size_t i = 0; i < n_fixed; ++i) fixed_list.push_back(pcl::Correspondence(fixed[i].index_query, fixed[i].index_match, fixed[i].distance));
  const bool normals = prm->with_normals || prm->estimator == OPE_EST_NORMAL_SHOOTING || prm->n_rejectors > 0 ||
                       prm->transformation != OPE_TE_SVD;
  if (!normals) {
    Open<pcl::IterativeClosestPoint<PointXYZ, PointXYZ> > icp;
    configure(icp, *prm);
    CloudXYZ::Ptr s = makeXYZ(src, ns), t = makeXYZ(tgt, nt);
    icp.setInputSource(s);
    icp.setInputTarget(t);
#ifndef REF_MODCORR
    if (n_fixed) icp.setFixedCorrespondences(&fixed_list);
#endif
    CloudXYZ out;
    icp.align(out, G);
    res->iterations = icp.iterations();
    finish(icp, out, res, icp.corr(), out_corr, out_aligned_xyz, fitness, align_strength);
#ifndef REF_MODCORR
    for (size_t i = 0; i < n_fixed; ++i) fixed[i].distance = fixed_list[i].distance;   // the loop rewrites the caller's list
#endif
    return 0;
  }
  if (!src_normals || !tgt_normals) return -1;
  CloudPN::Ptr s = makePN(src, src_normals, ns), t = makePN(tgt, tgt_normals, nt);
  pcl::Correspondences no_fixed;
  auto run = [&](auto& icp) {
    configure(icp, *prm);
    icp.setInputSource(s);
    icp.setInputTarget(t);
    if (prm->estimator == OPE_EST_NORMAL_SHOOTING) {
      typedef pcl::registration::CorrespondenceEstimationNormalShootingWeighted<PointNormal, PointNormal, PointNormal> NS;
      NS::Ptr est(new NS);
      est->setInputSource(s);                      // D&L/src/poseestimator.cpp:242-246
      est->setSourceNormals(CloudPN::ConstPtr(s));
      est->setInputTarget(t);
      est->setKSearch((unsigned)prm->k_search);
      est->setFixedCorrespondences(&no_fixed);     // this class dereferences the list unconditionally (:81)
      icp.setCorrespondenceEstimation(est);
    }
    for (int r = 0; r < prm->n_rejectors; ++r) {
      if (prm->rejector_kind[r] == OPE_REJ_SURFACE_NORMAL) {
        pcl::registration::CorrespondenceRejectorSurfaceNormal::Ptr rej(new pcl::registration::CorrespondenceRejectorSurfaceNormal);
        rej->setThreshold(prm->rejector_threshold[r]);
#ifdef REF_MODCORR   // the 1.7.1 loop hands the rejectors nothing: the application sets their clouds ONCE (D&L/src/poseestimator.cpp:264-273)
        rej->initializeDataContainer<PointNormal, PointNormal>();
        rej->setInputSource<PointNormal>(s);
        rej->setInputNormals<PointNormal, PointNormal>(s);
        rej->setInputTarget<PointNormal>(t);
        rej->setTargetNormals<PointNormal, PointNormal>(t);
#endif
        icp.addCorrespondenceRejector(rej);
      } else {
        pcl::registration::CorrespondenceRejectorSelfOccludedNormal::Ptr rej(new pcl::registration::CorrespondenceRejectorSelfOccludedNormal);
        rej->setThreshold(prm->rejector_threshold[r]);   // D&L/src/poseestimator.cpp:290-291
#ifdef REF_MODCORR
        rej->initializeDataContainer<PointNormal, PointNormal>();
        rej->setInputSource<PointNormal>(s);
        rej->setInputNormals<PointNormal, PointNormal>(s);
        rej->setInputTarget<PointNormal>(t);
        rej->setTargetNormals<PointNormal, PointNormal>(t);
#endif
        icp.addCorrespondenceRejector(rej);
      }
    }
    if (prm->transformation == OPE_TE_SVD)
      icp.setTransformationEstimation(pcl::registration::TransformationEstimationSVD<PointNormal, PointNormal>::Ptr(
          new pcl::registration::TransformationEstimationSVD<PointNormal, PointNormal>));          // D&L :306,341
    else if (prm->transformation == OPE_TE_POINT_TO_PLANE)
      icp.setTransformationEstimation(pcl::registration::TransformationEstimationPointToPlane<PointNormal, PointNormal>::Ptr(
          new pcl::registration::TransformationEstimationPointToPlane<PointNormal, PointNormal>));   // BM/src/regmeshpcd.cpp:162,193
    // OPE_TE_POINT_TO_PLANE_LLS is the constructor default of IterativeClosestPointWithNormals (VP/icp_mod.h:352-357)
#ifndef REF_MODCORR
    if (n_fixed && prm->estimator == OPE_EST_NEAREST) icp.setFixedCorrespondences(&fixed_list);
#endif
    CloudPN out;
    icp.align(out, G);
    res->iterations = icp.iterations();
    finish(icp, out, res, icp.corr(), out_corr, out_aligned_xyz, fitness, align_strength);
#ifndef REF_MODCORR
    if (prm->estimator == OPE_EST_NEAREST) for (size_t i = 0; i < n_fixed; ++i) fixed[i].distance = fixed_list[i].distance;
#endif
  };
  if (prm->with_normals) {
    Open<pcl::IterativeClosestPointWithNormals<PointNormal, PointNormal> > icp;
    run(icp);
  } else {
    Open<pcl::IterativeClosestPoint<PointNormal, PointNormal> > icp;
    run(icp);
  }
  return 0;
}

// CorrespondenceEstimation::determineCorrespondences / determineReciprocalCorrespondences on the clouds as given
// (VP/impl/correspondence_estimation_mod.hpp:127-303), optionally with fixed correspondences in front (:150-165).
__attribute__((visibility("default"))) int ref_correspondences(const float* src, size_t ns, const float* tgt, size_t nt,
                                                               double max_distance, int reciprocal, ope_correspondence* fixed,
                                                               size_t n_fixed, ope_correspondence* out, size_t* out_n) {
  pcl::registration::CorrespondenceEstimation<PointXYZ, PointXYZ> est;
  CloudXYZ::Ptr s = makeXYZ(src, ns), t = makeXYZ(tgt, nt);
  est.setInputSource(s);
  est.setInputTarget(t);
  pcl::Correspondences fixed_list, corr;
  for (size_t i = 0; i < n_fixed; ++i) fixed_list.push_back(pcl::Correspondence(fixed[i].index_query, fixed[i].index_match, fixed[i].distance));
  if (n_fixed) est.setFixedCorrespondences(&fixed_list);
  if (reciprocal) est.determineReciprocalCorrespondences(corr, max_distance);
  else est.determineCorrespondences(corr, max_distance);
  for (size_t i = 0; i < corr.size(); ++i) out[i] = ope_correspondence{corr[i].index_query, corr[i].index_match, corr[i].distance};
  *out_n = corr.size();
  return 0;
}

}  // extern "C"
