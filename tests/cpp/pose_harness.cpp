// tests/cpp/pose_harness.cpp — TEST PROGRAM for the drop-in boundary: application-style C++ written against PCL's own
// include paths and class names, compiled against include/ope_pcl (no PCL, Eigen or Boost installed) and linked to
// libope_cuda.so. It replays, through the PCL-style API only, the call sequences the reference applications make:
//
//   pose   the DetectAndLocalize frame path: down-sample + normals -> FPFH -> SAC-IA -> ICP-with-normals (normal
//          shooting, surface-normal + self-occluded rejectors, SVD) -> fitness / align strength -> dense SVD composition
//          (the sequence of D&L/src/poseestimator.cpp:16-448, parameters of SURVEY A.0)
//   icp    the plain point-to-point ICP of BM/src/regmeshpcd.cpp:8-44 (setMaxCorrespondenceDistance, 1e-16 epsilon)
//   chain  the BuildModel multi-view chain of BM/src/regmeshpcd.cpp:210-271 (pairwise ICP-with-normals, merge, repeat)
//
// Inputs are raw little-endian float32 xyz triples; results are printed as one JSON object so tests/test_cpp_shim.py can
// compare them with the CPU oracle running the same sequence.
#include <pcl/common/transforms.h>
#include <pcl/features/fpfh.h>
#include <pcl/features/normal_3d.h>
#include <pcl/filters/filter.h>
#include <pcl/keypoints/uniform_sampling.h>
#include <pcl/point_types.h>
#include <pcl/registration/correspondence_estimation_normal_shooting.h>
#include <pcl/registration/correspondence_rejection_self_occluded_normal.h>
#include <pcl/registration/correspondence_rejection_surface_normal.h>
#include <pcl/registration/ia_ransac.h>
#include <pcl/registration/icp.h>
#include <pcl/registration/transformation_estimation_svd.h>
#include <pcl/search/kdtree.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

typedef pcl::PointXYZRGB PointT;
typedef pcl::PointCloud<PointT> Cloud;
typedef pcl::PointXYZRGBNormal PointNT;
typedef pcl::PointCloud<PointNT> CloudN;

static Cloud::Ptr read_xyz(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  Cloud::Ptr c(new Cloud);
  if (!f) { std::fprintf(stderr, "cannot open %s\n", path.c_str()); return c; }
  float v[3];
  while (f.read(reinterpret_cast<char*>(v), sizeof(v))) {
    PointT p;
    p.x = v[0]; p.y = v[1]; p.z = v[2];
    c->push_back(p);
  }
  c->is_dense = false;
  return c;
}

static void print_mat(const char* name, const Eigen::Matrix4f& M, bool comma = true) {
  std::printf("\"%s\": [", name);
  for (int i = 0; i < 16; ++i) std::printf("%s%.9g", i ? ", " : "", M.data()[i]);  // column-major
  std::printf("]%s\n", comma ? "," : "");
}

// ---- the frame path, written the way an application would against PCL ------------------------------------------------
class FrameLocalizer {
 public:
  Cloud::Ptr aligned_, model_;
  double fine_fitness_ = 10.0, strength_ = 0.0;
  int frames_ = 0;
  int icp_iterations_ = 0, icp_state_ = 0;
  bool icp_converged_ = false;
  Eigen::Matrix4f coarse_ = Eigen::Matrix4f::Identity(), fine_ = Eigen::Matrix4f::Identity(), rigid_ = Eigen::Matrix4f::Identity();
  pcl::UniformSampling<PointT> sampler_;
  pcl::NormalEstimation<PointT, pcl::Normal> normal_est_;
  pcl::FPFHEstimation<PointT, pcl::Normal, pcl::FPFHSignature33> fpfh_est_;
  pcl::registration::TransformationEstimationSVD<PointT, PointT> svd_;

  void sampleWithNormals(const Cloud::Ptr& in, double leaf, Cloud::Ptr& sampled, pcl::PointCloud<pcl::Normal>::Ptr& normals) {
    pcl::PointCloud<int> picked;
    sampler_.setInputCloud(in);
    sampler_.setRadiusSearch(leaf);
    sampler_.compute(picked);
    sampled.reset(new Cloud);
    pcl::copyPointCloud(*in, picked.points, *sampled);
    pcl::search::KdTree<PointT>::Ptr tree(new pcl::search::KdTree<PointT>());
    normals.reset(new pcl::PointCloud<pcl::Normal>);
    normal_est_.setSearchMethod(tree);
    normal_est_.setKSearch(30);
    normal_est_.setInputCloud(sampled);
    normal_est_.compute(*normals);
  }

  void describe(const Cloud::Ptr& in, pcl::PointCloud<pcl::FPFHSignature33>::Ptr& features, Cloud::Ptr& keypoints) {
    pcl::PointCloud<pcl::Normal>::Ptr normals;
    sampleWithNormals(in, 0.01, keypoints, normals);
    features.reset(new pcl::PointCloud<pcl::FPFHSignature33>);
    fpfh_est_.setInputCloud(keypoints);
    fpfh_est_.setRadiusSearch(0.03);
    fpfh_est_.setInputNormals(normals);
    fpfh_est_.compute(*features);
  }

  Eigen::Matrix4f coarse(const Cloud::Ptr& source, const Cloud::Ptr& target) {
    pcl::PointCloud<pcl::FPFHSignature33>::Ptr fs, ft;
    Cloud::Ptr ks, kt;
    describe(source, fs, ks);
    describe(target, ft, kt);
    aligned_.reset(new Cloud);
    if (ft->size() < 10) {
      pcl::copyPointCloud(*source, *aligned_);
      return Eigen::Matrix4f::Identity();
    }
    pcl::SampleConsensusInitialAlignment<PointT, PointT, pcl::FPFHSignature33> sac;
    sac.setInputSource(ks);
    sac.setSourceFeatures(fs);
    sac.setInputTarget(kt);
    sac.setTargetFeatures(ft);
    sac.setMaximumIterations(400);
    sac.setNumberOfSamples(5);
    sac.setCorrespondenceRandomness(5);
    sac.setMaxCorrespondenceDistance(0.05);
    sac.setMinSampleDistance(0.01f);
    Cloud moved;
    sac.align(moved);
    const Eigen::Matrix4f T = sac.getFinalTransformation();
    pcl::transformPointCloud(*source, *aligned_, T);
    return T;
  }

  Eigen::Matrix4f fine(Cloud::Ptr& source, const Cloud::Ptr& target) {
    Cloud::Ptr s(new Cloud), t(new Cloud);
    std::vector<int> kept;
    pcl::removeNaNFromPointCloud(*source, *s, kept);
    pcl::removeNaNFromPointCloud(*target, *t, kept);
    Cloud::Ptr ss, ts;
    pcl::PointCloud<pcl::Normal>::Ptr ns, nt;
    sampleWithNormals(s, 0.008, ss, ns);
    sampleWithNormals(t, 0.008, ts, nt);
    CloudN::Ptr sn(new CloudN), tn(new CloudN);
    pcl::copyPointCloud(*ss, *sn);
    pcl::copyPointCloud(*ns, *sn);
    pcl::copyPointCloud(*ts, *tn);
    pcl::copyPointCloud(*nt, *tn);
    pcl::removeNaNNormalsFromPointCloud(*sn, *sn, kept);
    pcl::removeNaNNormalsFromPointCloud(*tn, *tn, kept);
    if (tn->size() < 100) return Eigen::Matrix4f::Identity();

    typedef pcl::registration::CorrespondenceEstimationNormalShooting<PointNT, PointNT, PointNT> Shooting;
    Shooting::Ptr shoot(new Shooting);
    shoot->setInputSource(sn);
    shoot->setSourceNormals(sn);
    shoot->setInputTarget(tn);
    shoot->setKSearch(20);
    pcl::registration::CorrespondenceRejectorSurfaceNormal::Ptr by_normal(new pcl::registration::CorrespondenceRejectorSurfaceNormal);
    by_normal->initializeDataContainer<PointNT, PointNT>();
    by_normal->setInputSource<PointNT>(sn);
    by_normal->setInputNormals<PointNT, PointNT>(sn);
    by_normal->setInputTarget<PointNT>(tn);
    by_normal->setTargetNormals<PointNT, PointNT>(tn);
    by_normal->setThreshold(0.7);
    pcl::registration::CorrespondenceRejectorSelfOccludedNormal::Ptr by_occlusion(new pcl::registration::CorrespondenceRejectorSelfOccludedNormal);
    by_occlusion->setThreshold(0.6);
    pcl::registration::TransformationEstimationSVD<PointNT, PointNT>::Ptr by_svd(new pcl::registration::TransformationEstimationSVD<PointNT, PointNT>);

    pcl::IterativeClosestPointWithNormals<PointNT, PointNT> icp;
    icp.setInputSource(sn);
    icp.setInputTarget(tn);
    icp.setMaximumIterations(100);
    icp.setTransformationEpsilon(1e-8);
    icp.setEuclideanFitnessEpsilon(1e-8);
    icp.setCorrespondenceEstimation(shoot);
    icp.addCorrespondenceRejector(by_normal);
#if (PCL_MINOR_VERSION >= 7 && PCL_REVISION_VERSION >= 2)
    icp.addCorrespondenceRejector(by_occlusion);
#endif
    icp.setTransformationEstimation(by_svd);
    CloudN moved;
    icp.align(moved);
    fine_fitness_ = icp.getFitnessScore();
    const Eigen::Matrix4f T = icp.getFinalTransformation();
    Cloud::Ptr full(new Cloud);
    pcl::transformPointCloud(*source, *full, T);
    *source = *full;
    strength_ = icp.getAlignStrength();
    icp_iterations_ = icp.getNumberOfIterations();
    icp_state_ = icp.getConvergenceState();
    icp_converged_ = icp.hasConverged();
    return T;
  }

  Eigen::Matrix4f localize(Cloud::Ptr& source, const Cloud::Ptr& target) {
    if (frames_++ == 0) model_ = source->makeShared();
    coarse_ = fine_ = Eigen::Matrix4f::Identity();
    if (!target->empty() && fine_fitness_ > 1e-4) coarse_ = coarse(source, target);
    if (!target->empty() && aligned_) fine_ = fine(aligned_, target);
    const Eigen::Matrix4f pose = coarse_ * fine_;
    svd_.estimateRigidTransformation(*model_, *source, rigid_);
    if (aligned_) *source = *aligned_;
    return rigid_ * pose;
  }
};

static int run_pose(int argc, char** argv) {
  if (argc < 4) return 2;
  Cloud::Ptr model = read_xyz(argv[2]);
  FrameLocalizer loc;
  std::printf("{\"frames\": [\n");
  for (int f = 3; f < argc; ++f) {
    Cloud::Ptr target = read_xyz(argv[f]);
    const Eigen::Matrix4f finalPose = loc.localize(model, target);
    std::printf("{");
    print_mat("final_pose", finalPose);
    print_mat("coarse_pose", loc.coarse_);
    print_mat("fine_pose", loc.fine_);
    print_mat("rigid_model_pose", loc.rigid_);
    std::printf("\"fitness\": %.17g, \"align_strength\": %.17g, \"icp_iterations\": %d, \"icp_state\": %d, \"icp_converged\": %d}%s\n",
                loc.fine_fitness_, loc.strength_, loc.icp_iterations_, loc.icp_state_, loc.icp_converged_ ? 1 : 0, f + 1 < argc ? "," : "");
  }
  std::printf("]}\n");
  return 0;
}

// ---- plain ICP (BM/src/regmeshpcd.cpp:8-44) ----------------------------------------------------------------------------
static int run_icp(int argc, char** argv) {
  if (argc < 6) return 2;
  Cloud::Ptr src = read_xyz(argv[2]), tgt = read_xyz(argv[3]);
  const double max_corr = std::atof(argv[4]);
  const int max_iter = std::atoi(argv[5]);
  pcl::IterativeClosestPoint<PointT, PointT> icp;
  icp.setInputSource(src);
  icp.setInputTarget(tgt);
  icp.setMaxCorrespondenceDistance(max_corr);
  icp.setMaximumIterations(max_iter);
  icp.setTransformationEpsilon(1e-16);
  icp.setRANSACOutlierRejectionThreshold(0.02);
  Cloud moved;
  icp.align(moved);
  std::printf("{");
  print_mat("T", icp.getFinalTransformation());
  double sx = 0, sy = 0, sz = 0;
  for (const PointT& p : moved) { sx += p.x; sy += p.y; sz += p.z; }
  std::printf("\"converged\": %d, \"fitness\": %.17g, \"iterations\": %d, \"state\": %d, \"n_out\": %zu, \"out_sum\": [%.9g, %.9g, %.9g]}\n",
              icp.hasConverged() ? 1 : 0, icp.getFitnessScore(), icp.getNumberOfIterations(), icp.getConvergenceState(), moved.size(), sx,
              sy, sz);
  // PCL re-reads *input_ on every align(): modify the source cloud IN PLACE (no setInputSource) and align again on the same
  // object, with a tighter distance — RegMeshPcd::getIcp2's second stage (BM/src/regmeshpcd.cpp:46-59) written the way the
  // reference's `*p_sourceCloud = *alignedCloud` pattern does it; plus reciprocal correspondences on the second stage
  if (std::getenv("OPE_ICP_SECOND_STAGE")) {
    *src = moved;
    icp.setMaxCorrespondenceDistance(0.005);
    icp.setMaximumIterations(500);
    icp.setUseReciprocalCorrespondences(std::atoi(std::getenv("OPE_ICP_SECOND_STAGE")) == 2);
    Cloud moved2;
    icp.align(moved2);
    std::printf("{");
    print_mat("T", icp.getFinalTransformation());
    std::printf("\"converged\": %d, \"fitness\": %.17g, \"iterations\": %d, \"state\": %d}\n", icp.hasConverged() ? 1 : 0,
                icp.getFitnessScore(), icp.getNumberOfIterations(), icp.getConvergenceState());
  }
  // error behaviour: align() without a target prints an error and leaves the transformation at identity
  pcl::IterativeClosestPoint<PointT, PointT> empty;
  empty.setInputSource(src);
  Cloud none;
  empty.align(none);
  if (!empty.getFinalTransformation().isIdentity() || empty.hasConverged()) { std::fprintf(stderr, "missing-target behaviour broken\n"); return 1; }
  return 0;
}

// ---- multi-view chain (BM/src/regmeshpcd.cpp:210-271): getIcpNormal per pair with the estimator the reference plugs in,
// TransformationEstimationPointToPlane (Levenberg-Marquardt, :162,193); OPE_CHAIN_TE=svd|lls selects another one -----
static CloudN::Ptr with_normals(const Cloud::Ptr& c, pcl::search::KdTree<PointT>::Ptr tree) {
  pcl::NormalEstimation<PointT, PointNT> ne;
  ne.setSearchMethod(tree);
  ne.setKSearch(12);
  ne.setInputCloud(c);
  CloudN::Ptr out(new CloudN);
  ne.compute(*out);
  pcl::copyPointCloud(*c, *out);
  return out;
}

static int run_chain(int argc, char** argv) {
  if (argc < 6) return 2;
  const double reject = std::atof(argv[2]);
  const int max_iter = std::atoi(argv[3]);
  Cloud::Ptr merged = read_xyz(argv[4]);
  std::printf("{\"pairs\": [\n");
  for (int v = 5; v < argc; ++v) {
    Cloud::Ptr target = read_xyz(argv[v]);
    pcl::search::KdTree<PointT>::Ptr tree(new pcl::search::KdTree<PointT>());
    CloudN::Ptr sn = with_normals(merged, tree), tn = with_normals(target, tree);
    typedef pcl::registration::CorrespondenceEstimationNormalShooting<PointNT, PointNT, PointNT> Shooting;
    Shooting::Ptr shoot(new Shooting);
    shoot->setInputSource(sn);
    shoot->setSourceNormals(sn);
    shoot->setInputTarget(tn);
    shoot->setKSearch(20);
    pcl::registration::CorrespondenceRejectorSurfaceNormal::Ptr by_normal(new pcl::registration::CorrespondenceRejectorSurfaceNormal);
    by_normal->setThreshold(reject);
    pcl::IterativeClosestPointWithNormals<PointNT, PointNT> icp;
    icp.setInputSource(sn);
    icp.setInputTarget(tn);
    icp.setMaximumIterations(max_iter);
    icp.setTransformationEpsilon(1e-8);
    icp.setEuclideanFitnessEpsilon(1e-8);
    icp.setCorrespondenceEstimation(shoot);
    icp.addCorrespondenceRejector(by_normal);
    const char* te = std::getenv("OPE_CHAIN_TE");
    if (te && std::string(te) == "svd")
      icp.setTransformationEstimation(pcl::registration::TransformationEstimationSVD<PointNT, PointNT>::Ptr(
          new pcl::registration::TransformationEstimationSVD<PointNT, PointNT>));
    else if (te && std::string(te) == "lls")
      icp.setTransformationEstimation(pcl::registration::TransformationEstimationPointToPlaneLLS<PointNT, PointNT>::Ptr(
          new pcl::registration::TransformationEstimationPointToPlaneLLS<PointNT, PointNT>));
    else
      icp.setTransformationEstimation(pcl::registration::TransformationEstimationPointToPlane<PointNT, PointNT>::Ptr(
          new pcl::registration::TransformationEstimationPointToPlane<PointNT, PointNT>));
    CloudN moved;
    icp.align(moved);
    const Eigen::Matrix4f T = icp.getFinalTransformation();
    Cloud::Ptr aligned(new Cloud);
    pcl::transformPointCloud(*merged, *aligned, T);
    *aligned += *target;
    *merged = *aligned;
    std::printf("{");
    print_mat("T", T);
    std::printf("\"fitness\": %.17g, \"iterations\": %d, \"converged\": %d, \"merged\": %zu}%s\n", icp.getFitnessScore(),
                icp.getNumberOfIterations(), icp.hasConverged() ? 1 : 0, merged->size(), v + 1 < argc ? "," : "");
  }
  std::printf("]}\n");
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: pose_harness pose|icp|chain ...\n"); return 2; }
  const std::string mode = argv[1];
  if (mode == "pose") return run_pose(argc, argv);
  if (mode == "icp") return run_icp(argc, argv);
  if (mode == "chain") return run_chain(argc, argv);
  return 2;
}
