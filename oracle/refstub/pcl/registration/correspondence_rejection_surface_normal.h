// TEST INFRASTRUCTURE (oracle/_ref build). CorrespondenceRejectorSurfaceNormal is [UPSTREAM]; its score is the vendored
// DataContainer::getCorrespondenceScoreFromNormals (VP/correspondence_rejection_mod.h:368-376) and its structure is that of
// the vendored self-occluded rejector (VP/correspondence_rejection_self_occluded_normal.h), which was derived from it: the
// blob setters fill a DataContainer<PointXYZ, Normal>, a correspondence is kept iff score > threshold.
#ifndef OPE_REFSTUB_CORRESPONDENCE_REJECTION_SURFACE_NORMAL_H_
#define OPE_REFSTUB_CORRESPONDENCE_REJECTION_SURFACE_NORMAL_H_
#include <pcl/registration/correspondence_rejection_mod.h>
namespace pcl {
namespace registration {
class CorrespondenceRejectorSurfaceNormal : public CorrespondenceRejector {
 public:
  typedef boost::shared_ptr<CorrespondenceRejectorSurfaceNormal> Ptr;
  CorrespondenceRejectorSurfaceNormal() : threshold_(1.0) { rejection_name_ = "CorrespondenceRejectorSurfaceNormal"; }
  void getRemainingCorrespondences(const pcl::Correspondences& in, pcl::Correspondences& out) {
    if (!data_container_) return;
    unsigned int n = 0;
    out.resize(in.size());
    for (size_t i = 0; i < in.size(); ++i)
      if (data_container_->getCorrespondenceScoreFromNormals(in[i]) > threshold_) out[n++] = in[i];
    out.resize(n);
  }
  void setThreshold(double t) { threshold_ = t; }
  double getThreshold() const { return threshold_; }
  // the application-side setters (D&L/src/poseestimator.cpp:264-273)
  template <typename PointT, typename NormalT> void initializeDataContainer() { data_container_.reset(new DataContainer<PointT, NormalT>); }
  template <typename PointT> void setInputSource(const typename pcl::PointCloud<PointT>::ConstPtr& c) {
    boost::static_pointer_cast<DataContainer<PointT> >(data_container_)->setInputSource(c);
  }
  template <typename PointT> void setInputTarget(const typename pcl::PointCloud<PointT>::ConstPtr& c) {
    boost::static_pointer_cast<DataContainer<PointT> >(data_container_)->setInputTarget(c);
  }
  template <typename PointT, typename NormalT> void setInputNormals(const typename pcl::PointCloud<NormalT>::ConstPtr& n) {
    boost::static_pointer_cast<DataContainer<PointT, NormalT> >(data_container_)->setInputNormals(n);
  }
  template <typename PointT, typename NormalT> void setTargetNormals(const typename pcl::PointCloud<NormalT>::ConstPtr& n) {
    boost::static_pointer_cast<DataContainer<PointT, NormalT> >(data_container_)->setTargetNormals(n);
  }
  // the blob setters the 1.7.2 loop calls (VP/impl/icp_mod.hpp:156-163,194-204)
  bool requiresSourceNormals() const { return (true); }
  bool requiresTargetNormals() const { return (true); }
  void setSourceNormals(pcl::PCLPointCloud2::ConstPtr cloud2) {
    if (!data_container_) initializeDataContainer<PointXYZ, Normal>();
    PointCloud<Normal>::Ptr cloud(new PointCloud<Normal>);
    fromPCLPointCloud2(*cloud2, *cloud);
    setInputNormals<PointXYZ, Normal>(cloud);
  }
  void setTargetNormals(pcl::PCLPointCloud2::ConstPtr cloud2) {
    if (!data_container_) initializeDataContainer<PointXYZ, Normal>();
    PointCloud<Normal>::Ptr cloud(new PointCloud<Normal>);
    fromPCLPointCloud2(*cloud2, *cloud);
    setTargetNormals<PointXYZ, Normal>(cloud);
  }
 protected:
  void applyRejection(pcl::Correspondences& c) { getRemainingCorrespondences(*input_correspondences_, c); }
  double threshold_;
  boost::shared_ptr<DataContainerInterface> data_container_;
};
}  // namespace registration
}  // namespace pcl
#endif
