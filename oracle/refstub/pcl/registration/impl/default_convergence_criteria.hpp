// TEST INFRASTRUCTURE (oracle/_ref build). DefaultConvergenceCriteria<Scalar>::hasConverged is [UPSTREAM]
// (registration/impl/default_convergence_criteria.hpp, PCL 1.7.2) — the vendored header
// VP/default_convergence_criteria_mod.h:64-282 declares it and includes this path. Restated here: iteration cap, then the
// transform test (cos of the rotation angle and squared translation of the LAST INCREMENT), then absolute and relative change
// of the mean correspondence distance, each guarded by the similar-transforms counter.
#ifndef OPE_REFSTUB_DEFAULT_CONVERGENCE_CRITERIA_HPP_
#define OPE_REFSTUB_DEFAULT_CONVERGENCE_CRITERIA_HPP_
template <typename Scalar> bool pcl::registration::DefaultConvergenceCriteria<Scalar>::hasConverged() {
  convergence_state_ = CONVERGENCE_CRITERIA_NOT_CONVERGED;
  if (iterations_ >= max_iterations_) {
    if (failure_after_max_iter_) return (false);
    convergence_state_ = CONVERGENCE_CRITERIA_ITERATIONS;
    return (true);
  }
  double cos_angle = 0.5 * (transformation_(0, 0) + transformation_(1, 1) + transformation_(2, 2) - 1);
  double translation_sqr = transformation_(0, 3) * transformation_(0, 3) + transformation_(1, 3) * transformation_(1, 3) +
                           transformation_(2, 3) * transformation_(2, 3);
  if (cos_angle >= rotation_threshold_ && translation_sqr <= translation_threshold_) {
    if (iterations_similar_transforms_ < max_iterations_similar_transforms_) { ++iterations_similar_transforms_; return (false); }
    iterations_similar_transforms_ = 0;
    convergence_state_ = CONVERGENCE_CRITERIA_TRANSFORM;
    return (true);
  }
  correspondences_cur_mse_ = calculateMSE(correspondences_);
  if (fabs(correspondences_cur_mse_ - correspondences_prev_mse_) < mse_threshold_absolute_) {
    if (iterations_similar_transforms_ < max_iterations_similar_transforms_) { ++iterations_similar_transforms_; return (false); }
    iterations_similar_transforms_ = 0;
    convergence_state_ = CONVERGENCE_CRITERIA_ABS_MSE;
    return (true);
  }
  if (fabs(correspondences_cur_mse_ - correspondences_prev_mse_) / correspondences_prev_mse_ < mse_threshold_relative_) {
    if (iterations_similar_transforms_ < max_iterations_similar_transforms_) { ++iterations_similar_transforms_; return (false); }
    iterations_similar_transforms_ = 0;
    convergence_state_ = CONVERGENCE_CRITERIA_REL_MSE;
    return (true);
  }
  correspondences_prev_mse_ = correspondences_cur_mse_;
  return (false);
}
#endif
