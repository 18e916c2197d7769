// TEST INFRASTRUCTURE (oracle/_ref build, icp_modCorr variant only). VP/icp_modCorr.h targets PCL 1.7.1 and includes the INSTALLED
// registration.h / correspondence_estimation.h / default_convergence_criteria.h / impl/icp.hpp; none exist here, so these paths
// forward to the vendored copies of the same classes (which define the very include guards of the files they stand in for).
#include <pcl/registration/default_convergence_criteria_mod.h>
