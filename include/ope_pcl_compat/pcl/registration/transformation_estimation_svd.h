// <pcl/registration/transformation_estimation_svd.h> forwarded to the B200 shim (include/ope_pcl/registration.h); see INTEGRATION.md.
#pragma once
#include "../pcl_config.h"
#include "../../../ope_pcl/registration.h"
