// <pcl/common/centroid.h> forwarded to the B200 shim (include/ope_pcl/filters.h); see INTEGRATION.md.
#pragma once
#include "../pcl_config.h"
#include "../../../ope_pcl/filters.h"
