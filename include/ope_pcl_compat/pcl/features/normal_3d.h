// <pcl/features/normal_3d.h> forwarded to the B200 shim (include/ope_pcl/features.h); see INTEGRATION.md.
#pragma once
#include "../pcl_config.h"
#include "../../../ope_pcl/features.h"
