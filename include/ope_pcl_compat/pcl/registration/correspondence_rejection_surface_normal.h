// <pcl/registration/correspondence_rejection_surface_normal.h> forwarded to the B200 shim (include/ope_pcl/registration.h); see INTEGRATION.md.
#pragma once
#include "../pcl_config.h"
#include "../../../ope_pcl/registration.h"
