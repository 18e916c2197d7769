// TEST INFRASTRUCTURE (oracle/_ref build). The reference vendors its own, extended copy of this header
// (VP/correspondence_rejection_mod.h: DataContainer gains getCorrespondenceScoreSelfOccludedNormal, :382-391) under the SAME
// include guard, so in the reference's build whichever is included first wins; here the vendored one always does.
#include <pcl/mock_pcl.h>
#include <pcl/registration/correspondence_rejection_mod.h>
