// TEST INFRASTRUCTURE (oracle/_ref build): forwards an [UPSTREAM] PCL include path to the mock in pcl/mock_pcl.h
#include <pcl/mock_pcl.h>
