// <pcl/pcl_config.h> stand-in: puts the shim classes in namespace `pcl` and reports the PCL release whose behaviour they
// reproduce (1.7.2: the branch of D&L/include/poseestimator.h:8-14 and D&L/CMakeLists.txt:36-41 that uses icp_mod.h and
// the self-occluded-normal rejector).
#pragma once
#ifndef OPE_PCL_NAMESPACE
#define OPE_PCL_NAMESPACE pcl
#endif
#define PCL_MAJOR_VERSION 1
#define PCL_MINOR_VERSION 7
#define PCL_REVISION_VERSION 2
#define PCL_VERSION_PRETTY "1.7.2 (ope_pcl shim over libope_cuda)"
#define PCL_VERSION_CALC(MAJ, MIN, PATCH) (MAJ * 100000 + MIN * 100 + PATCH)
#define PCL_VERSION PCL_VERSION_CALC(PCL_MAJOR_VERSION, PCL_MINOR_VERSION, PCL_REVISION_VERSION)
#define PCL_VERSION_COMPARE(OP, MAJ, MIN, PATCH) (PCL_VERSION OP PCL_VERSION_CALC(MAJ, MIN, PATCH))
