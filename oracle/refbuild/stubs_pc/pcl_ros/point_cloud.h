#include "pc_stubs.h"  // stand-in for <pcl_ros/point_cloud.h>: see oracle/refbuild/stubs_pc/pc_stubs.h
