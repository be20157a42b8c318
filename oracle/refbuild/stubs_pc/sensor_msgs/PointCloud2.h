#include "pc_stubs.h"  // stand-in for <sensor_msgs/PointCloud2.h>: see oracle/refbuild/stubs_pc/pc_stubs.h
