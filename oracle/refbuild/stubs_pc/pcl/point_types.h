#include "pc_stubs.h"  // stand-in for <pcl/point_types.h>: see oracle/refbuild/stubs_pc/pc_stubs.h
