#include "pc_stubs.h"  // stand-in for <pcl/io/pcd_io.h>: see oracle/refbuild/stubs_pc/pc_stubs.h
