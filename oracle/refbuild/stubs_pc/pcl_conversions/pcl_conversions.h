#include "pc_stubs.h"  // stand-in for <pcl_conversions/pcl_conversions.h>: see oracle/refbuild/stubs_pc/pc_stubs.h
