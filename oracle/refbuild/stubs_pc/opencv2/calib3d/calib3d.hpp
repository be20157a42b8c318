#include "pc_stubs.h"  // stand-in for <opencv2/calib3d/calib3d.hpp>: see oracle/refbuild/stubs_pc/pc_stubs.h
