#include "pc_stubs.h"  // stand-in for <stereo_msgs/DisparityImage.h>: see oracle/refbuild/stubs_pc/pc_stubs.h
