#include "pc_stubs.h"  // stand-in for <opencv2/features2d/features2d.hpp>: see oracle/refbuild/stubs_pc/pc_stubs.h
