#include "pc_stubs.h"  // stand-in for <cv_bridge/cv_bridge.h>: see oracle/refbuild/stubs_pc/pc_stubs.h
