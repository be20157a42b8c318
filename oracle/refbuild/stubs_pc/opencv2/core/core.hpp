#include "pc_stubs.h"  // stand-in for <opencv2/core/core.hpp>: see oracle/refbuild/stubs_pc/pc_stubs.h
