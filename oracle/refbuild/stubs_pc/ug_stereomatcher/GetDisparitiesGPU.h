#include "pc_stubs.h"  // stand-in for <ug_stereomatcher/GetDisparitiesGPU.h>: see oracle/refbuild/stubs_pc/pc_stubs.h
