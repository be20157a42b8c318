#include "pc_stubs.h"  // stand-in for <ug_stereomatcher/foveatedstack.h>: see oracle/refbuild/stubs_pc/pc_stubs.h
