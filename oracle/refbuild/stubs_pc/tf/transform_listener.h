#include "pc_stubs.h"  // stand-in for <tf/transform_listener.h>: see oracle/refbuild/stubs_pc/pc_stubs.h
