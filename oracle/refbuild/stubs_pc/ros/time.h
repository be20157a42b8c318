#include "pc_stubs.h"  // stand-in for <ros/time.h>: see oracle/refbuild/stubs_pc/pc_stubs.h
