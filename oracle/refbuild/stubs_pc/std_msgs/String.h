#include "pc_stubs.h"  // stand-in for <std_msgs/String.h>: see oracle/refbuild/stubs_pc/pc_stubs.h
