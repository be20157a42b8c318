#include "pc_stubs.h"  // stand-in for <sensor_msgs/JointState.h>: see oracle/refbuild/stubs_pc/pc_stubs.h
