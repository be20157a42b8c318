#include "pc_stubs.h"  // stand-in for <ros/ros.h>: see oracle/refbuild/stubs_pc/pc_stubs.h
