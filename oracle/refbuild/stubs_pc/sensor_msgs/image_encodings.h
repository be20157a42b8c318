#include "pc_stubs.h"  // stand-in for <sensor_msgs/image_encodings.h>: see oracle/refbuild/stubs_pc/pc_stubs.h
