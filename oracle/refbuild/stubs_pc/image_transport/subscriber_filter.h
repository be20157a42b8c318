#include "pc_stubs.h"  // stand-in for <image_transport/subscriber_filter.h>: see oracle/refbuild/stubs_pc/pc_stubs.h
