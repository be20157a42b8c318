#include "pc_stubs.h"  // stand-in for <image_transport/image_transport.h>: see oracle/refbuild/stubs_pc/pc_stubs.h
