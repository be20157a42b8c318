#include "pc_stubs.h"  // stand-in for <message_filters/subscriber.h>: see oracle/refbuild/stubs_pc/pc_stubs.h
