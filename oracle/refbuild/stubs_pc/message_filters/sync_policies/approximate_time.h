#include "pc_stubs.h"  // stand-in for <message_filters/sync_policies/approximate_time.h>: see oracle/refbuild/stubs_pc/pc_stubs.h
