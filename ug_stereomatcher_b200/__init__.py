"""ug_stereomatcher_b200 -- B200-native (sm_100a) implementation of the ug_stereomatcher gpu_matcher
hot path behind the reference's MatchGPULib / GetDisparitiesGPU contract.

  csrc/                hand-written CUDA kernels, the level driver and the C-ABI (include/ugsm.h),
                       the reference-named C++ facade (include/MatchGPULib.h), the headless driver
  capi.py              ctypes binding of the C-ABI (used by tests and bench.py)
  matcher.py           Python mirror of the MatchGPULib interface over the C-ABI
  synth.py             synthetic stereo pairs (BASELINE.json configs)

There is no CPU fallback: everything raises if libugsm.so or an sm_100 device is missing.
"""
from . import capi, synth  # noqa: F401
from .matcher import MatchGPULib  # noqa: F401
