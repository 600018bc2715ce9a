"""srrg2_slam_interfaces_b200 -- B200-native hot paths of srrg2_slam_interfaces behind its plugin API.

capi      ctypes binding of the C ABI (include/srrg2b.h, libsrrg2b.so; CUDA sm_100a, no CPU fallback)
synthetic seeded input generators for the BASELINE.json configurations
"""
from . import capi, synthetic  # noqa: F401
