"""brutus_b200 -- B200-native brute-force photometric likelihood sweep (joshspeagle/brutus hot path).

Host side is Python + ctypes over ``libbrutus_b200.so`` (hand-written sm_100a CUDA, C ABI declared in
``include/brutus_b200.h``).  There is no CPU fallback: importing :mod:`brutus_b200.fitting` works
without a GPU, but every compute call raises if the CUDA library or a device is missing.
"""
__version__ = "0.1.0"
