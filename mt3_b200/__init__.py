"""mt3_b200: B200-native (sm_100a) implementation of MT3's audio -> event-token hot path.

Host code is Python calling hand-written CUDA through the C ABI in include/mt3_b200.h
(ctypes); torch tensors are only the device-memory container.  There is no CPU
fallback: importing the kernels without the built library, or without a CUDA device,
raises.
"""
__version__ = "0.1.0"
