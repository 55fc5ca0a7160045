"""B200-native PTAM tracking front-end (pyramid + FAST-10, PatchFinder ZMSSD search, Tukey-WLS pose update).

The product is the CUDA library `libvslam_b200.so` behind the C-ABI of include/vslam_b200.h; `api` is a thin
ctypes binding used by the tests and the benchmark, `synth` produces synthetic frames / maps.
"""
__all__ = ["api", "synth"]
