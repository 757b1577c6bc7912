"""botlab_b200 -- B200-native Monte Carlo localization engine behind botLab's ParticleFilter API.

csrc/      CUDA kernels + the C ABI (include/mcl_cuda.h) -> libmcl_cuda.so
src/slam/  C++ host classes with the reference's public signatures, calling the C ABI
engine.py  ctypes plumbing for tests and bench.py;  synth.py  seeded synthetic workloads
"""
__all__ = ["engine", "synth"]
