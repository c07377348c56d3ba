"""dmcf_b200 -- B200-native (sm_100a) implementation of DMCF's per-step particle hot path.

Fixed-radius neighbour search + ContinuousConv / AntiSymmetricContinuousConv stack behind the reference's
layer / model API (utils/convolutions.py, models/*.py, pipelines/simulator.py of tum-pbs/DMCF), calling
hand-written CUDA kernels through the C ABI in include/dmcf_b200.h.  No CPU fallback.
"""
__version__ = "0.1.0"
