"""acg_b200: B200-native training hot path of action_conditioned_GANs.

Host code is Python; every tensor operation of the training step is a hand-written sm_100a CUDA kernel behind the
C-ABI of include/acg_b200.h (libacg_b200.so, built in-tree by `_build.build()`).  No CPU fallback.
"""
__version__ = "0.1.0"
