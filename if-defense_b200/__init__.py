"""ifdefense_b200 -- B200-native (sm_100a) implementation of IF-Defense's optimisation-based restoration
path (ConvONet/opt_defense.py, ONet/opt_defense.py of Wuziyi616/IF-Defense) behind the reference's own
Python seams.

  capi        ctypes binding of libifd_b200.so (the C ABI in include/ifd_b200.h)
  weights     reference state_dict -> packed kernel parameter blobs
  defense     drop-in mirrors of the reference's defense/ package (repulsion_loss, knn_point, SORDefense, FPS)
  convonet    decode / optimize_points for the ConvONet variant
  models      nn.Modules with the reference's state_dict names (checkpoint interface) for the encoders

There is no CPU fallback: every op raises RuntimeError if the CUDA library or a B200 is missing.
"""
__version__ = "0.1.0"
