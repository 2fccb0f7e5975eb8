"""deluca-b200: google/deluca's batched rollout hot path as hand-written sm_100a CUDA.

The package is a thin Python mirror of the reference's Env/Agent interface over the C ABI in
``include/deluca_b200.h``; see DESIGN.md.  There is no CPU implementation in here.
"""
__version__ = "0.1.0"
