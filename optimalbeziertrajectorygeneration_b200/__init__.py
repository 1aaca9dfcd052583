"""B200-native (sm_100a) Bezier constraint + Jacobian evaluation: a drop-in for
the hot path of caslabuiowa/OptimalBezierTrajectoryGeneration.

    from optimalbeziertrajectorygeneration_b200 import bezier, optimization

mirrors the reference's ``bezier`` / ``optimization`` modules.  The arithmetic
runs in libbezgpu.so (hand-written CUDA, C-ABI in include/bezgpu.h); importing
the sub-modules without the built library raises -- there is no CPU fallback.
"""
__version__ = "0.1.0"
