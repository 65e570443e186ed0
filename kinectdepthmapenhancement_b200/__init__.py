"""B200-native joint-bilateral depth enhancement (drop-in for the JointBilateralFilter /
guided cross-bilateral / Buffer2D path of stevesuyao/KinectDepthMapEnhancement).

All compute lives in libkdme_b200.so (hand-written CUDA for sm_100a behind the C ABI of
include/kdme_b200.h).  This package is the host-side mirror of the reference's classes.
"""
from ._lib import KdmeError, LIB_PATH, build  # noqa: F401


def __getattr__(name):
    if name in ("JointBilateralFilter", "projective_to_real"):
        from . import jbf
        return getattr(jbf, name)
    if name == "Buffer2D":
        from .buffer2d import Buffer2D
        return Buffer2D
    if name == "guided_fill":
        from .guided import guided_fill
        return guided_fill
    raise AttributeError(name)
