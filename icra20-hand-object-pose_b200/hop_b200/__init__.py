"""hop_b200 -- Python host mirror of the reference's operator interface over libhop.so (C ABI, include/hop_c_api.h).

The compute lives in hand-written sm_100a kernels inside libhop.so; this package only marshals buffers.  There is no
CPU path: importing works anywhere (the library loads without a GPU), creating a Context without a B200 raises.
"""
from .capi import (Context, Cloud, Mesh, RenderScene, FingerParams, CollisionParams, RenderParams, HandRemovalParams, HopError, IcpParams, LcpParams,
                   PoseRec, lib_path, load_library, build_library, declared_symbols)
from .pose_estimator import PoseEstimator, PoseHypo

__all__ = ["Context", "Cloud", "Mesh", "RenderScene", "CollisionParams", "RenderParams", "HandRemovalParams", "FingerParams", "HopError", "IcpParams", "LcpParams", "PoseRec", "PoseEstimator", "PoseHypo", "lib_path",
           "load_library", "build_library", "declared_symbols"]
