"""bcp_b200 -- B200-native (sm_100a) implementation of the BCP semi-supervised segmentation training hot path.

Importing the package never falls back to a CPU or library path: kernels live in libbcp_b200.so (built in-tree by
``bcp_b200.build``) and every op raises if the library or a CUDA device is missing.
"""
from . import _native  # noqa: F401

__all__ = ["ops", "networks", "utils", "pancreas", "step"]
