"""geoldm_b200 — B200-native (sm_100a) implementation of GeoLDM's EGNN-denoiser sampling hot path.

Only what that path needs lives here: ``csrc/`` (CUDA kernels + C ABI), the ctypes binding, the ragged
packer and host-side mirrors of the reference's module surface.  See DESIGN.md.
"""
from . import histograms  # noqa: F401

__all__ = ["histograms"]
