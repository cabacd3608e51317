"""B200-native spectral encoding front end (points -> 800-D descriptor).

Drop-in for ``encoding.spectral_encoder.SpectralEncoder`` of Kimun-Park/Neural-Spectral-Codec;
all arithmetic runs in hand-written sm_100a CUDA behind ``include/nsc_b200.h``.
"""
from .encoder import (RangeImageProjector, SpectralEncoder, interpolate_range_image,  # noqa: F401
                      test_rotation_invariance)

__all__ = ["SpectralEncoder", "RangeImageProjector", "interpolate_range_image",
           "test_rotation_invariance"]
