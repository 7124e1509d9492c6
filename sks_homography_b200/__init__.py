"""sks_homography_b200 -- B200-native batched 4-point homography engine.

The product is libsks_cuda.so (csrc/, C ABI in include/sks_cuda.h).  This
package is the thin Python host side: a ctypes binding (_lib), the mirror of the
reference's solver interface (api) and the one-process-per-GPU driver (dist).
"""
from ._lib import (DIST_DEEP, DIST_DEEP_INT, DIST_IMAGE, FLAG_NORMALIZE, LAYOUT_AOS, LAYOUT_SOA,
                   SksCuda, SksCudaError, declared_symbols, lib)

__all__ = ["SksCuda", "SksCudaError", "lib", "declared_symbols", "LAYOUT_AOS", "LAYOUT_SOA",
           "FLAG_NORMALIZE", "DIST_DEEP", "DIST_IMAGE", "DIST_DEEP_INT"]
