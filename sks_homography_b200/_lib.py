"""ctypes binding of libsks_cuda.so -- one Python method per C-ABI entry point.

The library is the product; this file only marshals pointers.  It fails loudly
when the shared object is missing or cannot be loaded: there is no CPU or
PyTorch fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsks_cuda.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "sks_cuda.h")

OK, ERR_INVALID_ARG, ERR_UNALIGNED, ERR_NO_DEVICE, ERR_NO_PEER_ACCESS = 0, -1, -2, -3, -4
LAYOUT_AOS, LAYOUT_SOA = 0, 1
FLAG_NORMALIZE = 1
DIST_DEEP, DIST_IMAGE, DIST_DEEP_INT = 0, 1, 2


class SksCudaError(RuntimeError):
    def __init__(self, status: int, what: str, msg: str):
        super().__init__(f"{what} failed: status {status} ({msg})")
        self.status = status


def declared_symbols(header: str = HEADER) -> list[str]:
    """Every function name declared in include/sks_cuda.h."""
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sks_(?:cuda|host)_\w+)\s*\(", text)))


_vp, _i64, _i32, _u32, _u64, _int = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.c_uint64, C.c_int
_f32, _f64 = C.c_float, C.c_double

_GENERAL = [_vp, _vp, _vp, _i64, _int, _i64, _int, _vp, _vp]
_SIGS = {
    "sks_cuda_abi_version": (_int, []),
    "sks_cuda_error_string": (C.c_char_p, [_int]),
    "sks_cuda_device_count": (_int, [C.POINTER(_int)]),
    "sks_cuda_aca_f32": (_int, _GENERAL),
    "sks_cuda_aca_f64": (_int, _GENERAL),
    "sks_cuda_sks_f32": (_int, _GENERAL),
    "sks_cuda_sks_f64": (_int, _GENERAL),
    "sks_cuda_ge_f32": (_int, _GENERAL),
    "sks_cuda_ge_f64": (_int, _GENERAL),
    "sks_cuda_gpt_f64": (_int, _GENERAL),
    "sks_cuda_aca_rect_f32": (_int, [_vp, _vp, _f32, _f32, _f32, _f32, _vp, _i64, _int, _i64, _int, _vp, _vp]),
    "sks_cuda_aca_rect_f64": (_int, [_vp, _vp, _f64, _f64, _f64, _f64, _vp, _i64, _int, _i64, _int, _vp, _vp]),
    "sks_cuda_aca_rect_planar_f32": (_int, [_vp, _vp, _f32, _f32, _f32, _f32, _vp, _i64, _int, _vp, _vp]),
    "sks_cuda_aca_rect_planar_f64": (_int, [_vp, _vp, _f64, _f64, _f64, _f64, _vp, _i64, _int, _vp, _vp]),
    "sks_cuda_aca_rect_planar_dev_f32": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _int, _vp, _vp]),
    "sks_cuda_aca_rect_planar_dev_f64": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _int, _vp, _vp]),
    "sks_host_aca_f32": (_int, [_vp, _vp, _vp, _i64, _int]),
    "sks_host_aca_f64": (_int, [_vp, _vp, _vp, _i64, _int]),
    "sks_host_sks_f32": (_int, [_vp, _vp, _vp, _i64, _int]),
    "sks_host_sks_f64": (_int, [_vp, _vp, _vp, _i64, _int]),
    "sks_host_ge_f32": (_int, [_vp, _vp, _vp, _i64, _int]),
    "sks_host_ge_f64": (_int, [_vp, _vp, _vp, _i64, _int]),
    "sks_host_aca_rect_f32": (_int, [_vp, _vp, _f32, _f32, _f32, _f32, _vp, _i64, _int]),
    "sks_host_aca_rect_f64": (_int, [_vp, _vp, _f64, _f64, _f64, _f64, _vp, _i64, _int]),
    "sks_host_ransac_aca_f32": (_int, [_vp, _i64, _i32, _vp, _u32, _u64, _f32, _vp, _vp, _vp, _vp]),
    "sks_host_ransac_aca_multi_f32": (_int, [_vp, _i64, _i32, _vp, _u32, _u64, _f32, _int, _vp, _vp, _vp, _vp]),
    "sks_cuda_ransac_aca_multi_f32": (_int, [_vp, _i64, _i32, _vp, _u32, _u64, _f32, _int, _vp, _vp, _vp, _vp, _vp]),
    "sks_host_set_device_count": (_int, [_int]),
    "sks_host_set_staging_copy": (_int, [_int]),
    "sks_host_set_staging_threads": (_int, [_int]),
    "sks_host_set_chunk_bytes": (_int, [_i64]),
    "sks_host_alloc_pinned": (_int, [C.POINTER(_vp), _i64]),
    "sks_host_free_pinned": (_int, [_vp]),
    "sks_host_register": (_int, [_vp, _i64]),
    "sks_host_unregister": (_int, [_vp]),
    "sks_cuda_gather_samples_f32": (_int, [_vp, _u32, _vp, _u64, _vp, _vp, _i64, _int, _i64, _vp]),
    "sks_cuda_gather_samples_f64": (_int, [_vp, _u32, _vp, _u64, _vp, _vp, _i64, _int, _i64, _vp]),
    "sks_cuda_gather_aca_f32": (_int, [_vp, _u32, _vp, _u64, _vp, _i64, _int, _i64, _int, _vp, _vp]),
    "sks_cuda_gather_aca_f64": (_int, [_vp, _u32, _vp, _u64, _vp, _i64, _int, _i64, _int, _vp, _vp]),
    "sks_cuda_gather_sks_f32": (_int, [_vp, _u32, _vp, _u64, _vp, _i64, _int, _i64, _int, _vp, _vp]),
    "sks_cuda_gather_sks_f64": (_int, [_vp, _u32, _vp, _u64, _vp, _i64, _int, _i64, _int, _vp, _vp]),
    "sks_cuda_ransac_score_f32": (_int, [_vp, _i64, _i32, _vp, _f32, _vp, _vp, _vp]),
    "sks_cuda_ransac_refit_f32": (_int, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp]),
    "sks_cuda_warp_grid_f32": (_int, [_vp, _i64, _f32, _f32, _f32, _f32, _i32, _i32, _vp, _vp]),
    "sks_cuda_aca_rect_warp_grid_f32": (_int, [_vp, _vp, _f32, _f32, _f32, _f32, _i64, _f32, _f32, _f32, _f32,
                                        _i32, _i32, _vp, _vp]),
    "sks_cuda_curand_mrg32k3a_u32": (_int, [_vp, _i64, _u64, _vp]),
    "sks_cuda_ransac_aca_f32": (_int, [_vp, _i64, _i32, _vp, _u32, _u32, _u32, _u64, _f32, _vp, _vp]),
    "sks_cuda_ransac_aca_shard_f32": (_int, [_vp, _i64, _i64, _i32, _vp, _u32, _u32, _u32, _u64, _f32, _vp, _vp]),
    "sks_cuda_ransac_finalize_shard_f32": (_int, [_vp, _i64, _i64, _i32, _vp, _u32, _u64, _f32, _vp, _vp, _vp, _vp, _vp]),
    "sks_cuda_ransac_finalize_f32": (_int, [_vp, _i64, _i32, _vp, _u32, _u64, _f32, _vp, _vp, _vp, _vp, _vp]),
    "sks_cuda_synth_quads_f32": (_int, [_vp, _vp, _i64, _i64, _u64, _int, _int, _i64, _vp]),
    "sks_cuda_synth_quads_f64": (_int, [_vp, _vp, _i64, _i64, _u64, _int, _int, _i64, _vp]),
    "sks_cuda_synth_corr_f32": (_int, [_vp, _i64, _i64, _i32, _u64, _int, _f32, _vp]),
    "sks_cuda_peer_alloc": (_int, [C.POINTER(_vp), _i64]),
    "sks_cuda_peer_free": (_int, [_vp]),
    "sks_cuda_peer_export": (_int, [_vp, _vp]),
    "sks_cuda_peer_open": (_int, [_vp, C.POINTER(_vp)]),
    "sks_cuda_peer_close": (_int, [_vp]),
    "sks_cuda_peer_push_max": (_int, [_vp, _i64, C.POINTER(_vp), _int, _int, _u64, _vp]),
    "sks_cuda_peer_wait": (_int, [_vp, _int, _u64, _vp, _i64, _vp, C.c_double, _vp]),
    "sks_cuda_shard_range": (_int, [_i64, _int, _int, C.POINTER(_i64), C.POINTER(_i64)]),
    "sks_cuda_launch_count": (_i64, []),
    "sks_cuda_reset_launch_count": (None, []),
    "sks_cuda_set_variant": (_int, [_int]),
    "sks_cuda_get_variant": (_int, []),
    "sks_cuda_set_tuning": (_int, [_int, _int, _int]),
    "sks_cuda_set_ransac_tuning": (_int, [_int, _int, _int]),
    "sks_cuda_ransac_chunk_plan": (_u32, [_i64, _u32, _int, _int, _int, _int]),
    "sks_cuda_shutdown": (_int, []),
}


class SksCuda:
    """Loaded libsks_cuda.so with typed entry points (`lib.c.<symbol>`)."""

    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found. Build it with `python -m sks_homography_b200.build` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        self.path = path
        self.c = C.CDLL(path)
        for name, (res, args) in _SIGS.items():
            fn = getattr(self.c, name)   # AttributeError if the ABI lost a symbol
            fn.restype, fn.argtypes = res, args

    def error_string(self, status: int) -> str:
        return self.c.sks_cuda_error_string(status).decode()

    def check(self, status: int, what: str) -> None:
        if status != OK:
            raise SksCudaError(status, what, self.error_string(status))

    def shard_range(self, n: int, rank: int, world: int) -> tuple[int, int]:
        b, c = _i64(), _i64()
        self.check(self.c.sks_cuda_shard_range(n, rank, world, C.byref(b), C.byref(c)), "shard_range")
        return b.value, c.value

    def device_count(self) -> int:
        n = _int()
        st = self.c.sks_cuda_device_count(C.byref(n))
        return n.value if st == OK else 0


_LIB: SksCuda | None = None


def lib() -> SksCuda:
    global _LIB
    if _LIB is None:
        _LIB = SksCuda()
    return _LIB
