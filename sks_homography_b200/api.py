"""Host-side mirror of the reference's solver interface over libsks_cuda.

Names and argument meaning follow the reference so its callers (and our parity
tests) read the same:

  runKernel_ACA / _ACA_double / _SKS / _SKS_double   MOD/ACA_SKS.hpp:17-20
      (src, tar, result) with AoS points x0,y0,..,x3,y3 -> row-major 3x3,
      h33-normalised, returns 0.  Here src/tar/result are batches [n,8]/[n,9].
  ACA_rect(TargetPts, M_x, M_y, width, ratio_rec)    ML/ACA_rect.m:22
  TensorACA_rect(bs, src, tar, scale, div)           PY.py:286  (returns H, un-normalised)
  ACA_vanilla(bs, src, tar)                          PY.py:312  (returns H, un-normalised)

torch is used for memory and streams only.  CUDA tensors go straight to the
device-pointer C ABI on the current stream; CPU tensors go through the
host-pointer C ABI (sks_host_*), i.e. they are still solved on the GPU.  There
is no CPU compute path: without the library or a GPU every call raises.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import FLAG_NORMALIZE, LAYOUT_AOS, LAYOUT_SOA, lib

_SUFFIX = {torch.float32: "f32", torch.float64: "f64"}


# Small batches (the deep-homography sizes, bs = 64 ... 4096) are call-overhead bound: one kernel of a
# few microseconds behind ~20 us of Python.  The helpers below keep that layer thin: the raw stream handle
# without building a torch.cuda.Stream object, and a device guard that does nothing when the tensor
# already lives on the current device.
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream_ptr(t: torch.Tensor) -> int:
    if _raw_stream is not None:
        idx = t.device.index
        return _raw_stream(torch.cuda.current_device() if idx is None else idx)
    return torch.cuda.current_stream(t.device).cuda_stream


_FN = {}


def _entry(name: str):
    """ctypes entry point by name, cached (attribute lookup on a CDLL is not free)."""
    fn = _FN.get(name)
    if fn is None:
        fn = _FN[name] = getattr(lib().c, name)
    return fn


class _on_device:
    """`with torch.cuda.device(dev)` without its cost in the common case (already current)."""
    __slots__ = ("idx", "prev")

    def __init__(self, device):
        self.idx = device.index
        self.prev = None

    def __enter__(self):
        if self.idx is not None:
            cur = torch.cuda.current_device()
            if cur != self.idx:
                self.prev = cur
                torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


def _ptr(t):
    return None if t is None else t.data_ptr()


def _check_pair(src, tar, dtype):
    if src.dtype != dtype or tar.dtype != dtype:
        raise TypeError(f"expected {dtype} tensors, got {src.dtype}/{tar.dtype}")
    if src.device != tar.device:
        raise ValueError("src and tar must live on the same device")


def solve(solver: str, src: torch.Tensor, tar: torch.Tensor, result: torch.Tensor | None = None,
          normalize: bool = True, layout: str = "aos", degenerate: torch.Tensor | None = None
          ) -> torch.Tensor:
    """Batched general 4-point solve.  solver in {'aca','sks'}.

    aos: src/tar [n,8] (or [n,4,2]) -> H [n,9];  soa: src/tar [8,n] -> H [9,n].
    """
    dtype = src.dtype
    if dtype not in _SUFFIX:
        raise TypeError("float32 or float64 only")
    _check_pair(src, tar, dtype)
    L = lib()
    src, tar = src.contiguous(), tar.contiguous()
    if layout == "aos":
        n = src.numel() // 8
        shape, lay, ld = (n, 9), LAYOUT_AOS, 0
    elif layout == "soa":
        if src.dim() != 2 or src.shape[0] != 8:
            raise ValueError("soa layout expects [8, n]")
        n = src.shape[1]
        shape, lay, ld = (9, n), LAYOUT_SOA, n
    else:
        raise ValueError(layout)
    if result is None:
        result = torch.empty(shape, dtype=dtype, device=src.device)
    elif result.numel() != 9 * n or result.dtype != dtype or not result.is_contiguous():
        raise ValueError("result must be a contiguous tensor of 9*n elements")
    flags = FLAG_NORMALIZE if normalize else 0
    name = f"{solver}_{_SUFFIX[dtype]}"
    if src.is_cuda:
        with _on_device(src.device):
            rc = _entry("sks_cuda_" + name)(src.data_ptr(), tar.data_ptr(), result.data_ptr(), n, lay, ld, flags,
                                           _ptr(degenerate), _stream_ptr(src))
            if rc:
                L.check(rc, f"sks_cuda_{name}")
    else:
        if layout != "aos" or degenerate is not None:
            raise ValueError("host tensors: AoS layout without flag output only")
        fn = getattr(L.c, f"sks_host_{name}")
        L.check(fn(_ptr(src), _ptr(tar), _ptr(result), n, flags), f"sks_host_{name}")
    return result


def runKernel_ACA(src, tar, result=None, **kw):
    _check_pair(src, tar, torch.float32)
    return solve("aca", src, tar, result, **kw)


def runKernel_ACA_double(src, tar, result=None, **kw):
    _check_pair(src, tar, torch.float64)
    return solve("aca", src, tar, result, **kw)


def runKernel_SKS(src, tar, result=None, **kw):
    _check_pair(src, tar, torch.float32)
    return solve("sks", src, tar, result, **kw)


def runKernel_SKS_double(src, tar, result=None, **kw):
    _check_pair(src, tar, torch.float64)
    return solve("sks", src, tar, result, **kw)


def runKernel_GE(src, tar, result=None, **kw):
    """Competitor solver cv::runKernel_GE (MOD/GE.hpp:9): RHO Gaussian elimination,
    fp32 in the reference; float64 tensors run the same elimination in fp64
    (the arithmetic of the reference's CUDA kernel cal_Homo_GE, GPU.cu:359-507)."""
    _check_pair(src, tar, src.dtype)
    return solve("ge", src, tar, result, **kw)


def runKernel_GPT(src, tar, result=None, **kw):
    """Competitor GPT-LU in the arithmetic of the reference's CUDA kernel cal_Homo_GPT
    (GPU.cu:242-357): float64 device tensors only (the reference's CPU form is OpenCV's
    getPerspectiveTransform and is not part of this library)."""
    _check_pair(src, tar, torch.float64)
    if not src.is_cuda:
        raise ValueError("runKernel_GPT takes device tensors")
    return solve("gpt", src, tar, result, **kw)


def aca_rect(tar: torch.Tensor, width: float, ratio: float, M_x: float = 0.0, M_y: float = 0.0,
             M: torch.Tensor | None = None, result: torch.Tensor | None = None,
             normalize: bool = True, layout: str = "aos",
             degenerate: torch.Tensor | None = None) -> torch.Tensor:
    """ACA-rect on interleaved target corners tar [n,8] (TL,TR,BL,BR)."""
    dtype = tar.dtype
    if dtype not in _SUFFIX:
        raise TypeError("float32 or float64 only")
    L = lib()
    tar = tar.contiguous()
    if layout == "aos":
        n = tar.numel() // 8
        shape, lay, ld = (n, 9), LAYOUT_AOS, 0
    else:
        n = tar.shape[1]
        shape, lay, ld = (9, n), LAYOUT_SOA, n
    if M is not None:
        M = M.contiguous()
        if M.dtype != dtype or M.numel() != 2 * n:
            raise ValueError("M must hold 2*n values of tar's dtype")
    if result is None:
        result = torch.empty(shape, dtype=dtype, device=tar.device)
    flags = FLAG_NORMALIZE if normalize else 0
    name = f"aca_rect_{_SUFFIX[dtype]}"
    if tar.is_cuda:
        with _on_device(tar.device):
            fn = getattr(L.c, f"sks_cuda_{name}")
            L.check(fn(_ptr(tar), _ptr(M), M_x, M_y, width, ratio, _ptr(result), n, lay, ld, flags,
                       _ptr(degenerate), _stream_ptr(tar)), f"sks_cuda_{name}")
    else:
        if layout != "aos" or degenerate is not None:
            raise ValueError("host tensors: AoS layout without flag output only")
        fn = getattr(L.c, f"sks_host_{name}")
        L.check(fn(_ptr(tar), _ptr(M), M_x, M_y, width, ratio, _ptr(result), n, flags),
                f"sks_host_{name}")
    return result


def ACA_rect(TargetPts: torch.Tensor, M_x: float, M_y: float, width: float, ratio_rec: float
             ) -> torch.Tensor:
    """MATLAB signature (ML/ACA_rect.m:22): TargetPts is 3x4 (or [bs,3,4])
    homogeneous, columns TL,TR,BL,BR; returns the h33-normalised 3x3."""
    single = TargetPts.dim() == 2
    T = TargetPts.reshape(-1, 3, 4)
    tar = T[:, :2, :].transpose(1, 2).reshape(-1, 8)
    H = aca_rect(tar, width, ratio_rec, M_x, M_y, normalize=True).reshape(-1, 3, 3)
    return H[0] if single else H


def TensorACA_rect(bs: int, src: torch.Tensor, tar: torch.Tensor, scale, div) -> torch.Tensor:
    """PY.py:286-309 with the tensor conventions of PY.py:24-37: src/tar are
    [bs,3,4] homogeneous (rows x,y,1; columns TL,TR,BL,BR); scale = width and
    div = width/height of the shared source square; the per-sample corner is
    src[:, 0:2, 0] (PY.py:302).  Returns the un-normalised H [bs,3,3] that the
    reference computes and discards."""
    if tar.is_cuda and tar.dtype in _SUFFIX and src.dtype == tar.dtype and tar.shape[1:] == (3, 4):
        # read the [bs,3,4] tensors in place (no transposing copy)
        L = lib()
        tar, src = tar.contiguous(), src.contiguous()
        H = torch.empty((bs, 3, 3), dtype=tar.dtype, device=tar.device)
        on_dev = all(torch.is_tensor(x) and x.device == tar.device and x.dtype == tar.dtype and x.numel() == 1
                     for x in (scale, div))
        with _on_device(tar.device):
            if on_dev:      # the reference's own calling convention: scale / div are device tensors; no sync
                fn = getattr(L.c, f"sks_cuda_aca_rect_planar_dev_{_SUFFIX[tar.dtype]}")
                L.check(fn(_ptr(tar), _ptr(src), _ptr(scale.contiguous()), _ptr(div.contiguous()), _ptr(H), bs, 0,
                           None, _stream_ptr(tar)), "sks_cuda_aca_rect_planar_dev")
            else:
                fn = getattr(L.c, f"sks_cuda_aca_rect_planar_{_SUFFIX[tar.dtype]}")
                L.check(fn(_ptr(tar), _ptr(src), 0.0, 0.0, float(scale), float(div), _ptr(H), bs, 0, None,
                           _stream_ptr(tar)), "sks_cuda_aca_rect_planar")
        return H
    tarq = tar[:, :2, :].transpose(1, 2).reshape(bs, 8)
    M = src[:, :2, 0].contiguous()
    H = aca_rect(tarq, float(scale), float(div), M=M, normalize=False)
    return H.reshape(bs, 3, 3)


def ACA_vanilla(bs: int, src: torch.Tensor, tar: torch.Tensor) -> torch.Tensor:
    """PY.py:312-388: src/tar [bs,4,2] -> un-normalised H [bs,3,3]."""
    H = solve("aca", src.reshape(bs, 8), tar.reshape(bs, 8), normalize=False)
    return H.reshape(bs, 3, 3)


# ------------------------------------------------------------ synthetic data
def synth_quads(n: int, seed: int = 11, dist: int = _lib.DIST_DEEP, dtype=torch.float32,
                device="cuda", begin: int = 0, layout: str = "aos"):
    """Device-generated quadruples [begin, begin+n) (bit-identical to the oracle's)."""
    L = lib()
    dev = torch.device(device)
    shape = (n, 8) if layout == "aos" else (8, n)
    src = torch.empty(shape, dtype=dtype, device=dev)
    tar = torch.empty(shape, dtype=dtype, device=dev)
    with _on_device(dev):
        fn = getattr(L.c, f"sks_cuda_synth_quads_{_SUFFIX[dtype]}")
        L.check(fn(_ptr(src), _ptr(tar), begin, n, seed, dist,
                   LAYOUT_AOS if layout == "aos" else LAYOUT_SOA, n, _stream_ptr(src)),
                "sks_cuda_synth_quads")
    return src, tar


def synth_corr(n_pairs: int, n_pts: int, seed: int = 11, inlier_permille: int = 500,
               noise: float = 0.5, device="cuda", pair_begin: int = 0) -> torch.Tensor:
    L = lib()
    dev = torch.device(device)
    corr = torch.empty((n_pairs, n_pts, 4), dtype=torch.float32, device=dev)
    with _on_device(dev):
        L.check(L.c.sks_cuda_synth_corr_f32(_ptr(corr), pair_begin, n_pairs, n_pts, seed,
                                            inlier_permille, noise, _stream_ptr(corr)),
                "sks_cuda_synth_corr_f32")
    return corr


def gather_samples(pool: torch.Tensor, n: int, seed: int = 11, rand4: torch.Tensor | None = None,
                   layout: str = "aos"):
    """GPU.cu:52-78: n minimal samples from a match pool [size,4] = (x,y,X,Y)."""
    L = lib()
    dtype = pool.dtype
    pool = pool.contiguous()
    shape = (n, 8) if layout == "aos" else (8, n)
    src = torch.empty(shape, dtype=dtype, device=pool.device)
    tar = torch.empty(shape, dtype=dtype, device=pool.device)
    with _on_device(pool.device):
        fn = getattr(L.c, f"sks_cuda_gather_samples_{_SUFFIX[dtype]}")
        L.check(fn(_ptr(pool), pool.shape[0], _ptr(rand4), seed, _ptr(src), _ptr(tar), n,
                   LAYOUT_AOS if layout == "aos" else LAYOUT_SOA, n, _stream_ptr(pool)),
                "sks_cuda_gather_samples")
    return src, tar


def gather_solve(solver: str, pool: torch.Tensor, n: int, seed: int = 11,
                 rand4: torch.Tensor | None = None, normalize: bool = True, layout: str = "aos"
                 ) -> torch.Tensor:
    """Fused get_rand_list + cal_Homo_* (GPU.cu:1449-1464): n hypotheses straight from
    a match pool [size,4]; the sampled quadruples never touch HBM."""
    L = lib()
    dtype = pool.dtype
    pool = pool.contiguous()
    H = torch.empty((n, 9) if layout == "aos" else (9, n), dtype=dtype, device=pool.device)
    with _on_device(pool.device):
        fn = getattr(L.c, f"sks_cuda_gather_{solver}_{_SUFFIX[dtype]}")
        L.check(fn(_ptr(pool), pool.shape[0], _ptr(rand4), seed, _ptr(H), n,
                   LAYOUT_AOS if layout == "aos" else LAYOUT_SOA, n,
                   FLAG_NORMALIZE if normalize else 0, None, _stream_ptr(pool)),
                "sks_cuda_gather_solve")
    return H


def warp_grid(H: torch.Tensor, gw: int, gh: int, x0: float = 0.0, y0: float = 0.0, dx: float = 1.0,
              dy: float = 1.0, out: torch.Tensor | None = None) -> torch.Tensor:
    """Sampling grid [n, gh, gw, 2] = H * (x0 + i*dx, y0 + j*dy, 1), dehomogenised; H [n,9] or
    [n,3,3] fp32 at any scale (the step after the solver in a deep-homography pipeline)."""
    L = lib()
    H = H.contiguous()
    n = H.numel() // 9
    if out is None:
        out = torch.empty((n, gh, gw, 2), dtype=torch.float32, device=H.device)
    with _on_device(H.device):
        L.check(L.c.sks_cuda_warp_grid_f32(_ptr(H), n, x0, y0, dx, dy, gw, gh, _ptr(out), _stream_ptr(H)),
                "sks_cuda_warp_grid_f32")
    return out


def aca_rect_warp_grid(tar: torch.Tensor, width: float, ratio: float, gw: int, gh: int,
                       M_x: float = 0.0, M_y: float = 0.0, M: torch.Tensor | None = None,
                       x0: float = 0.0, y0: float = 0.0, dx: float = 1.0, dy: float = 1.0,
                       out: torch.Tensor | None = None) -> torch.Tensor:
    """ACA-rect on tar [n,8] fused with warp_grid: the homography stays in registers (no
    normalisation, ML/ACA_rect.m:33-35) and only the grid [n, gh, gw, 2] is written."""
    L = lib()
    tar = tar.contiguous()
    n = tar.numel() // 8
    if M is not None:
        M = M.contiguous()
    if out is None:
        out = torch.empty((n, gh, gw, 2), dtype=torch.float32, device=tar.device)
    with _on_device(tar.device):
        L.check(L.c.sks_cuda_aca_rect_warp_grid_f32(_ptr(tar), _ptr(M), M_x, M_y, width, ratio, n, x0, y0,
                                                    dx, dy, gw, gh, _ptr(out), _stream_ptr(tar)),
                "sks_cuda_aca_rect_warp_grid_f32")
    return out


def curand_mrg32k3a(n: int, seed: int = 11, device="cuda", out: torch.Tensor | None = None) -> torch.Tensor:
    """n uint32 draws (as int32 bits) identical to cuRAND's host-API MRG32K3A generator
    with this seed -- the reference's sample list (GPU.cu:1443-1446) for n = 4*numsOfH;
    reshape to [4, numsOfH] and pass as ``rand4`` to gather_samples / gather_solve."""
    L = lib()
    if out is None:
        out = torch.empty(n, dtype=torch.int32, device=torch.device(device))
    elif out.numel() != n or out.dtype != torch.int32 or not out.is_contiguous():
        raise ValueError("out must be a contiguous int32 tensor of n elements")
    with _on_device(out.device):
        L.check(L.c.sks_cuda_curand_mrg32k3a_u32(_ptr(out), n, seed, _stream_ptr(out)),
                "sks_cuda_curand_mrg32k3a_u32")
    return out


# ------------------------------------------------------------------- RANSAC
def _check_samples(samples, corr, n_hyp: int):
    """An explicit sample list must be [P, n_hyp, 4] 32-bit integers, contiguous, on corr's device
    (the kernel reads it as uint4 rows; a wider dtype would be silently misread)."""
    if samples is None:
        return None
    P = corr.shape[0]
    if samples.dtype not in (torch.int32, torch.uint32):
        raise TypeError(f"samples must be int32/uint32, got {samples.dtype}")
    if samples.device != corr.device:
        raise ValueError(f"samples on {samples.device}, correspondences on {corr.device}")
    if samples.numel() != P * n_hyp * 4:
        raise ValueError(f"samples has {samples.numel()} elements, expected P*n_hyp*4 = {P * n_hyp * 4}")
    return samples.contiguous()


def ransac_keys(corr: torch.Tensor, n_hyp: int, seed: int, thr2: float,
                samples: torch.Tensor | None = None, hyp_begin: int = 0,
                hyp_count: int | None = None, out: torch.Tensor | None = None,
                pair_begin: int = 0) -> torch.Tensor:
    """Score hypothesis ids [hyp_begin, hyp_begin+hyp_count) of every pair and
    max-combine into `out` (int64 view of the packed uint64 keys).  pair_begin: global id of
    corr[0] when `corr` is one rank's shard of the image pairs (keys the sampler only)."""
    L = lib()
    corr = corr.contiguous()
    P, n_pts, _ = corr.shape
    samples = _check_samples(samples, corr, n_hyp)
    hyp_count = n_hyp - hyp_begin if hyp_count is None else hyp_count
    if out is None:
        out = torch.zeros(P, dtype=torch.int64, device=corr.device)
    with _on_device(corr.device):
        L.check(L.c.sks_cuda_ransac_aca_shard_f32(_ptr(corr), pair_begin, P, n_pts, _ptr(samples), n_hyp,
                                                  hyp_begin, hyp_count, seed, thr2, _ptr(out),
                                                  _stream_ptr(corr)),
                "sks_cuda_ransac_aca_shard_f32")
    return out


def ransac_host(corr: torch.Tensor, n_hyp: int, seed: int, thr2: float,
                samples: torch.Tensor | None = None, want_mask: bool = False, ngpu: int | None = None):
    """Whole estimate from HOST tensors through sks_host_ransac_aca_f32: corr [P, n_pts, 4] fp32 in
    host memory (pinned or pageable) -> (H [P,9], count [P], mask [P,n_pts] | None, keys [P])."""
    L = lib()
    if corr.is_cuda or corr.dtype != torch.float32:
        raise ValueError("ransac_host takes float32 host tensors")
    corr = corr.contiguous()
    P, n_pts, _ = corr.shape
    H = torch.empty((P, 9), dtype=torch.float32)
    cnt = torch.empty(P, dtype=torch.int32)
    keys = torch.empty(P, dtype=torch.int64)
    mask = torch.empty((P, n_pts), dtype=torch.uint8) if want_mask else None
    samples = _check_samples(samples, corr, n_hyp)
    if ngpu is None:       # one GPU, or whatever sks_host_set_device_count() says
        L.check(L.c.sks_host_ransac_aca_f32(_ptr(corr), P, n_pts, _ptr(samples), n_hyp, seed, thr2, _ptr(H),
                                            _ptr(cnt), _ptr(mask), _ptr(keys)), "sks_host_ransac_aca_f32")
    else:                  # in-library multi-GPU driver (0 = all visible GPUs)
        L.check(L.c.sks_host_ransac_aca_multi_f32(_ptr(corr), P, n_pts, _ptr(samples), n_hyp, seed, thr2, ngpu,
                                                  _ptr(H), _ptr(cnt), _ptr(mask), _ptr(keys)),
                "sks_host_ransac_aca_multi_f32")
    return H, cnt, mask, keys


def ransac_multi(corr: torch.Tensor, n_hyp: int, seed: int, thr2: float, ngpu: int = 0,
                 samples: torch.Tensor | None = None, want_mask: bool = False, finalize: bool = True):
    """Whole estimate over `ngpu` GPUs of THIS process (0 = all visible) through
    sks_cuda_ransac_aca_multi_f32: corr [P, n_pts, 4] lives on its own device, which also receives
    every output; the other devices read it over NVLink peer access and merge their winners with
    peer atomics (csrc/multi.cu).  Returns (H [P,9] | None, count [P] | None, mask | None, keys [P])."""
    L = lib()
    corr = corr.contiguous()
    P, n_pts, _ = corr.shape
    samples = _check_samples(samples, corr, n_hyp)
    dev = corr.device
    keys = torch.empty(P, dtype=torch.int64, device=dev)
    H = torch.empty((P, 9), dtype=torch.float32, device=dev) if finalize else None
    cnt = torch.empty(P, dtype=torch.int32, device=dev) if finalize else None
    mask = torch.empty((P, n_pts), dtype=torch.uint8, device=dev) if (finalize and want_mask) else None
    with _on_device(dev):
        L.check(L.c.sks_cuda_ransac_aca_multi_f32(_ptr(corr), P, n_pts, _ptr(samples), n_hyp, seed, thr2, ngpu,
                                                  _ptr(keys), _ptr(H), _ptr(cnt), _ptr(mask), _stream_ptr(corr)),
                "sks_cuda_ransac_aca_multi_f32")
    return H, cnt, mask, keys


def ransac_finalize(corr: torch.Tensor, n_hyp: int, seed: int, thr2: float, keys: torch.Tensor,
                    samples: torch.Tensor | None = None, want_mask: bool = False, pair_begin: int = 0):
    L = lib()
    corr = corr.contiguous()
    P, n_pts, _ = corr.shape
    samples = _check_samples(samples, corr, n_hyp)
    if keys.dtype != torch.int64 or keys.numel() != P or keys.device != corr.device or not keys.is_contiguous():
        raise ValueError("keys must be a contiguous int64 tensor of P elements on corr's device")
    H = torch.empty((P, 9), dtype=torch.float32, device=corr.device)
    cnt = torch.empty(P, dtype=torch.int32, device=corr.device)
    mask = torch.empty((P, n_pts), dtype=torch.uint8, device=corr.device) if want_mask else None
    with _on_device(corr.device):
        L.check(L.c.sks_cuda_ransac_finalize_shard_f32(_ptr(corr), pair_begin, P, n_pts, _ptr(samples), n_hyp,
                                                 seed, thr2, _ptr(keys), _ptr(H), _ptr(cnt), _ptr(mask),
                                                 _stream_ptr(corr)),
                "sks_cuda_ransac_finalize_f32")
    return H, cnt, mask


def ransac_score(corr: torch.Tensor, H: torch.Tensor, thr2: float, want_mask: bool = True):
    """Inlier count (and mask) of GIVEN models H [P,9] under the library's inlier rule."""
    L = lib()
    corr, H = corr.contiguous(), H.contiguous()
    P, n_pts, _ = corr.shape
    cnt = torch.empty(P, dtype=torch.int32, device=corr.device)
    mask = torch.empty((P, n_pts), dtype=torch.uint8, device=corr.device) if want_mask else None
    with _on_device(corr.device):
        L.check(L.c.sks_cuda_ransac_score_f32(_ptr(corr), P, n_pts, _ptr(H), thr2, _ptr(cnt), _ptr(mask),
                                              _stream_ptr(corr)), "sks_cuda_ransac_score_f32")
    return cnt, mask


def ransac_polish(corr: torch.Tensor, H: torch.Tensor, thr2: float, iters: int = 2):
    """LO-style polishing of RANSAC winners, entirely on the device: score -> least-squares refit on
    the inlier mask -> score again, keeping a refit only where it explains MORE matches; repeated
    `iters` times.  Returns (H [P,9], inlier_count [P], mask [P,n_pts])."""
    cnt, mask = ransac_score(corr, H, thr2)
    for _ in range(iters):
        H2, used = ransac_refit(corr, mask, H)
        cnt2, mask2 = ransac_score(corr, H2, thr2)
        better = cnt2 > cnt
        H = torch.where(better[:, None], H2, H)
        mask = torch.where(better[:, None], mask2, mask)
        cnt = torch.where(better, cnt2, cnt)
    return H, cnt, mask


def ransac_refit(corr: torch.Tensor, mask: torch.Tensor, H: torch.Tensor):
    """Least-squares re-estimation of every pair's homography from all matches flagged in
    `mask` [P, n_pts] (uint8, as returned by ransac_finalize(..., want_mask=True)); pairs
    without a usable solution keep `H`.  Returns (H_refit [P,9], n_used [P])."""
    L = lib()
    corr, mask, H = corr.contiguous(), mask.contiguous(), H.contiguous()
    P, n_pts, _ = corr.shape
    out = torch.empty((P, 9), dtype=torch.float32, device=corr.device)
    used = torch.empty(P, dtype=torch.int32, device=corr.device)
    with _on_device(corr.device):
        L.check(L.c.sks_cuda_ransac_refit_f32(_ptr(corr), P, n_pts, _ptr(mask), _ptr(H), _ptr(out), _ptr(used),
                                              _stream_ptr(corr)), "sks_cuda_ransac_refit_f32")
    return out, used


def decode_keys(keys: torch.Tensor):
    """packed key -> (inlier count, hypothesis id)"""
    count = keys >> 32
    hyp = 0xFFFFFFFF - (keys & 0xFFFFFFFF)
    return count, hyp
