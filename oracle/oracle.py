"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the CPU oracle.

``Oracle``  wraps oracle/libsks_oracle.so (our C restatement, sks_oracle.c).
``RefLib``  wraps oracle/_ref/libsks_ref.so (the reference's own
            "C++ Codes/modules/ACA_SKS.cpp", compiled in place by oracle/Makefile).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  The product package
(sks_homography_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libsks_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsks_ref.so")
REF_O3_SO = os.path.join(HERE, "_ref", "libsks_ref_o3.so")
REFGPU_SO = os.path.join(HERE, "_ref", "libsks_refgpu.so")
REFGPU_NOFMA_SO = os.path.join(HERE, "_ref", "libsks_refgpu_nofma.so")
REF_TORCH_PY = os.path.join(HERE, "_ref", "ref_torch_funcs.py")     # staged by oracle/stage_ref.py
REF_FIXTURE = os.path.join(HERE, "_ref", "orig_pts_wall.txt")         # CPU/orig_pts_wall.txt, staged likewise
REFERENCE_ROOT = "/root/reference"


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when the reference checkout exists)."""
    need = force or not os.path.exists(ORACLE_SO)
    if os.path.isdir(REFERENCE_ROOT) and not all(os.path.exists(f) for f in (
            REF_SO, REF_O3_SO, REFGPU_SO, REFGPU_NOFMA_SO, REF_TORCH_PY, REF_FIXTURE)):
        need = True
    if need:
        subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Oracle:
    """Our CPU restatement.  All arrays are AoS: src/tar [n,8], H [n,9]."""

    def __init__(self):
        build()
        self.lib = C.CDLL(ORACLE_SO)
        L = self.lib
        L.oracle_rng_u64.restype = C.c_uint64
        L.oracle_rng_u64.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        L.oracle_ransac_count_f32.restype = C.c_uint32

    def solve(self, solver: str, src, tar, normalize: bool = True) -> np.ndarray:
        dt = np.asarray(src).dtype
        assert dt in (np.float32, np.float64)
        src, tar = _c(src, dt).reshape(-1, 8), _c(tar, dt).reshape(-1, 8)
        n = src.shape[0]
        H = np.empty((n, 9), dtype=dt)
        fn = getattr(self.lib, f"oracle_{solver}_{'f32' if dt == np.float32 else 'f64'}")
        fn(_p(src), _p(tar), _p(H), C.c_int64(n), C.c_int(int(normalize)))
        return H

    def aca_rect(self, tar, mx, my, width, ratio, M=None, normalize: bool = True) -> np.ndarray:
        dt = np.asarray(tar).dtype
        tar = _c(tar, dt).reshape(-1, 8)
        n = tar.shape[0]
        H = np.empty((n, 9), dtype=dt)
        sc = C.c_float if dt == np.float32 else C.c_double
        Mp = None
        if M is not None:
            M = _c(M, dt).reshape(n, 2)
            Mp = _p(M)
        fn = getattr(self.lib, f"oracle_aca_rect_{'f32' if dt == np.float32 else 'f64'}")
        fn(_p(tar), Mp, sc(mx), sc(my), sc(width), sc(ratio), _p(H), C.c_int64(n),
           C.c_int(int(normalize)))
        return H

    def degenerate(self, H, normalized: bool = True) -> np.ndarray:
        dt = np.asarray(H).dtype
        H = _c(H, dt).reshape(-1, 9)
        out = np.empty(H.shape[0], dtype=np.uint8)
        fn = getattr(self.lib, f"oracle_degenerate_{'f32' if dt == np.float32 else 'f64'}")
        fn(_p(H), _p(out), C.c_int64(H.shape[0]), C.c_int(int(normalized)))
        return out

    def synth_quads(self, begin: int, count: int, seed: int, dist: int, dtype=np.float32):
        dt = np.dtype(dtype)
        src = np.empty((count, 8), dtype=dt)
        tar = np.empty((count, 8), dtype=dt)
        fn = getattr(self.lib, f"oracle_synth_quads_{'f32' if dt == np.float32 else 'f64'}")
        fn(_p(src), _p(tar), C.c_int64(begin), C.c_int64(count), C.c_uint64(seed), C.c_int(dist))
        return src, tar

    def ransac_refit(self, corr, mask, H_in):
        corr = _c(corr, np.float32)
        P, n_pts, _ = corr.shape
        mask = _c(mask, np.uint8).reshape(P, n_pts)
        H_in = _c(H_in, np.float32).reshape(P, 9)
        out = np.empty((P, 9), dtype=np.float32)
        used = np.empty(P, dtype=np.uint32)
        self.lib.oracle_ransac_refit_f32(_p(corr), C.c_int64(P), C.c_int32(n_pts), _p(mask), _p(H_in),
                                         _p(out), _p(used))
        return out, used

    def warp_grid(self, H, gw: int, gh: int, x0=0.0, y0=0.0, dx=1.0, dy=1.0) -> np.ndarray:
        H = _c(H, np.float32).reshape(-1, 9)
        out = np.empty((H.shape[0], gh, gw, 2), dtype=np.float32)
        self.lib.oracle_warp_grid_f32(_p(H), C.c_int64(H.shape[0]), C.c_float(x0), C.c_float(y0),
                                      C.c_float(dx), C.c_float(dy), C.c_int32(gw), C.c_int32(gh), _p(out))
        return out

    def curand_mrg32k3a(self, n: int, seed: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint32)
        self.lib.oracle_curand_mrg32k3a_u32(_p(out), C.c_int64(n), C.c_uint64(seed))
        return out

    def rng_u64(self, seed: int, ctr: int, lane: int) -> int:
        return int(self.lib.oracle_rng_u64(seed, ctr, lane))

    def ransac(self, corr, n_hyp: int, seed: int, thr2: float, samples=None, hyp_begin: int = 0,
               hyp_count: int | None = None, want_counts: bool = False, pair_begin: int = 0):
        """corr [P, n_pts, 4] float32 -> best_key [P] uint64 (+ counts [P, hyp_count]).
        pair_begin: global id of corr[0] when corr is a shard of the pairs (keys the sampler)."""
        corr = _c(corr, np.float32)
        P, n_pts, _ = corr.shape
        hyp_count = n_hyp - hyp_begin if hyp_count is None else hyp_count
        keys = np.zeros(P, dtype=np.uint64)
        counts = np.zeros((P, hyp_count), dtype=np.uint32) if want_counts else None
        sp = None
        if samples is not None:
            samples = _c(samples, np.uint32).reshape(P, n_hyp, 4)
            sp = _p(samples)
        self.lib.oracle_ransac_aca_shard_f32(_p(corr), C.c_int64(pair_begin), C.c_int64(P), C.c_int32(n_pts), sp,
                                       C.c_uint32(hyp_begin), C.c_uint32(hyp_count),
                                       C.c_uint32(n_hyp), C.c_uint64(seed), C.c_float(thr2),
                                       _p(keys), _p(counts) if want_counts else None)
        return (keys, counts) if want_counts else keys

    def ransac_sample(self, seed: int, pair: int, hyp: int, n_pts: int) -> np.ndarray:
        idx = np.zeros(4, dtype=np.uint32)
        self.lib.oracle_ransac_sample(C.c_uint64(seed), C.c_int64(pair), C.c_uint32(hyp),
                                      C.c_int32(n_pts), _p(idx))
        return idx

    def ransac_hypothesis(self, corr_pair, idx) -> np.ndarray:
        corr_pair = _c(corr_pair, np.float32)
        idx = _c(idx, np.uint32)
        H = np.empty(9, dtype=np.float32)
        self.lib.oracle_ransac_hypothesis_f32(_p(corr_pair), _p(idx), _p(H))
        return H

    def ransac_count(self, H, corr_pair, thr2: float) -> int:
        H = _c(H, np.float32)
        corr_pair = _c(corr_pair, np.float32)
        return int(self.lib.oracle_ransac_count_f32(_p(H), _p(corr_pair),
                                                    C.c_int32(corr_pair.shape[0]),
                                                    C.c_float(thr2)))


class RefLib:
    """The reference's own C++ (bit-exact ground truth for SKS / ACA)."""

    def __init__(self, o3: bool = False):
        """o3=True: the -O3 -march=x86-64-v3 -ffp-contract=fast build (speed comparator only:
        FMA contraction changes bits)."""
        build()
        so = REF_O3_SO if o3 else REF_SO
        if not os.path.exists(so):
            raise FileNotFoundError(
                f"{so} missing: it is built from /root/reference by `make -C oracle ref` in "
                "the authoring container and travels to the GPU box as a prebuilt file")
        self.lib = C.CDLL(so)
        self.lib.ref_hardware_threads.restype = C.c_int

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO) or os.path.isdir(REFERENCE_ROOT)

    def hardware_threads(self) -> int:
        return int(self.lib.ref_hardware_threads())

    def solve(self, solver: str, src, tar, threads: int = 1, out=None) -> np.ndarray:
        """solver in {aca, sks} (fp32/fp64) or ge (fp32 only, MOD/GE.cpp); always h33-normalised
        (MOD/ACA_SKS.cpp:94-98)."""
        dt = np.asarray(src).dtype
        src, tar = _c(src, dt).reshape(-1, 8), _c(tar, dt).reshape(-1, 8)
        n = src.shape[0]
        H = np.empty((n, 9), dtype=dt) if out is None else out
        fn = getattr(self.lib, f"ref_{solver}_{'f32' if dt == np.float32 else 'f64'}")
        fn(_p(src), _p(tar), _p(H), C.c_int64(n), C.c_int(threads))
        return H


class RefGpuLib:
    """The reference's own CUDA kernels (GPU.cu:81-240 cal_Homo_ACA/SKS, :359-507
    cal_Homo_GE, extracted and compiled for sm_100a at build time by oracle/Makefile)
    behind their reference launch shape <<<ceil(N/32),32>>>: fp64, SoA, un-normalised.
    Default build (FMA contraction on, as the reference ships): perf comparator only.
    ``nofma=True`` loads the -fmad=false build, whose rounding is that of the
    reference's C++: a GPU-side parity pin for our SoA fp64 kernels."""

    def __init__(self, nofma: bool = False):
        so = REFGPU_NOFMA_SO if nofma else REFGPU_SO
        if not os.path.exists(so):
            build()
        if not os.path.exists(so):
            raise FileNotFoundError(so)
        self.lib = C.CDLL(so)
        for name in ("refgpu_aca_f64", "refgpu_sks_f64", "refgpu_ge_f64", "refgpu_gpt_f64"):
            fn = getattr(self.lib, name)
            fn.restype = C.c_int
            fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]

    @staticmethod
    def available(nofma: bool = False) -> bool:
        return os.path.exists(REFGPU_NOFMA_SO if nofma else REFGPU_SO) or os.path.isdir(REFERENCE_ROOT)

    def run(self, solver: str, d_src: int, d_tar: int, d_H: int, n: int, stream: int) -> None:
        rc = getattr(self.lib, f"refgpu_{solver}_f64")(d_src, d_tar, d_H, n, stream)
        if rc != 0:
            raise RuntimeError(f"reference CUDA kernel launch failed: cudaError {rc}")


def ref_torch_funcs():
    """The reference's own torch statements (PY.py getInput / getTar / adjust, the timed bodies of
    TensorACA_rect :296-302 and ACA_vanilla :322-381), staged into oracle/_ref at build time.
    Returns the module, or raises FileNotFoundError where it was never staged."""
    import importlib.util
    if not os.path.exists(REF_TORCH_PY):
        raise FileNotFoundError(f"{REF_TORCH_PY} not staged (run `make -C oracle stage` where /root/reference exists)")
    spec = importlib.util.spec_from_file_location("ref_torch_funcs", REF_TORCH_PY)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
