"""CPU tests of the drop-in boundary: libsks_cuda.so builds with nvcc alone, loads,
exports every symbol include/sks_cuda.h declares, validates arguments, and fails
loudly (never silently falls back) when no GPU is present."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

import sks_homography_b200 as pkg
from sks_homography_b200 import _lib


def test_library_exports_every_declared_symbol(sks):
    declared = pkg.declared_symbols()
    assert len(declared) >= 30
    raw = C.CDLL(sks.path)
    missing = [n for n in declared if not hasattr(raw, n)]
    assert not missing, missing
    assert set(_lib._SIGS) == set(declared)          # the binding covers the whole ABI
    assert sks.c.sks_cuda_abi_version() == 2


def test_library_is_sm100a_with_bulk_copy_kernels(sks):
    """The shipped object carries sm_100a SASS with UBLKCP (cp.async.bulk) in the
    ring kernels -- evidence for the TMA-staged AoS path without needing a GPU."""
    exe = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-sass", sks.path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UBLKCP" in out and "SYNCS" in out
    assert "HMMA" not in out and "UTCHMMA" not in out       # deliberately no tensor cores


def test_shard_range_is_a_partition(sks):
    for n in (0, 1, 7, 64, 1000003, 2**28):
        for w in (1, 2, 3, 4, 8):
            pos = 0
            for r in range(w):
                b, c = sks.shard_range(n, r, w)
                assert b == pos and c in (n // w, n // w + 1)
                pos += c
            assert pos == n
    b, c = C.c_int64(), C.c_int64()
    assert sks.c.sks_cuda_shard_range(10, 3, 3, C.byref(b), C.byref(c)) == _lib.ERR_INVALID_ARG


def test_error_strings(sks):
    assert sks.error_string(0) == "success"
    assert "CPU fallback" in sks.error_string(_lib.ERR_NO_DEVICE)
    assert "aligned" in sks.error_string(_lib.ERR_UNALIGNED)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_device_is_a_loud_error(sks):
    a = np.zeros(32, np.float32)
    p = a.ctypes.data
    assert sks.c.sks_cuda_aca_f32(p, p, p, 1, 0, 0, 1, None, None) == _lib.ERR_NO_DEVICE
    assert sks.c.sks_host_sks_f32(p, p, p, 1, 1) == _lib.ERR_NO_DEVICE
    from sks_homography_b200 import api
    with pytest.raises(pkg.SksCudaError):
        api.runKernel_ACA(torch.zeros(1, 8), torch.zeros(1, 8))


def test_argument_validation_precedes_device_use(sks):
    a = np.zeros(32, np.float32)
    p = a.ctypes.data
    assert sks.c.sks_cuda_aca_f32(p, p, p, -1, 0, 0, 1, None, None) == _lib.ERR_INVALID_ARG
    assert sks.c.sks_cuda_aca_f32(p, p, p, 1, 7, 0, 1, None, None) == _lib.ERR_INVALID_ARG
    assert sks.c.sks_cuda_aca_f32(p, p, p, 1, 0, 0, 99, None, None) == _lib.ERR_INVALID_ARG
    assert sks.c.sks_cuda_aca_f32(None, p, p, 1, 0, 0, 1, None, None) == _lib.ERR_INVALID_ARG
    assert sks.c.sks_cuda_set_variant(5) == _lib.ERR_INVALID_ARG
    assert sks.c.sks_cuda_set_tuning(0, 1, 0) == _lib.ERR_INVALID_ARG
    assert sks.c.sks_cuda_set_tuning(0, 4, 0) == 0


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    root = os.path.dirname(os.path.abspath(pkg.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "libsks_oracle" not in text and "libsks_ref" not in text, f


def test_only_tests_bench_and_smoke_touch_the_oracle():
    """oracle/ is the checker: besides tests/, bench.py (cpu_baseline / reference arm) and
    __graft_entry__ (build + smoke), no Python file of the repository imports it."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    offenders = []
    for d, _, files in os.walk(root):
        rel = os.path.relpath(d, root)
        if rel.split(os.sep)[0] in ("tests", "oracle", ".git", "gpurun_out", "baseline", "sks-homography_b200"):
            continue
        for f in files:
            if f.endswith(".py") and not (rel == "." and f in ("bench.py", "__graft_entry__.py")):
                text = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, re.M):
                    offenders.append(os.path.join(rel, f))
    assert not offenders, offenders


def test_tuning_knobs_are_per_thread(sks):
    """sks_cuda_set_variant / _tuning act on the calling host thread only (no process-global state)."""
    import threading
    seen = {}
    assert sks.c.sks_cuda_set_variant(2) == 0
    try:
        def other():
            seen["fresh"] = sks.c.sks_cuda_get_variant()
            sks.c.sks_cuda_set_variant(3)
            seen["own"] = sks.c.sks_cuda_get_variant()
        t = threading.Thread(target=other)
        t.start(); t.join()
        assert seen == {"fresh": 0, "own": 3}
        assert sks.c.sks_cuda_get_variant() == 2
    finally:
        sks.c.sks_cuda_set_variant(0)


def test_ransac_launch_plan_fills_whole_waves(sks):
    """BASELINE configs[4] on 1/2/4/8 ranks (148 SMs x 3 resident CTAs): the plan halves the rounds
    per CTA exactly where a partial last wave would cost more than the extra tile loads."""
    plan = sks.c.sks_cuda_ransac_chunk_plan
    slots, round_ = 148 * 3, 256 * 2
    want_rounds = {1: 8, 2: 4, 4: 2, 8: 1}
    for world, rounds in want_rounds.items():
        chunk = plan(1024, 65536 // world, 256, 2, 8, slots)
        assert chunk == rounds * round_, (world, chunk)
        ctas = -(-(65536 // world) // chunk) * 1024
        waves = ctas / slots
        assert -(-ctas // slots) / waves < 1.01               # < 1 % lost to the partial last wave
    # few CTAs: one wave whatever the split -> the finest one (most parallelism)
    assert plan(4, 2048, 256, 2, 8, slots) == round_
    # always a multiple of a round, never zero, covers odd counts
    for hyp in (1, 511, 513, 70001):
        c = plan(3, hyp, 384, 4, 3, slots)
        assert c > 0 and c % (384 * 4) == 0


def test_sample_list_validation_needs_no_gpu(sks):
    """api.ransac_keys rejects a mis-typed / mis-sized explicit sample list before anything reaches CUDA."""
    import torch
    from sks_homography_b200 import api
    corr = torch.zeros((2, 50, 4))
    with pytest.raises(TypeError):
        api.ransac_keys(corr, 8, 1, 2.25, samples=torch.zeros((2, 8, 4), dtype=torch.int64))
    with pytest.raises(ValueError):
        api.ransac_keys(corr, 8, 1, 2.25, samples=torch.zeros((2, 7, 4), dtype=torch.int32))
