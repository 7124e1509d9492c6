"""The C++ drop-in header (include/sks_homography.hpp): a reference-style caller
compiles against it with plain g++ and links libsks_cuda.so."""
import os
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, sks):
    exe = str(tmp_path / "dropin")
    libdir = os.path.dirname(sks.path)
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "dropin_main.cpp"), "-o", exe,
                    "-L", libdir, "-lsks_cuda", f"-Wl,-rpath,{libdir}"], check=True)
    return exe


@pytest.mark.skipif(torch.cuda.is_available(), reason="no-GPU behaviour")
def test_cpp_caller_links_and_fails_loudly_without_gpu(tmp_path, sks):
    res = subprocess.run([_build(tmp_path, sks)], capture_output=True, text=True)
    assert "status -3" in res.stdout and res.returncode == 3      # SKS_ERR_NO_DEVICE, no fallback


@pytest.mark.gpu
def test_cpp_caller_reproduces_reference_kat(tmp_path, sks, oracle):
    res = subprocess.run([_build(tmp_path, sks)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    lines = dict(l.split(None, 1) for l in res.stdout.strip().splitlines() if l[:3] in ("ACA", "SKS"))
    # fp32 strings are the reference's own bit patterns
    aca = np.array(lines["ACA"].split(), dtype=np.float32)
    sks32 = np.array(lines["SKS"].split(), dtype=np.float32)
    assert np.array_equal(aca, np.array([1.72610629, 0.000940763159, 482, 1.04164827, 1.05091727,
                                         378.571411, 0.00147034752, -0.00092884578, 1], np.float32))
    assert np.array_equal(sks32, np.array([1.7261076, 0.00094215438, 482.000092, 1.0416491,
                                           1.05091894, 378.571411, 0.0014703495, -0.000928843336, 1],
                                          np.float32))
    src = np.array([[0, 0, 200, 0, 50, 139, 181, 93]], np.float64)
    tar = np.array([[482, 378.571428571429, 639.240222867399, 453.531347049346, 601.89683773457,
                     610.680390715948, 673.458433551996, 563.547292039412]])
    assert np.array_equal(np.array(lines["ACA64"].split(), dtype=np.float64), oracle.solve("aca", src, tar)[0])
    assert np.array_equal(np.array(lines["SKS64"].split(), dtype=np.float64), oracle.solve("sks", src, tar)[0])
    assert "batch status 0" in res.stdout
