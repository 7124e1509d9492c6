"""GPU tests of the in-library multi-GPU RANSAC driver (csrc/multi.cu) and of the multi-rank
paths (one process per GPU over NCCL, launched here with torchrun).  Everything that needs
more than one GPU skips on a single-GPU box; the ngpu = 1 forms always run."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_GPU = torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.fixture
def api(sks, cuda):
    from sks_homography_b200 import api as a
    return a


def test_multi_entry_with_one_gpu_equals_the_two_step_api(api, oracle, cuda):
    corr = api.synth_corr(5, 1001, seed=3, device=cuda)
    H, cnt, mask, keys = api.ransac_multi(corr, 777, 9, 2.25, ngpu=1, want_mask=True)
    want = api.ransac_keys(corr, 777, 9, 2.25)
    assert torch.equal(keys, want)
    assert np.array_equal(keys.cpu().numpy().view(np.uint64), oracle.ransac(corr.cpu().numpy(), 777, 9, 2.25))
    H2, cnt2, mask2 = api.ransac_finalize(corr, 777, 9, 2.25, want, want_mask=True)
    assert torch.equal(H.view(torch.int32), H2.view(torch.int32)) and torch.equal(cnt, cnt2) and torch.equal(mask, mask2)


@pytest.mark.skipif(N_GPU < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("ngpu", [2, 0])
def test_multi_gpu_device_entry_is_bit_identical(api, oracle, cuda, ngpu):
    """sks_cuda_ransac_aca_multi_f32: peers read the matches over NVLink peer access, winners merged
    by peer atomics; keys, models, counts and masks equal the single-GPU run bit for bit."""
    for n_pts, n_hyp in ((4096, 4096), (1001, 777), (9001, 300)):
        corr = api.synth_corr(7, n_pts, seed=5, device=cuda)
        H1, c1, m1, k1 = api.ransac_multi(corr, n_hyp, 11, 2.25, ngpu=1, want_mask=True)
        for _ in range(2):                               # second call: contexts and events reused
            Hn, cn, mn, kn = api.ransac_multi(corr, n_hyp, 11, 2.25, ngpu=ngpu, want_mask=True)
            assert torch.equal(kn, k1) and torch.equal(cn, c1) and torch.equal(mn, m1)
            assert torch.equal(Hn.view(torch.int32), H1.view(torch.int32))
    rng = np.random.default_rng(0)                       # explicit sample list, read over NVLink as well
    samples = torch.from_numpy(rng.integers(0, 2**31, size=(7, 300, 4), dtype=np.int64).astype(np.int32)).to(cuda)
    _, _, _, k1 = api.ransac_multi(corr, 300, 0, 2.25, ngpu=1, samples=samples)
    _, _, _, kn = api.ransac_multi(corr, 300, 0, 2.25, ngpu=ngpu, samples=samples)
    assert torch.equal(k1, kn)
    want = oracle.ransac(corr.cpu().numpy(), 300, 0, 2.25, samples=samples.cpu().numpy().view(np.uint32))
    assert np.array_equal(kn.cpu().numpy().view(np.uint64), want)


@pytest.mark.skipif(N_GPU < 2, reason="needs >= 2 GPUs")
def test_multi_gpu_host_entry_and_device_count_switch(api, sks, cuda):
    corr = api.synth_corr(9, 2048, seed=6, device=cuda).cpu()
    H1, c1, m1, k1 = api.ransac_host(corr, 2048, 4, 2.25, want_mask=True)
    Hn, cn, mn, kn = api.ransac_host(corr, 2048, 4, 2.25, want_mask=True, ngpu=0)
    assert torch.equal(k1, kn) and torch.equal(c1, cn) and torch.equal(m1, mn)
    assert torch.equal(H1.view(torch.int32), Hn.view(torch.int32))
    try:
        assert sks.c.sks_host_set_device_count(0) == 0   # the unchanged entry point now spans the box
        Hs, cs, _, ks = api.ransac_host(corr, 2048, 4, 2.25)
    finally:
        sks.c.sks_host_set_device_count(1)
    assert torch.equal(ks, k1) and torch.equal(Hs.view(torch.int32), H1.view(torch.int32))


def test_cpp_caller_multi_gpu_equals_single_gpu(tmp_path, sks):
    """g++-compiled caller, no Python on the path: all GPUs of the box vs one GPU (on a single-GPU
    box the multi entry degenerates to one device and must still agree)."""
    exe = str(tmp_path / "ransac_multi")
    libdir = os.path.dirname(sks.path)
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "ransac_multi_main.cpp"), "-o", exe,
                    "-L", libdir, "-lsks_cuda", f"-Wl,-rpath,{libdir}"], check=True)
    res = subprocess.run([exe, "48", "2048", "8192"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "identical 1" in res.stdout, res.stdout + res.stderr


@pytest.mark.skipif(N_GPU < 2, reason="needs >= 2 GPUs")
def test_multi_rank_nccl_and_peer_reduce_under_torchrun(tmp_path):
    """One process per GPU (the bench's layout): hypothesis-sharded RANSAC merged by the NCCL max
    all-reduce and by the hand-written NVLink PeerReducer, and pair-sharded RANSAC, all equal to
    the unsharded single-GPU keys (tests/tools/multi_rank_check.py)."""
    n = min(N_GPU, 4)
    env = dict(os.environ, PYTHONPATH=ROOT)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                          "--master-addr", "127.0.0.1", "--master-port", "29611",
                          os.path.join(ROOT, "tests", "tools", "multi_rank_check.py")],
                         capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0 and "MULTI_RANK_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.skipif(N_GPU < 2, reason="needs >= 2 GPUs")
def test_tensors_on_a_non_current_device(api, oracle):
    """A tensor on cuda:1 while cuda:0 is current: the call switches device for the launch (and the
    stream handle it passes is that device's current stream) and switches back."""
    assert torch.cuda.current_device() == 0
    d1 = torch.device("cuda", 1)
    s, t = oracle.synth_quads(0, 5000, 4, 1, np.float32)
    H = api.solve("aca", torch.from_numpy(s).to(d1), torch.from_numpy(t).to(d1))
    assert H.device == d1 and torch.cuda.current_device() == 0
    want = oracle.solve("aca", s, t)
    got = H.cpu().numpy()
    assert np.array_equal(np.nan_to_num(got).view(np.uint32), np.nan_to_num(want).view(np.uint32))
    corr = api.synth_corr(3, 600, seed=2, device=d1)
    keys = api.ransac_keys(corr, 256, 3, 2.25)
    assert keys.device == d1 and torch.cuda.current_device() == 0
    assert np.array_equal(keys.cpu().numpy().view(np.uint64), oracle.ransac(corr.cpu().numpy(), 256, 3, 2.25))


def test_multi_entry_edge_cases(api, sks, oracle, cuda):
    """ngpu larger than the box or than the hypothesis count is clamped, an empty batch is a no-op,
    the finalize step is optional, invalid arguments are rejected -- on any number of GPUs."""
    corr = api.synth_corr(3, 300, seed=4, device=cuda)
    want = oracle.ransac(corr.cpu().numpy(), 5, 2, 2.25)
    for ngpu in (0, 1, 99):
        H, cnt, mask, keys = api.ransac_multi(corr, 5, 2, 2.25, ngpu=ngpu)
        assert np.array_equal(keys.cpu().numpy().view(np.uint64), want)
    H, cnt, mask, keys = api.ransac_multi(corr, 1, 2, 2.25, ngpu=0)                 # one hypothesis: one device
    assert np.array_equal(keys.cpu().numpy().view(np.uint64), oracle.ransac(corr.cpu().numpy(), 1, 2, 2.25))
    H, cnt, mask, keys = api.ransac_multi(corr, 5, 2, 2.25, ngpu=0, finalize=False)
    assert H is None and cnt is None and np.array_equal(keys.cpu().numpy().view(np.uint64), want)
    st = torch.cuda.current_stream(cuda).cuda_stream
    k = torch.zeros(3, dtype=torch.int64, device=cuda)
    f = sks.c.sks_cuda_ransac_aca_multi_f32
    assert f(corr.data_ptr(), 0, 300, None, 5, 2, 2.25, 0, k.data_ptr(), None, None, None, st) == 0      # empty batch
    assert f(corr.data_ptr(), 3, 300, None, 0, 2, 2.25, 0, k.data_ptr(), None, None, None, st) == -1     # no hypotheses
    assert f(corr.data_ptr(), 3, 300, None, 5, 2, 2.25, -1, k.data_ptr(), None, None, None, st) == -1    # ngpu < 0
    assert f(None, 3, 300, None, 5, 2, 2.25, 0, k.data_ptr(), None, None, None, st) == -1
    h = sks.c.sks_host_ransac_aca_multi_f32
    Hh = torch.empty((3, 9))
    assert h(corr.cpu().data_ptr(), 3, 300, None, 0, 2, 2.25, 0, Hh.data_ptr(), None, None, None) == -1
