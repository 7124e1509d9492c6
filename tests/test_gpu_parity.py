"""GPU parity tests (run on the B200 with -m gpu): every streaming kernel, called
through the C ABI, against the CPU oracle on identical inputs.  The bar is
BIT-EXACT (the tolerances of BASELINE.json -- 1e-5 fp32 / 1e-12 fp64 relative,
1e-4 px reprojection -- are then met with error zero); non-finite outputs must be
non-finite in the same places (x86 and the GPU differ only in NaN payload)."""
import ctypes as C

import numpy as np
import pytest
import torch

from util import assert_same_bits, reproject_error

pytestmark = pytest.mark.gpu

TDT = {np.float32: torch.float32, np.float64: torch.float64}
VARIANTS = {"direct": (1, 0, 4, 0), "direct_16B": (1, 2, 4, 0), "ring": (2, 0, 4, 0),
            "ring_small_3": (2, 1, 3, 0), "ring_1cta_2": (2, 0, 2, 1), "wring": (3, 0, 4, 0)}


@pytest.fixture
def api(sks, cuda):
    from sks_homography_b200 import api as a
    yield a
    sks.c.sks_cuda_set_variant(0)
    sks.c.sks_cuda_set_tuning(0, 4, 0)


def set_variant(sks, name):
    v, small, stages, ctas = VARIANTS[name]
    assert sks.c.sks_cuda_set_variant(v) == 0
    assert sks.c.sks_cuda_set_tuning(small, stages, ctas) == 0


def dev(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("solver", ["aca", "sks"])
def test_aos_bit_exact(api, sks, oracle, cuda, solver, dtype, variant):
    set_variant(sks, variant)
    n = (1 << 20) + 37                      # ragged: last tile partial and not 16-byte sized
    for dist, normalize in ((0, True), (1, True), (1, False)):
        s, t = oracle.synth_quads(5 * n, n, 11 + dist, dist, dtype)
        flag = torch.full((n,), 7, dtype=torch.uint8, device=cuda)
        H = api.solve(solver, dev(s, cuda), dev(t, cuda), normalize=normalize, degenerate=flag)
        want = oracle.solve(solver, s, t, normalize=normalize)
        assert_same_bits(H.cpu().numpy(), want, f"{solver} {dtype.__name__} {variant} d{dist}")
        assert np.array_equal(flag.cpu().numpy(), oracle.degenerate(want, normalize))


@pytest.mark.parametrize("variant", ["direct", "direct_16B", "ring", "wring"])
@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 31, 63, 64, 65, 255, 256, 257, 511, 1000, 4099])
def test_ragged_sizes(api, sks, oracle, cuda, n, variant):
    set_variant(sks, variant)
    for dtype in (np.float32, np.float64):
        s, t = oracle.synth_quads(0, n, 3, 1, dtype)
        for solver in ("aca", "sks"):
            guard = torch.full((n * 9 + 64,), 123.0, dtype=TDT[dtype], device=cuda)
            out = guard[: n * 9].view(n, 9)
            api.solve(solver, dev(s, cuda).view(n, 8), dev(t, cuda).view(n, 8), result=out)
            assert_same_bits(out.cpu().numpy(), oracle.solve(solver, s, t).reshape(n, 9), f"n={n}")
            assert (guard[n * 9:] == 123.0).all(), "wrote past the end of H"


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("solver", ["aca", "sks"])
def test_soa_bit_exact(api, oracle, cuda, solver, dtype):
    """The reference GPU layout (GPU.cu:87-95,141-149), un-normalised like the
    reference kernels and normalised; vector path, scalar path (odd n) and tails."""
    for n in (1 << 18, (1 << 18) + 3, 5):
        s, t = oracle.synth_quads(77, n, 13, 1, dtype)
        for normalize in (False, True):
            H = api.solve(solver, dev(s.T, cuda), dev(t.T, cuda), normalize=normalize, layout="soa")
            want = oracle.solve(solver, s, t, normalize=normalize)
            assert_same_bits(H.cpu().numpy().T, want, f"soa {solver} n={n}")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("variant", ["direct", "direct_16B", "ring", "wring"])
def test_ge_competitor_bit_exact(api, sks, oracle, golden, cuda, dtype, variant):
    """RHO-GE (MOD/GE.cpp) through the same streaming kernels:
    AoS and SoA, ragged sizes, the reference's golden vectors, host-pointer path."""
    set_variant(sks, variant)
    for n, dist in (((1 << 19) + 5, 1), (4099, 0), (3, 1), (0, 1)):
        s, t = oracle.synth_quads(17, n, 29 + dist, dist, dtype)
        want = oracle.solve("ge", s, t).reshape(n, 9)
        flag = torch.full((n,), 7, dtype=torch.uint8, device=cuda)
        H = api.solve("ge", dev(s, cuda).view(n, 8), dev(t, cuda).view(n, 8), degenerate=flag)
        assert_same_bits(H.cpu().numpy(), want, f"ge aos {dtype.__name__} n={n}")
        assert np.array_equal(flag.cpu().numpy(), oracle.degenerate(want, True))
        if n:
            Hs = api.solve("ge", dev(s.T, cuda), dev(t.T, cuda), normalize=False, layout="soa")
            assert_same_bits(Hs.cpu().numpy().T, want, f"ge soa {dtype.__name__} n={n}")
    if dtype == np.float32:
        g = golden["ref_ge"]
        for case in ("d0", "d1", "deg", "kat"):
            H = api.runKernel_GE(dev(g[f"src_f32_{case}"], cuda), dev(g[f"tar_f32_{case}"], cuda))
            assert_same_bits(H.cpu().numpy(), g[f"ge_f32_{case}"], f"ge golden {case}")
        s, t = oracle.synth_quads(1, 300_001, 4, 1, np.float32)
        H = api.runKernel_GE(torch.from_numpy(s), torch.from_numpy(t))          # host pointers
        assert_same_bits(H.numpy(), oracle.solve("ge", s, t), "ge host path")


@pytest.mark.parametrize("variant", ["direct", "ring"])
def test_gpt_lu_competitor_bit_exact(api, sks, oracle, cuda, variant):
    """GPT-LU (fp64, the arithmetic of cal_Homo_GPT): AoS / SoA against the CPU restatement,
    including quadruples whose pivot search has to swap rows and exact degenerate ones."""
    set_variant(sks, variant)
    for n, dist in (((1 << 17) + 5, 1), (4099, 0), (2, 1), (0, 1)):
        s, t = oracle.synth_quads(23, n, 31 + dist, dist, np.float64)
        want = oracle.solve("gpt", s, t).reshape(n, 9)
        H = api.solve("gpt", dev(s, cuda).view(n, 8), dev(t, cuda).view(n, 8))
        assert_same_bits(H.cpu().numpy(), want, f"gpt aos n={n}")
        if n:
            Hs = api.solve("gpt", dev(s.T, cuda), dev(t.T, cuda), normalize=False, layout="soa")
            assert_same_bits(Hs.cpu().numpy().T, want, f"gpt soa n={n}")
    deg = np.array([[0, 0, 1, 0, 2, 0, 1, 1], [0, 0, 4, 0, 0, 3, 5, 4], [2, 2, 2, 2, 2, 2, 2, 2]], np.float64)
    H = api.solve("gpt", dev(deg, cuda), dev(deg[::-1].copy(), cuda))
    assert_same_bits(H.cpu().numpy(), oracle.solve("gpt", deg, deg[::-1].copy()), "gpt degenerate")
    s, t = oracle.synth_quads(0, 20_000, 7, 1, np.float64)
    g = api.solve("gpt", dev(s, cuda), dev(t, cuda)).cpu().numpy()
    assert reproject_error(g, s, t).max() < 1e-6


@pytest.mark.parametrize("solver", ["aca", "sks", "ge", "gpt"])
def test_soa_fp64_equals_reference_cuda_kernels_without_fma(api, cuda, solver):
    """GPU-side pin of rows a5/a6 (and fp64 GE): the reference's own kernels
    cal_Homo_ACA / cal_Homo_SKS / cal_Homo_GE / cal_Homo_GPT (GPU.cu:81-240, :359-507, :242-357) compiled with
    -fmad=false round every operation on its own, like the reference's C++; launched as
    the reference launches them (SoA fp64, block 32, un-normalised) they must produce
    the same bits as our SoA kernels on the same device buffers."""
    from oracle.oracle import RefGpuLib
    if not RefGpuLib.available(nofma=True):
        pytest.skip("oracle/_ref/libsks_refgpu_nofma.so not built")
    ref = RefGpuLib(nofma=True)
    for n, dist in (((1 << 20) + 3, 1), (100_000, 0)):
        src, tar = api.synth_quads(n, 41, dist, torch.float64, cuda, layout="soa")
        H_ref = torch.empty((9, n), dtype=torch.float64, device=cuda)
        ref.run(solver, src.data_ptr(), tar.data_ptr(), H_ref.data_ptr(), n,
                torch.cuda.current_stream().cuda_stream)
        H = api.solve(solver, src, tar, normalize=False, layout="soa")
        torch.cuda.synchronize()
        assert_same_bits(H.cpu().numpy(), H_ref.cpu().numpy(), f"{solver} vs reference CUDA kernel, dist {dist}")


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("solver", ["aca", "sks"])
def test_reference_golden_vectors(api, sks, golden, cuda, solver, tag):
    """Directly against outputs of the reference's own C++ (tests/golden)."""
    g = golden["ref_general"]
    for variant in ("direct", "ring"):
        set_variant(sks, variant)
        for case in ("d0", "d1", "d2", "deg"):
            s, t = g[f"src_{tag}_{case}"], g[f"tar_{tag}_{case}"]
            H = api.solve(solver, dev(s, cuda), dev(t, cuda))
            assert_same_bits(H.cpu().numpy(), g[f"{solver}_{tag}_{case}"], f"{solver} {tag} {case}")


def test_degenerate_flags_identical(api, oracle, golden, cuda):
    g = golden["ref_general"]
    for tag, tdt in (("f32", torch.float32), ("f64", torch.float64)):
        s, t = g[f"src_{tag}_deg"], g[f"tar_{tag}_deg"]
        for solver in ("aca", "sks"):
            for normalize in (True, False):
                flag = torch.zeros(len(s), dtype=torch.uint8, device=cuda)
                H = api.solve(solver, dev(s, cuda), dev(t, cuda), normalize=normalize, degenerate=flag)
                want = oracle.solve(solver, s, t, normalize=normalize)
                assert np.array_equal(np.isfinite(H.cpu().numpy()), np.isfinite(want))
                assert np.array_equal(flag.cpu().numpy(), oracle.degenerate(want, normalize))
    # reference-normalised degenerate rows are [nan x8, 1] with return code 0
    H = api.runKernel_ACA_double(dev(g["src_f64_deg"][:1], cuda), dev(g["tar_f64_deg"][:1], cuda))
    row = H.cpu().numpy()[0]
    assert np.isnan(row[:8]).all() and row[8] == 1.0


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("variant", ["direct", "ring", "wring"])
def test_aca_rect_bit_exact(api, sks, oracle, cuda, dtype, variant):
    set_variant(sks, variant)
    n = (1 << 19) + 5
    _, t = oracle.synth_quads(0, n, 17, 0, dtype)
    rngM = np.random.default_rng(4).integers(10, 30, size=(n, 2)).astype(dtype)
    for normalize in (True, False):
        H = api.aca_rect(dev(t, cuda), 128.0, 1.0, 36.0, 81.0, normalize=normalize)
        assert_same_bits(H.cpu().numpy(), oracle.aca_rect(t, 36.0, 81.0, 128.0, 1.0, normalize=normalize),
                         "rect shared corner")
        H = api.aca_rect(dev(t, cuda), 50.0, 1.25, M=dev(rngM, cuda), normalize=normalize)
        assert_same_bits(H.cpu().numpy(), oracle.aca_rect(t, 0, 0, 50.0, 1.25, M=rngM, normalize=normalize),
                         "rect per-sample corner")
    # SoA
    H = api.aca_rect(dev(t.T, cuda), 50.0, 1.25, M=dev(rngM.T, cuda), layout="soa")
    assert_same_bits(H.cpu().numpy().T, oracle.aca_rect(t, 0, 0, 50.0, 1.25, M=rngM), "rect soa")


def test_reference_torch_interfaces(api, golden, cuda):
    """TensorACA_rect / ACA_vanilla with the reference's tensor conventions against H
    obtained by executing the reference's torch statements (tests/golden/ref_torch)."""
    g = golden["ref_torch"]
    bs = g["src"].shape[0]
    H = api.TensorACA_rect(bs, dev(g["src_new"], cuda), dev(g["tar_new"], cuda),
                           g["scale"].item(), g["div"].item())
    assert_same_bits(H.cpu().numpy(), g["H_rect"], "TensorACA_rect")
    # the reference's own calling convention: scale / div as one-element DEVICE tensors (PY.py:33-35),
    # read by the kernel -- no host synchronisation
    for dt in (np.float32, np.float64):
        Hd = api.TensorACA_rect(bs, dev(g["src_new"].astype(dt), cuda), dev(g["tar_new"].astype(dt), cuda),
                                dev(g["scale"].astype(dt), cuda), dev(g["div"].astype(dt), cuda))
        Hs = api.TensorACA_rect(bs, dev(g["src_new"].astype(dt), cuda), dev(g["tar_new"].astype(dt), cuda),
                                float(g["scale"].astype(dt).item()), float(g["div"].astype(dt).item()))
        assert_same_bits(Hd.cpu().numpy(), Hs.cpu().numpy(), f"TensorACA_rect device scalars {dt.__name__}")
        if dt == np.float32:
            assert_same_bits(Hd.cpu().numpy(), g["H_rect"], "TensorACA_rect device scalars")
    Hv = api.ACA_vanilla(bs, dev(g["src"], cuda), dev(g["tar"], cuda))
    assert_same_bits(Hv.cpu().numpy(), g["H_vanilla"], "ACA_vanilla")


def test_matlab_signature_kat(api, golden, cuda):
    k = golden["kat_veri4pts"]
    mx, my, w, ratio = k["rect"]
    T = np.vstack([k["tar_rect"].reshape(4, 2).T, np.ones(4)])
    H = api.ACA_rect(dev(T, cuda), mx, my, w, ratio).cpu().numpy()
    ref = k["H_real"] / k["H_real"][2, 2]
    assert np.abs(H / ref - 1).max() < 1e-10
    s, t = dev(k["src_general"][None], cuda), dev(k["tar_general"][None], cuda)
    for fn in (api.runKernel_ACA_double, api.runKernel_SKS_double):
        Hn = fn(s, t).cpu().numpy().reshape(3, 3)
        assert np.abs(Hn - ref).max() / np.abs(ref).max() < 1e-13


def test_host_pointer_entry_points(api, oracle, cuda):
    """sks_host_*: host buffers in and out (pageable and pinned), chunked pipeline."""
    n = (1 << 21) + 11                       # > one 2^20-quad chunk: exercises the ring
    for dtype in (np.float32, np.float64):
        s, t = oracle.synth_quads(9, n, 23, 1, dtype)
        want = oracle.solve("aca", s, t)
        H = api.solve("aca", torch.from_numpy(s), torch.from_numpy(t))          # pageable
        assert not H.is_cuda
        assert_same_bits(H.numpy(), want, "host pageable")
        sp, tp = torch.from_numpy(s).pin_memory(), torch.from_numpy(t).pin_memory()
        out = torch.empty((n, 9), dtype=TDT[dtype]).pin_memory()
        api.solve("sks", sp, tp, result=out)
        assert_same_bits(out.numpy(), oracle.solve("sks", s, t), "host pinned")
    _, t = oracle.synth_quads(0, 70_001, 2, 0, np.float32)
    H = api.aca_rect(torch.from_numpy(t), 128.0, 1.0, 15.0, 12.0)
    assert_same_bits(H.numpy(), oracle.aca_rect(t, 15.0, 12.0, 128.0, 1.0), "host rect")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("dist", [0, 1, 2])
def test_device_generator_matches_oracle(api, oracle, cuda, dtype, dist):
    n = 100_003
    s, t = api.synth_quads(n, seed=11, dist=dist, dtype=TDT[dtype], device=cuda, begin=12345)
    ws, wt = oracle.synth_quads(12345, n, 11, dist, dtype)
    assert_same_bits(s.cpu().numpy(), ws, "src")
    assert_same_bits(t.cpu().numpy(), wt, "tar")
    s2, t2 = api.synth_quads(n, seed=11, dist=dist, dtype=TDT[dtype], device=cuda, begin=12345,
                             layout="soa")
    assert torch.equal(s2.T.contiguous(), s) and torch.equal(t2.T.contiguous(), t)


def test_gather_samples_matches_reference_semantics(api, cuda):
    """get_rand_list (GPU.cu:52-78): index = r % size, repeats allowed, SoA output."""
    rng = np.random.default_rng(0)
    pool = rng.uniform(0, 800, size=(2540, 4))
    n = 5000
    rand4 = rng.integers(0, 2**32, size=(4, n), dtype=np.uint32)
    src, tar = api.gather_samples(dev(pool, cuda), n, rand4=dev(rand4, cuda), layout="soa")
    idx = rand4 % 2540
    want_src = np.stack([pool[idx[k], c] for k in range(4) for c in (0, 1)])
    want_tar = np.stack([pool[idx[k], c] for k in range(4) for c in (2, 3)])
    assert np.array_equal(src.cpu().numpy(), want_src) and np.array_equal(tar.cpu().numpy(), want_tar)
    s1, t1 = api.gather_samples(dev(pool.astype(np.float32), cuda), n, seed=5)
    s2, t2 = api.gather_samples(dev(pool.astype(np.float32), cuda), n, seed=5)
    assert torch.equal(s1, s2) and s1.shape == (n, 8)


@pytest.mark.parametrize("gw,gh", [(32, 32), (33, 5), (1, 1), (128, 128)])
def test_warp_grid_consumer(api, oracle, cuda, gw, gh):
    """Sampling grid after the solver: from stored H and fused with
    ACA-rect (H never written); bit-exact against the CPU port, even and odd grid sizes."""
    n = 257
    _, t = oracle.synth_quads(3, n, 5, 0, np.float32)
    H = oracle.aca_rect(t, 15.0, 12.0, 128.0, 1.0, normalize=False)
    spec = dict(x0=15.0, y0=12.0, dx=128.0 / max(gw - 1, 1), dy=128.0 / max(gh - 1, 1))
    want = oracle.warp_grid(H, gw, gh, **spec)
    got = api.warp_grid(dev(H, cuda), gw, gh, **spec)
    assert_same_bits(got.cpu().numpy(), want, "warp_grid from H")
    guard = torch.full((n * gh * gw * 2 + 64,), 7.0, dtype=torch.float32, device=cuda)
    fused = api.aca_rect_warp_grid(dev(t, cuda), 128.0, 1.0, gw, gh, M_x=15.0, M_y=12.0,
                                   out=guard[: n * gh * gw * 2].view(n, gh, gw, 2), **spec)
    assert_same_bits(fused.cpu().numpy(), want, "fused ACA-rect + warp_grid")
    assert (guard[n * gh * gw * 2:] == 7.0).all()
    if gw > 1 and gh > 1:          # the grid's corners are the rectangle's corners: they land on tar
        c = fused.cpu().numpy()[:, [0, 0, -1, -1], [0, -1, 0, -1], :].reshape(n, 8)
        assert np.abs(c - t).max() < 2e-3


def test_curand_mrg32k3a_stream(api, oracle, golden, cuda):
    """Hand-written MRG32k3a kernel == cuRAND's host-API generator (the reference's sample
    list, GPU.cu:1443-1446): against the CPU restatement, the committed library output and,
    when libcurand is present on the box, the live library."""
    g = golden["curand_mrg32k3a"]
    for n in (0, 1, 5, 81919, 81920, 81921, 4 * (1 << 20) + 3):
        got = api.curand_mrg32k3a(n, 11, cuda).cpu().numpy().view(np.uint32)
        assert np.array_equal(got, oracle.curand_mrg32k3a(n, 11)), n
    for name, seed in zip(("s11", "s0", "sbig"), g["seeds"]):
        got = api.curand_mrg32k3a(200_000, int(seed), cuda).cpu().numpy().view(np.uint32)
        assert np.array_equal(got[:4096], g[f"{name}_head"])
        assert np.array_equal(got[81920 - 16:81920 + 4096], g[f"{name}_wrap"])
    try:
        lib = C.CDLL("libcurand.so.10")
    except OSError:
        return
    n = 1_000_003
    gen = C.c_void_p()
    assert lib.curandCreateGenerator(C.byref(gen), 121) == 0           # CURAND_RNG_PSEUDO_MRG32K3A
    assert lib.curandSetPseudoRandomGeneratorSeed(gen, C.c_ulonglong(11)) == 0
    buf = torch.empty(n, dtype=torch.int32, device=cuda)
    assert lib.curandGenerate(gen, C.c_void_p(buf.data_ptr()), C.c_size_t(n)) == 0
    torch.cuda.synchronize()
    assert torch.equal(buf, api.curand_mrg32k3a(n, 11, cuda)), "differs from live libcurand"
    lib.curandDestroyGenerator(gen)


def test_reference_gpu_flow_replayed(api, oracle, cuda):
    """The reference's whole GPU flow (GPU.cu:1443-1464) on its own fixture: cuRAND
    MRG32K3A(seed 11) sample list -> get_rand_list -> cal_Homo_ACA, here as list kernel +
    fused gather-solve, against the CPU restatement of each step (synthetic 2540-match pool:
    the fixture file itself does not travel to the GPU box)."""
    rng = np.random.default_rng(4)           # a 2540-match pool like the reference's fixture
    pool = (rng.random((2540, 4)) * 700 + 20).astype(np.float64)
    n = 100_000
    r = oracle.curand_mrg32k3a(4 * n, 11).reshape(4, n)
    idx = (r % pool.shape[0]).T                                # [n,4], as GPU.cu:55-58
    s = pool[idx][:, :, :2].reshape(n, 8)
    t = pool[idx][:, :, 2:].reshape(n, 8)
    want = oracle.solve("aca", s, t, normalize=False)
    rand4 = api.curand_mrg32k3a(4 * n, 11, cuda).view(4, n)
    H = api.gather_solve("aca", dev(pool, cuda), n, rand4=rand4, normalize=False, layout="soa")
    assert_same_bits(H.cpu().numpy().T, want, "replayed reference flow")


def test_fp64_accuracy_tier(api, oracle, cuda):
    """BASELINE config 4 tier on the GPU results: SKS64 == ACA64 to ~1e-12 and
    reprojection far below 1e-4 px."""
    n = 1 << 16
    s, t = oracle.synth_quads(0, n, 41, 1, np.float64)
    a = api.runKernel_ACA_double(dev(s, cuda), dev(t, cuda)).cpu().numpy()
    k = api.runKernel_SKS_double(dev(s, cuda), dev(t, cuda)).cpu().numpy()
    rel = np.abs(a - k).max(1) / np.abs(a).max(1)
    assert np.percentile(rel, 99) < 1e-10
    assert np.percentile(reproject_error(k, s, t), 99) < 1e-9
    assert reproject_error(k, s, t).max() < 1e-4


def test_16_byte_aligned_buffers_take_the_narrow_path(api, oracle, cuda):
    """Base pointers that are 16- but not 32-byte aligned cannot use LDG.256."""
    n = 5000
    s, t = oracle.synth_quads(0, n, 8, 1, np.float32)
    pad = torch.zeros(n * 8 + 4, dtype=torch.float32, device=cuda)
    pad2 = torch.zeros(n * 8 + 4, dtype=torch.float32, device=cuda)
    out = torch.zeros(n * 9 + 4, dtype=torch.float32, device=cuda)
    sv, tv, ov = pad[4:].view(n, 8), pad2[4:].view(n, 8), out[4:].view(n, 9)
    assert sv.data_ptr() % 32 == 16
    sv.copy_(torch.from_numpy(s)); tv.copy_(torch.from_numpy(t))
    api.solve("aca", sv, tv, result=ov)
    assert_same_bits(ov.cpu().numpy(), oracle.solve("aca", s, t), "16-byte aligned views")
    assert float(out[:4].abs().sum()) == 0.0


def test_launch_is_cuda_graph_capturable(api, oracle, cuda):
    """The device-pointer entry points only enqueue (no sync, no allocation), so a
    launch-bound loop of small solves can be captured once and replayed."""
    n = 777
    s, t = oracle.synth_quads(0, n, 19, 1, np.float32)
    ds, dt_ = dev(s, cuda), dev(t, cuda)
    outs = [torch.zeros((n, 9), dtype=torch.float32, device=cuda) for _ in range(4)]
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        api.solve("aca", ds, dt_, result=outs[0])          # warm-up outside capture
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for k, solver in enumerate(("aca", "sks", "aca", "sks")):
                api.solve(solver, ds, dt_, result=outs[k], normalize=(k < 2))
        for o in outs:
            o.zero_()
        g.replay()
        torch.cuda.synchronize()
    assert_same_bits(outs[0].cpu().numpy(), oracle.solve("aca", s, t), "graph aca")
    assert_same_bits(outs[1].cpu().numpy(), oracle.solve("sks", s, t), "graph sks")
    assert_same_bits(outs[3].cpu().numpy(), oracle.solve("sks", s, t, normalize=False), "graph sks raw")


def test_unaligned_pointer_is_rejected(api, sks, cuda):
    buf = torch.zeros(64, dtype=torch.float32, device=cuda)
    p = buf.data_ptr()
    st = sks.c.sks_cuda_aca_f32(p + 4, p, p, 1, 0, 0, 1, None, None)
    assert st == -2


def test_in_process_multi_gpu_host_driver(api, sks, oracle, cuda):
    """sks_host_set_device_count: a host-pointer batch sharded over every visible GPU
    by the library itself gives the same bytes as one GPU."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (covered by the gloo test and tools/gpu_multi.sh otherwise)")
    n = (1 << 21) + 77
    s, t = oracle.synth_quads(0, n, 29, 1, np.float32)
    want = oracle.solve("sks", s, t)
    try:
        assert sks.c.sks_host_set_device_count(0) == 0          # all visible GPUs
        H = api.solve("sks", torch.from_numpy(s).pin_memory(), torch.from_numpy(t).pin_memory())
        assert_same_bits(H.numpy(), want, "multi-GPU host path, pinned")
        H = api.solve("sks", torch.from_numpy(s), torch.from_numpy(t))
        assert_same_bits(H.numpy(), want, "multi-GPU host path, pageable")
    finally:
        sks.c.sks_host_set_device_count(1)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_rect_planar34_layout(api, sks, oracle, cuda, dtype):
    """sks_cuda_aca_rect_planar_*: the reference's [n,3,4] homogeneous tensors read in
    place (PY.py:24-37) == the interleaved path, with per-sample and shared corners."""
    for n in (1, 255, 100_003):
        _, t = oracle.synth_quads(0, n, 6, 0, dtype)
        M = np.random.default_rng(n).integers(10, 30, size=(n, 2)).astype(dtype)
        tar34 = np.ones((n, 3, 4), dtype=dtype)
        tar34[:, :2, :] = t.reshape(n, 4, 2).transpose(0, 2, 1)
        src34 = np.ones((n, 3, 4), dtype=dtype)
        src34[:, 0, :] = M[:, :1] + np.array([0, 128, 0, 128], dtype=dtype)
        src34[:, 1, :] = M[:, 1:] + np.array([0, 0, 128, 128], dtype=dtype)
        d_t, d_s = dev(tar34, cuda), dev(src34, cuda)
        name = "f32" if dtype == np.float32 else "f64"
        fn = getattr(sks.c, f"sks_cuda_aca_rect_planar_{name}")
        for normalize in (0, 1):
            H = torch.full((n * 9 + 16,), 5.0, dtype=TDT[dtype], device=cuda)
            flag = torch.zeros(n, dtype=torch.uint8, device=cuda)
            sks.check(fn(d_t.data_ptr(), d_s.data_ptr(), 0.0, 0.0, 128.0, 1.0, H.data_ptr(), n, normalize,
                         flag.data_ptr(), None), "planar per-sample")
            want = oracle.aca_rect(t, 0, 0, 128.0, 1.0, M=M, normalize=bool(normalize))
            assert_same_bits(H[: n * 9].view(n, 9).cpu().numpy(), want, f"planar per-sample n={n}")
            assert bool((H[n * 9:] == 5.0).all()) and int(flag.sum()) == 0
            sks.check(fn(d_t.data_ptr(), None, 36.0, 81.0, 50.0, 1.25, H.data_ptr(), n, normalize, None, None),
                      "planar shared")
            want = oracle.aca_rect(t, 36.0, 81.0, 50.0, 1.25, normalize=bool(normalize))
            assert_same_bits(H[: n * 9].view(n, 9).cpu().numpy(), want, f"planar shared n={n}")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("solver", ["aca", "sks"])
def test_fused_gather_solve_equals_gather_then_solve(api, oracle, cuda, solver, dtype):
    """One kernel for the reference's get_rand_list -> cal_Homo_* flow (GPU.cu:1449-1464)."""
    rng = np.random.default_rng(5)
    pool = rng.uniform(7, 790, size=(2540, 4)).astype(dtype)
    n = 100_003
    rand4 = rng.integers(0, 2**32, size=(4, n), dtype=np.uint32)
    d_pool, d_rand = dev(pool, cuda), dev(rand4, cuda)
    for r4 in (d_rand, None):
        src, tar = api.gather_samples(d_pool, n, seed=9, rand4=r4)
        want = oracle.solve(solver, src.cpu().numpy(), tar.cpu().numpy())
        H = api.gather_solve(solver, d_pool, n, seed=9, rand4=r4)
        assert_same_bits(H.cpu().numpy(), want, "fused aos")
        Hs = api.gather_solve(solver, d_pool, n, seed=9, rand4=r4, normalize=False, layout="soa")
        assert_same_bits(Hs.cpu().numpy().T, oracle.solve(solver, src.cpu().numpy(), tar.cpu().numpy(),
                                                            normalize=False), "fused soa")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_fused_gather_solve_shared_memory_pool(api, sks, oracle, cuda, dtype):
    """Large SoA batches take the persistent kernel that keeps the match pool in shared memory;
    it must agree with the L1-path kernel (forced by variant 2) and with gather + oracle."""
    rng = np.random.default_rng(6)
    pool = rng.uniform(7, 790, size=(2540, 4)).astype(dtype)
    n = 700_001
    d_pool = dev(pool, cuda)
    rand4 = api.curand_mrg32k3a(4 * n, 11, cuda).view(4, n)
    for r4 in (rand4, None):
        H = api.gather_solve("aca", d_pool, n, seed=9, rand4=r4, normalize=False, layout="soa")
        sks.c.sks_cuda_set_variant(2)
        H_l1 = api.gather_solve("aca", d_pool, n, seed=9, rand4=r4, normalize=False, layout="soa")
        sks.c.sks_cuda_set_variant(0)
        assert torch.equal(H.view(torch.int32 if dtype == np.float32 else torch.int64),
                           H_l1.view(torch.int32 if dtype == np.float32 else torch.int64))
        src, tar = api.gather_samples(d_pool, n, seed=9, rand4=r4)
        want = oracle.solve("aca", src[:50_000].cpu().numpy(), tar[:50_000].cpu().numpy(), normalize=False)
        assert_same_bits(H[:, :50_000].cpu().numpy().T, want, "shared-memory pool gather")
    big = dev(rng.uniform(7, 790, size=(20_000, 4)).astype(dtype), cuda)      # too large for shared memory
    Hb = api.gather_solve("sks", big, n, seed=2, normalize=True, layout="soa")
    src, tar = api.gather_samples(big, n, seed=2)
    assert_same_bits(Hb[:, -30_000:].cpu().numpy().T,
                     oracle.solve("sks", src[-30_000:].cpu().numpy(), tar[-30_000:].cpu().numpy()), "L1 fallback")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("solver", ["aca", "sks"])
def test_extreme_magnitudes_overflow_and_denormals(api, oracle, cuda, solver, dtype):
    """Un-normalised entries are degree 7-9 polynomials of the coordinates (README.md:54):
    huge inputs overflow to inf/NaN, tiny ones run through denormals.  IEEE arithmetic
    without flush-to-zero must reproduce the reference bit for bit in both regimes."""
    n = 20_000
    s, t = oracle.synth_quads(0, n, 3, 1, dtype)
    scales = [1e-12, 1e-7, 1e-4, 1e4, 1e7, 1e12] if dtype == np.float32 else [1e-60, 1e-35, 1e30, 1e45]
    for sc in scales:
        ss, tt = (s * dtype(sc)).astype(dtype), (t * dtype(sc)).astype(dtype)
        for normalize in (True, False):
            want = oracle.solve(solver, ss, tt, normalize=normalize)
            flag = torch.zeros(n, dtype=torch.uint8, device=cuda)
            H = api.solve(solver, dev(ss, cuda), dev(tt, cuda), normalize=normalize, degenerate=flag)
            assert_same_bits(H.cpu().numpy(), want, f"{solver} scale {sc} normalize={normalize}")
            assert np.array_equal(flag.cpu().numpy(), oracle.degenerate(want, normalize))
    # mixed: only the target plane is tiny (denormal products against normal ones)
    tt = (t * dtype(1e-30 if dtype == np.float32 else 1e-250)).astype(dtype)
    assert_same_bits(api.solve(solver, dev(s, cuda), dev(tt, cuda)).cpu().numpy(), oracle.solve(solver, s, tt),
                     "tiny target plane")


def test_concurrent_host_calls_from_two_threads(api, oracle, cuda):
    """Two host threads call the host-pointer entry points at the same time (streaming solver and
    fused RANSAC): the per-device contexts serialise them, results stay bit-exact."""
    import threading
    s, t = oracle.synth_quads(0, (1 << 20) + 5, 3, 1, np.float32)
    want = oracle.solve("aca", s, t)
    corr = api.synth_corr(6, 700, seed=2, device=cuda).cpu()
    wantk = oracle.ransac(corr.numpy(), 512, 5, 2.25)
    out, errs = {}, []

    def solver():
        try:
            for _ in range(3):
                out["H"] = api.runKernel_ACA(torch.from_numpy(s), torch.from_numpy(t))
        except Exception as e:
            errs.append(e)

    def ransac():
        try:
            for _ in range(3):
                out["k"] = api.ransac_host(corr, 512, 5, 2.25)[3]
        except Exception as e:
            errs.append(e)

    th = [threading.Thread(target=solver), threading.Thread(target=ransac)]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errs, errs
    assert_same_bits(out["H"].numpy(), want, "concurrent host solve")
    assert np.array_equal(out["k"].numpy().view(np.uint64), wantk)


def test_registering_a_caller_owned_buffer_takes_the_direct_dma_path(api, sks, oracle, cuda):
    """sks_host_register pins pageable caller memory in place; results are the same bits and the
    registered call must not be slower than the staged one."""
    import time
    n = (1 << 22) + 3
    s, t = oracle.synth_quads(0, n, 9, 1, np.float32)
    H = np.empty((n, 9), np.float32)
    ts, tt, tH = torch.from_numpy(s), torch.from_numpy(t), torch.from_numpy(H)

    def run():
        t0 = time.perf_counter()
        api.solve("aca", ts, tt, result=tH)
        return time.perf_counter() - t0
    run()
    staged = min(run() for _ in range(3))
    want = H.copy()
    bufs = [(s, s.nbytes), (t, t.nbytes), (H, H.nbytes)]
    try:
        for a, nb in bufs:
            assert sks.c.sks_host_register(a.ctypes.data, nb) == 0
            assert sks.c.sks_host_register(a.ctypes.data, nb) == 0        # idempotent
        H[:] = 0
        run()
        direct = min(run() for _ in range(3))
        assert_same_bits(H, want, "registered buffers")
        assert_same_bits(H, oracle.solve("aca", s, t), "registered buffers vs oracle")
        assert direct < staged * 1.5          # loose: shared hosts are noisy; typically 0.7x
    finally:
        for a, _ in bufs:
            sks.c.sks_host_unregister(a.ctypes.data)


def test_calls_are_cuda_graph_capturable(api, oracle, cuda):
    """The device-pointer entry points only enqueue on the current stream (no sync, no allocation in the
    library), so a launch-bound training loop can capture them in a CUDA graph and replay it:
    TensorACA_rect with device-resident scale / div and the general solver, replayed on new inputs."""
    bs = 256
    g = torch.Generator(device="cpu").manual_seed(3)
    def make():
        src = torch.randint(10, 30, (bs, 2), generator=g).float().unsqueeze(1).repeat(1, 4, 1)
        src[:, 1, 0] += 128; src[:, 2, 1] += 128; src[:, 3, 0] += 128; src[:, 3, 1] += 128
        tar = src + torch.randint(0, 32, (bs, 4, 2), generator=g).float()
        return src, tar
    def planar(a):
        return torch.cat((a.transpose(1, 2), torch.ones((bs, 1, 4))), dim=1).contiguous()
    src, tar = make()
    s34, t34 = planar(src).to(cuda), planar(tar).to(cuda)
    sq, tq = src.reshape(bs, 8).contiguous().to(cuda), tar.reshape(bs, 8).contiguous().to(cuda)
    scale = (s34[0, 0, 1:2] - s34[0, 0, 0:1]).clone()
    div = (scale / (s34[0, 1, 2:3] - s34[0, 1, 0:1])).clone()
    api.TensorACA_rect(bs, s34, t34, scale, div); api.solve("aca", sq, tq)      # warm-up outside the capture
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        H_rect = api.TensorACA_rect(bs, s34, t34, scale, div)
        H_aca = api.solve("aca", sq, tq)
    for _ in range(3):                                    # replay on fresh inputs copied into the static buffers
        src, tar = make()
        s34.copy_(planar(src)); t34.copy_(planar(tar)); sq.copy_(src.reshape(bs, 8)); tq.copy_(tar.reshape(bs, 8))
        graph.replay()
        torch.cuda.synchronize()
        want = oracle.solve("aca", src.reshape(bs, 8).numpy(), tar.reshape(bs, 8).numpy())
        assert_same_bits(H_aca.cpu().numpy(), want, "graph replay, general solver")
        eager = api.TensorACA_rect(bs, s34, t34, scale, div)
        assert torch.equal(H_rect.view(torch.int32), eager.view(torch.int32))
