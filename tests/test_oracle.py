"""CPU tests of the oracle itself: pinned against the reference's golden vectors
(tests/golden/, produced by the reference's own C++ / torch code), against the
live reference library when it is present, and against the veri_4Pts.m KATs."""
import numpy as np
import pytest

from util import assert_same_bits, reproject_error, same_bits

DT = {"f32": np.float32, "f64": np.float64}


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("case", ["d0", "d1", "d2", "deg"])
@pytest.mark.parametrize("solver", ["aca", "sks"])
def test_oracle_matches_reference_golden(oracle, golden, tag, case, solver):
    g = golden["ref_general"]
    H = oracle.solve(solver, g[f"src_{tag}_{case}"], g[f"tar_{tag}_{case}"], normalize=True)
    assert_same_bits(H, g[f"{solver}_{tag}_{case}"], f"{solver} {tag} {case}")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("dist", [0, 1, 2])
def test_oracle_matches_live_reference(oracle, reflib, dtype, dist):
    s, t = oracle.synth_quads(12345, 200_000, 7 + dist, dist, dtype)
    for solver in ("aca", "sks"):
        assert_same_bits(oracle.solve(solver, s, t), reflib.solve(solver, s, t),
                         f"{solver} {dtype.__name__} dist{dist}")


@pytest.mark.parametrize("case", ["d0", "d1", "deg", "kat"])
def test_ge_oracle_matches_reference_golden(oracle, golden, case):
    """Competitor solver RHO-GE against outputs of the reference's MOD/GE.cpp."""
    g = golden["ref_ge"]
    assert_same_bits(oracle.solve("ge", g[f"src_f32_{case}"], g[f"tar_f32_{case}"]), g[f"ge_f32_{case}"],
                     f"ge f32 {case}")


@pytest.mark.parametrize("dist", [0, 1, 2])
def test_ge_oracle_matches_live_reference(oracle, reflib, dist):
    s, t = oracle.synth_quads(31 * dist, 200_000, 5 + dist, dist, np.float32)
    assert_same_bits(oracle.solve("ge", s, t), reflib.solve("ge", s, t), f"ge dist{dist}")


def test_ge_fp64_agrees_with_aca_where_it_is_defined(oracle):
    """fp64 GE is the same type-generic elimination; on general quads it must agree with
    ACA to the accuracy the paper reports for GE (no pivoting: ~1e-7 relative at worst),
    and it has no answer for an axis-aligned source square (zero first pivot)."""
    s, t = oracle.synth_quads(0, 50_000, 3, 1, np.float64)
    ge, aca = oracle.solve("ge", s, t), oracle.solve("aca", s, t)
    scale = np.abs(aca).max(axis=1, keepdims=True)
    assert (np.abs(ge - aca) / scale).max() < 1e-9
    assert reproject_error(ge, s, t).max() < 1e-6
    sq = np.array([[10, 10, 138, 10, 10, 138, 138, 138]], dtype=np.float64)
    assert not np.isfinite(oracle.solve("ge", sq, sq + 3.0)[0, :8]).all()


def test_reference_threads_agree(reflib, oracle):
    s, t = oracle.synth_quads(0, 50_001, 3, 1, np.float32)
    a = reflib.solve("sks", s, t, threads=1)
    b = reflib.solve("sks", s, t, threads=4)
    assert_same_bits(a, b, "threaded reference loop")


def test_kat_general_quad(oracle, golden):
    """ML/veri_4Pts.m:9-12,70-77: both solvers recover H_real."""
    k = golden["kat_veri4pts"]
    Hn = (k["H_real"] / k["H_real"][2, 2]).ravel()
    s, t = k["src_general"][None], k["tar_general"][None]
    aca = oracle.solve("aca", s, t)[0]
    sks = oracle.solve("sks", s, t)[0]
    # h12 is ~1e-3 of the matrix scale, so its own relative error is ~1e-11
    assert np.abs(aca / Hn - 1).max() < 5e-11 and np.abs(aca - Hn).max() / np.abs(Hn).max() < 1e-13
    assert np.abs(sks / Hn - 1).max() < 5e-11 and np.abs(sks - Hn).max() / np.abs(Hn).max() < 1e-13
    # the reference's fp32 bit patterns listed in SURVEY.md A.2
    aca32 = oracle.solve("aca", s.astype(np.float32), t.astype(np.float32))[0]
    sks32 = oracle.solve("sks", s.astype(np.float32), t.astype(np.float32))[0]
    want_aca = np.array([1.72610629, 0.000940763159, 482, 1.04164827, 1.05091727, 378.571411,
                         0.00147034752, -0.00092884578, 1], dtype=np.float32)
    want_sks = np.array([1.7261076, 0.00094215438, 482.000092, 1.0416491, 1.05091894, 378.571411,
                         0.0014703495, -0.000928843336, 1], dtype=np.float32)
    assert np.array_equal(aca32, want_aca)
    assert np.array_equal(sks32, want_sks)


def test_kat_rectangle(oracle, golden):
    """ML/veri_4Pts.m:82-95 through ML/ACA_rect.m."""
    k = golden["kat_veri4pts"]
    mx, my, w, ratio = k["rect"]
    t = k["tar_rect"][None]
    H = oracle.aca_rect(t, mx, my, w, ratio, normalize=False)[0]
    want = np.array([212953.589614272, 115.924439385708, 59465407.8712903, 128510.470731783,
                     129653.851160308, 46705195.8680852, 181.400062388642, -114.593978674425,
                     123372.215500602])
    assert np.abs(H / want - 1).max() < 1e-9
    Hn = oracle.aca_rect(t, mx, my, w, ratio, normalize=True)[0]
    ref = (k["H_real"] / k["H_real"][2, 2]).ravel()
    assert np.abs(Hn / ref - 1).max() < 1e-10
    assert Hn[8] == 1.0


def test_rect_and_vanilla_match_reference_torch(oracle, golden):
    """Golden H obtained by executing PY.py:296-302 / :322-381 on torch-CPU."""
    g = golden["ref_torch"]
    tarq = g["tar_new"][:, :2, :].transpose(0, 2, 1).reshape(-1, 8)
    H = oracle.aca_rect(tarq, 0, 0, g["scale"].item(), g["div"].item(), M=g["src_new"][:, :2, 0],
                        normalize=False)
    assert_same_bits(H.reshape(-1, 3, 3), g["H_rect"], "TensorACA_rect")
    Hv = oracle.solve("aca", g["src"].reshape(-1, 8), g["tar"].reshape(-1, 8), normalize=False)
    assert_same_bits(Hv.reshape(-1, 3, 3), g["H_vanilla"], "ACA_vanilla")


@pytest.mark.parametrize("dtype,tol", [(np.float32, 2e-3), (np.float64, 1e-11)])
def test_rect_equals_general_aca_on_rectangle_corners(oracle, dtype, tol):
    """ACA-rect == runKernel_ACA fed the rectangle's corners, after
    normalisation."""
    _, t = oracle.synth_quads(0, 20_000, 5, 0, dtype)
    mx, my, w, h = 36.0, 81.0, 50.0, 40.0
    s = np.tile(np.array([mx, my, mx + w, my, mx, my + h, mx + w, my + h], dtype=dtype), (len(t), 1))
    a = oracle.solve("aca", s, t)
    r = oracle.aca_rect(t, mx, my, w, w / h)
    scale = np.abs(a).max(axis=1, keepdims=True)
    err = (np.abs(a - r) / scale).max(axis=1)
    assert err.max() < tol and np.median(err) < tol * 1e-2


def test_degenerate_flags_follow_nonfinite_output(oracle, golden):
    g = golden["ref_general"]
    for tag in ("f32", "f64"):
        for solver in ("aca", "sks"):
            H = g[f"{solver}_{tag}_deg"]
            flag = oracle.degenerate(H, normalized=True)
            assert np.array_equal(flag.astype(bool), ~np.isfinite(H[:, :8]).all(axis=1))
            assert flag[0] == 1 and flag[-1] == 0
    up = oracle.solve("aca", g["src_f32_deg"], g["tar_f32_deg"], normalize=False)
    f2 = oracle.degenerate(up, normalized=False)
    assert f2[0] == 1 and f2[-1] == 0     # un-normalised collinear quad is all zeros -> h33 == 0


def test_fp64_solvers_agree_and_reproject(oracle):
    s, t = oracle.synth_quads(0, 100_000, 21, 1, np.float64)
    a, k = oracle.solve("aca", s, t), oracle.solve("sks", s, t)
    rel = np.abs(a - k).max(1) / np.abs(a).max(1)
    assert np.median(rel) < 1e-13 and np.percentile(rel, 99) < 1e-10
    assert np.percentile(reproject_error(a, s, t), 99) < 1e-9
    assert np.percentile(reproject_error(k, s, t), 99) < 1e-9


def test_cv2_cross_check(oracle):
    """BASELINE config 4 accuracy tier: cv::getPerspectiveTransform (OpenCV is not
    vendored by the reference; python cv2 here) agrees up to its own conditioning."""
    cv2 = pytest.importorskip("cv2")
    s, t = oracle.synth_quads(0, 2048, 31, 1, np.float64)
    k = oracle.solve("sks", s, t)
    worst = 0.0
    for i in range(len(s)):
        G = cv2.getPerspectiveTransform(s[i].reshape(4, 2).astype(np.float32),
                                        t[i].reshape(4, 2).astype(np.float32))
        worst = max(worst, reproject_error(G.reshape(1, 9), s[i].astype(np.float32),
                                           t[i].astype(np.float32))[0])
    assert worst < 1e-3           # cv2's LU on raw pixel coordinates
    assert reproject_error(k, s, t).max() < 1e-6


def test_curand_mrg32k3a_restatement_matches_library_output(oracle, golden):
    """The reference's sample list (GPU.cu:1443-1446) comes from cuRAND's host-API MRG32K3A
    generator; the golden slices are libcurand 10.3 output on a B200 (tools/curand_dump.py),
    including the 81920-subsequence wrap-around of its output order."""
    g, T = golden["curand_mrg32k3a"], 81920
    for name, seed in zip(("s11", "s0", "sbig"), g["seeds"]):
        a = oracle.curand_mrg32k3a(2 * T + 1024, int(seed))
        assert np.array_equal(a[:4096], g[f"{name}_head"])
        assert np.array_equal(a[T - 16:T + 4096], g[f"{name}_wrap"])
        assert np.array_equal(a[2 * T - 16:2 * T + 1024], g[f"{name}_wrap2"])
    assert np.array_equal(oracle.curand_mrg32k3a(5, 11), g["s11_head"][:5])       # n < 81920
    assert oracle.curand_mrg32k3a(0, 11).size == 0


def test_refit_port_recovers_a_planted_model(oracle):
    """Post-RANSAC least-squares refit (our definition): noisy inliers of a known homography."""
    rng = np.random.default_rng(0)
    P, n = 3, 600
    Ht = np.array([[1.1, 0.05, 12], [-0.03, 0.95, -7], [1e-4, -2e-4, 1]])
    x = rng.uniform(10, 600, (P, n, 2))
    q = np.concatenate([x, np.ones((P, n, 1))], -1) @ Ht.T
    X = q[..., :2] / q[..., 2:] + rng.normal(0, 0.3, (P, n, 2))
    corr = np.concatenate([x, X], -1).astype(np.float32)
    mask = np.ones((P, n), np.uint8)
    mask[:, ::7] = 0
    corr[:, ::7, 2:] += 50                                   # gross outliers, masked out
    H0 = np.tile(np.eye(3, dtype=np.float32).ravel(), (P, 1))
    H, used = oracle.ransac_refit(corr, mask, H0)
    assert (used == mask.sum(1)).all()
    assert reproject_error(H, np.tile([10, 10, 600, 10, 10, 600, 600, 600], (P, 1)),
                           np.tile((np.array([[10, 10, 1], [600, 10, 1], [10, 600, 1], [600, 600, 1]]) @ Ht.T
                                    / (np.array([[10, 10, 1], [600, 10, 1], [10, 600, 1], [600, 600, 1]]) @ Ht.T)[:, 2:]
                                    )[:, :2].reshape(8), (P, 1))).max() < 0.2
    few = np.zeros_like(mask)
    few[:, :3] = 1                                           # fewer than 4 inliers: keep the input model
    H2, used2 = oracle.ransac_refit(corr, few, H0)
    assert (used2 == 0).all() and np.array_equal(H2, H0)


def test_warp_grid_port_maps_rectangle_corners_to_targets(oracle):
    _, t = oracle.synth_quads(0, 64, 9, 0, np.float32)
    H = oracle.aca_rect(t, 15.0, 12.0, 128.0, 1.0, normalize=False)
    g = oracle.warp_grid(H, 2, 2, x0=15.0, y0=12.0, dx=128.0, dy=128.0)
    assert np.abs(g.reshape(64, 8) - t).max() < 2e-3          # TL,TR,BL,BR order, fp32 solve
    Hn = oracle.aca_rect(t, 15.0, 12.0, 128.0, 1.0, normalize=True)
    assert np.abs(oracle.warp_grid(Hn, 2, 2, 15.0, 12.0, 128.0, 128.0) - g).max() < 2e-3   # scale-free


def test_synth_is_counter_based(oracle):
    a_s, a_t = oracle.synth_quads(100, 50, 11, 1, np.float32)
    b_s, b_t = oracle.synth_quads(0, 200, 11, 1, np.float32)
    assert np.array_equal(a_s, b_s[100:150]) and np.array_equal(a_t, b_t[100:150])
    c_s, _ = oracle.synth_quads(0, 200, 12, 1, np.float32)
    assert not np.array_equal(c_s, b_s)
    s, t = oracle.synth_quads(0, 4096, 11, 2, np.float32)
    assert ((s[:, 0] >= 10) & (s[:, 0] < 30)).all() and np.array_equal(s[:, 2], s[:, 0] + 128)
    off = t - s
    assert (off == np.floor(off)).all() and off.min() >= 0 and off.max() <= 31


def _scene(rng, n_pts, inlier_frac=0.6):
    """One RANSAC scene with a planted homography."""
    Hs = np.array([[1.1, 0.05, 6.0], [-0.04, 0.95, 3.0], [2e-4, -1e-4, 1.0]])
    xy = rng.uniform(10, 158, size=(n_pts, 2))
    p = np.c_[xy, np.ones(n_pts)] @ Hs.T
    XY = p[:, :2] / p[:, 2:3] + rng.uniform(-0.4, 0.4, size=(n_pts, 2))
    out = rng.random(n_pts) > inlier_frac
    XY[out] = rng.uniform(10, 190, size=(out.sum(), 2))
    return np.c_[xy, XY].astype(np.float32), ~out


def test_ransac_oracle_finds_planted_model(oracle):
    rng = np.random.default_rng(1)
    corr = np.stack([_scene(rng, 512)[0] for _ in range(3)])
    keys, counts = oracle.ransac(corr, 256, seed=5, thr2=4.0, want_counts=True)
    cnt, hyp = keys >> np.uint64(32), np.uint64(0xFFFFFFFF) - (keys & np.uint64(0xFFFFFFFF))
    assert (cnt > 0.5 * 512).all()
    for p in range(3):
        assert counts[p].max() == cnt[p]
        assert int(np.argmax(counts[p])) == int(hyp[p])          # lowest id wins ties
        idx = oracle.ransac_sample(5, p, int(hyp[p]), 512)
        H = oracle.ransac_hypothesis(corr[p], idx)
        assert oracle.ransac_count(H, corr[p], 4.0) == cnt[p]


def test_ransac_oracle_ranges_merge_by_max(oracle):
    rng = np.random.default_rng(2)
    corr = np.stack([_scene(rng, 256)[0] for _ in range(2)])
    full = oracle.ransac(corr, 300, seed=9, thr2=9.0)
    a = oracle.ransac(corr, 300, seed=9, thr2=9.0, hyp_begin=0, hyp_count=130)
    b = oracle.ransac(corr, 300, seed=9, thr2=9.0, hyp_begin=130, hyp_count=170)
    assert np.array_equal(np.maximum(a, b), full)
    # explicit sample list == counter RNG when the list holds the same indices
    samples = np.array([[oracle.ransac_sample(9, p, j, 256) for j in range(300)] for p in range(2)],
                       dtype=np.uint32)
    assert np.array_equal(oracle.ransac(corr, 300, seed=0, thr2=9.0, samples=samples), full)
