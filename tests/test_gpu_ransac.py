"""GPU tests of the fused ACA-RANSAC kernel: inlier counts and winners bit-exact
against the CPU oracle for a fixed seed and for an explicit sample list."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture
def api(sks, cuda):
    from sks_homography_b200 import api as a
    return a


def u64(t):
    return t.cpu().numpy().view(np.uint64)


@pytest.mark.parametrize("n_pts,n_hyp", [(4096, 2048), (1000, 777), (64, 5), (9000, 600)])
def test_keys_match_oracle_seeded(api, oracle, cuda, n_pts, n_hyp):
    P = 6
    corr = api.synth_corr(P, n_pts, seed=3, inlier_permille=550, noise=0.5, device=cuda)
    keys = api.ransac_keys(corr, n_hyp, seed=77, thr2=2.25)
    want = oracle.ransac(corr.cpu().numpy(), n_hyp, seed=77, thr2=2.25)
    assert np.array_equal(u64(keys), want)
    cnt, hyp = api.decode_keys(keys)
    assert (cnt.cpu().numpy() > 0.3 * n_pts).all() or n_hyp < 50


def test_per_hypothesis_counts_bit_exact(api, oracle, cuda):
    """Every hypothesis, not just the winner: score single-hypothesis ranges."""
    P, n_pts, n_hyp = 2, 512, 96
    corr = api.synth_corr(P, n_pts, seed=9, device=cuda)
    _, counts = oracle.ransac(corr.cpu().numpy(), n_hyp, seed=5, thr2=4.0, want_counts=True)
    got = np.zeros_like(counts)
    for j in range(n_hyp):
        k = api.ransac_keys(corr, n_hyp, seed=5, thr2=4.0, hyp_begin=j, hyp_count=1)
        got[:, j] = (u64(k) >> np.uint64(32)).astype(np.uint32)
    assert np.array_equal(got, counts)


def test_explicit_sample_list(api, oracle, cuda):
    P, n_pts, n_hyp = 3, 2048, 1024
    corr = api.synth_corr(P, n_pts, seed=4, device=cuda)
    rng = np.random.default_rng(8)
    samples = rng.integers(0, n_pts, size=(P, n_hyp, 4), dtype=np.uint32)
    samples[:, 5] = samples[:, 5, :1]                     # a repeated-index (degenerate) sample
    keys = api.ransac_keys(corr, n_hyp, seed=0, thr2=4.0, samples=torch.from_numpy(samples).to(cuda))
    want = oracle.ransac(corr.cpu().numpy(), n_hyp, seed=0, thr2=4.0, samples=samples)
    assert np.array_equal(u64(keys), want)


def test_ranges_merge_and_finalize(api, oracle, cuda):
    P, n_pts, n_hyp = 5, 4096, 4096
    corr = api.synth_corr(P, n_pts, seed=6, device=cuda)
    full = api.ransac_keys(corr, n_hyp, seed=1, thr2=2.25)
    keys = torch.zeros(P, dtype=torch.int64, device=cuda)
    for b, c in ((0, 1000), (1000, 2000), (3000, 1096)):        # what 3 GPUs would each score
        api.ransac_keys(corr, n_hyp, seed=1, thr2=2.25, hyp_begin=b, hyp_count=c, out=keys)
    assert torch.equal(keys, full)
    H, cnt, mask = api.ransac_finalize(corr, n_hyp, 1, 2.25, keys, want_mask=True)
    kc, hyp = api.decode_keys(keys)
    assert torch.equal(cnt.long(), kc) and torch.equal(mask.sum(1).long(), kc)
    c_np = corr.cpu().numpy()
    for p in range(P):
        idx = oracle.ransac_sample(1, p, int(hyp[p]), n_pts)
        Hw = oracle.ransac_hypothesis(c_np[p], idx)
        assert np.array_equal(H[p].cpu().numpy().view(np.uint32), Hw.view(np.uint32))
        assert oracle.ransac_count(Hw, c_np[p], 2.25) == int(cnt[p])


def test_dist_driver_single_rank(api, cuda):
    from sks_homography_b200 import dist as sd
    corr = api.synth_corr(4, 1024, seed=2, device=cuda)
    H, cnt, hyp, _ = sd.ransac_aca(corr, 512, seed=3, thr2=4.0)
    keys = api.ransac_keys(corr, 512, seed=3, thr2=4.0)
    kc, kh = api.decode_keys(keys)
    assert torch.equal(cnt.long(), kc) and torch.equal(hyp, kh)


@pytest.mark.parametrize("packed", [0, 1, 2, 3])
@pytest.mark.parametrize("n_pts,n_hyp", [(4096, 1024), (1001, 777), (9001, 600), (3, 40)])
def test_kernel_variants_match_oracle(api, sks, oracle, cuda, n_pts, n_hyp, packed):
    """Scalar, match-pair FFMA2, hypothesis-pair FFMA2 and FP-pipe-count (3) scorers, 2 or 4 hypotheses per thread: same keys."""
    corr = api.synth_corr(4, n_pts, seed=31, inlier_permille=600, device=cuda)
    want = oracle.ransac(corr.cpu().numpy(), n_hyp, seed=13, thr2=2.25)
    try:
        for hpt in (2, 4):
            assert sks.c.sks_cuda_set_ransac_tuning(hpt, 3, packed) == 0
            assert np.array_equal(u64(api.ransac_keys(corr, n_hyp, seed=13, thr2=2.25)), want)
    finally:
        sks.c.sks_cuda_set_ransac_tuning(2, 8, 3)


@pytest.mark.parametrize("hpt,packed", [(2, 0), (4, 0), (2, 1), (4, 1), (2, 2), (4, 2), (2, 3), (4, 3)])
def test_nonfinite_scores_are_never_inliers(api, sks, oracle, cuda, hpt, packed):
    """Overflowing / NaN hypotheses and matches: the GPU reads the inlier bit off the
    sign of acc, the oracle evaluates acc < 0; both must ignore NaN and +-inf alike."""
    assert sks.c.sks_cuda_set_ransac_tuning(hpt, 8, packed) == 0
    try:
        rng = np.random.default_rng(12)
        P, n_pts, n_hyp = 3, 1024, 1536
        corr = api.synth_corr(P, n_pts, seed=21, device=cuda).cpu().numpy()
        corr[0, ::7] *= 1e18          # products overflow to inf, inf - inf = NaN
        corr[1, ::5, 2:] = np.inf
        corr[2, ::3] = np.nan
        corr[2, 1::3, :2] = corr[2, 4, :2] if np.isfinite(corr[2, 4, 0]) else 1.0   # repeated points
        c = torch.from_numpy(corr).to(cuda)
        keys = api.ransac_keys(c, n_hyp, seed=2, thr2=4.0)
        want, counts = oracle.ransac(corr, n_hyp, seed=2, thr2=4.0, want_counts=True)
        assert np.array_equal(u64(keys), want)
        got = np.zeros_like(counts)
        for j in range(0, n_hyp, 97):
            k = api.ransac_keys(c, n_hyp, seed=2, thr2=4.0, hyp_begin=j, hyp_count=1)
            got[:, j] = (u64(k) >> np.uint64(32)).astype(np.uint32)
            assert np.array_equal(got[:, j], counts[:, j])
    finally:
        sks.c.sks_cuda_set_ransac_tuning(2, 8, 3)


def test_more_pairs_than_one_grid_dimension(api, oracle, cuda):
    """n_pairs > 65535 goes out in blocks of pairs; sampling must not depend on blocking."""
    P, n_pts, n_hyp = 70_001, 8, 6
    rng = np.random.default_rng(3)
    corr = rng.uniform(10, 150, size=(P, n_pts, 4)).astype(np.float32)
    corr[:, :, 2:] = corr[:, :, :2] * 1.02 + 2.0 + rng.uniform(-0.3, 0.3, size=(P, n_pts, 2)).astype(np.float32)
    keys = api.ransac_keys(torch.from_numpy(corr).to(cuda), n_hyp, seed=4, thr2=4.0)
    want = oracle.ransac(corr, n_hyp, seed=4, thr2=4.0)
    assert np.array_equal(u64(keys), want)


def test_peer_reducer_single_rank_epochs(api, cuda):
    """csrc/peer.cuh on one GPU (world = 1): buffers alternate by epoch parity and the
    arrival counter is monotonic, so repeated reduces need no reset."""
    from sks_homography_b200 import dist as sd
    red = sd.PeerReducer(1000, cuda)
    try:
        g = torch.Generator(device="cpu").manual_seed(0)
        for epoch in range(7):
            keys = torch.randint(0, 2**62, (1000,), generator=g, dtype=torch.int64).to(cuda)
            want = keys.clone()
            red.max_reduce_(keys)
            assert torch.equal(keys, want) and not red.timed_out()
    finally:
        red.close()


def test_peer_wait_timeout_is_fatal_and_leaves_keys_untouched(api, sks, cuda):
    """A wait for more arrivals than ever come (world = 2 on one rank) must time out, set the
    sticky status flag, leave keys_out as it was, and make the reducer refuse further work."""
    import ctypes as C
    from sks_homography_b200 import dist as sd
    red = sd.PeerReducer(64, cuda, timeout_s=0.05)
    try:
        keys = torch.arange(1, 65, dtype=torch.int64, device=cuda)
        red.max_reduce_(keys)                                   # epoch 0, world 1: fine
        assert not red.timed_out()
        out = torch.full((64,), -7, dtype=torch.int64, device=cuda)
        st = torch.cuda.current_stream(cuda).cuda_stream
        sks.check(sks.c.sks_cuda_peer_push_max(keys.data_ptr(), 64, red.blocks, 1, 0, red.epoch, st), "push")
        sks.check(sks.c.sks_cuda_peer_wait(red.own, 2, red.epoch, out.data_ptr(), 64, red.status.data_ptr(),
                                           C.c_double(0.05), st), "wait")        # a second rank never arrives
        assert red.timed_out()
        assert bool((out == -7).all()), "keys_out must be untouched after a timeout"
        with pytest.raises(RuntimeError, match="did not arrive"):
            red.max_reduce_(keys)
    finally:
        red.close(check=False)


def test_explicit_samples_are_reduced_modulo_n_pts(api, oracle, cuda):
    """A raw 32-bit stream (here the cuRAND-compatible generator) is a valid sample list: entries
    are taken % n_pts like the reference's get_rand_list (GPU.cu:55-58), on the GPU and in the oracle."""
    P, n_pts, n_hyp = 3, 777, 400
    corr = api.synth_corr(P, n_pts, seed=8, device=cuda)
    raw = api.curand_mrg32k3a(P * n_hyp * 4, seed=11, device=cuda).view(P, n_hyp, 4)
    assert int((raw.cpu().numpy().view(np.uint32) >= n_pts).sum()) > 0
    keys = api.ransac_keys(corr, n_hyp, seed=0, thr2=2.25, samples=raw)
    want = oracle.ransac(corr.cpu().numpy(), n_hyp, 0, 2.25, samples=raw.cpu().numpy().view(np.uint32))
    assert np.array_equal(u64(keys), want)
    reduced = torch.from_numpy((raw.cpu().numpy().view(np.uint32) % n_pts).astype(np.int32)).to(cuda)
    assert torch.equal(api.ransac_keys(corr, n_hyp, seed=0, thr2=2.25, samples=reduced), keys)
    H, cnt, _ = api.ransac_finalize(corr, n_hyp, 0, 2.25, keys, samples=raw)
    kc, _ = api.decode_keys(keys)
    assert torch.equal(cnt.long(), kc)


def test_finalize_of_an_unscored_pair_is_no_model(api, cuda):
    """best_key == 0 (nothing scored) must not decode to hypothesis 0xFFFFFFFF: NaN model, count 0."""
    corr = api.synth_corr(2, 100, seed=8, device=cuda)
    keys = api.ransac_keys(corr, 64, seed=1, thr2=2.25)
    keys[1] = 0
    H, cnt, mask = api.ransac_finalize(corr, 64, 1, 2.25, keys, want_mask=True)
    assert bool(torch.isfinite(H[0]).all()) and int(cnt[0]) == int(keys[0] >> 32)
    assert bool(torch.isnan(H[1]).all()) and int(cnt[1]) == 0 and int(mask[1].sum()) == 0
    samples = torch.zeros((2, 64, 4), dtype=torch.int32, device=cuda)
    H2, cnt2, _ = api.ransac_finalize(corr, 64, 1, 2.25, keys, samples=samples)
    assert bool(torch.isnan(H2[1]).all()) and int(cnt2[1]) == 0


def test_sample_list_validation(api, cuda):
    corr = api.synth_corr(2, 100, seed=8, device=cuda)
    with pytest.raises(TypeError):
        api.ransac_keys(corr, 16, 1, 2.25, samples=torch.zeros((2, 16, 4), dtype=torch.int64, device=cuda))
    with pytest.raises(ValueError):
        api.ransac_keys(corr, 16, 1, 2.25, samples=torch.zeros((2, 15, 4), dtype=torch.int32, device=cuda))
    with pytest.raises(ValueError):
        api.ransac_keys(corr, 16, 1, 2.25, samples=torch.zeros((2, 16, 4), dtype=torch.int32))


def test_host_pointer_entry_point(api, oracle, cuda):
    """sks_host_ransac_aca_f32: matches in host memory in, models / counts / masks / keys out."""
    P, n_pts, n_hyp = 3, 2000, 1500
    corr = api.synth_corr(P, n_pts, seed=8, device=cuda).cpu()
    H, cnt, mask, keys = api.ransac_host(corr, n_hyp, seed=4, thr2=2.25, want_mask=True)
    want = oracle.ransac(corr.numpy(), n_hyp, seed=4, thr2=2.25)
    assert np.array_equal(u64(keys), want)
    kc, hyp = api.decode_keys(keys)
    assert torch.equal(cnt.long(), kc) and torch.equal(mask.sum(1).long(), kc)
    for p in range(P):
        Hw = oracle.ransac_hypothesis(corr[p].numpy(), oracle.ransac_sample(4, p, int(hyp[p]), n_pts))
        assert np.array_equal(H[p].numpy().view(np.uint32), Hw.view(np.uint32))
    samples = torch.from_numpy(np.array([[oracle.ransac_sample(4, p, j, n_pts) for j in range(n_hyp)]
                                         for p in range(P)], dtype=np.uint32).view(np.int32))
    _, _, _, k2 = api.ransac_host(corr.pin_memory(), n_hyp, seed=0, thr2=2.25, samples=samples)
    assert torch.equal(k2, keys)


def test_pair_shards_equal_rows_of_the_unsharded_run(api, sks, oracle, cuda):
    """Multi-GPU variant B (pairs sharded, no collective): a shard called with its global
    pair_begin reproduces exactly its rows of the unsharded result (and of the oracle)."""
    from sks_homography_b200 import dist as sd
    P, n_pts, n_hyp = 7, 1500, 900
    corr = api.synth_corr(P, n_pts, seed=5, device=cuda)
    full = api.ransac_keys(corr, n_hyp, seed=3, thr2=2.25)
    assert np.array_equal(u64(full), oracle.ransac(corr.cpu().numpy(), n_hyp, seed=3, thr2=2.25))
    Hf, cf, _ = api.ransac_finalize(corr, n_hyp, 3, 2.25, full)
    got_H, got_c, got_k = [], [], []
    for r in range(3):
        b, c = sks.shard_range(P, r, 3)
        shard = api.synth_corr(c, n_pts, seed=5, device=cuda, pair_begin=b)     # a rank generates only its pairs
        assert torch.equal(shard, corr[b:b + c])
        H, cnt, hyp, _ = sd.ransac_aca_pairs(shard, b, n_hyp, 3, 2.25)
        got_H.append(H); got_c.append(cnt); got_k.append(hyp)
    assert torch.equal(torch.cat(got_H), Hf) and torch.equal(torch.cat(got_c), cf)
    assert torch.equal(torch.cat(got_k), api.decode_keys(full)[1])
    # without the global pair id the samples differ (the id keys the RNG)
    assert not torch.equal(api.ransac_keys(corr[3:].contiguous(), n_hyp, 3, 2.25), full[3:])


def test_refit_on_inlier_mask(api, oracle, cuda):
    """Post-RANSAC hook: least-squares refit of each winner on its inlier mask, bit-exact with the
    CPU port (same lane partial sums, same shuffle-tree order, same LU), and at least as good."""
    P, n_pts, n_hyp = 9, 3000, 2048
    corr = api.synth_corr(P, n_pts, seed=12, device=cuda)
    keys = api.ransac_keys(corr, n_hyp, seed=6, thr2=2.25)
    H, cnt, mask = api.ransac_finalize(corr, n_hyp, 6, 2.25, keys, want_mask=True)
    Hr, used = api.ransac_refit(corr, mask, H)
    want, want_used = oracle.ransac_refit(corr.cpu().numpy(), mask.cpu().numpy(), H.cpu().numpy())
    assert np.array_equal(Hr.cpu().numpy().view(np.uint32), want.view(np.uint32))
    assert np.array_equal(used.cpu().numpy().astype(np.uint32), want_used)
    assert torch.equal(used.long(), cnt.long())
    c_np, m_np = corr.cpu().numpy().astype(np.float64), mask.cpu().numpy().astype(bool)

    def rms(Hs):                                 # transfer error over each pair's mask, in pixels
        out = []
        for p in range(P):
            x = c_np[p][m_np[p]]
            q = np.concatenate([x[:, :2], np.ones((len(x), 1))], 1) @ Hs[p].astype(np.float64).reshape(3, 3).T
            out.append(np.sqrt((((q[:, :2] / q[:, 2:]) - x[:, 2:]) ** 2).sum(1).mean()))
        return np.array(out)
    e_before, e_after = rms(H.cpu().numpy()), rms(Hr.cpu().numpy())
    assert (e_after < e_before).all() and e_after.max() < 0.6     # noise is +-0.5 px uniform
    after = sum(oracle.ransac_count(Hr[p].cpu().numpy(), corr[p].cpu().numpy(), 2.25) for p in range(P))
    assert after >= 0.98 * int(cnt.sum())        # and it still explains the winner's inliers
    mask[0] = 0
    mask[0, :3] = 1                              # too few inliers: the 4-point winner is kept
    Hk, uk = api.ransac_refit(corr, mask, H)
    assert torch.equal(Hk[0], H[0]) and int(uk[0]) == 0


def test_score_given_models_and_polish(api, oracle, cuda):
    """sks_cuda_ransac_score_f32 == the oracle's count for arbitrary models; the LO-style polish
    (score / refit / score, accept only improvements) never loses inliers and reproduces the
    same loop run with the CPU ports."""
    P, n_pts, n_hyp = 6, 2500, 1024
    corr = api.synth_corr(P, n_pts, seed=14, device=cuda)
    keys = api.ransac_keys(corr, n_hyp, seed=8, thr2=2.25)
    H, cnt, mask = api.ransac_finalize(corr, n_hyp, 8, 2.25, keys, want_mask=True)
    c2, m2 = api.ransac_score(corr, H, 2.25)
    assert torch.equal(c2, cnt) and torch.equal(m2, mask)
    Hp, cp, mp = api.ransac_polish(corr, H, 2.25, iters=2)
    assert (cp >= cnt).all() and torch.equal(mp.sum(1).int(), cp)
    # the same loop with the CPU ports
    c_np = corr.cpu().numpy()
    Hc, cc = H.cpu().numpy().copy(), cnt.cpu().numpy().astype(np.int64).copy()
    mc = mask.cpu().numpy().copy()
    for _ in range(2):
        H2, _ = oracle.ransac_refit(c_np, mc, Hc)
        for p in range(P):
            n2 = oracle.ransac_count(H2[p], c_np[p], 2.25)
            if n2 > cc[p]:
                Hc[p], cc[p] = H2[p], n2
                mc[p] = api.ransac_score(corr[p:p + 1], torch.from_numpy(H2[p:p + 1]).to(cuda), 2.25)[1][0].cpu().numpy()
    assert np.array_equal(Hp.cpu().numpy().view(np.uint32), Hc.view(np.uint32))
    assert np.array_equal(cp.cpu().numpy().astype(np.int64), cc)


def test_real_matches_fixture_gather_ransac_refit(api, oracle, cuda):
    """The reference's real-data fixture (2540 wall matches, CPU/orig_pts_wall.txt, staged into the
    git-ignored oracle/_ref at build time): minimal-sample gather -> solve replays the reference's GPU
    flow (GPU.cu:1443-1464) on real matches, the fused RANSAC finds the wall's homography with the
    oracle's exact keys (~63 % of the matches at 3 px), and the refit keeps it."""
    import os
    from oracle.oracle import REF_FIXTURE
    from sks_homography_b200.io import read_points
    if not os.path.exists(REF_FIXTURE):
        pytest.skip("fixture not staged (oracle/_ref/orig_pts_wall.txt; run `make -C oracle stage`)")
    pool_np = read_points(REF_FIXTURE)
    assert pool_np.shape == (2540, 4)
    corr = torch.from_numpy(pool_np)[None].contiguous().to(cuda)
    n_hyp, seed, thr2 = 4096, 11, 9.0
    keys = api.ransac_keys(corr, n_hyp, seed, thr2)
    want = oracle.ransac(pool_np[None], n_hyp, seed, thr2)
    assert np.array_equal(u64(keys), want)
    H, cnt, mask = api.ransac_finalize(corr, n_hyp, seed, thr2, keys, want_mask=True)
    frac = int(cnt[0]) / 2540
    assert 0.55 < frac < 0.70 and int(mask.sum()) == int(cnt[0])
    H2, used = api.ransac_refit(corr, mask, H)
    cnt2, _ = api.ransac_score(corr, H2, thr2)
    assert int(used[0]) == int(cnt[0]) and int(cnt2[0]) >= 0.97 * int(cnt[0])
    # hypothesis generation from the pool with the reference's cuRAND sample list (seed 11, GPU.cu:1445)
    n = 8192
    rand4 = api.curand_mrg32k3a(4 * n, seed=11, device=cuda).view(4, n)
    Hg = api.gather_solve("aca", corr[0].double(), n, rand4=rand4)
    idx = (rand4.cpu().numpy().view(np.uint32) % 2540).T                     # [n, 4]
    sel = pool_np.astype(np.float64)[idx]                                    # [n, 4, 4]
    src, tar = sel[:, :, :2].reshape(n, 8), sel[:, :, 2:].reshape(n, 8)
    wantH = oracle.solve("aca", src, tar)
    got = Hg.cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(wantH))
    assert np.array_equal(np.nan_to_num(got).view(np.uint64), np.nan_to_num(wantH).view(np.uint64))
