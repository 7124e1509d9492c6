"""bench.py output contract: one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
             "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def run_bench(*args, env=None):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True,
                         text=True, env={**os.environ, **(env or {})}, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, res.stdout
    return json.loads(lines[0])


def test_reference_arm_line(reflib):
    """`--impl reference`: the reference's own C++ on the host cores, same line shape."""
    d = run_bench("--impl", "reference", "--steps", "2", "--warmup", "3", "--ref-log2n", "16")
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["unit"] == "homographies/s" and d["value"] > 1e5 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None and "workload" in d["config"]


def test_both_arms_describe_the_same_config(reflib):
    """The driver compares the two arms' `config` dicts: the reference arm must print exactly what our arm
    prints for the headline workload (BASELINE configs[1]), whatever its bounded sample is."""
    sys.path.insert(0, ROOT)
    import bench
    solver, dt, bph, log2n, dist = bench.WORKLOADS["aca_f32"]
    ours = bench.stream_config(solver, dt, log2n, "aos", True, False, 1 << log2n, bph, 11, dist)
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "3", "--ref-log2n", "14")
    assert d["config"] == ours and "2^26" in ours["workload"] and ours["quadruples_per_gpu"] == 1 << 26


def test_reference_arm_non_zero_ranks_exit_quietly():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env={**os.environ, "RANK": "1", "WORLD_SIZE": "2"},
                         timeout=120)
    assert res.returncode == 0 and res.stdout.strip() == ""


@pytest.mark.gpu
def test_ours_arm_line(cuda):
    d = run_bench("--steps", "3", "--warmup", "3", "--log2n", "20", "--cpu-log2n", "18", "--no-gpu-baseline")
    assert d["impl"] == "ours" and BASE_KEYS | {"roofline", "clocks"} <= set(d)
    assert d["n_gpus"] == 1 and d["dtype"] == "f32" and d["scaling"] == "weak" and d["gpu_launches"] == 3
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == (1 << 20) * 64 and e["d2h_bytes_per_step"] == (1 << 20) * 36
    assert e["parity_vs_device_path"] == "bit-exact" and 0 < e["value"] < d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["parity_gpu_vs_cpu"]["mismatching_elements"] == 0


@pytest.mark.gpu
def test_ours_arm_legs_for_the_other_configs(cuda):
    """The `configs` legs (BASELINE configs[2..4]) at test sizes: RANSAC with its per-part breakdown and
    oracle parity, rect strong, SKS f64 with the accuracy tier -- the same code path the full-size run takes."""
    d = run_bench("--steps", "3", "--warmup", "3", "--log2n", "20", "--cpu-log2n", "16",
                  "--force-legs", "--leg-log2n", "20", "--pairs", "16", "--points", "512", "--hyps", "2048",
                  "--sustained-s", "0.2", "--ransac-cpu-evals", "2e6", "--e2e-steps", "3")
    c = d["configs"]
    r = c["ransac"]
    assert r["parity"]["keys_equal_cpu_oracle"] is True and r["roofline"]["bound"] == "fp32"
    assert set(r["breakdown"]) >= {"zero_keys_ms", "score_kernel_ms", "reduce_incl_wait_for_slowest_rank_ms", "finalize_ms"}
    assert c["rect_2p28_strong"]["quadruples_per_gpu"] == 1 << 20 and c["rect_2p28_strong"]["scaling"] == "strong"
    a = c["sks_f64_2p25"]["accuracy_tier"]
    assert a["mismatching_elements_vs_runKernel_SKS_double"] == 0 and a["reprojection_px"]["p99"] < 1e-9
    assert d["sustained"]["launches"] >= 200 and d["clocks"]["samples"] >= 5
    assert d["e2e"]["pageable"]["parity_vs_device_path"] == "bit-exact" and d["e2e"]["link"]["h2d_GBps"] > 5
    # same-box comparators: the reference's CUDA kernels and its torch-eager functions (both staged into
    # oracle/_ref at build time; "unavailable" where they were not)
    g = d["gpu_baseline"]
    if "rows" in g:
        assert all(r["ours_us"] > 0 and r["reference_us"] > 0 for r in g["rows"])
    te = g.get("torch_eager_fp32", {})
    if "rows" in te:
        good = [r for r in te["rows"] if "error" not in r]
        assert good and all(r["bit_identical"] for r in good) and all(r["speedup"] > 1 for r in good)
