#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched 4-point homography path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch of synthetic quadruples.
Default workload = BASELINE.json configs[1]: batched ACA, general quad-to-quad,
2^26 random quadruples per GPU, fp32, AoS in/out, h33-normalised (weak scaling:
every rank solves its own contiguous 2^26-quad shard of the global index space,
no data-path collective).  Rank 0 prints ONE JSON line.

  value     whole-job homographies/s, inputs resident in HBM, CUDA-event timed
  e2e       same metric through the host-pointer C ABI (sks_host_*): pinned host
            buffers, H2D and D2H copies inside the timed region
  roofline  algorithmic bytes (100 B/homography fp32 general) per
            launch / mean launch duration, against the measured HBM copy peak
  cpu_baseline  the reference's own C++ (oracle/_ref) on the host cores, bounded sample

--impl reference times the reference's CPU implementation of the same path on
the host cores (all hardware threads) and prints the same line with
"impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ACA/SKS homographies/s"
UNIT = "homographies/s"

# The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner,
# torchrun notices), so stdout's file descriptor is pointed at stderr for the whole run and the
# line goes to the saved descriptor.
_REAL_STDOUT = None


def _capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


WORKLOADS = {
    # name: (solver, dtype, bytes per homography (algorithmic), default log2 n, dist)
    "aca_f32": ("aca", "f32", 100, 26, 0),     # BASELINE configs[1]  (headline)
    "sks_f32": ("sks", "f32", 100, 26, 0),
    "aca_f64": ("aca", "f64", 200, 25, 1),
    "sks_f64": ("sks", "f64", 200, 25, 1),     # BASELINE configs[3]
    "rect_f32": ("rect", "f32", 68, 26, 0),    # BASELINE configs[2] per-GPU shard at 4 GPUs
    # competitor solver in the same harness; dist 1 because GE has no
    # pivoting and fails on the axis-aligned source squares of dist 0
    "ge_f32": ("ge", "f32", 100, 26, 1),
    "ge_f64": ("ge", "f64", 200, 25, 1),
    "gpt_f64": ("gpt", "f64", 200, 24, 1),     # GPT-LU, the arithmetic of cal_Homo_GPT (compute-bound)
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload: str):
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(workload)
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons, pw = [], [], set(), []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(self.NAMES, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args):
    """Reference arm: the reference's own C++ CPU path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from oracle.oracle import Oracle, RefLib
    if args.workload == "ransac":
        emit({"impl": "reference", "unavailable": "the reference has no RANSAC scorer"})
        return 0
    solver, dt, bytes_per_h, log2n, dist = WORKLOADS[args.workload]
    if solver == "rect":
        emit({"impl": "reference", "unavailable": "the reference has no C++ ACA-rect"})
        return 0
    o = Oracle()
    kind = "reference"
    try:
        if (solver == "ge" and dt == "f64") or solver == "gpt":
            raise LookupError("the reference's C++ GE is fp32 only; its GPT is OpenCV")
        ref = RefLib()
        threads = ref.hardware_threads()
        run = lambda s, t, out: ref.solve(solver, s, t, threads=threads, out=out)
    except Exception:
        kind, threads = "port", 1
        run = lambda s, t, out: out.__setitem__(slice(None), o.solve(solver, s, t))
    dtype = np.float32 if dt == "f32" else np.float64
    S = 1 << args.ref_log2n
    s, t = o.synth_quads(0, S, args.seed, dist, dtype)
    out = np.empty((S, 9), dtype=dtype)
    for _ in range(args.warmup):
        run(s, t, out)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(s, t, out)
    el = time.perf_counter() - t0
    value = S * args.steps / el
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": dt, "data": "synthetic",
        "config": {"workload": f"batched {solver.upper()} {dt} 2^{args.log2n or log2n} quadruples per GPU, AOS in/out, "
                               f"h33-normalised -- reference C++ on the host cores, each step a bounded sample "
                               f"of 2^{args.ref_log2n} distinct quadruples streamed from memory",
                   "threads": threads, "compiler": "g++ -O2 -ffp-contract=off"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"2^{args.ref_log2n} quadruples per step x {args.steps} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="aca_f32", choices=sorted(WORKLOADS) + ["ransac"])
    ap.add_argument("--pairs", type=int, default=1024)
    ap.add_argument("--points", type=int, default=4096)
    ap.add_argument("--hyps", type=int, default=65536)
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--hpt", type=int, default=0, help="RANSAC hypotheses per thread (2|4)")
    ap.add_argument("--rounds", type=int, default=8, help="RANSAC rounds per CTA")
    ap.add_argument("--packed", type=int, default=-1, help="RANSAC FFMA2 scoring (0|1)")
    ap.add_argument("--ransac-shard", default="hypotheses", choices=["hypotheses", "pairs"],
                    help="multi-GPU RANSAC: shard the hypothesis ids of every pair (one max all-reduce; "
                         "north_star's variant) or the image pairs (no collective at all)")
    ap.add_argument("--peer-reduce", action="store_true",
                    help="RANSAC: merge the winners with the hand-written NVLink peer max-reduce "
                         "instead of the NCCL all-reduce")
    ap.add_argument("--log2n", type=int, default=None, help="quadruples per GPU = 2^log2n")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: 2^log2n quadruples in TOTAL, sharded contiguously over the ranks "
                         "(BASELINE configs[2]: --workload rect_f32 --log2n 28 --strong)")
    ap.add_argument("--layout", default="aos", choices=["aos", "soa"])
    ap.add_argument("--no-normalize", action="store_true")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--small-tile", type=int, default=0)
    ap.add_argument("--stages", type=int, default=4)
    ap.add_argument("--ctas", type=int, default=0)
    ap.add_argument("--seed", type=int, default=11)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-log2n", type=int, default=None)
    ap.add_argument("--cpu-log2n", type=int, default=24)
    ap.add_argument("--ref-log2n", type=int, default=24)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="multi-rank runs: do not pin the rank to its GPU's NUMA node")
    args = ap.parse_args()
    _capture_stdout()
    if args.warmup < 3:
        args.warmup = max(args.warmup, 3)          # timing rules: W >= 3
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from sks_homography_b200 import api, lib
    from sks_homography_b200 import dist as sdist
    L = lib()
    # one process per GPU: keep the rank's host threads and pinned buffers on its GPU's socket
    args.numa = sdist.bind_to_gpu_numa_node(local) if (world > 1 and not args.no_numa_bind) else {"bound": False}
    L.check(L.c.sks_cuda_set_variant(args.variant), "set_variant")
    L.check(L.c.sks_cuda_set_tuning(args.small_tile, args.stages, args.ctas), "set_tuning")

    if args.workload == "ransac":
        if args.hpt or args.packed >= 0:
            L.check(L.c.sks_cuda_set_ransac_tuning(args.hpt or 2, args.rounds, max(args.packed, 0)),
                    "set_ransac_tuning")
        return run_ransac(args, api, L, dev, rank, world, local)

    solver, dt, bytes_per_h, log2n, dist_id = WORKLOADS[args.workload]
    log2n = args.log2n if args.log2n is not None else log2n
    if args.strong:
        begin, n = L.shard_range(1 << log2n, rank, world)     # contiguous shard of a fixed total
    else:
        n = 1 << log2n
        begin = rank * n                              # this rank's shard of the global index space
    tdt = torch.float32 if dt == "f32" else torch.float64
    normalize = not args.no_normalize

    # ---- inputs resident in HBM, generated on the device -----------------------
    src, tar = api.synth_quads(n, seed=args.seed, dist=dist_id, dtype=tdt, device=dev, begin=begin,
                               layout=args.layout)
    H = torch.empty((n, 9) if args.layout == "aos" else (9, n), dtype=tdt, device=dev)

    def step():
        if solver == "rect":
            api.aca_rect(tar, 128.0, 1.0, 15.0, 12.0, result=H, normalize=normalize, layout=args.layout)
        else:
            api.solve(solver, src, tar, result=H, normalize=normalize, layout=args.layout)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    L.c.sks_cuda_reset_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    torch.cuda.synchronize()
    launches = int(L.c.sks_cuda_launch_count())
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop() if sampler else None
    barrier()
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max = float(tmax.item())
    ms_per_step = total_ms_max / args.steps
    n_total = (1 << log2n) if args.strong else world * n
    value = n_total / (ms_per_step * 1e-3)

    peak, peak_src = peaks()
    mean_launch_ms = statistics.mean(per_launch_ms)
    achieved = n * bytes_per_h / (mean_launch_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(args.workload),
                "peak_source": peak_src, "bytes_per_homography": bytes_per_h,
                "launch_ms_mean": mean_launch_ms, "launch_ms_min": min(per_launch_ms),
                "frac_of_nominal_8TBs": achieved / 8000.0}

    # ---- end to end through the host-pointer C ABI ------------------------------
    e2e = None
    if not args.no_e2e and args.layout == "aos":
        # 2^25 quadruples per GPU per step by default: 3.4 GB of pinned host memory per
        # rank, so 8 ranks stay far below the host's RAM (the kernel-only `value` keeps 2^26)
        ne = 1 << (args.e2e_log2n if args.e2e_log2n is not None else min(log2n, 25))
        ne = min(ne, n)
        hs = torch.empty((ne, 8), dtype=tdt, pin_memory=True)
        ht = torch.empty((ne, 8), dtype=tdt, pin_memory=True)
        hH = torch.empty((ne, 9), dtype=tdt, pin_memory=True)
        hs.copy_(src[:ne]); ht.copy_(tar[:ne])
        torch.cuda.synchronize()

        def host_step():
            if solver == "rect":
                api.aca_rect(ht, 128.0, 1.0, 15.0, 12.0, result=hH, normalize=normalize)
            else:
                api.solve(solver, hs, ht, result=hH, normalize=normalize)

        host_step()                                   # warm-up: ring buffers, streams
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            host_step()                               # synchronous: returns with H in host memory
        el = time.perf_counter() - t0
        te = torch.tensor([el], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        el = float(te.item())
        esz = 4 if dt == "f32" else 8
        in_elems = 8 if solver == "rect" else 16
        e2e = {"value": world * ne * args.e2e_steps / el, "unit": UNIT,
               "h2d_bytes_per_step": ne * in_elems * esz, "d2h_bytes_per_step": ne * 9 * esz,
               "steps": args.e2e_steps, "quadruples_per_step_per_gpu": ne,
               "api": "sks_host_* (pinned host buffers, chunked H2D/kernel/D2H ring)",
               "bound": "PCIe: every homography moves %d B host->device and %d B back; at the ~50 GB/s a Gen5 x16 "
                        "link sustains with both directions busy that is a ceiling of ~%.2f G H/s per GPU, "
                        "whatever the kernel does (DESIGN.md section 3, host-pointer path)"
                        % (in_elems * esz, 9 * esz, 50.0 / (in_elems * esz)),
               "numa_binding": args.numa}
        # the host path must give the same bytes as the device path
        if not torch.equal(hH.view(torch.int32 if dt == "f32" else torch.int64),
                           H[:ne].cpu().view(torch.int32 if dt == "f32" else torch.int64)):
            same_nan = torch.equal(torch.isnan(hH), torch.isnan(H[:ne].cpu()))
            e2e["parity_vs_device_path"] = "nan-pattern-equal" if same_nan else "MISMATCH"
        else:
            e2e["parity_vs_device_path"] = "bit-exact"
        del hs, ht, hH

    # ---- CPU baseline beside it (rank 0, N = 1): the reference's own C++ ---------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and args.layout == "aos" and solver != "rect":
        from oracle.oracle import Oracle, RefLib
        S = min(n, 1 << args.cpu_log2n)
        s_h, t_h = src[:S].cpu().numpy(), tar[:S].cpu().numpy()
        out = np.empty((S, 9), dtype=s_h.dtype)
        try:
            if (solver == "ge" and dt == "f64") or solver == "gpt":
                raise LookupError("the reference's C++ GE is fp32 only; its GPT is OpenCV")
            ref = RefLib()
            threads, kind = ref.hardware_threads(), "reference"
            run = lambda: ref.solve(solver, s_h, t_h, threads=threads, out=out)
        except Exception:
            o = Oracle()
            threads, kind = 1, "port"
            run = lambda: out.__setitem__(slice(None), o.solve(solver, s_h, t_h))
        run()
        best = 1e30
        for _ in range(5):
            t0 = time.perf_counter(); run(); best = min(best, time.perf_counter() - t0)
        got = H[:S].cpu().numpy() if normalize else None
        parity = None
        if got is not None:
            v = np.uint32 if dt == "f32" else np.uint64
            ok = (got.view(v) == out.view(v)) | (np.isnan(got) & np.isnan(out))
            parity = {"checked_quadruples": int(S), "mismatching_elements": int((~ok).sum())}
        o3 = None
        if kind == "reference":
            try:                       # speed-only: the same files at -O3 with AVX2/FMA (bits differ)
                ref3 = RefLib(o3=True)
                out3 = np.empty_like(out)
                ref3.solve(solver, s_h, t_h, threads=threads, out=out3)
                b3 = 1e30
                for _ in range(5):
                    t0 = time.perf_counter(); ref3.solve(solver, s_h, t_h, threads=threads, out=out3)
                    b3 = min(b3, time.perf_counter() - t0)
                o3 = {"value": S / b3, "flags": "g++ -O3 -march=x86-64-v3 -ffp-contract=fast",
                      # relative to each matrix's largest element (FMA contraction changes low bits)
                      "max_diff_vs_O2_rel_to_matrix_scale": float(np.nanmax(
                          np.abs(out3 - out) / np.nanmax(np.abs(out), axis=1, keepdims=True)))}
            except Exception as e:
                o3 = {"unavailable": str(e)[:100]}
        cpu = {"value": S / best, "unit": UNIT, "cores": threads, "kind": kind, "o3_fma_build": o3,
               "sample": f"first 2^{args.cpu_log2n} quadruples of the workload, best of 5 passes, "
                         f"{'MOD/GE.cpp' if solver == 'ge' else 'GPU.cu:242-357' if solver == 'gpt' else 'MOD/ACA_SKS.cpp'} g++ -O2 -ffp-contract=off"
                         f"{'' if kind == 'reference' else ' (oracle port)'}, {threads} threads",
               "parity_gpu_vs_cpu": parity}

    gpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu and not args.no_gpu_baseline:
        gpu_base = reference_gpu_kernels(api, dev)

    if rank == 0:
        line = {
            "impl": "ours", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
            "dtype": dt, "data": "synthetic",
            "config": {"workload": f"batched {solver.upper()} {dt} 2^{log2n} quadruples "
                                   f"{'in total, sharded over the ranks' if args.strong else 'per GPU'}, "
                                   f"{args.layout.upper()} in/out, "
                                   f"{'h33-normalised' if normalize else 'up to scale'}",
                       "quadruples_per_gpu": n, "layout": args.layout, "variant": args.variant,
                       "l2": f"inputs+outputs {n * bytes_per_h / 1e9:.2f} GB per step >> 126 MB L2, no flush needed",
                       "seed": args.seed, "dist": dist_id},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks, "gpu_baseline": gpu_base,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def _event_ms(fn, iters, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2]


def reference_gpu_kernels(api, dev):
    """Same-box GPU comparator: the reference's own CUDA kernels (GPU.cu:81-240,
    compiled for sm_100a at build time into oracle/_ref) launched as the reference
    does -- fp64, SoA, <<<ceil(N/32),32>>>, un-normalised -- next to ours on the same
    buffers.  N = 2^20 is the largest row of the paper's Table 8; 2^25 is config 4."""
    import torch
    try:
        from oracle.oracle import RefGpuLib
        ref = RefGpuLib()
    except Exception as e:                                   # not built: report, do not fail
        return {"unavailable": str(e)[:120]}
    out = {"kernels": "GPU_Runtime Test.cu:81-240 cal_Homo_ACA/SKS and :359-507 cal_Homo_GE, nvcc -O3 "
                      "sm_100a, block 32, fp64 SoA un-normalised", "rows": []}
    for log2n in (20, 25):
        n = 1 << log2n
        src, tar = api.synth_quads(n, 11, 1, torch.float64, dev, layout="soa")
        H = torch.empty((9, n), dtype=torch.float64, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        for solver in ("aca", "sks", "ge", "gpt"):
            if solver == "gpt" and log2n > 20:
                continue                      # compute-bound: the Table-8 row is enough
            t_ref = _event_ms(lambda: ref.run(solver, src.data_ptr(), tar.data_ptr(), H.data_ptr(), n, st), 20)
            t_our = _event_ms(lambda: api.solve(solver, src, tar, result=H, normalize=False, layout="soa"), 20)
            out["rows"].append({"solver": solver, "n": n, "reference_us": 1e3 * t_ref, "ours_us": 1e3 * t_our,
                                "reference_GHps": n / t_ref / 1e6, "ours_GHps": n / t_our / 1e6,
                                "reference_GBps": n * 200 / t_ref / 1e6, "ours_GBps": n * 200 / t_our / 1e6})
        del src, tar, H
    return out


def run_ransac(args, api, L, dev, rank, world, local):
    """BASELINE configs[4]: fused ACA-RANSAC, P pairs x n_pts matches x n_hyp
    hypotheses.  Multi-GPU variant (A) of SURVEY.md 8(e): every rank holds all
    pairs' matches and scores its contiguous shard of the hypothesis ids; ONE
    int64 max-all-reduce (NCCL over NVLink) merges the winners; the winning model
    is recomputed locally.  Strong scaling: total work is fixed."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from sks_homography_b200 import dist as sd

    P, n_pts, n_hyp = args.pairs, args.points, args.hyps
    thr2 = 2.25
    by_pairs = args.ransac_shard == "pairs" and world > 1
    if by_pairs:
        # SURVEY.md 8(e) variant B: this rank owns pairs [pb, pb + pc) and scores every hypothesis
        # id of them; nothing is exchanged (the throughput-optimal split for many pairs)
        pb, pc = sd.shard_range(P, rank, world)
        corr = api.synth_corr(pc, n_pts, seed=args.seed, inlier_permille=500, noise=0.5, device=dev,
                              pair_begin=pb)
        hb, hc = 0, n_hyp
    else:
        pb, pc = 0, P
        corr = api.synth_corr(P, n_pts, seed=args.seed, inlier_permille=500, noise=0.5, device=dev)
        hb, hc = sd.shard_range(n_hyp, rank, world)
    keys = torch.zeros(pc, dtype=torch.int64, device=dev)
    res = {}
    reducer = sd.PeerReducer(P, dev) if (args.peer_reduce and not by_pairs) else None

    def step():
        keys.zero_()
        api.ransac_keys(corr, n_hyp, args.seed, thr2, None, hb, hc, out=keys, pair_begin=pb)
        if by_pairs:
            pass
        elif reducer is not None:
            reducer.max_reduce_(keys)
        else:
            sd.merge_keys(keys)
        res["H"], res["cnt"], _ = api.ransac_finalize(corr, n_hyp, args.seed, thr2, keys, pair_begin=pb)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    L.c.sks_cuda_reset_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    torch.cuda.synchronize()
    launches = int(L.c.sks_cuda_launch_count())
    total_ms = ev[0].elapsed_time(ev[-1])
    clocks = sampler.stop() if sampler else None
    barrier()
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_per_step = float(tmax.item()) / args.steps
    value = P * n_hyp / (ms_per_step * 1e-3)

    # FP32 roofline: 21 flop per hypothesis x point (8 FMA + 1 MUL + 2 FMA once the threshold is
    # folded into the operands; SURVEY.md 8(d) counted 22 for the unfolded form) + 97 + 6 per hypothesis
    flops = P * float(n_hyp) * (103.0 + 21.0 * n_pts)
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    peak = 148 * 128 * 2 * mhz * 1e6 / 1e12 * world
    achieved = flops / (ms_per_step * 1e-3) / 1e12
    roofline = {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": None,
                "peak_source": f"148 SM x 128 lanes x 2 x {mhz:.0f} MHz (SM clock sampled under load) x {world} GPU",
                "flop_per_hypothesis_point": 21, "flop_per_hypothesis": 103,
                # tools/ubench/fma_peak.cu on this pool's B200s: 121.8 of the nominal 128 FMA/clk/SM
                # are attainable with reuse-friendly operands, 84.7 with three fresh register pairs
                "frac_of_measured_fma_peak": achieved / (peak * 121.8 / 128.0)}

    # parity at full size: a few pairs against the CPU oracle on a hypothesis prefix,
    # and (multi-GPU) merged keys == single-GPU keys on the same pairs
    parity = None
    if rank == 0 and not args.no_cpu:
        from oracle.oracle import Oracle
        o = Oracle()
        sel = [0, pc // 2, pc - 1]
        sub = corr[sel].contiguous()
        nh = min(n_hyp, 2048)
        got = api.ransac_keys(sub, n_hyp, args.seed, thr2, None, 0, nh).cpu().numpy().view(np.uint64)
        want = o.ransac(sub.cpu().numpy(), nh, args.seed, thr2)      # same pair ids 0..2, hyps 0..nh-1
        full1 = None
        if world > 1 and not by_pairs:
            full1 = api.ransac_keys(sub, n_hyp, args.seed, thr2)
            k2 = torch.zeros(3, dtype=torch.int64, device=dev)
            for r in range(world):
                b, c = sd.shard_range(n_hyp, r, world)
                api.ransac_keys(sub, n_hyp, args.seed, thr2, None, b, c, out=k2)
            full1 = bool(torch.equal(full1, k2))
        parity = {"pairs_checked": 3, "hypotheses_checked": nh,
                  "keys_equal_cpu_oracle": bool(np.array_equal(got, want)),
                  "sharded_equals_unsharded": full1}
    # e2e: the same step with the matches in pinned HOST memory -- H2D of every pair's matches,
    # scoring, winner merge, finalize, D2H of the models and inlier counts inside the timed region
    e2e = None
    if not args.no_e2e:
        h_corr = corr.cpu().pin_memory()
        d_corr = torch.empty_like(corr)
        h_H = torch.empty((pc, 9), dtype=torch.float32).pin_memory()
        h_cnt = torch.empty(pc, dtype=torch.int32).pin_memory()

        def e2e_step():
            if world == 1 and reducer is None:      # the host-pointer C-ABI call, as a C++ caller would make it
                H, c, _, _ = api.ransac_host(h_corr, n_hyp, args.seed, thr2)
                h_H.copy_(H); h_cnt.copy_(c)
                return
            d_corr.copy_(h_corr, non_blocking=True)
            keys.zero_()
            api.ransac_keys(d_corr, n_hyp, args.seed, thr2, None, hb, hc, out=keys, pair_begin=pb)
            if by_pairs:
                pass
            elif reducer is not None:
                reducer.max_reduce_(keys)
            else:
                sd.merge_keys(keys)
            H, c, _ = api.ransac_finalize(d_corr, n_hyp, args.seed, thr2, keys, pair_begin=pb)
            h_H.copy_(H, non_blocking=True)
            h_cnt.copy_(c, non_blocking=True)

        e2e_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.e2e_steps):
            e2e_step()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item()) / args.e2e_steps
        e2e = {"value": P * n_hyp / (e2e_ms * 1e-3), "unit": "hypotheses/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": corr.numel() * 4, "d2h_bytes_per_step": pc * 9 * 4 + pc * 4,
               "api": "sks_host_ransac_aca_f32 (host matches in, models out)" if world == 1 and reducer is None
                      else "per-rank H2D + sks_cuda_ransac_aca_f32 on the rank's hypothesis shard + max-reduce + finalize",
               "models_equal_device_path": bool(torch.equal(h_H, res["H"].cpu()))}

    # CPU baseline: the oracle's scalar port of the same definition on one host core, on a bounded
    # sample (one pair, a prefix of its hypothesis ids)
    cpu = None
    if rank == 0 and not args.no_cpu:
        import time
        from oracle.oracle import Oracle
        o = Oracle()
        sub = corr[:1].cpu().numpy()
        nh = min(n_hyp, max(1024, int(1e9 // max(n_pts, 1))))
        t0 = time.perf_counter()
        o.ransac(sub, nh, args.seed, thr2)
        dt = time.perf_counter() - t0
        cpu = {"value": nh / dt, "unit": "hypotheses/s", "cores": 1, "kind": "port",
               "sample": f"1 pair x {n_pts} matches x {nh} hypotheses, oracle/sks_oracle.c "
                         f"(gcc -O2 -ffp-contract=off), {dt:.1f} s"}

    if rank == 0:
        cnt = res["cnt"].float()
        line = {
            "impl": "ours", "metric": "ACA-RANSAC hypotheses/s (homographies solved and scored)",
            "value": value, "unit": "hypotheses/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"fused ACA-RANSAC {P} pairs x {n_pts} matches x {n_hyp} hypotheses, "
                                   + (f"image pairs sharded over {world} GPU(s), no collective" if by_pairs else
                                      f"hypotheses sharded over {world} GPU(s), "
                                      + ("winners merged by NVLink peer atomics (csrc/peer.cuh)" if reducer
                                         else "one int64 max all-reduce (NCCL)")),
                       "thr2": thr2, "seed": args.seed,
                       "l2": "compute-bound; matches (64 KiB/pair) live in shared memory"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks, "parity": parity,
            "mean_inlier_fraction_of_winner": float(cnt.mean().item()) / n_pts,
            "peer_reduce_timed_out": reducer.timed_out() if reducer else None,
        }
        emit(line)
    if reducer is not None:
        reducer.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
