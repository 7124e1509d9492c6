#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched 4-point homography path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch of synthetic quadruples.
Headline workload = BASELINE.json configs[1]: batched ACA, general quad-to-quad,
2^26 random quadruples per GPU, fp32, AoS in/out, h33-normalised (weak scaling:
every rank solves its own contiguous 2^26-quad shard of the global index space,
no data-path collective).  Rank 0 prints ONE JSON line.

  value     whole-job homographies/s, inputs resident in HBM, CUDA-event timed
  e2e       same metric, same 2^26 quadruples, through the host-pointer C ABI
            (sks_host_*): pinned host buffers, H2D and D2H copies inside the timed
            region; e2e.pageable = the same from plain malloc'd arrays (what the
            reference's CPU/main.cpp:47-58 holds); e2e.link = the PCIe ceiling
            measured in the same run
  roofline  algorithmic bytes (100 B/homography fp32 general) per
            launch / mean launch duration, against the measured HBM copy peak
  sustained the headline kernel back to back for >= 2.5 s with the clock sampler
            at 20 ms (clocks come from this window plus the timed steps)
  cpu_baseline  the reference's own C++ (oracle/_ref) on the host cores, bounded sample
  gpu_baseline  the reference's existing GPU implementations on the same box: its fp64
            SoA CUDA kernels recompiled for sm_100a, and its fp32 torch-eager functions
            (TensorACA_rect, ACA_vanilla) executed unmodified on cuda tensors
  configs   the other BASELINE configs as short legs in the same process group:
            ransac (configs[4], hypothesis-sharded, NCCL max all-reduce, strong
            scaling), rect_2p28_strong (configs[2]), sks_f64_2p25 (configs[3]), and
            ransac_inprocess_multi (the C-ABI single-process multi-GPU driver)

--impl reference times the reference's CPU implementation of the same path on
the host cores (all hardware threads), each step one pass over the SAME 2^26
quadruples, and prints the same line with "impl": "reference".
--workload X benches one of the other workloads as the top-level line instead.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
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


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


WORKLOADS = {
    # name: (solver, dtype, bytes per homography (algorithmic), default log2 n, dist)
    "aca_f32": ("aca", "f32", 100, 26, 0),     # BASELINE configs[1]  (headline)
    "sks_f32": ("sks", "f32", 100, 26, 0),
    "aca_f64": ("aca", "f64", 200, 25, 1),
    "sks_f64": ("sks", "f64", 200, 25, 1),     # BASELINE configs[3]
    "rect_f32": ("rect", "f32", 68, 26, 0),    # BASELINE configs[2]: --log2n 28 --strong
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


def stream_config(solver, dt, log2n, layout, normalize, strong, n, bytes_per_h, seed, dist_id) -> dict:
    """The workload description BOTH arms print (the driver compares the two dicts)."""
    return {"workload": f"batched {solver.upper()} {dt} 2^{log2n} quadruples "
                        f"{'in total, sharded over the ranks' if strong else 'per GPU'}, "
                        f"{layout.upper()} in/out, {'h33-normalised' if normalize else 'up to scale'}",
            "quadruples_per_gpu": n, "layout": layout,
            "l2": f"inputs+outputs {n * bytes_per_h / 1e9:.2f} GB per step >> 126 MB L2, no flush needed",
            "seed": seed, "dist": dist_id}


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING a measured region: an NVML polling
    thread at `period_s` (default 20 ms); nvidia-smi -lms as the fallback."""
    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
             0x80: "hw_power_brake_slowdown"}

    def __init__(self, gpu_index: int, period_s: float = 0.02):
        self.period = period_s
        self.sm, self.pw, self.reasons, self.mx = [], [], set(), None
        self.stop_flag = threading.Event()
        self.thread = None
        self.smi = None
        self.marks = []
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = gpu_index
            if vis and all(x.strip().isdigit() for x in vis.split(",")):
                idx = int(vis.split(",")[gpu_index])
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
            try:
                q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                     "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                     "clocks_event_reasons.sw_power_cap")
                self.smi = subprocess.Popen(
                    ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms",
                     str(max(20, int(period_s * 1000))), "-i", str(gpu_index)],
                    stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            except Exception:
                self.smi = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.pw.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for b, name in self.NAMES.items():
                    if bits & b:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def mark(self, name):
        """Remember how many samples existed at this point (to count samples per window)."""
        self.marks.append((name, len(self.sm)))

    def stop(self):
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            sm = self.sm
            out = {"sm_mhz": statistics.median(sm) if sm else None, "sm_mhz_min": min(sm) if sm else None,
                   "sm_max_mhz": self.mx, "power_w_max": max(self.pw) if self.pw else None,
                   "power_w_median": statistics.median(self.pw) if self.pw else None,
                   "samples": len(sm), "period_ms": 1e3 * self.period, "source": "NVML polling thread",
                   "reasons": sorted(self.reasons)}
            if self.marks:
                out["samples_at"] = {k: v for k, v in self.marks}
            return out
        if self.smi is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no NVML, no nvidia-smi"]}
        time.sleep(0.15)
        self.smi.terminate()
        try:
            out, _ = self.smi.communicate(timeout=5)
        except Exception:
            self.smi.kill()
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
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "source": "nvidia-smi -lms",
                "reasons": sorted(reasons)}


def synth_quads_mt(o, n, seed, dist, dtype, threads):
    """The oracle's generator over [0, n) on `threads` host threads (ctypes releases the GIL)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    s = np.empty((n, 8), dtype=dtype)
    t = np.empty((n, 8), dtype=dtype)
    step = max(1 << 16, (n + threads - 1) // threads)

    def part(b):
        c = min(step, n - b)
        ps, pt = o.synth_quads(b, c, seed, dist, dtype)
        s[b:b + c] = ps
        t[b:b + c] = pt

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(part, range(0, n, step)))
    return s, t


def run_reference(args):
    """Reference arm: the reference's own C++ CPU path on the host cores, every step one pass over
    the same 2^log2n quadruples our arm solves per GPU."""
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
    log2n = args.log2n if args.log2n is not None else log2n
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
    ref_log2n = args.ref_log2n if args.ref_log2n is not None else log2n
    S = 1 << ref_log2n
    t0 = time.perf_counter()
    s, t = synth_quads_mt(o, S, args.seed, dist, dtype, max(1, threads))
    log(f"reference arm: generated 2^{ref_log2n} quadruples in {time.perf_counter() - t0:.1f} s")
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
        "config": stream_config(solver, dt, log2n, "aos", True, False, 1 << log2n, bytes_per_h, args.seed, dist),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"every step = one pass over 2^{ref_log2n} distinct quadruples streamed from "
                                   f"memory ({'the full per-GPU workload' if ref_log2n == log2n else 'a bounded sample'}), "
                                   f"{args.steps} steps, {threads} host threads",
                         "compiler": "g++ -O2 -ffp-contract=off (MOD/ACA_SKS.cpp compiled in place)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# --------------------------------------------------------------------------------------------
class Ctx:
    """What every measurement needs: torch, the product API, rank layout."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback "
                             "(use --impl reference for the CPU arm)")
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        from sks_homography_b200 import api, lib
        from sks_homography_b200 import dist as sdist
        self.api, self.L, self.sd = api, lib(), sdist
        self.args = args

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def host_wait_for_rank0(self, tag: str, work=None):
        """Rank 0 runs `work()` while the other ranks block on the HOST (the rendezvous store), so that
        their GPUs are idle -- an NCCL barrier would park a spinning kernel on every waiting GPU."""
        self.barrier()
        out = None
        if self.world == 1:
            return work() if work else None
        store = self.dist.distributed_c10d._get_default_store()
        if self.rank == 0:
            try:
                out = work() if work else None
            finally:
                store.set(tag, "done")
        else:
            store.wait([tag])
        self.barrier()
        return out

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def timed_steps(cx, step, steps, warmup):
    """W warm-up steps, then EXACTLY `steps` steps between barrier + synchronize on both sides,
    CUDA events per launch; returns (ms_per_step as max over ranks, per-launch ms of this rank,
    launches counted by the library)."""
    torch = cx.torch
    for _ in range(warmup):
        step()
    cx.barrier()
    cx.L.c.sks_cuda_reset_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        step()
        ev[i + 1].record()
    torch.cuda.synchronize()
    launches = int(cx.L.c.sks_cuda_launch_count())
    total_ms = ev[0].elapsed_time(ev[-1])
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    cx.barrier()
    return cx.max_over_ranks(total_ms) / steps, per, launches


def measure_stream(cx, workload, log2n=None, strong=False, layout="aos", normalize=True, steps=20, warmup=3,
                   sustained_s=0.0, sampler=None, keep=False):
    """One streaming-solver workload with inputs resident in HBM (generated on the device)."""
    torch, api, L = cx.torch, cx.api, cx.L
    solver, dt, bytes_per_h, dlog2n, dist_id = WORKLOADS[workload]
    log2n = dlog2n if log2n is None else log2n
    if strong:
        begin, n = L.shard_range(1 << log2n, cx.rank, cx.world)      # contiguous shard of a fixed total
    else:
        n = 1 << log2n
        begin = cx.rank * n                                          # this rank's shard of the global index space
    tdt = torch.float32 if dt == "f32" else torch.float64
    src, tar = api.synth_quads(n, seed=cx.args.seed, dist=dist_id, dtype=tdt, device=cx.dev, begin=begin, layout=layout)
    H = torch.empty((n, 9) if layout == "aos" else (9, n), dtype=tdt, device=cx.dev)

    def step():
        if solver == "rect":
            api.aca_rect(tar, 128.0, 1.0, 15.0, 12.0, result=H, normalize=normalize, layout=layout)
        else:
            api.solve(solver, src, tar, result=H, normalize=normalize, layout=layout)

    sustained = None
    if sustained_s > 0:
        # the same launch back to back for >= sustained_s seconds: does the burst figure hold, and
        # what do clock and power do meanwhile (the sampler runs through this window)
        for _ in range(warmup):
            step()
        cx.barrier()
        if sampler:
            sampler.mark("sustained_begin")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        k = 0
        while time.perf_counter() - t0 < sustained_s or k < 200:
            for _ in range(100):
                step()
            k += 100
        e1.record()
        torch.cuda.synchronize()
        if sampler:
            sampler.mark("sustained_end")
        ms = e0.elapsed_time(e1) / k
        sustained = {"launches": k, "seconds": e0.elapsed_time(e1) / 1e3, "ms_per_step": ms,
                     "GBps": n * bytes_per_h / ms / 1e6, "GHps_per_gpu": n / ms / 1e6}
    ms_per_step, per, launches = timed_steps(cx, step, steps, warmup)
    if sampler:
        sampler.mark("timed_end")
    n_total = (1 << log2n) if strong else cx.world * n
    peak, peak_src = peaks()
    mean_ms = statistics.mean(per)
    achieved = n * bytes_per_h / (mean_ms * 1e-3) / 1e9
    res = {
        "value": n_total / (ms_per_step * 1e-3), "ms_per_step": ms_per_step, "launches": launches, "n": n,
        "n_total": n_total, "begin": begin, "log2n": log2n, "dt": dt, "solver": solver, "dist": dist_id,
        "bytes_per_h": bytes_per_h, "strong": strong, "layout": layout, "normalize": normalize,
        "sustained": sustained,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(workload), "peak_source": peak_src, "bytes_per_homography": bytes_per_h,
                     "launch_ms_mean": mean_ms, "launch_ms_min": min(per), "frac_of_nominal_8TBs": achieved / 8000.0},
    }
    if sustained:
        res["roofline"]["sustained_frac"] = sustained["GBps"] / peak
    if keep:
        res["tensors"] = (src, tar, H)
    return res


def link_probe(cx, mib=1024):
    """PCIe ceiling of this box, measured here: one large pinned H2D copy and one D2H copy, alone."""
    torch = cx.torch
    n = mib << 20
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device=cx.dev)
    out = {}
    for name, (dst, src) in (("h2d", (d, h)), ("d2h", (h, d))):
        best = 1e30
        for _ in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); dst.copy_(src, non_blocking=True); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[name + "_GBps"] = n / best / 1e6
    del h, d
    return out


def measure_e2e(cx, m, src, tar, H):
    """The same workload through the host-pointer C ABI: host buffers in, host buffers out, H2D and
    D2H copies inside the timed region -- from pinned buffers and from plain pageable arrays."""
    torch, api, args = cx.torch, cx.api, cx.args
    solver, dt, n = m["solver"], m["dt"], m["n"]
    tdt = torch.float32 if dt == "f32" else torch.float64
    ne = min(n, 1 << (args.e2e_log2n if args.e2e_log2n is not None else m["log2n"]))
    esz = 4 if dt == "f32" else 8
    in_elems = 8 if solver == "rect" else 16
    link = link_probe(cx)
    ceiling = link["h2d_GBps"] * 1e9 / (in_elems * esz)
    out = {"unit": UNIT, "h2d_bytes_per_step": ne * in_elems * esz, "d2h_bytes_per_step": ne * 9 * esz,
           "quadruples_per_step_per_gpu": ne,
           "api": "sks_host_* (host buffers, chunked H2D/kernel/D2H ring)",
           "link": {**link, "ceiling_GHps_per_gpu": ceiling / 1e9,
                    "note": f"every homography moves {in_elems * esz} B host->device and {9 * esz} B back; the H2D "
                            "direction alone, measured above with one large pinned copy, bounds the rate per GPU "
                            "whatever the kernel does"},
           "numa_binding": getattr(args, "numa", None)}

    def run_host(hs, ht, hH, steps):
        def host_step():
            if solver == "rect":
                api.aca_rect(ht, 128.0, 1.0, 15.0, 12.0, result=hH, normalize=m["normalize"])
            else:
                api.solve(solver, hs, ht, result=hH, normalize=m["normalize"])
        host_step()                                   # warm-up: ring buffers, streams
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            host_step()                               # synchronous: returns with H in host memory
        el = cx.max_over_ranks(time.perf_counter() - t0)
        return cx.world * ne * steps / el

    view = torch.int32 if dt == "f32" else torch.int64
    want = H[:ne].cpu()
    for kind, steps in (("pinned", args.e2e_steps), ("pageable", max(2, args.e2e_steps // 3))):
        pin = kind == "pinned"
        hs = torch.empty((ne, 8), dtype=tdt, pin_memory=pin) if solver != "rect" else None
        ht = torch.empty((ne, 8), dtype=tdt, pin_memory=pin)
        hH = torch.empty((ne, 9), dtype=tdt, pin_memory=pin)
        if hs is not None:
            hs.copy_(src[:ne])
        ht.copy_(tar[:ne])
        torch.cuda.synchronize()
        v = run_host(hs, ht, hH, steps)
        same = torch.equal(hH.view(view), want.view(view))
        if not same:
            same = "nan-pattern-equal" if torch.equal(torch.isnan(hH), torch.isnan(want)) and torch.equal(
                torch.nan_to_num(hH).view(view), torch.nan_to_num(want).view(view)) else "MISMATCH"
        else:
            same = "bit-exact"
        if pin:
            out.update(value=v, steps=steps, buffers="pinned (cudaHostAlloc)", parity_vs_device_path=same,
                       frac_of_link_ceiling=v / cx.world / ceiling)
        else:
            out["pageable"] = {"value": v, "unit": UNIT, "steps": steps, "parity_vs_device_path": same,
                               "buffers": "plain malloc'd arrays, as the reference's CPU/main.cpp:47-58 holds them; "
                                          "staged through the library's pinned ring by host memcpy threads"}
        del hs, ht, hH
    return out


def cpu_baseline_stream(cx, m, src, tar, H):
    """The reference's own C++ on this box's host cores, bounded sample, + bit parity of the GPU results."""
    import numpy as np
    from oracle.oracle import Oracle, RefLib
    args = cx.args
    solver, dt = m["solver"], m["dt"]
    S = min(m["n"], 1 << args.cpu_log2n)
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
    parity = None
    if m["normalize"]:
        got = H[:S].cpu().numpy()
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
    src_name = 'MOD/GE.cpp' if solver == 'ge' else 'GPU.cu:242-357' if solver == 'gpt' else 'MOD/ACA_SKS.cpp'
    return {"value": S / best, "unit": UNIT, "cores": threads, "kind": kind, "o3_fma_build": o3,
            "sample": f"first 2^{args.cpu_log2n} quadruples of the workload, best of 5 passes, {src_name} "
                      f"g++ -O2 -ffp-contract=off{'' if kind == 'reference' else ' (oracle port)'}, {threads} threads",
            "parity_gpu_vs_cpu": parity}


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


def reference_torch_eager(api, dev, log2ns=(6, 10, 14, 20, 22, 24)):
    """The reference's existing fp32 GPU implementation for configs[1]/[2]: its torch-eager
    functions TensorACA_rect (PY.py:286-309, math :296-302) and ACA_vanilla (:312-388, math
    :322-381), EXECUTED UNMODIFIED on cuda tensors (statements staged from the reference checkout
    at build time, oracle/_ref/ref_torch_funcs.py), next to sks_cuda_aca_rect_planar_f32 /
    sks_cuda_aca_f32 on the SAME tensors.  Inputs come from the reference's own generator (adjust)."""
    import torch
    try:
        from oracle.oracle import ref_torch_funcs
        R = ref_torch_funcs()
    except Exception as e:
        return {"unavailable": str(e)[:160]}
    out = {"what": "reference torch eager (Modules_Runtime_Test.py:296-302, :322-381) vs libsks_cuda on the same "
                   "cuda tensors; median of 10 calls (40 below 2^16: the deep-homography batch sizes, where both sides are "
                   "launch-bound -- ~15 / ~60 eager launches against one), CUDA events",
           "rows": []}
    for log2n in log2ns:
        bs = 1 << log2n
        try:
            torch.manual_seed(11)
            src, tar, src_new, tar_new, scale, div = R.adjust(dev, bs)
            # --- TensorACA_rect: [bs,3,4] homogeneous tensors, up to scale -----------------------
            H_ref = R.TensorACA_rect_body(bs, src_new, tar_new, scale, div)
            iters = 10 if log2n >= 16 else 40
            t_ref = _event_ms(lambda: R.TensorACA_rect_body(bs, src_new, tar_new, scale, div), iters)
            H_our = api.TensorACA_rect(bs, src_new, tar_new, scale, div)
            t_our = _event_ms(lambda: api.TensorACA_rect(bs, src_new, tar_new, scale, div), iters)
            same_rect = bool(torch.equal(H_ref.reshape(bs, 9).view(torch.int32), H_our.reshape(bs, 9).view(torch.int32)))
            graph_us = None
            if log2n < 16:      # launch-bound sizes: the same call captured once in a CUDA graph and replayed
                try:
                    torch.cuda.synchronize()
                    gr = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gr):
                        api.TensorACA_rect(bs, src_new, tar_new, scale, div)
                    graph_us = 1e3 * _event_ms(gr.replay, iters)
                    del gr
                except Exception as e:
                    graph_us = f"{type(e).__name__}: {e}"[:80]
            out["rows"].append({"fn": "TensorACA_rect", "bs": bs, "reference_us": 1e3 * t_ref, "ours_us": 1e3 * t_our,
                                "ours_cuda_graph_replay_us": graph_us,
                                "speedup": t_ref / t_our, "reference_GHps": bs / t_ref / 1e6, "ours_GHps": bs / t_our / 1e6,
                                "ours_GBps_tensor_layout": bs * (48 + 36) / t_our / 1e6, "bit_identical": same_rect})
            del H_ref, H_our
            # --- ACA_vanilla: [bs,4,2] AoS tensors, up to scale --------------------------------------
            H_ref = R.ACA_vanilla_body(bs, src, tar)
            t_ref = _event_ms(lambda: R.ACA_vanilla_body(bs, src, tar), iters)
            H_our = api.ACA_vanilla(bs, src, tar)
            t_our = _event_ms(lambda: api.ACA_vanilla(bs, src, tar), iters)
            same_van = bool(torch.equal(H_ref.reshape(bs, 9).view(torch.int32), H_our.reshape(bs, 9).view(torch.int32)))
            out["rows"].append({"fn": "ACA_vanilla", "bs": bs, "reference_us": 1e3 * t_ref, "ours_us": 1e3 * t_our,
                                "speedup": t_ref / t_our, "reference_GHps": bs / t_ref / 1e6, "ours_GHps": bs / t_our / 1e6,
                                "ours_GBps": bs * 100 / t_our / 1e6, "bit_identical": same_van})
            del H_ref, H_our, src, tar, src_new, tar_new
            torch.cuda.empty_cache()
        except Exception as e:                              # e.g. the eager graph runs out of memory at 2^24
            out["rows"].append({"bs": bs, "error": f"{type(e).__name__}: {e}"[:160]})
            torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------
def measure_ransac(cx, steps, warmup, shard="hypotheses", peer_reduce=False, breakdown=True, do_e2e=True, do_cpu=True):
    """BASELINE configs[4]: fused ACA-RANSAC, P pairs x n_pts matches x n_hyp
    hypotheses.  Multi-GPU variant (A) of SURVEY.md 8(e): every rank holds all
    pairs' matches and scores its contiguous shard of the hypothesis ids; ONE
    int64 max-all-reduce (NCCL over NVLink) merges the winners; the winning model
    is recomputed locally.  Strong scaling: total work is fixed."""
    import numpy as np
    torch, api, L, sd, dist, args = cx.torch, cx.api, cx.L, cx.sd, cx.dist, cx.args
    rank, world, dev = cx.rank, cx.world, cx.dev
    P, n_pts, n_hyp = args.pairs, args.points, args.hyps
    thr2 = 2.25
    by_pairs = shard == "pairs" and world > 1
    if by_pairs:
        # SURVEY.md 8(e) variant B: this rank owns pairs [pb, pb + pc) and scores every hypothesis
        # id of them; nothing is exchanged (the throughput-optimal split for many pairs)
        pb, pc = sd.shard_range(P, rank, world)
        corr = api.synth_corr(pc, n_pts, seed=args.seed, inlier_permille=500, noise=0.5, device=dev, pair_begin=pb)
        hb, hc = 0, n_hyp
    else:
        pb, pc = 0, P
        corr = api.synth_corr(P, n_pts, seed=args.seed, inlier_permille=500, noise=0.5, device=dev)
        hb, hc = sd.shard_range(n_hyp, rank, world)
    keys = torch.zeros(pc, dtype=torch.int64, device=dev)
    res = {}
    reducer = sd.PeerReducer(P, dev) if (peer_reduce and not by_pairs) else None
    seg = []      # per step: events around zero | score | reduce | finalize

    def step(record=False):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if record else None
        if ev: ev[0].record()
        keys.zero_()
        if ev: ev[1].record()
        api.ransac_keys(corr, n_hyp, args.seed, thr2, None, hb, hc, out=keys, pair_begin=pb)
        if ev: ev[2].record()
        if by_pairs:
            pass
        elif reducer is not None:
            reducer.max_reduce_(keys)
        else:
            sd.merge_keys(keys)
        if ev: ev[3].record()
        res["H"], res["cnt"], _ = api.ransac_finalize(corr, n_hyp, args.seed, thr2, keys, pair_begin=pb)
        if ev:
            ev[4].record()
            seg.append(ev)

    sampler = ClockSampler(cx.local) if rank == 0 else None      # SM clock / power DURING the scoring steps
    ms_per_step, per, launches = timed_steps(cx, step, steps, warmup)
    clocks = sampler.stop() if sampler else None
    value = P * n_hyp / (ms_per_step * 1e-3)

    parts = None
    if breakdown:
        # the same step once more with events between its four parts (a separate, untimed pass: the
        # extra event records are not part of the number above); max over ranks per part
        cx.barrier()
        for _ in range(3):
            step(record=True)
        torch.cuda.synchronize()
        names = ["zero_keys", "score_kernel", "reduce_incl_wait_for_slowest_rank", "finalize"]
        parts = {}
        for i, name in enumerate(names):
            parts[name + "_ms"] = cx.max_over_ranks(statistics.median(e[i].elapsed_time(e[i + 1]) for e in seg))
        parts["score_kernel_ms_min_over_ranks"] = -cx.max_over_ranks(
            -statistics.median(e[1].elapsed_time(e[2]) for e in seg))
        # the collective alone, ranks aligned by a barrier first: its own cost without the skew
        if world > 1 and not by_pairs:
            k2 = keys.clone()
            ts = []
            for _ in range(10):
                cx.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if reducer is not None:
                    reducer.max_reduce_(k2)
                else:
                    sd.merge_keys(k2)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            parts["reduce_alone_ms"] = cx.max_over_ranks(statistics.median(ts))
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        parts["note"] = ("strong scaling = score_kernel at 1/world of the hypothesis ids + reduce + finalize + "
                         "zero_keys; the scorer picks its rounds per CTA so that the grid fills whole waves of "
                         f"{sms} SMs x 3 CTAs (csrc/capi.cu)")

    # FP32 roofline: 21 flop per hypothesis x point (8 FMA + 1 MUL + 2 FMA once the threshold is
    # folded into the operands; SURVEY.md 8(d) counted 22 for the unfolded form) + 97 + 6 per hypothesis
    flops = P * float(n_hyp) * (103.0 + 21.0 * n_pts)
    mhz = 1965.0
    try:
        mhz = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["sm_max_mhz"])
    except Exception:
        pass
    peak = 148 * 128 * 2 * mhz * 1e6 / 1e12 * world
    achieved = flops / (ms_per_step * 1e-3) / 1e12
    roofline = {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": None,
                "peak_source": f"148 SM x 128 lanes x 2 x {mhz:.0f} MHz (max SM clock, MEASURED_PEAKS.json) x {world} GPU",
                "flop_per_hypothesis_point": 21, "flop_per_hypothesis": 103,
                "fma_pipe_instructions_per_hypothesis_point": 12,
                "ceiling_note": "11 scoring FMAs/MULs + 1 counting FMA (FFMA2.RM) per evaluation all run on the FMA "
                                "pipe, 10.5 of the 12 are counted flops: 0.875 is the most this metric can show; ncu "
                                "has the pipe 93 % busy (profiles/r02_ransac_fpcount.txt)"}

    # parity at full size: a few pairs against the CPU oracle on a hypothesis prefix,
    # and (multi-GPU) merged keys == single-GPU keys on the same pairs
    parity = None
    if rank == 0 and do_cpu:
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
                  "sharded_equals_unsharded": full1,
                  "note": "scoring rule is this project's definition (parity unpinned by construction); "
                          "hypotheses are the reference's ACA, bit-exact"}
    # e2e: the same step with the matches in pinned HOST memory -- H2D of every pair's matches,
    # scoring, winner merge, finalize, D2H of the models and inlier counts inside the timed region
    e2e = None
    if do_e2e:
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
        cx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            e2e_step()
        e1.record()
        torch.cuda.synchronize()
        e2e_ms = cx.max_over_ranks(e0.elapsed_time(e1)) / 3
        e2e = {"value": P * n_hyp / (e2e_ms * 1e-3), "unit": "hypotheses/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": corr.numel() * 4, "d2h_bytes_per_step": pc * 9 * 4 + pc * 4,
               "api": "sks_host_ransac_aca_f32 (host matches in, models out)" if world == 1 and reducer is None
                      else "per-rank H2D + sks_cuda_ransac_aca_f32 on the rank's hypothesis shard + max-reduce + finalize",
               "models_equal_device_path": bool(torch.equal(h_H, res["H"].cpu()))}
        del h_corr, d_corr

    # CPU baseline: the oracle's scalar port of the same definition on one host core, on a bounded
    # sample (one pair, a prefix of its hypothesis ids)
    cpu = None
    if rank == 0 and do_cpu:
        from oracle.oracle import Oracle
        o = Oracle()
        sub = corr[:1].cpu().numpy()
        nh = min(n_hyp, max(1024, int(args.ransac_cpu_evals // max(n_pts, 1))))
        t0 = time.perf_counter()
        o.ransac(sub, nh, args.seed, thr2)
        dt = time.perf_counter() - t0
        cpu = {"value": nh / dt, "unit": "hypotheses/s", "cores": 1, "kind": "port",
               "sample": f"1 pair x {n_pts} matches x {nh} hypotheses, oracle/sks_oracle.c "
                         f"(gcc -O2 -ffp-contract=off), {dt:.1f} s"}
    out = {
        "metric": "ACA-RANSAC hypotheses/s (homographies solved and scored)", "value": value, "unit": "hypotheses/s",
        "ms_per_step": ms_per_step, "steps": steps, "scaling": "strong", "dtype": "f32",
        "config": {"workload": f"fused ACA-RANSAC {P} pairs x {n_pts} matches x {n_hyp} hypotheses, "
                               + (f"image pairs sharded over {world} GPU(s), no collective" if by_pairs else
                                  f"hypotheses sharded over {world} GPU(s), "
                                  + ("winners merged by NVLink peer atomics (csrc/peer.cuh)" if reducer
                                     else "one int64 max all-reduce (NCCL)")),
                   "thr2": thr2, "seed": args.seed,
                   "l2": "compute-bound; matches (64 KiB/pair) live in shared memory"},
        "roofline": roofline, "breakdown": parts, "parity": parity, "e2e": e2e, "cpu_baseline": cpu,
        "gpu_launches": launches, "clocks": clocks,
        "mean_inlier_fraction_of_winner": float(res["cnt"].float().mean().item()) / n_pts,
        "peer_reduce_timed_out": reducer.timed_out() if reducer else None,
    }
    if reducer is not None:
        reducer.close()
    del corr, keys
    torch.cuda.empty_cache()
    return out


def measure_ransac_inprocess(cx, steps=5):
    """The single-process multi-GPU driver behind the C ABI (sks_cuda_ransac_aca_multi_f32, csrc/multi.cu):
    rank 0 alone drives `world` GPUs -- matches on its own device, broadcast to the others over NVLink,
    winners merged with peer atomics -- while the other ranks block on the host with their GPUs idle."""
    torch, api, args = cx.torch, cx.api, cx.args

    def work():
        try:
            ngpu = min(cx.world, torch.cuda.device_count())
            P, n_pts, n_hyp, thr2 = args.pairs, args.points, args.hyps, 2.25
            corr = api.synth_corr(P, n_pts, seed=args.seed, inlier_permille=500, noise=0.5, device=cx.dev)
            H1, c1, _, k1 = api.ransac_multi(corr, n_hyp, args.seed, thr2, ngpu=1)

            def step():
                return api.ransac_multi(corr, n_hyp, args.seed, thr2, ngpu=ngpu)
            for _ in range(2):
                Hn, cn, _, kn = step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            return {"api": "sks_cuda_ransac_aca_multi_f32 (one process, one enqueueing thread, matches broadcast by a "
                           "binomial tree of NVLink peer copies, winners merged by peer atomics, CUDA-event ordering; "
                           "no NCCL, no IPC)", "gpus": ngpu, "steps": steps,
                    "ms_per_step": ms, "value": P * n_hyp / (ms * 1e-3), "unit": "hypotheses/s",
                    "other_ranks": "blocked on the host (rendezvous store), GPUs idle",
                    "bit_identical_to_one_gpu": bool(torch.equal(kn, k1) and torch.equal(cn, c1) and
                                                     torch.equal(Hn.view(torch.int32), H1.view(torch.int32)))}
        except Exception as e:
            return {"error": f"{type(e).__name__}: {e}"[:200]}

    return cx.host_wait_for_rank0("ransac_inprocess_multi_done", work)


def accuracy_tier_f64(cx, m, src, tar, H):
    """configs[3] accuracy tier on rank 0: SKS fp64 bit-compared with the reference's runKernel_SKS_double,
    ACA_double == SKS_double to rounding, 4-point reprojection error."""
    import numpy as np
    torch, api = cx.torch, cx.api
    S = min(m["n"], 1 << 20)
    out = {"checked_quadruples": S}
    try:
        from oracle.oracle import RefLib
        ref = RefLib()
        s_h, t_h = src[:S].cpu().numpy(), tar[:S].cpu().numpy()
        want = ref.solve("sks", s_h, t_h, threads=ref.hardware_threads())
        got = H[:S].cpu().numpy()
        ok = (got.view(np.uint64) == want.view(np.uint64)) | (np.isnan(got) & np.isnan(want))
        out["mismatching_elements_vs_runKernel_SKS_double"] = int((~ok).sum())
    except Exception as e:
        out["reference_unavailable"] = str(e)[:100]
    Ha = api.solve("aca", src[:S], tar[:S])
    Hs = H[:S]
    scale = Hs.abs().amax(dim=1, keepdim=True)
    d = ((Ha - Hs).abs() / scale)
    fin = torch.isfinite(d).all(dim=1)
    out["aca64_vs_sks64_rel_to_matrix_scale"] = {"median": float(d[fin].amax(dim=1).median()),
                                                 "max": float(d[fin].amax(dim=1).max())}
    s4 = src[:S].view(S, 4, 2)
    t4 = tar[:S].view(S, 4, 2)
    Hm = Hs.view(S, 3, 3)
    p = torch.cat([s4, torch.ones(S, 4, 1, dtype=s4.dtype, device=s4.device)], dim=2) @ Hm.transpose(1, 2)
    err = ((p[..., :2] / p[..., 2:3]) - t4).norm(dim=2).amax(dim=1)
    err = err[torch.isfinite(err)]
    out["reprojection_px"] = {"median": float(err.median()), "p99": float(err.kthvalue(int(0.99 * err.numel())).values)}
    return out


# --------------------------------------------------------------------------------------------
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
    ap.add_argument("--packed", type=int, default=-1, help="RANSAC scorer (0 scalar | 1 FFMA2 | 2 hyp pairs | 3 FP count)")
    ap.add_argument("--ransac-shard", default="hypotheses", choices=["hypotheses", "pairs"],
                    help="multi-GPU RANSAC: shard the hypothesis ids of every pair (one max all-reduce; "
                         "north_star's variant) or the image pairs (no collective at all)")
    ap.add_argument("--peer-reduce", action="store_true",
                    help="RANSAC: merge the winners with the hand-written NVLink peer max-reduce "
                         "instead of the NCCL all-reduce")
    ap.add_argument("--ransac-steps", type=int, default=5)
    ap.add_argument("--ransac-cpu-evals", type=float, default=2.5e8,
                    help="size of the CPU-oracle RANSAC baseline sample (hypothesis x match evaluations)")
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
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--e2e-log2n", type=int, default=None, help="default: the workload's own size")
    ap.add_argument("--cpu-log2n", type=int, default=24)
    ap.add_argument("--ref-log2n", type=int, default=None, help="reference arm: quadruples per step (default: --log2n)")
    ap.add_argument("--sustained-s", type=float, default=2.5, help="length of the sustained leg (0 = skip)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the `configs` legs (other BASELINE configs)")
    ap.add_argument("--legs", default="ransac,rect_2p28_strong,sks_f64_2p25,ransac_inprocess_multi")
    ap.add_argument("--force-legs", action="store_true", help="run the legs although the top level is not the headline "
                                                              "workload (tests: small sizes)")
    ap.add_argument("--leg-log2n", type=int, default=None, help="shrink the rect / SKS-f64 legs (tests only)")
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="multi-rank runs: do not pin the rank to its GPU's NUMA node")
    args = ap.parse_args()
    _capture_stdout()
    if args.warmup < 3:
        args.warmup = max(args.warmup, 3)          # timing rules: W >= 3
    if args.impl == "reference":
        return run_reference(args)

    cx = Ctx(args)
    torch, dist, L, api = cx.torch, cx.dist, cx.L, cx.api
    rank, world, local = cx.rank, cx.world, cx.local
    # one process per GPU: keep the rank's host threads and pinned buffers on its GPU's socket
    args.numa = cx.sd.bind_to_gpu_numa_node(local) if (world > 1 and not args.no_numa_bind) else {"bound": False}
    L.check(L.c.sks_cuda_set_variant(args.variant), "set_variant")
    L.check(L.c.sks_cuda_set_tuning(args.small_tile, args.stages, args.ctas), "set_tuning")
    if args.hpt or args.packed >= 0:
        L.check(L.c.sks_cuda_set_ransac_tuning(args.hpt or 2, args.rounds, args.packed if args.packed >= 0 else 3),
                "set_ransac_tuning")

    if args.workload == "ransac":
        r = measure_ransac(cx, args.steps, args.warmup, args.ransac_shard, args.peer_reduce,
                           do_e2e=not args.no_e2e, do_cpu=not args.no_cpu)
        if rank == 0:
            line = {"impl": "ours", "n_gpus": world, "warmup": args.warmup, "higher_is_better": True,
                    "vs_baseline": None, "data": "synthetic", **r}
            emit(line)
        if world > 1:
            dist.destroy_process_group()
        return 0

    headline = args.workload == "aca_f32" and args.log2n is None and not args.strong and args.layout == "aos" \
        and not args.no_normalize
    sampler = ClockSampler(local) if rank == 0 else None
    m = measure_stream(cx, args.workload, args.log2n, args.strong, args.layout, not args.no_normalize, args.steps,
                       args.warmup, sustained_s=args.sustained_s, sampler=sampler, keep=True)
    clocks = sampler.stop() if sampler else None
    src, tar, H = m.pop("tensors")
    log(f"{args.workload}: {m['value'] / 1e9:.2f} G H/s, {m['roofline']['achieved']:.0f} GB/s")

    e2e = None
    if not args.no_e2e and args.layout == "aos":
        e2e = measure_e2e(cx, m, src, tar, H)
        log(f"e2e pinned {e2e['value'] / 1e9:.3f} G H/s, pageable {e2e['pageable']['value'] / 1e9:.3f} G H/s, "
            f"link ceiling {e2e['link']['ceiling_GHps_per_gpu']:.3f} G H/s per GPU")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and args.layout == "aos" and m["solver"] != "rect":
        cpu = cpu_baseline_stream(cx, m, src, tar, H)
    del src, tar, H
    torch.cuda.empty_cache()

    gpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu and not args.no_gpu_baseline:
        gpu_base = reference_gpu_kernels(api, cx.dev)
        gpu_base["torch_eager_fp32"] = reference_torch_eager(api, cx.dev)
        torch.cuda.empty_cache()

    # ---- the other BASELINE configs, as short legs in the same process group --------------------
    legs = None
    if (headline or args.force_legs) and not args.no_legs:
        want = [x for x in args.legs.split(",") if x]
        legs = {}
        if "ransac" in want:
            legs["ransac"] = measure_ransac(cx, max(5, args.ransac_steps), 3, "hypotheses", False,
                                            do_e2e=False, do_cpu=not args.no_cpu)
            log(f"ransac leg: {legs['ransac']['ms_per_step']:.2f} ms/step, frac {legs['ransac']['roofline']['frac']:.3f}")
        if "rect_2p28_strong" in want:
            r = measure_stream(cx, "rect_f32", args.leg_log2n or 28, True, steps=max(5, args.steps), warmup=3)
            legs["rect_2p28_strong"] = {
                "config": "BASELINE configs[2]: ACA-rect fp32, 2^28 quadruples in TOTAL, shared source rectangle, "
                          f"contiguous shards over {world} GPU(s), no collective",
                "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "scaling": "strong",
                "quadruples_per_gpu": r["n"], "roofline": r["roofline"], "gpu_launches": r["launches"]}
            torch.cuda.empty_cache()
            log(f"rect 2^28 strong: {r['value'] / 1e9:.1f} G H/s")
        if "sks_f64_2p25" in want:
            r = measure_stream(cx, "sks_f64", args.leg_log2n or 25, False, steps=max(5, args.steps), warmup=3, keep=True)
            s2, t2, H2 = r.pop("tensors")
            legs["sks_f64_2p25"] = {
                "config": f"BASELINE configs[3]: SKS fp64, 2^25 quadruples per GPU on {world} GPU(s), AoS, h33-normalised",
                "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "scaling": "weak",
                "roofline": r["roofline"], "gpu_launches": r["launches"],
                "accuracy_tier": accuracy_tier_f64(cx, r, s2, t2, H2) if rank == 0 else None}
            del s2, t2, H2
            torch.cuda.empty_cache()
            log(f"sks f64 2^25: {r['value'] / 1e9:.1f} G H/s")
        if "ransac_inprocess_multi" in want and world > 1:
            legs["ransac_inprocess_multi"] = measure_ransac_inprocess(cx)

    if rank == 0:
        line = {
            "impl": "ours", "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_per_step"],
            "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
            "dtype": m["dt"], "data": "synthetic",
            "config": stream_config(m["solver"], m["dt"], m["log2n"], args.layout, m["normalize"], args.strong, m["n"],
                                    m["bytes_per_h"], args.seed, m["dist"]),
            "roofline": m["roofline"], "sustained": m["sustained"], "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": m["launches"], "clocks": clocks, "gpu_baseline": gpu_base, "configs": legs,
            "kernel_variant": args.variant,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
