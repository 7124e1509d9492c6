"""Regenerates tests/golden/*.npz from the REFERENCE ITSELF (run in the authoring
container, where /root/reference exists; the fixtures are committed because the
reference cannot travel to the GPU box).

  ref_general.npz  outputs of the reference's own C++ (MOD/ACA_SKS.cpp compiled in
                   place -> oracle/_ref/libsks_ref.so) on seeded quadruples, fp32 and
                   fp64, three input distributions + exact degenerate cases.
  ref_ge.npz       outputs of the reference's competitor solver cv::runKernel_GE (MOD/GE.cpp,
                   same library, fp32 only) on the image-uniform distribution, the
                   veri_4Pts.m general quad and the degenerate cases.
  ref_torch.npz    outputs obtained by EXECUTING the reference's torch statements
                   (PY.py getInput/getTar/adjust, the body of TensorACA_rect
                   :296-302 and of ACA_vanilla :322-381) on torch-CPU.  The source
                   text is read from /root/reference at generation time and
                   exec'd; nothing is copied into this repository.
  kat_veri4pts.npz the two known-answer cases of ML/veri_4Pts.m (camera of
                   :28-46, general quad :9-12, rectangle :82-91), H_real computed
                   here in float64 from those parameters.

usage: python tests/golden/make_golden.py
"""
from __future__ import annotations

import ast
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))
PY_REF = "/root/reference/PyTorch Codes/Modules_Runtime_Test.py"


def degenerate_cases(dtype):
    """Small-integer quadruples with exactly collinear / repeated points (exact in
    fp32, so the flags do not depend on rounding) plus healthy ones."""
    good = [0, 0, 4, 0, 0, 3, 5, 4]
    good2 = [1, 1, 6, 2, 0, 7, 8, 9]
    cases = [
        ([0, 0, 1, 0, 2, 0, 1, 1], [0, 0, 1, 0, 2, 0, 1, 1]),   # M,N,P collinear, both planes
        ([0, 0, 1, 0, 2, 0, 1, 1], good),                       # source only
        (good, [0, 0, 1, 0, 2, 0, 1, 1]),                       # target only
        ([0, 0, 2, 2, 1, 3, 4, 4], good),                       # M,N,Q collinear
        ([0, 0, 3, 1, 1, 2, 3, 6], good),                       # M,P,Q collinear
        ([5, 5, 1, 0, 2, 3, 4, 4], good),                       # N,P,Q collinear? (no: healthy)
        ([0, 0, 0, 0, 0, 3, 5, 4], good),                       # M == N
        ([0, 0, 4, 0, 0, 3, 0, 3], good),                       # P == Q
        ([2, 2, 2, 2, 2, 2, 2, 2], good),                       # all equal
        (good, good), (good, good2), (good2, good),             # healthy
        ([0, 0, 1, 0, 0, 1, 1, 1], [0, 0, 1, 0, 0, 1, 1, 1]),   # unit square -> identity
    ]
    s = np.array([c[0] for c in cases], dtype=dtype)
    t = np.array([c[1] for c in cases], dtype=dtype)
    return s, t


def make_ref_general():
    from oracle.oracle import Oracle, RefLib
    o, r = Oracle(), RefLib()
    out = {}
    for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
        for dist in (0, 1, 2):
            s, t = o.synth_quads(1000 * dist, 768, 11 + dist, dist, dt)
            out[f"src_{tag}_d{dist}"], out[f"tar_{tag}_d{dist}"] = s, t
            out[f"aca_{tag}_d{dist}"] = r.solve("aca", s, t)
            out[f"sks_{tag}_d{dist}"] = r.solve("sks", s, t)
        s, t = degenerate_cases(dt)
        out[f"src_{tag}_deg"], out[f"tar_{tag}_deg"] = s, t
        out[f"aca_{tag}_deg"] = r.solve("aca", s, t)
        out[f"sks_{tag}_deg"] = r.solve("sks", s, t)
    np.savez_compressed(os.path.join(OUT, "ref_general.npz"), **out)
    print("ref_general.npz", {k: v.shape for k, v in list(out.items())[:4]}, "...")


def make_ref_ge():
    from oracle.oracle import Oracle, RefLib
    o, r = Oracle(), RefLib()
    out = {}
    for dist in (0, 1):      # dist 0: axis-aligned source squares, GE's first pivot is zero
        s, t = o.synth_quads(4000 * dist, 768, 21 + dist, dist, np.float32)
        out[f"src_f32_d{dist}"], out[f"tar_f32_d{dist}"] = s, t
        out[f"ge_f32_d{dist}"] = r.solve("ge", s, t)
    s, t = degenerate_cases(np.float32)
    out["src_f32_deg"], out["tar_f32_deg"], out["ge_f32_deg"] = s, t, r.solve("ge", s, t)
    k = np.load(os.path.join(OUT, "kat_veri4pts.npz"))
    s, t = k["src_general"][None].astype(np.float32), k["tar_general"][None].astype(np.float32)
    out["src_f32_kat"], out["tar_f32_kat"], out["ge_f32_kat"] = s, t, r.solve("ge", s, t)
    np.savez_compressed(os.path.join(OUT, "ref_ge.npz"), **out)
    print("ref_ge.npz", {k: v.shape for k, v in list(out.items())[:3]}, "...")


def _function_nodes(path):
    text = open(path).read()
    tree = ast.parse(text)
    return text, {n.name: n for n in tree.body if isinstance(n, ast.FunctionDef)}


def _exec_function_defs(text, nodes, names, ns):
    for name in names:
        exec(compile(ast.Module(body=[nodes[name]], type_ignores=[]), PY_REF, "exec"), ns)


def _timed_body(text, node):
    """Statements inside the reference function's timing loop, minus the
    synchronize / clock / bookkeeping lines."""
    loop = next(n for n in node.body if isinstance(n, ast.For))
    keep = []
    for st in loop.body:
        seg = ast.get_source_segment(text, st)
        if any(w in seg for w in ("synchronize", "perf_counter", "time_list")):
            continue
        keep.append(st)
    return ast.Module(body=keep, type_ignores=[])


def make_ref_torch():
    import torch
    text, nodes = _function_nodes(PY_REF)
    ns = {"torch": torch}
    _exec_function_defs(text, nodes, ["getInput", "getTar", "adjust"], ns)
    torch.manual_seed(20251018)
    bs = 512
    src, tar, src_new, tar_new, scale, div = ns["adjust"]("cpu", bs)
    # body of TensorACA_rect (PY.py:296-302)
    env = {"torch": torch, "bs": bs, "src": src_new, "tar": tar_new, "scale": scale, "div": div}
    exec(compile(_timed_body(text, nodes["TensorACA_rect"]), PY_REF, "exec"), env)
    H_rect = env["H"].clone()
    # body of ACA_vanilla (PY.py:322-381)
    env = {"torch": torch, "bs": bs, "src": src, "tar": tar}
    exec(compile(_timed_body(text, nodes["ACA_vanilla"]), PY_REF, "exec"), env)
    H_van = env["H"].clone()
    np.savez_compressed(
        os.path.join(OUT, "ref_torch.npz"),
        src=src.numpy(), tar=tar.numpy(), src_new=src_new.numpy(), tar_new=tar_new.numpy(),
        scale=scale.numpy(), div=div.numpy(), H_rect=H_rect.numpy(), H_vanilla=H_van.numpy())
    print("ref_torch.npz", H_rect.shape, H_van.shape, float(scale), float(div))


def make_kat():
    # ML/veri_4Pts.m:28-53
    fu = fv = 900.0
    u0, v0 = 500.0, 400.0
    K = np.array([[fu, 0, u0], [0, fv, v0], [0, 0, 1]])
    rx = ry = -np.pi / 8 * np.sqrt(5)
    rz = -np.pi / 16 * np.sqrt(5)
    Rx = np.array([[1, 0, 0], [0, np.cos(rx), -np.sin(rx)], [0, np.sin(rx), np.cos(rx)]])
    Ry = np.array([[np.cos(ry), 0, np.sin(ry)], [0, 1, 0], [-np.sin(ry), 0, np.cos(ry)]])
    Rz = np.array([[np.cos(rz), -np.sin(rz), 0], [np.sin(rz), np.cos(rz), 0], [0, 0, 1]])
    R = Rx @ Ry @ Rz
    T = np.array([-10.5, -12.5, 525.0])
    H_real = K @ np.column_stack([R[:, 0], R[:, 1], T])

    def project(P):
        Q = H_real @ P
        return Q[:2] / Q[2]

    src1 = np.array([[0, 200, 50, 181], [0, 0, 139, 93], [1, 1, 1, 1]], dtype=np.float64)
    tar1 = project(src1)
    w, h, mx, my = 50.0, 40.0, 36.0, 81.0
    src2 = np.array([[mx, mx + w, mx, mx + w], [my, my, my + h, my + h], [1, 1, 1, 1]])
    tar2 = project(src2)
    np.savez_compressed(
        os.path.join(OUT, "kat_veri4pts.npz"), H_real=H_real,
        src_general=src1[:2].T.reshape(8), tar_general=tar1.T.reshape(8),
        rect=np.array([mx, my, w, w / h]), tar_rect=tar2.T.reshape(8))
    print("kat_veri4pts.npz", (H_real / H_real[2, 2]).ravel())


if __name__ == "__main__":
    make_ref_general()
    make_ref_torch()
    make_kat()
    make_ref_ge()
