#!/usr/bin/env python
"""TEST / BENCH INFRASTRUCTURE ONLY -- stages two things from the reference checkout into the
git-ignored oracle/_ref/ at build time (like the Makefile does for the reference's CUDA
kernels), because /root/reference does not exist on the GPU box:

  _ref/ref_torch_funcs.py   the reference's OWN torch statements -- getInput / getTar / adjust and
                            the timed bodies of TensorACA_rect (PY.py:296-302) and ACA_vanilla
                            (PY.py:322-381) -- extracted with `ast` and wrapped as functions, so
                            that bench.py's gpu_baseline can run them unmodified on cuda tensors
                            (torch eager: the reference's existing fp32 GPU implementation)
  _ref/orig_pts_wall.txt    the real-data fixture (2540 wall matches, CPU/orig_pts_wall.txt) for the
                            GPU RANSAC test on real matches

Nothing staged here is committed, and nothing in the product imports it.
usage: python oracle/stage_ref.py [REFERENCE_ROOT]
"""
import ast
import os
import shutil
import sys
import textwrap

HERE = os.path.dirname(os.path.abspath(__file__))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
PY_REF = os.path.join(REF, "PyTorch Codes", "Modules_Runtime_Test.py")
FIXTURE = os.path.join(REF, "C++ Codes", "Runtime Test", "CPU_Runtime Test", "orig_pts_wall.txt")
OUT = os.path.join(HERE, "_ref")


def timed_body(text, node):
    loop = next(n for n in node.body if isinstance(n, ast.For))
    keep = []
    for st in loop.body:
        seg = ast.get_source_segment(text, st)
        if any(w in seg for w in ("synchronize", "perf_counter", "time_list")):
            continue
        keep.append(seg)
    return keep


def main():
    os.makedirs(OUT, exist_ok=True)
    text = open(PY_REF).read()
    nodes = {n.name: n for n in ast.parse(text).body if isinstance(n, ast.FunctionDef)}
    out = ['"""GENERATED at build time by oracle/stage_ref.py from the reference checkout',
           '(PyTorch Codes/Modules_Runtime_Test.py); git-ignored, never committed."""', "import torch", ""]
    for name in ("getInput", "getTar", "adjust"):
        out += [ast.get_source_segment(text, nodes[name]), "", ""]
    for name, sig in (("TensorACA_rect", "bs, src, tar, scale, div"), ("ACA_vanilla", "bs, src, tar")):
        body = "\n".join(timed_body(text, nodes[name]))
        out += [f"def {name}_body({sig}):", textwrap.indent(body, "    "), "    return H", "", ""]
    with open(os.path.join(OUT, "ref_torch_funcs.py"), "w") as f:
        f.write("\n".join(out))
    shutil.copyfile(FIXTURE, os.path.join(OUT, "orig_pts_wall.txt"))
    print("staged", os.path.join(OUT, "ref_torch_funcs.py"), "and orig_pts_wall.txt")


if __name__ == "__main__":
    main()
