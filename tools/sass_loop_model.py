#!/usr/bin/env python
"""Static issue-cycle model of a kernel's hottest loop, read off the SASS.

    python tools/sass_loop_model.py [lib.so] --kernel k_ransac_acaILi2ELi3ELi256

Finds the innermost backward branch with the most FFMA2 in its body and charges
every instruction of that body the scheduler cycles the measured rules give
(B300_MICROARCH.md "RF banking", tools/ubench/fma_peak.cu, issue_mix.cu):

    FFMA2 / FMUL2 / FADD2      max(2, distinct even registers, distinct odd registers)
                               (operands flagged .reuse by the PREVIOUS instruction in the
                               same slot come from the operand-reuse cache and are free)
    FFMA / FMUL / FADD         max(1, even, odd)
    ALU pipe (LEA, IADD3, LOP3, SHF, PRMT, ISETP, FSETP, MOV, SEL ...)   ALU_COST (default 1.65)
    LDS / everything else      1

It prints the body, the cycles, and the fraction of them that are "useful" FP32
lane-cycles (2 per packed op, 1 per scalar op) -- the ceiling the loop can reach
with perfect latency hiding.  A reading aid for kernel work on the CPU box; the
numbers that count are measured on the GPU.
"""
import argparse
import re
import subprocess
import sys

ALU = ("LEA", "IADD", "IADD3", "LOP3", "SHF", "PRMT", "ISETP", "FSETP", "MOV", "SEL", "IMNMX", "FMNMX",
       "VIADD", "POPC", "FLO", "BREV", "I2F", "F2I", "IABS", "VIMNMX")
FP2 = ("FFMA2", "FMUL2", "FADD2")
FP1 = ("FFMA", "FMUL", "FADD")


def sass_of(lib, kernel):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    lines, on = [], False
    for l in out.splitlines():
        if "Function :" in l:
            on = kernel in l
            continue
        if on:
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*", l)
            if m:
                lines.append((int(m.group(1), 16), m.group(2).strip()))
    return lines


def regs_of(op, width):
    """registers an operand reads: R12.F32x2.HI_LO -> R12,R13; R24.F32 -> R24."""
    m = re.match(r"[-|~]*R(\d+)", op)
    if not m:
        return []
    r = int(m.group(1))
    if ".F32x2" in op or width == 2 and ".F32" not in op:
        return [r, r + 1]
    return [r]


def cost(ins, prev_reuse, alu_cost):
    pred = re.match(r"(@!?U?P\d+\s+)?(\S+)\s*(.*)", ins)
    opc, rest = pred.group(2), pred.group(3)
    base = opc.split(".")[0]
    ops = [o.strip() for o in rest.split(",")] if rest else []
    srcs = ops[1:]
    reuse_now = {}
    if base in FP2 or base in FP1:
        width = 2 if base in FP2 else 1
        even, odd = set(), set()
        for slot, o in enumerate(srcs):
            rs = regs_of(o, width if base in FP2 else 1)
            key = re.sub(r"\.reuse", "", o).lstrip("-|")
            if ".reuse" in o:
                reuse_now[slot] = key
            if prev_reuse.get(slot) == key:
                continue                      # served by the reuse cache
            for r in rs:
                (even if r % 2 == 0 else odd).add(r)
        c = max(2 if base in FP2 else 1, len(even), len(odd))
        return c, (2 if base in FP2 else 1), reuse_now
    if base in ALU:
        return alu_cost, 0, reuse_now
    return 1, 0, reuse_now


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("lib", nargs="?", default="sks_homography_b200/libsks_cuda.so")
    ap.add_argument("--kernel", required=True)
    ap.add_argument("--alu-cost", type=float, default=1.65)
    ap.add_argument("--show", action="store_true")
    ap.add_argument("--require", default=None, help="only loops whose body contains this opcode (e.g. FFMA2.RM)")
    a = ap.parse_args()
    ins = sass_of(a.lib, a.kernel)
    if not ins:
        sys.exit("kernel not found")
    addr = {x: i for i, (x, _) in enumerate(ins)}
    best = None
    for i, (x, t) in enumerate(ins):
        m = re.search(r"BRA\S*\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) in addr and int(m.group(1), 16) <= x:
            lo = addr[int(m.group(1), 16)]
            n2 = sum(1 for _, q in ins[lo:i + 1] if re.match(r"(@\S+\s+)?(FFMA2|FMUL2)", q))
            inner = not any(re.search(r"BRA", q) for _, q in ins[lo:i])
            if a.require and not any(a.require in q for _, q in ins[lo:i + 1]):
                continue
            if inner and (best is None or n2 > best[2]):
                best = (lo, i, n2)
    lo, hi, n2 = best
    cyc = useful = 0.0
    prev = {}
    hist = {}
    for x, t in ins[lo:hi + 1]:
        c, u, prev = cost(t, prev, a.alu_cost)
        cyc += c
        useful += u
        k = re.match(r"(@\S+\s+)?(\S+)", t).group(2).split(".")[0]
        hist.setdefault(k, [0, 0.0])
        hist[k][0] += 1
        hist[k][1] += c
        if a.show:
            print(f"{x:05x} {c:4.2f}  {t}")
    print(f"loop {ins[lo][0]:#x}..{ins[hi][0]:#x}: {hi - lo + 1} instructions, {n2} packed FP32")
    for k, (n, c) in sorted(hist.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:8s} x{n:4d}  {c:7.1f} cycles")
    print(f"model: {cyc:.1f} issue cycles, {useful:.0f} FP32 lane-cycles -> ceiling {useful / cyc:.3f} of the FP32 pipe")


if __name__ == "__main__":
    main()
