#!/usr/bin/env python
"""scripts/sass_hot_loop.py — static instruction mix of the element kernel's hot loop, read off the SASS of the
object that is linked into libnsm_b200.so (run by nimblesm_b200/csrc/Makefile; the result is compiled into the
library and returned by nsm_b200_kernel_info(), so the figures `bench.py` reports are those of the binary it runs).

usage: sass_hot_loop.py <object or .so> <sources whose hash is recorded ...> > kernel_info.json

Per element_force_kernel<MAT, ORDERED, MODE> instance:
  loop        = the outermost backward branch's [target, branch] range (the persistent warp's `while (g < n_groups)`)
  cold path   = every range inside the loop that a forward branch skips and that holds > 400 instructions (the
                plain-IEEE redo of an integration point, taken only when an operand leaves the branch-free window:
                nsm_b200_cold_points counts 0 in every benchmark run)
  hot pass    = loop minus cold path: one pass of a warp over a group of 4 elements
  dp          = DADD + DMUL + DFMA + DSETP in the hot pass; x 32 lanes / 4 elements = DP lane-instructions per
                element-update (the executed count of ncu's sm__inst_executed_pipe_fp64 agrees to the instruction:
                profiles/r01A_ncu_elem_neo_f2_summary.txt 1121 per pass)
"""
import collections
import hashlib
import json
import re
import subprocess
import sys

DP_OPS = ("DADD", "DMUL", "DFMA", "DSETP")


def functions(sass):
    cur, out = None, collections.OrderedDict()
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;", ln)
        if m and cur:
            out[cur].append((int(m.group(1), 16), m.group(2)))
    return out


def opcode(text):
    m = re.match(r"(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", text)
    return m.group(1) if m else "?"


def analyse(ins):
    index = {a: i for i, (a, _) in enumerate(ins)}
    branches = []
    for i, (a, t) in enumerate(ins):
        if opcode(t) != "BRA" or ".DIV" in t:
            continue
        m = re.search(r"0x([0-9a-f]+)\s*$", t)
        if m and int(m.group(1), 16) in index:
            branches.append((i, index[int(m.group(1), 16)], t.startswith("@")))
    back = [(i, j) for i, j, pred in branches if pred and j <= i and i - j > 200]  # (unpredicated ones return from BRA.DIV stubs)
    if not back:
        return None
    hi, lo = max(back, key=lambda b: b[0] - b[1])
    cold = [(i + 1, j) for i, j, _ in branches if lo <= i < hi and j > i and j - i > 400 and j <= hi + 1]
    hot = [k for k in range(lo, hi + 1) if not any(a <= k < b for a, b in cold)]
    mix = collections.Counter(opcode(ins[k][1]) for k in hot)
    dp = sum(mix[o] for o in DP_OPS)
    return {"loop_instructions": hi - lo + 1, "cold_instructions": (hi - lo + 1) - len(hot), "hot_instructions": len(hot),
            "dp": dp, "other": len(hot) - dp, "dp_lane_instr_per_element": dp * 8,
            "mix": dict(mix.most_common(24))}


def main():
    obj, sources = sys.argv[1], sys.argv[2:]
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", obj], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*(.*)", res):
        usage[m.group(1)] = dict((k, int(v)) for k, v in re.findall(r"(REG|SHARED|LOCAL|STACK):(\d+)", m.group(2)))
    h = hashlib.sha256()
    for s in sources:
        h.update(open(s, "rb").read())
    out = {"source_sha": h.hexdigest()[:16], "sources": [s.split("/")[-1] for s in sources], "kernels": {}}
    for name, ins in functions(sass).items():
        m = re.search(r"element_force_kernelILi(\d+)ELb([01])ELi(\d+)E", name)
        if not m:
            continue
        a = analyse(ins)
        if a is None:
            continue
        a.update({k.lower(): v for k, v in usage.get(name, {}).items()})
        out["kernels"]["mat%s_ordered%s_mode%s" % m.groups()] = a
    json.dump(out, sys.stdout, indent=1, sort_keys=True)
    print()


if __name__ == "__main__":
    main()
