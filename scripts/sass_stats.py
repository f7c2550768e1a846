#!/usr/bin/env python
"""scripts/sass_stats.py — static SASS opcode histogram of one kernel of libnsm_b200.so (cuobjdump -sass).
usage: sass_stats.py <so> <kernel-substring> [--dump]"""
import collections, re, subprocess, sys
so, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, hist, lines = None, collections.Counter(), []
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            hist[m.group(1).split(".")[0]] += 1
            lines.append(ln)
tot = sum(hist.values())
dp = sum(v for k, v in hist.items() if k in ("DADD", "DMUL", "DFMA", "DSETP"))
print("total", tot, "DP", dp)
print(" ".join("%s:%d" % kv for kv in hist.most_common(40)))
if "--dump" in sys.argv:
    print("\n".join(lines))
