#!/usr/bin/env python
"""scripts/ncu_sass_mix.py <report.ncu-rep> [n_groups] — dynamic SASS opcode mix of the profiled kernel from the
ncu source page (per-instruction executed counts), plus the stall samples per opcode class."""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
ngroups = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
iS, iE, iT, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
mix, samples, stalls = collections.Counter(), collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iT:
        continue
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS])
    if not m:
        continue
    op = m.group(1)
    n = int(r[iE] or 0)
    mix[op] += n
    tot += n
    samples[op] += int(r[iSamp] or 0)
    for i in stall_cols:
        stalls[hdr[i]] += int(r[i] or 0)
dp = sum(mix[k] for k in ("DADD", "DMUL", "DFMA", "DSETP"))
print("warp-instructions %d, DP %d (%.1f%%)" % (tot, dp, 100.0 * dp / tot))
if ngroups:
    print("per 4-element group: %.1f instr, %.1f DP -> DP lane-ops per element %.1f" % (tot / ngroups, dp / ngroups, dp / ngroups * 8))
    print(" ".join("%s:%.0f" % (k, v / ngroups) for k, v in mix.most_common(36)))
ts = sum(samples.values())
print("samples by opcode:", " ".join("%s:%.1f%%" % (k, 100.0 * v / ts) for k, v in samples.most_common(14)))
tt = sum(stalls.values())
print("stall samples:", " ".join("%s:%.1f%%" % (k[6:], 100.0 * v / tt) for k, v in stalls.most_common(12)))
