#!/usr/bin/env python
"""scripts/ncu_traffic.py <out.json> <kernel_key>=<report.ncu-rep>:<elements> ... — DRAM bytes per element of the
element kernel from `ncu --set full` captures, stamped with the kernel-source hash of the library that was profiled
(nsm_b200_kernel_info), so that bench.py can refuse the figure once the kernels change (run on the GPU box by
scripts/ncu_traffic.sh)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nimblesm_b200 import capi  # noqa: E402

out = {"source_sha": capi.kernel_info()["source_sha"], "kernels": {}, "report": "profiles/ (ncu --set full --clock-control none)"}
for spec in sys.argv[2:]:
    key, rest = spec.split("=", 1)
    rep, elems = rest.rsplit(":", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    d = dict(zip(hdr, rows[2]))

    def gbytes(name):
        v, u = float(d[name].replace(",", "")), units[hdr.index(name)].lower()
        return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]

    tot = gbytes("dram__bytes_read.sum") + gbytes("dram__bytes_write.sum")
    out["kernels"][key] = {"dram_bytes_per_element": tot / float(elems), "elements": int(elems), "dram_bytes": tot,
                           "kernel": d.get("Kernel Name"), "duration": d.get("gpu__time_duration.sum"),
                           "fp64_pipe_pct": d.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                           "report": os.path.basename(rep)}
json.dump(out, open(sys.argv[1], "w"), indent=1, sort_keys=True)
print(json.dumps(out))
