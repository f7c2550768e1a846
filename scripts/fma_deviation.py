#!/usr/bin/env python
"""scripts/fma_deviation.py — how far a build of the element kernel WITH FMA contraction (-fmad=true; the parity build
is -fmad=false) moves from the reference, and what it buys (SURVEY.md §7 "hard part 1": decide with numbers).

Run once per library (NSM_B200_LIB selects it): nodal force and sigma of a 12^3 cube with random displacements of
relative size eps against the CPU oracle (max-norm relative error), for both materials.  Never the headline: the
product, the bench default and every parity claim stay on the contraction-free kernel."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nimblesm_b200 import capi  # noqa: E402
from oracle import hex8 as oracle  # noqa: E402
from tests.conftest import perturbed_cube  # noqa: E402

K, G, RHO = 1.6e12, 0.8e12, 7.8
rows = []
for material, kind in (("elastic", oracle.ELASTIC), ("neohookean", oracle.NEOHOOKEAN)):
    for eps in (1e-6, 1e-4, 1e-3, 1e-2, 1e-1):
        mesh, ref, disp = perturbed_cube(12, eps)
        conn = mesh["conn"][1]
        f_want, ed_want = oracle.internal_force(kind, K, G, ref, disp, conn)
        c = capi.Context(0)
        c.set_nodes(mesh["x"], mesh["y"], mesh["z"])
        c.add_block(1, conn, material, K, G, RHO)
        c.finalize(capi.ASSEMBLY_ORDERED, 2)
        f = c.internal_force_host(disp, store_ipt=True)
        ed = c.element_data(1)
        c.close()
        sig, sig_want = ed[..., 9:], ed_want[..., 9:]
        rows.append({"material": material, "eps": eps,
                     "force_rel": float(np.abs(f - f_want).max() / np.abs(f_want).max()),
                     "sigma_rel": float(np.abs(sig - sig_want).max() / np.abs(sig_want).max()),
                     "F_rel": float(np.abs(ed[..., :9] - ed_want[..., :9]).max())})
print(json.dumps({"library": os.environ.get("NSM_B200_LIB", "nimblesm_b200/lib/libnsm_b200.so"), "kernel_info_sha": capi.kernel_info()["source_sha"],
                  "dp_per_element": {k: v["dp_lane_instr_per_element"] for k, v in capi.kernel_info()["kernels"].items() if k.endswith("ordered0_mode2")},
                  "rows": rows}))
