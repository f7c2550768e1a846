#!/usr/bin/env python
"""scripts/contact_driver_run.py [edge] [steps] — penalty contact through the C++ driver at scale: the two-body mesh of
`bench.py --workload contact` (2 x edge x edge x edge/2 neohookean elements, the upper body falling onto the lower one)
written as a Genesis file, a deck with a `contact:` line, NimbleSM_b200 end to end (Genesis reader -> host
ContactManager: skinning of both blocks, entity lists -> device steps with the contact term -> Exodus output with
contact_force).  Prints the driver's own phase report and checks the written contact force of the last output step
against the oracle on a window of contact nodes (the file's own displacement) and action = reaction over the surface.
Run on a GPU box: python scripts/contact_driver_run.py 200 100 > profiles/..."""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from nimblesm_b200.exodus_py import read_results, write_genesis  # noqa: E402

EXE = os.path.join(ROOT, "nimblesm_b200", "lib", "NimbleSM_b200")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    work = tempfile.mkdtemp(prefix="nsm_contact_")
    t0 = time.perf_counter()
    mesh, ent, h = bench.contact_stack(n)
    n_elem = sum(len(c) for c in mesh["conn"].values())
    mesh["all_block_ids"] = [1, 2]
    mesh["node_gid"] = np.arange(len(mesh["x"]), dtype=np.int32)
    mesh["elem_gid"] = {1: np.arange(len(mesh["conn"][1]), dtype=np.int32),
                        2: np.arange(len(mesh["conn"][1]), len(mesh["conn"][1]) + len(mesh["conn"][2]), dtype=np.int32)}
    mesh["node_sets"] = {1: mesh["node_sets"]["bottom"], 2: mesh["node_sets"]["upper"]}
    write_genesis(os.path.join(work, "stack.g"), mesh)
    print("mesh: 2 bodies of %d x %d x %d elements = %d elements, %d nodes; Genesis file written in %.1f s"
          % (n, n, n // 2, n_elem, len(mesh["x"]), time.perf_counter() - t0))
    dt = float(0.2 * h / np.sqrt(bench.BULK / bench.RHO))
    penalty = bench.BULK * h
    deck = "\n".join([
        "genesis input file: stack.g", "exodus output file: stack.e", "final time: %r" % (steps * dt),
        "number of load steps: %d" % steps, "output frequency: %d" % steps, "output fields: displacement contact_force",
        "material parameters: material_1 neohookean density %r bulk_modulus %r shear_modulus %r" % (bench.RHO, bench.BULK, bench.SHEAR),
        "element block: block_1 material_1", "element block: block_2 material_1",
        "boundary condition: prescribed_velocity nodelist_1 x 0.0", "boundary condition: prescribed_velocity nodelist_1 y 0.0",
        "boundary condition: prescribed_velocity nodelist_1 z 0.0", "boundary condition: initial_velocity nodelist_2 z -1000.0",
        "contact: primary_blocks block_2 secondary_blocks block_1 penalty_parameter %r" % penalty]) + "\n"
    open(os.path.join(work, "case.in"), "w").write(deck)
    t0 = time.perf_counter()
    r = subprocess.run([EXE, "--assembly", "atomic", "case.in"], cwd=work, capture_output=True, text=True)
    wall = time.perf_counter() - t0
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    keep = [l for l in r.stdout.splitlines() if any(k in l for k in ("Contact", "contact", "number of", "Total step", " --- ", "element-updates"))]
    print("NimbleSM_b200, %d steps, ATOMIC assembly: %.1f s wall for the whole process\n  " % (steps, wall) + "\n  ".join(keep))
    res = read_results(os.path.join(work, "stack.out.e"))
    u = np.stack([res["nod"]["displacement_" + c][-1] for c in "xyz"], 1)
    fc = np.stack([res["nod"]["contact_force_" + c][-1] for c in "xyz"], 1)
    X = np.stack([mesh["x"], mesh["y"], mesh["z"]], 1)
    from oracle import contact as contact_oracle

    cn = ent["contact_nodes"]
    cur = X[cn] + u[cn]
    w = 8.0 * h
    near = (np.abs(cur[:, 0] - 0.5) < w) & (np.abs(cur[:, 1] - 0.5) < w) & (np.abs(cur[:, 2] - 0.5) < w)
    qc = (X[ent["primary_quads"]] + u[ent["primary_quads"]]).mean(1)
    qsel = (np.abs(qc[:, 0] - 0.5) < w + 3 * h) & (np.abs(qc[:, 1] - 0.5) < w + 3 * h) & (np.abs(qc[:, 2] - 0.5) < w + 3 * h)
    want = np.zeros_like(X)
    pairs = contact_oracle._lib().h8o_contact_force(
        penalty, len(X), np.ascontiguousarray(X), np.ascontiguousarray(u), int(qsel.sum()),
        np.ascontiguousarray(ent["primary_quads"][qsel]).reshape(-1), np.ascontiguousarray(ent["primary_char_len"][qsel]), int(near.sum()),
        np.ascontiguousarray(cn[near]), np.ascontiguousarray(ent["contact_node_char_len"][near]), want, None)
    scale = np.abs(fc).max()
    err = np.abs(fc[cn[near]] - want[cn[near]]).max() / scale
    resid = float(np.abs(fc.astype(np.longdouble).sum(0)).max() / scale)
    print("output step %d (t = %.4e): max|contact force| %.4e; window of %d contact nodes, %d pairs: max|file - oracle| / max = %.2e; "
          "sum of contact forces / largest = %.2e" % (len(res["times"]) - 1, res["times"][-1], scale, int(near.sum()), pairs, err, resid))
    ok = err <= 1e-12 and resid <= 1e-10 and pairs > 0
    print("contact through the driver at %d elements:" % n_elem, "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
