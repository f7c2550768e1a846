#!/usr/bin/env python
"""scripts/config1_driver_run.py — BASELINE.json configs[1] through the C++ driver, end to end, at full size:
synthetic structured hex8 cube 200^3 = 8 M elements, linear elastic, 1000 explicit steps on 1 B200
(Genesis file -> NimbleSM_b200 -> Exodus output), plus a full-size parity run: 20 steps of the same mesh against the
CPU oracle (oracle/hex8_oracle.c on all host cores), displacement and velocity within 1e-9 * max.
Run on a GPU box: python scripts/config1_driver_run.py [edge] > profiles/..."""
import os
import re
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nimblesm_b200.exodus_py import read_results, write_genesis  # noqa: E402
from nimblesm_b200.mesh import structured_cube  # noqa: E402

EXE = os.path.join(ROOT, "nimblesm_b200", "lib", "NimbleSM_b200")
RHO, BULK, SHEAR = 7.8, 1.6e12, 0.8e12


def deck(n_steps, dt, out_freq, fields, prescribed):
    s = ["genesis input file: cube.g", "exodus output file: cube.e", "final time: %r" % (n_steps * dt),
         "number of load steps: %d" % n_steps, "output frequency: %d" % out_freq, "output fields: " + fields,
         "material parameters: material_1 elastic density %r bulk_modulus %r shear_modulus %r" % (RHO, BULK, SHEAR),
         "element block: block_1 material_1", 'boundary condition: initial_velocity nodelist_1 x "1000.0*x"']
    if prescribed:
        s += ["boundary condition: prescribed_velocity nodelist_2 %s 0.0" % c for c in "xyz"]
    return "\n".join(s) + "\n"


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    import tempfile

    work = tempfile.mkdtemp(prefix="nsm_config1_")  # several GB of mesh and results: not under gpurun_out/
    t0 = time.perf_counter()
    mesh = structured_cube(n)
    mesh["all_block_ids"] = [1]
    write_genesis(os.path.join(work, "cube.g"), mesh)
    print("mesh: %d^3 = %d elements, %d nodes; Genesis file written in %.1f s" % (n, n ** 3, len(mesh["x"]), time.perf_counter() - t0))
    dt = float(0.2 * (1.0 / n) / np.sqrt(BULK / RHO))

    # ---- full-size parity: 20 steps, free vibration, vs the CPU oracle
    steps = 20
    open(os.path.join(work, "case.in"), "w").write(deck(steps, dt, steps, "displacement velocity", False))
    t0 = time.perf_counter()
    r = subprocess.run([EXE, "--quiet", "case.in"], cwd=work, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    print("driver, %d steps (ORDERED assembly, output at 0 and %d): %.1f s wall" % (steps, steps, time.perf_counter() - t0))
    res = read_results(os.path.join(work, "cube.out.e"))
    from oracle import hex8 as port

    ref = np.ascontiguousarray(np.stack([mesh["x"], mesh["y"], mesh["z"]], 1))
    conn = mesh["conn"][1]
    mass = port.lumped_mass(RHO, ref, conn)
    u, v, a = np.zeros_like(ref), np.zeros_like(ref), np.zeros_like(ref)
    v[:, 0] = 1000.0 * ref[:, 0]
    t0 = time.perf_counter()
    # the driver accumulates time (t += dt_user; dt = t - t_prev); 20 equal-length steps differ from dt by < 1 ulp of t
    port.bench_steps(port.ELASTIC, BULK, SHEAR, ref, conn, mass, u, v, a, (steps * dt) / steps, steps, os.cpu_count() or 1)
    print("oracle (plain-C restatement, %d threads), %d steps: %.1f s" % (os.cpu_count() or 1, steps, time.perf_counter() - t0))
    ok = True
    for lbl, want in (("displacement", u), ("velocity", v)):
        err = 0.0
        for i, c in enumerate("xyz"):
            got = res["nod"]["%s_%s" % (lbl, c)][-1]
            err = max(err, np.abs(got - want[:, i]).max())
        rel = err / np.abs(want).max()
        print("  times", res["times"].tolist(), "max|driver| per plane", [float(np.abs(res["nod"][lbl + "_x"][k]).max()) for k in range(len(res["times"]))],
              "max|oracle|", float(np.abs(want).max()))
        print("  %s after %d steps at %d elements: max|driver - oracle| / max|oracle| = %.3e" % (lbl, steps, n ** 3, rel))
        ok = ok and rel <= 1e-9

    print("full-size parity:", "PASS" if ok else "FAIL")

    # ---- configs[1]: 1000 steps, prescribed velocity on x = 0, outputs at 0 / 500 / 1000
    steps = 1000
    open(os.path.join(work, "case.in"), "w").write(deck(steps, dt, 500, "displacement velocity stress", True))
    t0 = time.perf_counter()
    r = subprocess.run([EXE, "--assembly", "atomic", "case.in"], cwd=work, capture_output=True, text=True)
    wall = time.perf_counter() - t0
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    m = re.search(r"Total step time = ([0-9.eE+-]+) s", r.stdout)
    loop = float(m.group(1))
    print("configs[1] through NimbleSM_b200 (ATOMIC assembly): %d elements x %d steps" % (n ** 3, steps))
    print("  step loop incl. 2 output steps (volume-averaged stress + nodal fields to Exodus): %.2f s = %.3e element-updates/s"
          % (loop, n ** 3 * steps / loop))
    print("  whole process (read Genesis, setup, lumped mass, 3 Exodus time planes, teardown): %.2f s = %.3e element-updates/s"
          % (wall, n ** 3 * steps / wall))
    res = read_results(os.path.join(work, "cube.out.e"))
    print("  output times:", res["times"].tolist(), " max |u| at the end: %.3e" % max(np.abs(res["nod"]["displacement_" + c][-1]).max() for c in "xyz"))
    import shutil

    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
