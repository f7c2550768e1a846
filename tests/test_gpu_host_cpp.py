"""GPU tests of the C++ host layer: the NimbleSM_b200 driver (nimblesm_b200/host) runs the reference's own
regression decks end to end — deck -> Genesis mesh -> B200 ModelData -> ExplicitTimeIntegrator -> Exodus output —
and the `.out.e` it writes is compared with (1) the reference's gold file under the reference's exodiff rules and
(2) snapshots of the reference's serial code on the same deck at 1e-9 * max (tests/golden)."""
import os
import subprocess

import numpy as np
import pytest

from tests.conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "nimblesm_b200", "lib", "NimbleSM_b200")
CASES = ["wave_in_bar", "notched_plate_native_neohookean", "notched_plate_native_hypoelastic", "brick_with_fibers",
         "single_elem_complex_displacement", "single_elem_native_neohookean", "rigid_body_motion", "simple_deformation_modes"]


def _run(tmp_path, case, extra=(), pieces=None):
    import re

    from nimblesm_b200.exodus_py import write_genesis

    deck, mesh, gold, ref, all_pieces = load_golden(case)
    base = re.search(r"genesis input file:\s*(\S+)", deck).group(1)
    if pieces:
        P = pieces
        for r in range(P):
            write_genesis(str(tmp_path / ("%s.%d.%d" % (base, P, r))), all_pieces[(P, r)])
    else:
        write_genesis(str(tmp_path / base), mesh)
    (tmp_path / "case.in").write_text(deck)
    r = subprocess.run([EXE, "--quiet", *extra, "case.in"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = re.search(r"exodus output file:\s*(\S+)", deck).group(1)
    stem = out[:-2] if out.endswith(".e") else out
    return deck, mesh, gold, ref, all_pieces, str(tmp_path / (stem + ".out.e"))


def _check_against_reference(mesh, gold, ref, res):
    from nimblesm_b200 import exodiff

    assert np.array_equal(res["times"], ref["times"])
    for lbl in ("displacement", "velocity", "acceleration", "internal_force"):
        want = ref["node_" + lbl]
        for i, c in enumerate("xyz"):
            key = "%s_%s" % (lbl, c)
            if key in res["nod"]:
                tol = 1e-9 * max(np.abs(want).max(), 1e-300)
                assert np.abs(res["nod"][key] - want[:, :, i]).max() <= tol, key
    if "lumped_mass" in res["nod"]:
        assert np.abs(res["nod"]["lumped_mass"] - ref["node_lumped_mass"]).max() <= 1e-9 * np.abs(ref["node_lumped_mass"]).max()
    for bi, b in enumerate(mesh["all_block_ids"]):
        if b not in mesh["block_ids"]:
            continue
        last = ref["elem_last_%d" % b]
        for key in ref:
            pre = "derived_%d_" % b
            if key.startswith(pre) and (key[len(pre):], bi) in res["elem"]:
                lab = key[len(pre):]
                w, g = ref[key], res["elem"][(lab, bi)]
                scale = np.abs(w).max()
                if lab.startswith("stress"):
                    scale = max(scale, np.abs(last[..., 9:15]).max())
                assert np.abs(g - w).max() <= 1e-9 * max(scale, 1e-300), lab
    fails = exodiff.compare(gold["exodiff"], gold, res)
    assert not fails, fails[:5]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("assembly", ["ordered", "atomic"])
def test_driver_runs_reference_decks(case, assembly, tmp_path):
    """Both assembly modes at the same bars: ORDERED is the driver's default, ATOMIC is what bench.py times."""
    from nimblesm_b200.exodus_py import read_results

    _deck, mesh, gold, ref, _pieces, out = _run(tmp_path, case, extra=("--assembly", assembly))
    _check_against_reference(mesh, gold, ref, read_results(out))


@pytest.mark.parametrize("case", ["wave_in_bar", "brick_with_fibers", "single_elem_complex_displacement"])
def test_reference_sequence_equals_fused_steps(case, tmp_path):
    """The loop issued call by call through the ModelDataBase virtuals on host views (what a drop-in ModelData sees
    inside the reference's unmodified integrator) gives, bit for bit, the fields of the fused device stepping."""
    from nimblesm_b200.exodus_py import read_results

    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    *_x, out_a = _run(tmp_path / "a", case)
    *_y, out_b = _run(tmp_path / "b", case, extra=("--reference_sequence",))
    ra, rb = read_results(out_a), read_results(out_b)
    assert np.array_equal(ra["times"], rb["times"])
    for k in ra["nod"]:
        assert np.array_equal(ra["nod"][k].view(np.int64), rb["nod"][k].view(np.int64)), k
    for k in ra["elem"]:
        assert np.array_equal(ra["elem"][k].view(np.int64), rb["elem"][k].view(np.int64)), k


def test_driver_errors_like_the_reference(tmp_path):
    (tmp_path / "bad.in").write_text("genesis input file: nothing.g\nno such key: 1\n")
    r = subprocess.run([EXE, "bad.in"], cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "unknown key no such key" in r.stderr
    r = subprocess.run([EXE], cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "Usage" in r.stderr


def _num_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except Exception:
        return 0


def _rank_flags(P):
    """One GPU per rank where the box has P GPUs (shared-node sums over NVLink peer memory).  On a smaller box the P
    ranks share GPU 0 in lockstep (`--devices 0,0,...`, RankGroup::SetLockstep): same partition tables, same pack /
    wait / rank-ordered unpack kernels, same boundary-first element schedule -- the peer stores then land in the same
    device's memory instead of crossing NVLink -- so the value checks of the multi-rank path run wherever the GPU
    suite runs."""
    if _num_gpus() < P:
        if _num_gpus() < 1:
            pytest.skip("needs a GPU")
        return ("--gpus", str(P), "--devices", ",".join(["0"] * P))
    return ("--gpus", str(P))


def _join_pieces(tmp_path, stem, P, pieces, mesh):
    """epu: per-rank results -> global arrays by global node / element id (shared nodes must agree bit for bit)."""
    from nimblesm_b200.exodus_py import read_results

    parts = [read_results(str(tmp_path / ("%s.out.e.%d.%d" % (stem, P, r)))) for r in range(P)]
    out = {"times": parts[0]["times"], "nod": {}, "elem": {}}
    n_nodes = len(mesh["x"])
    gid_to_local = {int(g): i for i, g in enumerate(mesh["node_gid"])}
    for k in parts[0]["nod"]:
        a = np.full((len(out["times"]), n_nodes), np.nan)
        for r, pr in enumerate(parts):
            idx = np.array([gid_to_local[int(g)] for g in pieces[(P, r)]["node_gid"]])
            prev = a[:, idx]
            new = pr["nod"][k]
            seen = ~np.isnan(prev)
            assert np.array_equal(prev[seen].view(np.int64), new[seen].view(np.int64)), "replicas of a shared node differ: " + k
            a[:, idx] = new
        out["nod"][k] = a
    for bi, b in enumerate(mesh["all_block_ids"]):
        egid_to_local = {int(g): i for i, g in enumerate(mesh["elem_gid"][b])} if b in mesh["elem_gid"] else {}
        keys = set(k for pr in parts for k in pr["elem"] if k[1] == bi)
        for k in keys:
            a = np.full((len(out["times"]), len(egid_to_local)), np.nan)
            for r, pr in enumerate(parts):
                if k in pr["elem"] and b in pieces[(P, r)]["elem_gid"]:
                    idx = np.array([egid_to_local[int(g)] for g in pieces[(P, r)]["elem_gid"][b]])
                    a[:, idx] = pr["elem"][k]
            out["elem"][k] = a
    return out


@pytest.mark.parametrize("case,P", [("wave_in_bar", 2), ("brick_with_fibers", 2), ("wave_in_bar", 4), ("notched_plate_native_neohookean", 4)])
def test_driver_on_decomposed_meshes(case, P, tmp_path):
    """The reference's -np2 / -np4 regression runs: one rank (thread + GPU) per Nemesis piece, shared-node forces
    summed over NVLink peer memory, per-rank outputs joined by global id and compared with the SERIAL gold file and
    the serial reference snapshots.  Needs P GPUs (gpurun --gpus P)."""
    import re

    deck, mesh, gold, ref, pieces, _out = _run(tmp_path, case, extra=_rank_flags(P), pieces=P)
    out = re.search(r"exodus output file:\s*(\S+)", deck).group(1)
    stem = out[:-2] if out.endswith(".e") else out
    res = _join_pieces(tmp_path, stem, P, pieces, mesh)
    _check_against_reference(mesh, gold, ref, res)


@pytest.mark.parametrize("case,P", [("brick_with_fibers", 2), ("wave_in_bar", 2), ("notched_plate_native_neohookean", 4)])
def test_driver_decomposes_a_serial_mesh(case, P, tmp_path):
    """`NimbleSM_b200 --gpus P` on a SERIAL Genesis file with no Nemesis pieces on disk: the driver bisects the
    elements itself (GenesisMesh::RcbElementPartition), runs one rank per part and writes per-rank outputs whose
    id maps join into the serial result (gold file + reference snapshots)."""
    import re

    from nimblesm_b200.exodus_py import read_results

    deck, mesh, gold, ref, _pieces, _out = _run(tmp_path, case, extra=_rank_flags(P))
    out = re.search(r"exodus output file:\s*(\S+)", deck).group(1)
    stem = out[:-2] if out.endswith(".e") else out
    pieces = {}
    for r in range(P):
        pr = read_results(str(tmp_path / ("%s.out.e.%d.%d" % (stem, P, r))))
        eg, k = {}, 0
        for b, n in zip(pr["block_ids"], pr["num_el_in_blk"]):
            if n:
                eg[b] = pr["elem_gid"][k:k + n]
            k += n
        pieces[(P, r)] = {"node_gid": pr["node_gid"], "elem_gid": eg}
    assert sum(len(p["node_gid"]) for p in pieces.values()) > len(mesh["x"])  # shared nodes are duplicated
    res = _join_pieces(tmp_path, stem, P, pieces, mesh)
    _check_against_reference(mesh, gold, ref, res)


def test_driver_time_dependent_bc_on_device_equals_host_evaluation(tmp_path):
    """A deck with time-dependent expression BCs: the driver compiles them into device programs (only the
    sub-expressions of t are evaluated on the host, one scalar per step); with NSM_B200_HOST_BC=1 it evaluates one
    magnitude per boundary node per step on the host as the reference does.  Both runs write the same bytes."""
    import re

    from nimblesm_b200.exodus_py import read_results, write_genesis

    deck, mesh, _gold, _ref, _pieces = load_golden("wave_in_bar")
    deck = re.sub(r"\n*$", "\n", deck)
    deck += 'boundary condition:  prescribed_velocity nodelist_2 y "0.5*(1.0-cos(t*3.141592653589793/1.0e-6))*(1.0+z)"\n'
    deck += 'boundary condition:  prescribed_displacement nodelist_2 z "1.0e-3*t*(y+2)/(x+3)"\n'
    # a position-only sub-tree through libm and pow: per-entry constants on the device (NSM_BCOP_ENTRYCONST)
    deck += 'boundary condition:  prescribed_velocity nodelist_2 x "1.0e-2*sin(300*y)*cos(t*1.0e6) + z^2*t"\n'
    base = re.search(r"genesis input file:\s*(\S+)", deck).group(1)
    out = re.search(r"exodus output file:\s*(\S+)", deck).group(1)
    stem = out[:-2] if out.endswith(".e") else out
    blobs = {}
    for mode in ("device", "host"):
        d = tmp_path / mode
        d.mkdir()
        write_genesis(str(d / base), mesh)
        (d / "case.in").write_text(deck)
        env = dict(os.environ)
        if mode == "host":
            env["NSM_B200_HOST_BC"] = "1"
        r = subprocess.run([EXE, "--quiet", "case.in"], cwd=d, capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        blobs[mode] = (d / (stem + ".out.e")).read_bytes()
    assert blobs["device"] == blobs["host"]
    res = read_results(str(tmp_path / "device" / (stem + ".out.e")))
    ns2 = mesh["node_sets"][2]
    vy = res["nod"]["velocity_y"][:, ns2]
    assert np.abs(vy).max() > 0  # the prescribed velocity really acted (it returns to 0 at the final time)


def test_driver_timing_summary_and_log(tmp_path):
    """`write timing data file: on` (src/nimble_parser.cc): the driver prints the reference's closing timing summary
    (explicit_time_integrator.cc:307-322) and writes nimble_timing_data_n<ranks>_<stamp>.log in the reference's
    tab-separated layout (src/nimble_timing_utils.cc:70-94): ranks, simulation, force, contact, exodus write, reduction."""
    import glob
    import re

    from nimblesm_b200.exodus_py import write_genesis

    deck, mesh, _gold, _ref, _pieces = load_golden("wave_in_bar")
    deck = re.sub(r"\n*$", "\n", deck) + "write timing data file: on\n"
    base = re.search(r"genesis input file:\s*(\S+)", deck).group(1)
    write_genesis(str(tmp_path / base), mesh)
    (tmp_path / "case.in").write_text(deck)
    r = subprocess.run([EXE, "case.in"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for line in ("======== Timing data: ========", "Total step time = ", " --- Update A, V, U: ", " --- Force: ", " --- Exodus Write = "):
        assert line in r.stdout, line
    logs = glob.glob(str(tmp_path / "nimble_timing_data_n1_*.log"))
    assert len(logs) == 1
    cols = open(logs[0]).read().split()
    assert len(cols) == 6 and int(cols[0]) == 1
    sim, force, contact, exo, red = (float(c) for c in cols[1:])
    assert sim > 0 and 0 < force <= sim and contact == 0.0 and exo >= 0.0 and red == 0.0
