"""GPU tests of the penalty contact (second half): csrc/contact.cuh behind nsm_b200_set_contact /
nsm_b200_contact_force / the contact term of nsm_b200_step, and the host ContactManager inside the NimbleSM_b200 driver.

Checker: oracle/contact_oracle.c (pinned bit for bit to the reference's own ContactEntity objects and to the reference's
gold files, tests/test_oracle.py).  Bars: contact force of a given displacement within 1e-12 (max-norm relative; each
pair's force has the oracle's bits, only the order of the sum over pairs differs) with the SAME set of accepted pairs;
fields after N steps within 1e-9 * max; the reference's gold files under the reference's exodiff rules."""
import ctypes as C
import os

import numpy as np
import pytest

from tests.conftest import ROOT, load_golden
from tests.test_gpu_host_cpp import EXE, _run
from tests.test_host_cpp import LIB as HOST_LIB
from tests.test_host_cpp import host_contact_entities

pytestmark = pytest.mark.gpu


def _rel(a, b):
    s = np.abs(b).max()
    return np.abs(a - b).max() / (s if s > 0 else 1.0)


@pytest.fixture(scope="module")
def host():
    return C.CDLL(HOST_LIB)


def _context(mesh, deck_obj, flags=0, assembly=None):
    from nimblesm_b200 import capi

    c = capi.Context(0)
    c.set_nodes(mesh["x"], mesh["y"], mesh["z"])
    for b in mesh["block_ids"]:
        m = deck_obj.block_material(b)
        c.add_block(b, mesh["conn"][b], m.model, m.bulk_modulus, m.shear_modulus, m.density)
    c.finalize(capi.ASSEMBLY_ORDERED if assembly is None else assembly, flags)
    return c


@pytest.mark.parametrize("flags,ordered", [(0, False), (8, False), (0, True), (8, True)])
def test_contact_force_vs_oracle(oracle, host, tmp_path, flags, ordered):
    """One-shot evaluations on the cubes_contact mesh: entity lists from the C++ ContactManager, displacement fields in
    contact, separated, and pushed through by more than the facets' characteristic length.  Same accepted pairs and
    active entities as the oracle's all-pairs walk, force within 1e-12; also with the nodes renumbered inside the context.
    ORDERED assembly files the pairs' contributions and adds them in the serial order: the force is then BIT-IDENTICAL to
    the oracle's (and to the walk over the reference's own ContactEntity objects, which the oracle equals bit for bit)."""
    from nimblesm_b200.deck import parse_deck
    from nimblesm_b200.exodus_py import write_genesis
    from oracle.model import OracleModel

    deck, mesh, *_ = load_golden("cubes_contact")
    g = str(tmp_path / "c.g")
    write_genesis(g, mesh)
    ent = host_contact_entities(host, g, deck)
    om = OracleModel(deck, mesh)
    n = len(mesh["x"])
    in_b1 = np.zeros(n, bool)
    in_b1[np.unique(mesh["conn"][1])] = True
    rng = np.random.default_rng(5)
    total = 0
    from nimblesm_b200 import capi

    with _context(mesh, parse_deck(deck), flags, capi.ASSEMBLY_ORDERED if ordered else capi.ASSEMBLY_ATOMIC) as c:
        c.set_contact(ent["penalty"], ent["primary_quads"], ent["primary_char_len"], ent["contact_nodes"], ent["contact_node_char_len"])
        for amp, shift in ((0.0, 0.0), (1e-4, 0.0), (1e-3, -2e-3), (1e-2, -0.02), (1e-3, 0.05), (0.0, -1.0), (1e-3, -3e-3)):
            u = amp * (2.0 * rng.random((n, 3)) - 1.0)
            u[in_b1, 2] += shift
            want, pairs, status = om.contact.force(u, want_status=True)
            got = c.contact_force_host(u)
            st = c.contact_stats()
            assert st["pairs"] == pairs, (amp, shift, st)
            assert st["active_faces"] == int(status[:4 * len(ent["primary_quads"])].sum())
            assert st["active_nodes"] == int(status[4 * len(ent["primary_quads"]):].sum())
            assert st["box_tested"] >= pairs
            # contact_status of every entity, in the entity order of the reference (triangle k of face f at 4 f + k)
            face_status, node_status = c.contact_status(len(ent["primary_quads"]), len(ent["contact_nodes"]))
            assert np.array_equal(np.concatenate([face_status, node_status]), status)
            if pairs:
                assert _rel(got, want) <= 1e-12, (amp, shift, _rel(got, want))
            else:
                assert not got.any()
            if ordered:
                assert st["ordered_overflow_pairs"] == 0
                assert np.array_equal(got.view(np.int64), want.view(np.int64)), "ORDERED contact force must be bit-identical"
            total += pairs
            # the device-resident call gives the same field
            c.upload("displacement", u)
            c.contact_force()
            assert _rel(c.download("contact_force"), want) <= 1e-12
        # switching contact off clears the field
        c.set_contact(0.0, np.zeros((0, 4), np.int32), np.zeros(0), np.zeros(0, np.int32), np.zeros(0))
        assert not c.download("contact_force").any() and c.contact_stats()["pairs"] == 0
    assert total > 100


def test_contact_argument_errors(host, tmp_path):
    from nimblesm_b200 import capi
    from nimblesm_b200.deck import parse_deck
    from nimblesm_b200.exodus_py import write_genesis

    deck, mesh, *_ = load_golden("cubes_contact")
    g = str(tmp_path / "c.g")
    write_genesis(g, mesh)
    ent = host_contact_entities(host, g, deck)
    with _context(mesh, parse_deck(deck)) as c:
        with pytest.raises(capi.NsmError):  # no entities yet
            c.contact_force()
        with pytest.raises(capi.NsmError):
            c.download("contact_force")
        with pytest.raises(capi.NsmError) as e:  # ComputeContactForce: invalid penalty_parameter
            c.set_contact(0.0, ent["primary_quads"], ent["primary_char_len"], ent["contact_nodes"], ent["contact_node_char_len"])
        assert "penalty_parameter" in str(e.value)
        bad = ent["primary_quads"].copy()
        bad[3, 2] = len(mesh["x"])
        with pytest.raises(capi.NsmError):
            c.set_contact(1.0, bad, ent["primary_char_len"], ent["contact_nodes"], ent["contact_node_char_len"])
        # a refused call leaves the entities in place as they were; replacing and dropping them returns their memory
        bare = c.device_bytes
        args = (ent["primary_quads"], ent["primary_char_len"], ent["contact_nodes"], ent["contact_node_char_len"])
        c.set_contact(1.0e5, *args)
        c.contact_force()
        with_contact, fc = c.device_bytes, c.download("contact_force")
        for refused in ((0.0,) + args, (1.0, bad) + args[1:]):
            with pytest.raises(capi.NsmError):
                c.set_contact(*refused)
            c.contact_force()
            assert np.array_equal(c.download("contact_force"), fc) and c.device_bytes == with_contact
        c.set_contact(2.0e5, *args)
        assert c.device_bytes == with_contact
        c.set_contact(0.0, args[0][:0], args[1][:0], args[2][:0], args[3][:0])
        fc_bytes = 3 * 8 * len(mesh["x"])  # the nodal contact-force field stays (zeroed) once it exists
        assert c.device_bytes <= bare + fc_bytes + 256 and not c.download("contact_force").any()


@pytest.mark.parametrize("assembly", ["ordered", "atomic"])
def test_contact_steps_vs_oracle(oracle, host, tmp_path, assembly):
    """The explicit loop with the contact term on the device (nsm_b200_step: predict, elements, contact, correct) against
    the oracle's loop on cubes_contact: 100 steps in runs of 1 / 7 / the rest, contact force and fields at 1e-9 * max
    (measured ~1e-14), the same number of enforced pairs at the end."""
    from nimblesm_b200 import capi
    from nimblesm_b200.deck import parse_deck
    from nimblesm_b200.exodus_py import write_genesis
    from oracle.model import OracleModel

    deck, mesh, *_ = load_golden("cubes_contact")
    g = str(tmp_path / "c.g")
    write_genesis(g, mesh)
    ent = host_contact_entities(host, g, deck)
    d = parse_deck(deck)
    om = OracleModel(deck, mesh)
    om.begin()
    om.advance(100)
    tn, tc, tv = [], [], []
    v0 = np.zeros((len(mesh["x"]), 3))
    for bc in d.boundary_conditions:
        ns = mesh["node_sets"][bc.node_set_id]
        if bc.kind == "initial_velocity":
            v0[ns, bc.coordinate] = bc.magnitude
        else:
            tn.append(ns), tc.append(np.full(len(ns), bc.coordinate, np.int32)), tv.append(np.full(len(ns), bc.magnitude))
    tn, tc, tv = np.concatenate(tn).astype(np.int32), np.concatenate(tc), np.concatenate(tv)
    asm = capi.ASSEMBLY_ORDERED if assembly == "ordered" else capi.ASSEMBLY_ATOMIC
    with _context(mesh, d, 2, asm) as c:
        c.compute_lumped_mass()
        c.set_contact(ent["penalty"], ent["primary_quads"], ent["primary_char_len"], ent["contact_nodes"], ent["contact_node_char_len"])
        c.set_bc_table(tn, tc, np.zeros(len(tn), np.int32))
        c.set_bc_values(tv)
        for k in range(len(tn)):  # ApplyKinematicConditions at t = 0 (later entries win)
            v0[tn[k], tc[k]] = tv[k]
        c.upload("velocity", v0)
        dt = (d.final_time - d.initial_time) / d.num_load_steps
        t = c.step(1, 0.0, dt)
        t = c.step(7, t, dt)
        t = c.step(92, t, dt, store_ipt_last=True)  # the last step is an output step: boundary conditions once more
        assert t == om.time
        for lbl, want in (("displacement", om.u), ("velocity", om.v), ("acceleration", om.a), ("contact_force", om.fcontact)):
            assert _rel(c.download(lbl), want) <= 1e-9, lbl
        assert c.contact_stats()["pairs"] == om.contact_pairs > 0
        # Forces on the SAME displacement (the 1e-12 bar).  The force of the oracle's own trajectory is a weaker check in
        # this deck: x = X + u carries 1e-16 of rounding, the strain 1e-15, the stress K * 1e-15 = 1.6e-4 -- one ulp of u
        # moves nodal forces by ~2e-4 of a maximum of 9e4 (2e-9), whoever computes them.
        ug = c.download("displacement")
        f_same = np.zeros_like(ug)
        for b in sorted(mesh["block_ids"]):
            m = d.block_material(b)
            fb, _ed = oracle.internal_force(oracle.NEOHOOKEAN, m.bulk_modulus, m.shear_modulus, om.ref, ug, mesh["conn"][b], False)
            f_same += fb
        assert _rel(c.download("internal_force"), f_same) <= 1e-12
        assert _rel(c.download("contact_force"), om.contact.force(ug)[0]) <= 1e-12
        assert _rel(c.download("internal_force"), om.f) <= 1e-7
        if assembly == "ordered":  # internal and contact forces both summed in the serial order: the same trajectory, bit for bit
            for lbl, want in (("displacement", om.u), ("velocity", om.v), ("acceleration", om.a), ("internal_force", om.f),
                              ("contact_force", om.fcontact)):
                assert np.array_equal(c.download(lbl).view(np.int64), want.view(np.int64)), lbl
            assert c.contact_stats()["ordered_overflow_pairs"] == 0
        # the host-state step takes the plain schedule with contact and carries the same term
        U, V, A, Fo = (c.download(l) for l in ("displacement", "velocity", "acceleration", "internal_force"))
        t2 = c.step_host(t, dt, U, V, A, Fo)
    om.advance(1)
    assert t2 == om.time and _rel(U, om.u) <= 1e-9 and _rel(A, om.a) <= 1e-9


@pytest.mark.parametrize("case", ["cubes_contact", "sphere_plate_contact", "sliding_contact"])
@pytest.mark.parametrize("extra", [(), ("--assembly", "atomic"), ("--reference_sequence",), ("--flags", "14")])
def test_driver_runs_contact_decks(case, extra, tmp_path):
    """NimbleSM_b200 on the reference's contact decks (the reference runs them only in its Kokkos + ArborX / BVH builds):
    deck -> Genesis mesh -> ContactManager (skinning, entities) -> device steps with the contact term -> Exodus output.
    The output is compared with the reference's gold file under the reference's exodiff rules and with snapshots of the
    reference's serial code + ContactEntity objects (tests/golden) at 1e-9 * max, contact_force included; fused stepping,
    ATOMIC assembly, the call-by-call reference sequence (ComputeContactForce on host views), and with the elements
    walked and the nodes numbered along Morton curves inside the context (--flags 14: the entity lists keep the caller's ids)."""
    from nimblesm_b200 import exodiff
    from nimblesm_b200.exodus_py import read_results

    if case == "sliding_contact" and extra:
        pytest.skip("1000 steps of a 1149-element deck: one schedule is enough")
    _deck, mesh, gold, ref, _pieces, out = _run(tmp_path, case, extra=extra)
    res = read_results(out)
    idx = ref["snapshot_index"] if "snapshot_index" in ref else np.arange(len(ref["times"]))
    assert np.array_equal(res["times"][idx], ref["times"])
    for lbl in ("displacement", "velocity", "internal_force", "contact_force"):
        want = ref["node_" + lbl]
        # internal force of a TRAJECTORY: one ulp of u moves it by K * eps * |X| (see test_contact_steps_vs_oracle); the
        # tight bar for it is the same-displacement check below
        bar = 1e-7 if lbl == "internal_force" else 1e-9
        for i, comp in enumerate("xyz"):
            key = "%s_%s" % (lbl, comp)
            if key in res["nod"]:
                assert np.abs(res["nod"][key][idx] - want[:, :, i]).max() <= bar * np.abs(want).max(), key
    assert "contact_force_z" in res["nod"] and np.abs(res["nod"]["contact_force_x"]).max() > 0
    if "atomic" not in extra:
        # ORDERED assembly (the driver's default): internal force summed in the serial element order, contact force in the
        # serial pair order -- every nodal field the file holds is BIT-IDENTICAL to the serial reference-entity run
        for lbl in ("displacement", "velocity", "internal_force", "contact_force"):
            for i, comp in enumerate("xyz"):
                key = "%s_%s" % (lbl, comp)
                if key in res["nod"]:
                    assert np.array_equal(res["nod"][key][idx].view(np.int64), ref["node_" + lbl][:, :, i].view(np.int64)), key
    # forces of the file's own displacement at the last output step, recomputed by the oracle: 1e-12
    from nimblesm_b200.deck import parse_deck
    from oracle import contact as contact_oracle
    from oracle import hex8

    d = parse_deck(_deck)
    X = np.ascontiguousarray(np.stack([mesh["x"], mesh["y"], mesh["z"]], 1))
    u_last = np.ascontiguousarray(np.stack([res["nod"]["displacement_" + comp][-1] for comp in "xyz"], 1))
    prim, sec, penalty = contact_oracle.parse_contact_command(d.contact_string)
    ids = lambda names: [int(nm.rsplit("_", 1)[1]) for nm in names]
    fc_want, pairs = contact_oracle.ContactSetup(mesh, ids(prim), ids(sec), penalty).force(u_last)
    fc_got = np.stack([res["nod"]["contact_force_" + comp][-1] for comp in "xyz"], 1)
    assert np.abs(fc_got - fc_want).max() <= 1e-12 * max(np.abs(fc_want).max(), 1e-300) and (pairs > 0 or not fc_got.any())
    if "internal_force_x" in res["nod"]:
        f_want = np.zeros_like(X)
        for b in sorted(mesh["block_ids"]):
            m = d.block_material(b)
            fb, _ed = hex8.internal_force(hex8.NEOHOOKEAN, m.bulk_modulus, m.shear_modulus, X, u_last, mesh["conn"][b], False)
            f_want += fb
        f_got = np.stack([res["nod"]["internal_force_" + comp][-1] for comp in "xyz"], 1)
        assert np.abs(f_got - f_want).max() <= 1e-12 * np.abs(f_want).max()
    if len(gold["times"]):
        fails = exodiff.compare(gold["exodiff"], gold, res, rounding_floor=1e-11)  # (absolute 2e-4 on forces of 3e8: see exodiff.compare)
        assert not fails, fails[:5]


@pytest.mark.parametrize("extra", [(), ("--assembly", "atomic")])
def test_driver_runs_contact_entity_creation_deck(extra, tmp_path):
    """test/contact/contact_entity_creation through NimbleSM_b200: five blocks of three materials, primary block_1 block_2,
    secondary block_3 block_5, block_4 outside the contact definition, `contact visualization` on.
    Blocks 1 and 5 fly at 1e7 for 2e-9 s and meet nothing: the contact force stays zero, the nodal fields equal the
    reference-entity snapshots (bit for bit in ORDERED assembly), and the displacement the contact ENTITIES see at the end
    -- facet vertices, fictitious face-centre nodes, contact nodes, in the reference's entity order -- meets the last
    record of the reference's gold visualisation database under contact_entity_creation.exodiff (2e-8 / 1e-12 / 1e-12)."""
    from nimblesm_b200.deck import parse_deck
    from nimblesm_b200.exodus_py import read_results
    from oracle import contact as contact_oracle

    deck, mesh, gold, ref, _pieces, out = _run(tmp_path, "contact_entity_creation", extra=extra)
    res = read_results(out)
    assert np.array_equal(res["times"], ref["times"])
    for i, comp in enumerate("xyz"):
        got, want = res["nod"]["displacement_" + comp], ref["node_displacement"][:, :, i]
        assert np.abs(got - want).max() <= 1e-9 * np.abs(ref["node_displacement"]).max()
        if "atomic" not in extra:
            assert np.array_equal(got.view(np.int64), want.view(np.int64))
        assert not res["nod"]["contact_force_" + comp].any()
    prim, sec, penalty = contact_oracle.parse_contact_command(parse_deck(deck).contact_string)
    ids = lambda names: [int(nm.rsplit("_", 1)[1]) for nm in names]
    cs = contact_oracle.ContactSetup(mesh, ids(prim), ids(sec), penalty)
    u_last = np.stack([res["nod"]["displacement_" + comp][-1] for comp in "xyz"], 1)
    seen = cs.entity_vertices(cs.ref + u_last) - cs.entity_vertices(cs.ref)
    assert len(seen) == len(gold["vis"]["coordx"]) == 6778
    for i, (comp, tol) in enumerate(zip("xyz", (2.0e-8, 1.0e-12, 1.0e-12))):
        assert np.abs(seen[:, i] - gold["nod"]["displacement_" + comp][-1]).max() <= tol, comp
    assert np.abs(seen[:, 0]).max() == pytest.approx(0.02, rel=1e-9)
    # ... and the reference's own contract for this test (run_exodiff_test.py): the contact visualisation database the
    # deck asks for, <file_name>.out.e, against contact_entity_creation.gold.e under contact_entity_creation.exodiff
    from nimblesm_b200 import exodiff

    vis = read_results(str(tmp_path / "contact_entity_creation.out.e"))
    fails = exodiff.compare(gold["exodiff"], gold, vis)
    assert not fails, fails[:5]
    assert np.array_equal(vis["nod"]["displacement_x"][-1], seen[:, 0]) and not vis["nod"]["contact_status"].any()


def _check_contact_run(mesh, gold, ref, res):
    from nimblesm_b200 import exodiff

    idx = ref["snapshot_index"] if "snapshot_index" in ref else np.arange(len(ref["times"]))
    assert np.array_equal(res["times"][idx], ref["times"])
    for lbl in ("displacement", "velocity", "internal_force", "contact_force"):
        want = ref["node_" + lbl]
        bar = 1e-7 if lbl == "internal_force" else 1e-9  # (a trajectory's internal force: see test_contact_steps_vs_oracle)
        for i, comp in enumerate("xyz"):
            key = "%s_%s" % (lbl, comp)
            if key in res["nod"]:
                assert np.abs(res["nod"][key][idx] - want[:, :, i]).max() <= bar * np.abs(want).max(), key
    assert np.abs(res["nod"]["contact_force_x"]).max() > 0
    if len(gold["times"]):
        fails = exodiff.compare(gold["exodiff"], gold, res, rounding_floor=1e-11)  # (absolute 2e-4 on forces of 3e8: see exodiff.compare)
        assert not fails, fails[:5]


@pytest.mark.parametrize("case,P,how", [("sphere_plate_contact", 2, "pieces"), ("sphere_plate_contact", 4, "pieces"),
                                        ("sliding_contact", 2, "pieces"), ("cubes_contact", 2, "rcb"), ("cubes_contact", 4, "rcb")])
def test_driver_contact_across_partitions(case, P, how, tmp_path):
    """The reference's -np2 / -np4 contact runs (test/contact/*/CMakeLists.txt): one rank per Nemesis piece -- or per
    part of the driver's own bisection of the serial mesh -- with the contact surface replicated: every rank learns the
    whole skin (partition cuts drop out), rank 0 evaluates the contact force of the whole surface on its GPU from the
    pooled displacements, every rank picks its nodes' entries.  Per-rank outputs are joined by global id (replicas of a
    shared node must agree bit for bit, contact_force included) and compared with the SERIAL reference-entity snapshots
    at 1e-9 * max and with the reference's serial gold file under its exodiff rules."""
    import re

    from nimblesm_b200.exodus_py import read_results
    from tests.test_gpu_host_cpp import _join_pieces, _rank_flags

    deck, mesh, gold, ref, pieces, _out = _run(tmp_path, case, extra=_rank_flags(P), pieces=P if how == "pieces" else None)
    out = re.search(r"exodus output file:\s*(\S+)", deck).group(1)
    stem = out[:-2] if out.endswith(".e") else out
    if how == "rcb":
        pieces = {}
        for r in range(P):
            pr = read_results(str(tmp_path / ("%s.out.e.%d.%d" % (stem, P, r))))
            eg, k = {}, 0
            for b, n in zip(pr["block_ids"], pr["num_el_in_blk"]):
                if n:
                    eg[b] = pr["elem_gid"][k:k + n]
                k += n
            pieces[(P, r)] = {"node_gid": pr["node_gid"], "elem_gid": eg}
    res = _join_pieces(tmp_path, stem, P, pieces, mesh)
    _check_contact_run(mesh, gold, ref, res)


@pytest.mark.parametrize("seed", range(6))
def test_contact_search_never_misses_a_pair(oracle, seed):
    """The hashed-grid search against the oracle's all-pairs walk on random two-body configurations: body size, length
    scale (1e-3 ... 1e3), position (negative coordinates, far from the origin), random and rigid displacements that open,
    close and over-close the gap, so that cell pitch, grid anchor and hash buckets differ from case to case.  Same number
    of accepted pairs, same active entities, force within 1e-12 -- every time."""
    import bench
    from nimblesm_b200 import capi
    from oracle.contact import ContactSetup

    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(2, 9)) * 2
    mesh, _ent, h = bench.contact_stack(n)
    scale = 10.0 ** rng.uniform(-3, 3)
    shift = scale * rng.uniform(-50, 50, 3) * (rng.random(3) < 0.7)
    for k, c_ in zip("xyz", range(3)):
        mesh[k] = np.ascontiguousarray(mesh[k] * scale + shift[c_])
    h *= scale
    cs = ContactSetup(mesh, [2], [1], 1.0e9 * scale)  # entity lists (pinned to the host C++ lists in tests/test_host_cpp.py)
    nn = len(mesh["x"])
    upper = mesh["node_sets"]["upper"]
    c = capi.Context(0)
    c.set_nodes(mesh["x"], mesh["y"], mesh["z"])
    for b in (1, 2):
        c.add_block(b, mesh["conn"][b], "elastic", 1.6e12, 0.8e12, 7.8)
    c.finalize(capi.ASSEMBLY_ATOMIC, 0)
    c.set_contact(cs.penalty, cs.primary_quads, cs.primary_char_len, cs.contact_nodes, cs.contact_node_char_len)
    total = 0
    for trial in range(8):
        u = h * 10.0 ** rng.uniform(-4, -0.5) * (2.0 * rng.random((nn, 3)) - 1.0)
        u[upper] += h * np.array([rng.uniform(-1.5, 1.5), rng.uniform(-1.5, 1.5), rng.choice([-0.3, -0.05, -0.004, 0.0, 0.2, -1.4, -2.5])])
        want, pairs, status = cs.force(u, want_status=True)
        got = c.contact_force_host(u)
        st = c.contact_stats()
        assert st["pairs"] == pairs, (seed, trial, st, pairs)
        assert st["active_faces"] == int(status[:4 * len(cs.primary_quads)].sum()) and st["active_nodes"] == int(status[4 * len(cs.primary_quads):].sum())
        if pairs:
            assert _rel(got, want) <= 1e-12, (seed, trial)
        else:
            assert not got.any()
        total += pairs
    c.close()
    assert total > 0


def test_contact_workload_of_the_bench_checks_itself():
    """`bench.py --workload contact` at a size the suite affords (2 x 96 x 96 x 48 elements, 147 k contact triangles): the
    line's own parity block -- a window of contact nodes against the oracle on the device's displacement (1e-12), the
    whole surface's action = reaction (1e-10) -- must hold, and the step with contact must not cost more than a third
    over the step without."""
    import json
    import subprocess
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "contact", "--n", "96", "--steps", "5"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stdout[-1000:] + r.stderr[-2000:]
    d = json.loads(lines[-1])
    assert d["parity"]["ok"] and d["parity"]["max_rel_fc"] <= 1e-12 and d["parity"]["window_pairs"] > 0, d["parity"]
    assert d["contact"]["pairs_enforced"] > 1000
    assert d["contact"]["step_ms_with_contact"] < 1.34 * d["contact"]["step_ms_without_contact"], d["contact"]
    assert r.returncode in (0, 3), r.stderr[-2000:]  # (3: no nvidia-smi clock sample fell inside a 5-step timed region)


def test_driver_reports_contact_time(tmp_path):
    """The reference times contact as its own region and writes it into the timing log
    (src/integrators/explicit_time_integrator.cc:233-236, 301, 312-316; src/nimble_timing_utils.cc:70-94): the driver's
    closing summary and nimble_timing_data_*.log carry the device time of the contact kernels."""
    import glob
    import re
    import subprocess

    from nimblesm_b200.exodus_py import write_genesis

    deck, mesh, *_ = load_golden("cubes_contact")
    deck = re.sub(r"\n*$", "\n", deck) + "write timing data file: on\n"
    write_genesis(str(tmp_path / "cubes_contact.g"), mesh)
    (tmp_path / "case.in").write_text(deck)
    r = subprocess.run([EXE, "case.in"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    m = re.search(r" --- Contact time: ([0-9.eE+-]+)", r.stdout)
    assert m and float(m.group(1)) > 0.0, r.stdout[-1500:]
    assert "number of triangular contact facets (primary blocks): 1536" in r.stdout
    logs = glob.glob(str(tmp_path / "nimble_timing_data_n1_*.log"))
    assert len(logs) == 1
    cols = open(logs[0]).read().split()
    sim, force, contact = float(cols[1]), float(cols[2]), float(cols[3])
    assert 0.0 < contact < sim and force > 0.0
