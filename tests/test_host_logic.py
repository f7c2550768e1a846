"""Host-side logic on CPU: deck surface, synthetic meshes, element partitions, shared-node tables, exodiff rules,
and the world_size-2 rendezvous of the peer-exchange tables over gloo."""
import os
import socket
import sys

import numpy as np
import pytest

from tests.conftest import ROOT, load_golden


def test_deck_parser_matches_reference_decks():
    from nimblesm_b200.deck import parse_deck

    deck, mesh, gold, ref, _ = load_golden("brick_with_fibers")
    d = parse_deck(deck)
    assert d.num_load_steps == 23 and d.output_frequency == 2 and d.final_time == 1.5e-8
    assert d.blocks == {1: "material_1", 2: "material_2"}
    m2 = d.block_material(2)
    assert (m2.model, m2.density, m2.bulk_modulus, m2.shear_modulus) == ("neohookean", 5.0, 1.333e12, 0.1379e12)
    kinds = [bc.kind for bc in d.boundary_conditions]
    assert kinds == ["initial_velocity"] + ["prescribed_velocity"] * 3
    assert d.boundary_conditions[0].node_set_id == 1 and d.boundary_conditions[0].magnitude == 1000.0


def test_deck_parser_errors_like_the_reference():
    from nimblesm_b200.deck import parse_boundary_condition, parse_deck, parse_material

    with pytest.raises(ValueError, match="unknown key"):
        parse_deck("no such key: 1\n")
    with pytest.raises(ValueError):
        parse_material("plastic density 1 bulk_modulus 1 shear_modulus 1")
    with pytest.raises(ValueError):
        parse_material("elastic density 1 bulk_modulus 1 youngs_modulus 1")
    bc = parse_boundary_condition('prescribed_velocity nodelist_3 Y "0.5*x + t"')
    assert bc.coordinate == 1 and bc.expression == "0.5*x + t" and bc.node_set_id == 3
    with pytest.raises(ValueError, match="quotes"):
        parse_boundary_condition('prescribed_velocity nodelist_3 y "0.5')


def test_structured_cube_layout():
    from nimblesm_b200.mesh import structured_cube

    n = 5
    m = structured_cube(n)
    assert len(m["x"]) == (n + 1) ** 3 and m["conn"][1].shape == (n ** 3, 8)
    c = m["conn"][1][0]
    p = np.stack([m["x"][c], m["y"][c], m["z"][c]], 1) * n
    assert np.array_equal(p, [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]])
    assert len(m["node_sets"][2]) == (n + 1) ** 2 and np.all(m["x"][m["node_sets"][2]] == 0.0)
    # every interior node belongs to 8 elements
    cnt = np.bincount(m["conn"][1].ravel(), minlength=len(m["x"]))
    assert cnt.max() == 8 and cnt.min() == 1


@pytest.mark.parametrize("world", [2, 4, 8])
def test_cube_partition_and_shared_tables(world):
    from nimblesm_b200.mesh import brick_surface_gids, cube_partition, shared_node_tables, structured_cube

    n = 6
    whole = structured_cube(n)
    parts = [cube_partition(n, world, r) for r in range(world)]
    assert sum(len(p["conn"][1]) for p in parts) == n ** 3
    egid = np.sort(np.concatenate([p["elem_gid"][1] for p in parts]))
    assert np.array_equal(egid, np.arange(n ** 3))
    for p in parts:  # coordinates of duplicated nodes are bit-identical to the whole mesh
        assert np.array_equal(p["x"], whole["x"][p["node_gid"]])
    cands = [brick_surface_gids(p) for p in parts]
    full = [p["node_gid"] for p in parts]
    tables = [shared_node_tables(r, cands, parts[r]["node_gid"]) for r in range(world)]
    tables_full = [shared_node_tables(r, full, parts[r]["node_gid"]) for r in range(world)]
    for r in range(world):
        peers, offs, nodes = tables[r]
        pf, of, nf = tables_full[r]
        assert np.array_equal(peers, pf) and np.array_equal(offs, of) and np.array_equal(nodes, nf)
        for i, p in enumerate(peers):
            mine = parts[r]["node_gid"][nodes[offs[i]:offs[i + 1]]]
            pp, po, pn = tables[p]
            j = list(pp).index(r)
            theirs = parts[p]["node_gid"][pn[po[j]:po[j + 1]]]
            assert np.array_equal(mine, theirs) and np.all(np.diff(mine) > 0)  # same order on both sides


def test_rcb_partition_matches_reference_pieces_semantics():
    """RCB pieces keep every node their elements touch; node sets are restricted to local nodes; shared nodes are
    exactly the nodes with the same global id (Appendix B of SURVEY.md)."""
    from nimblesm_b200.mesh import rcb_partition

    deck, mesh, gold, ref, pieces = load_golden("wave_in_bar")
    parts = rcb_partition(mesh, 2)
    assert sum(len(p["conn"][1]) for p in parts) == len(mesh["conn"][1])
    shared = np.intersect1d(parts[0]["node_gid"], parts[1]["node_gid"])
    ref_shared = np.intersect1d(pieces[(2, 0)]["node_gid"], pieces[(2, 1)]["node_gid"])
    assert len(shared) == len(ref_shared) == 4
    for p in parts:
        gx = dict(zip(mesh["node_gid"].tolist(), mesh["x"].tolist()))
        assert all(gx[g] == x for g, x in zip(p["node_gid"].tolist(), p["x"].tolist()))
        for sid, ns in p["node_sets"].items():
            assert set(p["node_gid"][ns]) <= set(np.asarray(mesh["node_gid"])[mesh["node_sets"][sid]])


def test_exodiff_rules():
    from nimblesm_b200 import exodiff

    spec = ("TIME STEPS relative 1.e-6 floor 0.0\nNODAL VARIABLES relative 1.e-6 floor 0.0\n"
            "\tdisplacement_x absolute 1.0e-3\n\tvelocity_x\n"
            "ELEMENT VARIABLES relative 1.e-6 floor 1.0\n")
    gold = {"times": np.array([0.0, 1.0]), "nod": {"displacement_x": np.array([[1.0]]), "velocity_x": np.array([[2.0]]),
                                                    "skipped": np.array([[5.0]])},
            "elem": {("s", 0): np.array([[0.5, 10.0]])}}
    test = {"times": np.array([0.0, 1.0]), "nod": {"displacement_x": np.array([[1.0005]]),
                                                    "velocity_x": np.array([[2.0 + 1e-7]]), "skipped": np.array([[50.0]])},
            "elem": {("s", 0): np.array([[0.9, 10.0 + 1e-7]])}}
    assert exodiff.compare(spec, gold, test) == []
    test["nod"]["velocity_x"] = np.array([[2.1]])
    assert len(exodiff.compare(spec, gold, test)) == 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from nimblesm_b200.mesh import brick_surface_gids, cube_partition, reference_shared_sum, shared_node_tables

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    part = cube_partition(n, world, rank)
    cand = brick_surface_gids(part)
    sizes = [None] * world
    dist.all_gather_object(sizes, int(len(cand)))
    buf = torch.full((max(sizes),), -1, dtype=torch.int64)
    buf[:len(cand)] = torch.from_numpy(np.ascontiguousarray(cand))
    allb = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(allb, buf)
    cands = [b.numpy()[:s] for b, s in zip(allb, sizes)]
    peers, offs, nodes = shared_node_tables(rank, cands, part["node_gid"])
    # emulate the device exchange on the host: send partial values to peers, sum in ascending rank order
    rng = np.random.default_rng(100 + rank)
    val = rng.random((len(part["x"]), 3))
    gathered = [None] * world
    dist.all_gather_object(gathered, (part["node_gid"], val))
    total = val.copy()
    contrib = {}
    for i, p in enumerate(peers):
        pg, pv = gathered[p]
        loc = nodes[offs[i]:offs[i + 1]]
        pos = np.searchsorted(pg, part["node_gid"][loc])
        for ln, pp in zip(loc, pos):
            contrib.setdefault(int(ln), []).append((int(p), pv[pp]))
    for ln, lst in contrib.items():
        lst.append((rank, val[ln]))
        lst.sort(key=lambda t: t[0])
        s = lst[0][1].copy()
        for _, x in lst[1:]:
            s = s + x
        total[ln] = s
    want = reference_shared_sum([g[1] for g in gathered], [g[0] for g in gathered])[rank]
    q.put((rank, bool(np.array_equal(total, want)), len(peers), int(offs[-1])))
    dist.barrier()
    dist.destroy_process_group()


def test_shared_node_sum_world_size_2_gloo():
    """N > 1 host path on CPU: two ranks rendezvous over gloo, exchange surface candidates, build the peer tables,
    and the rank-ordered shared-node sum equals VectorReduction's result on every holder (bit-identical replicas)."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world, n = 2, 4
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == [True, True]
    assert res[0][2] == 1 and res[0][3] == (n + 1) ** 2  # one peer, one shared face of (n+1)^2 nodes


def test_lattice_window_determines_its_complete_nodes(oracle):
    """mesh.lattice_window (the sampled parity checks of the 64 M-element runs): the oracle force of a window equals
    the oracle force of the whole mesh on every node the window marks `complete`, bit for bit."""
    from nimblesm_b200.mesh import lattice_window
    from tests.conftest import perturbed_cube

    mesh, ref, disp = perturbed_cube(7, 1e-2)
    conn = mesh["conn"][1]
    f_all, ed_all = oracle.internal_force(oracle.NEOHOOKEAN, 1.6e12, 0.8e12, ref, disp, conn, True)
    for lo, w in (((0, 0, 0), 3), ((2, 1, 3), 4), ((4, 4, 4), 3), ((0, 3, 4), 3)):
        nodes, elems, conn_w, complete = lattice_window((7, 7, 7), lo, w)
        assert np.array_equal(nodes[conn_w], conn[elems])
        f_w, ed_w = oracle.internal_force(oracle.NEOHOOKEAN, 1.6e12, 0.8e12, np.ascontiguousarray(ref[nodes]),
                                          np.ascontiguousarray(disp[nodes]), conn_w, True)
        assert np.array_equal(ed_w.view(np.int64), ed_all[elems].view(np.int64))
        assert complete.any() and not complete.all()
        assert np.array_equal(f_w[complete].view(np.int64), f_all[nodes][complete].view(np.int64))
        assert not np.array_equal(f_w[~complete], f_all[nodes][~complete])


class _FakeContext:
    """Stands in for capi.Context.download in bench.parity_check (no GPU in the CPU suite)."""

    def __init__(self, fields):
        self.fields = fields

    def download(self, name):
        return self.fields[name]


def _parity_worker(rank, world, port, n, corrupt, q):
    sys.path.insert(0, ROOT)
    import argparse

    import torch.distributed as dist

    import bench
    from oracle import hex8 as oracle

    dist_mod = None
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        dist_mod = dist
    grid = (world, 1, 1)
    bricks = [bench.weak_brick(n, grid, (r, 0, 0)) for r in range(world)]
    # the global mesh, joined by global node id, and a smooth displacement of the global coordinates
    n_glob = int(max(b["node_gid"].max() for b in bricks)) + 1
    X = np.zeros((n_glob, 3))
    conn = []
    for b in bricks:
        X[b["node_gid"]] = np.stack([b["x"], b["y"], b["z"]], 1)
        conn.append(b["node_gid"][b["conn"][1]])
    conn = np.ascontiguousarray(np.concatenate(conn), dtype=np.int32)
    u = 1e-3 * np.stack([np.sin(3 * X[:, 1]) * X[:, 0], X[:, 2] * X[:, 0] ** 2, np.cos(2 * X[:, 0]) * X[:, 1]], 1)
    f, _ = oracle.internal_force(oracle.NEOHOOKEAN, bench.BULK, bench.SHEAR, X, u, conn, False)
    mine = bricks[rank]
    g = mine["node_gid"]
    fields = {"displacement": u[g].copy(), "velocity": (2.0 * u[g]).copy(), "internal_force": f[g].copy()}
    if corrupt == "replica" and rank == world - 1:
        fields["velocity"][mine["surface_idx"][0], 1] += 1e-300  # one replica of one shared node, one bit pattern off
    if corrupt == "force":  # every rank alike: replicas stay equal, the oracle comparison must catch it
        fields["internal_force"] *= 1.0 + 1e-9
    args = argparse.Namespace(workload="cube", material="neohookean")
    res = bench.parity_check(_FakeContext(fields), mine, args, n, grid, (rank, 0, 0), rank, world, dist_mod)
    q.put((rank, res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("world,corrupt,ok", [(1, None, True), (2, None, True), (2, "replica", False), (2, "force", False)])
def test_bench_parity_block(world, corrupt, ok):
    """bench.parity_check (the `parity` object of the bench line, N = 1 and N = 2 over gloo): passes on the oracle's
    own forces distributed over the ranks' bricks, fails when one replica of a shared node differs by one bit
    pattern or when the forces are off by 1e-9."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_parity_worker, args=(r, world, port, 12, corrupt, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    res = got[0]
    assert res["checked"] and res["ok"] is ok, res
    if world > 1:
        assert res["shared_node_replicas_compared"] == 13 * 13 and res["replicas_bit_equal"] is (corrupt != "replica")
        assert "straddling" in res["window"]
    if corrupt is None:
        assert res["max_rel_f"] <= 1e-13
    assert got.get(1) is None


def test_bench_contact_stack_entities_match_the_generic_skinner():
    """bench.py builds the contact entities of its two-body workload analytically (six sides per body, Exodus face orders);
    they must be the skin the oracle's generic skinner finds -- same quads with the same cyclic orientation (outward
    normals), same contact nodes, same characteristic lengths."""
    import bench
    from oracle.contact import ContactSetup

    mesh, ent, h = bench.contact_stack(10)
    cs = ContactSetup(mesh, [2], [1], 1.0)
    want = {tuple(sorted(q)): list(q) for q in cs.primary_quads.tolist()}
    assert len(ent["primary_quads"]) == len(want)
    for q, l in zip(ent["primary_quads"].tolist(), ent["primary_char_len"]):
        w = want[tuple(sorted(q))]
        i = w.index(q[0])
        assert w[i:] + w[:i] == q
        assert l == cs.primary_char_len[[tuple(sorted(x)) for x in cs.primary_quads.tolist()].index(tuple(sorted(q)))]
    assert dict(zip(ent["contact_nodes"].tolist(), ent["contact_node_char_len"])) == dict(zip(cs.contact_nodes.tolist(), cs.contact_node_char_len))
    assert abs(h - 0.1) < 1e-15


def test_exodiff_rounding_floor():
    """An absolute tolerance far below the rounding of the data (the contact decks ask for 2e-4 on forces of 3e8) can be
    floored at a multiple of the variable's magnitude; relative tolerances and honest absolute ones are untouched."""
    from nimblesm_b200 import exodiff

    spec = "NODAL VARIABLES relative 1.e-6 floor 0.0\n\tf_x absolute 2.0e-4\n\tu_x absolute 1.0e-8\n"
    gold = {"times": np.array([0.0, 1.0]), "nod": {"f_x": np.array([[0.0, 3.0e8], [1.0e8, 2.0e8]]), "u_x": np.array([[0.0, 1.0e-5], [0.0, 2.0e-5]])}, "elem": {}}
    test = {"times": gold["times"].copy(), "nod": {"f_x": gold["nod"]["f_x"] + 3.0e-4, "u_x": gold["nod"]["u_x"] + 5.0e-9}, "elem": {}}
    assert exodiff.compare(spec, gold, test)  # 3e-4 > 2e-4
    assert not exodiff.compare(spec, gold, test, rounding_floor=1e-11)  # 3e-4 < 1e-11 * 3e8 = 3e-3
    test["nod"]["u_x"] = gold["nod"]["u_x"] + 2.0e-8
    fails = exodiff.compare(spec, gold, test, rounding_floor=1e-11)
    assert len(fails) == 1 and "u_x" in fails[0]  # (1e-11 * 2e-5 is far below the 1e-8 the file asks for)
