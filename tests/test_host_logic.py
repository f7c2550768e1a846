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
