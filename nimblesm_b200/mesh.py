"""nimblesm_b200/mesh.py — hex8 meshes for the B200 path: synthetic structured cubes, element partitions and
the shared-node tables of the peer exchange.

Mesh dict layout (same as tests/golden/make_golden.py::read_genesis, i.e. what GenesisMesh holds,
src/nimble_genesis_mesh.h:328-355): x, y, z [n_nodes] fp64; node_gid [n_nodes] 0-based global ids;
block_ids [..]; conn {block_id: int32 [n_elem, 8]} 0-based LOCAL node ids in Exodus hex8 order;
elem_gid {block_id: [n_elem]}; node_sets {id: int32 local node ids}.

Partitioning follows the reference's domain decomposition model: elements are
split, every part keeps all nodes its elements touch, shared nodes are duplicated and identified only through
their global ids (src/nimble.mpi.reduction.cc:50-123).
"""
from __future__ import annotations

import numpy as np

# Exodus hex8 corner offsets (i, j, k): bottom face counter-clockwise, then top (src/nimble_element.cc:113-120)
HEX_CORNERS = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]])


def structured_brick(n_global, lo=(0, 0, 0), hi=None, block_of_element=None, length=1.0):
    """Sub-brick [lo, hi) (element indices) of the n_global^3-element cube [0, length]^3.

    Node (i, j, k) of the global lattice has id i + (n+1)(j + (n+1)k) and coordinate (i/n, j/n, k/n)*length;
    element (i, j, k) has id i + n(j + nk).  Node sets: 1 = all nodes, 2 = face x = 0,
    3 = face x = length, 4 = y = 0, 5 = z = 0.  `block_of_element(ei, ej, ek) -> block id array` splits blocks.
    """
    n = int(n_global)
    hi = (n, n, n) if hi is None else hi
    ex, ey, ez = (hi[d] - lo[d] for d in range(3))
    nx, ny, nz = ex + 1, ey + 1, ez + 1
    gi = np.arange(lo[0], hi[0] + 1, dtype=np.int64)
    gj = np.arange(lo[1], hi[1] + 1, dtype=np.int64)
    gk = np.arange(lo[2], hi[2] + 1, dtype=np.int64)
    K, J, I = np.meshgrid(gk, gj, gi, indexing="ij")  # local node index = i + nx (j + ny k)
    node_gid = (I + (n + 1) * (J + (n + 1) * K)).ravel()
    x = (I.ravel() / n) * length
    y = (J.ravel() / n) * length
    z = (K.ravel() / n) * length
    ek, ej, ei = np.meshgrid(np.arange(ez), np.arange(ey), np.arange(ex), indexing="ij")
    ei, ej, ek = ei.ravel(), ej.ravel(), ek.ravel()
    conn = np.empty((len(ei), 8), dtype=np.int32)
    for c, (di, dj, dk) in enumerate(HEX_CORNERS):
        conn[:, c] = (ei + di) + nx * ((ej + dj) + ny * (ek + dk))
    elem_gid = (ei + lo[0]) + n * ((ej + lo[1]) + n * (ek + lo[2]))
    if block_of_element is None:
        blk = np.ones(len(ei), dtype=np.int64)
    else:
        blk = np.asarray(block_of_element(ei + lo[0], ej + lo[1], ek + lo[2]))
    block_ids = sorted(int(b) for b in np.unique(blk))
    mesh = dict(x=x, y=y, z=z, node_gid=node_gid, block_ids=block_ids, all_block_ids=block_ids, conn={}, elem_gid={},
                node_sets={})
    for b in block_ids:
        sel = blk == b
        mesh["conn"][b] = np.ascontiguousarray(conn[sel])
        mesh["elem_gid"][b] = elem_gid[sel]
    Ir, Jr, Kr = I.ravel(), J.ravel(), K.ravel()
    allnodes = np.arange(len(x), dtype=np.int32)
    mesh["node_sets"] = {1: allnodes, 2: allnodes[Ir == 0], 3: allnodes[Ir == n], 4: allnodes[Jr == 0],
                         5: allnodes[Kr == 0]}
    mesh["lattice"] = dict(n=n, lo=tuple(lo), hi=tuple(hi))
    return mesh


def structured_cube(n, block_of_element=None, length=1.0):
    return structured_brick(n, (0, 0, 0), (n, n, n), block_of_element, length)


def brick_grid(world_size):
    """(px, py, pz), px >= py >= pz, px*py*pz = world_size, as cubic as possible (what RCB of a cube yields):
    1 -> (1,1,1), 2 -> (2,1,1), 4 -> (2,2,1), 8 -> (2,2,2)."""
    best = None
    for px in range(1, world_size + 1):
        for py in range(1, px + 1):
            for pz in range(1, py + 1):
                if px * py * pz == world_size and (best is None or px - pz < best[0] - best[2]):
                    best = (px, py, pz)
    return best


def cube_partition(n_global, world_size, rank, block_of_element=None, length=1.0):
    """Rank's brick of the n_global^3 cube split into brick_grid(world_size) parts (weak-scaling layouts build
    n_global = n_per_gpu * grid)."""
    px, py, pz = brick_grid(world_size)
    rx, ry, rz = rank % px, (rank // px) % py, rank // (px * py)
    n = n_global

    def span(r, p):
        return (n * r) // p, (n * (r + 1)) // p

    (x0, x1), (y0, y1), (z0, z1) = span(rx, px), span(ry, py), span(rz, pz)
    return structured_brick(n, (x0, y0, z0), (x1, y1, z1), block_of_element, length)


def brick_surface_gids(mesh):
    """Global ids of the brick's nodes that can be shared with another part: nodes on a brick face that is
    interior to the global cube (candidate set exchanged between ranks instead of all node ids)."""
    lat = mesh["lattice"]
    n, lo, hi = lat["n"], lat["lo"], lat["hi"]
    nx, ny, nz = (hi[d] - lo[d] + 1 for d in range(3))
    idx = np.arange(nx * ny * nz)
    i, j, k = idx % nx, (idx // nx) % ny, idx // (nx * ny)
    m = np.zeros(len(idx), dtype=bool)
    for loc, size, d in ((i, nx, 0), (j, ny, 1), (k, nz, 2)):
        if lo[d] > 0:
            m |= loc == 0
        if hi[d] < n:
            m |= loc == size - 1
    return mesh["node_gid"][m]


def rcb_partition(mesh, n_parts):
    """Recursive coordinate bisection of ELEMENTS by centroid (what SEACAS decomp produces for the reference,
    test/_wip/scaling_study/decomp.sh).  Returns one local mesh dict per part; local nodes/elements keep
    ascending global order."""
    cent, owner_blk, owner_idx = [], [], []
    for b in mesh["block_ids"]:
        c = mesh["conn"][b]
        cent.append(np.stack([mesh["x"][c].mean(1), mesh["y"][c].mean(1), mesh["z"][c].mean(1)], 1))
        owner_blk.append(np.full(len(c), b))
        owner_idx.append(np.arange(len(c)))
    cent = np.concatenate(cent) if cent else np.zeros((0, 3))
    owner_blk, owner_idx = np.concatenate(owner_blk), np.concatenate(owner_idx)
    part = np.zeros(len(cent), dtype=np.int64)

    def split(ids, p0, np_):
        if np_ == 1:
            part[ids] = p0
            return
        ext = cent[ids].max(0) - cent[ids].min(0)
        d = int(np.argmax(ext))
        order = ids[np.argsort(cent[ids, d], kind="stable")]
        left = np_ // 2
        cut = (len(order) * left) // np_
        split(order[:cut], p0, left)
        split(order[cut:], p0 + left, np_ - left)

    split(np.arange(len(cent)), 0, n_parts)
    out = []
    for p in range(n_parts):
        sel = part == p
        local = dict(block_ids=[], all_block_ids=list(mesh["all_block_ids"]), conn={}, elem_gid={}, node_sets={})
        used = []
        for b in mesh["block_ids"]:
            eidx = np.sort(owner_idx[sel & (owner_blk == b)])
            if len(eidx) == 0:
                continue
            local["block_ids"].append(b)
            local["conn"][b] = mesh["conn"][b][eidx]
            local["elem_gid"][b] = mesh["elem_gid"][b][eidx]
            used.append(local["conn"][b].ravel())
        nodes = np.unique(np.concatenate(used)) if used else np.zeros(0, np.int64)
        remap = -np.ones(len(mesh["x"]), dtype=np.int64)
        remap[nodes] = np.arange(len(nodes))
        for b in local["block_ids"]:
            local["conn"][b] = np.ascontiguousarray(remap[local["conn"][b]], dtype=np.int32)
        local["x"], local["y"], local["z"] = mesh["x"][nodes], mesh["y"][nodes], mesh["z"][nodes]
        local["node_gid"] = np.asarray(mesh["node_gid"])[nodes]
        for sid, ns in mesh["node_sets"].items():
            keep = remap[ns]
            local["node_sets"][sid] = np.ascontiguousarray(keep[keep >= 0], dtype=np.int32)
        out.append(local)
    return out


def shared_node_tables(rank, candidate_gids_by_rank, local_gid):
    """Peer tables of `rank` for nsm_b200_comm_init.

    candidate_gids_by_rank[r]: global ids rank r may share (all its node ids, or its surface candidates);
    local_gid: this rank's node_gid array.  Returns (peer_ranks, pair_offsets, pair_local_nodes): for each peer
    the local ids of the common nodes sorted by GLOBAL id, so that both sides enumerate them alike
    (src/nimble.mpi.reduction.cc:114-120).
    """
    local_gid = np.asarray(local_gid)
    order = np.argsort(local_gid, kind="stable")
    sorted_gid = local_gid[order]
    mine = np.asarray(candidate_gids_by_rank[rank])
    peers, offs, nodes = [], [0], []
    for r, theirs in enumerate(candidate_gids_by_rank):
        if r == rank:
            continue
        common = np.intersect1d(mine, np.asarray(theirs))  # sorted unique
        if len(common) == 0:
            continue
        pos = np.searchsorted(sorted_gid, common)
        assert np.all(sorted_gid[pos] == common), "candidate id not present among this rank's nodes"
        peers.append(r)
        nodes.append(order[pos])
        offs.append(offs[-1] + len(common))
    pair_nodes = np.concatenate(nodes).astype(np.int32) if nodes else np.zeros(0, np.int32)
    return np.asarray(peers, dtype=np.int32), np.asarray(offs, dtype=np.int64), pair_nodes


def reference_shared_sum(values_by_rank, gids_by_rank):
    """Host statement of VectorCommunicator::VectorReduction for tests: every shared node receives the sum of
    all holders' values, added in ascending rank order (the order the device exchange fixes)."""
    acc = {}
    for r, (vals, gids) in enumerate(zip(values_by_rank, gids_by_rank)):
        for g, v in zip(np.asarray(gids).tolist(), np.asarray(vals)):
            acc[g] = v.copy() if g not in acc else acc[g] + v
    return [np.stack([acc[g] for g in np.asarray(gids).tolist()]) if len(gids) else np.asarray(vals)
            for vals, gids in zip(values_by_rank, gids_by_rank)]


def lattice_window(n_elem_axes, lo, w, at_global_lo=(True, True, True), at_global_hi=(True, True, True)):
    """A w^3-element window of a structured brick (local node id i + nx (j + ny k), local element id
    ei + ex (ej + ey ek), as structured_brick / bench.weak_brick number them) with its lower corner at local
    element (lo[0], lo[1], lo[2]).

    Returns (node_ids [(w+1)^3], elem_ids [w^3], conn [w^3, 8] into node_ids, complete [(w+1)^3] bool): `complete`
    marks the window nodes ALL of whose elements lie inside the window -- strictly inside it, or on a window face
    that coincides with a face of the whole body (at_global_lo / at_global_hi say whether the brick's own faces are
    such) -- i.e. the nodes whose assembled internal force the window alone determines.  Used by the sampled parity
    checks at sizes where the CPU oracle cannot run the whole mesh (tests/test_gpu_parity.py, bench.py)."""
    ex, ey, ez = (int(v) for v in n_elem_axes)
    nx, ny = ex + 1, ey + 1
    a = np.arange(w + 1, dtype=np.int64)
    K, J, I = np.meshgrid(a + lo[2], a + lo[1], a + lo[0], indexing="ij")
    node_ids = (I + nx * (J + ny * K)).ravel()
    e = np.arange(w, dtype=np.int64)
    EK, EJ, EI = np.meshgrid(e, e, e, indexing="ij")
    EI, EJ, EK = EI.ravel(), EJ.ravel(), EK.ravel()
    elem_ids = (EI + lo[0]) + ex * ((EJ + lo[1]) + ey * (EK + lo[2]))
    conn = np.empty((w ** 3, 8), dtype=np.int32)
    for c, (di, dj, dk) in enumerate(HEX_CORNERS):
        conn[:, c] = (EI + di) + (w + 1) * ((EJ + dj) + (w + 1) * (EK + dk))
    complete = np.ones(node_ids.shape, dtype=bool)
    for loc, l0, n_ax, glo, ghi in ((I, lo[0], ex, at_global_lo[0], at_global_hi[0]), (J, lo[1], ey, at_global_lo[1], at_global_hi[1]),
                                    (K, lo[2], ez, at_global_lo[2], at_global_hi[2])):
        loc = loc.ravel()
        inside = (loc > l0) & (loc < l0 + w)
        inside |= (loc == l0) & (l0 == 0) & bool(glo)
        inside |= (loc == l0 + w) & (l0 + w == n_ax) & bool(ghi)
        complete &= inside
    return node_ids, elem_ids, conn, complete
