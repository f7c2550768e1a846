"""nimblesm_b200/exodus_py.py — Genesis / Exodus II files from Python (scipy's NetCDF-3 64-bit-offset writer/reader).

Used by the tests and tools to (1) materialise the mesh of a golden fixture (tests/golden/*.npz) as a `.g` file
that the C++ host layer (nimblesm_b200/host) reads exactly like a cubit / SEACAS-decomp file, and (2) read the
`.out.e` files the C++ ExodusOutput writes.  Variable names follow SURVEY.md Appendix A.
"""
from __future__ import annotations

import numpy as np
from scipy.io import netcdf_file


def _put_names(f, var, dims, names, length):
    v = f.createVariable(var, "c", dims)
    arr = np.zeros((len(names), length), dtype="S1")
    for i, nm in enumerate(names):
        b = nm.encode()[:length - 1]
        arr[i, :len(b)] = np.frombuffer(b, dtype="S1")
    v[:] = arr


def write_genesis(path, mesh, name_sets=False):
    """mesh: dict as produced by tests/golden/make_golden.read_genesis (0-based ids)."""
    f = netcdf_file(path, "w", version=2)
    n_nodes = len(mesh["x"])
    all_ids = list(mesh.get("all_block_ids", mesh["block_ids"]))
    n_elem = int(sum(len(mesh["conn"][b]) for b in mesh["block_ids"]))
    f.title = "nimblesm_b200 fixture"
    f.api_version = np.float32(6.05)
    f.version = np.float32(6.05)
    f.floating_point_word_size = np.int32(8)
    f.file_size = np.int32(1)
    f.createDimension("time_step", None)  # scipy wants the record dimension first
    f.createDimension("len_string", 33)
    f.createDimension("len_line", 81)
    f.createDimension("four", 4)
    f.createDimension("len_name", 33)
    f.createDimension("num_dim", 3)
    f.createDimension("num_nodes", n_nodes)
    if n_elem:
        f.createDimension("num_elem", n_elem)
    f.createDimension("num_el_blk", len(all_ids))
    ns_ids = list(mesh["node_sets"].keys())
    if ns_ids:
        f.createDimension("num_node_sets", len(ns_ids))
    f.createVariable("time_whole", "d", ("time_step",))
    v = f.createVariable("eb_prop1", "i", ("num_el_blk",))
    v.name = "ID"
    v[:] = np.array(all_ids, dtype=np.int32)
    f.createVariable("eb_status", "i", ("num_el_blk",))[:] = np.ones(len(all_ids), np.int32)
    _put_names(f, "eb_names", ("num_el_blk", "len_name"), ["" for _ in all_ids], 33)
    for k, nm in enumerate(("coordx", "coordy", "coordz")):
        f.createVariable(nm, "d", ("num_nodes",))[:] = np.asarray(mesh["xyz"[k]], dtype=np.float64)
    f.createVariable("node_num_map", "i", ("num_nodes",))[:] = np.asarray(mesh["node_gid"], dtype=np.int32) + 1
    if n_elem:
        egid = np.concatenate([np.asarray(mesh["elem_gid"][b]) for b in mesh["block_ids"]]).astype(np.int32) + 1
        f.createVariable("elem_num_map", "i", ("num_elem",))[:] = egid
    for i, b in enumerate(all_ids):
        if b not in mesh["conn"] or len(mesh["conn"][b]) == 0:
            continue
        c = np.asarray(mesh["conn"][b], dtype=np.int32) + 1
        f.createDimension("num_el_in_blk%d" % (i + 1), c.shape[0])
        f.createDimension("num_nod_per_el%d" % (i + 1), c.shape[1])
        v = f.createVariable("connect%d" % (i + 1), "i", ("num_el_in_blk%d" % (i + 1), "num_nod_per_el%d" % (i + 1)))
        v.elem_type = "HEX8"
        v[:] = c
    if ns_ids:
        v = f.createVariable("ns_prop1", "i", ("num_node_sets",))
        v.name = "ID"
        v[:] = np.array(ns_ids, dtype=np.int32)
        f.createVariable("ns_status", "i", ("num_node_sets",))[:] = np.ones(len(ns_ids), np.int32)
        _put_names(f, "ns_names", ("num_node_sets", "len_name"),
                   [("nodelist_%d" % s) if name_sets else "" for s in ns_ids], 33)
        for i, s in enumerate(ns_ids):
            nodes = np.asarray(mesh["node_sets"][s], dtype=np.int32)
            if len(nodes) == 0:
                continue  # a set without local nodes keeps only its id
            f.createDimension("num_nod_ns%d" % (i + 1), len(nodes))
            f.createVariable("node_ns%d" % (i + 1), "i", ("num_nod_ns%d" % (i + 1),))[:] = nodes + 1
            f.createVariable("dist_fact_ns%d" % (i + 1), "d", ("num_nod_ns%d" % (i + 1),))[:] = np.ones(len(nodes))
    f.close()


def _names(var):
    out = []
    for row in var.data:
        out.append(b"".join(row).split(b"\x00")[0].decode().strip())
    return out


def read_results(path):
    """-> dict(times, nod{name: [T, n]}, elem{(name, 0-based block index): [T, ne]}, node_gid, elem_gid)."""
    f = netcdf_file(path, "r", mmap=False)
    v = f.variables
    out = {"times": np.array(v["time_whole"].data, dtype=np.float64), "nod": {}, "elem": {}}
    if "name_nod_var" in v:
        for k, nm in enumerate(_names(v["name_nod_var"])):
            out["nod"][nm] = np.array(v["vals_nod_var%d" % (k + 1)].data, dtype=np.float64)
    if "name_elem_var" in v:
        for k, nm in enumerate(_names(v["name_elem_var"])):
            for b in range(f.dimensions["num_el_blk"]):
                key = "vals_elem_var%deb%d" % (k + 1, b + 1)
                if key in v:
                    out["elem"][(nm, b)] = np.array(v[key].data, dtype=np.float64)
    out["node_gid"] = np.array(v["node_num_map"].data, dtype=np.int64) - 1 if "node_num_map" in v else None
    out["elem_gid"] = np.array(v["elem_num_map"].data, dtype=np.int64) - 1 if "elem_num_map" in v else None
    out["block_ids"] = [int(b) for b in v["eb_prop1"].data] if "eb_prop1" in v else []
    out["num_el_in_blk"] = [int(f.dimensions.get("num_el_in_blk%d" % (i + 1), 0)) for i in range(len(out["block_ids"]))]
    f.close()
    return out
