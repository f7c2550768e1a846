#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference's own regression cases (run in the build container only).

For every hot-path deck under /root/reference/test/dynamics this script stores, in one compressed npz:
  * the mesh read from <case>.g (NetCDF-3 / Exodus II; read with scipy.io.netcdf_file),
  * the deck text (<case>.in),
  * the reference's gold results <case>.gold.e (times, nodal and element variables, by name)
    -- these pin the oracle at the reference's own exodiff tolerance (1e-6 * max|gold|),
  * `ref_*`: snapshots produced HERE by the reference's serial code (oracle/_ref/libnimble_ref.so) on the
    same deck -- these pin the GPU path at the north-star tolerances when /root/reference is absent.
Decomposed pieces (<case>.g.<P>.<r>) are stored for the multi-rank shared-node tests.

Usage:  python tests/golden/make_golden.py [/root/reference]
"""
import glob
import os
import sys

import numpy as np
from scipy.io import netcdf_file

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CASES = [
    ("wave_in_bar", "wave_in_bar", "wave_in_bar"),
    ("notched_plate_native_neohookean", "notched_plate_native_neohookean", "notched_plate_native_neohookean"),
    ("notched_plate_native_hypoelastic", "notched_plate_native_hypoelastic", "notched_plate_native_hypoelastic"),
    ("brick_with_fibers", "brick_with_fibers", "brick_with_fibers"),
    ("simple_deformation_modes", "simple_deformation_modes", "simple_deformation_modes"),
    ("rigid_body_motion", "rigid_body_motion", "rigid_body_motion"),
    ("single_elem_complex_displacement", "single_elem_complex_displacement", "single_elem_complex_displacement"),
    ("single_elem_complex_displacement", "single_elem_native_neohookean", "single_elem"),
]
# penalty-contact decks (test/contact; the reference runs them only in its Kokkos + ArborX / BVH builds).  Their ref_*
# snapshots come from the reference's serial model data + the reference's own ContactEntity objects behind the
# restated pair loop (oracle/ref_contact.cc); sliding_contact ships no gold file.
CONTACT_CASES = [
    ("cubes_contact", "cubes_contact", "cubes_contact"),
    ("sphere_plate_contact", "sphere_plate_contact", "sphere_plate_contact"),
    ("sliding_contact", "sliding_contact", "sliding_contact"),
    # five blocks, two primary and two secondary; its gold file is the reference's CONTACT VISUALISATION database (one
    # triangle element per contact facet with its own three nodes, one sphere element per contact node, the entities'
    # displacement): the reference's own record of which entities CreateContactEntities makes, in which order
    ("contact_entity_creation", "contact_entity_creation", "contact_entity_creation"),
]


def _names(var):
    out = []
    for row in var.data:
        s = b"".join(row).split(b"\x00")[0].decode().strip()
        out.append(s)
    return out


def read_genesis(path):
    """-> mesh dict: x,y,z, node_gid (0-based), block_ids, conn{b}[ne,8] 0-based, elem_gid{b}, node_sets{id}."""
    f = netcdf_file(path, "r", mmap=False)
    v = f.variables
    n_nodes = f.dimensions["num_nodes"]
    if "coordx" in v:
        x, y, z = (np.array(v[k].data, dtype=np.float64) for k in ("coordx", "coordy", "coordz"))
    else:
        c = np.array(v["coord"].data, dtype=np.float64)
        x, y, z = c[0], c[1], c[2]
    node_gid = (np.array(v["node_num_map"].data, dtype=np.int32) - 1 if "node_num_map" in v
                else np.arange(n_nodes, dtype=np.int32))
    # decomp's original_global_id_map wins (src/nimble_genesis_mesh.cc:123-149)
    if "nmap_names" in v:
        for i, nm in enumerate(_names(v["nmap_names"])):
            if nm == "original_global_id_map":
                node_gid = np.array(v["node_map%d" % (i + 1)].data, dtype=np.int32) - 1
    n_elem = f.dimensions["num_elem"]
    elem_gid_all = (np.array(v["elem_num_map"].data, dtype=np.int32) - 1 if "elem_num_map" in v
                    else np.arange(n_elem, dtype=np.int32))
    block_ids = [int(b) for b in v["eb_prop1"].data]
    conn, elem_gid = {}, {}
    off = 0
    local_blocks = []
    for i, b in enumerate(block_ids):
        key = "connect%d" % (i + 1)
        if key not in v:
            continue
        c = np.array(v[key].data, dtype=np.int32) - 1
        conn[b] = np.ascontiguousarray(c)
        elem_gid[b] = elem_gid_all[off:off + len(c)].copy()
        off += len(c)
        local_blocks.append(b)
    node_sets = {}
    if "ns_prop1" in v:
        for i, sid in enumerate(v["ns_prop1"].data):
            key = "node_ns%d" % (i + 1)
            node_sets[int(sid)] = (np.array(v[key].data, dtype=np.int32) - 1 if key in v
                                   else np.zeros(0, np.int32))
    f.close()
    return dict(x=x, y=y, z=z, node_gid=node_gid, block_ids=local_blocks, all_block_ids=block_ids, conn=conn,
                elem_gid=elem_gid, node_sets=node_sets)


def read_results(path):
    f = netcdf_file(path, "r", mmap=False)
    v = f.variables
    times = np.array(v["time_whole"].data, dtype=np.float64)
    nod = {}
    if "name_nod_var" in v:
        for k, nm in enumerate(_names(v["name_nod_var"])):
            nod[nm] = np.array(v["vals_nod_var%d" % (k + 1)].data, dtype=np.float64)
    elem = {}
    if "name_elem_var" in v:
        nblk = f.dimensions["num_el_blk"]
        for k, nm in enumerate(_names(v["name_elem_var"])):
            for b in range(nblk):
                key = "vals_elem_var%deb%d" % (k + 1, b + 1)
                if key in v:
                    elem[(nm, b)] = np.array(v[key].data, dtype=np.float64)
    f.close()
    return times, nod, elem


def pack_mesh(prefix, m, out):
    out[prefix + "x"], out[prefix + "y"], out[prefix + "z"] = m["x"], m["y"], m["z"]
    out[prefix + "node_gid"] = m["node_gid"]
    out[prefix + "block_ids"] = np.array(m["block_ids"], dtype=np.int32)
    out[prefix + "all_block_ids"] = np.array(m["all_block_ids"], dtype=np.int32)
    for b in m["block_ids"]:
        out[prefix + "conn_%d" % b] = m["conn"][b]
        out[prefix + "elem_gid_%d" % b] = m["elem_gid"][b]
    out[prefix + "ns_ids"] = np.array(list(m["node_sets"].keys()), dtype=np.int32)
    for sid, nodes in m["node_sets"].items():
        out[prefix + "ns_%d" % sid] = nodes


def unpack_mesh(prefix, z):
    m = dict(x=z[prefix + "x"], y=z[prefix + "y"], z=z[prefix + "z"], node_gid=z[prefix + "node_gid"],
             block_ids=[int(b) for b in z[prefix + "block_ids"]],
             all_block_ids=[int(b) for b in z[prefix + "all_block_ids"]], conn={}, elem_gid={}, node_sets={})
    for b in m["block_ids"]:
        m["conn"][b] = z[prefix + "conn_%d" % b]
        m["elem_gid"][b] = z[prefix + "elem_gid_%d" % b]
    for sid in z[prefix + "ns_ids"]:
        m["node_sets"][int(sid)] = z[prefix + "ns_%d" % sid]
    return m


def load_case(name):
    """Loader used by the tests: -> (deck_text, mesh, gold dict, ref dict, pieces dict)."""
    z = np.load(os.path.join(HERE, name + ".npz"), allow_pickle=False)
    mesh = unpack_mesh("mesh_", z)
    gold = {"times": z["gold_times"], "nod": {}, "elem": {}}
    gold["vis"] = {k[len("gold_vis_"):]: z[k] for k in z.files if k.startswith("gold_vis_")}  # contact visualisation database
    for k in z.files:
        if k.startswith("gold_nod_"):
            gold["nod"][k[len("gold_nod_"):]] = z[k]
        elif k.startswith("gold_elem_"):
            nm, b = k[len("gold_elem_"):].rsplit("_eb", 1)
            gold["elem"][(nm, int(b))] = z[k]
    ref = {k[len("ref_"):]: z[k] for k in z.files if k.startswith("ref_")}
    pieces = {}
    for k in z.files:
        if k.startswith("piece_") and k.endswith("_x"):
            tag = k[len("piece_"):-2]  # "P.r"
            P, r = tag.split(".")
            pieces[(int(P), int(r))] = unpack_mesh("piece_%s_" % tag, z)
    gold["exodiff"] = str(z["exodiff"]) if "exodiff" in z.files else ""
    return str(z["deck"]), mesh, gold, ref, pieces


STATE_CUBE_DECK = """genesis input file:               state_cube.g
exodus output file:               state_cube.e
final time:                       4.4e-6
number of load steps:             60
output frequency:                 20
output fields:                    displacement velocity internal_force stress equivalent_plastic_strain ipt03_equivalent_plastic_strain ipt08_von_mises_stress ipt01_stress volume
material parameters:              material_1 j2_plasticity density 7.8 bulk_modulus 1.6e12 shear_modulus 0.8e12 yield_stress 5.0e8 hardening_modulus 2.0e10
material parameters:              material_2 neohookean density 7.8 bulk_modulus 1.6e12 shear_modulus 0.8e12
element block:                    block_1 material_1
element block:                    block_2 material_2
boundary condition:               initial_velocity nodelist_1 x "1000.0*x"
boundary condition:               prescribed_velocity nodelist_2 x 0.0
boundary condition:               prescribed_velocity nodelist_2 y 0.0
boundary condition:               prescribed_velocity nodelist_2 z 0.0
"""


def make_state_case():
    """state_cube: NOT a deck of the reference (it ships no material with state variables).  A 6^3 cube whose lower
    half (block 1) is the history-dependent test material behind the reference's own Material virtuals
    (oracle/ref_state_material.cc) and whose upper half (block 2) is neohookean; the snapshots come from the
    reference's unmodified block / element-data / UpdateStates code, so they pin the B200 state-variable slot."""
    from nimblesm_b200.mesh import structured_cube
    from oracle import refdrive

    mesh = structured_cube(6, block_of_element=lambda i, j, k: np.where(k < 3, 1, 2))
    out = {}
    pack_mesh("mesh_", mesh, out)
    out["deck"] = np.array(STATE_CUBE_DECK)
    out["exodiff"] = np.array("")
    out["gold_times"] = np.zeros(0)
    run = refdrive.RefRun(STATE_CUBE_DECK, mesh, keep_snapshots=True)
    out["ref_critical_dt"] = np.array(run.begin())
    run.advance(60)
    snaps = run.snapshots()
    out["ref_times"] = np.array([s["time"] for s in snaps])
    for lbl in snaps[0]["node"]:
        out["ref_node_" + lbl] = np.stack([s["node"][lbl] for s in snaps])
    for b in mesh["block_ids"]:
        out["ref_elem_%d" % b] = np.stack([s["elem"][b] for s in snaps])  # every output step: [times][ne][8][stride]
        out["ref_elem_last_%d" % b] = snaps[-1]["elem"][b]
        for lbl in snaps[0]["derived"][b]:
            out["ref_derived_%d_%s" % (b, lbl)] = np.stack([s["derived"][b][lbl] for s in snaps])
    run.close()
    path = os.path.join(HERE, "state_cube.npz")
    np.savez_compressed(path, **out)
    eqps = out["ref_elem_last_1"][:, :, 15]
    print("%-40s %8.1f kB  nodes=%d elems=%d  yielded points %.0f%%  max eqps %.3e" % (
        "state_cube.npz", os.path.getsize(path) / 1e3, len(mesh["x"]), sum(len(c) for c in mesh["conn"].values()),
        100.0 * (eqps > 0).mean(), eqps.max()))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    refroot = args[0] if args else "/root/reference"
    from oracle import refdrive

    if "--contact-only" not in sys.argv:
        make_state_case()
    if "--state-only" in sys.argv:
        return

    cases = [("dynamics",) + c for c in CASES] if "--contact-only" not in sys.argv else []
    contact_cases = [c for c in CONTACT_CASES if "--entity-creation-only" not in sys.argv or c[0] == "contact_entity_creation"]
    for sub, d, deck, g in ([] if "--entity-creation-only" in sys.argv else cases) + [("contact",) + c for c in contact_cases]:
        base = os.path.join(refroot, "test", sub, d)
        out = {}
        mesh = read_genesis(os.path.join(base, g + ".g"))
        pack_mesh("mesh_", mesh, out)
        out["deck"] = np.array(open(os.path.join(base, deck + ".in")).read())
        # the reference's own per-variable tolerances for this case (exodiff -f <case>.exodiff)
        out["exodiff"] = np.array(open(os.path.join(base, deck + ".exodiff")).read())
        if os.path.exists(os.path.join(base, deck + ".gold.e")):
            times, nod, elem = read_results(os.path.join(base, deck + ".gold.e"))
        else:
            times, nod, elem = np.zeros(0), {}, {}
        out["gold_times"] = times
        if nod and len(next(iter(nod.values()))[0]) != len(mesh["x"]):
            # not a results file of the mesh: the contact visualisation database (src/nimble_contact_manager.cc:431-590)
            f = netcdf_file(os.path.join(base, deck + ".gold.e"), "r", mmap=False)
            for k in ("coordx", "coordy", "coordz", "elem_num_map", "node_num_map", "connect1", "connect2"):
                out["gold_vis_" + k] = np.array(f.variables[k].data)
        for k, a in nod.items():
            out["gold_nod_" + k] = a
        for (k, b), a in elem.items():
            if sub == "contact" and not k.startswith(("stress_xx", "deformation_gradient_xx")):
                continue  # the contact .exodiff files compare nodal variables only; keep the fixture small
            out["gold_elem_%s_eb%d" % (k, b)] = a
        for p in sorted(glob.glob(os.path.join(base, g + ".g.*.*"))):
            tag = p.split(".g.")[1]
            if sub == "contact" and int(tag.split(".")[0]) > 4:
                continue  # (the np8 pieces of sphere_plate_contact: keep the fixture small)
            pack_mesh("piece_%s_" % tag, read_genesis(p), out)
        # reference-code snapshots at the deck's output steps (tight oracle when /root/reference is absent)
        run = refdrive.RefRun(str(out["deck"]), mesh, keep_snapshots=True)
        out["ref_critical_dt"] = np.array(run.begin())
        import re
        nsteps = int(re.search(r"number of load steps:\s*(\d+)", str(out["deck"])).group(1))
        run.advance(nsteps)
        snaps = run.snapshots()
        if sub == "contact" and len(snaps) > 16:  # sliding_contact: 102 output steps; keep every tenth and the last
            keep = sorted(set(range(0, len(snaps), 10)) | {len(snaps) - 1})
            out["ref_snapshot_index"] = np.array(keep)
            snaps = [snaps[i] for i in keep]
        out["ref_times"] = np.array([s["time"] for s in snaps])
        for lbl in snaps[0]["node"]:
            if sub == "contact" and lbl in ("reference_coordinate", "external_force", "acceleration"):
                continue
            out["ref_node_" + lbl] = np.stack([s["node"][lbl] for s in snaps]) if lbl != "lumped_mass" else snaps[0]["node"][lbl][None]
        for b in mesh["block_ids"]:
            # full per-ipt F/sigma only at the final output step (keeps the fixture small)
            out["ref_elem_last_%d" % b] = snaps[-1]["elem"][b]
            for lbl in snaps[0]["derived"][b]:
                if sub == "contact" and lbl not in ("stress_xx", "stress_zz", "volume"):
                    continue
                out["ref_derived_%d_%s" % (b, lbl)] = np.stack([s["derived"][b][lbl] for s in snaps])
        ce = run.contact_entities()
        if ce is not None:
            for k, a in ce.items():
                out["ref_contact_" + k] = a
        run.close()
        path = os.path.join(HERE, deck + ".npz")
        np.savez_compressed(path, **out)
        print("%-40s %8.1f kB  nodes=%d elems=%d times=%d" % (
            deck + ".npz", os.path.getsize(path) / 1e3, len(mesh["x"]),
            sum(len(c) for c in mesh["conn"].values()), len(times)))


if __name__ == "__main__":
    main()
