"""CPU tests of the C++ host layer (nimblesm_b200/host) through its C shim (lib/libnsm_host_c.so): deck parser,
Genesis reader, Exodus writer, expression evaluator, boundary-condition tables and shared-node discovery — the
reference-facing surfaces either side of the hot path.  No device work happens here."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from tests.conftest import ROOT, load_golden

LIB = os.path.join(ROOT, "nimblesm_b200", "lib", "libnsm_host_c.so")
CASES = ["wave_in_bar", "notched_plate_native_neohookean", "notched_plate_native_hypoelastic", "brick_with_fibers",
         "simple_deformation_modes", "rigid_body_motion", "single_elem_complex_displacement", "single_elem_native_neohookean"]


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(LIB):
        import __graft_entry__ as g

        g.build()
    return C.CDLL(LIB)


def _call_json(fn, *args):
    out, err = C.create_string_buffer(1 << 20), C.create_string_buffer(2048)
    rc = fn(*args, out, len(out), err, len(err))
    assert rc == 0, err.value.decode()
    return json.loads(out.value.decode())


EXPRESSIONS = [" 0.0635 * (-0.5*cos(t*3.141592653589793/2.0e-4) + 0.5)", "-0.0635 * (-0.5*cos(t*3.141592653589793/2.0e-4) + 0.5)",
               " 0.01 * t", "-0.01 * t", "0.0005 * (-0.5*cos(t*3.141592653589793/2.0e-6) + 0.5)", "exp(0.2*t)", "exp(-0.2*t)",
               "1000.0*x", "x*y/z*t", "x+y-z+t", "x-y+z", "2*x/3*y", "x^2 + y*z - 3", "-(x+1)^2", "sqrt(x*x+y*y)+abs(-z)",
               "t>0.5?x:y", "x*(y+z)^3/2", "sin(x)*cos(y)+tan(z)-log(t+2)", "1.5e-3*x", "PI*x", "x % 0.3", "floor(10*x)+ceil(y)+round(z)",
               "cbrt(x)+erf(y)-erfc(z)+log10(t+1)+asin(x)+acos(y)+atan(z)", "e^x"]
POINTS = [(0.3, 0.7, 1.1, 1.0e-4), (1.25, -0.5, 2.0, 0.75), (0.0, 0.0, 0.5, 0.0), (0.9, 0.1, 3.3, 2.0e-6)]


def test_expression_tree_matches_the_reference_parser(host, refdrive):
    """Same value, bit for bit, as ExpressionParsing::BoundaryConditionFunctor of the reference (compiled into
    oracle/_ref) on expressions that exercise its association rules (a*b/c = a*(b/c), a+b-c = a+(b-c), ...)."""
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libnimble_ref.so"))
    ref.nsmref_expression_eval.argtypes = [C.c_char_p] + [C.c_double] * 4 + [C.POINTER(C.c_double)]
    host.nsmh_expression_eval.argtypes = [C.c_char_p] + [C.c_double] * 4 + [C.POINTER(C.c_double), C.c_char_p, C.c_int]
    for ex in EXPRESSIONS:
        for p in POINTS:
            a, b = C.c_double(), C.c_double()
            err = C.create_string_buffer(512)
            ra = host.nsmh_expression_eval(ex.encode(), *p, C.byref(a), err, 512)
            rb = ref.nsmref_expression_eval(ex.encode(), *p, C.byref(b))
            assert ra == rb, (ex, err.value)
            if ra == 0:
                assert np.float64(a.value).view(np.int64) == np.float64(b.value).view(np.int64) or (
                    np.isnan(a.value) and np.isnan(b.value)), (ex, p, a.value, b.value)


def test_expression_known_answers(host):
    """Runs everywhere (no reference needed): values of the reference decks' BC expressions at fixed points,
    generated with the reference parser in the build container."""
    host.nsmh_expression_eval.argtypes = [C.c_char_p] + [C.c_double] * 4 + [C.POINTER(C.c_double), C.c_char_p, C.c_int]
    known = json.load(open(os.path.join(ROOT, "tests", "golden", "expressions.json")))
    assert len(known) >= 40
    for ex, p, bits in known:
        a = C.c_double()
        err = C.create_string_buffer(512)
        assert host.nsmh_expression_eval(ex.encode(), *p, C.byref(a), err, 512) == 0, err.value
        assert int(np.float64(a.value).view(np.int64)) == bits, (ex, p)


def test_expression_programs_reproduce_host_evaluation(host):
    """Expression::compile -> device program (nsm_bc_op): interpreted with IEEE doubles it returns the bits of the
    host evaluation, with every sub-expression of t alone supplied as a host-evaluated slot; expressions whose
    position-dependent part needs libm are refused (the host keeps evaluating them)."""
    from tests import bc_program as bp

    compiled = 0
    for ex in EXPRESSIONS + bp.EXTRA_EXPRESSIONS + bp.ENTRY_CONSTANT_EXPRESSIONS:
        for (x, y, z, t) in POINTS:
            prog = bp.compile_expression(host, ex, t)
            if prog is None:
                assert ex not in bp.ENTRY_CONSTANT_EXPRESSIONS, ex
                continue
            compiled += 1
            entry = bp.entry_constants(host, ex, x, y, z)
            got, want = bp.interpret(*prog, x, y, z, entry), bp.host_eval(host, ex, x, y, z, t)
            assert np.float64(got).view(np.int64) == np.float64(want).view(np.int64) or (got != got and want != want), (ex, x, y, z, t)
    assert compiled >= 4 * (len(bp.EXTRA_EXPRESSIONS) + len(bp.ENTRY_CONSTANT_EXPRESSIONS) + 10)
    for ex in bp.NOT_COMPILABLE:
        assert bp.compile_expression(host, ex, 0.5) is None, ex
    # a position-only libm sub-tree is ONE per-entry constant, shared when it repeats
    code, consts, slots = bp.compile_expression(host, "sin(3*x)*cos(t) + sin(3*x)*t", 0.25)
    assert sum(1 for w in code if (int(w) & 0xff) == bp.ENTRYCONST) == 2 and len(bp.entry_constants(host, "sin(3*x)*cos(t) + sin(3*x)*t", 0.1, 0.2, 0.3)) == 1
    # ... and the MAXIMAL position-only sub-tree is taken: x*sin(3*x) is one constant, not sin(3*x) times a device product
    assert len(bp.entry_constants(host, "x*sin(3*x)*t", 0.1, 0.2, 0.3)) == 1
    # a function of t alone is ONE slot and nothing else crosses per step
    code, consts, slots = bp.compile_expression(host, " 0.0635 * (-0.5*cos(t*3.141592653589793/2.0e-4) + 0.5)", 1.0e-4)
    assert len(code) == 1 and (code[0] & 0xff) == bp.SLOT and len(slots) == 1
    # ... and repeated sub-expressions share their slot
    code, consts, slots = bp.compile_expression(host, "x*cos(t) + y*cos(t)", 0.3)
    assert len(slots) == 1


def test_material_params_known_answers_of_the_reference_unit_test(host):
    """unit_tests/test_nimble_material_params.cc:53-147 on the host MaterialParameters / MaterialFactoryBase: an unknown key
    throws, the three doubles of a neohookean string parse, extra double and string parameter names can be registered."""

    def parse(text, doubles=b"", strings=b""):
        out, err = C.create_string_buffer(4096), C.create_string_buffer(2048)
        rc = host.nsmh_material_params(text.encode(), doubles, strings, out, len(out), err, len(err))
        return (json.loads(out.value.decode()) if rc == 0 else None), err.value.decode()

    # invalid_parameter_throws
    got, err = parse("neohookean some_stuff 1.0e6 stuff 0.27 density 1.0e3")
    assert got is None and "Invalid material parameter encountered: 'some_stuff'" in err
    # parse_double_parameters
    got, err = parse("neohookean bulk_modulus 1.0e6 shear_modulus 5.e5 density 1.0e3")
    assert got["name"] == "neohookean" and got["upper"] == "NEOHOOKEAN" and got["n_doubles"] == 3 and got["n_strings"] == 0
    assert got["doubles"] == {"bulk_modulus": 1.0e6, "shear_modulus": 5.0e5, "density": 1.0e3} and "testParam" not in got["doubles"]
    # register_new_test_property / register_new_test_string_property
    text = "neohookean bulk_modulus 1.0e6 shear_modulus 5.e5 density 1.0e3 test_property "
    assert parse(text + "2.0")[0] is None  # not registered: refused
    got, err = parse(text + "2.0", doubles=b"test_property")
    assert got["doubles"]["test_property"] == 2.0 and got["n_doubles"] == 4
    got, err = parse(text + "custom_property_val", strings=b"test_property")
    assert got["strings"] == {"test_property": "custom_property_val"} and got["n_strings"] == 1 and got["n_doubles"] == 3
    # (src/nimble_material_factory_base.cc:85: std::map::insert keeps the first value of a repeated key)
    assert parse("elastic density 1 density 2 bulk_modulus 3 shear_modulus 4")[0]["doubles"]["density"] == 1.0
    assert parse("elastic density")[0] is None and parse("elastic")[0] is None


@pytest.mark.parametrize("case", CASES)
def test_parser_and_material_factory_on_reference_decks(host, case):
    from nimblesm_b200.deck import parse_deck

    deck, mesh, _gold, _ref, _pieces = load_golden(case)
    ids = (C.c_int * len(mesh["all_block_ids"]))(*mesh["all_block_ids"])
    s = _call_json(host.nsmh_deck_summary, deck.encode(), ids, len(mesh["all_block_ids"]))
    d = parse_deck(deck)
    assert s["scheme"] == "explicit" and s["num_load_steps"] == d.num_load_steps and s["output_frequency"] == d.output_frequency
    assert s["final_time"] == d.final_time and s["initial_time"] == d.initial_time
    assert s["genesis"] == d.genesis_file and s["exodus"] == d.exodus_file
    assert s["n_bc"] == len(d.boundary_conditions)
    for b in mesh["block_ids"]:
        m, want = s["materials"][str(b)], d.block_material(b)
        assert (m["model"], m["density"], m["bulk_modulus"], m["shear_modulus"], m["num_state"]) == (
            want.model, want.density, want.bulk_modulus, want.shear_modulus, 0)


def test_parser_errors_like_the_reference(host):
    ids = (C.c_int * 1)(1)
    out, err = C.create_string_buffer(4096), C.create_string_buffer(2048)
    base = "genesis input file: a.g\nexodus output file: a.e\nfinal time: 1.0\nnumber of load steps: 1\noutput fields: displacement\n"
    rc = host.nsmh_deck_summary((base + "no such key: 1\n").encode(), ids, 1, out, 4096, err, 2048)
    assert rc != 0 and b"unknown key no such key" in err.value
    rc = host.nsmh_deck_summary((base + "material parameters: m neohookean density 1 youngs 3\nelement block: block_1 m\n").encode(),
                                ids, 1, out, 4096, err, 2048)
    assert rc != 0 and b"Invalid material parameter encountered: 'youngs'" in err.value
    rc = host.nsmh_deck_summary((base + "material parameters: m plastic density 1 bulk_modulus 2 shear_modulus 3\nelement block: block_1 m\n").encode(),
                                ids, 1, out, 4096, err, 2048)
    assert rc != 0 and b"invalid material model name" in err.value
    rc = host.nsmh_deck_summary("genesis input file: a.g\n".encode(), ids, 1, out, 4096, err, 2048)
    assert rc != 0 and b"output fields not found" in err.value


def test_io_file_name(host):
    out = C.create_string_buffer(512)
    for args, want in ((("wave.g", "g", "", 0, 1), "wave.g"), (("wave.g", "g", "", 3, 16), "wave.g.16.03"),
                       (("wave.e", "e", "out", 0, 1), "wave.out.e"), (("wave.e", "e", "out", 1, 2), "wave.out.e.2.1"),
                       (("none", "e", "out", 1, 2), "none")):
        assert host.nsmh_io_file_name(args[0].encode(), args[1].encode(), args[2].encode(), args[3], args[4], out, 512) == 0
        assert out.value.decode() == want


def _mesh_checks(s, mesh):
    assert s["dim"] == 3 and s["num_nodes"] == len(mesh["x"])
    assert s["num_elements"] == sum(len(mesh["conn"][b]) for b in mesh["block_ids"])
    assert s["sum_x"] == float(np.sum(np.cumsum(mesh["x"])[-1:])) or np.isclose(s["sum_x"], mesh["x"].sum(), rtol=1e-13)
    assert s["node_gid_sum"] == int(np.asarray(mesh["node_gid"], dtype=np.int64).sum())
    for b in mesh["block_ids"]:
        c = np.asarray(mesh["conn"][b], dtype=np.int64).ravel()
        w = (np.arange(len(c)) % 7 + 1)
        assert s["blocks"][str(b)]["num_elements"] == len(mesh["conn"][b])
        assert s["blocks"][str(b)]["conn_checksum"] == int((w * c).sum())
        assert s["blocks"][str(b)]["name"] == "block_%d" % b
    for sid, nodes in mesh["node_sets"].items():
        e = s["node_sets"][str(sid)]
        assert e["size"] == len(nodes) and e["sum"] == int(np.asarray(nodes, dtype=np.int64).sum())
        assert e["name"] == "nodelist_%d" % sid


@pytest.mark.parametrize("case", CASES)
def test_genesis_reader_on_fixture_meshes_and_pieces(host, case, tmp_path):
    """The mesh of every fixture (and every decomposed piece) is written as a NetCDF-3 Genesis file and read back
    by the C++ reader: counts, id maps, connectivity checksums, node-set contents, default names."""
    from nimblesm_b200.exodus_py import write_genesis

    _deck, mesh, _gold, _ref, pieces = load_golden(case)
    p = str(tmp_path / "m.g")
    write_genesis(p, mesh)
    _mesh_checks(_call_json(host.nsmh_mesh_summary, p.encode()), mesh)
    for (P, r), piece in list(pieces.items())[:6]:
        pp = str(tmp_path / ("m.g.%d.%d" % (P, r)))
        write_genesis(pp, piece)
        _mesh_checks(_call_json(host.nsmh_mesh_summary, pp.encode()), piece)


def test_genesis_reader_on_the_reference_files(host):
    """The cubit / SEACAS-written files of the reference themselves (only where /root/reference exists)."""
    base = "/root/reference/test/dynamics"
    if not os.path.isdir(base):
        pytest.skip("reference tree not present")
    from tests.golden.make_golden import read_genesis

    n = 0
    for d in sorted(os.listdir(base)):
        if not os.path.isdir(os.path.join(base, d)):
            continue
        for fn in sorted(os.listdir(os.path.join(base, d))):
            if fn.endswith(".g") or ".g." in fn:
                path = os.path.join(base, d, fn)
                _mesh_checks(_call_json(host.nsmh_mesh_summary, path.encode()), read_genesis(path))
                n += 1
    assert n >= 20


@pytest.mark.parametrize("case,P", [("brick_with_fibers", 2), ("notched_plate_native_neohookean", 4), ("wave_in_bar", 3),
                                    ("simple_deformation_modes", 8)])
def test_driver_side_rcb_decomposition(host, tmp_path, case, P):
    """GenesisMesh::RcbElementPartition + KeepPart (what `NimbleSM_b200 --gpus P` does to a serial mesh when no
    Nemesis pieces exist) against the Python mirror nimblesm_b200.mesh.rcb_partition: same parts, same local
    numbering, node sets restricted alike; every element lands in exactly one part."""
    from nimblesm_b200.exodus_py import write_genesis
    from nimblesm_b200.mesh import rcb_partition

    _deck, mesh, _gold, _ref, _pieces = load_golden(case)
    p = str(tmp_path / "m.g")
    write_genesis(p, mesh)
    want = rcb_partition(mesh, P)
    seen = {b: [] for b in mesh["block_ids"]}
    for r in range(P):
        got = _call_json(host.nsmh_mesh_part, p.encode(), P, r)
        w = want[r]
        assert got["block_ids"] == list(w["block_ids"]) and got["all_block_ids"] == list(mesh["all_block_ids"])
        assert got["node_gid"] == [int(g) for g in w["node_gid"]]
        assert np.array_equal(np.array(got["x"]), w["x"])
        for b in w["block_ids"]:
            assert got["elem_gid"][str(b)] == [int(g) for g in w["elem_gid"][b]]
            assert got["conn"][str(b)] == [int(n) for n in w["conn"][b].ravel()]
            seen[b] += got["elem_gid"][str(b)]
        for sid, ns in w["node_sets"].items():
            assert got["node_sets"][str(sid)] == [int(n) for n in ns]
    for b in mesh["block_ids"]:
        assert sorted(seen[b]) == sorted(int(g) for g in mesh["elem_gid"][b])


def test_exodus_writer_round_trip(host, tmp_path):
    from nimblesm_b200.exodus_py import read_results, write_genesis

    _deck, mesh, _gold, _ref, _pieces = load_golden("brick_with_fibers")
    g, e = str(tmp_path / "b.g"), str(tmp_path / "b.out.e")
    write_genesis(g, mesh)
    err = C.create_string_buffer(2048)
    assert host.nsmh_exodus_roundtrip(g.encode(), e.encode(), 3, err, 2048) == 0, err.value
    r = read_results(e)
    assert np.array_equal(r["times"], [0.0, 0.25, 0.5])
    for k, c in enumerate("xyz"):
        want = np.stack([(s + 1) * np.asarray(mesh[c]) for s in range(3)])
        assert np.array_equal(r["nod"]["displacement_" + c], want)
    assert r["block_ids"] == mesh["all_block_ids"]
    for bi, b in enumerate(mesh["all_block_ids"]):
        ne = len(mesh["conn"][b])
        assert np.array_equal(r["elem"][("volume", bi)], np.stack([np.arange(ne) + s for s in range(3)]))
        assert np.array_equal(r["elem"][("ipt01_stress_xx", bi)], np.stack([10.0 * np.arange(ne) + s for s in range(3)]))
    assert np.array_equal(r["node_gid"], mesh["node_gid"])
    # the file is a well-formed Exodus database for scipy's independent NetCDF reader: mesh payload intact
    from scipy.io import netcdf_file

    f = netcdf_file(e, "r", mmap=False)
    assert f.variables["coordx"].data.shape == (len(mesh["x"]),) and np.array_equal(f.variables["coordx"].data, mesh["x"])
    for bi, b in enumerate(mesh["all_block_ids"]):
        assert np.array_equal(f.variables["connect%d" % (bi + 1)].data - 1, mesh["conn"][b])
    names = [b"".join(row).split(b"\x00")[0].decode() for row in f.variables["name_elem_var"].data]
    assert names == sorted(names) == ["ipt01_stress_xx", "volume"]
    f.close()


def test_boundary_condition_tables(host, tmp_path):
    from nimblesm_b200.exodus_py import write_genesis

    deck, mesh, _gold, _ref, _pieces = load_golden("notched_plate_native_neohookean")
    g = str(tmp_path / "n.g")
    write_genesis(g, mesh)
    cap = 100000
    n, td = C.c_int(), C.c_int()
    node, comp, kind = (C.c_int * cap)(), (C.c_int * cap)(), (C.c_int * cap)()
    val = (C.c_double * cap)()
    err = C.create_string_buffer(2048)
    t = 7.0e-7
    assert host.nsmh_bc_table(g.encode(), deck.encode(), C.c_double(t), cap, C.byref(n), node, comp, kind, val, C.byref(td), err, 2048) == 0, err.value
    from nimblesm_b200.deck import parse_deck

    d = parse_deck(deck)
    want_nodes, want_comp, want_kind = [], [], []
    for bc in d.boundary_conditions:
        if bc.kind == "initial_velocity":
            continue
        ns = mesh["node_sets"][bc.node_set_id]
        want_nodes += list(ns)
        want_comp += [bc.coordinate] * len(ns)
        want_kind += [0 if bc.kind == "prescribed_velocity" else 1] * len(ns)
    assert n.value == len(want_nodes) and list(node[:n.value]) == want_nodes
    assert list(comp[:n.value]) == want_comp and list(kind[:n.value]) == want_kind
    assert td.value == 0  # constants only
    assert np.array_equal(np.array(val[:n.value]), np.zeros(n.value))
    # a prescribed displacement that depends on t and x is flagged time dependent and evaluated per entry
    deck2 = deck + '\nboundary condition: prescribed_displacement nodelist_1 y "0.5*x*cos(t*3.0/2.0e-6)"\n'
    assert host.nsmh_bc_table(g.encode(), deck2.encode(), C.c_double(t), cap, C.byref(n), node, comp, kind, val, C.byref(td), err, 2048) == 0, err.value
    ns1 = mesh["node_sets"][1]
    assert td.value == 1 and n.value == len(want_nodes) + len(ns1)
    tail_nodes = np.array(node[len(want_nodes):n.value])
    assert np.array_equal(tail_nodes, ns1) and set(kind[len(want_nodes):n.value]) == {1}
    want = np.array([0.5 * x * np.cos(t * (3.0 / 2.0e-6)) for x in mesh["x"][ns1]])
    assert np.allclose(np.array(val[len(want_nodes):n.value]), want, rtol=1e-15, atol=0)


def test_bc_manager_chooses_device_programs(host, tmp_path, monkeypatch):
    """BoundaryConditionManager::GetDevicePrograms: every time-dependent expression compiles -> the step loop evaluates
    the magnitudes on the device (programs per BC, shared slots for the sub-expressions of t); one expression that needs
    libm at a position (sin(x*t)) or NSM_B200_HOST_BC=1 -> the host evaluates per node, as the reference does."""
    import math

    from nimblesm_b200.exodus_py import write_genesis

    deck, mesh, _gold, _ref, _pieces = load_golden("wave_in_bar")
    g = str(tmp_path / "w.g")
    write_genesis(g, mesh)
    n2, n1 = len(mesh["node_sets"][2]), len(mesh["node_sets"][1])
    t = 3.0e-7
    s = _call_json(host.nsmh_bc_programs, g.encode(), deck.encode(), C.c_double(t))
    assert s["active"] is False and s["time_dependent"] is False and s["n_entries"] == 3 * n2  # constants only
    td = deck + '\nboundary condition: prescribed_velocity nodelist_2 y "0.5*(1.0-cos(t*2.0e6))*(1.0+z)"\n' \
              + 'boundary condition: prescribed_displacement nodelist_1 z "x*cos(t*2.0e6) + 1.0e-3*t"\n'
    s = _call_json(host.nsmh_bc_programs, g.encode(), td.encode(), C.c_double(t))
    assert s["active"] is True and s["n_programs"] == 2 and s["entries_with_program"] == n2 + n1
    assert s["n_entries"] == 4 * n2 + n1
    # slots: 0.5*(1.0-cos(t*2.0e6)), cos(t*2.0e6), 1.0e-3*t -- the cosine itself is NOT shared between the first two
    # (the first slot is the whole t-only factor), the values are glibc's
    assert s["n_slots"] == 3
    assert s["slots_at_t"] == [0.5 * (1.0 - math.cos(t * 2.0e6)), math.cos(t * 2.0e6), 1.0e-3 * t]
    bad = td + 'boundary condition: prescribed_velocity nodelist_2 x "sin(x*t)"\n'
    s = _call_json(host.nsmh_bc_programs, g.encode(), bad.encode(), C.c_double(t))
    assert s["active"] is False and s["time_dependent"] is True and s["n_programs"] == 0
    monkeypatch.setenv("NSM_B200_HOST_BC", "1")
    s = _call_json(host.nsmh_bc_programs, g.encode(), td.encode(), C.c_double(t))
    assert s["active"] is False and s["time_dependent"] is True


@pytest.mark.parametrize("case,P", [("wave_in_bar", 2), ("wave_in_bar", 4), ("brick_with_fibers", 4)])
def test_vector_communicator_tables_match_python(host, case, P):
    """VectorCommunicator::Initialize on P rank threads == mesh.shared_node_tables (used by bench.py): for every
    peer the shared nodes in ascending GLOBAL id order."""
    from nimblesm_b200.mesh import shared_node_tables

    _deck, _mesh, _gold, _ref, pieces = load_golden(case)
    parts = [pieces[(P, r)] for r in range(P)]
    gids = np.concatenate([np.asarray(p["node_gid"], dtype=np.int32) for p in parts])
    off = np.concatenate([[0], np.cumsum([len(p["node_gid"]) for p in parts])]).astype(np.int32)
    cap = 1 << 16
    for r in range(P):
        npeers = C.c_int()
        peers, poff, pnodes = (C.c_int * 64)(), (C.c_longlong * 65)(), (C.c_int * cap)()
        err = C.create_string_buffer(2048)
        rc = host.nsmh_shared_node_tables(P, gids.ctypes.data_as(C.POINTER(C.c_int)), off.ctypes.data_as(C.POINTER(C.c_int)), r, cap,
                                          C.byref(npeers), peers, poff, pnodes, err, 2048)
        assert rc == 0, err.value
        wp, wo, wn = shared_node_tables(r, [np.asarray(p["node_gid"]) for p in parts], np.asarray(parts[r]["node_gid"]))
        assert list(peers[:npeers.value]) == list(wp)
        assert list(poff[:npeers.value + 1]) == list(wo)
        assert list(pnodes[:poff[npeers.value]]) == list(wn)


def host_contact_entities(host, genesis_path, deck):
    """ContactManager::BuildEntityLists through the C shim -> dict (mesh node ids)."""
    n = (C.c_longlong * 2)()
    pen = C.c_double()
    err = C.create_string_buffer(2048)
    assert host.nsmh_contact_entities(genesis_path.encode(), deck.encode(), n, C.byref(pen), None, None, None, None, None, err, 2048) == 0, err.value
    nf, nn = int(n[0]), int(n[1])
    quads, ids, flen = np.zeros((nf, 4), np.int32), np.zeros(nf, np.int32), np.zeros(nf)
    nodes, nlen = np.zeros(nn, np.int32), np.zeros(nn)
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    assert host.nsmh_contact_entities(genesis_path.encode(), deck.encode(), n, C.byref(pen), quads.ctypes.data_as(ip), ids.ctypes.data_as(ip),
                                      flen.ctypes.data_as(dp), nodes.ctypes.data_as(ip), nlen.ctypes.data_as(dp), err, 2048) == 0, err.value
    return {"penalty": pen.value, "primary_quads": quads, "primary_entity_ids": ids, "primary_char_len": flen, "contact_nodes": nodes,
            "contact_node_char_len": nlen}


@pytest.mark.parametrize("case", ["cubes_contact", "sphere_plate_contact", "sliding_contact", "contact_entity_creation"])
def test_contact_manager_entity_lists(host, tmp_path, case):
    """ContactManager::SkinBlocks / CreateContactEntities (host side, C++): the skin quads of the primary blocks in the
    reference's std::map order, the contact nodes of the secondary blocks in order of first appearance, and both
    characteristic lengths -- equal, value for value, to what the glue around the reference's own ContactEntity objects
    built from the same deck (tests/golden, oracle/ref_contact.cc)."""
    from nimblesm_b200.exodus_py import write_genesis

    deck, mesh, _gold, ref, _ = load_golden(case)
    g = str(tmp_path / "c.g")
    write_genesis(g, mesh)
    got = host_contact_entities(host, g, deck)
    assert got["penalty"] == (1.0e13 if case == "contact_entity_creation" else float("0.33333333333333333333333333333e12"))
    for k in ("primary_quads", "contact_nodes", "primary_char_len", "contact_node_char_len"):
        assert np.array_equal(got[k], ref["contact_" + k]), k
    # entity ids: (element global id + 1 + largest node global id) << 5 | face ordinal << 2, unique per face
    assert len(set(got["primary_entity_ids"].tolist())) == len(got["primary_quads"])
    assert np.all((got["primary_entity_ids"] & 3) == 0) and np.all(((got["primary_entity_ids"] >> 2) & 7) < 6)


def test_contact_manager_entities_equal_the_reference_visualisation_database(host, tmp_path):
    """The reference's own record of CreateContactEntities: the gold file of test/contact/contact_entity_creation is the
    contact visualisation database (src/nimble_contact_manager.cc:431-590) -- every facet with its three vertices and
    every contact node, in entity order, with the entity ids as Exodus maps.  The C++ ContactManager of this repository
    (five blocks: primary block_1 block_2, secondary block_3 block_5) makes the same entities: 6 778 vertex coordinates
    bit for bit in the same order, the same entity id per triangle ((element id + offset) << 5 | face ordinal << 2 |
    triangle ordinal, :924-927, :1106-1176), the same ids for the visualisation nodes (:525-535) and contact nodes (:321)."""
    from nimblesm_b200.exodus_py import write_genesis

    deck, mesh, gold, _ref, _ = load_golden("contact_entity_creation")
    vis = gold["vis"]
    g = str(tmp_path / "c.g")
    write_genesis(g, mesh)
    ent = host_contact_entities(host, g, deck)
    n_tri = 4 * len(ent["primary_quads"])
    assert (n_tri, len(ent["contact_nodes"])) == (len(vis["connect1"]), len(vis["connect2"])) == (2112, 442)
    X = np.stack([mesh["x"], mesh["y"], mesh["z"]], 1)
    c = X[ent["primary_quads"]]
    centre = (((0.0 + c[:, 0]) + c[:, 1]) + c[:, 2] + c[:, 3]) / 4
    tri = np.stack([np.stack([c[:, k], c[:, (k + 1) % 4], centre], 1) for k in range(4)], 1)  # [nf,4,3,3]
    got = np.vstack([tri.reshape(-1, 3), X[ent["contact_nodes"]]])
    want = np.stack([vis["coordx"], vis["coordy"], vis["coordz"]], 1)
    assert np.array_equal(got.view(np.int64), want.astype(np.float64).view(np.int64))
    tri_ids = np.repeat(ent["primary_entity_ids"], 4) | np.tile(np.arange(4, dtype=np.int32), len(ent["primary_quads"]))
    node_ids = np.asarray(mesh["node_gid"])[ent["contact_nodes"]] + 1
    assert np.array_equal(vis["elem_num_map"], np.concatenate([tri_ids, node_ids]) + 1)  # (Exodus maps are written 1-based)
    top = max(tri_ids.max(), node_ids.max())
    vis_nodes = np.stack([3 * tri_ids + top + 9, 3 * tri_ids + top + 10, 3 * tri_ids + top + 11], 1).ravel()
    assert np.array_equal(vis["node_num_map"], np.concatenate([vis_nodes, node_ids]) + 1)


def test_contact_visualisation_database_meets_the_reference_gold_file(host, tmp_path):
    """The reference's own contract for test/contact/contact_entity_creation (run_exodiff_test.py: exodiff -f
    contact_entity_creation.exodiff contact_entity_creation.gold.e contact_entity_creation.out.e): the file named by the
    deck's `contact visualization` line.  The host ContactVisualizationDatabase, fed with the oracle's displacement at
    the reference's output steps, writes that file: same Exodus mesh (coordinates bit for bit, connectivity, both id
    maps, block layout), time planes and nodal displacement within the reference's exodiff rules."""
    from scipy.io import netcdf_file

    from nimblesm_b200 import exodiff
    from nimblesm_b200.exodus_py import read_results, write_genesis
    from oracle.model import OracleModel

    deck, mesh, gold, ref, _ = load_golden("contact_entity_creation")
    g = str(tmp_path / "c.g")
    write_genesis(g, mesh)
    m = OracleModel(deck, mesh)
    m.begin()
    planes, at = [], 0
    for step in (0, 1, 21, 41, 60):
        m.advance(step - at)
        at = step
        planes.append(m.u.copy())
    disp = np.ascontiguousarray(np.stack(planes))
    times = np.ascontiguousarray(ref["times"], dtype=np.float64)
    evaluated = np.array([0, 1, 1, 1, 1], np.int32)
    out, err = str(tmp_path / "contact_entity_creation.out.e"), C.create_string_buffer(2048)
    rc = host.nsmh_contact_visualization(g.encode(), deck.encode(), out.encode(), 5, times.ctypes.data_as(C.POINTER(C.c_double)),
                                         disp.ctypes.data_as(C.POINTER(C.c_double)), evaluated.ctypes.data_as(C.POINTER(C.c_int)), err, 2048)
    assert rc == 0, err.value
    f = netcdf_file(out, "r", mmap=False)
    vis = gold["vis"]
    for k in ("coordx", "coordy", "coordz"):
        assert np.array_equal(np.array(f.variables[k].data, dtype=np.float64).view(np.int64), vis[k].astype(np.float64).view(np.int64)), k
    for k in ("connect1", "connect2", "elem_num_map", "node_num_map"):
        assert np.array_equal(np.array(f.variables[k].data), vis[k]), k
    assert f.dimensions["num_el_blk"] == 2 and f.dimensions["num_nod_per_el1"] == 3 and f.dimensions["num_nod_per_el2"] == 1
    names = [b"".join(r).split(b"\x00")[0].decode().strip() for r in f.variables["name_nod_var"].data]
    assert names == ["displacement_x", "displacement_y", "displacement_z", "contact_status"]
    f.close()
    res = read_results(out)
    fails = exodiff.compare(gold["exodiff"], gold, res)
    assert not fails, fails[:5]
    assert np.abs(res["nod"]["displacement_x"][-1]).max() == pytest.approx(0.02, rel=1e-12) and not res["nod"]["contact_status"].any()
    # a deck without the line is refused; a malformed line keeps the reference's parser error
    rc = host.nsmh_contact_visualization(g.encode(), "\n".join(l for l in deck.splitlines() if not l.startswith("contact visualization")).encode(),
                                         out.encode(), 0, None, None, None, err, 2048)
    assert rc != 0 and b"no contact visualization" in err.value
    rc = host.nsmh_contact_visualization(g.encode(), deck.replace("visualize_bounding_boxes off", "visualize_boxes off").encode(), out.encode(),
                                         0, None, None, None, err, 2048)
    assert rc != 0 and b'unexpected value for "contact visualization"' in err.value


def test_contact_command_errors(host, tmp_path):
    """ParseContactCommand keeps the reference's error behaviour (src/nimble_contact_manager.cc:95-149)."""
    from nimblesm_b200.exodus_py import write_genesis

    deck, mesh, *_ = load_golden("cubes_contact")
    g = str(tmp_path / "c.g")
    write_genesis(g, mesh)
    n, pen, err = (C.c_longlong * 2)(), C.c_double(), C.create_string_buffer(2048)
    base = "\n".join(l for l in deck.splitlines() if not l.startswith("contact:"))
    for line, msg in (("contact: main_blocks block_2 secondary_blocks block_1 penalty_parameter 1.0", b"unknown key: main_blocks"),
                      ("contact: primary_blocks block_2 penalty_parameter 1.0", b"secondary_blocks"),
                      ("contact: primary_blocks block_2 secondary_blocks block_1", b"penalty_parameter")):
        rc = host.nsmh_contact_entities(g.encode(), (base + "\n" + line + "\n").encode(), n, C.byref(pen), None, None, None, None, None, err, 2048)
        assert rc == 1 and msg in err.value, err.value
    # the deprecated spelling is still read
    ok = base + "\ncontact: master_blocks block_2 slave_blocks block_1 penalty_parameter 2.5e11\n"
    assert host.nsmh_contact_entities(g.encode(), ok.encode(), n, C.byref(pen), None, None, None, None, None, err, 2048) == 0, err.value
    assert pen.value == 2.5e11 and n[0] == 384 and n[1] == 518


@pytest.mark.parametrize("seed", range(5))
def test_skin_blocks_on_random_block_assignments(host, tmp_path, seed):
    """ContactManager::SkinBlocks (sort-based) against the oracle's independent skinner on cubes whose elements are dealt
    at random to a primary block, a secondary block and a bystander block -- ragged skins, interfaces between blocks that
    share nodes, faces inside a block: same quads in the same order, same contact nodes, same characteristic lengths."""
    from nimblesm_b200.exodus_py import write_genesis
    from nimblesm_b200.mesh import structured_cube
    from oracle.contact import ContactSetup

    rng = np.random.default_rng(40 + seed)
    n = int(rng.integers(3, 7))
    lab = rng.integers(1, 4, size=(n, n, n))
    mesh = structured_cube(n, block_of_element=lambda i, j, k: lab[i, j, k])
    if len(mesh["block_ids"]) < 3:
        pytest.skip("a block came out empty")
    # distort the lattice so that edge lengths differ
    for c in "xyz":
        mesh[c] = mesh[c] + 0.2 / n * (rng.random(len(mesh[c])) - 0.5)
    mesh["node_gid"] = np.arange(len(mesh["x"]), dtype=np.int32)
    g = str(tmp_path / "r.g")
    write_genesis(g, mesh)
    deck = ("genesis input file: r.g\nexodus output file: r.e\nfinal time: 1.0\nnumber of load steps: 1\n"
            "material parameters: m neohookean density 7.8 bulk_modulus 1.6e12 shear_modulus 0.8e12\n"
            "element block: block_1 m\nelement block: block_2 m\nelement block: block_3 m\n"
            "contact: primary_blocks block_2 secondary_blocks block_1 penalty_parameter 1.0e9\n")
    got = host_contact_entities(host, g, deck)
    want = ContactSetup(mesh, [2], [1], 1.0e9)
    assert np.array_equal(got["primary_quads"], want.primary_quads)
    assert np.array_equal(got["contact_nodes"], want.contact_nodes)
    assert np.array_equal(got["primary_char_len"], want.primary_char_len)
    assert np.array_equal(got["contact_node_char_len"], want.contact_node_char_len)
    assert len(got["primary_quads"]) > 0 and len(got["contact_nodes"]) > 0


@pytest.mark.parametrize("case,P", [("sphere_plate_contact", 2), ("sphere_plate_contact", 4), ("sliding_contact", 4),
                                    ("contact_entity_creation", 2), ("contact_entity_creation", 3), ("contact_entity_creation", 4)])
def test_replicated_contact_sub_model_on_rank_threads(host, tmp_path, case, P):
    """Contact across mesh partitions, host side (ContactManager::BuildReplicatedSubModel on P rank threads, one Nemesis
    piece each, no device): every rank ends up with the SAME sub-model, and that sub-model is the serial one -- the same
    skin quads (partition cuts dropped), the same contact nodes, the same characteristic lengths as the glue around the
    reference's own ContactEntity objects built from the undecomposed mesh (tests/golden), all in global node ids."""
    from nimblesm_b200.exodus_py import write_genesis

    deck, mesh, _gold, ref, pieces = load_golden(case)
    paths = []
    for r in range(P):
        path = str(tmp_path / ("p.g.%d.%d" % (P, r)))
        write_genesis(path, pieces[(P, r)])
        paths.append(path)
    j = _call_json(host.nsmh_contact_replicated, "\n".join(paths).encode(), deck.encode(), P)
    assert j["identical"] is True
    gid = np.asarray(mesh["node_gid"])
    quads_want = gid[ref["contact_primary_quads"]]
    quads_got = np.array(j["primary_quads_gid"]).reshape(-1, 4)
    # same set of faces, each with the same cyclic orientation (the owning element may list it from another corner)
    key = lambda q: tuple(sorted(q))
    want = {key(q): list(q) for q in quads_want.tolist()}
    assert len(quads_got) == len(quads_want) and {key(q) for q in quads_got.tolist()} == set(want)
    for q in quads_got.tolist():
        w = want[key(q)]
        i = w.index(q[0])
        assert w[i:] + w[:i] == q
    len_want = {key(q): l for q, l in zip(quads_want.tolist(), ref["contact_primary_char_len"])}
    assert all(len_want[key(q)] == l for q, l in zip(quads_got.tolist(), j["primary_char_len"]))
    nodes_want = dict(zip(gid[ref["contact_contact_nodes"]].tolist(), ref["contact_contact_node_char_len"]))
    assert dict(zip(j["contact_nodes_gid"], j["contact_node_char_len"])) == nodes_want
    assert j["n_surface"] == len(set(quads_want.ravel().tolist()) | set(gid[np.unique(ref["contact_secondary_quads"])].tolist()))
    assert sum(j["held"]) >= j["n_surface"] and all(h > 0 for h in j["held"])
