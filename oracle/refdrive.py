"""oracle/refdrive.py — TEST INFRASTRUCTURE (never imported by nimblesm_b200/).

ctypes driver for oracle/_ref/libnimble_ref.so: the reference's own serial CPU path compiled from
/root/reference/src by oracle/Makefile (glue: oracle/ref_glue.cc).  Used by tests/ as the tight
(1e-12 / 1e-9) oracle, by tests/golden/make_golden.py to generate fixtures, and by bench.py's
cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libnimble_ref.so")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def available() -> bool:
    return os.path.exists(REF_LIB)


# tests/ref_binding/_build/libref_binding.so: the SAME glue compiled with the B200 binding class as the model data
BINDING_LIB = os.path.join(os.path.dirname(_HERE), "tests", "ref_binding", "_build", "libref_binding.so")
_libs = {}


def binding_available() -> bool:
    return os.path.exists(BINDING_LIB)


def lib(path=None):
    path = path or REF_LIB
    if path not in _libs:
        L = C.CDLL(path)
        L.nsmref_open.restype = C.c_void_p
        L.nsmref_open.argtypes = [C.c_char_p, C.c_int, _ip, _dp, _dp, _dp, C.c_int, _ip, _ip, _ip, _ip,
                                  C.c_int, _ip, _ip, _ip, C.c_int]
        L.nsmref_close.argtypes = [C.c_void_p]
        L.nsmref_last_error.restype = C.c_char_p
        L.nsmref_last_error.argtypes = [C.c_void_p]
        L.nsmref_begin.restype = C.c_double
        L.nsmref_begin.argtypes = [C.c_void_p]
        L.nsmref_advance.restype = C.c_double
        L.nsmref_advance.argtypes = [C.c_void_p, C.c_int]
        L.nsmref_internal_force.argtypes = [C.c_void_p]
        L.nsmref_num_nodes.restype = C.c_int
        L.nsmref_num_nodes.argtypes = [C.c_void_p]
        L.nsmref_node_field.restype = C.POINTER(C.c_double)
        L.nsmref_node_field.argtypes = [C.c_void_p, C.c_char_p]
        L.nsmref_elem_data.restype = C.c_long
        L.nsmref_elem_data.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.nsmref_num_snapshots.restype = C.c_int
        L.nsmref_num_snapshots.argtypes = [C.c_void_p]
        L.nsmref_snapshot_time.restype = C.c_double
        L.nsmref_snapshot_time.argtypes = [C.c_void_p, C.c_int]
        L.nsmref_snapshot_node.restype = C.c_long
        L.nsmref_snapshot_node.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p]
        L.nsmref_snapshot_elem.restype = C.c_long
        L.nsmref_snapshot_elem.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.nsmref_derived_labels.restype = C.c_long
        L.nsmref_derived_labels.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_long]
        L.nsmref_snapshot_derived.restype = C.c_long
        L.nsmref_snapshot_derived.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.nsmref_elem_stride.restype = C.c_int
        L.nsmref_elem_stride.argtypes = [C.c_void_p, C.c_int]
        L.nsmref_material_stress.restype = C.c_int
        L.nsmref_material_stress.argtypes = [C.c_char_p, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]
        L.nsmref_bench_steps.restype = C.c_double
        L.nsmref_bench_steps.argtypes = [C.c_char_p, C.c_int, _dp, C.c_int, _ip, _dp, _dp, _dp, _dp, _dp,
                                         C.c_double, C.c_int, C.c_int]
        L.nsmref_enable_output.restype = C.c_int
        L.nsmref_enable_output.argtypes = [C.c_void_p, C.c_char_p]
        L.nsmref_device_launches.restype = C.c_long
        L.nsmref_device_launches.argtypes = [C.c_void_p]
        L.nsmref_contact_ints.restype = C.c_long
        L.nsmref_contact_ints.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.nsmref_contact_doubles.restype = C.c_long
        L.nsmref_contact_doubles.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.nsmref_contact_force.restype = C.c_long
        L.nsmref_contact_force.argtypes = [C.c_void_p, _dp, _dp]
        L.nsmref_contact_pairs_last.restype = C.c_long
        L.nsmref_contact_pairs_last.argtypes = [C.c_void_p]
        _libs[path] = L
    return _libs[path]


class RefRun:
    """One reference run: deck text + mesh dict (see meshio.py for the mesh dict layout)."""

    def __init__(self, deck_text: str, mesh: dict, keep_snapshots: bool = True, lib_path: str = None):
        """lib_path: None = the serial reference (libnimble_ref.so); BINDING_LIB = the same glue over the B200 binding."""
        self._lib_path = lib_path
        L = lib(lib_path)
        self._tmp = tempfile.NamedTemporaryFile("w", suffix=".in", delete=False)
        self._tmp.write(deck_text)
        self._tmp.close()
        self.mesh = mesh
        bids = np.ascontiguousarray(mesh["block_ids"], dtype=np.int32)
        nel = np.ascontiguousarray([len(mesh["conn"][b]) for b in mesh["block_ids"]], dtype=np.int32)
        conn = np.ascontiguousarray(np.concatenate([mesh["conn"][b].reshape(-1, 8) for b in mesh["block_ids"]]),
                                    dtype=np.int32)
        egid = np.ascontiguousarray(np.concatenate([mesh["elem_gid"][b] for b in mesh["block_ids"]]), dtype=np.int32)
        ns_ids = np.ascontiguousarray(list(mesh["node_sets"].keys()), dtype=np.int32)
        ns_sizes = np.ascontiguousarray([len(v) for v in mesh["node_sets"].values()], dtype=np.int32)
        ns_nodes = (np.ascontiguousarray(np.concatenate(list(mesh["node_sets"].values())), dtype=np.int32)
                    if len(ns_ids) else np.zeros(0, np.int32))
        self.h = L.nsmref_open(self._tmp.name.encode(), len(mesh["x"]),
                               np.ascontiguousarray(mesh["node_gid"], dtype=np.int32),
                               np.ascontiguousarray(mesh["x"], dtype=np.float64),
                               np.ascontiguousarray(mesh["y"], dtype=np.float64),
                               np.ascontiguousarray(mesh["z"], dtype=np.float64),
                               len(bids), bids, nel, conn, egid, len(ns_ids), ns_ids, ns_sizes, ns_nodes,
                               1 if keep_snapshots else 0)
        err = L.nsmref_last_error(self.h)
        if err:
            raise RuntimeError(err.decode())
        self.n_nodes = L.nsmref_num_nodes(self.h)

    def close(self):
        if self.h:
            lib(self._lib_path).nsmref_close(self.h)
            self.h = None
            os.unlink(self._tmp.name)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def enable_output(self, filename: str):
        """every snapshot also goes through ModelData::WriteExodusOutput -> the reference's ExodusOutput (text form)"""
        if lib(self._lib_path).nsmref_enable_output(self.h, filename.encode()):
            raise RuntimeError(lib(self._lib_path).nsmref_last_error(self.h).decode())

    def device_launches(self) -> int:
        return int(lib(self._lib_path).nsmref_device_launches(self.h))

    def begin(self) -> float:
        return lib(self._lib_path).nsmref_begin(self.h)

    def advance(self, n: int = 1) -> float:
        return lib(self._lib_path).nsmref_advance(self.h, n)

    def internal_force(self):
        lib(self._lib_path).nsmref_internal_force(self.h)

    def contact_entities(self):
        """What the contact glue built from the deck's `contact:` line (oracle/ref_contact.cc) -> dict of the primary
        quads [nf,4], secondary quads, contact node ids (mesh node ids) and the characteristic lengths; None without contact."""
        L = lib(self._lib_path)
        if L.nsmref_contact_ints(self.h, 0, None) < 0:
            return None
        out = {}
        for which, key, width in ((0, "primary_quads", 4), (1, "secondary_quads", 4), (2, "contact_nodes", 1)):
            n = L.nsmref_contact_ints(self.h, which, None)
            a = np.empty(n * width, np.int32)
            L.nsmref_contact_ints(self.h, which, a.ctypes.data)
            out[key] = a.reshape(-1, 4) if width == 4 else a
        for which, key in ((0, "primary_char_len"), (1, "contact_node_char_len")):
            n = L.nsmref_contact_doubles(self.h, which, None)
            a = np.empty(n)
            L.nsmref_contact_doubles(self.h, which, a.ctypes.data)
            out[key] = a
        return out

    def contact_force(self, displacement):
        """-> (contact force [n,3], enforced node-face pairs) of a displacement field, through the reference's
        ContactEntity objects"""
        d = np.ascontiguousarray(displacement, dtype=np.float64)
        f = np.zeros_like(d)
        pairs = lib(self._lib_path).nsmref_contact_force(self.h, d, f)
        return f, int(pairs)

    def contact_pairs_last(self) -> int:
        return int(lib(self._lib_path).nsmref_contact_pairs_last(self.h))

    def field(self, label: str) -> np.ndarray:
        """Live (writable) view of a nodal field, AoS [n,3] (or [n] for lumped_mass)."""
        p = lib(self._lib_path).nsmref_node_field(self.h, label.encode())
        if not p:
            raise KeyError(label)
        n = self.n_nodes * (1 if label == "lumped_mass" else 3)
        a = np.ctypeslib.as_array(p, shape=(n,))
        return a if label == "lumped_mass" else a.reshape(-1, 3)

    def elem_data(self, block_id: int, which: int = 0) -> np.ndarray:
        n = lib(self._lib_path).nsmref_elem_data(self.h, block_id, which, None)
        out = np.empty(n)
        lib(self._lib_path).nsmref_elem_data(self.h, block_id, which, out.ctypes.data)
        return out.reshape(-1, 8, self.elem_stride(block_id))

    def elem_stride(self, block_id: int) -> int:
        """doubles per integration point: 15 + the material's state variables"""
        return lib(self._lib_path).nsmref_elem_stride(self.h, block_id)

    def snapshots(self):
        L = lib()
        res = []
        for i in range(L.nsmref_num_snapshots(self.h)):
            s = {"time": L.nsmref_snapshot_time(self.h, i), "node": {}, "elem": {}, "derived": {}}
            for lbl in ("lumped_mass", "reference_coordinate", "displacement", "velocity", "acceleration",
                        "internal_force", "external_force", "contact_force"):
                n = L.nsmref_snapshot_node(self.h, i, lbl.encode(), None)
                if n < 0:
                    continue
                a = np.empty(n)
                L.nsmref_snapshot_node(self.h, i, lbl.encode(), a.ctypes.data)
                s["node"][lbl] = a if lbl == "lumped_mass" else a.reshape(-1, 3)
            for b in self.mesh["block_ids"]:
                n = L.nsmref_snapshot_elem(self.h, i, int(b), None)
                a = np.empty(n)
                L.nsmref_snapshot_elem(self.h, i, int(b), a.ctypes.data)
                s["elem"][int(b)] = a.reshape(-1, 8, self.elem_stride(int(b)))
                buf = C.create_string_buffer(65536)
                L.nsmref_derived_labels(self.h, int(b), buf, 65536)
                labels = [x for x in buf.value.decode().split("\n") if x]
                d = {}
                for k, lbl in enumerate(labels):
                    n = L.nsmref_snapshot_derived(self.h, i, int(b), k, None)
                    a = np.empty(n)
                    L.nsmref_snapshot_derived(self.h, i, int(b), k, a.ctypes.data)
                    d[lbl] = a
                s["derived"][int(b)] = d
            res.append(s)
        return res


def material_stress(material_string: str, F_n, F_np1, s_n, state_n):
    """nimble::Material::GetStress of the reference (incl. the test-only state material) on n points ->
    (sigma_np1 [n,6], state_np1 [n,n_state])."""
    F_n, F_np1, s_n = (np.ascontiguousarray(a, dtype=np.float64) for a in (F_n, F_np1, s_n))
    state_n = np.ascontiguousarray(state_n, dtype=np.float64)
    s = np.empty((len(F_n), 6))
    st = np.zeros_like(state_n) if state_n.size else np.zeros((len(F_n), 0))
    ns = lib().nsmref_material_stress(material_string.encode(), len(F_n), F_n, F_np1, s_n, s,
                                      state_n if state_n.size else np.zeros(1), st if st.size else np.zeros(1))
    if ns < 0:
        raise RuntimeError("reference material failed: " + material_string)
    return s, st


def bench_steps(material: str, ref_coord, conn, lumped_mass, u, v, a, dt: float, steps: int, threads: int):
    """Reference per-element code + node loop on `threads` element chunks; returns (seconds, f)."""
    f = np.zeros_like(u)
    n_nodes = ref_coord.shape[0]
    n_elem = conn.shape[0]
    t = lib().nsmref_bench_steps(material.encode(), n_nodes, np.ascontiguousarray(ref_coord), n_elem,
                                 np.ascontiguousarray(conn, dtype=np.int32), np.ascontiguousarray(lumped_mass),
                                 u, v, a, f, dt, steps, threads)
    return t, f
