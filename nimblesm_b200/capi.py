"""nimblesm_b200/capi.py — ctypes mirror of include/nsm_b200.h (the C-ABI drop-in boundary).

Loads the in-tree ``nimblesm_b200/lib/libnsm_b200.so`` (built by ``__graft_entry__.build()`` /
``nimblesm_b200/csrc/Makefile``).  There is deliberately no fallback of any kind: a missing library raises
``NsmLibraryError`` and a missing GPU makes ``Context()`` raise ``NsmError`` (NSM_ERR_CUDA).
Nothing in this package imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# NSM_B200_LIB selects another build of the SAME library (kernel-variant experiments); never a fallback.
LIB_PATH = os.environ.get("NSM_B200_LIB") or os.path.join(_HERE, "lib", "libnsm_b200.so")

OK, ERR_ARG, ERR_CUDA, ERR_JACOBIAN, ERR_MATERIAL, ERR_COMM = range(6)
MAT_ELASTIC, MAT_NEOHOOKEAN, MAT_J2_PLASTICITY = 0, 1, 2
MATERIAL_KINDS = {"elastic": MAT_ELASTIC, "neohookean": MAT_NEOHOOKEAN, "j2_plasticity": MAT_J2_PLASTICITY}
(FIELD_LUMPED_MASS, FIELD_REFERENCE_COORDINATE, FIELD_DISPLACEMENT, FIELD_VELOCITY, FIELD_ACCELERATION,
 FIELD_INTERNAL_FORCE, FIELD_EXTERNAL_FORCE, FIELD_CONTACT_FORCE) = range(8)
FIELDS = {"lumped_mass": 0, "reference_coordinate": 1, "displacement": 2, "velocity": 3, "acceleration": 4,
          "internal_force": 5, "external_force": 6, "contact_force": 7}
BC_PRESCRIBED_VELOCITY, BC_PRESCRIBED_DISPLACEMENT = 0, 1
ASSEMBLY_ATOMIC, ASSEMBLY_ORDERED = 0, 1
FLAG_STORE_IPT_EVERY_STEP, FLAG_CACHE_REF_JACOBIAN, FLAG_REORDER_ELEMENTS, FLAG_RENUMBER_NODES = 0x1, 0x2, 0x4, 0x8
COMM_HANDLE_BYTES = 192

# every symbol include/nsm_b200.h declares (tests/test_abi.py checks header <-> library <-> this list)
SYMBOLS = [
    "nsm_b200_create", "nsm_b200_destroy", "nsm_b200_last_error", "nsm_b200_version", "nsm_b200_set_nodes",
    "nsm_b200_add_block", "nsm_b200_finalize", "nsm_b200_num_nodes", "nsm_b200_num_elements",
    "nsm_b200_device_bytes", "nsm_b200_upload_field", "nsm_b200_download_field", "nsm_b200_upload_field_async",
    "nsm_b200_download_field_async", "nsm_b200_sync", "nsm_b200_host_alloc", "nsm_b200_host_free",
    "nsm_b200_compute_lumped_mass", "nsm_b200_internal_force", "nsm_b200_internal_force_host",
    "nsm_b200_compute_stress", "nsm_b200_set_bc_table", "nsm_b200_set_bc_values", "nsm_b200_apply_kinematic_bc",
    "nsm_b200_step", "nsm_b200_get_element_data", "nsm_b200_derived_element_data", "nsm_b200_comm_init",
    "nsm_b200_comm_export", "nsm_b200_comm_attach", "nsm_b200_comm_ready", "nsm_b200_timer_start",
    "nsm_b200_timer_stop", "nsm_b200_launch_count", "nsm_b200_profile", "nsm_b200_profile_read",
    "nsm_b200_fp64_peak", "nsm_b200_cold_points", "nsm_b200_set_bc_values_steps", "nsm_b200_set_bc_programs",
    "nsm_b200_set_bc_slots_steps", "nsm_b200_get_element_components", "nsm_b200_step_host",
    "nsm_b200_get_element_data_subset", "nsm_b200_comm_set_timeout", "nsm_b200_kernel_info",
    "nsm_b200_add_block_params", "nsm_b200_material_num_state", "nsm_b200_material_num_params",
    "nsm_b200_material_state_label", "nsm_b200_material_state_initial_value", "nsm_b200_compute_stress_state",
    "nsm_b200_element_data_stride", "nsm_b200_update_states", "nsm_b200_get_element_data_previous",
    "nsm_b200_set_element_data", "nsm_b200_set_bc_entry_constants", "nsm_b200_comm_set_host_barrier",
    "nsm_b200_set_host_step_chunks", "nsm_b200_effective_flags", "nsm_b200_fp64_peak_sustained",
    "nsm_b200_set_contact", "nsm_b200_contact_force", "nsm_b200_contact_force_host", "nsm_b200_contact_stats",
    "nsm_b200_profile_read_contact", "nsm_b200_contact_status",
]


class NsmLibraryError(RuntimeError):
    pass


class NsmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("nsm_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """The loaded C-ABI library (raises NsmLibraryError when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NsmLibraryError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int32)
    lp = C.POINTER(C.c_int64)
    sig = {
        "nsm_b200_create": (i32, [i32, C.POINTER(vp)]),
        "nsm_b200_destroy": (None, [vp]),
        "nsm_b200_last_error": (C.c_char_p, [vp]),
        "nsm_b200_version": (C.c_char_p, []),
        "nsm_b200_set_nodes": (i32, [vp, i64, dp, dp, dp]),
        "nsm_b200_add_block": (i32, [vp, i32, i64, ip, i32, dbl, dbl, dbl]),
        "nsm_b200_finalize": (i32, [vp, i32, C.c_uint]),
        "nsm_b200_num_nodes": (i64, [vp]),
        "nsm_b200_num_elements": (i64, [vp, i32]),
        "nsm_b200_device_bytes": (i64, [vp]),
        "nsm_b200_upload_field": (i32, [vp, i32, vp]),
        "nsm_b200_download_field": (i32, [vp, i32, vp]),
        "nsm_b200_upload_field_async": (i32, [vp, i32, vp]),
        "nsm_b200_download_field_async": (i32, [vp, i32, vp]),
        "nsm_b200_sync": (i32, [vp]),
        "nsm_b200_host_alloc": (vp, [i64]),
        "nsm_b200_host_free": (None, [vp]),
        "nsm_b200_compute_lumped_mass": (i32, [vp, dp]),
        "nsm_b200_internal_force": (i32, [vp, i32]),
        "nsm_b200_internal_force_host": (i32, [vp, vp, vp, i32]),
        "nsm_b200_compute_stress": (i32, [vp, i32, dbl, dbl, i64, dp, dp]),
        "nsm_b200_set_bc_table": (i32, [vp, i64, ip, ip, ip]),
        "nsm_b200_set_bc_values": (i32, [vp, i64, dp]),
        "nsm_b200_set_bc_values_steps": (i32, [vp, i32, i64, dp]),
        "nsm_b200_set_bc_programs": (i32, [vp, i32, ip, ip, i32, dp, i32, i64, ip]),
        "nsm_b200_set_bc_slots_steps": (i32, [vp, i32, i32, dp]),
        "nsm_b200_apply_kinematic_bc": (i32, [vp, dbl, dbl]),
        "nsm_b200_step": (i32, [vp, i32, dp, dbl, i32]),
        "nsm_b200_step_host": (i32, [vp, dp, dbl, vp, vp, vp, vp]),
        "nsm_b200_get_element_data": (i32, [vp, i32, dp]),
        "nsm_b200_derived_element_data": (i32, [vp, i32, dp]),
        "nsm_b200_get_element_components": (i32, [vp, i32, i32, ip, dp]),
        "nsm_b200_comm_init": (i32, [vp, i32, i32, i32, ip, lp, ip]),
        "nsm_b200_comm_export": (i32, [vp, C.c_char_p]),
        "nsm_b200_comm_attach": (i32, [vp, i32, C.c_char_p]),
        "nsm_b200_comm_ready": (i32, [vp]),
        "nsm_b200_timer_start": (i32, [vp]),
        "nsm_b200_timer_stop": (i32, [vp, C.POINTER(C.c_float)]),
        "nsm_b200_launch_count": (i64, [vp]),
        "nsm_b200_profile": (i32, [vp, i32]),
        "nsm_b200_profile_read": (i32, [vp, dp, dp, lp]),
        "nsm_b200_fp64_peak": (i32, [vp, dp, dp]),
        "nsm_b200_cold_points": (i64, [vp]),
        "nsm_b200_get_element_data_subset": (i32, [vp, i32, i64, lp, dp]),
        "nsm_b200_comm_set_timeout": (i32, [vp, dbl]),
        "nsm_b200_kernel_info": (C.c_char_p, []),
        "nsm_b200_add_block_params": (i32, [vp, i32, i64, ip, i32, i32, dp]),
        "nsm_b200_material_num_state": (i32, [i32]),
        "nsm_b200_material_num_params": (i32, [i32]),
        "nsm_b200_material_state_label": (C.c_char_p, [i32, i32]),
        "nsm_b200_material_state_initial_value": (dbl, [i32, i32]),
        "nsm_b200_compute_stress_state": (i32, [vp, i32, i32, dp, i64, dp, dp, dp, dp, dp, dp]),
        "nsm_b200_element_data_stride": (i32, [vp, i32]),
        "nsm_b200_update_states": (i32, [vp]),
        "nsm_b200_get_element_data_previous": (i32, [vp, i32, dp]),
        "nsm_b200_set_element_data": (i32, [vp, i32, i32, dp]),
        "nsm_b200_set_bc_entry_constants": (i32, [vp, i32, i64, dp]),
        "nsm_b200_comm_set_host_barrier": (i32, [vp, vp, vp]),
        "nsm_b200_set_host_step_chunks": (i32, [vp, i32]),
        "nsm_b200_effective_flags": (C.c_uint, [vp]),
        "nsm_b200_fp64_peak_sustained": (i32, [vp, dbl, dp]),
        "nsm_b200_set_contact": (i32, [vp, dbl, i64, ip, dp, i64, ip, dp]),
        "nsm_b200_contact_force": (i32, [vp]),
        "nsm_b200_contact_force_host": (i32, [vp, vp, vp]),
        "nsm_b200_contact_stats": (i32, [vp, lp]),
        "nsm_b200_contact_status": (i32, [vp, vp, vp]),
        "nsm_b200_profile_read_contact": (i32, [vp, dp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _lptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


class PinnedArray:
    """A numpy view over page-locked host memory from nsm_b200_host_alloc."""

    def __init__(self, shape, dtype=np.float64):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._p = lib().nsm_b200_host_alloc(n)
        if not self._p:
            raise NsmError(ERR_CUDA, "pinned host allocation of %d bytes failed" % n)
        buf = (C.c_char * n).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self._p:
            self.array = None
            lib().nsm_b200_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One GPU's share of a model: thin object wrapper over the nsm_b200_* calls."""

    def __init__(self, device: int = 0):
        self._L = lib()
        h = C.c_void_p()
        rc = self._L.nsm_b200_create(device, C.byref(h))
        if rc:
            raise NsmError(rc, self._L.nsm_b200_last_error(None).decode())
        self._h = h
        self.device = device
        self.n_nodes = 0
        self.block_ids = []
        self.block_nelem = {}
        self.block_stride = {}  # doubles per integration point: 15 + the material's state variables

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.nsm_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc:
            raise NsmError(rc, self._L.nsm_b200_last_error(self._h).decode())

    # -- model --------------------------------------------------------------------------------------
    def set_nodes(self, x, y, z):
        x, y, z = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z))
        self.n_nodes = len(x)
        self._ck(self._L.nsm_b200_set_nodes(self._h, len(x), _dptr(x), _dptr(y), _dptr(z)))

    def add_block(self, block_id, conn, material, bulk_modulus, shear_modulus, density, *extra):
        """extra: the model's own parameters after (bulk, shear, density) -- j2_plasticity: yield_stress, hardening_modulus"""
        conn = np.ascontiguousarray(conn, dtype=np.int32).reshape(-1, 8)
        kind = MATERIAL_KINDS[material] if isinstance(material, str) else int(material)
        params = np.array([bulk_modulus, shear_modulus, density, *extra], dtype=np.float64)
        self._ck(self._L.nsm_b200_add_block_params(self._h, int(block_id), len(conn), _iptr(conn), kind, len(params), _dptr(params)))
        self.block_ids.append(int(block_id))
        self.block_nelem[int(block_id)] = len(conn)
        self.block_stride[int(block_id)] = 15 + int(self._L.nsm_b200_material_num_state(kind))

    def finalize(self, assembly=ASSEMBLY_ATOMIC, flags=0):
        self._ck(self._L.nsm_b200_finalize(self._h, assembly, flags))

    @property
    def n_elements(self):
        return int(self._L.nsm_b200_num_elements(self._h, -1))

    @property
    def effective_flags(self):
        return int(self._L.nsm_b200_effective_flags(self._h))

    @property
    def device_bytes(self):
        return int(self._L.nsm_b200_device_bytes(self._h))

    # -- fields -------------------------------------------------------------------------------------
    def upload(self, field, host):
        fid = FIELDS[field] if isinstance(field, str) else field
        host = np.ascontiguousarray(host, dtype=np.float64)
        assert host.size == self.n_nodes * (1 if fid == 0 else 3), "field size mismatch"
        self._ck(self._L.nsm_b200_upload_field(self._h, fid, host.ctypes.data))

    def download(self, field, out=None):
        fid = FIELDS[field] if isinstance(field, str) else field
        if out is None:
            out = np.empty(self.n_nodes if fid == 0 else (self.n_nodes, 3))
        self._ck(self._L.nsm_b200_download_field(self._h, fid, out.ctypes.data))
        return out

    def upload_async(self, field, host):
        fid = FIELDS[field] if isinstance(field, str) else field
        self._ck(self._L.nsm_b200_upload_field_async(self._h, fid, host.ctypes.data))

    def download_async(self, field, out):
        fid = FIELDS[field] if isinstance(field, str) else field
        self._ck(self._L.nsm_b200_download_field_async(self._h, fid, out.ctypes.data))

    def sync(self):
        self._ck(self._L.nsm_b200_sync(self._h))

    # -- path ---------------------------------------------------------------------------------------
    def compute_lumped_mass(self) -> float:
        dt = C.c_double()
        self._ck(self._L.nsm_b200_compute_lumped_mass(self._h, C.byref(dt)))
        return dt.value

    def internal_force(self, store_ipt=False):
        self._ck(self._L.nsm_b200_internal_force(self._h, 1 if store_ipt else 0))

    def internal_force_host(self, displacement, out=None, store_ipt=False):
        displacement = np.ascontiguousarray(displacement, dtype=np.float64)
        if out is None:
            out = np.empty((self.n_nodes, 3))
        self._ck(self._L.nsm_b200_internal_force_host(self._h, displacement.ctypes.data, out.ctypes.data,
                                                       1 if store_ipt else 0))
        return out

    def set_contact(self, penalty, primary_quads, primary_char_len, contact_nodes, contact_node_char_len):
        """Contact entities (ContactManager::CreateContactEntities): skin quads of the primary blocks and contact nodes
        of the secondary blocks with their characteristic lengths; empty lists switch contact off."""
        q = np.ascontiguousarray(primary_quads, dtype=np.int32).reshape(-1, 4)
        ql = np.ascontiguousarray(primary_char_len, dtype=np.float64)
        cn = np.ascontiguousarray(contact_nodes, dtype=np.int32)
        cl = np.ascontiguousarray(contact_node_char_len, dtype=np.float64)
        assert len(q) == len(ql) and len(cn) == len(cl)
        self._ck(self._L.nsm_b200_set_contact(self._h, float(penalty), len(q), _iptr(q), _dptr(ql), len(cn), _iptr(cn), _dptr(cl)))

    def contact_force(self):
        self._ck(self._L.nsm_b200_contact_force(self._h))

    def contact_force_host(self, displacement=None, out=None):
        if displacement is not None:
            displacement = np.ascontiguousarray(displacement, dtype=np.float64)
        if out is None:
            out = np.empty((self.n_nodes, 3))
        self._ck(self._L.nsm_b200_contact_force_host(self._h, displacement.ctypes.data if displacement is not None else None,
                                                      out.ctypes.data))
        return out

    def contact_stats(self):
        """-> dict of the last evaluation's counters (pairs enforced / box-tested, active faces / nodes)"""
        st = np.zeros(5, np.int64)
        self._ck(self._L.nsm_b200_contact_stats(self._h, st.ctypes.data_as(C.POINTER(C.c_int64))))
        return {"pairs": int(st[0]), "box_tested": int(st[1]), "active_faces": int(st[2]), "active_nodes": int(st[3]),
                "ordered_overflow_pairs": int(st[4])}

    def contact_status(self, n_faces, n_contact_nodes):
        """-> (status of the 4 * n_faces triangles, status of the contact nodes) after the last evaluation (uint8, 1 = in contact)"""
        face, node = np.zeros(4 * int(n_faces), np.uint8), np.zeros(int(n_contact_nodes), np.uint8)
        self._ck(self._L.nsm_b200_contact_status(self._h, face.ctypes.data, node.ctypes.data))
        return face, node

    def compute_stress(self, material, bulk_modulus, shear_modulus, def_grad):
        kind = MATERIAL_KINDS[material] if isinstance(material, str) else int(material)
        F = np.ascontiguousarray(def_grad, dtype=np.float64).reshape(-1, 9)
        s = np.empty((len(F), 6))
        self._ck(self._L.nsm_b200_compute_stress(self._h, kind, bulk_modulus, shear_modulus, len(F), _dptr(F),
                                                  _dptr(s)))
        return s

    def set_bc_table(self, node, comp, kind):
        node, comp, kind = (np.ascontiguousarray(a, dtype=np.int32) for a in (node, comp, kind))
        self._n_bc = len(node)
        self._ck(self._L.nsm_b200_set_bc_table(self._h, len(node), _iptr(node), _iptr(comp), _iptr(kind)))

    def set_bc_values(self, value):
        value = np.ascontiguousarray(value, dtype=np.float64)
        self._ck(self._L.nsm_b200_set_bc_values(self._h, len(value), _dptr(value)))

    def set_bc_values_steps(self, values):
        """values[r][k]: magnitude of table entry k at the time of step r of the next step() call."""
        values = np.ascontiguousarray(values, dtype=np.float64)
        assert values.ndim == 2
        self._ck(self._L.nsm_b200_set_bc_values_steps(self._h, values.shape[0], values.shape[1], _dptr(values)))

    def set_bc_programs(self, offsets, code, consts, n_slots, program_of_entry):
        """Device-evaluated magnitudes (include/nsm_b200.h, nsm_bc_op): postfix programs, one id (or -1) per entry."""
        offsets, code, poe = (np.ascontiguousarray(a, dtype=np.int32) for a in (offsets, code, program_of_entry))
        consts = np.ascontiguousarray(consts, dtype=np.float64)
        self._ck(self._L.nsm_b200_set_bc_programs(self._h, len(offsets) - 1, _iptr(offsets), _iptr(code), len(consts),
                                                   _dptr(consts), int(n_slots), len(poe), _iptr(poe)))

    def set_bc_entry_constants(self, values):
        """values[j][k]: host-evaluated per-entry constant j (a function of the position alone) for table entry k."""
        values = np.ascontiguousarray(values, dtype=np.float64)
        assert values.ndim == 2
        self._ck(self._L.nsm_b200_set_bc_entry_constants(self._h, values.shape[0], values.shape[1], _dptr(values)))

    def set_bc_slots_steps(self, slots):
        """slots[r][s]: host-evaluated scalar s (a function of t alone) at the time of step r of the next step() call."""
        slots = np.ascontiguousarray(slots, dtype=np.float64)
        assert slots.ndim == 2
        self._ck(self._L.nsm_b200_set_bc_slots_steps(self._h, slots.shape[0], slots.shape[1], _dptr(slots)))

    def apply_kinematic_bc(self, time_current, time_previous):
        self._ck(self._L.nsm_b200_apply_kinematic_bc(self._h, time_current, time_previous))

    def step(self, n_steps, time, dt_user, store_ipt_last=False) -> float:
        t = C.c_double(time)
        self._ck(self._L.nsm_b200_step(self._h, int(n_steps), C.byref(t), dt_user, 1 if store_ipt_last else 0))
        return t.value

    def set_host_step_chunks(self, n_chunks):
        """node chunks of the pipelined step_host (-1 automatic, <= 1 the plain schedule); before the first step_host"""
        self._ck(self._L.nsm_b200_set_host_step_chunks(self._h, int(n_chunks)))

    def step_host(self, time, dt_user, displacement, velocity, acceleration, internal_force) -> float:
        """One explicit step on host-resident [n][3] float64 arrays, updated in place (pinned arrays overlap copies)."""
        for a in (displacement, velocity, acceleration, internal_force):
            assert a is None or (a.dtype == np.float64 and a.flags.c_contiguous and a.shape == (self.n_nodes, 3))
        t = C.c_double(time)
        self._ck(self._L.nsm_b200_step_host(self._h, C.byref(t), dt_user, displacement.ctypes.data, velocity.ctypes.data,
                                             acceleration.ctypes.data, internal_force.ctypes.data if internal_force is not None else None))
        return t.value

    def element_data(self, block_id, previous=False):
        """[n_elem][8][15 + n_state] records: the most recently computed ones, or (previous) the N records."""
        out = np.empty((self.block_nelem[block_id], 8, self.block_stride[block_id]))
        fn = self._L.nsm_b200_get_element_data_previous if previous else self._L.nsm_b200_get_element_data
        self._ck(fn(self._h, block_id, _dptr(out)))
        return out

    def set_element_data(self, block_id, data, previous=False):
        data = np.ascontiguousarray(data, dtype=np.float64)
        assert data.shape == (self.block_nelem[block_id], 8, self.block_stride[block_id])
        self._ck(self._L.nsm_b200_set_element_data(self._h, block_id, 1 if previous else 0, _dptr(data)))

    def update_states(self):
        self._ck(self._L.nsm_b200_update_states(self._h))

    def compute_stress_state(self, material, params, F_n, F_np1, s_n, state_n):
        """Full material seam: params = [bulk, shear, density, model-specific...] -> (sigma_np1, state_np1)."""
        kind = MATERIAL_KINDS[material] if isinstance(material, str) else int(material)
        params = np.ascontiguousarray(params, dtype=np.float64)
        F_n, F_np1, s_n, state_n = (np.ascontiguousarray(a, dtype=np.float64) for a in (F_n, F_np1, s_n, state_n))
        s, st = np.empty((len(F_np1), 6)), np.empty_like(state_n)
        self._ck(self._L.nsm_b200_compute_stress_state(self._h, kind, len(params), _dptr(params), len(F_np1), _dptr(F_n), _dptr(F_np1),
                                                        _dptr(s_n), _dptr(state_n), _dptr(s), _dptr(st)))
        return s, st

    def element_components(self, block_id, offsets):
        """out[k][e] = integration-point value offsets[k] (0..119 = 15*point + field) of element e, split on the device."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        out = np.empty((len(offsets), self.block_nelem[block_id]))
        self._ck(self._L.nsm_b200_get_element_components(self._h, block_id, len(offsets), _iptr(offsets), _dptr(out)))
        return out

    def element_data_subset(self, block_id, elements):
        """[len(elements)][8][15] integration-point records of the listed elements (file-order indices in the block)."""
        el = np.ascontiguousarray(elements, dtype=np.int64)
        out = np.empty((len(el), 8, self.block_stride[block_id]))
        self._ck(self._L.nsm_b200_get_element_data_subset(self._h, block_id, len(el), _lptr(el), _dptr(out)))
        return out

    def derived_element_data(self, block_id):
        out = np.empty((1 + self.block_stride[block_id], self.block_nelem[block_id]))
        self._ck(self._L.nsm_b200_derived_element_data(self._h, block_id, _dptr(out)))
        return out

    # -- peer exchange ------------------------------------------------------------------------------
    def comm_init(self, rank, world_size, peer_ranks, pair_offsets, pair_local_nodes):
        pr = np.ascontiguousarray(peer_ranks, dtype=np.int32)
        po = np.ascontiguousarray(pair_offsets, dtype=np.int64)
        pn = np.ascontiguousarray(pair_local_nodes, dtype=np.int32)
        self._ck(self._L.nsm_b200_comm_init(self._h, rank, world_size, len(pr), _iptr(pr), _lptr(po), _iptr(pn)))

    def comm_export(self) -> bytes:
        buf = C.create_string_buffer(COMM_HANDLE_BYTES)
        self._ck(self._L.nsm_b200_comm_export(self._h, buf))
        return buf.raw

    def comm_attach(self, peer_rank, blob: bytes):
        assert len(blob) == COMM_HANDLE_BYTES
        self._ck(self._L.nsm_b200_comm_attach(self._h, peer_rank, blob))

    def comm_ready(self):
        self._ck(self._L.nsm_b200_comm_ready(self._h))

    def comm_set_timeout(self, seconds):
        self._ck(self._L.nsm_b200_comm_set_timeout(self._h, float(seconds)))

    # -- measurement --------------------------------------------------------------------------------
    def timer_start(self):
        self._ck(self._L.nsm_b200_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._ck(self._L.nsm_b200_timer_stop(self._h, C.byref(ms)))
        return ms.value

    @property
    def launch_count(self):
        return int(self._L.nsm_b200_launch_count(self._h))

    def profile(self, enable=True):
        self._ck(self._L.nsm_b200_profile(self._h, 1 if enable else 0))

    def profile_read(self):
        e, n, k = C.c_double(), C.c_double(), C.c_int64()
        self._ck(self._L.nsm_b200_profile_read(self._h, C.byref(e), C.byref(n), C.byref(k)))
        return e.value, n.value, k.value

    def profile_read_contact(self):
        """average device time of the contact evaluation per profiled step, ms"""
        m = C.c_double()
        self._ck(self._L.nsm_b200_profile_read_contact(self._h, C.byref(m)))
        return m.value

    @property
    def cold_points(self):
        return int(self._L.nsm_b200_cold_points(self._h))

    def fp64_peak(self):
        a, b = C.c_double(), C.c_double()
        self._ck(self._L.nsm_b200_fp64_peak(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value


def _fp64_peak_sustained(self, seconds=2.0):
    """DADD+DMUL rate of the last quarter of `seconds` of back-to-back launches (1e12 lane-ops/s)"""
    a = C.c_double()
    self._ck(self._L.nsm_b200_fp64_peak_sustained(self._h, float(seconds), C.byref(a)))
    return a.value


Context.fp64_peak_sustained = _fp64_peak_sustained


def version() -> str:
    return lib().nsm_b200_version().decode()


def kernel_info() -> dict:
    """nsm_b200_kernel_info(): the element kernels' static instruction mix, read off this binary's SASS at build time."""
    import json

    return json.loads(lib().nsm_b200_kernel_info().decode())
