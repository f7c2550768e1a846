"""oracle/hex8.py — TEST INFRASTRUCTURE: ctypes loader for oracle/_ref/libhex8_oracle.so (hex8_oracle.c)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libhex8_oracle.so")
ELASTIC, NEOHOOKEAN, J2_PLASTICITY = 0, 1, 2

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lib = None


def available() -> bool:
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        d, i, l, vp = C.c_double, C.c_int, C.c_long, C.c_void_p
        L.h8o_shape_tables.argtypes = [_dp, _dp]
        L.h8o_def_grad.argtypes = [_dp, _dp, _dp]
        L.h8o_stress_elastic.argtypes = [d, d, _dp, _dp]
        L.h8o_stress_neohookean.argtypes = [d, d, _dp, _dp]
        L.h8o_polar_left_stretch.argtypes = [_dp, _dp]
        L.h8o_invert_full33.restype = d
        L.h8o_invert_full33.argtypes = [_dp, _dp]
        L.h8o_nodal_forces.argtypes = [_dp, _dp, _dp]
        L.h8o_lumped_mass.argtypes = [d, _dp, _dp]
        L.h8o_char_length.restype = d
        L.h8o_char_length.argtypes = [_dp]
        L.h8o_block_internal_force.argtypes = [i, d, d, _dp, _dp, l, _ip, _dp, vp]
        L.h8o_block_lumped_mass.argtypes = [d, _dp, l, _ip, _dp]
        L.h8o_block_critical_dt.restype = d
        L.h8o_block_critical_dt.argtypes = [d, d, _dp, _dp, l, _ip]
        L.h8o_block_derived.argtypes = [_dp, _dp, l, _ip, _dp, _dp]
        L.h8o_stress_j2.argtypes = [_dp, _dp, _dp, _dp, _dp, _dp, _dp]
        L.h8o_num_state.restype = i
        L.h8o_num_state.argtypes = [i]
        L.h8o_block_internal_force_state.argtypes = [i, _dp, _dp, _dp, l, _ip, _dp, vp, vp]
        L.h8o_block_derived_stride.argtypes = [_dp, _dp, l, _ip, _dp, i, _dp]
        L.h8o_axpy.argtypes = [l, d, _dp, _dp]
        L.h8o_accel.argtypes = [l, _dp, _dp, vp, _dp]
        L.h8o_bench_steps.restype = d
        L.h8o_bench_steps.argtypes = [i, d, d, l, _dp, l, _ip, _dp, _dp, _dp, _dp, _dp, d, i, i]
        _lib = L
    return _lib


def internal_force(material, bulk, shear, ref, disp, conn, want_elem_data=True):
    """-> (f [n,3], elem_data [ne,8,15] or None); AoS inputs."""
    ref = np.ascontiguousarray(ref, dtype=np.float64)
    disp = np.ascontiguousarray(disp, dtype=np.float64)
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    f = np.zeros_like(ref)
    ed = np.empty((len(conn), 8, 15)) if want_elem_data else None
    lib().h8o_block_internal_force(material, bulk, shear, ref, disp, len(conn), conn, f,
                                   ed.ctypes.data if ed is not None else None)
    return f, ed


def internal_force_state(material, params, ref, disp, conn, elem_data_n=None, f=None):
    """Any material incl. the history-dependent one -> (f [n,3], elem_data_np1 [ne,8,15+n_state]).  params =
    [bulk, shear, material-specific...]; elem_data_n is the previous record array (required when n_state > 0).
    f: accumulate into this array (the serial reference adds block after block into one array,
    src/nimble_model_data.cc:636-659) instead of a fresh zero one."""
    ref = np.ascontiguousarray(ref, dtype=np.float64)
    disp = np.ascontiguousarray(disp, dtype=np.float64)
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    params = np.ascontiguousarray(params, dtype=np.float64)
    ns = lib().h8o_num_state(material)
    if f is None:
        f = np.zeros_like(ref)
    ed = np.empty((len(conn), 8, 15 + ns))
    if ns:
        elem_data_n = np.ascontiguousarray(elem_data_n, dtype=np.float64)
        assert elem_data_n.shape == ed.shape
    lib().h8o_block_internal_force_state(material, params, ref, disp, len(conn), conn, f,
                                         elem_data_n.ctypes.data if ns else None, ed.ctypes.data)
    return f, ed


def initial_elem_data(material, n_elem):
    """Block::InitializeElementData (src/nimble_block.cc:148-207): F = I, sigma = 0, state = initial values (0)."""
    ed = np.zeros((n_elem, 8, 15 + lib().h8o_num_state(material)))
    ed[:, :, :3] = 1.0
    return ed


def stress_j2(params, Fn, Fnp1, sn, state_n):
    """Point-wise seam of the history-dependent material -> (sigma_np1 [n,6], state_np1 [n,2])."""
    params = np.ascontiguousarray(params, dtype=np.float64)
    n = len(Fn)
    s, st = np.empty((n, 6)), np.empty((n, 2))
    for k in range(n):
        lib().h8o_stress_j2(params, np.ascontiguousarray(Fn[k]), np.ascontiguousarray(Fnp1[k]), np.ascontiguousarray(sn[k]),
                            np.ascontiguousarray(state_n[k]), s[k], st[k])
    return s, st


def derived_stride(ref, disp, conn, elem_data):
    """-> [1 + stride, ne]: volume, then volume averages of every per-point field of the record."""
    stride = elem_data.shape[-1]
    out = np.empty((1 + stride, len(conn)))
    lib().h8o_block_derived_stride(np.ascontiguousarray(ref), np.ascontiguousarray(disp), len(conn),
                                   np.ascontiguousarray(conn, dtype=np.int32), np.ascontiguousarray(elem_data), stride, out)
    return out


def lumped_mass(density, ref, conn):
    ref = np.ascontiguousarray(ref, dtype=np.float64)
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    m = np.zeros(len(ref))
    lib().h8o_block_lumped_mass(density, ref, len(conn), conn, m)
    return m


def critical_dt(bulk, density, ref, disp, conn):
    return lib().h8o_block_critical_dt(bulk, density, np.ascontiguousarray(ref), np.ascontiguousarray(disp),
                                       len(conn), np.ascontiguousarray(conn, dtype=np.int32))


def derived(ref, disp, conn, elem_data):
    """-> [16, ne]: volume, then volume averages of F(9) and sigma(6)."""
    out = np.empty((16, len(conn)))
    lib().h8o_block_derived(np.ascontiguousarray(ref), np.ascontiguousarray(disp), len(conn),
                            np.ascontiguousarray(conn, dtype=np.int32), np.ascontiguousarray(elem_data), out)
    return out


def bench_steps(material, bulk, shear, ref, conn, mass, u, v, a, dt, steps, threads):
    f = np.zeros_like(u)
    t = lib().h8o_bench_steps(material, bulk, shear, len(ref), np.ascontiguousarray(ref), len(conn),
                              np.ascontiguousarray(conn, dtype=np.int32), np.ascontiguousarray(mass), u, v, a, f,
                              dt, steps, threads)
    return t, f
