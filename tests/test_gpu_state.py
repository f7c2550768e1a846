"""GPU parity of the state-variable slot: per-integration-point state N / N+1 carried through the
element kernel, rolled by nsm_b200_update_states / nsm_b200_step, exposed through the element-data getters and the
C++ driver's Exodus output.

The oracle is the REFERENCE'S OWN plumbing: a test-only nimble::Material subclass with two state variables
(oracle/ref_state_material.cc, "j2_plasticity") runs behind the unmodified nimble_block.cc / nimble_model_data.cc
(label order src/nimble_block.cc:84-108, initial values :168-183, F_n / sigma_n / state_n hand-over :297-368,
swap src/nimble_model_data.h:104-107); tests/golden/state_cube.npz holds its snapshots, and oracle/hex8_oracle.c
restates it bit for bit (tests/test_oracle.py).  Bars: bit equality of sigma, state and (ORDERED) nodal force.
"""
import os
import re
import subprocess

import numpy as np
import pytest

from tests.conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu

K, G, RHO, Y, H = 1.6e12, 0.8e12, 7.8, 5.0e8, 2.0e10
PARAMS_ORACLE = [K, G, Y, H]
EXE = os.path.join(ROOT, "nimblesm_b200", "lib", "NimbleSM_b200")


def _bits(a):
    return np.ascontiguousarray(a).view(np.int64)


def test_state_material_seam_bitwise(oracle):
    """nsm_b200_compute_stress_state == h8o_stress_j2 on elastic, plastic, virgin (zero stress) and unloading points."""
    from nimblesm_b200 import capi

    rng = np.random.default_rng(11)
    n = 20000
    Fn = np.zeros((n, 9))
    Fn[:, :3] = 1.0
    Fn += 1e-3 * (2 * rng.random((n, 9)) - 1)
    Fnp1 = Fn + 5e-4 * (2 * rng.random((n, 9)) - 1) * rng.random((n, 1)) ** 3
    sn = 1.5e8 * (2 * rng.random((n, 6)) - 1)
    st = np.stack([1e-3 * rng.random(n), np.zeros(n)], 1)
    # virgin points (the first step of every run): F_n = I, sigma_n = 0, no increment at all / a tiny one
    Fn[:200] = 0.0
    Fn[:200, :3] = 1.0
    sn[:200] = 0.0
    st[:200] = 0.0
    Fnp1[:100] = Fn[:100]
    want_s, want_st = oracle.stress_j2(PARAMS_ORACLE, Fn, Fnp1, sn, st)
    with capi.Context(0) as c:
        got_s, got_st = c.compute_stress_state("j2_plasticity", [K, G, RHO, Y, H], Fn, Fnp1, sn, st)
        assert c.cold_points in (0, -1)
    assert np.array_equal(_bits(got_s), _bits(want_s))
    assert np.array_equal(_bits(got_st), _bits(want_st))
    plastic = want_st[:, 0] > st[:, 0]
    assert 0.2 < plastic.mean() < 0.9
    from nimblesm_b200.capi import lib

    L = lib()
    assert L.nsm_b200_material_num_state(2) == 2 and L.nsm_b200_material_num_state(1) == 0
    assert L.nsm_b200_material_state_label(2, 0) == b"equivalent_plastic_strain"
    assert L.nsm_b200_material_state_label(2, 1) == b"von_mises_stress" and L.nsm_b200_material_num_params(2) == 5


def _oracle_loop(oracle, mesh, n_steps, dt_user, blocks, is_output=lambda step: False):
    """The explicit loop on the plain-C oracle with per-block records; yields after every step.  On an output step the
    kinematic conditions are applied once more after the second half-kick (explicit_time_integrator.cc:266-269)."""
    L = oracle.lib()
    ref = np.ascontiguousarray(np.stack([mesh["x"], mesh["y"], mesh["z"]], 1))
    m = np.zeros(len(ref))
    for b in sorted(blocks):
        L.h8o_block_lumped_mass(RHO, ref, len(mesh["conn"][b]), np.ascontiguousarray(mesh["conn"][b], dtype=np.int32), m)
    u, v, a = np.zeros_like(ref), np.zeros_like(ref), np.zeros_like(ref)
    v[:, 0] = 1000.0 * ref[:, 0]
    face = mesh["node_sets"][2]
    v[face] = 0.0
    ed = {b: oracle.initial_elem_data(blocks[b][0], len(mesh["conn"][b])) for b in blocks}
    t = 0.0
    for step in range(n_steps):
        t_prev, t = t, t + dt_user
        d = t - t_prev
        L.h8o_axpy(u.size, 0.5 * d, a.ravel(), v.ravel())
        v[face] = 0.0
        L.h8o_axpy(u.size, d, v.ravel(), u.ravel())
        v[face] = 0.0
        f = np.zeros_like(ref)
        for b in sorted(blocks):  # ascending block id, all into one array: the serial summation order
            kind, params = blocks[b]
            f, ed[b] = oracle.internal_force_state(kind, params, ref, u, mesh["conn"][b], ed[b], f)
        L.h8o_accel(len(ref), m, f, None, a)
        L.h8o_axpy(u.size, 0.5 * d, a.ravel(), v.ravel())
        if is_output(step):
            v[face] = 0.0
        yield step, t, u, v, a, f, ed


def _two_block_cube(n):
    from nimblesm_b200.mesh import structured_cube

    return structured_cube(n, block_of_element=lambda i, j, k: np.where(k < n // 2, 1, 2))


def _ctx(mesh, assembly, flags):
    from nimblesm_b200 import capi

    c = capi.Context(0)
    c.set_nodes(mesh["x"], mesh["y"], mesh["z"])
    c.add_block(1, mesh["conn"][1], "j2_plasticity", K, G, RHO, Y, H)
    c.add_block(2, mesh["conn"][2], "neohookean", K, G, RHO)
    c.finalize(assembly, flags)
    c.compute_lumped_mass()
    ref = np.stack([mesh["x"], mesh["y"], mesh["z"]], 1)
    v0 = np.zeros_like(ref)
    v0[:, 0] = 1000.0 * ref[:, 0]
    c.upload("velocity", v0)
    face = mesh["node_sets"][2]
    c.set_bc_table(np.repeat(face, 3), np.tile(np.arange(3, dtype=np.int32), len(face)), np.zeros(3 * len(face), np.int32))
    c.set_bc_values(np.zeros(3 * len(face)))
    c.apply_kinematic_bc(0.0, 0.0)
    return c


@pytest.mark.parametrize("flags", [0, 2])
def test_state_trajectory_bitwise_ordered(oracle, flags):
    """50 steps of a 10^3 cube (block 1 = the material with state, block 2 = neohookean; ragged against the 4-element
    warp groups), ORDERED assembly: nodal force every step, and sigma / state records, u, v, a at the end equal the
    oracle's bit for bit -- stepping one step per call, in one call, and in chunks; the N records equal the oracle's
    previous records; most points yield."""
    from nimblesm_b200 import capi

    n = 10
    mesh = _two_block_cube(n)
    dt = 0.2 * (1.0 / n) / float(np.sqrt(K / RHO))
    blocks = {1: (oracle.J2_PLASTICITY, PARAMS_ORACLE), 2: (oracle.NEOHOOKEAN, [K, G])}
    c = _ctx(mesh, capi.ASSEMBLY_ORDERED, flags)
    assert c.block_stride == {1: 17, 2: 15}
    t = 0.0
    prev = None
    for step, t_o, u, v, a, f, ed in _oracle_loop(oracle, mesh, 50, dt, blocks, lambda s: s % 10 == 9):
        t = c.step(1, t, dt, store_ipt_last=(step % 10 == 9))  # an output step: records stored, BCs re-applied
        assert t == t_o
        assert np.array_equal(_bits(c.download("internal_force")), _bits(f)), step
        if step % 10 == 9:
            assert np.array_equal(_bits(c.element_data(1)), _bits(ed[1])), step
            assert np.array_equal(_bits(c.element_data(2)), _bits(ed[2])), step
            assert np.array_equal(_bits(c.element_data(1, previous=True)), _bits(prev)), step
        prev = ed[1].copy()
        final = (u.copy(), v.copy(), a.copy(), f.copy(), {b: ed[b].copy() for b in ed})
    for lbl, want in zip(("displacement", "velocity", "acceleration"), final[:3]):
        assert np.array_equal(_bits(c.download(lbl)), _bits(want)), lbl
    eqps = final[4][1][:, :, 15]
    assert (eqps > 0).mean() > 0.5 and eqps.max() > 1e-4
    # derived (volume-averaged) state of the block
    der = c.derived_element_data(1)
    ref = np.stack([mesh["x"], mesh["y"], mesh["z"]], 1)
    want = oracle.derived_stride(ref, final[0], mesh["conn"][1], final[4][1])
    assert der.shape == (18, len(mesh["conn"][1])) and np.array_equal(_bits(der), _bits(want))
    comps = c.element_components(1, [15, 16, 17 * 7 + 15, 17 * 2 + 9])
    rec = final[4][1].reshape(len(mesh["conn"][1]), -1)
    for k, o in enumerate([15, 16, 17 * 7 + 15, 17 * 2 + 9]):
        assert np.array_equal(_bits(comps[k]), _bits(rec[:, o]))
    sub = c.element_data_subset(1, [0, 5, len(rec) - 1])
    assert np.array_equal(_bits(sub), _bits(final[4][1][[0, 5, len(rec) - 1]]))
    c.close()
    # one call / chunks: the per-step roll inside nsm_b200_step
    for chunks in ([10] * 5, [3, 7, 10, 10, 1, 9, 10]):  # every call ends where the first run had its output steps
        c = _ctx(mesh, capi.ASSEMBLY_ORDERED, flags)
        t, done = 0.0, 0
        for k in chunks:
            done += k
            t = c.step(k, t, dt, store_ipt_last=(done % 10 == 0))
        for lbl, want in zip(("displacement", "velocity", "acceleration", "internal_force"), final[:4]):
            assert np.array_equal(_bits(c.download(lbl)), _bits(want)), (chunks, lbl)
        assert np.array_equal(_bits(c.element_data(1)), _bits(final[4][1])), chunks
        c.close()


def test_state_explicit_sequence_and_restart(oracle):
    """nsm_b200_internal_force + nsm_b200_update_states sequenced by the caller (the reference's loop body) gives the
    bits of the fused stepping; records written back with nsm_b200_set_element_data resume a run bit for bit; ATOMIC
    assembly stays within 1e-12 of the oracle's force on the same state."""
    from nimblesm_b200 import capi

    n = 6
    mesh = _two_block_cube(n)
    dt = 0.2 * (1.0 / n) / float(np.sqrt(K / RHO))
    blocks = {1: (oracle.J2_PLASTICITY, PARAMS_ORACLE), 2: (oracle.NEOHOOKEAN, [K, G])}
    traj = [(u.copy(), f.copy(), ed[1].copy()) for _s, _t, u, _v, _a, f, ed in _oracle_loop(oracle, mesh, 12, dt, blocks)]
    # caller-sequenced: upload u_k, force, UpdateStates -- repeated evaluation without a roll recomputes the same N+1
    c = _ctx(mesh, capi.ASSEMBLY_ORDERED, 2)
    for k, (u, f, ed1) in enumerate(traj):
        c.upload("displacement", u)
        c.internal_force(store_ipt=True)
        if k == 3:
            c.internal_force(store_ipt=True)  # same N records: same result
        assert np.array_equal(_bits(c.download("internal_force")), _bits(f)), k
        assert np.array_equal(_bits(c.element_data(1)), _bits(ed1)), k
        c.update_states()
    c.close()
    # restart at step 6 from saved records
    c = _ctx(mesh, capi.ASSEMBLY_ORDERED, 2)
    c.set_element_data(1, traj[5][2], previous=True)
    for k in range(6, 12):
        c.upload("displacement", traj[k][0])
        c.internal_force()
        assert np.array_equal(_bits(c.element_data(1)), _bits(traj[k][2])), k
        c.update_states()
    c.close()
    # ATOMIC
    c = _ctx(mesh, capi.ASSEMBLY_ATOMIC, 2)
    for k, (u, f, ed1) in enumerate(traj):
        c.upload("displacement", u)
        c.internal_force()
        fg = c.download("internal_force")
        assert np.abs(fg - f).max() <= 1e-12 * np.abs(f).max(), k
        assert np.array_equal(_bits(c.element_data(1)), _bits(ed1)), k
        c.update_states()
    c.close()


@pytest.mark.parametrize("extra", [(), ("--reference_sequence",), ("--assembly", "atomic")])
def test_driver_state_deck_vs_reference_plumbing(extra, tmp_path):
    """The C++ driver on a deck whose block 1 uses the material with state: Exodus output with the state fields
    (volume-averaged `equivalent_plastic_strain`, `ipt03_equivalent_plastic_strain`, `ipt08_von_mises_stress`) against
    snapshots the reference's own block / element-data / UpdateStates code produced (tests/golden/state_cube.npz)."""
    from nimblesm_b200.exodus_py import read_results, write_genesis

    deck, mesh, _gold, ref, _pieces = load_golden("state_cube")
    base = re.search(r"genesis input file:\s*(\S+)", deck).group(1)
    write_genesis(str(tmp_path / base), mesh)
    (tmp_path / "case.in").write_text(deck)
    r = subprocess.run([EXE, "--quiet", *extra, "case.in"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = read_results(str(tmp_path / "state_cube.out.e"))
    assert np.array_equal(res["times"], ref["times"])
    exact = "atomic" not in extra

    def same(got, want, what, scale=None):
        if exact:
            assert np.array_equal(_bits(got), _bits(want)), what
        else:
            assert np.abs(got - want).max() <= 1e-9 * max(scale if scale else np.abs(want).max(), 1e-300), what

    for lbl in ("displacement", "velocity", "internal_force"):
        for i, cmp_ in enumerate("xyz"):
            same(res["nod"]["%s_%s" % (lbl, cmp_)], ref["node_" + lbl][:, :, i], lbl, np.abs(ref["node_" + lbl]).max())
    rec1 = ref["elem_1"]  # [times][ne][8][17]
    bi = {b: i for i, b in enumerate(mesh["all_block_ids"])}
    smax = np.abs(rec1[..., 9:15]).max()
    same(res["elem"][("ipt03_equivalent_plastic_strain", bi[1])], rec1[:, :, 2, 15], "ipt03 eqps")
    same(res["elem"][("ipt08_von_mises_stress", bi[1])], rec1[:, :, 7, 16], "ipt08 von Mises", smax)
    same(res["elem"][("ipt01_stress_xx", bi[1])], rec1[:, :, 0, 9], "ipt01 stress", smax)
    same(res["elem"][("equivalent_plastic_strain", bi[1])], ref["derived_1_equivalent_plastic_strain"], "averaged eqps")
    same(res["elem"][("stress_xx", bi[1])], ref["derived_1_stress_xx"], "averaged stress", smax)
    same(res["elem"][("volume", bi[1])], ref["derived_1_volume"], "volume")
    same(res["elem"][("stress_xx", bi[2])], ref["derived_2_stress_xx"], "block 2 stress", smax)
    # the state fields exist on the block that carries them only (src/nimble_model_data.cc:284-306)
    assert ("equivalent_plastic_strain", bi[2]) not in res["elem"] or np.all(res["elem"][("equivalent_plastic_strain", bi[2])] == 0)
    assert rec1[-1, :, :, 15].max() > 1e-4
