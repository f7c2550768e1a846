"""oracle/model.py — TEST INFRASTRUCTURE: the explicit loop of the reference on the CPU oracle (hex8_oracle.c).

Mirrors src/integrators/explicit_time_integrator.cc:123-278 + src/nimble_model_data.cc:495-667 +
src/nimble_boundary_condition_manager.h:93-204 for one rank, calling the plain-C restatement for all arithmetic.
Checked in tests/test_oracle.py against snapshots of the reference's own compiled code (tests/golden ref_*) bit for
bit.  Only tests/ may import this file.
"""
from __future__ import annotations

import numpy as np

from nimblesm_b200.deck import parse_deck
from nimblesm_b200.model import IPT_F_LABELS, IPT_S_LABELS, eval_expression

from . import contact as contact_oracle
from . import hex8


class OracleModel:
    def __init__(self, deck, mesh):
        self.deck = parse_deck(deck) if isinstance(deck, str) else deck
        self.mesh = mesh
        self.ref = np.ascontiguousarray(np.stack([mesh["x"], mesh["y"], mesh["z"]], 1))
        n = len(self.ref)
        self.u, self.v, self.a = (np.zeros((n, 3)) for _ in range(3))
        self.f, self.fext = np.zeros((n, 3)), np.zeros((n, 3))
        self.mass = np.zeros(n)
        self.elem = {b: None for b in mesh["block_ids"]}
        self.snapshots = []
        # penalty contact (explicit_time_integrator.cc:76-92): block names of the `contact:` line -> entities
        self.contact, self.fcontact, self.contact_pairs = None, np.zeros((n, 3)), 0
        if getattr(self.deck, "contact_string", ""):
            prim, sec, penalty = contact_oracle.parse_contact_command(self.deck.contact_string)
            ids = lambda names: [int(nm.rsplit("_", 1)[1]) for nm in names]
            self.contact = contact_oracle.ContactSetup(mesh, ids(prim), ids(sec), penalty)

    def _kind(self, b):
        return hex8.ELASTIC if self.deck.block_material(b).model == "elastic" else hex8.NEOHOOKEAN

    def _apply_bc(self, t, t_prev):
        dt = t - t_prev
        m = self.mesh
        for bc in self.deck.boundary_conditions:
            ns = m["node_sets"].get(bc.node_set_id)
            if ns is None or bc.kind == "initial_velocity":
                continue
            mag = (eval_expression(bc.expression, m["x"][ns], m["y"][ns], m["z"][ns], t) if bc.expression
                   else bc.magnitude)
            if bc.kind == "prescribed_velocity":
                self.v[ns, bc.coordinate] = mag
            elif bc.kind == "prescribed_displacement" and dt > 0.0:
                self.v[ns, bc.coordinate] = (mag - self.u[ns, bc.coordinate]) / dt

    def internal_force(self):
        self.f[:] = 0.0
        for b in sorted(self.mesh["block_ids"]):
            mat = self.deck.block_material(b)
            conn = np.ascontiguousarray(self.mesh["conn"][b], dtype=np.int32)
            ed = np.empty((len(conn), 8, 15))
            hex8.lib().h8o_block_internal_force(self._kind(b), mat.bulk_modulus, mat.shear_modulus, self.ref, self.u,
                                                len(conn), conn, self.f, ed.ctypes.data)
            self.elem[b] = ed

    def begin(self, keep_snapshots=False):
        d = self.deck
        self.mass[:] = 0.0
        crit = np.inf
        for b in sorted(self.mesh["block_ids"]):
            mat = d.block_material(b)
            conn = np.ascontiguousarray(self.mesh["conn"][b], dtype=np.int32)
            hex8.lib().h8o_block_lumped_mass(mat.density, self.ref, len(conn), conn, self.mass)
            crit = min(crit, hex8.critical_dt(mat.bulk_modulus, mat.density, self.ref, self.u, conn))
            ed = np.zeros((len(conn), 8, 15))
            ed[:, :, :3] = 1.0
            self.elem[b] = ed
        self.time = self.time_prev = d.initial_time
        self.dt_user = (d.final_time - d.initial_time) / d.num_load_steps if d.num_load_steps else 0.0
        self.step_index = 0
        m = self.mesh
        for bc in d.boundary_conditions:
            ns = m["node_sets"].get(bc.node_set_id)
            if bc.kind == "initial_velocity" and ns is not None:
                self.v[ns, bc.coordinate] = (eval_expression(bc.expression, m["x"][ns], m["y"][ns], m["z"][ns], 0.0)
                                             if bc.expression else bc.magnitude)
        self._apply_bc(0.0, 0.0)
        self.keep = keep_snapshots
        if keep_snapshots:
            self.snapshots.append(self.snapshot())
        return crit

    def advance(self, n):
        L = hex8.lib()
        d = self.deck
        for _ in range(n):
            step = self.step_index
            out = d.output_frequency != 0 and (step % d.output_frequency == 0 or step == d.num_load_steps - 1)
            self.time_prev = self.time
            self.time += self.dt_user
            dt = self.time - self.time_prev
            hdt = 0.5 * dt
            L.h8o_axpy(self.v.size, hdt, self.a.ravel(), self.v.ravel())
            self._apply_bc(self.time, self.time_prev)
            L.h8o_axpy(self.u.size, dt, self.v.ravel(), self.u.ravel())
            self._apply_bc(self.time, self.time_prev)
            self.fext[:] = 0.0
            self.internal_force()
            if self.contact is not None:  # explicit_time_integrator.cc:232-249
                self.fcontact, self.contact_pairs = self.contact.force(self.u)
                contact_oracle.accel_contact(self.mass, self.f, self.fext, self.fcontact, self.a)
            else:
                L.h8o_accel(len(self.ref), self.mass, self.f, self.fext.ctypes.data, self.a)
            L.h8o_axpy(self.v.size, hdt, self.a.ravel(), self.v.ravel())
            if out:
                self._apply_bc(self.time, self.time_prev)
                if self.keep:
                    self.snapshots.append(self.snapshot())
            self.step_index += 1
        return self.time

    def snapshot(self):
        s = {"time": self.time, "node": {"lumped_mass": self.mass.copy(), "reference_coordinate": self.ref.copy(),
                                         "displacement": self.u.copy(), "velocity": self.v.copy(),
                                         "acceleration": self.a.copy(), "internal_force": self.f.copy(),
                                         "external_force": self.fext.copy(), "contact_force": self.fcontact.copy()},
             "elem": {}, "derived": {}}
        for b in self.mesh["block_ids"]:
            s["elem"][b] = self.elem[b].copy()
            dd = hex8.derived(self.ref, self.u, self.mesh["conn"][b], self.elem[b])
            lab = {"volume": dd[0]}
            for i, c in enumerate(IPT_F_LABELS):
                lab["deformation_gradient_" + c] = dd[1 + i]
            for i, c in enumerate(IPT_S_LABELS):
                lab["stress_" + c] = dd[10 + i]
            s["derived"][b] = lab
        return s
