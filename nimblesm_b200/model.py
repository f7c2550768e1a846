"""nimblesm_b200/model.py — Python host mirror of the reference's driver surface for the hex8 explicit path.

`ExplicitModel` plays the roles of DataManager + ModelData + ExplicitTimeIntegrator for one rank
(src/nimble_data_manager.cc:69-160, src/nimble_model_data.cc:416-667, src/integrators/
explicit_time_integrator.cc:60-330) on top of the C ABI (nimblesm_b200/capi.py): all arithmetic runs in the
CUDA kernels; this file only sequences calls, evaluates boundary-condition magnitudes and keeps the
output-step bookkeeping.  The C++ twin for drop-in use is nimblesm_b200/host.
"""
from __future__ import annotations

import math

import numpy as np

from . import capi
from .deck import Deck, parse_deck

IPT_F_LABELS = ["xx", "yy", "zz", "xy", "yz", "zx", "yx", "zy", "xz"]
IPT_S_LABELS = ["xx", "yy", "zz", "xy", "yz", "zx"]

_EXPR_FUNCS = {k: getattr(math, k) for k in ("sin", "cos", "tan", "exp", "log", "sqrt", "fabs", "atan", "asin", "acos")}
_EXPR_FUNCS["abs"] = abs
_EXPR_FUNCS["pow"] = pow


def eval_expression(expr: str, x, y, z, t):
    """Boundary-condition magnitude expression in x, y, z, t (src/nimble_expression_parser.h); evaluated on the
    host per node exactly where the reference evaluates it (src/nimble_boundary_condition_manager.h:104-113)."""
    code = compile(expr.replace("^", "**"), "<bc>", "eval")
    out = np.empty(len(x))
    for i in range(len(x)):
        out[i] = eval(code, {"__builtins__": {}}, dict(_EXPR_FUNCS, x=float(x[i]), y=float(y[i]), z=float(z[i]), t=t))
    return out


class ExplicitModel:
    def __init__(self, deck, mesh, device=0, assembly=capi.ASSEMBLY_ATOMIC, flags=0):
        self.deck: Deck = parse_deck(deck) if isinstance(deck, str) else deck
        if self.deck.time_integration_scheme != "explicit":
            raise ValueError("only the explicit scheme is on the B200 path")
        self.mesh = mesh
        self.ctx = capi.Context(device)
        c = self.ctx
        c.set_nodes(mesh["x"], mesh["y"], mesh["z"])
        for b in mesh["block_ids"]:
            if b not in self.deck.blocks:
                raise ValueError("block %d has no 'element block' line in the deck" % b)
            m = self.deck.block_material(b)
            c.add_block(b, mesh["conn"][b], m.model, m.bulk_modulus, m.shear_modulus, m.density)
        c.finalize(assembly, flags)
        self.n_nodes = len(mesh["x"])
        self.time = self.time_prev = self.deck.initial_time
        self.step_index = 0
        self.dt_user = 0.0
        self._build_bcs()
        self.snapshots = []
        self.keep_snapshots = False

    # ---- boundary conditions --------------------------------------------------------------------
    def _build_bcs(self):
        node, comp, kind, self._bc_src = [], [], [], []
        for bc in self.deck.boundary_conditions:
            if bc.kind == "initial_velocity":
                continue
            ns = self.mesh["node_sets"].get(bc.node_set_id)
            if ns is None:  # set absent on this rank: BC skipped (src/nimble_boundary_condition_manager.cc:72-76)
                continue
            node.append(np.asarray(ns, dtype=np.int32))
            comp.append(np.full(len(ns), bc.coordinate, dtype=np.int32))
            k = capi.BC_PRESCRIBED_VELOCITY if bc.kind == "prescribed_velocity" else capi.BC_PRESCRIBED_DISPLACEMENT
            kind.append(np.full(len(ns), k, dtype=np.int32))
            self._bc_src.append((bc, np.asarray(ns)))
        self.n_bc = int(sum(len(a) for a in node))
        self._bc_time_dependent = any(bc.expression and "t" in bc.expression.replace("sqrt", "").replace("tan", "")
                                      for bc, _ in self._bc_src)
        if self.n_bc:
            self.ctx.set_bc_table(np.concatenate(node), np.concatenate(comp), np.concatenate(kind))
            self._set_bc_values(self.time)

    def _set_bc_values(self, t):
        vals = []
        for bc, ns in self._bc_src:
            if bc.expression:
                m = self.mesh
                vals.append(eval_expression(bc.expression, m["x"][ns], m["y"][ns], m["z"][ns], t))
            else:
                vals.append(np.full(len(ns), bc.magnitude))
        self.ctx.set_bc_values(np.concatenate(vals))

    def _apply_initial_conditions(self):
        v = np.zeros((self.n_nodes, 3))
        touched = False
        for bc in self.deck.boundary_conditions:
            if bc.kind != "initial_velocity":
                continue
            ns = self.mesh["node_sets"].get(bc.node_set_id)
            if ns is None:
                continue
            m = self.mesh
            v[ns, bc.coordinate] = (eval_expression(bc.expression, m["x"][ns], m["y"][ns], m["z"][ns], 0.0)
                                    if bc.expression else bc.magnitude)
            touched = True
        if touched:
            self.ctx.upload("velocity", v)

    # ---- integrator -------------------------------------------------------------------------------
    def begin(self, keep_snapshots=False) -> float:
        """Pre-loop part of ExplicitTimeIntegrator::Integrate (:123-149); returns the critical time step."""
        d = self.deck
        self.critical_dt = self.ctx.compute_lumped_mass()
        self.time = self.time_prev = d.initial_time
        self.dt_user = (d.final_time - d.initial_time) / d.num_load_steps if d.num_load_steps else 0.0
        self.step_index = 0
        self._apply_initial_conditions()
        if self.n_bc:
            self._set_bc_values(0.0)
            self.ctx.apply_kinematic_bc(0.0, 0.0)
        self.keep_snapshots = keep_snapshots
        if keep_snapshots:
            self.snapshots.append(self.snapshot())
        return self.critical_dt

    def _is_output_step(self, step):
        f = self.deck.output_frequency
        return f != 0 and (step % f == 0 or step == self.deck.num_load_steps - 1)

    def advance(self, n: int) -> float:
        """n passes of the loop body (:177-278).  Steps between output steps are issued as one device call."""
        done = 0
        while done < n:
            # run = consecutive steps up to and including the next output step (one step at a time when a BC
            # magnitude depends on t: the host re-evaluates it like the reference does every step)
            run, out = 0, False
            while done + run < n:
                is_out = self._is_output_step(self.step_index + run)
                run += 1
                if is_out:
                    out = True
                    break
                if self._bc_time_dependent:
                    break
            if self._bc_time_dependent:
                self._set_bc_values(self.time + self.dt_user)
            t_before = self.time
            self.time = self.ctx.step(run, self.time, self.dt_user, store_ipt_last=out and self.keep_snapshots)
            tp = t_before  # time_previous of the last step, accumulated like the loop does
            for _ in range(run - 1):
                tp += self.dt_user
            self.time_prev = tp
            self.step_index += run
            done += run
            if out and self.keep_snapshots:
                self.snapshots.append(self.snapshot())
        return self.time

    # ---- data -------------------------------------------------------------------------------------
    def field(self, label):
        return self.ctx.download(label)

    def snapshot(self):
        s = {"time": self.time, "node": {}, "elem": {}, "derived": {}}
        for lbl in ("lumped_mass", "reference_coordinate", "displacement", "velocity", "acceleration",
                    "internal_force", "external_force"):
            s["node"][lbl] = self.ctx.download(lbl)
        for b in self.mesh["block_ids"]:
            s["elem"][b] = self.ctx.element_data(b)
            d = self.ctx.derived_element_data(b)
            lab = {"volume": d[0]}
            for i, c in enumerate(IPT_F_LABELS):
                lab["deformation_gradient_" + c] = d[1 + i]
            for i, c in enumerate(IPT_S_LABELS):
                lab["stress_" + c] = d[10 + i]
            s["derived"][b] = lab
        return s

    def close(self):
        self.ctx.close()


def exodus_variables(snaps, mesh):
    """Snapshots -> Exodus variable dictionaries named as ModelData::SpecifyOutputFields / ExodusOutput write them
    (src/nimble_model_data.cc:215-368, src/nimble_exodus_output.cc:259-272): nodal `<field>_{x,y,z}` and
    `lumped_mass`; element `iptNN_deformation_gradient_<c>`, `iptNN_stress_<c>` and the volume-averaged
    `deformation_gradient_<c>`, `stress_<c>`, `volume`, keyed (name, 0-based block index)."""
    out = {"times": np.array([s["time"] for s in snaps]), "nod": {}, "elem": {}}
    for lbl in snaps[0]["node"]:
        a = np.stack([s["node"][lbl] for s in snaps])
        if a.ndim == 2:
            out["nod"][lbl] = a
        else:
            for i, c in enumerate("xyz"):
                out["nod"]["%s_%s" % (lbl, c)] = a[:, :, i]
    for bi, b in enumerate(mesh["all_block_ids"] if "all_block_ids" in mesh else mesh["block_ids"]):
        if b not in snaps[0]["elem"]:
            continue
        ed = np.stack([s["elem"][b] for s in snaps])  # [T, ne, 8, 15]
        for q in range(8):
            for i, c in enumerate(IPT_F_LABELS):
                out["elem"][("ipt%02d_deformation_gradient_%s" % (q + 1, c), bi)] = ed[:, :, q, i]
            for i, c in enumerate(IPT_S_LABELS):
                out["elem"][("ipt%02d_stress_%s" % (q + 1, c), bi)] = ed[:, :, q, 9 + i]
        for lab in snaps[0]["derived"][b]:
            out["elem"][(lab, bi)] = np.stack([s["derived"][b][lab] for s in snaps])
    return out
