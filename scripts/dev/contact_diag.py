"""diagnostic: device contact loop vs oracle on cubes_contact, error per field after k steps"""
import ctypes as C
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.conftest import load_golden
from tests.test_host_cpp import LIB as HOST_LIB, host_contact_entities
from nimblesm_b200 import capi
from nimblesm_b200.deck import parse_deck
from nimblesm_b200.exodus_py import write_genesis
from oracle.model import OracleModel
import tempfile

deck, mesh, *_ = load_golden("cubes_contact")
tmp = tempfile.mkdtemp()
g = os.path.join(tmp, "c.g")
write_genesis(g, mesh)
host = C.CDLL(HOST_LIB)
ent = host_contact_entities(host, g, deck)
d = parse_deck(deck)
tn, tc, tv = [], [], []
v0 = np.zeros((len(mesh["x"]), 3))
for bc in d.boundary_conditions:
    ns = mesh["node_sets"][bc.node_set_id]
    if bc.kind == "initial_velocity":
        v0[ns, bc.coordinate] = bc.magnitude
    else:
        tn.append(ns), tc.append(np.full(len(ns), bc.coordinate, np.int32)), tv.append(np.full(len(ns), bc.magnitude))
tn, tc, tv = np.concatenate(tn).astype(np.int32), np.concatenate(tc), np.concatenate(tv)
for k in range(len(tn)):
    v0[tn[k], tc[k]] = tv[k]
def rel(a, b):
    s = np.abs(b).max()
    return np.abs(a - b).max() / (s if s > 0 else 1.0)
om = OracleModel(deck, mesh)
om.begin()
c = capi.Context(0)
c.set_nodes(mesh["x"], mesh["y"], mesh["z"])
for b in mesh["block_ids"]:
    m = d.block_material(b)
    c.add_block(b, mesh["conn"][b], m.model, m.bulk_modulus, m.shear_modulus, m.density)
c.finalize(capi.ASSEMBLY_ORDERED, 2)
c.compute_lumped_mass()
print("mass", rel(c.download("lumped_mass"), om.mass), "v0", rel(v0, om.v))
c.set_contact(ent["penalty"], ent["primary_quads"], ent["primary_char_len"], ent["contact_nodes"], ent["contact_node_char_len"])
c.set_bc_table(tn, tc, np.zeros(len(tn), np.int32))
c.set_bc_values(tv)
c.upload("velocity", v0)
dt = (d.final_time - d.initial_time) / d.num_load_steps
t = 0.0
for k in range(100):
    t = c.step(1, t, dt)
    om.advance(1)
    u = c.download("displacement")
    fc_dev = c.download("contact_force")
    fc_or_on_dev_u, pairs = om.contact.force(u)
    if k % 10 != 9:
        continue
    fi = c.download("internal_force")
    f_same = np.zeros_like(fi)
    from oracle import hex8
    for b in sorted(mesh["block_ids"]):
        m_ = d.block_material(b)
        fb, _ = hex8.internal_force(hex8.NEOHOOKEAN, m_.bulk_modulus, m_.shear_modulus, om.ref, u, mesh["conn"][b], False)
        f_same += fb
    print("   |f| max %.3e  abs err vs oracle loop %.3e  vs oracle on device u %.3e ; |fc| %.3e abs err %.3e ; |u| %.3e abs err %.3e" % (
        np.abs(om.f).max(), np.abs(fi - om.f).max(), np.abs(fi - f_same).max(), np.abs(om.fcontact).max(), np.abs(fc_dev - om.fcontact).max(),
        np.abs(om.u).max(), np.abs(u - om.u).max()))
    print(k, "t", t == om.time, " ".join("%s %.2e" % (l[:4], rel(c.download(l), w)) for l, w in (("displacement", om.u), ("velocity", om.v), ("acceleration", om.a), ("internal_force", om.f), ("contact_force", om.fcontact))),
          "| fc(dev u): dev vs oracle %.2e" % rel(fc_dev, fc_or_on_dev_u), "pairs", c.contact_stats()["pairs"], pairs, om.contact_pairs,
          "bitdiff u", int((u.view(np.int64) != om.u.view(np.int64)).sum()))
