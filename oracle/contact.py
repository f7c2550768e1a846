"""oracle/contact.py — TEST INFRASTRUCTURE: penalty contact of the reference on the CPU oracle (contact_oracle.c).

Entity creation follows ContactManager::CreateContactEntities / SkinBlocks / CreateContactNodesAndFaces
(src/nimble_contact_manager.cc:184-393, 788-934, 1043-1205) for one rank; the force evaluation is
oracle/contact_oracle.c.  Pinned in tests/test_oracle.py against oracle/ref_contact.cc (the reference's own
ContactEntity objects) bit for bit and against the reference's gold files.  Only tests/ may import this file.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import hex8

# Exodus hex8 face -> local nodes (src/nimble_contact_manager.cc:820-905)
FACE_NODES = np.array([[0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [0, 4, 7, 3], [0, 3, 2, 1], [4, 5, 6, 7]])
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_bound = False


def _lib():
    global _bound
    L = hex8.lib()
    if not _bound:
        L.h8o_contact_char_lengths.argtypes = [_dp, C.c_long, _ip, _dp, C.c_long, _ip, C.c_long, _dp]
        L.h8o_contact_force.restype = C.c_long
        L.h8o_contact_force.argtypes = [C.c_double, C.c_long, _dp, _dp, C.c_long, _ip, _dp, C.c_long, _ip, _dp, _dp, C.c_void_p]
        L.h8o_accel_contact.argtypes = [C.c_long, _dp, _dp, C.c_void_p, _dp, _dp]
        L.h8o_contact_projection.restype = C.c_int
        L.h8o_contact_projection.argtypes = [_dp, _dp, C.c_double, _dp, _dp, _dp]
        _bound = True
    return L


def projection(node, tri, char_len):
    """ContactManager::Projection (src/nimble_contact_manager.cc:1549-1620) of one node on one triangular facet ->
    (in, gap, normal[3], barycentric[3]); the last three are NaN when the projection falls outside the facet (the
    reference leaves its outputs untouched there)."""
    gap, normal, bary = np.full(1, np.nan), np.full(3, np.nan), np.full(3, np.nan)
    inside = _lib().h8o_contact_projection(np.ascontiguousarray(node, dtype=np.float64), np.ascontiguousarray(tri, dtype=np.float64).reshape(-1),
                                          float(char_len), gap, normal, bary)
    return bool(inside), float(gap[0]), normal, bary


def parse_contact_command(command: str):
    """`contact:` deck line -> (primary block names, secondary block names, penalty); ParseContactCommand (:95-149)."""
    tok = command.split()
    if not tok or tok[0] not in ("primary_blocks", "master_blocks"):
        raise ValueError("Error processing contact command, unknown key: " + (tok[0] if tok else ""))
    primary, secondary, penalty, stage = [], [], None, 0
    it = iter(tok[1:])
    for t in it:
        if stage == 0 and t in ("secondary_blocks", "slave_blocks"):
            stage = 1
        elif stage == 1 and t == "penalty_parameter":
            penalty = float(next(it))
            stage = 2
            break
        else:
            (primary if stage == 0 else secondary).append(t)
    if stage != 2:
        raise ValueError('Error processing contact command, expected "secondary_blocks" ... "penalty_parameter"')
    return primary, secondary, penalty


def skin_faces(conns):
    """Faces that occur exactly once in the given blocks' connectivity -> [nf,4] node ids in the owning element's
    Exodus face order, sorted lexicographically by their sorted node lists (the reference's std::map order)."""
    conns = [np.asarray(c).reshape(-1, 8) for c in conns if len(c)]
    if not conns:
        return np.zeros((0, 4), np.int32)
    faces = np.concatenate([c[:, FACE_NODES].reshape(-1, 4) for c in conns])
    key = np.sort(faces, axis=1)
    order = np.lexsort(key.T[::-1])
    key, faces = key[order], faces[order]
    new = np.ones(len(key), bool)
    new[1:] = np.any(key[1:] != key[:-1], axis=1)
    start = np.flatnonzero(new)
    count = np.diff(np.append(start, len(key)))
    if np.any(count > 2):
        raise ValueError("Error in mesh skinning routine, face found more than two times!")
    return np.ascontiguousarray(faces[start[count == 1]], dtype=np.int32)


class ContactSetup:
    """The contact entities of one rank: primary quads (4 triangles each), contact nodes, characteristic lengths."""

    def __init__(self, mesh, primary_block_ids, secondary_block_ids, penalty):
        self.penalty = float(penalty)
        self.ref = np.ascontiguousarray(np.stack([mesh["x"], mesh["y"], mesh["z"]], 1))
        n = len(self.ref)
        self.primary_quads = skin_faces([mesh["conn"][b] for b in primary_block_ids if b in mesh["conn"]])
        self.secondary_quads = skin_faces([mesh["conn"][b] for b in secondary_block_ids if b in mesh["conn"]])
        # contact nodes in order of first appearance over the secondary faces (:288-330)
        flat = self.secondary_quads.ravel()
        _u, first = np.unique(flat, return_index=True)
        self.contact_nodes = np.ascontiguousarray(flat[np.sort(first)], dtype=np.int32)
        self.primary_char_len = np.zeros(len(self.primary_quads))
        node_len = np.zeros(n)
        _lib().h8o_contact_char_lengths(self.ref, len(self.primary_quads), self.primary_quads.reshape(-1), self.primary_char_len,
                                        len(self.secondary_quads), self.secondary_quads.reshape(-1), n, node_len)
        self.contact_node_char_len = np.ascontiguousarray(node_len[self.contact_nodes])

    def entity_vertices(self, coord):
        """The coordinates the contact entities hold for the nodal coordinates `coord` [n,3], in the order of the
        reference's contact visualisation database (src/nimble_contact_manager.cc:515-563): three vertices per
        triangular facet (node k, node k+1, the fictitious node at the mean of the face's four nodes,
        CreateContactNodesAndFaces :1083-1088 and ContactEntity::SetCoordinates), then the contact nodes."""
        c = np.asarray(coord)[self.primary_quads]  # [nf,4,3]
        centre = (((0.0 + c[:, 0]) + c[:, 1]) + c[:, 2] + c[:, 3]) / 4
        tri = np.empty((len(c), 4, 3, 3))
        for k in range(4):
            tri[:, k, 0], tri[:, k, 1], tri[:, k, 2] = c[:, k], c[:, (k + 1) % 4], centre
        return np.vstack([tri.reshape(-1, 3), np.asarray(coord)[self.contact_nodes]])

    def force(self, disp, want_status=False):
        """-> (contact force [n,3], enforced pairs[, status flags of the 4*nf triangles then the contact nodes])"""
        disp = np.ascontiguousarray(disp, dtype=np.float64)
        f = np.zeros_like(disp)
        status = np.zeros(4 * len(self.primary_quads) + len(self.contact_nodes), np.uint8) if want_status else None
        pairs = _lib().h8o_contact_force(self.penalty, len(self.ref), self.ref, disp, len(self.primary_quads),
                                         self.primary_quads.reshape(-1), self.primary_char_len, len(self.contact_nodes),
                                         self.contact_nodes, self.contact_node_char_len, f,
                                         status.ctypes.data if want_status else None)
        return (f, int(pairs), status) if want_status else (f, int(pairs))


def accel_contact(mass, f_int, f_ext, f_contact, a):
    _lib().h8o_accel_contact(len(mass), mass, f_int, f_ext.ctypes.data if f_ext is not None else None, f_contact, a)
