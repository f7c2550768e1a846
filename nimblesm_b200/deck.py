"""nimblesm_b200/deck.py — NimbleSM input-deck surface (Python mirror used by tests and bench.py).

Follows nimble::Parser (src/nimble_parser.cc:196-360): ``key: value`` lines, ``#`` comments, the same keys,
defaults (:168-190) and error behaviour (unknown key -> error).  Material strings are parsed like
MaterialFactoryBase::ParseMaterialParametersString (src/nimble_material_factory_base.cc:62-98) and boundary
conditions like BoundaryCondition::Initialize (src/nimble_boundary_condition.cc:65-160).  The C++ twin lives in
nimblesm_b200/host/parser.{h,cc}.
"""
from __future__ import annotations

from dataclasses import dataclass, field

KNOWN_KEYS = {
    "genesis input file", "exodus output file", "use two level mesh decomposition", "write timing data file",
    "time integration scheme", "nonlinear solver relative tolerance", "nonlinear solver maximum iterations",
    "initial time", "final time", "number of load steps", "output frequency", "contact", "contact backend",
    "contact visualization", "material parameters", "element block", "boundary condition", "output fields",
    "contact dicing", "contact splitting",
}
MATERIAL_MODELS = ("neohookean", "elastic")
MATERIAL_KEYS = ("bulk_modulus", "shear_modulus", "density")


@dataclass
class Material:
    model: str
    density: float
    bulk_modulus: float
    shear_modulus: float


@dataclass
class BoundaryCondition:
    kind: str  # initial_velocity | prescribed_velocity | prescribed_displacement
    node_set_name: str
    node_set_id: int
    coordinate: int
    magnitude: float = 0.0
    expression: str | None = None


@dataclass
class Deck:
    genesis_file: str = "none"
    exodus_file: str = "none"
    time_integration_scheme: str = "explicit"
    initial_time: float = 0.0
    final_time: float = 0.0
    num_load_steps: int = 0
    output_frequency: int = 1
    output_fields: str = ""
    materials: dict = field(default_factory=dict)  # key -> Material
    blocks: dict = field(default_factory=dict)  # block id -> material key
    boundary_conditions: list = field(default_factory=list)
    contact_string: str = ""  # `contact:` line (src/nimble_parser.cc: contact_string_); empty = no contact

    def block_material(self, block_id: int) -> Material:
        return self.materials[self.blocks[block_id]]


def parse_material(props: str) -> Material:
    tok = props.split()
    if not tok:
        raise ValueError("empty material string")
    model = tok[0]
    if model not in MATERIAL_MODELS:
        raise ValueError("Invalid material model name: " + model)
    vals = {}
    rest = tok[1:]
    if len(rest) % 2:
        raise ValueError("material parameters must be name/value pairs: " + props)
    for k, v in zip(rest[::2], rest[1::2]):
        if k not in MATERIAL_KEYS:
            raise ValueError("Invalid material parameter for %s: %s" % (model, k))
        vals[k] = float(v)
    for k in MATERIAL_KEYS:
        if k not in vals:
            raise ValueError("material parameter %s missing in: %s" % (k, props))
    return Material(model, vals["density"], vals["bulk_modulus"], vals["shear_modulus"])


def parse_boundary_condition(s: str) -> BoundaryCondition:
    tok = s.split()
    kind = tok[0]
    if kind not in ("initial_velocity", "prescribed_velocity", "prescribed_displacement", "prescribed_traction"):
        raise ValueError("Error processing boundary condition, unknown boundary condition type: " + kind)
    if kind == "prescribed_traction":
        raise ValueError("prescribed_traction is outside the hex8 explicit path (side sets)")
    name = tok[1]
    coord = tok[2].lower()
    if coord not in "xyz" or len(coord) != 1:
        raise ValueError("Error processing boundary condition, unknown coordinate: " + coord)
    nq = s.count('"')
    bc = BoundaryCondition(kind, name, int(name.rsplit("_", 1)[1]) if "_" in name else -1, "xyz".index(coord))
    if nq == 2:
        bc.expression = s[s.find('"') + 1:s.rfind('"')]
    elif nq == 0:
        bc.magnitude = float(tok[3])
    else:
        raise ValueError("Error processing boundary condition, illegal number of quotes: " + s)
    return bc


def parse_deck(text: str) -> Deck:
    d = Deck()
    for raw in text.splitlines():
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        if ":" not in line:
            raise ValueError("**** Error in Parser::ReadFile(), line without key: " + raw)
        key, value = line.split(":", 1)
        key, value = " ".join(key.split()), value.strip()
        if key not in KNOWN_KEYS:
            raise ValueError("**** Error in Parser::ReadFile(), unknown key " + key)
        if key == "genesis input file":
            d.genesis_file = value
        elif key == "exodus output file":
            d.exodus_file = value
        elif key == "time integration scheme":
            d.time_integration_scheme = value
        elif key == "initial time":
            d.initial_time = float(value)
        elif key == "final time":
            d.final_time = float(value)
        elif key == "number of load steps":
            d.num_load_steps = int(value)
        elif key == "output frequency":
            d.output_frequency = int(value)
        elif key == "output fields":
            d.output_fields = value
        elif key == "material parameters":
            mk, props = value.split(None, 1)
            d.materials[mk] = parse_material(props)
        elif key == "element block":
            bname, mk = value.split(None, 1)
            d.blocks[int(bname.rsplit("_", 1)[1])] = mk.strip()
        elif key == "boundary condition":
            d.boundary_conditions.append(parse_boundary_condition(value))
        elif key == "contact":
            d.contact_string = value
    return d
