"""nimblesm_b200/exodiff.py — an exodiff-equivalent comparator that reads the reference's ``<case>.exodiff``
command files (test/dynamics/*/*.exodiff) and applies them the way ``exodiff -f`` does
(test/run_exodiff_test.py:180-190).

Supported subset (everything the reference's files use): the section headers ``TIME STEPS``, ``NODAL VARIABLES``,
``ELEMENT VARIABLES`` (``GLOBAL``/``NODESET``/``SIDESET`` are parsed and ignored: NimbleSM writes none) with a default
``relative|absolute <tol> [floor <f>]``, followed by indented per-variable lines ``name [relative|absolute <tol>]
[floor <f>]``.  When a section lists variables only those are compared; a section header without a list compares
all variables with the default.  Difference measures (exodiff manual):
    absolute: |a - b|            relative: |a - b| / max(|a|, |b|)
Values whose magnitudes are both below ``floor`` are skipped.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Tol:
    kind: str = "relative"
    value: float = 1e-6
    floor: float = 0.0


@dataclass
class Section:
    default: Tol = field(default_factory=Tol)
    variables: dict = field(default_factory=dict)  # name -> Tol
    present: bool = False


def _parse_tol(tokens, base: Tol) -> Tol:
    t = Tol(base.kind, base.value, base.floor)
    i = 0
    while i < len(tokens):
        k = tokens[i].lower()
        if k in ("relative", "absolute") and i + 1 < len(tokens):
            t.kind, t.value = k, float(tokens[i + 1])
            i += 2
        elif k == "floor" and i + 1 < len(tokens):
            t.floor = float(tokens[i + 1])
            i += 2
        else:
            i += 1
    return t


def parse(text: str) -> dict:
    """-> {'time': Section, 'nodal': Section, 'element': Section, 'coordinates': Section}"""
    sec = {"time": Section(), "nodal": Section(), "element": Section(), "coordinates": Section(),
           "global": Section(), "nodeset": Section(), "sideset": Section()}
    heads = {"TIME STEPS": "time", "NODAL VARIABLES": "nodal", "ELEMENT VARIABLES": "element",
             "COORDINATES": "coordinates", "GLOBAL VARIABLES": "global", "NODESET VARIABLES": "nodeset",
             "SIDESET VARIABLES": "sideset"}
    cur = None
    for raw in text.splitlines():
        line = raw.split("#", 1)[0].rstrip()
        if not line.strip():
            continue
        head = next((h for h in heads if line.upper().startswith(h)), None)
        if head and not raw[:1].isspace():
            cur = sec[heads[head]]
            cur.present = True
            cur.default = _parse_tol(line[len(head):].split(), Tol())
            continue
        if cur is not None and raw[:1].isspace():
            tok = line.split()
            cur.variables[tok[0]] = _parse_tol(tok[1:], cur.default)
    return sec


def difference(a, b, tol: Tol):
    """Largest violation measure over all entries and whether it passes."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return float("inf"), False
    keep = ~((np.abs(a) < tol.floor) & (np.abs(b) < tol.floor)) if tol.floor > 0 else np.ones(a.shape, bool)
    d = np.abs(a - b)
    if tol.kind == "relative":
        m = np.maximum(np.abs(a), np.abs(b))
        with np.errstate(invalid="ignore", divide="ignore"):
            d = np.where(m > 0, d / m, 0.0)
    d = np.where(keep, d, 0.0)
    worst = float(d.max()) if d.size else 0.0
    return worst, bool(worst <= tol.value)


def compare(spec_text: str, gold: dict, test: dict, rounding_floor: float = 0.0):
    """gold/test: {'times': [T], 'nod': {name: [T, n]}, 'elem': {(name, block_index): [T, ne]}}.
    Returns a list of failure strings (empty == files are the same in exodiff's sense).
    rounding_floor > 0: an ABSOLUTE tolerance is never taken below rounding_floor * max|gold variable| (the contact decks'
    .exodiff files ask for 2e-4 absolute on forces of 3e8, i.e. 15 digits -- less than a different summation order of
    the same node-face pairs moves them, the reference's own atomic scatter included)."""
    spec = parse(spec_text)
    fails = []
    if spec["time"].present or True:
        w, ok = difference(gold["times"], test["times"], spec["time"].default if spec["time"].present else Tol())
        if not ok:
            fails.append("time steps differ: %.3e" % w)
    for kind, key in (("nodal", "nod"), ("element", "elem")):
        s = spec[kind]
        if not s.present:
            continue
        names = sorted({k if kind == "nodal" else k[0] for k in gold[key]})
        for nm in names:
            if s.variables and nm not in s.variables:
                continue
            tol = s.variables.get(nm, s.default)
            entries = [k for k in gold[key] if (k if kind == "nodal" else k[0]) == nm]
            for k in entries:
                if k not in test[key]:
                    fails.append("%s variable %s missing from the test data" % (kind, k))
                    continue
                if rounding_floor > 0 and tol.kind == "absolute":
                    scale = float(np.abs(np.asarray(gold[key][k], dtype=np.float64)).max()) if np.size(gold[key][k]) else 0.0
                    tol = Tol(tol.kind, max(tol.value, rounding_floor * scale), tol.floor)
                w, ok = difference(gold[key][k], test[key][k], tol)
                if not ok:
                    fails.append("%s %s: %s diff %.3e > %.3e" % (kind, k, tol.kind, w, tol.value))
    return fails
