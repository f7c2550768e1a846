"""tests/bc_program.py — test helper: compile an expression with the C++ host layer (Expression::compile through
lib/libnsm_host_c.so) and interpret the device program (include/nsm_b200.h, nsm_bc_op) with plain IEEE doubles.
Only the tests use it; the product evaluates programs in bc_program_kernel."""
import ctypes as C
import math

import numpy as np

(CONST, X, Y, Z, SLOT, ADD, SUB, MUL, DIV, FMOD, NEG, SQRT, ABS, FLOOR, CEIL, ROUND, LT, LE, GT, GE, EQ, AND, OR, XOR, NOT,
 SELECT, ENTRYCONST) = range(27)

EXTRA_EXPRESSIONS = ["x*cos(t*3.0e5)", "cos(t*3.141592653589793/2.0e-4)*x + y/3", "sqrt(x*x+y*y)*exp(-0.2*t) - abs(z)*t",
                     "x % 0.3 + t", "t>1.0e-6 ? 10*x : -y", "(x+1)*(y+2)/(z+3)*log(t+2)", "floor(10*x)+ceil(y)+round(z)+t",
                     "0.01*t", "x<0.5 ? -t*y : 2*y+t", "1000.0*x*t"]
# position-only sub-trees through libm / pow: compiled, the host supplies one value per BC entry (NSM_BCOP_ENTRYCONST)
ENTRY_CONSTANT_EXPRESSIONS = ["x^2*t", "exp(x)*t", "cbrt(x)+t", "e^x * t", "sin(3*x)*cos(t)", "exp(-y)*(1+t)",
                              "sin(3*x)*cos(t) + x*sin(3*x)*t", "log(z+2)^2*t - y"]
# libm / pow of a MIX of position and time has no bit-exact device form: the host evaluates these per node per step
NOT_COMPILABLE = ["sin(x*t)", "(x*t)^2", "exp(x+t)", "cos(t+y)*x"]


def compile_expression(host, text, t):
    """-> (code int32[], consts float64[], slot values at time t float64[]) or None when it has no device form."""
    host.nsmh_expression_compile.argtypes = [C.c_char_p, C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int,
                                             C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int),
                                             C.POINTER(C.c_double), C.c_char_p, C.c_int]
    code, consts, slots = (C.c_int * 256)(), (C.c_double * 64)(), (C.c_double * 64)()
    nw, nc, ns = C.c_int(), C.c_int(), C.c_int()
    err = C.create_string_buffer(512)
    rc = host.nsmh_expression_compile(text.encode(), t, 256, C.byref(nw), code, 64, C.byref(nc), consts, 64, C.byref(ns), slots,
                                      err, 512)
    if rc == 2:
        return None
    assert rc == 0, err.value
    return (np.array(code[:nw.value], dtype=np.int32), np.array(consts[:nc.value]), np.array(slots[:ns.value]))


def entry_constants(host, text, x, y, z):
    """the per-entry constants of compile_expression(text) evaluated at one point (host, glibc)"""
    host.nsmh_expression_entry_constants.argtypes = [C.c_char_p, C.c_double, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_int),
                                                     C.POINTER(C.c_double), C.c_char_p, C.c_int]
    vals, n = (C.c_double * 64)(), C.c_int()
    err = C.create_string_buffer(512)
    rc = host.nsmh_expression_entry_constants(text.encode(), x, y, z, 64, C.byref(n), vals, err, 512)
    assert rc == 0, (rc, err.value)
    return np.array(vals[:n.value])


def interpret(code, consts, slots, x, y, z, entry=()):
    st = []
    for w in code:
        op, arg = int(w) & 0xff, int(w) >> 8
        if op == CONST:
            st.append(float(consts[arg]))
        elif op in (X, Y, Z):
            st.append(float((x, y, z)[op - X]))
        elif op == SLOT:
            st.append(float(slots[arg]))
        elif op == ENTRYCONST:
            st.append(float(entry[arg]))
        elif op in (NEG, SQRT, ABS, FLOOR, CEIL, ROUND, NOT):
            a = st.pop()
            st.append({NEG: lambda: -a, SQRT: lambda: math.sqrt(a) if a >= 0 else float("nan"), ABS: lambda: abs(a),
                       FLOOR: lambda: float(math.floor(a)), CEIL: lambda: float(math.ceil(a)),
                       ROUND: lambda: float(math.copysign(math.floor(abs(a) + 0.5), a)),
                       NOT: lambda: 0.0 if a != 0.0 else 1.0}[op]())
        elif op == SELECT:
            c, b, a = st.pop(), st.pop(), st.pop()
            st.append(b if a != 0.0 else c)
        else:
            b, a = st.pop(), st.pop()
            st.append({ADD: lambda: a + b, SUB: lambda: a - b, MUL: lambda: a * b,
                       DIV: lambda: float(np.float64(a) / np.float64(b)), FMOD: lambda: math.fmod(a, b),
                       LT: lambda: float(a < b), LE: lambda: float(a <= b), GT: lambda: float(a > b), GE: lambda: float(a >= b),
                       EQ: lambda: float(a == b), AND: lambda: float((a != 0.0) and (b != 0.0)),
                       OR: lambda: float((a != 0.0) or (b != 0.0)), XOR: lambda: float((a != 0.0) != (b != 0.0))}[op]())
    assert len(st) == 1
    return st[0]


def host_eval(host, text, x, y, z, t):
    host.nsmh_expression_eval.argtypes = [C.c_char_p] + [C.c_double] * 4 + [C.POINTER(C.c_double), C.c_char_p, C.c_int]
    a = C.c_double()
    err = C.create_string_buffer(512)
    assert host.nsmh_expression_eval(text.encode(), x, y, z, t, C.byref(a), err, 512) == 0, err.value
    return a.value
