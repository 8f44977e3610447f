"""OpenQASM 2.0 ingest: QASM text -> circuit object the engine can run.

Both CPU paths of the reference go through ``circuit.to_qasm()``
(``/root/reference/src/qibotn/eval_qu.py:39-41``, ``backends/quimb.py:169-171``); accepting the same
text makes the engine usable without qibo objects.  Supported: one or more ``qreg`` (concatenated in
declaration order), the ``qelib1.inc`` gates ``id x y z h s sdg t tdg sx rx ry rz u1 u2 u3 p cx cz
cy swap crz cu1 cp rzz rxx``, ``barrier``, ``measure`` (kept as a measurement marker, which the
converters skip) and ``pi`` arithmetic in parameters.  ``gate`` definitions, ``if`` and ``reset`` raise.
"""

from __future__ import annotations

import ast
import math
import operator
import re

import numpy as np

try:  # pragma: no cover - qibo is not in this image
    from qibo import Circuit, gates
    _unitary = lambda u, *q: gates.Unitary(u, *q)          # noqa: E731
except Exception:
    from . import qibo_compat as gates
    from .qibo_compat import Circuit
    _unitary = lambda u, *q: gates.Unitary(u, *q, name="unitary")      # noqa: E731

__all__ = ["from_openqasm2_str"]

_OPS = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv,
        ast.Pow: operator.pow, ast.USub: operator.neg, ast.UAdd: operator.pos}
_FUNCS = {"sin": math.sin, "cos": math.cos, "tan": math.tan, "exp": math.exp, "ln": math.log, "sqrt": math.sqrt}


def _split_statement(st):
    """``name(params) args`` -> (name, params or None, args); the parameter list is taken by matching
    balanced parentheses, so that ``rz(cos(pi/4)) q[0]`` keeps its inner call."""
    m = re.match(r"\s*(\w+)\s*", st)
    if not m:
        raise ValueError(f"cannot parse statement {st!r}")
    name, rest = m.group(1).lower(), st[m.end():]
    if not rest.startswith("("):
        return name, None, rest.strip()
    depth = 0
    for pos, ch in enumerate(rest):
        depth += ch == "("
        depth -= ch == ")"
        if depth == 0:
            return name, rest[1:pos], rest[pos + 1:].strip()
    raise ValueError(f"unbalanced parentheses in {st!r}")


def _split_top_level(params):
    """Split a parameter list at the commas outside any parentheses."""
    out, depth, cur = [], 0, []
    for ch in params:
        depth += ch == "("
        depth -= ch == ")"
        if ch == "," and depth == 0:
            out.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    out.append("".join(cur))
    return out


def _eval(expr: str) -> float:
    """Arithmetic in gate parameters: numbers, pi, + - * / ^ and the QASM math functions."""
    def ev(node):
        if isinstance(node, ast.Expression):
            return ev(node.body)
        if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
            return float(node.value)
        if isinstance(node, ast.Name) and node.id == "pi":
            return math.pi
        if isinstance(node, ast.BinOp) and type(node.op) in _OPS:
            return _OPS[type(node.op)](ev(node.left), ev(node.right))
        if isinstance(node, ast.UnaryOp) and type(node.op) in _OPS:
            return _OPS[type(node.op)](ev(node.operand))
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id in _FUNCS:
            return _FUNCS[node.func.id](*[ev(a) for a in node.args])
        raise ValueError(f"unsupported expression in QASM parameter: {expr!r}")
    try:
        tree = ast.parse(expr.strip().replace("^", "**"), mode="eval")
    except SyntaxError:
        raise ValueError(f"cannot parse QASM parameter {expr!r}") from None
    return ev(tree)


def _u3(theta, phi, lam):
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    return np.array([[c, -np.exp(1j * lam) * s], [np.exp(1j * phi) * s, np.exp(1j * (phi + lam)) * c]])


def from_openqasm2_str(text: str):
    """Parse OpenQASM 2.0 and return a circuit (qibo's when installed, the stand-in otherwise)."""
    text = re.sub(r"//[^\n]*", "", text)
    stmts = [s.strip() for s in text.split(";") if s.strip()]
    regs, offset, ops = {}, 0, []
    for st in stmts:
        if st.startswith("OPENQASM") or st.startswith("include") or st.startswith("creg"):
            continue
        m = re.fullmatch(r"qreg\s+(\w+)\s*\[\s*(\d+)\s*\]", st)
        if m:
            regs[m.group(1)] = (offset, int(m.group(2)))
            offset += int(m.group(2))
            continue
        if re.match(r"(gate|if|reset|opaque)\b", st):
            raise NotImplementedError(f"OpenQASM statement not supported: {st.split()[0]}")
        ops.append(st)
    if not regs:
        raise ValueError("no qreg declared")
    circ = Circuit(offset)

    def qubits_of(arg):
        m = re.fullmatch(r"(\w+)\s*\[\s*(\d+)\s*\]", arg)
        if m:
            base, size = regs[m.group(1)]
            if int(m.group(2)) >= size:
                raise ValueError(f"qubit index out of range: {arg}")
            return [base + int(m.group(2))]
        if arg in regs:                                   # whole register: broadcast
            base, size = regs[arg]
            return list(range(base, base + size))
        raise ValueError(f"unknown qubit argument {arg!r}")

    one = {"id": gates.I, "x": gates.X, "y": gates.Y, "z": gates.Z, "h": gates.H, "s": gates.S,
           "sdg": gates.SDG, "t": gates.T, "tdg": gates.TDG, "sx": gates.SX}
    rot = {"rx": gates.RX, "ry": gates.RY, "rz": gates.RZ, "u1": gates.U1, "p": gates.U1}
    two = {"cx": gates.CNOT, "cz": gates.CZ, "swap": gates.SWAP}
    rot2 = {"crz": gates.CRZ, "cu1": gates.CU1, "cp": gates.CU1, "rzz": gates.RZZ, "rxx": gates.RXX}
    for st in ops:
        name, params, args = _split_statement(st)
        if name == "barrier":
            continue
        if name == "measure":
            src = args.split("->")[0].strip()
            circ.add(gates.M(*qubits_of(src)))
            continue
        par = [_eval(p) for p in _split_top_level(params)] if params else []
        lists = [qubits_of(a.strip()) for a in args.split(",")]
        width = max(len(q) for q in lists)
        for i in range(width):                            # register broadcast, as qelib semantics
            q = [l[i] if len(l) > 1 else l[0] for l in lists]
            if name in one and len(q) == 1 and not par:
                circ.add(one[name](*q))
            elif name in rot and len(q) == 1 and len(par) == 1:
                circ.add(rot[name](q[0], theta=par[0]))
            elif name == "u2" and len(q) == 1 and len(par) == 2:
                circ.add(_unitary(_u3(math.pi / 2, par[0], par[1]), q[0]))
            elif name in ("u3", "u") and len(q) == 1 and len(par) == 3:
                circ.add(_unitary(_u3(*par), q[0]))
            elif name in two and len(q) == 2 and not par:
                circ.add(two[name](*q))
            elif name == "cy" and len(q) == 2 and not par:
                circ.add(gates.Unitary(np.array([[0, -1j], [1j, 0]]), q[1], controls=(q[0],), name="cy")
                         if gates.__name__.endswith("qibo_compat") else gates.CY(*q))
            elif name in rot2 and len(q) == 2 and len(par) == 1:
                circ.add(rot2[name](q[0], q[1], theta=par[0]))
            else:
                raise NotImplementedError(f"OpenQASM gate not supported: {st}")
    return circ
