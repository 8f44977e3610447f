"""OpenQASM 2 ingest (the wire format of the reference's CPU paths, eval_qu.py:39-41) against the
independent dense simulator and hand-built circuits."""
import math

import numpy as np
import pytest

from oracle import statevector as osv
from qibotn_b200 import circuits
from qibotn_b200.qasm import from_openqasm2_str

QFT3 = """OPENQASM 2.0;
include "qelib1.inc";
qreg q[3];
creg c[3];
h q[0];
cu1(pi/2) q[1],q[0];   // controlled phase
cu1(pi/4) q[2],q[0];
h q[1];
cu1(pi/2) q[2],q[1];
h q[2];
swap q[0],q[2];
barrier q;
measure q -> c;
"""


def test_qft_text_matches_the_builtin_qft():
    got = osv.simulate(from_openqasm2_str(QFT3))
    want = osv.simulate(circuits.qft(3, True))
    assert np.abs(got - want).max() < 1e-14


def test_gate_set_parameters_and_broadcast():
    text = """OPENQASM 2.0; include "qelib1.inc"; qreg a[2]; qreg b[1];
    h a; rx(0.3) a[0]; ry(-pi/3) a[1]; rz(2*pi/5) b[0]; u3(0.1,0.2,0.3) a[0]; u2(0.4,-0.5) b[0];
    cx a[0],b[0]; cz a[1],a[0]; cy a[0],a[1]; crz(0.7) b[0],a[1]; rzz(0.9) a[0],a[1]; sdg b[0]; t a[1]; sx a[0];"""
    circ = from_openqasm2_str(text)
    assert circ.nqubits == 3
    psi = osv.simulate(circ)
    assert abs(np.vdot(psi, psi) - 1) < 1e-13
    # the same circuit by hand on a dense vector
    g = circuits.gates
    ref = circuits.Circuit(3)
    for q in (0, 1):
        ref.add(g.H(q))
    ref.add(g.RX(0, theta=0.3)); ref.add(g.RY(1, theta=-math.pi / 3)); ref.add(g.RZ(2, theta=2 * math.pi / 5))
    from qibotn_b200.qasm import _u3        # qelib1's u3 (no global phase), not qibo's U3
    ref.add(g.Unitary(_u3(0.1, 0.2, 0.3), 0)); ref.add(g.Unitary(_u3(math.pi / 2, 0.4, -0.5), 2))
    ref.add(g.CNOT(0, 2)); ref.add(g.CZ(1, 0))
    ref.add(g.Unitary(np.array([[0, -1j], [1j, 0]]), 1, controls=(0,), name="cy"))
    ref.add(g.CRZ(2, 1, theta=0.7)); ref.add(g.RZZ(0, 1, theta=0.9)); ref.add(g.SDG(2)); ref.add(g.T(1)); ref.add(g.SX(0))
    assert np.abs(psi - osv.simulate(ref)).max() < 1e-13


def test_unsupported_statements_raise():
    with pytest.raises(NotImplementedError):
        from_openqasm2_str("OPENQASM 2.0; qreg q[1]; gate foo a { h a; } foo q[0];")
    with pytest.raises(NotImplementedError):
        from_openqasm2_str("OPENQASM 2.0; qreg q[3]; ccx q[0],q[1],q[2];")
    with pytest.raises(ValueError):
        from_openqasm2_str("OPENQASM 2.0; h q[0];")


def test_function_calls_in_parameters():
    """Parameters are cut at balanced parentheses: ``rz(cos(pi/4))`` keeps its inner call, a list may mix
    calls and commas, and a malformed expression is a ValueError."""
    import math

    from qibotn_b200 import qasm

    src = 'OPENQASM 2.0;\ninclude "qelib1.inc";\nqreg q[2];\nrz(cos(pi/4)) q[0];\nu3(sin(pi/2), ln(exp(1)), sqrt(4)/2) q[1];\n'
    circ = qasm.from_openqasm2_str(src)
    g0, g1 = circ.queue[0], circ.queue[1]
    want = np.diag([np.exp(-0.5j * math.cos(math.pi / 4)), np.exp(0.5j * math.cos(math.pi / 4))])
    assert np.allclose(g0.matrix(), want)
    assert np.allclose(g1.matrix(), qasm._u3(1.0, 1.0, 1.0))
    with pytest.raises(ValueError):
        qasm.from_openqasm2_str('OPENQASM 2.0;\nqreg q[1];\nrz(cos(pi/4) q[0];\n')


def test_to_qasm_round_trip_of_the_stand_in_circuit():
    """``circuit.to_qasm()`` -> ``from_openqasm2_str`` gives the same state (the reference's CPU paths ship circuits this
    way: tests/test_quimb_backend.py:51, eval_qu.py:39-41); gates without a qelib1 name raise."""
    g = circuits.gates
    if not g.__name__.endswith("qibo_compat"):
        pytest.skip("qibo is installed: its own to_qasm is used")
    c = circuits.Circuit(4)
    c.add([g.H(0), g.X(1), g.Y(2), g.Z(3), g.S(0), g.SDG(1), g.T(2), g.TDG(3), g.SX(0), g.I(1)])
    c.add([g.RX(0, theta=0.3), g.RY(1, theta=-1.1), g.RZ(2, theta=2.2), g.U1(3, theta=0.7)])
    c.add([g.CNOT(0, 2), g.CZ(3, 1), g.SWAP(1, 2), g.CU1(3, 0, theta=math.pi / 8), g.CRZ(1, 3, theta=-0.4),
           g.RZZ(0, 3, theta=0.9), g.RXX(2, 1, theta=1.3)])
    c.add(g.M(0, 2))
    text = c.to_qasm()
    assert "cu1(0.39269908169872414) q[3],q[0];" in text and "creg c[2];" in text and "measure q[2] -> c[1];" in text
    back = from_openqasm2_str(text)
    assert back.nqubits == 4 and len(back.queue) == len(c.queue) + 1            # M(0, 2) comes back as two measures
    assert np.abs(osv.simulate(back) - osv.simulate(c)).max() < 1e-14
    q = circuits.qft(5, True)
    assert np.abs(osv.simulate(from_openqasm2_str(q.to_qasm())) - osv.simulate(q)).max() < 1e-14
    bad = circuits.Circuit(2)
    bad.add(g.fSim(0, 1, theta=0.1, phi=0.2))
    with pytest.raises(NotImplementedError):
        bad.to_qasm()
