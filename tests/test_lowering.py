"""CPU check of the whole lowering (path -> kernel plans -> arena) by emulating every
node's address algebra with numpy and comparing with the oracle contraction."""
import numpy as np
import pytest

import emu
from oracle import contract as ocontract
from oracle import network as onet
from oracle import statevector as osv
from qibotn_b200 import circuits, executor
from qibotn_b200.network import QiboCircuitToEinsum
from qibotn_b200.planner import find_path


def emulate(low, leaves, slice_ids):
    dt = np.complex64 if low.dtype == "complex64" else np.complex128
    bufs = {}
    for i, arr in enumerate(leaves):
        bufs[i] = np.ascontiguousarray(arr, dtype=dt).reshape(-1)
    root = low.tensors[low.root]
    out = np.zeros(1 << len(root.labels), dtype=dt)
    esz = np.dtype(dt).itemsize
    # intermediates live at their planned offsets inside shared arenas, filled with NaN,
    # so that aliasing or an undersized arena shows up as a wrong result
    arenas = {"hoist": np.full(low.hoist_bytes // esz + 1, np.nan, dtype=dt),
              "slice": np.full(low.slice_bytes // esz + 1, np.nan, dtype=dt)}
    def run(depends, sid):
        for node in low.nodes:
            if node.depends != depends:
                continue
            tc = low.tensors[node.c]
            if tc.where == "out":
                c = out
            else:
                n = 1 << len(tc.labels)
                assert tc.offset % esz == 0 and tc.offset + n * esz <= getattr(low, tc.where + "_bytes")
                c = arenas[tc.where][tc.offset // esz: tc.offset // esz + n]
                c[:] = 0
            emu.run_plan(node.plan, bufs[node.a], bufs[node.b], c, sid)
            bufs[node.c] = c
    run(False, 0)
    if any(n.depends for n in low.nodes):
        for s in slice_ids:
            run(True, s)
    return out.reshape((2,) * len(root.labels))


@pytest.mark.parametrize("n,layers,min_slices", [(4, 1, 1), (5, 2, 1), (5, 2, 4), (6, 2, 8)])
def test_dense_vector_lowering_matches_statevector(n, layers, min_slices):
    circ = circuits.ry_rz_cnot_ring(n, layers, seed=7)
    conv = QiboCircuitToEinsum(circ, "complex128")
    net = conv.state_vector_network()
    info = find_path(net.modes, net.out, samples=2, seed=1, min_slices=min_slices)
    assert info.num_slices >= min_slices
    low = executor.lower(net.modes, net.out, info, "complex128")
    got = emulate(low, net.tensors, range(info.num_slices)).reshape(-1)
    want = osv.simulate(circ)
    assert np.allclose(got, want, atol=1e-12)


@pytest.mark.parametrize("n,min_slices", [(2, 1), (5, 1), (10, 1), (6, 4)])
def test_dense_initial_state_enters_as_one_leaf(n, min_slices):
    """BASELINE config 1 (reference tests/test_quimb_backend.py:23-66): QFT with swaps on a random normalised
    initial state; the state is one rank-n leaf of the network (reference eval_qu.py:34-41)."""
    rng = np.random.default_rng(n)
    init = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    init /= np.linalg.norm(init)
    circ = circuits.qft(n, True)
    net = QiboCircuitToEinsum(circ, "complex128").state_vector_network(init)
    assert net.modes[0] == list(range(n)) and len(net.out) == n
    info = find_path(net.modes, net.out, samples=2, seed=1, min_slices=min_slices)
    low = executor.lower(net.modes, net.out, info, "complex128")
    got = emulate(low, net.tensors, range(info.num_slices)).reshape(-1)
    want = osv.simulate(circ, initial_state=init)
    assert np.allclose(got, want, atol=1e-12)
    with pytest.raises(ValueError):
        QiboCircuitToEinsum(circ, "complex128").state_vector_network(init[:-1])


@pytest.mark.parametrize("min_slices", [1, 4])
def test_amplitude_and_expectation_lowering(min_slices):
    circ = circuits.brickwork(6, 3, seed=3)
    conv = QiboCircuitToEinsum(circ, "complex128")
    psi = osv.simulate(circ)
    net = conv.amplitude_network("010110")
    info = find_path(net.modes, net.out, samples=2, min_slices=min_slices)
    low = executor.lower(net.modes, net.out, info, "complex128")
    got = emulate(low, net.tensors, range(info.num_slices))
    assert np.allclose(got, psi[int("010110", 2)], atol=1e-12)
    # same network through the oracle with the same path and slices
    ops, modes, out = ocontract.split_interleaved(net.interleaved())
    want = ocontract.contract_sliced(net.interleaved(), info.path, info.sliced_labels,
                                     range(info.num_slices))
    assert np.allclose(got, want, atol=1e-12)
    from qibotn_b200.observables import get_ham_gates
    term = [(0, "X", 0.5), (1, "Z", 1.0), (4, "Y", 1.0)]
    enet = conv.expectation_network(get_ham_gates(term))
    info = find_path(enet.modes, enet.out, samples=2, min_slices=min_slices)
    low = executor.lower(enet.modes, enet.out, info, "complex128")
    got = emulate(low, enet.tensors, range(info.num_slices))
    want = osv.pauli_expectation(psi, [(0.5, [("X", 0), ("Z", 1), ("Y", 4)])], 6)
    assert np.allclose(got, want, atol=1e-12)


@pytest.mark.parametrize("min_slices", [1, 4, 16])
def test_hyper_indexed_diagonal_gates_lower_to_the_same_value(min_slices):
    """QAOA expectation networks with the diagonal gates (RZZ) and the ZZ observable emitted as
    hyper-indexed diagonals: planned (variable elimination is among the candidates), sliced -- hyper
    labels included -- lowered and emulated; equal to the dense simulator and to the reference-style
    network with full rank-4 gate tensors."""
    from qibotn_b200.observables import check_observable, extract_gates_and_qubits, get_ham_gates
    n = 10
    circ, edges = circuits.qaoa_maxcut(n, 2, seed=1)
    psi = osv.simulate(circ)
    conv = QiboCircuitToEinsum(circ, "complex128")
    ham = check_observable(circuits.zz_terms(edges), n)
    for term in list(extract_gates_and_qubits(ham))[:3]:
        hg = get_ham_gates(term, dtype="complex128")
        want = osv.pauli_expectation(psi, [(1.0, [(l, q) for q, l, _ in term])], n) * np.prod([c for _, _, c in term])
        plain = conv.expectation_network(hg)
        net = conv.expectation_network(hg, hyper=True)
        assert len(net.tensors) == len(plain.tensors)
        assert sum(len(m) for m in net.modes) < sum(len(m) for m in plain.modes)
        info = find_path(net.modes, net.out, samples=2, seed=0, min_slices=min_slices)
        assert info.num_slices >= min_slices
        low = executor.lower(net.modes, net.out, info, "complex128")
        got = emulate(low, net.tensors, range(info.num_slices)).reshape(-1)[0]
        assert abs(got - want) < 1e-12
        ref = ocontract.contract(plain.interleaved())
        assert abs(np.asarray(ref).reshape(-1)[0] - want) < 1e-12

