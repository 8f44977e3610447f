"""GPU parity of the tensor-network path through the backend surface, against the
closed forms the reference's own tests pin and against the CPU oracle."""
import math

import numpy as np
import pytest

from oracle import contract as ocontract
from oracle import statevector as osv
from qibotn_b200 import MetaBackend, circuits, engine
from qibotn_b200 import eval as ev
from qibotn_b200.network import QiboCircuitToEinsum

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

# BASELINE.json north_star: results within 1e-5 (complex64) / 1e-10 (complex128) RELATIVE of the CPU path.
# The norm each comparison uses, with no extra multipliers:
#   state vectors      ||got - want||_2 / ||want||_2                      (rel2)
#   single amplitudes  |got - want| / |want|
#   expectation values |got - want| / sum_i |c_i|  -- the operator-norm bound of H = sum_i c_i P_i, which is
#                      what the rounding of a sandwich <psi|H|psi> scales with (<H> itself may be ~0)
TOL = {"complex64": 1e-5, "complex128": 1e-10}


def rel2(got, want):
    got, want = np.asarray(got).reshape(-1), np.asarray(want).reshape(-1)
    return float(np.linalg.norm(got - want) / np.linalg.norm(want))


def _backend(dtype):
    b = MetaBackend.load("cutensornet")
    b.set_precision("single" if dtype == "complex64" else "double")
    return b


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("nqubits", [1, 2, 5, 10])
def test_eval_qft_uniform(nqubits, dtype):
    """reference tests/test_cuquantum_cutensor_backend.py:46-85: QFT|0..0> = uniform vector."""
    circ = circuits.qft(nqubits, True)
    b = _backend(dtype)
    got = b.execute_circuit(circuit=circ).statevector.flatten().get()
    want = np.full(1 << nqubits, 2.0 ** (-nqubits / 2))
    assert np.allclose(got, want, rtol=1e-5, atol=1e-8 if dtype == "complex128" else 1e-6)
    b.configure_tn_simulation({"MPI_enabled": False, "MPS_enabled": False, "NCCL_enabled": False,
                               "expectation_enabled": False})
    got = b.execute_circuit(circuit=circ, return_array=True).statevector.get()
    assert np.allclose(got, want, rtol=1e-5, atol=1e-8 if dtype == "complex128" else 1e-6)


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("nqubits,layers", [(2, 2), (5, 2), (7, 2), (12, 3), (16, 4), (20, 3)])
def test_random_circuit_state_vs_oracle(nqubits, layers, dtype):
    circ = circuits.ry_rz_cnot_ring(nqubits, layers, seed=42)
    want = osv.simulate(circ)
    got = ev.dense_vector_tn(circ, dtype).flatten().get()
    assert rel2(got, want) <= TOL[dtype]


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("nqubits", [2, 5, 10])
def test_expectation_three_runcard_forms(nqubits, dtype):
    """reference tests/test_cuquantum_cutensor_backend.py:145-199 (value is 0 on QFT|0>) and a
    discriminating random circuit (value is not 0)."""
    from qibotn_b200.observables import build_observable
    terms = {"terms": [{"coefficient": 0.5, "operators": [("X", i), ("Z", (i + 1) % nqubits)]}
                       for i in range(nqubits)]}
    for circ in (circuits.qft(nqubits, True), circuits.ry_rz_cnot_ring(nqubits, 2, seed=42)):
        psi = osv.simulate(circ)
        want = osv.pauli_expectation(
            psi, [(0.5, [("X", i), ("Z", (i + 1) % nqubits)]) for i in range(nqubits)], nqubits).real
        b = _backend(dtype)
        for value in (True, build_observable(nqubits), terms):
            b.configure_tn_simulation({"MPI_enabled": False, "MPS_enabled": False,
                                       "NCCL_enabled": False, "expectation_enabled": value})
            res = b.execute_circuit(circuit=circ)
            # the reference's own bound (abs 1e-7, tests/test_cuquantum_cutensor_backend.py:10) in double;
            # 1e-5 of the operator-norm bound n * 0.5 in single
            bound = 1e-7 if dtype == "complex128" else TOL[dtype] * 0.5 * nqubits
            assert math.isclose(want, res.real.get().item(), abs_tol=bound)
            assert math.isclose(want, float(res[0]), abs_tol=bound)


def test_pauli_string_pattern_runcard():
    n = 6
    circ = circuits.brickwork(n, 4, seed=1)
    psi = osv.simulate(circ)
    want = osv.pauli_expectation(psi, [(1.0, [("XZ"[q % 2], q) for q in range(n)])], n).real
    b = _backend("complex128")
    b.configure_tn_simulation({"MPI_enabled": False, "MPS_enabled": False, "NCCL_enabled": False,
                               "expectation_enabled": {"pauli_string_pattern": "XZ"}})
    assert math.isclose(want, float(b.execute_circuit(circ)[0]), abs_tol=1e-9)


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_amplitude_sliced_equals_unsliced_and_oracle(dtype):
    circ = circuits.brickwork(14, 8, seed=2)
    psi = osv.simulate(circ)
    bits = "01101001110100"
    want = psi[int(bits, 2)]
    conv = QiboCircuitToEinsum(circ, dtype)
    net = conv.amplitude_network(bits)
    vals = []
    for min_slices, graph in ((1, True), (8, True), (8, False), (64, True)):
        engine.clear_caches()
        out = engine.contract(net, dtype, samples=4, min_slices=min_slices, use_graph=graph)
        vals.append(complex(out.cpu().numpy().reshape(-1)[0]))
    for v in vals:
        assert abs(v - want) <= TOL[dtype] * abs(want), (v, want)
    # partial slice ranges add up (what each rank of a multi-GPU run computes)
    engine.clear_caches()
    info, low = engine.plan(net, dtype, samples=4, min_slices=8)
    parts = [engine.contract(net, dtype, info=info, slice_ids=ev.compute_slices(info.num_slices, r, 3))
             .cpu().numpy().reshape(-1)[0] for r in range(3)]
    assert abs(sum(parts) - want) <= TOL[dtype] * abs(want)
    # and each partial equals the oracle's sum over the same slice ids on the same path
    r1 = ev.compute_slices(info.num_slices, 1, 3)
    o = ocontract.contract_sliced(net.interleaved(), info.path, info.sliced_labels, r1,
                                  dtype=np.complex128)
    # a partial sum is not an amplitude: its error is relative to the terms it adds up, bounded by |want| + |o|
    assert abs(parts[1] - o) <= TOL[dtype] * (abs(o) + abs(want))


def test_dense_output_drops_inactive_qubits_and_initial_state_rejected():
    from qibotn_b200.qibo_compat import Circuit, H, CNOT
    c = Circuit(5)
    c.add(H(1)); c.add(CNOT(1, 3))
    b = _backend("complex128")
    res = b.execute_circuit(c)
    assert res.statevector.shape == (2, 2)          # only qubits 1 and 3 are active
    assert np.allclose(res.statevector.flatten().get(), [2 ** -0.5, 0, 0, 2 ** -0.5])
    with pytest.raises(NotImplementedError):
        b.execute_circuit(c, initial_state=np.ones(32))


# ----------------------------------------------------------------------------- golden fixtures
import os

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.npz"))
CLOSED = np.load(os.path.join(os.path.dirname(__file__), "golden", "closed_forms.npz"))


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_against_committed_golden_vectors(dtype):
    for n, layers in ((5, 2), (8, 3), (12, 3)):
        circ = circuits.ry_rz_cnot_ring(n, layers, seed=42)
        got = ev.dense_vector_tn(circ, dtype).flatten().get()
        assert rel2(got, GOLD[f"ring_state_{n}_{layers}"]) <= TOL[dtype]
        exp = float(ev.expectation_tn(circ, dtype, None)[0])
        assert abs(exp - float(GOLD[f"ring_xz_{n}_{layers}"])) <= TOL[dtype] * 0.5 * n
    circ = circuits.brickwork(12, 6, seed=0)
    z = float(ev.expectation_tn(circ, dtype, {"pauli_string_pattern": "Z"})[0])
    assert abs(z - float(GOLD["brick12x6_zall"])) <= TOL[dtype]
    # 53-qubit Sycamore-like circuit, 4 cycles: single amplitudes (the C4 code path at reduced depth)
    circ = circuits.sycamore_rqc(4, seed=0)
    for bits in ("0" * 53, "01" * 26 + "1"):
        want = complex(GOLD[f"syc53_d4_amp_{bits[:4]}"])
        got = complex(ev.amplitude_tn(circ, dtype, bits, n_samples=4).get().reshape(-1)[0])
        assert abs(got - want) <= TOL[dtype] * abs(want) + 1e-30


def test_qft_on_zero_state_is_uniform_at_reference_sizes():
    """reference tests: allclose(rtol 1e-5, atol 1e-8) against the uniform vector."""
    for n in (1, 2, 5, 10):
        got = ev.dense_vector_tn(circuits.qft(n, True), "complex128").flatten().get()
        assert np.allclose(got, CLOSED[f"qft_zero_{n}"], rtol=1e-5, atol=1e-8)


def test_c4_amplitude_properties_at_depth_8():
    """Full-width properties the domain offers (no dense oracle at 53 qubits): the same amplitude
    from two different contraction trees, from the tensor-core and the FFMA kernels, in complex64
    and complex128, and as a sum of slice partials."""
    from qibotn_b200 import lib
    circ = circuits.sycamore_rqc(8, seed=0)
    bits = "0" * 53
    net64 = QiboCircuitToEinsum(circ, "complex64").amplitude_network(bits)
    net128 = QiboCircuitToEinsum(circ, "complex128").amplitude_network(bits)
    engine.clear_caches()
    ref = complex(engine.contract(net128, "complex128", samples=4, seed=0).cpu().numpy().reshape(-1)[0])
    assert abs(ref) > 0
    vals = {}
    for seed in (0, 1):
        engine.clear_caches()
        vals[f"tc{seed}"] = complex(engine.contract(net64, "complex64", samples=4, seed=seed, min_slices=4)
                                    .cpu().numpy().reshape(-1)[0])
    lib.set_option("tc_enable", 0)
    try:
        engine.clear_caches()
        vals["ffma"] = complex(engine.contract(net64, "complex64", samples=4, seed=0).cpu().numpy().reshape(-1)[0])
    finally:
        lib.set_option("tc_enable", 1)
    for k, v in vals.items():
        assert abs(v - ref) <= 1e-5 * abs(ref), (k, v, ref)


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_fused_pauli_kernel_route_equals_network_route(dtype):
    """north_star (d): dense state + one-pass Pauli kernel vs one sandwich contraction per term."""
    n = 9
    circ = circuits.brickwork(n, 5, seed=4)
    psi = osv.simulate(circ)
    terms = [(0.7, [("X", 0), ("Y", 3), ("Z", 8)]), (-1.3, [("Y", 1), ("Y", 2)]), (0.4, [("Z", 4)]),
             (2.0, [("X", 5), ("X", 6), ("Z", 7), ("Y", 0)])]
    want = osv.pauli_expectation(psi, terms, n)
    obs = {"terms": [{"coefficient": c, "operators": f} for c, f in terms]}
    vals = {}
    for route in ("network", "state", "auto"):
        ev.EXPECTATION_ROUTE = route
        try:
            vals[route] = complex(ev.expectation_tn(circ, dtype, obs).get()[0])
        finally:
            ev.EXPECTATION_ROUTE = "auto"
    for route, v in vals.items():
        assert abs(v - want) <= TOL[dtype] * sum(abs(c) for c, _ in terms), (route, v, want)
    # inactive qubits stay |0>: Z acts as +1, X kills the term
    from qibotn_b200.qibo_compat import Circuit, H, CNOT
    c = Circuit(5)
    c.add(H(1)); c.add(CNOT(1, 3))
    obs = {"terms": [{"coefficient": 1.0, "operators": [("Z", 0), ("Z", 1), ("Z", 3)]},
                     {"coefficient": 5.0, "operators": [("X", 2), ("Z", 1)]},
                     {"coefficient": 0.5, "operators": [("X", 1), ("X", 3)]}]}
    for route in ("network", "state"):
        ev.EXPECTATION_ROUTE = route
        try:
            v = complex(ev.expectation_tn(c, dtype, obs).get()[0])
        finally:
            ev.EXPECTATION_ROUTE = "auto"
        assert abs(v - 1.5) <= TOL[dtype] * 6.5, (route, v)


@pytest.mark.parametrize("options", [{}, {"tc_gather": 2}, {"tc_gather": 1}, {"tc_prepack_a": 0}])
def test_c4_full_amplitude_within_tolerance_of_complex128(options):
    """The benchmark contraction itself, all slices: the complex64 tensor-core result must agree
    with the complex128 result of the same engine (FP64 FMA kernels, 70 s on a B200; value recorded
    in tests/golden/c4_amplitude.json together with the complex64 FFMA and the other-plan values)
    to the north-star tolerance of 1e-5 relative."""
    import json
    from qibotn_b200 import executor, plans, workloads
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "c4_amplitude.json")))
    want = complex(*ref["complex128"])
    wl = workloads.build("c4")
    net = wl.network()
    info = plans.load_plan("c4", net.modes, net.out)
    assert info is not None, "plans/c4.json does not match the workload"
    # the switches of the staggered kernel's A path (bulk-copied raw chunks, address-ordered cp.async gather,
    # no A pre-pack) on the real layouts and tile counts of the plan: races there do not show on small tensors
    from qibotn_b200 import lib
    saved = {key: lib.get_option(key) for key in options}
    try:
        for key, val in options.items():
            lib.set_option(key, val)
        low = executor.lower(net.modes, net.out, info, "complex64")
    finally:
        for key, val in saved.items():
            lib.set_option(key, val)
    runner = executor.ContractionRunner(low, use_graph=True)
    runner.upload(net.tensors)
    got = complex(runner.run(range(info.num_slices)).reshape(-1)[0].item())
    del runner
    torch.cuda.empty_cache()
    assert abs(got - want) <= 1e-5 * abs(want), (options, got, want, abs(got - want) / abs(want))


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_c4_full_circuit_slices_against_the_cpu_oracle_fixture(dtype):
    """Engine-independent parity on the benchmark circuit itself: single slices of the CPU-sized plan
    (plans/c4_cpu.json, width 2^24, 2^21 slices) of the 53-qubit depth-14 amplitude, against the partial sums the
    numpy oracle produced for the same slices in complex128 (tests/golden/c4_slice_partials.json, written by
    scripts/make_golden_c4_slices.py).  Tolerance 1e-10 (complex128) / 1e-5 (complex64) relative to the largest
    |partial| of the fixture (the partials are within a factor 4 of each other; each is itself a sum with
    cancellations, so its own magnitude is not the scale of its rounding error)."""
    import json
    from qibotn_b200 import engine, plans, workloads
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "c4_slice_partials.json")))
    wl = workloads.build("c4")
    net = QiboCircuitToEinsum(wl.circuit, dtype=dtype).amplitude_network(wl.bitstring)
    info = plans.load_plan("c4_cpu", net.modes, net.out)
    assert info is not None and info.num_slices == gold["num_slices"]
    scale = max(abs(complex(re, im)) for re, im in gold["slices"].values())
    for sid, (re, im) in gold["slices"].items():
        want = complex(re, im)
        got = complex(engine.contract(net, dtype, info=info, slice_ids=[int(sid)]).reshape(-1)[0].item())
        assert abs(got - want) <= TOL[dtype] * scale, (dtype, sid, got, want, abs(got - want) / scale)
    want = complex(*gold["sum_0_to_3"])
    got = complex(engine.contract(net, dtype, info=info, slice_ids=range(4)).reshape(-1)[0].item())
    assert abs(got - want) <= TOL[dtype] * scale


def test_c5_p4_terms_from_the_committed_plans_in_both_precisions():
    """Config 5 as named (40-qubit QAOA MaxCut, p = 4): light-cone terms of the committed plan store (plans/c5_p4,
    hyper-indexed networks of 540-800 tensors planned by the hyper-aware tree search) contracted in complex64 -- batch
    labels on the tcgen05 kernel and on the K reduction -- and in complex128 with the SAME trees; |<ZZ>| <= 1, so the
    difference is relative to the operator norm of a term: <= 1e-5.  Terms 4 and 3 are the cheapest (10^8.2, 10^8.6
    flops), 0 and 9 cost 10^11.2-10^11.5, 10 is sliced (2 slices, 10^12.8 flops)."""
    from qibotn_b200 import engine as eng
    store = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "plans", "c5_p4")
    circ, edges = circuits.qaoa_maxcut(40, 4, seed=0)
    obs = circuits.zz_terms(edges)
    old_store = eng.PLAN_STORE
    eng.PLAN_STORE = store
    try:
        nets64 = list(ev._term_networks(circ, "complex64", obs))
        nets128 = list(ev._term_networks(circ, "complex128", obs))
        for k in (4, 3, 0, 9, 10):
            info = eng._stored_plan(nets64[k], eng.default_width("complex64"), 1)
            assert info is not None, f"plans/c5_p4 holds no plan for term {k}"
            if info.log2_width > 30 and k != 10:
                continue
            v64 = complex(eng.contract(nets64[k], "complex64", info=info).reshape(()).item())
            eng.clear_caches()
            if info.log2_width <= 29:
                v128 = complex(eng.contract(nets128[k], "complex128", info=info).reshape(()).item())
                eng.clear_caches()
                assert abs(v64 - v128) <= 1e-5, (k, v64, v128)
                assert abs(v128.imag) <= 1e-10 and abs(v128.real) <= 1.0 + 1e-10
            else:
                # sliced term: both slices one by one must add up to the whole, whatever the order
                parts = [complex(eng.contract(nets64[k], "complex64", info=info, slice_ids=[s]).reshape(()).item())
                         for s in range(info.num_slices)]
                eng.clear_caches()
                assert abs(sum(parts) - v64) <= 1e-5 and abs(v64.imag) <= 1e-5 and abs(v64.real) <= 1.0 + 1e-5
    finally:
        eng.PLAN_STORE = old_store
        eng.clear_caches()
        torch.cuda.empty_cache()


def test_c2_full_size_expectation_properties():
    """BASELINE config 2 at full size (30-qubit brickwork depth 12, <Z...Z>, one closed sandwich
    network of 1158 tensors): no dense oracle reaches 30 qubits here, so the value is checked through
    path independence (two planner seeds), precision independence (complex64 vs complex128) and
    kernel independence (tensor-core vs FFMA).  Tolerance: 1e-5 of max(|value|, 1e-3) -- the
    expectation of a 30-body Pauli string on a scrambled state is itself tiny."""
    from qibotn_b200 import lib, workloads
    wl = workloads.build("c2")
    vals = {}
    for dtype in ("complex128", "complex64"):
        wl.dtype = dtype
        net = wl.network()
        for seed in (0, 1):
            engine.clear_caches()
            vals[(dtype, seed)] = complex(engine.contract(net, dtype, samples=4, seed=seed).cpu().numpy().reshape(-1)[0])
    wl.dtype = "complex64"
    net = wl.network()
    lib.set_option("tc_enable", 0)
    try:
        engine.clear_caches()
        vals[("complex64", "ffma")] = complex(engine.contract(net, "complex64", samples=4, seed=0).cpu().numpy().reshape(-1)[0])
    finally:
        lib.set_option("tc_enable", 1)
    ref = vals[("complex128", 0)]
    assert abs(ref.imag) < 1e-9                      # a Hermitian observable
    scale = max(abs(ref), 1e-3)
    assert abs(vals[("complex128", 1)] - ref) < 1e-10 * scale
    for key, v in vals.items():
        if key[0] == "complex64":
            assert abs(v - ref) < 1e-5 * scale, (key, v, ref)


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_edge_cases(dtype):
    """Empty and degenerate inputs: no gates, measurement markers only, one gate, repeated gates
    on one qubit, a qubit untouched in the middle, expectation on an empty circuit."""
    from qibotn_b200.qibo_compat import Circuit, H, X, Z, CNOT, M, RX
    b = _backend(dtype)
    c = Circuit(3)
    assert b.execute_circuit(c).statevector.flatten().get().tolist() == [1]         # no active qubit
    c.add(M(0, 1, 2))
    assert b.execute_circuit(c).statevector.flatten().get().tolist() == [1]
    c = Circuit(1); c.add(X(0))
    assert np.allclose(b.execute_circuit(c).statevector.flatten().get(), [0, 1])
    c = Circuit(4)
    for _ in range(5):
        c.add(RX(2, theta=0.3))
    c.add(H(0)); c.add(CNOT(0, 3))
    got = b.execute_circuit(c).statevector.flatten().get()                          # qubits 0, 2, 3 active
    want = osv.simulate(c).reshape(2, 2, 2, 2)[:, 0, :, :].reshape(-1)
    assert rel2(got, want) <= TOL[dtype]
    # expectation of Z0 on an empty circuit is <0|Z|0> = 1, X0 gives 0
    for letter, want in (("Z", 1.0), ("X", 0.0)):
        b.configure_tn_simulation({"MPI_enabled": False, "MPS_enabled": False, "NCCL_enabled": False,
                                   "expectation_enabled": {"terms": [{"coefficient": 1.0, "operators": [(letter, 0)]}]}})
        assert abs(float(b.execute_circuit(Circuit(2))[0]) - want) < 1e-6
    # MPS on the same degenerate circuits
    b.configure_tn_simulation({"MPI_enabled": False, "MPS_enabled": True, "NCCL_enabled": False,
                               "expectation_enabled": False})
    got = b.execute_circuit(Circuit(3)).statevector.flatten().get()
    assert np.allclose(got, np.eye(8)[0])
    got = b.execute_circuit(c).statevector.flatten().get()
    assert rel2(got, osv.simulate(c)) <= TOL[dtype]


@pytest.mark.parametrize("ansatz", ["mps", None])
def test_quimb_platform_api(ansatz):
    """The quimb platform's calls (reference backends/quimb.py:123-261) on this engine: dense state,
    symbolic expectation, sampling."""
    n = 7
    circ = circuits.ry_rz_cnot_ring(n, 2, seed=3)
    psi = osv.simulate(circ)
    b = MetaBackend.load("quimb")
    b.configure_tn_simulation(ansatz=ansatz, max_bond_dimension=None, svd_cutoff=1e-12)
    res = b.execute_circuit(circ, return_array=True)
    assert np.abs(res.statevector.get() - psi).max() < 1e-9
    ops, sites, coeffs = ["xz", "yy", "z", "zxz"], [(0, 1), (2, 5), (6,), (1, 3, 4)], [0.5, -1.25, 2.0, 0.75]
    want = osv.pauli_expectation(psi, [(c, [(l.upper(), s) for l, s in zip(o, st)])
                                       for o, st, c in zip(ops, sites, coeffs)], n).real
    got = b.exp_value_observable_symbolic(circ, ops, sites, coeffs, n)
    assert abs(got - want) < 1e-8
    res = b.execute_circuit(circ, nshots=4000)
    assert sum(res.frequencies().values()) == 4000 and res.statevector is None
    probs = np.abs(psi) ** 2
    for state, p in res.probabilities().items():
        assert abs(p - probs[int(state, 2)]) < 1e-9
    top = max(res.frequencies(), key=res.frequencies().get)
    assert abs(res.frequencies()[top] / 4000 - probs[int(top, 2)]) < 5 * np.sqrt(probs[int(top, 2)] / 4000)


def test_arena_pool_changes_no_value():
    """executor.ArenaPool (one grow-only device buffer for the arenas of successive terms, scripts/run_c5_p4.py):
    the terms of an observable contracted one after the other give the same values with and without it, and match
    the dense simulator."""
    from qibotn_b200 import executor
    from qibotn_b200.observables import get_ham_gates
    n, dtype = 12, "complex64"
    circ = circuits.brickwork(n, 4, seed=9)
    psi = osv.simulate(circ)
    conv = QiboCircuitToEinsum(circ, dtype)
    terms = ([(5, "Z", 1.0)], [(0, "X", 1.0), (11, "Y", 1.0)], [(6, "Z", 1.0), (7, "Z", 1.0)], [(2, "Y", 1.0)])
    nets = [conv.expectation_network(get_ham_gates(t, dtype)) for t in terms]

    def run_all():
        vals = []
        for net in nets:
            vals.append(complex(engine.contract(net, dtype, samples=4).cpu().numpy().reshape(-1)[0]))
            engine.clear_caches()                      # one live runner at a time, as the pool requires
        return vals

    engine.clear_caches()
    plain = run_all()
    assert executor.ARENA_POOL is None
    executor.ARENA_POOL = executor.ArenaPool(reserve_bytes=1 << 20)
    try:
        pooled = run_all()
        assert executor.ARENA_POOL.buf is not None and executor.ARENA_POOL.grown >= 1
    finally:
        executor.ARENA_POOL = None
        engine.clear_caches()
    for t, a, b in zip(terms, plain, pooled):
        want = osv.pauli_expectation(psi, [(1.0, [(l, q) for q, l, _ in t])], n)
        assert abs(a - b) <= 1e-6 and abs(b - want) <= TOL[dtype]


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_light_cone_cancellation_changes_no_value(dtype):
    """Dropping the gate / inverse-gate pairs outside the observable's reverse light cone is exact:
    same expectation with and without it, and equal to the dense simulator's."""
    from qibotn_b200.observables import get_ham_gates
    n = 12
    circ = circuits.brickwork(n, 4, seed=9)
    psi = osv.simulate(circ)
    conv = QiboCircuitToEinsum(circ, dtype)
    for term in ([(5, "Z", 0.7)], [(0, "X", 1.0), (11, "Y", -2.0)], [(6, "Z", 1.0), (7, "Z", 1.0)]):
        hg = get_ham_gates(term, dtype)
        full = conv.expectation_network(hg, light_cone=False)
        cone = conv.expectation_network(hg)
        assert len(cone.tensors) < len(full.tensors)
        engine.clear_caches()
        a = complex(engine.contract(full, dtype, samples=4).cpu().numpy().reshape(-1)[0])
        b = complex(engine.contract(cone, dtype, samples=4).cpu().numpy().reshape(-1)[0])
        coeff = np.prod([c for _, _, c in term])
        want = coeff * osv.pauli_expectation(psi, [(1.0, [(l, q) for q, l, _ in term])], n)
        assert abs(a - want) <= TOL[dtype] * abs(coeff) and abs(b - want) <= TOL[dtype] * abs(coeff)


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_qaoa_maxcut_expectation(dtype):
    """BASELINE config 5's kind of workload at sizes a dense oracle reaches: QAOA MaxCut on a
    3-regular graph, sum of Z_a Z_b over the edges (one light-cone-reduced contraction per term,
    plans shared between terms of equal structure)."""
    n, p = 14, 3
    circ, edges = circuits.qaoa_maxcut(n, p, seed=1)
    psi = osv.simulate(circ)
    want = osv.pauli_expectation(psi, [(1.0, [("Z", a), ("Z", b)]) for a, b in edges], n).real
    # "network" twice: diagonal gates as hyper-indexed diagonals (the default) and as the reference's
    # full rank-4 tensors -- same value
    for route, hyper in (("network", True), ("network", False), ("state", True)):
        ev.EXPECTATION_ROUTE, samples, hyper_default = route, ev.PLAN_SAMPLES, ev.HYPER_DIAGONAL
        ev.PLAN_SAMPLES = 4                      # 21 differently shaped terms: keep the planner's share of the test small
        ev.HYPER_DIAGONAL = hyper
        engine.clear_caches()
        try:
            got = float(ev.expectation_tn(circ, dtype, circuits.zz_terms(edges)).real.get().reshape(-1)[0])
        finally:
            ev.EXPECTATION_ROUTE, ev.PLAN_SAMPLES, ev.HYPER_DIAGONAL = "auto", samples, hyper_default
        assert abs(got - want) <= TOL[dtype] * len(edges), (route, hyper, got, want)     # sum_i |c_i| = #edges


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_results_survive_a_second_run_of_the_same_structure(dtype):
    """A parametrised circuit run twice reuses the plan, the runner and its output buffer: the first result
    must be the caller's own copy (the reference's `contract` returns a fresh array every time)."""
    from qibotn_b200.qibo_compat import Circuit, H, CNOT, RY

    def circuit(theta):
        c = Circuit(6)
        for q in range(6):
            c.add(H(q))
        for q in range(5):
            c.add(CNOT(q, q + 1))
            c.add(RY(q + 1, theta=theta * (q + 1)))
        return c

    first = ev.dense_vector_tn(circuit(0.3), dtype)
    keep = first.flatten().get().copy()
    second = ev.dense_vector_tn(circuit(1.1), dtype)
    assert rel2(first.flatten().get(), keep) == 0.0
    assert rel2(first.flatten().get(), osv.simulate(circuit(0.3))) <= TOL[dtype]
    assert rel2(second.flatten().get(), osv.simulate(circuit(1.1))) <= TOL[dtype]
    b = _backend(dtype)
    r1 = b.execute_circuit(circuit(0.3))
    r2 = b.execute_circuit(circuit(1.1))
    assert rel2(r1.statevector.flatten().get(), keep) <= TOL[dtype]
    assert rel2(r2.statevector.flatten().get(), osv.simulate(circuit(1.1))) <= TOL[dtype]


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_fused_multi_term_pauli_kernel(dtype):
    """qtn_pauli_expectation_host: terms sharing an xmask are evaluated up to 16 per pass over psi; every term must
    equal the one-term-per-pass kernel and the numpy value, whatever the grouping (more than 16 terms with one
    xmask, single-term groups, identity, Y phases)."""
    import ctypes as C
    from qibotn_b200 import lib

    n = 11
    rng = np.random.default_rng(3)
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    psi = (psi / np.linalg.norm(psi)).astype(dtype)
    terms = []
    for _ in range(37):                                   # 37 Z-only strings: one xmask, three passes
        terms.append({q: "Z" for q in rng.choice(n, size=rng.integers(1, 5), replace=False)})
    for _ in range(9):                                    # XX/YY/XY-type strings on a few pairs: shared xmasks
        a, b = rng.choice(n, size=2, replace=False)
        for letters in (("X", "X"), ("Y", "Y"), ("X", "Y")):
            t = {int(a): letters[0], int(b): letters[1]}
            if rng.random() < 0.5:
                t[int((a + 3) % n)] = "Z" if (a + 3) % n not in (a, b) else t[int((a + 3) % n)]
            terms.append(t)
    terms.append({})                                      # identity
    terms.append({0: "Y", 5: "Y", 7: "Y", 2: "X"})
    xm, zm, ny = [], [], []
    for t in terms:
        fx = sum(1 << (n - 1 - q) for q, l in t.items() if l in "XY")
        fz = sum(1 << (n - 1 - q) for q, l in t.items() if l in "ZY")
        xm.append(fx); zm.append(fz); ny.append(sum(l == "Y" for l in t.values()))
    want = [complex(osv.pauli_expectation(psi.astype(np.complex128), [(1.0, [(l, q) for q, l in t.items()])], n))
            for t in terms]
    h = lib.load()
    st = torch.cuda.current_stream().cuda_stream
    d_psi = torch.from_numpy(psi).cuda()
    nt = len(terms)
    out_h = torch.empty(nt, dtype=torch.complex128, device="cuda")
    lib.check(h.qtn_pauli_expectation_host(lib.dtype_code(dtype), n, d_psi.data_ptr(), nt, (C.c_uint64 * nt)(*xm),
                                           (C.c_uint64 * nt)(*zm), (C.c_int32 * nt)(*ny), out_h.data_ptr(), st))
    out_d = torch.empty(nt, dtype=torch.complex128, device="cuda")
    # the mask arrays must outlive the launch (a temporary's block is handed to the next allocation)
    d_xm = torch.tensor(xm, dtype=torch.int64, device="cuda")
    d_zm = torch.tensor(zm, dtype=torch.int64, device="cuda")
    d_ny = torch.tensor(ny, dtype=torch.int32, device="cuda")
    lib.check(h.qtn_pauli_expectation(lib.dtype_code(dtype), n, d_psi.data_ptr(), nt, d_xm.data_ptr(),
                                      d_zm.data_ptr(), d_ny.data_ptr(), out_d.data_ptr(), st))
    got_h, got_d = out_h.cpu().numpy(), out_d.cpu().numpy()
    assert np.abs(got_h - got_d).max() <= 1e-12
    assert np.abs(got_h - np.array(want)).max() <= TOL[dtype]      # |<P>| <= 1: absolute = relative to the norm bound
