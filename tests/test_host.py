"""Host-side logic that needs no GPU: the plugin boundary (runcard type rules, dispatch errors,
result container), observable normalisation, the C-ABI surface (every symbol the header declares
is exported; ctypes mirrors match the C structs), and that compute paths fail loudly without CUDA."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from qibotn_b200 import MetaBackend, circuits, lib
from qibotn_b200 import eval as ev
from qibotn_b200 import observables as obs
from qibotn_b200.network import QiboCircuitToEinsum
from qibotn_b200.result import TensorNetworkResult

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plugin_alias_and_platforms():
    import qibotn
    assert qibotn.MetaBackend is MetaBackend
    b = MetaBackend.load("cutensornet")
    assert (b.name, b.platform, b.supports_multigpu) == ("qibotn", "cutensornet", True)
    assert MetaBackend.load("qutensornet").platform == "cutensornet"        # README's spelling
    with pytest.raises(NotImplementedError):
        MetaBackend.load("qmatchatea")
    assert MetaBackend().list_available() == {"cutensornet": True, "qutensornet": True, "quimb": True}


def test_quimb_platform_surface():
    """reference backends/quimb.py:40-83,123-135,205-236: constructor, configuration, validation."""
    b = MetaBackend.load("quimb", quimb_backend="numpy", contraction_optimizer="greedy")
    assert (b.name, b.platform, b.backend, b.contractions_optimizer) == ("qibotn", "quimb", "numpy", "greedy")
    assert (b.ansatz, b.max_bond_dimension, b.svd_cutoff, b.n_most_frequent_states) == ("mps", None, 1e-10, 100)
    b.configure_tn_simulation(ansatz=None, max_bond_dimension=16, svd_cutoff=1e-6, n_most_frequent_states=5)
    assert (b.ansatz, b.max_bond_dimension) == (None, 16)
    assert b._gate_algo()["svd_method"] == {"partition": "UV", "abs_cutoff": 1e-6, "max_extent": 16}
    with pytest.raises(ValueError):
        MetaBackend.load("quimb", quimb_backend="tensorflow")
    with pytest.raises(ValueError):
        b.execute_circuit(circuits.qft(2, True), initial_state=np.ones(4))
    with pytest.raises(ValueError):
        b.exp_value_observable_symbolic(circuits.qft(3, True), ["zz"], [(1, 1)], [1.0], 3)


def test_runcard_type_rules():
    """reference backends/cutensornet.py:24-69."""
    b = MetaBackend.load("cutensornet")
    assert not (b.MPI_enabled or b.NCCL_enabled or b.MPS_enabled or b.expectation_enabled)
    b.configure_tn_simulation({"MPI_enabled": False, "NCCL_enabled": True, "MPS_enabled": False,
                               "expectation_enabled": True})
    assert b.NCCL_enabled and b.expectation_enabled and b.observable is None
    b.configure_tn_simulation({"MPS_enabled": True, "expectation_enabled": False})
    assert b.gate_algo == {"qr_method": False, "svd_method": {"partition": "UV", "abs_cutoff": 1e-12}}
    custom = {"qr_method": False, "svd_method": {"partition": "UV", "abs_cutoff": 1e-6, "max_extent": 8}}
    b.configure_tn_simulation({"MPS_enabled": custom, "expectation_enabled": False})
    assert b.gate_algo is custom
    b.configure_tn_simulation({"MPS_enabled": False, "expectation_enabled": {"pauli_string_pattern": "IXZ"}})
    assert b.expectation_enabled and b.observable == {"pauli_string_pattern": "IXZ"}
    with pytest.raises(TypeError):
        b.configure_tn_simulation({"MPS_enabled": False})                      # expectation_enabled missing
    with pytest.raises(TypeError):
        b.configure_tn_simulation({"MPS_enabled": "yes", "expectation_enabled": False})
    with pytest.raises(TypeError):
        b.configure_tn_simulation({"MPS_enabled": False, "expectation_enabled": 3})


def test_dispatch_errors():
    """reference backends/cutensornet.py:87-88,151-152."""
    b = MetaBackend.load("cutensornet")
    circ = circuits.qft(3, True)
    with pytest.raises(NotImplementedError):
        b.execute_circuit(circ, initial_state=np.ones(8))
    b.configure_tn_simulation({"MPI_enabled": True, "NCCL_enabled": False, "MPS_enabled": True,
                               "expectation_enabled": False})
    with pytest.raises(NotImplementedError):
        b.execute_circuit(circ)
    for meth in ("apply_gate", "apply_gate_density_matrix"):
        with pytest.raises(NotImplementedError):
            getattr(b, meth)(None, None, 3)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    b = MetaBackend.load("cutensornet")
    with pytest.raises(lib.QtnError):
        b.execute_circuit(circuits.qft(3, True))
    b.configure_tn_simulation({"MPS_enabled": True, "expectation_enabled": False})
    with pytest.raises(lib.QtnError):
        b.execute_circuit(circuits.qft(3, True))


def test_result_container():
    """reference result.py:12-66."""
    r = TensorNetworkResult(nqubits=3, backend=None, measures=None, measured_probabilities=None,
                            prob_type=None, statevector=np.arange(8))
    assert (r.state() == np.arange(8)).all()
    with pytest.raises(ValueError):
        r.frequencies()
    big = TensorNetworkResult(nqubits=20, backend=None, measures={"00": 3}, measured_probabilities={"U": {"0": (0.1, 0.4)}},
                              prob_type="U", statevector=None)
    with pytest.raises(NotImplementedError):
        big.state()
    assert big.frequencies() == {"00": 3}
    assert abs(big.probabilities()["0"] - 0.3) < 1e-15


def test_observable_helpers():
    """reference eval.py:15-150."""
    ham = obs.check_observable(None, 3)
    terms = obs.extract_gates_and_qubits(ham)
    assert terms == [[(0, "X", 0.5), (1, "Z", 1.0)], [(1, "X", 0.5), (2, "Z", 1.0)], [(2, "X", 0.5), (0, "Z", 1.0)]] \
        or sorted(map(sorted, terms)) == sorted(map(sorted, [[(0, "X", 0.5), (1, "Z", 1.0)], [(1, "X", 0.5), (2, "Z", 1.0)],
                                                             [(0, "Z", 0.5), (2, "X", 1.0)]]))
    d = {"terms": [{"coefficient": 1.5, "operators": [("Y", 2)]}]}
    t = obs.extract_gates_and_qubits(obs.check_observable(d, 3))
    assert len(t) == 1 and [(q, l) for q, l, _ in t[0]] == [(0, "I"), (1, "I"), (2, "Y")]
    assert abs(np.prod([c for _, _, c in t[0]]) - 1.5) < 1e-15
    gates = obs.get_ham_gates(t[0], "complex64")
    assert gates[2][0].dtype == np.complex64 and gates[2][1] == (2,)
    with pytest.raises(ValueError):
        obs.get_ham_gates([(0, "Q", 1.0)])
    with pytest.raises(ValueError):
        obs.create_hamiltonian_from_dict({"terms": []}, 3)
    with pytest.raises(TypeError):
        obs.check_observable(1.0, 3)
    p = obs.extract_gates_and_qubits(obs.check_observable({"pauli_string_pattern": "XZ"}, 5))
    assert [(q, l) for q, l, _ in p[0]] == [(0, "X"), (1, "Z"), (2, "X"), (3, "Z"), (4, "X")]


def test_converter_operand_conventions():
    """reference circuit_convertor.py:28-58 (active qubits only) and :198-246 (all qubits)."""
    c = circuits.Circuit(4)
    c.add(circuits.gates.H(1)); c.add(circuits.gates.CNOT(1, 3)); c.add(circuits.gates.M(0, 1, 2, 3))
    conv = QiboCircuitToEinsum(c, "complex64")
    assert list(conv.active_qubits) == [1, 3] and len(conv.gate_tensors) == 2
    inter = conv.state_vector_operands()
    assert len(inter) == 2 * 4 + 1 and len(inter[-1]) == 2 and inter[0].dtype == np.complex64
    assert conv.gate_tensors[1][0].shape == (2, 2, 2, 2) and conv.gate_tensors[1][1] == (1, 3)
    exp = conv.expectation_operands(obs.get_ham_gates([(0, "Z", 1.0)], "complex64"))
    assert len(exp) == 2 * (4 + 2 + 1 + 2 + 4)
    # the engine's own network drops gate / inverse pairs outside the observable's light cone
    assert len(conv.expectation_network(obs.get_ham_gates([(0, "Z", 1.0)], "complex64")).tensors) == 4 + 1 + 4
    assert len(conv.expectation_network(obs.get_ham_gates([(3, "Z", 1.0)], "complex64")).tensors) == 4 + 2 + 1 + 2 + 4
    with pytest.raises(ValueError):
        conv.amplitude_operands("012")


def test_compute_slices_matches_reference_formula():
    """reference eval.py:197-205."""
    for ns, size in ((32, 1), (32, 3), (33, 4), (5, 8), (512, 8)):
        got = [ev.compute_slices(ns, r, size) for r in range(size)]
        assert [x for p in got for x in p] == list(range(ns))
        chunk, extra = divmod(ns, size)
        assert [len(p) for p in got] == [chunk + (1 if r < extra else 0) for r in range(size)]


def test_header_symbols_are_exported_and_structs_match():
    text = open(os.path.join(ROOT, "include", "qibotn_b200.h")).read()
    declared = set(re.findall(r"\b(qtn_[a-z0-9_]+)\s*\(", text))
    h = lib.load()
    for name in declared:
        assert hasattr(h, name), f"{name} declared in the header but not exported"
    assert declared == set(lib.exported_symbols()) | {"qtn_sizeof"} or declared >= set(lib.exported_symbols())
    h.qtn_sizeof.restype, h.qtn_sizeof.argtypes = C.c_int64, [C.c_char_p]
    for cname, ctype in (("qtn_pair_desc", lib.PairDesc), ("qtn_pair_plan_t", lib.PairPlan),
                         ("qtn_svd_item", lib.SvdItem), ("qtn_svd_trunc", lib.SvdTrunc)):
        assert h.qtn_sizeof(cname.encode()) == C.sizeof(ctype), cname
    assert h.qtn_sizeof(b"nonsense") == -1
    assert b"sm_100a" in h.qtn_version()
    with pytest.raises(lib.QtnError):
        lib.set_option("no_such_option", 1)


def test_partition_units_covers_every_term_slice_once():
    """dist.partition_units: the flattened (term, slice) list cut by the reference's slice rule
    (eval.py:197-205) -- contiguous, complete, disjoint, low ranks take the remainder."""
    from qibotn_b200 import dist
    for counts in ([1] * 60, [1, 4, 1, 1, 32, 1, 2], [64], [], [3, 0, 5]):
        for size in (1, 2, 3, 8, 70):
            seen, sizes = [], []
            for r in range(size):
                mine = dist.partition_units(counts, r, size)
                units = [(t, s) for t, rg in mine.items() for s in rg]
                sizes.append(len(units))
                seen += units
            assert seen == [(t, s) for t, c in enumerate(counts) for s in range(c)]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_partition_units_by_cost():
    """With per-slice costs the flattened list is cut where the cumulative cost crosses multiples of total / size:
    still contiguous, complete and disjoint, and no rank carries more than its share plus one unit."""
    from qibotn_b200 import dist
    rng = np.random.default_rng(5)
    for counts in ([1] * 60, [1, 4, 1, 1, 32, 1, 2], [64], [3, 0, 5], [1, 128, 1, 512, 2]):
        weights = [float(10.0 ** rng.uniform(-3, 3)) for _ in counts]
        total = sum(c * w for c, w in zip(counts, weights))
        for size in (1, 2, 3, 8):
            seen, loads = [], []
            for r in range(size):
                mine = dist.partition_units(counts, r, size, weights)
                units = [(t, s) for t, rg in mine.items() for s in rg]
                loads.append(sum(weights[t] for t, _ in units))
                seen += units
            assert seen == [(t, s) for t, c in enumerate(counts) for s in range(c)]
            assert max(loads) <= total / size + max(weights) * (1 + 1e-9)
    assert dist.partition_units([2, 3], 1, 2, [0.0, 0.0]) == dist.partition_units([2, 3], 1, 2)


def test_hyper_tree_legs_match_the_path_walk():
    """treeopt.HyperTree keeps a label open until ALL its owners are inside a subtree (owner counts as bit-sliced
    counters); its costs must be those planner.path_cost computes for the same pairs, before and after subtree
    reconfiguration and annealing, and find_path must use it for hyper-indexed (QAOA) networks."""
    import random
    from qibotn_b200 import circuits, planner
    from qibotn_b200 import eval as ev
    from qibotn_b200.treeopt import HyperTree
    circ, edges = circuits.qaoa_maxcut(12, 2, seed=0)
    net = list(ev._term_networks(circ, "complex64", circuits.zz_terms(edges)))[3]
    sn = planner.SymbolicNetwork(net.modes, net.out)
    assert sn.hyper_mask
    pre, live, nxt = planner._absorb_small(sn)
    tail = planner._elimination_path(live, sn.out_mask, sn.hyper_mask, nxt, random.Random(0), 0.0)
    pre_cost = planner.path_cost(sn, pre + tail, 0)[0] - HyperTree(live, tail, nxt, sn.out_mask).total_cost()
    tree = HyperTree(live, tail, nxt, sn.out_mask)
    start = tree.total_cost()
    tree.reconfigure(rounds=4, k=8)
    tree.anneal(5000, random.Random(1))
    path = pre + tree.ssa_pairs(nxt)
    cost, width, _, _ = planner.path_cost(sn, path, 0)
    assert abs(cost - pre_cost - tree.total_cost()) <= 1e-9 * cost and width == max(width, tree.width())
    assert tree.total_cost() <= start
    info = planner.find_path(net.modes, net.out, samples=2, seed=0, target_log2_width=30)
    assert info.opt_cost <= cost * 1.5 and info.log2_width <= 30


def test_plan_store_round_trip(tmp_path, monkeypatch):
    """engine.PLAN_STORE: a plan found for a network structure is written to disk and a fresh cache reads
    it back instead of searching again (offline planning of many-term observables, scripts/plan_c5.py)."""
    from qibotn_b200 import circuits, engine
    from qibotn_b200 import planner
    from qibotn_b200.network import QiboCircuitToEinsum
    from qibotn_b200.observables import get_ham_gates
    circ, edges = circuits.qaoa_maxcut(8, 2, seed=1)
    conv = QiboCircuitToEinsum(circ, "complex64")
    net = conv.expectation_network(get_ham_gates([(edges[0][0], "Z", 1.0), (edges[0][1], "Z", 1.0)], dtype="complex64"),
                                   hyper=True)
    monkeypatch.setattr(engine, "PLAN_STORE", str(tmp_path))
    engine.clear_caches()
    info = engine.find_info(net, "complex64", samples=2)
    files = list(tmp_path.iterdir())
    assert len(files) == 1 and files[0].name.endswith("_w31_s1.json")
    engine.clear_caches()
    calls = []
    monkeypatch.setattr(engine, "find_path", lambda *a, **k: calls.append(1))
    again = engine.find_info(net, "complex64", samples=2)
    assert not calls and again.path == info.path and again.sliced_labels == info.sliced_labels
    # the committed plans of config 5 at p = 3 belong to the networks the converter builds today
    import os
    store = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "plans", "c5_p3")
    if os.path.isdir(store):
        from qibotn_b200 import eval as ev
        monkeypatch.setattr(engine, "PLAN_STORE", store)
        circ, edges = circuits.qaoa_maxcut(40, 3, seed=0)
        nets = list(ev._term_networks(circ, "complex64", circuits.zz_terms(edges)))
        assert len(nets) == 60
        for net in nets[:5]:
            stored = engine._stored_plan(net, engine.default_width("complex64"), 1)
            assert stored is not None
            cost, width, _, _ = planner.path_cost(planner.SymbolicNetwork(net.modes, net.out), stored.path,
                                                  planner.SymbolicNetwork(net.modes, net.out).to_mask(stored.sliced_labels)
                                                  if stored.sliced_labels else 0)
            assert width <= 31 and abs(cost - stored.opt_cost) <= 1e-6 * stored.opt_cost



def test_gate_algo_rules():
    """``gate_algo`` follows cuQuantum's ContractDecomposeAlgorithm: QR split by default, SVD needs a
    partition (the reference's ``a, _, b = contract_decompose(...)`` would drop separate singular values)."""
    from qibotn_b200 import mps

    for algo in (None, {}, {"qr_method": {}}, {"qr_method": {}, "svd_method": False}):
        tr, part, qr_only, assist = mps.parse_algorithm(algo)
        assert qr_only and not assist and tr.max_extent == 0
    tr, part, qr_only, assist = mps.parse_algorithm({"qr_method": False,
                                                     "svd_method": {"partition": "UV", "abs_cutoff": 1e-12}})
    assert (part, qr_only, assist, tr.abs_cutoff) == (3, False, False, 1e-12)
    assert mps.parse_algorithm({"svd_method": {"partition": "U", "max_extent": 8}})[3] is True     # qr_method default {}
    with pytest.raises(ValueError):
        mps.parse_algorithm({"qr_method": False, "svd_method": False})
    with pytest.raises(ValueError):
        mps.parse_algorithm({"svd_method": {"abs_cutoff": 1e-12}})            # partition None
    with pytest.raises(ValueError):
        mps.parse_algorithm({"svd_method": {"partition": "UV", "cutoff": 1}})
    with pytest.raises(ValueError):
        mps.parse_algorithm({"svd": {}})


def test_reference_module_names_resolve_to_this_engine():
    """What the reference's tests and backends import (tests/test_quimb_backend.py:39, test_cuquantum_cutensor_backend.py,
    backends/cutensornet.py:85): every name is the one qibotn_b200 module object, not a second copy."""
    import importlib

    import qibotn
    import qibotn_b200
    from qibotn.backends.abstract import QibotnBackend
    from qibotn.backends.cutensornet import CuTensorNet
    from qibotn.circuit_convertor import QiboCircuitToEinsum
    from qibotn.circuit_to_mps import QiboCircuitToMPS
    from qibotn.mps_contraction_helper import MPSContractionHelper
    from qibotn.mps_utils import apply_gate, initial, mps_site_right_swap
    from qibotn.result import TensorNetworkResult

    assert importlib.import_module("qibotn.eval") is importlib.import_module("qibotn_b200.eval")
    assert importlib.import_module("qibotn.eval_qu") is importlib.import_module("qibotn_b200.eval_qu")
    assert importlib.import_module("qibotn.backends.quimb") is importlib.import_module("qibotn_b200.backends.quimb")
    assert qibotn.MetaBackend is qibotn_b200.MetaBackend and issubclass(CuTensorNet, QibotnBackend)
    import qibotn_b200.mps as m
    import qibotn_b200.network as nw
    assert QiboCircuitToEinsum is nw.QiboCircuitToEinsum and QiboCircuitToMPS is m.QiboCircuitToMPS
    assert (MPSContractionHelper, apply_gate, initial, mps_site_right_swap) == (
        m.MPSContractionHelper, m.apply_gate, m.initial, m.mps_site_right_swap)
    assert m.__spec__.name == "qibotn_b200.mps" and TensorNetworkResult.__module__ == "qibotn_b200.result"
    assert {"init_state_tn", "dense_vector_tn_qu"} <= set(dir(importlib.import_module("qibotn.eval_qu")))
    with pytest.raises(ImportError):
        importlib.import_module("qibotn.backends.qmatchatea")           # out of scope


def test_arena_pool_carves_aligned_disjoint_views_and_only_grows():
    """executor.ArenaPool: one buffer for the hoist / slice / workspace arenas of successive runners."""
    import torch

    from qibotn_b200 import executor
    pool = executor.ArenaPool()
    a, b, c = pool.carve(torch, torch.device("cpu"), [1000, 0, 70000])
    assert (a.numel(), b.numel(), c.numel()) == (1000, 256, 70000) and pool.grown == 1
    base = pool.buf.data_ptr()
    offs = [t.data_ptr() - base for t in (a, b, c)]
    assert offs == [0, 4096, 8192] and all(o % executor.ArenaPool.ALIGN == 0 for o in offs)
    first = pool.buf
    x, y, z = pool.carve(torch, torch.device("cpu"), [4096, 4096, 4096])            # fits: same buffer
    assert pool.buf is first and pool.grown == 1 and z.data_ptr() - base == 8192
    pool.carve(torch, torch.device("cpu"), [1 << 20, 1 << 20, 1 << 20])             # larger: one reallocation
    assert pool.grown == 2 and pool.buf.numel() >= 3 << 20
    assert executor.ARENA_POOL is None                                               # opt-in
    big = executor.ArenaPool(reserve_bytes=1 << 16)
    big.carve(torch, torch.device("cpu"), [100, 100, 100])
    assert big.buf.numel() == 1 << 16 and big.grown == 1                             # first allocation = the reserve


def test_eval_qu_gate_opts_translation_and_input_checks():
    """quimb gate_opts (reference tests/test_quimb_backend.py:54-57) -> this engine's gate_algo; input validation of
    dense_vector_tn_qu happens before any device work."""
    from qibotn_b200 import eval_qu
    f = eval_qu._gate_algo
    assert f({"method": "svd", "cutoff": 1e-6, "cutoff_mode": "abs"}) == {
        "qr_method": False, "svd_method": {"partition": "UV", "abs_cutoff": 1e-6}}
    assert f({"cutoff": 1e-8, "max_bond": 32}) == {
        "qr_method": False, "svd_method": {"partition": "UV", "rel_cutoff": 1e-8, "max_extent": 32}}   # quimb's default mode
    assert f({"method": "qr"}) == {"qr_method": {}, "svd_method": False}
    assert f({"method": "svd", "absorb": "both"})["svd_method"]["rel_cutoff"] == 1e-10
    with pytest.raises(NotImplementedError):
        f({"method": "eig"})
    with pytest.raises(NotImplementedError):
        f({"cutoff_mode": "rsum2"})
    with pytest.raises(ValueError):
        f({"renorm": True})
    text = 'OPENQASM 2.0; include "qelib1.inc"; qreg q[3]; h q[0]; cx q[0],q[1];'
    with pytest.raises(ValueError):
        eval_qu.dense_vector_tn_qu(text, np.ones(4), None)                 # 2^3 amplitudes expected
    with pytest.raises(ValueError):
        eval_qu.init_state_tn(3, np.ones(4))
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(lib.QtnError):
            eval_qu.dense_vector_tn_qu(text, None, {"method": "svd"})      # no device here: no CPU fallback
