"""The CPU oracle against everything that pins it (no GPU needed):

* the closed forms the reference's own tests hold (tests/golden/closed_forms.npz);
* the survey's independent known answers (tests/golden/survey_known_answers.json);
* the committed oracle vectors (a change in the oracle's answers must show up as a diff);
* self-consistency: TN restatement == MPS restatement == independent dense simulator,
  path independence, slice-sum invariance, the reference's slice partition."""
import json
import os

import numpy as np
import pytest

from oracle import contract as oc
from oracle import mps as omps
from oracle import network as onet
from oracle import statevector as osv
from qibotn_b200 import circuits
from qibotn_b200.planner import find_path

GOLD = os.path.join(os.path.dirname(__file__), "golden")
closed = np.load(os.path.join(GOLD, "closed_forms.npz"))
vectors = np.load(os.path.join(GOLD, "oracle_vectors.npz"))
ALGO = {"qr_method": False, "svd_method": {"partition": "UV", "abs_cutoff": 1e-12}}


@pytest.mark.parametrize("n,tol", [(1, 1e-6), (2, 1e-6), (5, 1e-3), (10, 1e-3)])
def test_qft_on_random_state_is_the_inverse_fft(n, tol):
    """reference tests/test_quimb_backend.py:23-66, same sizes and tolerances (and far tighter)."""
    init, want = closed[f"qft_in_{n}"], closed[f"qft_out_{n}"]
    got = osv.simulate(circuits.qft(n, True), initial_state=init)
    assert np.allclose(got, want, atol=tol)
    assert np.abs(got - want).max() < 1e-12


@pytest.mark.parametrize("n", [1, 2, 5, 10])
def test_qft_on_zero_state_three_ways(n):
    """reference tests/test_cuquantum_cutensor_backend.py:46-142: dense TN and MPS give the
    uniform vector; every MPS bond stays 1."""
    circ = circuits.qft(n, True)
    want = closed[f"qft_zero_{n}"]
    tn = oc.contract(onet.state_vector_operands(circ)).reshape(-1)
    assert np.allclose(tn, want, rtol=1e-5, atol=1e-8)
    sites = omps.evolve(circ, ALGO)
    assert np.allclose(omps.to_dense(sites).reshape(-1), want, rtol=1e-5, atol=1e-8)
    assert all(t.shape[2] == 1 for t in sites)


@pytest.mark.parametrize("n", [2, 5, 10])
def test_default_observable_vanishes_on_qft_state(n):
    """reference tests/test_cuquantum_cutensor_backend.py:145-199 (abs_tol 1e-7)."""
    circ = circuits.qft(n, True)
    total = 0.0
    for term in onet.observable_terms(None, n):
        total += oc.contract(onet.expectation_operands(circ, onet.ham_gates(term)))
    assert abs(complex(total)) < 1e-7


def test_survey_known_answers():
    ka = json.load(open(os.path.join(GOLD, "survey_known_answers.json")))
    for n, want in ka["values"].items():
        n = int(n)
        circ = circuits.ry_rz_cnot_ring(n, 2, seed=42)
        psi = osv.simulate(circ)
        dense = osv.pauli_expectation(psi, [(0.5, [("X", i), ("Z", (i + 1) % n)]) for i in range(n)], n).real
        tn = sum(complex(oc.contract(onet.expectation_operands(circ, onet.ham_gates(t))))
                 for t in onet.observable_terms(None, n)).real
        assert abs(dense - want) < ka["abs_tol"] and abs(tn - want) < ka["abs_tol"]


def test_committed_oracle_vectors_are_reproduced():
    for n, layers in ((5, 2), (8, 3)):
        psi = osv.simulate(circuits.ry_rz_cnot_ring(n, layers, seed=42))
        assert np.abs(psi - vectors[f"ring_state_{n}_{layers}"]).max() < 1e-13
    circ = circuits.brickwork(10, 8, seed=11)
    algo = {"qr_method": False, "svd_method": {"partition": "UV", "abs_cutoff": 1e-12, "max_extent": 6}}
    got = omps.to_dense(omps.evolve(circ, algo)).reshape(-1)
    assert np.abs(got - vectors["brick10x8_mps_chi6"]).max() < 1e-9
    bonds = [t.shape[2] for t in omps.evolve(circuits.tfim_trotter(12, 4), ALGO)[:-1]]
    assert bonds == list(vectors["tfim12x4_bonds"])


@pytest.mark.parametrize("n,layers", [(3, 2), (6, 3), (9, 2)])
def test_tn_mps_and_dense_agree_on_random_circuits(n, layers):
    circ = circuits.ry_rz_cnot_ring(n, layers, seed=7)          # includes the distant (n-1, 0) CNOT
    want = osv.simulate(circ)
    tn = oc.contract(onet.state_vector_operands(circ)).reshape(-1)
    mps = omps.to_dense(omps.evolve(circ, ALGO)).reshape(-1)
    assert np.abs(tn - want).max() < 1e-12 and np.abs(mps - want).max() < 1e-10
    bits = "".join("01"[(i * 5 + 1) % 2] for i in range(n))
    amp = complex(oc.contract(onet.state_vector_operands(circ, bitstring=bits)))
    assert abs(amp - want[int(bits, 2)]) < 1e-12


def test_path_independence_and_slice_sum():
    circ = circuits.brickwork(8, 5, seed=2)
    inter = onet.state_vector_operands(circ, bitstring="0" * 8)
    ops, modes, out = oc.split_interleaved(inter)
    want = osv.simulate(circ)[0]
    vals = []
    for seed in (0, 1):
        info = find_path(modes, out, samples=2, seed=seed, min_slices=8, target_log2_width=6)
        assert info.num_slices >= 8
        total = oc.contract_sliced(inter, info.path, info.sliced_labels, range(info.num_slices))
        vals.append(complex(total))
        # the reference's partition: contiguous ranges, remainder to the low ranks (eval.py:197-205)
        parts = [onet.slice_range(info.num_slices, r, 3) for r in range(3)]
        assert [x for p in parts for x in p] == list(range(info.num_slices))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
        by_rank = sum(complex(oc.contract_sliced(inter, info.path, info.sliced_labels, p)) for p in parts if len(p))
        assert abs(by_rank - total) < 1e-12
    assert abs(vals[0] - want) < 1e-12 and abs(vals[1] - want) < 1e-12


def test_observable_spellings_agree():
    n = 4
    a = onet.observable_terms({"pauli_string_pattern": "XZ"}, n)
    assert a == [[(0, "X", 1.0), (1, "Z", 1.0), (2, "X", 1.0), (3, "Z", 1.0)]]
    b = onet.observable_terms({"terms": [{"coefficient": 2.0, "operators": [("Z", 1), ("X", 3)]}]}, n)
    assert b == [[(0, "I", 2.0), (1, "Z", 1.0), (2, "I", 1.0), (3, "X", 1.0)]]
    with pytest.raises(ValueError):
        onet.observable_terms({"terms": []}, n)
    with pytest.raises(TypeError):
        onet.observable_terms(3.0, n)


def test_truncation_rank_rules():
    s = np.array([3.0, 2.0, 1.0, 0.5, 0.2, 1e-3, 1e-4, 0.0])
    assert omps.truncation_rank(s, {"abs_cutoff": 1e-6}) == 7
    assert omps.truncation_rank(s, {"rel_cutoff": 0.1}) == 4
    assert omps.truncation_rank(s, {"max_extent": 3, "abs_cutoff": 1e-6}) == 3
    assert omps.truncation_rank(s, {"discarded_weight_cutoff": 0.01}) == 4
    assert omps.truncation_rank(s, {"abs_cutoff": 10.0}) == 1
    assert omps.truncation_rank(s, {}) == 8


def test_oracle_mps_sampler_against_the_dense_state():
    """oracle.mps.sample (the checker of the GPU sampler): per-shot |amplitude|^2 and the empirical
    distribution agree with the dense simulator."""
    n, nshots = 6, 3000
    circ = circuits.ry_rz_cnot_ring(n, 2, seed=3)
    algo = {"qr_method": False, "svd_method": {"partition": "UV", "abs_cutoff": 1e-12}}
    psi = osv.simulate(circ)
    bits, probs = omps.sample(omps.evolve(circ, algo), np.random.default_rng(0).random((n, nshots)))
    idx = (bits.astype(np.int64) << np.arange(n - 1, -1, -1)).sum(axis=1)
    assert np.abs(probs - np.abs(psi[idx]) ** 2).max() < 1e-12
    emp = np.bincount(idx, minlength=1 << n) / nshots
    assert np.abs(emp - np.abs(psi) ** 2).max() < 0.03

