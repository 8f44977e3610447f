"""Regenerate tests/golden/*.  Run from the repo root: ``python scripts/make_golden.py``.

The reference (qiboteam/qibotn v0.0.7) ships no golden vectors and cannot be imported in this
image (qibo, quimb, cotengra, cuQuantum are absent and there is no network), so the fixtures are

* ``closed_forms.npz``   -- values with a closed form that the reference's own tests pin:
  QFT(n)|s> = ifft(s, norm="ortho") (tests/test_quimb_backend.py:23-66, random normalised |s>),
  QFT(n)|0..0> = uniform (tests/test_cuquantum_cutensor_backend.py:46-142);
* ``survey_known_answers.json`` -- sum_i 0.5 <X_i Z_{i+1}> on the RY/RZ/CNOT-ring circuit of
  tests/test_expectation.py:19-31 (seed 42, 2 layers), computed by an INDEPENDENT throw-away
  simulator during the survey -- not by this repo's oracle;
* ``oracle_vectors.npz`` -- outputs of this repo's CPU oracle (complex128) on the seeded circuits
  the GPU tests use, so that a GPU run can be checked without re-running the oracle and so that a
  change of the oracle's answers shows up as a diff.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import contract as oc, mps as omps, network as onet, statevector as osv  # noqa: E402
from qibotn_b200 import circuits  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)

closed = {}
for n in (1, 2, 5, 10):
    rng = np.random.default_rng(100 + n)
    s = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    s /= np.linalg.norm(s)
    closed[f"qft_in_{n}"] = s
    closed[f"qft_out_{n}"] = np.fft.ifft(s, norm="ortho")
    closed[f"qft_zero_{n}"] = np.full(1 << n, 2.0 ** (-n / 2), dtype=np.complex128)
np.savez_compressed(os.path.join(OUT, "closed_forms.npz"), **closed)

json.dump({"circuit": "ry_rz_cnot_ring(n, 2 layers, seed=42)", "observable": "sum_i 0.5 X_i Z_{(i+1) mod n}",
           "source": "SURVEY.md section 8c [PROBE], independent simulator", "abs_tol": 1e-11,
           "values": {"2": 0.147461548281, "5": -0.122980371916, "7": -0.215778731456}},
          open(os.path.join(OUT, "survey_known_answers.json"), "w"), indent=1)

vec = {}
for n, layers in ((5, 2), (8, 3), (12, 3)):
    c = circuits.ry_rz_cnot_ring(n, layers, seed=42)
    psi = osv.simulate(c)
    vec[f"ring_state_{n}_{layers}"] = psi
    vec[f"ring_xz_{n}_{layers}"] = np.array(osv.pauli_expectation(
        psi, [(0.5, [("X", i), ("Z", (i + 1) % n)]) for i in range(n)], n).real)
c = circuits.brickwork(10, 8, seed=11)
algo = {"qr_method": False, "svd_method": {"partition": "UV", "abs_cutoff": 1e-12, "max_extent": 6}}
vec["brick10x8_exact"] = osv.simulate(c)
vec["brick10x8_mps_chi6"] = omps.to_dense(omps.evolve(c, algo)).reshape(-1)
c = circuits.tfim_trotter(12, 4)
vec["tfim12x4_state"] = osv.simulate(c)
vec["tfim12x4_bonds"] = np.array([t.shape[2] for t in omps.evolve(
    c, {"qr_method": False, "svd_method": {"partition": "UV", "abs_cutoff": 1e-12}})[:-1]])
c = circuits.brickwork(12, 6, seed=0)
psi = osv.simulate(c)
vec["brick12x6_zall"] = np.array(osv.pauli_expectation(psi, [(1.0, [("Z", q) for q in range(12)])], 12).real)
c = circuits.sycamore_rqc(4, seed=0)             # 53 qubits, 4 cycles: amplitudes by TN contraction only
for bits in ("0" * 53, "01" * 26 + "1"):
    vec[f"syc53_d4_amp_{bits[:4]}"] = np.array(oc.contract(onet.state_vector_operands(c, bitstring=bits)))
np.savez_compressed(os.path.join(OUT, "oracle_vectors.npz"), **vec)
print("wrote", sorted(os.listdir(OUT)))
