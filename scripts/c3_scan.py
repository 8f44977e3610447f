"""Scan of the TFIM Trotter angles for BASELINE config 3 (64 sites, 20 steps, chi <= 256) on the GPU engine:
which (J dt, h dt) make the bulk bonds reach 256 while the truncated state keeps its norm ("choose
J dt and h dt large enough that bulk bonds actually reach 256 ... record the realised bond profile")."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from qibotn_b200 import circuits, mps as gmps  # noqa: E402

algo = {"qr_method": False, "svd_method": {"partition": "UV", "abs_cutoff": 1e-12, "max_extent": 256}}
grid = [(float(j), float(h)) for j, h in (x.split(",") for x in sys.argv[1:])] or \
    [(j, h) for j in (0.10, 0.15, 0.20, 0.25, 0.30) for h in (0.2, 0.3, 0.4, 0.6)]
for j_dt, h_dt in grid:
    circ = circuits.tfim_trotter(64, 20, j_dt=j_dt, h_dt=h_dt)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    conv = gmps.QiboCircuitToMPS(circ, algo, dtype="complex128")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    helper = gmps.MPSContractionHelper(64)
    norm = float(helper.contract_norm(conv.mps_tensors).cpu())
    bonds = [int(t.shape[2]) for t in conv.mps_tensors[:-1]]
    print(json.dumps({"j_dt": j_dt, "h_dt": h_dt, "norm": norm, "max_bond": max(bonds), "bonds_at_256": sum(b == 256 for b in bonds),
                      "mid_bond": bonds[31], "seconds": dt, "gates_per_s": len(circ.queue) / dt}), flush=True)
