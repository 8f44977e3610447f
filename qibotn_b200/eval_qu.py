"""The legacy quimb executor's call, served by the B200 engine (BASELINE config 1).

Mirrors ``/root/reference/src/qibotn/eval_qu.py``: ``init_state_tn`` (``:5-18``) and
``dense_vector_tn_qu(qasm, initial_state, mps_opts, backend)`` (``:21-46``) -- same names and argument
meaning, so that ``/root/reference/tests/test_quimb_backend.py:23-66`` reads the same against this module.

* ``mps_opts`` falsy: the circuit (with the dense initial state as one extra leaf) is contracted by the
  tensor-network engine -- where the reference builds ``quimb.tensor.Circuit`` and calls ``to_dense``;
* ``mps_opts`` a dict of quimb ``gate_opts`` (``method``, ``cutoff``, ``cutoff_mode``, ``max_bond``): the MPS gate
  loop of ``qibotn_b200.mps`` with the equivalent truncation -- where the reference builds ``CircuitMPS``.

There is no CPU path: both routes run the CUDA kernels and raise without a device.
"""

from __future__ import annotations

import numpy as np

from . import qasm as _qasm

__all__ = ["init_state_tn", "dense_vector_tn_qu"]


def init_state_tn(nqubits, init_state_sv, dtype="complex128"):
    """Dense vector -> list of ``nqubits`` MPS site tensors (l, 2, r) on the GPU (eval_qu.py:5-18)."""
    from .mps import from_dense

    state = np.asarray(init_state_sv).reshape(-1)
    if state.size != 1 << int(nqubits):
        raise ValueError(f"initial state has {state.size} amplitudes, expected 2^{nqubits}")
    return from_dense(state, dtype)


def _gate_algo(mps_opts):
    """quimb ``gate_opts`` -> this engine's gate_algo (cuTensorNet SVDMethod vocabulary).
    quimb's ``cutoff_mode``: "abs" drops s < cutoff; "rel" drops s < cutoff * s_max (its default is "rel")."""
    opts = dict(mps_opts)
    method = opts.pop("method", "svd")
    if method not in ("svd", "qr"):
        raise NotImplementedError(f"gate_opts method {method!r}: this engine splits by 'svd' or 'qr'")
    cutoff = float(opts.pop("cutoff", 1e-10))
    mode = opts.pop("cutoff_mode", "rel")
    max_bond = opts.pop("max_bond", None)
    opts.pop("absorb", None)                        # where the singular values go does not change the state
    if opts:
        raise ValueError(f"unknown gate_opts key(s): {sorted(opts)}")
    if method == "qr":
        return {"qr_method": {}, "svd_method": False}
    svd = {"partition": "UV"}
    if mode in ("abs", 1):
        svd["abs_cutoff"] = cutoff
    elif mode in ("rel", 2):
        svd["rel_cutoff"] = cutoff
    else:
        raise NotImplementedError(f"cutoff_mode {mode!r}: 'abs' and 'rel' are supported")
    if max_bond and int(max_bond) > 0:
        svd["max_extent"] = int(max_bond)
    return {"qr_method": False, "svd_method": svd}


def dense_vector_tn_qu(qasm: str, initial_state, mps_opts, backend="numpy", dtype="complex128"):
    """Amplitudes of the final state of the QASM program (eval_qu.py:21-46): an array of 2^n values, qubit 0 most
    significant, on the host for ``backend="numpy"`` (the reference's default) and as a CUDA torch tensor for
    ``backend="torch"``."""
    circuit = _qasm.from_openqasm2_str(qasm)
    n = int(circuit.nqubits)
    if initial_state is not None:
        initial_state = np.asarray(initial_state).reshape(-1)
        if initial_state.size != 1 << n:
            raise ValueError(f"initial state has {initial_state.size} amplitudes, expected 2^{n}")
    if mps_opts:
        from .mps import MPSContractionHelper, QiboCircuitToMPS

        sites = QiboCircuitToMPS(circuit, _gate_algo(mps_opts), dtype=dtype, initial_state=initial_state).mps_tensors
        psi = MPSContractionHelper(n).contract_state_vector(sites).reshape(-1)
    else:
        import torch

        from . import engine
        from .network import QiboCircuitToEinsum

        conv = QiboCircuitToEinsum(circuit, dtype=dtype)
        if initial_state is None:
            active = [int(q) for q in conv.active_qubits]
            psi = engine.contract(conv.state_vector_network(), dtype).reshape((2,) * len(active))
            if len(active) < n:                             # idle qubits stay |0>
                full = torch.zeros((2,) * n, dtype=psi.dtype, device=psi.device)
                full[tuple(slice(None) if q in active else 0 for q in range(n))] = psi
                psi = full
            psi = psi.reshape(-1)
        else:
            psi = engine.contract(conv.state_vector_network(initial_state), dtype).reshape(-1)
    if backend == "numpy":
        return psi.cpu().numpy()
    if backend == "torch":
        return psi
    raise ValueError(f"Unsupported backend: {backend}")
