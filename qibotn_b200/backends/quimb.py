"""The ``quimb`` platform's API served by the B200 engine.

Mirrors ``/root/reference/src/qibotn/backends/quimb.py`` (constructor ``:40-56``,
``configure_tn_simulation`` ``:59-83``, ``execute_circuit`` ``:123-202``,
``exp_value_observable_symbolic`` ``:205-261``): same method names, arguments and result object, so
that ``qibo.set_backend(backend="qibotn", platform="quimb")`` and
``hamiltonian.expectation(circuit)`` of newer qibo keep working.  The arithmetic is ours:

* ``ansatz="mps"`` -> the MPS engine (``qibotn_b200.mps``) with ``max_bond_dimension`` as the bond cap
  and ``svd_cutoff`` as an absolute cutoff (quimb's ``CircuitMPS`` keeps a canonical form and cuts
  relatively; at the default near-lossless cutoff both give the same state);
* any other ansatz -> the tensor-network contraction engine.

Sampling (``nshots``): with the MPS ansatz, perfect sampling site by site on the GPU
(``MPSContractionHelper.sample``, any number of qubits); with a contraction ansatz the dense
distribution is drawn from, which is limited to 26 qubits.
"""

from __future__ import annotations

from collections import Counter

import numpy as np

from ..result import TensorNetworkResult
from .abstract import QibotnBackend
from .cutensornet import NumpyBackend, raise_error

_MAX_SAMPLING_QUBITS = 26


class QuimbBackend(QibotnBackend, NumpyBackend):
    def __init__(self, quimb_backend="numpy", contraction_optimizer="auto-hq"):
        super().__init__()
        self.name = "qibotn"
        self.platform = "quimb"
        self.backend = quimb_backend
        self.ansatz = None
        self.max_bond_dimension = None
        self.svd_cutoff = None
        self.n_most_frequent_states = None
        self.configure_tn_simulation()
        self.setup_backend_specifics(quimb_backend=quimb_backend, contractions_optimizer=contraction_optimizer)

    def configure_tn_simulation(self, ansatz="mps", max_bond_dimension=None, svd_cutoff=1e-10,
                                n_most_frequent_states=100):
        self.ansatz = ansatz
        self.max_bond_dimension = max_bond_dimension
        self.svd_cutoff = svd_cutoff
        self.n_most_frequent_states = n_most_frequent_states

    def setup_backend_specifics(self, quimb_backend="numpy", contractions_optimizer="auto-hq"):
        if quimb_backend not in ("numpy", "torch", "jax"):
            raise_error(ValueError, f"Unsupported quimb backend: {quimb_backend}")
        self.backend = quimb_backend
        self.contractions_optimizer = contractions_optimizer
        self.contraction_optimizer = contractions_optimizer     # the reference reads both spellings

    # ------------------------------------------------------------------ helpers
    def _gate_algo(self):
        svd = {"partition": "UV", "abs_cutoff": float(self.svd_cutoff or 0.0)}
        if self.max_bond_dimension:
            svd["max_extent"] = int(self.max_bond_dimension)
        return {"qr_method": False, "svd_method": svd}

    def _mps(self, circuit, initial_state=None):
        from ..mps import QiboCircuitToMPS
        return QiboCircuitToMPS(circuit, self._gate_algo(), dtype=self.dtype, initial_state=initial_state).mps_tensors

    def _dense(self, circuit, initial_state=None):
        """Dense state over ALL qubits, qubit 0 most significant."""
        if self.ansatz == "mps":
            from ..mps import MPSContractionHelper
            return MPSContractionHelper(circuit.nqubits).contract_state_vector(
                self._mps(circuit, initial_state)).reshape(-1)
        import torch

        from .. import engine
        from ..network import QiboCircuitToEinsum

        conv = QiboCircuitToEinsum(circuit, dtype=self.dtype)
        active = [int(q) for q in conv.active_qubits]
        psi = engine.contract(conv.state_vector_network(), self.dtype).reshape((2,) * len(active))
        n = circuit.nqubits
        if len(active) < n:                                   # idle qubits stay |0>
            full = torch.zeros((2,) * n, dtype=psi.dtype, device=psi.device)
            index = tuple(slice(None) if q in active else 0 for q in range(n))
            full[index] = psi
            psi = full
        return psi.reshape(-1)

    # ------------------------------------------------------------------ API
    def execute_circuit(self, circuit, initial_state=None, nshots=None, return_array=False):
        n = circuit.nqubits
        if initial_state is not None and self.ansatz != "mps":
            raise_error(ValueError, "Initial state not None supported only for MPS ansatz.")
        if initial_state is not None:
            # dense vector -> MPS on the GPU (the reference: MatrixProductState.from_dense, :162-165)
            initial_state = np.asarray(initial_state).reshape(-1)
            if initial_state.size != 1 << n:
                raise_error(ValueError, f"Initial state has {initial_state.size} amplitudes, expected 2^{n}.")
        frequencies = measured_probabilities = None
        psi = None
        if nshots and self.ansatz == "mps":
            from ..mps import MPSContractionHelper
            bits, probs = MPSContractionHelper(n).sample(self._mps(circuit, initial_state), int(nshots),
                                                         seed=getattr(self, "sampling_seed", None))
            strings = ["".join("1" if b else "0" for b in row) for row in bits]
            frequencies = Counter(strings)
            main = dict(frequencies.most_common(self.n_most_frequent_states))
            first = {}
            for st_, pr in zip(strings, probs):
                first.setdefault(st_, float(pr))
            measured_probabilities = {state: first[state] for state in main}
        elif nshots:
            if n > _MAX_SAMPLING_QUBITS:
                raise_error(NotImplementedError, f"sampling is limited to {_MAX_SAMPLING_QUBITS} qubits "
                                                 "for contraction ansaetze; use ansatz='mps'")
            import torch

            psi = self._dense(circuit)
            probs = (psi.real.double() ** 2 + psi.imag.double() ** 2)
            draws = torch.multinomial(probs / probs.sum(), int(nshots), replacement=True).cpu().numpy()
            frequencies = Counter(format(int(i), f"0{n}b") for i in draws)
            main = dict(frequencies.most_common(self.n_most_frequent_states))
            host = probs.cpu().numpy()
            measured_probabilities = {state: float(host[int(state, 2)]) for state in main}
        statevector = None
        if return_array:
            from ..result import DeviceArray
            statevector = DeviceArray(psi if psi is not None else self._dense(circuit, initial_state))
        return TensorNetworkResult(nqubits=n, backend=self, measures=frequencies,
                                   measured_probabilities=measured_probabilities, prob_type="default",
                                   statevector=statevector)

    def exp_value_observable_symbolic(self, circuit, operators_list, sites_list, coeffs_list, nqubits):
        for sites in sites_list:
            if len(sites) != len(set(sites)):
                raise_error(ValueError,
                            f"Invalid Hamiltonian term sites {sites}: repeated qubit indices are not allowed "
                            "within a single term (e.g. (0,0,0) is invalid).")
        terms = []
        for opstr, sites, coeff in zip(operators_list, sites_list, coeffs_list):
            if len(opstr) != len(sites):
                raise_error(ValueError, f"operator string {opstr!r} and sites {sites} differ in length")
            letters = [c.upper() for c in opstr]
            if any(c not in "IXYZ" for c in letters):
                raise_error(ValueError, "pauli string character must be one of I/X/Y/Z")
            terms.append((complex(coeff).real, letters, [int(s) for s in sites]))
        if self.ansatz == "mps":
            from ..mps import MPSContractionHelper
            sites_t = self._mps(circuit)
            helper = MPSContractionHelper(circuit.nqubits)
            total = 0.0
            for coeff, letters, sites in terms:
                val = helper.pauli_expectation(sites_t, dict(zip(sites, letters)))
                total += coeff * float(val.real.cpu())
            return total
        from .. import eval as ev

        obs = {"terms": [{"coefficient": c, "operators": [(l, s) for l, s in zip(letters, sites) if l != "I"]}
                         for c, letters, sites in terms]}
        return float(ev.expectation_tn(circuit, self.dtype, obs).real.get().reshape(-1)[0])
