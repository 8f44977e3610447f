"""The ``cutensornet`` platform of qibotn, served by the B200 engine.

Same class name, constructor, runcard rules, dispatch and return conventions as
``/root/reference/src/qibotn/backends/cutensornet.py:10-169`` so that
``qibo.set_backend(backend="qibotn", platform="cutensornet", runcard=...)`` and the
reference's tests work unchanged; the arithmetic is ``qibotn_b200.eval``.
"""

from __future__ import annotations

import numpy as np

from .. import __version__
from ..result import TensorNetworkResult
from .abstract import QibotnBackend

try:  # pragma: no cover - qibo is not in this image
    from qibo import hamiltonians
    from qibo.backends import NumpyBackend
    from qibo.config import raise_error
except Exception:
    from .. import hamiltonians_compat as hamiltonians
    from ..qibo_compat import raise_error

    class NumpyBackend:
        """The few attributes of ``qibo.backends.NumpyBackend`` the engine reads."""

        def __init__(self):
            self.name = "numpy"
            self.precision = "double"
            self.dtype = "complex128"
            self.device = "/GPU:0"
            self.versions = {"numpy": np.__version__}
            self.nthreads = 1
            self.supports_multigpu = False

        def set_precision(self, precision):
            if precision not in ("single", "double"):
                raise_error(ValueError, f"Unknown precision {precision}.")
            self.precision = precision
            self.dtype = "complex64" if precision == "single" else "complex128"


class CuTensorNet(QibotnBackend, NumpyBackend):
    """qibotn backend whose contractions run on hand-written sm_100a kernels."""

    def __init__(self, runcard=None):
        super().__init__()
        self.name = "qibotn"
        self.platform = "cutensornet"
        self.versions["qibotn_b200"] = __version__
        self.supports_multigpu = True
        self.configure_tn_simulation(runcard)

    def configure_tn_simulation(self, runcard):
        self.rank = None
        if runcard is not None:
            self.MPI_enabled = runcard.get("MPI_enabled", False)
            self.NCCL_enabled = runcard.get("NCCL_enabled", False)

            value = runcard.get("expectation_enabled")
            if value is True:
                self.expectation_enabled, self.observable = True, None
            elif value is False:
                self.expectation_enabled = False
            elif isinstance(value, dict) or isinstance(value, hamiltonians.SymbolicHamiltonian):
                self.expectation_enabled, self.observable = True, value
            else:
                raise TypeError("expectation_enabled has an unexpected type")

            value = runcard.get("MPS_enabled")
            if value is True:
                self.MPS_enabled = True
                self.gate_algo = {"qr_method": False,
                                  "svd_method": {"partition": "UV", "abs_cutoff": 1e-12}}
            elif value is False:
                self.MPS_enabled = False
            elif isinstance(value, dict):
                self.MPS_enabled, self.gate_algo = True, value
            else:
                raise TypeError("MPS_enabled has an unexpected type")
        else:
            self.MPI_enabled = False
            self.MPS_enabled = False
            self.NCCL_enabled = False
            self.expectation_enabled = False

    def execute_circuit(self, circuit, initial_state=None, nshots=None, return_array=False):
        """Run ``circuit`` with the configured computation; see the reference's
        seven-way dispatch at ``backends/cutensornet.py:90-152``."""
        from .. import eval as ev

        if initial_state is not None:
            raise_error(NotImplementedError, "QiboTN cannot support initial state.")
        flags = (bool(self.MPI_enabled), bool(self.MPS_enabled), bool(self.NCCL_enabled),
                 bool(self.expectation_enabled))
        if flags == (False, False, False, False):
            state = ev.dense_vector_tn(circuit, self.dtype)
        elif flags == (False, True, False, False):
            state = ev.dense_vector_mps(circuit, self.gate_algo, self.dtype)
        elif flags == (True, False, False, False):
            state, self.rank = ev.dense_vector_tn_MPI(circuit, self.dtype, 32)
        elif flags == (False, False, True, False):
            state, self.rank = ev.dense_vector_tn_nccl(circuit, self.dtype, 32)
        elif flags == (False, False, False, True):
            state = ev.expectation_tn(circuit, self.dtype, self.observable)
        elif flags == (True, False, False, True):
            state, self.rank = ev.expectation_tn_MPI(circuit, self.dtype, self.observable, 32)
        elif flags == (False, False, True, True):
            state, self.rank = ev.expectation_tn_nccl(circuit, self.dtype, self.observable, 32)
        else:
            raise_error(NotImplementedError, "Compute type not supported.")
        if self.rank is not None and self.rank > 0:
            state = np.array(0)

        if self.expectation_enabled:
            return state.flatten().real
        statevector = state.flatten() if return_array else state
        return TensorNetworkResult(nqubits=circuit.nqubits, backend=self, measures=None,
                                   measured_probabilities=None, prob_type=None,
                                   statevector=statevector)
