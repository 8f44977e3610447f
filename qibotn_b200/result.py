"""Result containers.

``TensorNetworkResult`` mirrors ``/root/reference/src/qibotn/result.py:12-66``
(same fields, same ``probabilities`` / ``frequencies`` / ``state`` behaviour).
``DeviceArray`` wraps the ``torch.Tensor`` the engine produces and offers the
handful of cupy-style calls the reference's tests make on returned arrays
(``.flatten()``, ``.real``, ``.get()``, ``.item()``, indexing, ``np.asarray``).
"""

from __future__ import annotations

from copy import deepcopy
from dataclasses import dataclass
from typing import Any, Union

import numpy as np

try:  # pragma: no cover
    from qibo.config import raise_error
except Exception:
    from .qibo_compat import raise_error


class DeviceArray:
    """A GPU-resident array; data moves to the host only on ``get()`` / ``__array__``."""

    __slots__ = ("tensor",)

    def __init__(self, tensor):
        self.tensor = tensor

    # cupy-like surface -------------------------------------------------------
    def get(self):
        return self.tensor.detach().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.get()
        return a.astype(dtype) if dtype is not None else a

    def flatten(self):
        return DeviceArray(self.tensor.reshape(-1))

    def reshape(self, *shape):
        return DeviceArray(self.tensor.reshape(*shape))

    @property
    def real(self):
        t = self.tensor
        return DeviceArray(t.real if t.is_complex() else t)

    @property
    def imag(self):
        import torch

        t = self.tensor
        return DeviceArray(t.imag if t.is_complex() else torch.zeros_like(t))

    def item(self):
        return self.tensor.item()

    def __getitem__(self, idx):
        return DeviceArray(self.tensor[idx])

    def __float__(self):
        return float(self.tensor.reshape(-1)[0].real.item()) if self.tensor.is_complex() \
            else float(self.tensor.item())

    def __complex__(self):
        return complex(self.tensor.item())

    def __len__(self):
        return self.tensor.shape[0]

    @property
    def shape(self):
        return tuple(self.tensor.shape)

    @property
    def size(self):
        return self.tensor.numel()

    @property
    def ndim(self):
        return self.tensor.dim()

    @property
    def dtype(self):
        return np.dtype(str(self.tensor.dtype).replace("torch.", ""))

    @property
    def device(self):
        return self.tensor.device

    def __repr__(self):
        return f"DeviceArray({self.tensor!r})"

    def _bin(self, other, op):
        o = other.tensor if isinstance(other, DeviceArray) else other
        return DeviceArray(op(self.tensor, o))

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __abs__(self): return DeviceArray(self.tensor.abs())
    def max(self): return DeviceArray(self.tensor.max())
    def conj(self): return DeviceArray(self.tensor.conj())
    __radd__, __rmul__ = __add__, __mul__


@dataclass
class TensorNetworkResult:
    """Output of a tensor-network simulation (reference ``result.py:12-66``)."""

    nqubits: int
    backend: Any
    measures: dict
    measured_probabilities: Union[dict, np.ndarray]
    prob_type: str
    statevector: Any

    def __post_init__(self):
        if self.measured_probabilities is None:
            self.measured_probabilities = {"default": self.measured_probabilities}

    def probabilities(self):
        if self.prob_type == "U":
            probs = deepcopy(self.measured_probabilities)
            for bitstring, prob in self.measured_probabilities[self.prob_type].items():
                probs[self.prob_type][bitstring] = prob[1] - prob[0]
            return probs[self.prob_type]
        return self.measured_probabilities

    def frequencies(self):
        if self.measures is None:
            raise_error(ValueError, "To access frequencies, circuit has to be executed with a "
                                    "given number of shots != None")
        return self.measures

    def state(self):
        if self.nqubits < 20:
            return self.statevector
        raise_error(NotImplementedError,
                    "Tensor network simulation cannot be used to reconstruct statevector for >= 20 .")
