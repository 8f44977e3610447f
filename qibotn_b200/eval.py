"""Execution drivers: the seven entry points of
``/root/reference/src/qibotn/eval.py`` (``dense_vector_tn`` :301,
``dense_vector_tn_MPI`` :238, ``dense_vector_tn_nccl`` :269, ``expectation_tn``
:455, ``expectation_tn_MPI`` :385, ``expectation_tn_nccl`` :316,
``dense_vector_mps`` :480) with the same names, arguments and return shapes,
running on the B200 engine instead of cuQuantum.

Differences from the reference, all in how work is scheduled, not in results:

* ranks rebuild operands deterministically from the circuit instead of
  receiving pickled arrays from rank 0 (C4);
* each rank searches paths with its own seed and the cheapest plan is adopted
  (same vote as ``compute_optimal_path``, eval.py:180-194) -- once per network
  *structure*, not once per Hamiltonian term;
* the distributed expectation does ONE collective over all terms
  (a ``2 * n_terms`` real vector) instead of one per term (eval.py:352-380).

Returned arrays are ``DeviceArray`` (a thin ``torch.Tensor`` wrapper offering the
cupy calls the reference's tests use: ``.flatten()``, ``.real``, ``.get()``,
``.item()``); results stay on the GPU until asked for.
"""

from __future__ import annotations

import numpy as np

from . import dist, engine
from .network import QiboCircuitToEinsum, TensorNetwork
from .observables import (build_observable, check_observable, create_hamiltonian_from_dict,
                          extract_gates_and_qubits, get_ham_gates)
from .planner import compute_slices
from .result import DeviceArray

__all__ = [
    "dense_vector_tn", "dense_vector_tn_MPI", "dense_vector_tn_nccl", "expectation_tn",
    "expectation_tn_MPI", "expectation_tn_nccl", "dense_vector_mps", "amplitude_tn", "expectation_from_state",
    "check_observable", "build_observable", "create_hamiltonian_from_dict", "get_ham_gates",
    "extract_gates_and_qubits", "compute_slices", "initialize_mpi", "initialize_nccl",
    "compute_optimal_path", "reduce_result",
]


# ---------------------------------------------------------------- distributed bits
def initialize_mpi():
    """``(comm, rank, size, device_id)`` as in eval.py:153-160; ``comm`` is the dist module."""
    rank, size, device_id = dist.initialize()
    return dist, rank, size, device_id


def initialize_nccl(comm_mpi=None, rank=None, size=None):
    """NCCL is bootstrapped by ``dist.initialize`` (torch TCP store instead of an
    MPI broadcast of the unique id, eval.py:163-167)."""
    dist.initialize()
    return dist


def compute_optimal_path(net: TensorNetwork, n_samples, size, dtype):
    """Per-rank path search with forced slicing, then the cost vote (eval.py:180-194)."""
    rank, _ = dist.rank_size()
    info, _ = engine.plan(net, dtype, samples=n_samples, seed=rank, min_slices=max(32, size))
    best, _winner = dist.vote_best(info.opt_cost, info)
    return best


def reduce_result(result, comm=None, method="MPI", root=0):
    return dist.reduce_result(result, method=method, root=root)


def _distributed_contract(net, dtype, n_samples, method):
    _, rank, size, _ = initialize_mpi()
    info = compute_optimal_path(net, n_samples, size, dtype)
    slices = compute_slices(info.num_slices, rank, size)
    partial = engine.contract(net, dtype, info=info, slice_ids=slices)
    return dist.reduce_result(partial, method=method), rank


# ------------------------------------------------------------------- dense vector
def dense_vector_tn(qibo_circ, datatype):
    """Dense state vector over the active qubits on one GPU (eval.py:301-313)."""
    conv = QiboCircuitToEinsum(qibo_circ, dtype=datatype)
    return DeviceArray(engine.contract(conv.state_vector_network(), datatype, samples=PLAN_SAMPLES))


def dense_vector_tn_MPI(qibo_circ, datatype, n_samples=8):
    """Sliced, reduce-to-root variant (eval.py:238-266); returns ``(state, rank)``."""
    conv = QiboCircuitToEinsum(qibo_circ, dtype=datatype)
    state, rank = _distributed_contract(conv.state_vector_network(), datatype, n_samples, "MPI")
    return DeviceArray(state), rank


def dense_vector_tn_nccl(qibo_circ, datatype, n_samples=8):
    """Sliced, NCCL all-reduce variant (eval.py:269-298); returns ``(state, rank)``."""
    conv = QiboCircuitToEinsum(qibo_circ, dtype=datatype)
    state, rank = _distributed_contract(conv.state_vector_network(), datatype, n_samples, "NCCL")
    return DeviceArray(state), rank


def amplitude_tn(qibo_circ, datatype, bitstring, n_samples=32, distributed=False, min_slices=1):
    """<bitstring|U|0..0> as one (optionally slice-parallel) contraction.  Not in the
    reference; needed for BASELINE config 4."""
    conv = QiboCircuitToEinsum(qibo_circ, dtype=datatype)
    net = conv.amplitude_network(bitstring)
    if distributed:
        amp, rank = _distributed_contract(net, datatype, n_samples, "NCCL")
        return DeviceArray(amp), rank
    return DeviceArray(engine.contract(net, datatype, samples=n_samples, min_slices=min_slices))


# -------------------------------------------------------------------- expectation
# Diagonal gates (RZ, CZ, CU1, RZZ, ...) and diagonal Pauli factors enter the expectation networks as
# hyper-indexed diagonals instead of full rank-2k tensors (network.QiboCircuitToEinsum._thread): same
# value, much smaller treewidth when the entangling gates are diagonal (QAOA: BASELINE config 5).
HYPER_DIAGONAL = True


def _term_networks(qibo_circ, datatype, observable):
    conv = QiboCircuitToEinsum(qibo_circ, dtype=datatype)
    ham = check_observable(observable, qibo_circ.nqubits)
    for term in extract_gates_and_qubits(ham):
        yield conv.expectation_network(get_ham_gates(term, dtype=datatype), hyper=HYPER_DIAGONAL)


# Planner effort of the single-GPU drivers (seeds tried per network structure; the distributed
# drivers receive n_samples from the backend, 32, like the reference's cutensornet.py:110-147).
PLAN_SAMPLES = 32

# How expectation_tn evaluates a Hamiltonian: "network" = one closed <0|U^dag P U|0> contraction
# per term, as the reference does; "state" = contract U|0> once to a dense vector and run the fused
# Pauli kernel (one streaming pass per term); "auto" = "state" when the active qubits fit easily
# and there are at least two terms.  Same value either way (tests/test_gpu_tn.py).
EXPECTATION_ROUTE = "auto"
_STATE_ROUTE_AUTO_QUBITS = 26
_STATE_ROUTE_MAX_QUBITS = 31


def expectation_from_state(qibo_circ, datatype, terms):
    """sum_t coeff_t <psi|P_t|psi> with |psi> = U|0..0> dense over the active qubits and
    ``qtn_pauli_expectation`` doing every term in one pass each.  ``terms`` as produced by
    ``extract_gates_and_qubits``.  Inactive qubits stay |0>: I and Z act as 1, X and Y give 0."""
    import ctypes as C

    import torch

    from . import lib

    conv = QiboCircuitToEinsum(qibo_circ, dtype=datatype)
    active = [int(q) for q in conv.active_qubits]
    n = len(active)
    if n == 0 or n > _STATE_ROUTE_MAX_QUBITS:
        raise ValueError("state route needs between 1 and %d active qubits" % _STATE_ROUTE_MAX_QUBITS)
    bit = {q: n - 1 - k for k, q in enumerate(active)}          # qubit -> address bit of the dense state
    state = engine.contract(conv.state_vector_network(), datatype, samples=PLAN_SAMPLES).reshape(-1)
    xm, zm, ny, coeffs = [], [], [], []
    for term in terms:
        coeff, fx, fz, cy, dead = 1.0, 0, 0, 0, False
        for qubit, letter, c in term:
            coeff = coeff * c
            if letter not in ("I", "X", "Y", "Z"):
                raise ValueError("pauli string character must be one of I/X/Y/Z")
            if qubit not in bit:
                dead = dead or letter in ("X", "Y")
                continue
            if letter in ("X", "Y"):
                fx |= 1 << bit[qubit]
            if letter in ("Z", "Y"):
                fz |= 1 << bit[qubit]
            cy += letter == "Y"
        if dead:
            continue
        xm.append(fx); zm.append(fz); ny.append(cy); coeffs.append(complex(coeff))
    cdtype = torch.complex64 if str(datatype) == "complex64" else torch.complex128
    if not coeffs:
        return DeviceArray(torch.zeros(1, dtype=cdtype, device=state.device))
    dev = state.device
    h_x = (C.c_uint64 * len(xm))(*xm)
    h_z = (C.c_uint64 * len(zm))(*zm)
    h_y = (C.c_int32 * len(ny))(*ny)
    out = torch.empty(len(coeffs), dtype=torch.complex128, device=dev)
    # terms sharing an xmask (e.g. every ZZ term of a MaxCut Hamiltonian) are fused 16 to a pass over psi
    lib.check(lib.load().qtn_pauli_expectation_host(lib.dtype_code(str(datatype)), n, state.data_ptr(), len(coeffs),
                                                    h_x, h_z, h_y, out.data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream))
    total = (out * torch.tensor(coeffs, dtype=torch.complex128, device=dev)).sum().reshape(1)
    return DeviceArray(total.to(cdtype))


def expectation_tn(qibo_circ, datatype, observable):
    """sum over Hamiltonian terms of <0|U^dag P U|0> (eval.py:455-477): one contraction per term,
    or -- see EXPECTATION_ROUTE -- one dense contraction plus the fused Pauli kernel."""
    import torch

    if EXPECTATION_ROUTE in ("auto", "state"):
        ham = check_observable(observable, qibo_circ.nqubits)
        terms = extract_gates_and_qubits(ham)
        if not terms:
            raise ValueError("No valid Hamiltonian terms were added.")
        n_active = len(QiboCircuitToEinsum(qibo_circ, dtype=datatype).active_qubits)
        limit = _STATE_ROUTE_MAX_QUBITS if EXPECTATION_ROUTE == "state" else _STATE_ROUTE_AUTO_QUBITS
        if 1 <= n_active <= limit and (EXPECTATION_ROUTE == "state" or len(terms) >= 2):
            return expectation_from_state(qibo_circ, datatype, terms)

    total = None
    for net in _term_networks(qibo_circ, datatype, observable):
        val = engine.contract(net, datatype, samples=PLAN_SAMPLES).reshape(1)
        total = val.clone() if total is None else total + val
    if total is None:
        raise ValueError("No valid Hamiltonian terms were added.")
    return DeviceArray(total)


def unit_seconds(nets, infos, datatype):
    """Estimated seconds of ONE (term, slice) unit of every term (planner.estimate_seconds; the slice-independent
    part of a term is charged to its slices in equal parts: a rank that owns any slice of a term runs it once)."""
    from .planner import SymbolicNetwork, estimate_seconds

    out = []
    for net, info in zip(nets, infos):
        sn = SymbolicNetwork(net.modes, net.out)
        per_slice, hoisted = estimate_seconds(sn, info.path, sn.to_mask(info.sliced_labels), datatype)
        out.append(per_slice + hoisted / max(1, info.num_slices))
    return out


def _expectation_distributed(qibo_circ, datatype, observable, n_samples, method):
    """Terms AND slices are sharded: the unit of work is (term, slice), the flat list is cut into ``size``
    contiguous ranges (``dist.partition_units``), each rank contracts its units and ONE all-reduce of the
    per-term partials ends the job (the reference repeats plan-vote + reduce per term, eval.py:352-380).
    With fewer terms than ranks every term is sliced into >= max(32, size) slices as the reference does
    (eval.py:186); with many terms the slicing is left to the width limit alone, so that small light-cone
    networks are not cut into 32 pieces just to be spread."""
    import torch

    _, rank, size, _ = initialize_mpi()
    nets = list(_term_networks(qibo_circ, datatype, observable))
    if not nets:
        raise ValueError("No valid Hamiltonian terms were added.")
    forced = max(32, size) if len(nets) < size else 1
    plans, infos = {}, []
    for net in nets:
        key = (tuple(tuple(m) for m in net.modes),)
        if key not in plans:                       # one vote per network structure
            rk, _ = dist.rank_size()
            info = engine.find_info(net, datatype, samples=n_samples, seed=rk, min_slices=forced)
            plans[key], _winner = dist.vote_best(info.opt_cost, info)   # lowered by its owners only
        infos.append(plans[key])
    # cut by estimated device time, not by unit count: the light cones of one observable differ by orders of
    # magnitude in cost and in how memory-bound their nodes are
    mine = dist.partition_units([info.num_slices for info in infos], rank, size,
                                weights=unit_seconds(nets, infos, datatype))
    cdtype = torch.complex64 if str(datatype) == "complex64" else torch.complex128
    partials = torch.zeros(len(nets), dtype=cdtype, device="cuda")
    for t, slices in mine.items():
        partials[t] = engine.contract(nets[t], datatype, info=infos[t], slice_ids=slices).reshape(())
    partials = dist.reduce_result(partials, method=method)        # one collective for all terms
    return DeviceArray(partials.sum().reshape(1)), rank


def expectation_tn_nccl(qibo_circ, datatype, observable, n_samples=8):
    """Distributed expectation, NCCL all-reduce (eval.py:316-382); ``(exp, rank)``."""
    return _expectation_distributed(qibo_circ, datatype, observable, n_samples, "NCCL")


def expectation_tn_MPI(qibo_circ, datatype, observable, n_samples=8):
    """Distributed expectation, reduce to root (eval.py:385-452); ``(exp, rank)``."""
    return _expectation_distributed(qibo_circ, datatype, observable, n_samples, "MPI")


# ---------------------------------------------------------------------------- MPS
def dense_vector_mps(qibo_circ, gate_algo, datatype):
    """MPS evolution then dense read-out (eval.py:480-497)."""
    from .mps import QiboCircuitToMPS, MPSContractionHelper

    conv = QiboCircuitToMPS(qibo_circ, gate_algo, dtype=datatype)
    helper = MPSContractionHelper(conv.num_qubits)
    return DeviceArray(helper.contract_state_vector(conv.mps_tensors))
