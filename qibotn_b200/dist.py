"""One-process-per-GPU plumbing over ``torch.distributed``.

Replaces the MPI + ``cupy.cuda.nccl`` pair of the reference
(``/root/reference/src/qibotn/eval.py:153-235``): there is no MPI in this image,
so ranks are started by ``torchrun`` (``RANK`` / ``LOCAL_RANK`` / ``WORLD_SIZE`` /
``MASTER_*`` in the environment) and bootstrap NCCL through torch's TCP store.
Without those variables everything degrades to a single rank with no
communicator, so the ``MPI_enabled`` / ``NCCL_enabled`` runcards also work in a
plain ``python`` process.

The slice-partial sum is the only data-path collective (C2 in SURVEY.md): one
in-place all-reduce (``NCCL_enabled``) or reduce-to-root (``MPI_enabled``) of
``2 * result.size`` reals on the stream the last contraction ran on.
"""

from __future__ import annotations

import os
import pickle

import numpy as np

from .planner import compute_slices  # re-exported: eval.compute_slices

__all__ = ["initialize", "rank_size", "vote_best", "reduce_result", "compute_slices", "barrier",
           "is_initialized"]

_STATE = {"initialized": False, "rank": 0, "size": 1, "device_id": 0, "backend": None}


def partition_units(counts, rank, size, weights=None):
    """Work of a many-term observable is the flat list of (term, slice) units, term-major; rank ``rank``
    takes the contiguous range ``compute_slices(sum(counts), rank, size)`` of it (the reference's
    partition rule, eval.py:197-205, applied to the flattened list instead of once per term as
    eval.py:352-380 does).  ``counts[t]`` = number of slices of term t.  Returns ``{term: range}`` with the
    slice ids of each term this rank owns (terms it does not touch are absent).

    ``weights[t]`` = cost of ONE slice of term t: the list is then cut where the cumulative cost crosses
    multiples of total / size instead of by unit count (light-cone terms of one observable differ by
    orders of magnitude in cost: QAOA p = 4 on 40 qubits spans 10^8 .. 10^15 flops per term).  A unit
    belongs to the rank in whose share its cost midpoint falls: still contiguous, every unit owned once."""
    if weights is not None:
        total = float(sum(float(w) * int(c) for w, c in zip(weights, counts)))
        if total <= 0.0:
            return partition_units(counts, rank, size)
        share = total / size
        out, before = {}, 0.0
        for t, (c, w) in enumerate(zip(counts, weights)):
            c, w = int(c), float(w)
            lo = hi = None
            for s in range(c):
                owner = min(size - 1, int((before + (s + 0.5) * w) / share))
                if owner == rank:
                    if lo is None:
                        lo = s
                    hi = s + 1
            if lo is not None:
                out[t] = range(lo, hi)
            before += c * w
        return out
    total = int(sum(counts))
    mine = compute_slices(total, rank, size) if total else range(0)
    out, begin = {}, 0
    for t, c in enumerate(counts):
        lo, hi = max(mine.start, begin), min(mine.stop, begin + int(c))
        if lo < hi:
            out[t] = range(lo - begin, hi - begin)
        begin += int(c)
    return out


def is_initialized():
    return _STATE["initialized"]


def initialize(backend=None):
    """Equivalent of ``initialize_mpi`` + ``initialize_nccl`` (eval.py:153-167).

    Returns ``(rank, size, device_id)``; ``device_id = LOCAL_RANK % device_count``
    mirrors ``rank % getDeviceCount()`` (eval.py:158).
    """
    import torch
    import torch.distributed as td

    if _STATE["initialized"]:
        return _STATE["rank"], _STATE["size"], _STATE["device_id"]
    size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    have_cuda = torch.cuda.is_available()
    device_id = local % torch.cuda.device_count() if have_cuda else 0
    if have_cuda:
        torch.cuda.set_device(device_id)
    if size > 1 and not td.is_initialized():
        if backend is None:
            backend = "nccl" if have_cuda else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", device_id)
        td.init_process_group(backend=backend, rank=rank, world_size=size, **kw)
    if td.is_initialized():
        rank, size = td.get_rank(), td.get_world_size()
        backend = td.get_backend()
    _STATE.update(initialized=True, rank=rank, size=size, device_id=device_id, backend=backend)
    return rank, size, device_id


def rank_size():
    return _STATE["rank"], _STATE["size"]


def _comm_device():
    import torch

    if _STATE["backend"] == "nccl":
        return torch.device("cuda", _STATE["device_id"])
    return torch.device("cpu")


def barrier():
    import torch.distributed as td

    if _STATE["size"] > 1:
        td.barrier()


def vote_best(cost, payload):
    """Every rank proposes ``(cost, payload)``; all ranks get the cheapest payload.

    The reference's distributed path search (eval.py:180-194): each rank samples
    paths independently, ``allreduce(MINLOC)`` on ``opt_cost`` picks the winner
    and the winner broadcasts its ``OptimizerInfo``.
    """
    import torch
    import torch.distributed as td

    rank, size = rank_size()
    if size == 1:
        return payload, 0
    dev = _comm_device()
    costs = torch.zeros(size, dtype=torch.float64, device=dev)
    mine = torch.tensor([float(cost)], dtype=torch.float64, device=dev)
    td.all_gather_into_tensor(costs, mine)
    winner = int(torch.argmin(costs).item())            # lowest rank wins ties, like MINLOC
    blob = pickle.dumps(payload) if rank == winner else b""
    n = torch.tensor([len(blob)], dtype=torch.int64, device=dev)
    td.broadcast(n, src=winner)
    buf = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    if rank == winner:
        buf.copy_(torch.frombuffer(bytearray(blob), dtype=torch.uint8))
    td.broadcast(buf, src=winner)
    return pickle.loads(buf.cpu().numpy().tobytes()), winner


def reduce_result(result, method="NCCL", root=0):
    """Sum the slice partials (eval.py:208-235).  ``result`` is a complex torch
    tensor; it is viewed as ``2 * size`` reals (float32 for complex64, float64
    for complex128) and reduced in place on the current stream.

    ``method="NCCL"``: all-reduce (every rank ends with the sum);
    ``method="MPI"``: reduce to ``root`` only.
    """
    import torch
    import torch.distributed as td

    if method not in ("MPI", "NCCL"):
        raise ValueError(f"Unknown reduce method: {method}")
    if result.dtype not in (torch.complex64, torch.complex128):
        raise TypeError(f"Unsupported dtype for NCCL reduce: {result.dtype}")
    if _STATE["size"] == 1:
        return result
    reals = torch.view_as_real(result.contiguous())
    if reals.device != _comm_device():
        staged = reals.to(_comm_device())
    else:
        staged = reals
    if method == "NCCL":
        td.all_reduce(staged, op=td.ReduceOp.SUM)
    else:
        td.reduce(staged, dst=root, op=td.ReduceOp.SUM)
    if staged is not reals:
        reals.copy_(staged)
    return torch.view_as_complex(reals).reshape(result.shape)
