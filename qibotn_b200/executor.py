"""Pairwise contraction executor: lowers (network, path, slices) to a list of
launch-ready kernel plans plus a static memory arena, then runs it on the GPU.

This is the part of ``cuquantum.tensornet.Network.contract(slices=...)`` the
reference relies on (``/root/reference/src/qibotn/eval.py:265,297,313,476``):

* every tree node is ONE fused permute+contract launch (``qtn_pair_run``); no
  transposes are materialised -- the C layout of each intermediate is chosen by
  the node that produces it so that its own stores are contiguous;
* sliced labels never cause copies: a leaf that carries a sliced label keeps its
  full buffer and the kernel adds ``bit(slice_id) << address_bit`` to its base;
  the slice id lives in device memory, so the per-slice launch sequence is a
  CUDA graph replayed once per slice;
* nodes that depend on no sliced label are executed once (hoisted);
* all leaves travel to the device in one packed host->device copy (K10).

There is no CPU path: without the CUDA library or a GPU this module raises.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import lib
from .planner import PathInfo, SymbolicNetwork, _walk


@dataclass
class _Tensor:
    labels: list            # labels present (sliced ones removed)
    pos: list               # address bit of each label
    slices: list            # [(slice bit index, address bit)] for sliced labels (leaves only)
    where: str              # "leaf" | "hoist" | "slice" | "out"
    offset: int = 0         # byte offset in its arena
    nbytes: int = 0
    depends: bool = False


@dataclass
class _Node:
    a: int
    b: int
    c: int
    plan: lib.PairPlan
    depends: bool


@dataclass
class LoweredContraction:
    """Everything needed to run a contraction; reusable across calls with new leaf data."""

    dtype: str
    info: PathInfo
    tensors: dict = field(default_factory=dict)
    nodes: list = field(default_factory=list)
    leaf_bytes: int = 0
    hoist_bytes: int = 0
    slice_bytes: int = 0
    workspace_bytes: int = 0
    out_rank: int = 0
    n_leaves: int = 0
    root: int = -1
    flops_hoisted: float = 0.0
    flops_per_slice: float = 0.0
    bytes_hoisted: float = 0.0
    bytes_per_slice: float = 0.0
    out_labels: list = field(default_factory=list)

    @property
    def launches_per_slice(self):
        return sum(1 + (1 if (n.plan.split_bits or n.plan.kernel == lib.KERNEL_REDUCE) else 0)
                   for n in self.nodes if n.depends)

    @property
    def launches_hoisted(self):
        return sum(1 + (1 if (n.plan.split_bits or n.plan.kernel == lib.KERNEL_REDUCE) else 0)
                   for n in self.nodes if not n.depends)


class _Arena:
    """First-fit offset allocator evaluated at plan time (256-byte granules)."""

    def __init__(self):
        self.free = []          # sorted list of (offset, size)
        self.top = 0
        self.peak = 0           # high-water mark: the buffer size to allocate

    def alloc(self, size):
        size = (size + 255) & ~255
        for i, (off, sz) in enumerate(self.free):
            if sz >= size:
                if sz == size:
                    self.free.pop(i)
                else:
                    self.free[i] = (off + size, sz - size)
                return off
        off = self.top
        self.top += size
        self.peak = max(self.peak, self.top)
        return off

    def release(self, off, size):
        size = (size + 255) & ~255
        self.free.append((off, size))
        self.free.sort()
        merged = []
        for o, s in self.free:
            if merged and merged[-1][0] + merged[-1][1] == o:
                merged[-1] = (merged[-1][0], merged[-1][1] + s)
            else:
                merged.append((o, s))
        if merged and merged[-1][0] + merged[-1][1] == self.top:
            self.top = merged[-1][0]
            merged.pop()
        self.free = merged


def lower(modes, out, info: PathInfo, dtype="complex64", sm_count=148) -> LoweredContraction:
    """Turn a path into kernel plans and arena offsets.  Pure host work (no GPU)."""
    esz = 8 if dtype == "complex64" else 16
    net = SymbolicNetwork(modes, out)
    sliced_labels = list(info.sliced_labels)
    nsl = len(sliced_labels)
    # most significant slice bit first, as in oracle.contract.contract_ssa
    slice_bit = {lab: nsl - 1 - k for k, lab in enumerate(sliced_labels)}
    sliced_mask = net.to_mask(sliced_labels) if sliced_labels else 0
    low = LoweredContraction(dtype=dtype, info=info, n_leaves=len(modes), out_rank=len(out),
                             out_labels=list(out))

    leaf_off = 0
    for i, m in enumerate(modes):
        r = len(m)
        labels, pos, slices = [], [], []
        for ax, lab in enumerate(m):
            p = r - 1 - ax                      # dense row-major, first mode slowest
            if lab in slice_bit:
                slices.append((slice_bit[lab], p))
            else:
                labels.append(lab)
                pos.append(p)
        if len(set(m)) != len(m):
            raise ValueError("repeated label inside one tensor is not supported")
        nbytes = esz << r
        low.tensors[i] = _Tensor(labels, pos, slices, "leaf", leaf_off, nbytes, bool(slices))
        leaf_off += (nbytes + 255) & ~255
    low.leaf_bytes = leaf_off

    if len(modes) == 1:
        t = low.tensors[0]
        if t.slices:
            raise ValueError("cannot slice a single-tensor network")
        low.root = 0
        return low

    steps = list(_walk(net, info.path, sliced_mask))
    consumers = {}
    for k, (i, j) in enumerate(info.path):
        consumers[i] = k
        consumers[j] = k
    root_id = steps[-1][0]
    hoist, per_slice = _Arena(), _Arena()
    for k, ((i, j), (cid, _a, _b, cmask, dep)) in enumerate(zip(info.path, steps)):
        ta, tb = low.tensors[i], low.tensors[j]
        c_labels = net.to_labels(cmask)
        is_root = cid == root_id
        if is_root:
            live_out = [lab for lab in out if lab not in slice_bit]
            if sorted(live_out) != sorted(c_labels):
                raise ValueError("path does not end in the requested output labels")
            rc = len(live_out)
            c_spec = [(lab, rc - 1 - ax) for ax, lab in enumerate(live_out)]
        else:
            c_spec = c_labels
        desc = lib.make_pair_desc(
            dtype, list(zip(ta.labels, ta.pos)), list(zip(tb.labels, tb.pos)), c_spec,
            accumulate=bool(is_root and dep), slice_a=ta.slices, slice_b=tb.slices)
        plan = lib.pair_plan(desc, sm_count)
        rc = len(c_labels)
        labels_c = [desc.label_c[x] for x in range(rc)]
        pos_c = [desc.pos_c[x] for x in range(rc)]
        nbytes = esz << rc
        where = "out" if is_root else ("slice" if dep else "hoist")
        tc = _Tensor(labels_c, pos_c, [], where, 0, nbytes, dep)
        if where == "hoist":
            tc.offset = hoist.alloc(nbytes)
        elif where == "slice":
            tc.offset = per_slice.alloc(nbytes)
        low.tensors[cid] = tc
        low.nodes.append(_Node(i, j, cid, plan, dep))
        low.workspace_bytes = max(low.workspace_bytes, plan.workspace_bytes)
        if dep:
            low.flops_per_slice += plan.flops
            low.bytes_per_slice += plan.bytes
        else:
            low.flops_hoisted += plan.flops
            low.bytes_hoisted += plan.bytes
        # release operands whose last use this was; hoisted tensors read by a per-slice node
        # must survive the whole slice loop, so they are never released
        for t_id in (i, j):
            t = low.tensors[t_id]
            if t.where == "slice":
                per_slice.release(t.offset, t.nbytes)
            elif t.where == "hoist" and not dep:
                hoist.release(t.offset, t.nbytes)
    low.hoist_bytes, low.slice_bytes = _peak(hoist), _peak(per_slice)
    low.root = root_id
    return low


def _peak(arena):
    return arena.peak


class ArenaPool:
    """ONE grow-only device buffer from which successive runners carve their hoist / slice / workspace arenas.

    An expectation value over many differently shaped terms (config 5: 60 light cones, arenas of up to 89 GiB)
    otherwise pays a cudaMalloc and a cudaFree of tens of GiB per term.  Only one runner that uses the pool may be
    live at a time: the caller drops a runner (``engine.clear_caches()``) before the next one is built.  Opt-in:
    ``executor.ARENA_POOL = executor.ArenaPool()``."""

    ALIGN = 4096

    def __init__(self, reserve_bytes=0):
        self.buf = None
        self.grown = 0
        self.reserve_bytes = int(reserve_bytes)            # size of the first allocation: growing costs a free + a malloc

    def carve(self, torch, device, sizes):
        offsets, total = [], 0
        for nbytes in sizes:
            offsets.append(total)
            total += (max(int(nbytes), 256) + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        if self.buf is None or self.buf.numel() < total:
            self.buf = None
            torch.cuda.empty_cache()                       # the old block goes back before the larger one is asked for
            self.buf = torch.empty(max(total, self.reserve_bytes), dtype=torch.uint8, device=device)
            self.grown += 1
        return [self.buf[o:o + max(int(nbytes), 256)] for o, nbytes in zip(offsets, sizes)]


ARENA_POOL = None


class ContractionRunner:
    """Owns the device buffers of one lowered contraction and runs it."""

    def __init__(self, lowered: LoweredContraction, device=None, use_graph=True):
        import torch

        if not torch.cuda.is_available():
            raise lib.QtnError("qibotn_b200 needs a CUDA device; there is no CPU fallback")
        self.torch = torch
        self.low = lowered
        self.lib = lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.cdtype = torch.complex64 if lowered.dtype == "complex64" else torch.complex128
        self.np_dtype = np.complex64 if lowered.dtype == "complex64" else np.complex128
        u8 = torch.uint8
        self.leaf_buf = torch.empty(max(lowered.leaf_bytes, 256), dtype=u8, device=self.device)
        self.host_buf = torch.empty(max(lowered.leaf_bytes, 256), dtype=u8, pin_memory=True)
        if ARENA_POOL is not None:
            self.hoist_buf, self.slice_buf, self.ws = ARENA_POOL.carve(
                torch, self.device, [lowered.hoist_bytes, lowered.slice_bytes, lowered.workspace_bytes])
        else:
            self.hoist_buf = torch.empty(max(lowered.hoist_bytes, 256), dtype=u8, device=self.device)
            self.slice_buf = torch.empty(max(lowered.slice_bytes, 256), dtype=u8, device=self.device)
            self.ws = torch.empty(max(lowered.workspace_bytes, 256), dtype=u8, device=self.device)
        self.out = torch.zeros(1 << self._live_out_rank(), dtype=self.cdtype, device=self.device)
        self.slice_id = torch.zeros(1, dtype=torch.int32, device=self.device)
        # Magnitude words (qtn_pair_run_scaled): one device word per tensor that an fp16-split plan reads,
        # bounding max(|re|, |im|) of that tensor.  Leaves are measured after every upload, intermediates
        # by the kernel that produces them; the words of per-slice tensors are cleared at the start of
        # every slice (inside the captured graph), so a slice's result does not depend on its predecessors.
        self._amax_slot = {}
        needs = set()
        for node in lowered.nodes:
            if node.plan.kernel == lib.KERNEL_TC and node.plan.tc_kind == 1:
                needs.update((node.a, node.b))
        groups = {"leaf": [], "hoist": [], "slice": [], "out": []}
        for tid in sorted(needs):
            groups[lowered.tensors[tid].where].append(tid)
        self._amax = {}
        for where, tids in groups.items():
            self._amax[where] = torch.zeros(max(len(tids), 1), dtype=torch.int32, device=self.device)
            for k, tid in enumerate(tids):
                self._amax_slot[tid] = self._amax[where].data_ptr() + 4 * k
        self._amax_leaves = groups["leaf"]
        self._amax_used = {where: bool(tids) for where, tids in groups.items()}
        self.use_graph = use_graph
        self._hoist_valid = False
        self._graph = None
        self._graph_hoist = None
        self._copied = None
        self.launches = 0

    def _live_out_rank(self):
        low = self.low
        if len(low.tensors) == 1 or low.root < 0:
            return low.out_rank
        return len(low.tensors[low.root].labels)

    # ---------------------------------------------------------------- data path
    def upload(self, leaves):
        """Pack every leaf into pinned host memory and ship it with one async copy."""
        low = self.low
        if self._copied is not None:
            self._copied.synchronize()      # previous async copy still reads the staging buffer
        host = self.host_buf.numpy()
        for i, arr in enumerate(leaves):
            t = low.tensors[i]
            a = np.ascontiguousarray(np.asarray(arr), dtype=self.np_dtype).reshape(-1)
            host[t.offset:t.offset + a.nbytes] = a.view(np.uint8)
        self.leaf_buf.copy_(self.host_buf, non_blocking=True)
        self._copied = self.torch.cuda.Event()
        self._copied.record()
        self._hoist_valid = False            # the slice-independent intermediates belong to the previous leaves
        return low.leaf_bytes

    def _ptr(self, tid):
        t = self.low.tensors[tid]
        if t.where == "leaf":
            return self.leaf_buf.data_ptr() + t.offset
        if t.where == "hoist":
            return self.hoist_buf.data_ptr() + t.offset
        if t.where == "slice":
            return self.slice_buf.data_ptr() + t.offset
        return self.out.data_ptr()

    def _run_nodes(self, depends):
        st = self.torch.cuda.current_stream().cuda_stream
        n = 0
        slot = self._amax_slot
        where = "slice" if depends else "hoist"
        if self._amax_used[where]:
            self._amax[where].zero_()
        for node in self.low.nodes:
            if node.depends != depends:
                continue
            lib.check(self.lib.qtn_pair_run_scaled(C.byref(node.plan), self._ptr(node.a), self._ptr(node.b),
                                                   self._ptr(node.c), self.ws.data_ptr(),
                                                   self.slice_id.data_ptr(), slot.get(node.a), slot.get(node.b),
                                                   slot.get(node.c), st))
            n += 2 if (node.plan.split_bits or node.plan.kernel == lib.KERNEL_REDUCE) else 1
        return n

    def time_nodes(self, slice_id=0, reps=3, flush=None):
        """Per-node device times (ms, CUDA events on the launch stream) of one slice, nodes run in tree order so
        that every node sees its real inputs; each node is launched ``reps`` times in place and the mean of
        all but the first launch is reported (``flush``: a device buffer larger than L2, zeroed before every
        launch).  Measurement helper for bench.py / scripts: returns [(node, ms)] for the per-slice nodes."""
        torch = self.torch
        st = torch.cuda.current_stream().cuda_stream
        slot = self._amax_slot
        self._measure_leaves()
        self._run_hoisted()
        self._hoist_valid = True
        lib.check(self.lib.qtn_set_i32(self.slice_id.data_ptr(), slice_id, st))
        out = []
        if self._amax_used["slice"]:
            self._amax["slice"].zero_()
        for node in self.low.nodes:
            if not node.depends:
                continue
            ts = []
            for _ in range(reps):
                if flush is not None:
                    flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                lib.check(self.lib.qtn_pair_run_scaled(C.byref(node.plan), self._ptr(node.a), self._ptr(node.b),
                                                       self._ptr(node.c), self.ws.data_ptr(),
                                                       self.slice_id.data_ptr(), slot.get(node.a), slot.get(node.b),
                                                       slot.get(node.c), st))
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            out.append((node, sum(ts[1:]) / (len(ts) - 1) if len(ts) > 1 else ts[0]))
        return out

    def _measure_leaves(self):
        """Magnitude words of the leaves that an fp16-split plan reads directly (rare: such plans mostly
        read intermediates); one small streaming kernel per leaf, after every upload."""
        if not self._amax_leaves:
            return
        st = self.torch.cuda.current_stream().cuda_stream
        self._amax["leaf"].zero_()
        for tid in self._amax_leaves:
            t = self.low.tensors[tid]
            lib.check(self.lib.qtn_absmax(lib.QTN_C64, t.nbytes // 8, self._ptr(tid), self._amax_slot[tid], st))

    def run(self, slice_ids=None):
        """Hoisted nodes once, then the per-slice nodes for every id in ``slice_ids``
        (default: all).  Returns the output tensor (sum over those slices) on the device."""
        torch, low = self.torch, self.low
        st = torch.cuda.current_stream().cuda_stream
        if len(low.nodes) == 0:
            return self._single_leaf()
        ns = low.info.num_slices
        if slice_ids is None:
            slice_ids = range(ns)
        slice_ids = list(slice_ids)
        launches = 0
        any_dep = any(n.depends for n in low.nodes)
        if any_dep:
            self.out.zero_()
        if not self._hoist_valid or not any_dep:
            # slice-independent nodes: once per upload (their results stay valid while the leaves do -- every
            # rank of a sliced run would otherwise replay the same sub-tree at every call)
            self._measure_leaves()
            launches += self._run_hoisted()
            self._hoist_valid = True
        if any_dep and slice_ids:
            contiguous = slice_ids == list(range(slice_ids[0], slice_ids[0] + len(slice_ids)))
            if self.use_graph and contiguous and len(slice_ids) > 1:
                if self._graph is None:
                    self._capture()
                lib.check(self.lib.qtn_set_i32(self.slice_id.data_ptr(), slice_ids[0], st))
                for _ in slice_ids:
                    self._graph.replay()
                launches += (low.launches_per_slice + 1) * len(slice_ids) + 1
            else:
                for s in slice_ids:
                    lib.check(self.lib.qtn_set_i32(self.slice_id.data_ptr(), s, st))
                    launches += 1 + self._run_nodes(True)
        self.launches = launches
        return self.out.reshape((2,) * self._live_out_rank())

    def _run_hoisted(self):
        """Slice-independent nodes (mostly tiny gate-against-gate products): hundreds of
        launches, replayed as one CUDA graph after the first call."""
        n = self.low.launches_hoisted
        if not self.use_graph or n < 8:
            return self._run_nodes(False)
        torch = self.torch
        if self._graph_hoist is None:
            self._run_nodes(False)                      # warm-up outside capture
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run_nodes(False)
            self._graph_hoist = g
        self._graph_hoist.replay()
        return n

    def _capture(self):
        torch = self.torch
        # warm-up outside capture (sets function attributes, loads modules)
        saved = self.out.clone()
        self._run_nodes(True)
        torch.cuda.synchronize()
        self.out.copy_(saved)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            st = torch.cuda.current_stream().cuda_stream
            self._run_nodes(True)
            lib.check(self.lib.qtn_add_i32(self.slice_id.data_ptr(), 1, st))
        self.out.copy_(saved)
        self._graph = g

    def _single_leaf(self):
        low = self.low
        t = low.tensors[0]
        out_labels = list(low.out_labels)
        r = len(t.labels)
        pos_out = [r - 1 - out_labels.index(lab) for lab in t.labels]
        st = self.torch.cuda.current_stream().cuda_stream
        lib.check(self.lib.qtn_permute_bits(lib.dtype_code(low.dtype), r, lib.u8(t.pos), lib.u8(pos_out),
                                            self._ptr(0), self.out.data_ptr(), 0, st))
        self.launches = 1
        return self.out.reshape((2,) * r)
