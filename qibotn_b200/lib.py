"""ctypes binding of ``libqibotn_b200.so`` (the C ABI in ``include/qibotn_b200.h``).

The library is built in-tree by ``qibotn_b200/csrc/Makefile`` (``__graft_entry__.build()``).
There is no fallback: if it is missing, every compute entry point raises.
"""

from __future__ import annotations

import ctypes as C
import os
from functools import lru_cache

QTN_C64, QTN_C128 = 0, 1
QTN_MAX_RANK, QTN_MAX_SLICE, QTN_LO_MAX, QTN_HI_MAX = 48, 24, 16, 40
KERNEL_TILE, KERNEL_REDUCE, KERNEL_TC, KERNEL_KRED = 0, 1, 2, 3

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqibotn_b200.so")


class PairDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("rank_a", C.c_int32), ("rank_b", C.c_int32), ("rank_c", C.c_int32),
        ("label_a", C.c_int32 * QTN_MAX_RANK), ("pos_a", C.c_uint8 * QTN_MAX_RANK),
        ("label_b", C.c_int32 * QTN_MAX_RANK), ("pos_b", C.c_uint8 * QTN_MAX_RANK),
        ("label_c", C.c_int32 * QTN_MAX_RANK), ("pos_c", C.c_uint8 * QTN_MAX_RANK),
        ("c_layout_free", C.c_int32),
        ("conj_a", C.c_int32), ("conj_b", C.c_int32),
        ("accumulate", C.c_int32),
        ("n_slice_a", C.c_int32), ("n_slice_b", C.c_int32),
        ("slice_bit_a", C.c_uint8 * QTN_MAX_SLICE), ("slice_pos_a", C.c_uint8 * QTN_MAX_SLICE),
        ("slice_bit_b", C.c_uint8 * QTN_MAX_SLICE), ("slice_pos_b", C.c_uint8 * QTN_MAX_SLICE),
    ]


class PairPlan(C.Structure):
    _fields_ = [
        ("kernel", C.c_int32), ("dtype", C.c_int32), ("swapped", C.c_int32),
        ("tmb", C.c_int32), ("tnb", C.c_int32), ("tkb", C.c_int32),
        ("m_real", C.c_int32), ("n_real", C.c_int32), ("k_real", C.c_int32),
        ("n_mhi", C.c_int32), ("n_nhi", C.c_int32), ("n_khi", C.c_int32), ("n_batch", C.c_int32),
        ("split_bits", C.c_int32),
        ("conj_a", C.c_int32), ("conj_b", C.c_int32), ("accumulate", C.c_int32),
        ("na_lo", C.c_int32), ("nb_lo", C.c_int32),
        ("a_lo_pos", C.c_uint8 * QTN_LO_MAX), ("a_lo_sstr", C.c_uint16 * QTN_LO_MAX),
        ("b_lo_pos", C.c_uint8 * QTN_LO_MAX), ("b_lo_sstr", C.c_uint16 * QTN_LO_MAX),
        ("sa_stride", C.c_int32), ("sb_stride", C.c_int32),
        ("nc_lo", C.c_int32),
        ("c_lo_pos", C.c_uint8 * QTN_LO_MAX),
        ("c_m_sidx", C.c_uint16 * QTN_LO_MAX), ("c_n_sidx", C.c_uint16 * QTN_LO_MAX),
        ("a_mhi", C.c_uint8 * QTN_HI_MAX), ("c_mhi", C.c_uint8 * QTN_HI_MAX),
        ("b_nhi", C.c_uint8 * QTN_HI_MAX), ("c_nhi", C.c_uint8 * QTN_HI_MAX),
        ("a_khi", C.c_uint8 * QTN_HI_MAX), ("b_khi", C.c_uint8 * QTN_HI_MAX),
        ("a_bat", C.c_uint8 * QTN_HI_MAX), ("b_bat", C.c_uint8 * QTN_HI_MAX),
        ("c_bat", C.c_uint8 * QTN_HI_MAX),
        ("n_slice_a", C.c_int32), ("n_slice_b", C.c_int32),
        ("slice_bit_a", C.c_uint8 * QTN_MAX_SLICE), ("slice_pos_a", C.c_uint8 * QTN_MAX_SLICE),
        ("slice_bit_b", C.c_uint8 * QTN_MAX_SLICE), ("slice_pos_b", C.c_uint8 * QTN_MAX_SLICE),
        ("red_chunk_bits", C.c_int32),
        ("grid", C.c_int64), ("block", C.c_int32), ("smem_bytes", C.c_int32),
        ("workspace_bytes", C.c_int64), ("c_elems", C.c_int64),
        ("flops", C.c_double), ("bytes", C.c_double),
        ("tc_stages", C.c_int32), ("tc_plane_a", C.c_int32), ("tc_plane_b", C.c_int32),
        ("tc_reserved", C.c_int32),
        ("tc_a_it_g", C.c_int64 * 8), ("tc_b_it_g", C.c_int64 * 16),
        ("tc_a_it_s", C.c_int32 * 8), ("tc_b_it_s", C.c_int32 * 16),
        ("tc_terms", C.c_int32), ("tc_pad", C.c_int32),
        ("tc_prepack", C.c_int32), ("tc_pair", C.c_int32), ("tc_prepack_off", C.c_int64),
        ("tc_kind", C.c_int32), ("tc_gather", C.c_int32),
        ("tc_a_gbit", C.c_uint8 * QTN_LO_MAX), ("tc_a_rank", C.c_uint8 * QTN_LO_MAX),
        ("tc_a_it_gg", C.c_int64 * 8), ("tc_a_it_raw", C.c_int32 * 8),
        ("tile_stages", C.c_int32), ("kred_inner", C.c_int32), ("kred_c_in", C.c_uint8 * 4),
        ("tc_prepack_a", C.c_int32), ("tc_run_bits", C.c_int32),
        ("tc_prepack_a_off", C.c_int64),
    ]


class Options(C.Structure):
    """``qtn_options``: per-call switches of the planner and the SVD driver (include/qibotn_b200.h)."""
    _fields_ = [(n, C.c_int32) for n in ("tc_enable", "tc_min_log2", "tc_persistent", "tc_terms", "tc_accum_split",
                                         "tc_prepack", "tc_pair", "tc_stagger", "tc_fp16", "svd_wide",
                                         "svd_group_mb", "svd_reg", "tc_gather", "tc_prepack_a", "tc_gather_min_run")] + [("reserved", C.c_int32 * 1)]


class GemmItem(C.Structure):
    _fields_ = [("a", C.c_void_p), ("b", C.c_void_p), ("c", C.c_void_p), ("m", C.c_int32), ("n", C.c_int32),
                ("k", C.c_int32), ("lda", C.c_int32), ("ldb", C.c_int32), ("ldc", C.c_int32),
                ("opa", C.c_int32), ("opb", C.c_int32)]


class SvdItem(C.Structure):
    _fields_ = [("w", C.c_void_p), ("n", C.c_int32), ("m", C.c_int32),
                ("block", C.c_int32), ("nblocks", C.c_int32)]


class SvdTrunc(C.Structure):
    _fields_ = [("abs_cutoff", C.c_double), ("rel_cutoff", C.c_double),
                ("discarded_weight_cutoff", C.c_double), ("max_extent", C.c_int32),
                ("normalization", C.c_int32)]


OP_N, OP_C, OP_T, OP_H = 0, 1, 2, 3


class QtnError(RuntimeError):
    pass


_SIGNATURES = {
    "qtn_version": (C.c_char_p, []),
    "qtn_last_error": (C.c_char_p, []),
    "qtn_device_info": (C.c_int, [C.c_int] + [C.POINTER(C.c_int)] * 4),
    "qtn_pair_plan": (C.c_int, [C.POINTER(PairDesc), C.c_int, C.POINTER(PairPlan)]),
    "qtn_pair_plan_opts": (C.c_int, [C.POINTER(PairDesc), C.c_int, C.POINTER(Options), C.POINTER(PairPlan)]),
    "qtn_options_default": (None, [C.POINTER(Options)]),
    "qtn_options_current": (None, [C.POINTER(Options)]),
    "qtn_svd_rows_opts": (C.c_int, [C.c_int, C.c_int, C.POINTER(SvdItem), C.c_double, C.c_int,
                                    C.POINTER(SvdTrunc), C.c_void_p, C.c_void_p, C.c_int64,
                                    C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_int32),
                                    C.c_void_p, C.c_int64, C.POINTER(Options), C.c_void_p]),
    "qtn_pair_run": (C.c_int, [C.POINTER(PairPlan), C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p]),
    "qtn_pair_run_scaled": (C.c_int, [C.POINTER(PairPlan), C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qtn_absmax": (C.c_int, [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qtn_permute_bits": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8),
                                   C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "qtn_permute": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                              C.c_void_p, C.c_void_p, C.c_void_p]),
    "qtn_axpy": (C.c_int, [C.c_int, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                           C.c_int, C.c_void_p]),
    "qtn_gemm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                           C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "qtn_mps_apply_1q": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qtn_mps_apply_2q": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qtn_gemm_batched": (C.c_int, [C.c_int, C.c_int, C.POINTER(GemmItem), C.c_int, C.c_void_p]),
    "qtn_mps_sample_step": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qtn_svd_workspace_bytes": (C.c_int64, [C.c_int, C.c_int64]),
    "qtn_svd_lq_workspace_bytes": (C.c_int64, [C.c_int]),
    "qtn_svd_lq": (C.c_int, [C.c_int, C.c_int, C.POINTER(SvdItem), C.POINTER(SvdItem), C.c_void_p, C.c_int64,
                             C.c_void_p]),
    "qtn_svd_rows": (C.c_int, [C.c_int, C.c_int, C.POINTER(SvdItem), C.c_double, C.c_int,
                               C.POINTER(SvdTrunc), C.c_void_p, C.c_void_p, C.c_int64,
                               C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_int32),
                               C.c_void_p, C.c_int64, C.c_void_p]),
    "qtn_svd_emit": (C.c_int, [C.c_int, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_double,
                               C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qtn_pauli_expectation": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "qtn_pauli_expectation_host": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p]),
    "qtn_microbench_fma": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_double), C.c_void_p]),
    "qtn_sizeof": (C.c_int64, [C.c_char_p]),
    "qtn_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "qtn_get_option": (C.c_int, [C.c_char_p, C.POINTER(C.c_int)]),
    "qtn_set_i32": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "qtn_add_i32": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
}


@lru_cache(maxsize=1)
def load():
    """Load the shared library; raises ``QtnError`` with build advice when it is absent."""
    if not os.path.exists(LIB_PATH):
        raise QtnError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C qibotn_b200/csrc`). qibotn_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    # the ctypes mirrors must be the structs this build of the library was compiled with
    for cname, mirror in (("qtn_pair_desc", PairDesc), ("qtn_pair_plan_t", PairPlan), ("qtn_options", Options),
                          ("qtn_gemm_item", GemmItem)):
        want = int(lib.qtn_sizeof(cname.encode()))
        if want != C.sizeof(mirror):
            raise QtnError(f"{LIB_PATH} is stale: sizeof({cname}) = {want} there, {C.sizeof(mirror)} in lib.py; "
                           "rebuild it (`make -C qibotn_b200/csrc`)")
    # A/B runs of whole scripts: QTN_OPTIONS="tc_gather=0,svd_reg=5" edits the process-wide option set once
    for item in filter(None, os.environ.get("QTN_OPTIONS", "").split(",")):
        key, _, val = item.partition("=")
        if lib.qtn_set_option(key.strip().encode(), int(val)) != 0:
            raise QtnError(f"QTN_OPTIONS: unknown option {key!r}")
    return lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(status):
    if status != 0:
        raise QtnError(f"qibotn_b200 error {status}: {load().qtn_last_error().decode()}")


def set_option(name, value):
    """Library-wide switch (``tc_enable``, ``tc_min_log2``); see include/qibotn_b200.h."""
    check(load().qtn_set_option(name.encode(), int(value)))


def get_option(name):
    v = C.c_int(0)
    check(load().qtn_get_option(name.encode(), C.byref(v)))
    return v.value


def dtype_code(dtype) -> int:
    import numpy as np

    s = str(np.dtype(dtype)) if not isinstance(dtype, str) else dtype
    s = s.replace("torch.", "")
    if s == "complex64":
        return QTN_C64
    if s == "complex128":
        return QTN_C128
    raise TypeError(f"unsupported dtype {dtype!r}")


def u8(values):
    return (C.c_uint8 * len(values))(*values)


def make_pair_desc(dtype, a, b, c=None, accumulate=False, conj_a=False, conj_b=False,
                   slice_a=(), slice_b=()):
    """``a``/``b``: lists of (label, address_bit).  ``c``: list of (label, address_bit), or a
    plain list of labels to let the planner choose C's layout."""
    d = PairDesc()
    d.dtype = dtype_code(dtype)
    if len(a) > QTN_MAX_RANK or len(b) > QTN_MAX_RANK:
        raise QtnError("tensor rank exceeds QTN_MAX_RANK")
    d.rank_a, d.rank_b = len(a), len(b)
    for i, (lab, pos) in enumerate(a):
        d.label_a[i], d.pos_a[i] = lab, pos
    for i, (lab, pos) in enumerate(b):
        d.label_b[i], d.pos_b[i] = lab, pos
    free = c is None or (len(c) > 0 and not isinstance(c[0], (tuple, list)))
    if c is None:
        sa, sb = {l for l, _ in a}, {l for l, _ in b}
        c = sorted(sa ^ sb)
    d.rank_c = len(c)
    d.c_layout_free = 1 if free else 0
    for i, item in enumerate(c):
        if free:
            d.label_c[i] = item
        else:
            d.label_c[i], d.pos_c[i] = item
    d.accumulate, d.conj_a, d.conj_b = int(accumulate), int(conj_a), int(conj_b)
    d.n_slice_a, d.n_slice_b = len(slice_a), len(slice_b)
    for i, (bit, pos) in enumerate(slice_a):
        d.slice_bit_a[i], d.slice_pos_a[i] = bit, pos
    for i, (bit, pos) in enumerate(slice_b):
        d.slice_bit_b[i], d.slice_pos_b[i] = bit, pos
    return d


def options(**overrides):
    """The library's default ``qtn_options`` with ``overrides`` applied (e.g. ``options(tc_fp16=0)``): handed to
    ``pair_plan`` / ``qtn_svd_rows_opts`` it takes the place of the process-wide switches for that call."""
    opt = Options()
    load().qtn_options_default(C.byref(opt))
    for key, val in overrides.items():
        if not hasattr(opt, key) or key == "reserved":
            raise QtnError(f"unknown option {key!r}")
        setattr(opt, key, int(val))
    return opt


def pair_plan(desc, sm_count=148, options=None):
    """Launch-ready plan of one pairwise contraction; ``options``: a ``qtn_options`` (see ``lib.options``) for this
    call only, None = the process-wide switches."""
    plan = PairPlan()
    if options is None:
        check(load().qtn_pair_plan(C.byref(desc), sm_count, C.byref(plan)))
    else:
        check(load().qtn_pair_plan_opts(C.byref(desc), sm_count, C.byref(options), C.byref(plan)))
    return plan
