"""GPU parity of the MPS path: general GEMM, site kernels, batched Jacobi SVD with truncation,
the gate loop of mps_utils.py and the read-outs of mps_contraction_helper.py, against the CPU
oracle (oracle/mps.py, oracle/statevector.py) and LAPACK singular values."""
import ctypes as C

import numpy as np
import pytest

from oracle import mps as omps
from oracle import statevector as osv
from oracle.network import gate_tensors
from qibotn_b200 import MetaBackend, circuits, lib
from qibotn_b200 import mps as gmps

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL = {"complex64": 2e-5, "complex128": 1e-10}


def _rand(rng, *shape, dtype="complex128"):
    return (rng.normal(size=shape) + 1j * rng.normal(size=shape)).astype(dtype)


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("shape", [(1, 1, 1), (5, 7, 3), (64, 64, 16), (130, 70, 33), (2, 300, 129)])
@pytest.mark.parametrize("ops", [(0, 0), (2, 0), (0, 3), (3, 1), (1, 2)])
def test_gemm_all_ops(shape, ops, dtype):
    m, n, k = shape
    opa, opb = ops
    rng = np.random.default_rng(m * 1000 + n * 10 + k + opa * 7 + opb)
    a = _rand(rng, *((k, m) if opa & 2 else (m, k)), dtype=dtype)
    b = _rand(rng, *((n, k) if opb & 2 else (k, n)), dtype=dtype)
    A = a.T if opa & 2 else a
    A = A.conj() if opa & 1 else A
    B = b.T if opb & 2 else b
    B = B.conj() if opb & 1 else B
    want = A.astype(np.complex128) @ B.astype(np.complex128)
    h = lib.load()
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    dc = torch.zeros(m, n, dtype=da.dtype, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    lib.check(h.qtn_gemm(lib.dtype_code(dtype), opa, opb, m, n, k, da.data_ptr(), a.shape[1],
                         db.data_ptr(), b.shape[1], dc.data_ptr(), n, 0, st))
    got = dc.cpu().numpy()
    tol = 1e-5 if dtype == "complex64" else 1e-12
    assert np.abs(got - want).max() <= tol * (np.abs(want).max() + 1)
    lib.check(h.qtn_gemm(lib.dtype_code(dtype), opa, opb, m, n, k, da.data_ptr(), a.shape[1],
                         db.data_ptr(), b.shape[1], dc.data_ptr(), n, 1, st))
    assert np.abs(dc.cpu().numpy() - 2 * want).max() <= 2 * tol * (np.abs(want).max() + 1)


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_gemm_batched_mixed_shapes_and_ops(dtype):
    """qtn_gemm_batched: 40 products of different shapes and operand forms (two launches of <= 32)."""
    rng = np.random.default_rng(3)
    h = lib.load()
    shapes = [(1, 1, 1), (5, 7, 3), (64, 64, 16), (130, 70, 33), (2, 300, 129), (96, 200, 64), (0, 4, 4), (65, 1, 257)]
    ops = [(0, 0), (2, 0), (0, 3), (3, 1), (1, 2)]
    items = (lib.GemmItem * 40)()
    keep, wants = [], []
    for i in range(40):
        m, n, k = shapes[i % len(shapes)]
        opa, opb = ops[i % len(ops)]
        a = _rand(rng, *((k, m) if opa & 2 else (m, k)), dtype=dtype)
        b = _rand(rng, *((n, k) if opb & 2 else (k, n)), dtype=dtype)
        A = a.T if opa & 2 else a
        A = A.conj() if opa & 1 else A
        B = b.T if opb & 2 else b
        B = B.conj() if opb & 1 else B
        wants.append(A.astype(np.complex128) @ B.astype(np.complex128))
        da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        dc = torch.full((m, n), float("nan"), dtype=da.dtype, device="cuda")
        keep.append((da, db, dc))
        it = items[i]
        it.a, it.b, it.c = da.data_ptr() or 1, db.data_ptr() or 1, dc.data_ptr() or 1
        it.m, it.n, it.k, it.lda, it.ldb, it.ldc = m, n, k, max(a.shape[1], 1), max(b.shape[1], 1), max(n, 1)
        it.opa, it.opb = opa, opb
    lib.check(h.qtn_gemm_batched(lib.dtype_code(dtype), 40, items, 0, torch.cuda.current_stream().cuda_stream))
    tol = 1e-5 if dtype == "complex64" else 1e-12
    for (_, _, dc), want in zip(keep, wants):
        got = dc.cpu().numpy()
        assert got.shape == want.shape
        if want.size:
            assert np.abs(got - want).max() <= tol * (np.abs(want).max() + 1)


def _svd_split(theta, dtype, algorithm):
    """theta (p x q numpy) -> (left p x k, right k x q, sigma) through the C ABI, the way
    MPSEngine.flush drives it."""
    # partition None (singular values kept apart) is refused by the MPS update, which would drop them
    # (parse_algorithm); the C ABI below still offers it, so it is set on the engine directly
    method = dict(algorithm["svd_method"])
    part = method.pop("partition", None)
    eng = gmps.MPSEngine(2, dtype, {"qr_method": False, "svd_method": dict(method, partition="UV")})
    eng.partition = gmps._PARTITION[part]
    p, q = theta.shape
    t = torch.from_numpy(theta.astype(dtype)).cuda().reshape(-1)
    transposed = p > q
    w = eng._transpose(t, p, q) if transposed else t.clone()
    items = (lib.SvdItem * 1)()
    items[0].w, items[0].n, items[0].m = w.data_ptr(), (q if transposed else p), (p if transposed else q)
    ld = items[0].n
    sigma = torch.empty(ld, dtype=torch.float64, device="cuda")
    perm = torch.empty(ld, dtype=torch.int32, device="cuda")
    nbytes = int(eng.lib.qtn_svd_workspace_bytes(1, ld))
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    rank, scale, sweeps = (C.c_int32 * 1)(), (C.c_double * 1)(), C.c_int32(0)
    st = torch.cuda.current_stream().cuda_stream
    lib.check(eng.lib.qtn_svd_rows(eng.code, 1, items, 0.0, 40, C.byref(eng.trunc), sigma.data_ptr(),
                                   perm.data_ptr(), ld, rank, scale, C.byref(sweeps), ws.data_ptr(), nbytes, st))
    k = int(rank[0])
    mcols = items[0].m
    o1 = torch.empty(k * mcols, dtype=t.dtype, device="cuda")
    o2 = torch.empty(k * mcols, dtype=t.dtype, device="cuda")
    lib.check(eng.lib.qtn_svd_emit(eng.code, w.data_ptr(), mcols, perm.data_ptr(), sigma.data_ptr(),
                                   float(scale[0]), k, eng.partition, int(transposed), o1.data_ptr(),
                                   o2.data_ptr(), st))
    if not transposed:
        right = o1.view(k, q)
        left = torch.empty(p * k, dtype=t.dtype, device="cuda")
        eng._gemm(lib.OP_N, lib.OP_H, p, k, q, t, q, o2, q, left, k)
        left = left.view(p, k)
    else:
        left = eng._transpose(o1, k, p).view(p, k)
        right = torch.empty(k * q, dtype=t.dtype, device="cuda")
        eng._gemm(lib.OP_C, lib.OP_N, k, q, p, o2, p, t, q, right, q)
        right = right.view(k, q)
    return left.cpu().numpy(), right.cpu().numpy(), sigma.cpu().numpy()[:k] * scale[0], int(sweeps.value)


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("shape", [(2, 2), (4, 4), (6, 10), (10, 6), (32, 32), (64, 48), (70, 200), (160, 160),
                                   (257, 130)])
def test_jacobi_svd_matches_lapack(shape, dtype):
    rng = np.random.default_rng(shape[0] * 31 + shape[1])
    theta = _rand(rng, *shape)
    s_ref = np.linalg.svd(theta.astype(dtype).astype(np.complex128), compute_uv=False)
    # north-star tolerances, relative to the largest singular value (the 2-norm of theta)
    tol = 1e-5 if dtype == "complex64" else 1e-10
    for part in (None, "U", "V", "UV"):
        left, right, sig, sweeps = _svd_split(theta, dtype, {"svd_method": {"partition": part}})
        assert sweeps < 40
        assert sig.shape == s_ref.shape
        assert np.abs(sig - s_ref).max() <= tol * s_ref[0]
        if part is None:
            rec = (left * sig) @ right
            # isometries: left^H left = 1, right right^H = 1.  A singular vector is determined to
            # eps * sigma_1 / sigma_i (any SVD), so the bound carries the condition number, nothing else
            cond = max(1.0, s_ref[0] / s_ref[-1])
            assert np.abs(left.conj().T @ left - np.eye(len(sig))).max() <= tol * cond
            assert np.abs(right @ right.conj().T - np.eye(len(sig))).max() <= tol * cond
        else:
            rec = left @ right
        assert np.abs(rec - theta).max() <= tol * s_ref[0]


@pytest.mark.parametrize("graded", [False, True])
def test_batched_512_blocks_match_lapack(graded):
    """The chi = 256 shape of BASELINE config 3: a layer of 512 x 512 complex128 two-site blocks through
    MPSEngine.flush (batched GEMM, LQ reduction, Jacobi sweeps, truncation to 256, factor recovery)
    against numpy's LAPACK SVD: kept singular values to 1e-10 of sigma_1 and the truncated product
    to 1e-10 of sigma_1, for a flat spectrum and one graded over 12 decades."""
    rng = np.random.default_rng(11 + graded)
    nblk, chi = 3, 256
    eng = gmps.MPSEngine(2 * nblk, "complex128",
                         {"qr_method": False, "svd_method": {"partition": "V", "max_extent": chi, "abs_cutoff": 1e-12}})
    thetas = []
    for k in range(nblk):
        u, _ = np.linalg.qr(_rand(rng, 2 * chi, 2 * chi))
        v, _ = np.linalg.qr(_rand(rng, 2 * chi, 2 * chi))
        s = 10.0 ** (-12.0 * np.arange(2 * chi) / (2 * chi)) if graded else 1.0 + rng.random(2 * chi)
        s = np.sort(s)[::-1]
        a = (u * s)                                    # (2 chi) x (2 chi) = A reshaped (chi, 2, 2 chi)
        eng.sites[2 * k] = torch.from_numpy(a.reshape(chi, 2, 2 * chi)).cuda().contiguous()
        eng.sites[2 * k + 1] = torch.from_numpy(v.conj().T.reshape(2 * chi, 2, chi).copy()).cuda().contiguous()
        thetas.append(((u * s) @ v.conj().T, s))
    for k in range(nblk):
        eng.apply_gate(np.eye(4).reshape(2, 2, 2, 2), (2 * k, 2 * k + 1))
    eng.flush()
    for k, (theta, s) in enumerate(thetas):
        left = eng.sites[2 * k].reshape(2 * chi, -1).cpu().numpy()
        right = eng.sites[2 * k + 1].reshape(-1, 2 * chi).cpu().numpy()
        kept = left.shape[1]
        s_ref = np.linalg.svd(theta, compute_uv=False)
        want_rank = min(chi, int((s_ref >= 1e-12).sum()))
        assert kept == want_rank
        u, sv, vh = np.linalg.svd(theta)
        trunc = (u[:, :kept] * sv[:kept]) @ vh[:kept]
        assert np.abs(np.linalg.norm(right, axis=1) - s_ref[:kept]).max() <= 1e-10 * s_ref[0]
        if not graded:
            # a flat spectrum has no gap at the truncation rank: compare the projectors' effect on theta only
            assert np.abs(left.conj().T @ left - np.eye(kept)).max() <= 1e-10
            assert abs(np.linalg.norm(left @ right) ** 2 - (sv[:kept] ** 2).sum()) <= 1e-9 * (sv ** 2).sum()
        else:
            assert np.abs(left @ right - trunc).max() <= 1e-10 * s_ref[0]


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_truncation_rules(dtype):
    rng = np.random.default_rng(7)
    u, _ = np.linalg.qr(_rand(rng, 24, 24))
    v, _ = np.linalg.qr(_rand(rng, 40, 40))
    s = np.array([3.0, 2.0, 1.0, 0.5, 0.2, 1e-3, 1e-4] + [1e-9] * 5 + [0.0] * 12)
    theta = (u * s) @ v[:24]
    cases = [({"abs_cutoff": 1e-6}, 7), ({"abs_cutoff": 0.3}, 4), ({"rel_cutoff": 0.1}, 4),
             ({"max_extent": 3}, 3), ({"abs_cutoff": 1e-6, "max_extent": 5}, 5),
             ({"discarded_weight_cutoff": 0.01}, 4), ({"abs_cutoff": 10.0}, 1)]
    for method, keep in cases:
        algo = {"svd_method": dict(method, partition="UV")}
        left, right, sig, _ = _svd_split(theta, dtype, algo)
        assert omps.truncation_rank(s, algo["svd_method"]) == keep
        assert len(sig) == keep, (method, len(sig))
        want = (u[:, :keep] * s[:keep]) @ v[:keep]
        tol = 1e-5 if dtype == "complex64" else 1e-10
        assert np.abs(left @ right - want).max() <= tol * s[0]
    for norm, f in (("L1", lambda x: x / x.sum()), ("L2", lambda x: x / np.sqrt((x ** 2).sum())),
                    ("LInf", lambda x: x / x[0])):
        _, _, sig, _ = _svd_split(theta, dtype, {"svd_method": {"max_extent": 4, "normalization": norm}})
        assert np.allclose(sig, f(s[:4]), rtol=1e-4 if dtype == "complex64" else 1e-10)


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("lmr", [(64, 128, 64), (48, 100, 80), (80, 100, 48), (128, 40, 128)])
@pytest.mark.parametrize("graded", [False, True])
@pytest.mark.parametrize("svd_reg", [0, 1, 3, 4, 5])
def test_lq_preconditioned_split(lmr, graded, dtype, svd_reg, monkeypatch):
    """Two-site blocks of >= LQ_MIN_ROWS rows go through qtn_svd_lq (LQ preconditioner) before the
    Jacobi sweeps; the split must agree with LAPACK and with the un-preconditioned path, for every
    cross-step kernel (``svd_reg``: 0 shared-memory, 1 register-resident, 3 streaming, 4 two-pass)."""
    default = lib.get_option("svd_reg")
    lib.set_option("svd_reg", svd_reg)
    try:
        _lq_split_case(lmr, graded, dtype, monkeypatch)
    finally:
        lib.set_option("svd_reg", default)


def _lq_split_case(lmr, graded, dtype, monkeypatch):
    l, m, r = lmr
    rng = np.random.default_rng(l * 7 + m * 3 + r + graded)
    # theta = U diag(s) V^H with a known spectrum: condition number 10, or 10 decades (5 for complex64);
    # rank min(2l, m, 2r) (the last shape is rank-deficient)
    m = min(2 * l, m, 2 * r)
    u, _ = np.linalg.qr(_rand(rng, 2 * l, m))
    v, _ = np.linalg.qr(_rand(rng, 2 * r, m))
    span = (10 if dtype == "complex128" else 5) if graded else 1
    a, b = u * 10.0 ** (-span * np.arange(m) / m), v.conj().T
    theta = (a.astype(dtype) @ b.astype(dtype)).astype(np.complex128)
    s_ref = np.linalg.svd(theta, compute_uv=False)
    tol = 1e-5 if dtype == "complex64" else 1e-10
    results = {}
    for lq_rows in (0, 96):
        monkeypatch.setattr(gmps, "LQ_MIN_ROWS", lq_rows)
        for part in ("UV", "U", "V"):
            eng = gmps.MPSEngine(2, dtype, {"qr_method": False, "svd_method": {"partition": part}})
            eng.sites = [torch.from_numpy(a.astype(dtype)).cuda().reshape(l, 2, m).contiguous(),
                         torch.from_numpy(b.astype(dtype)).cuda().reshape(m, 2, r).contiguous()]
            eng.apply_gate(np.eye(4).reshape(2, 2, 2, 2), (0, 1))
            eng.flush()
            left = eng.sites[0].reshape(2 * l, -1).cpu().numpy().astype(np.complex128)
            right = eng.sites[1].reshape(-1, 2 * r).cpu().numpy().astype(np.complex128)
            k = left.shape[1]
            floor = 16 * (1.1e-16 if dtype == "complex128" else 6e-8) * s_ref[0]
            assert int((s_ref > 8 * floor).sum()) <= k <= m
            assert np.abs(left @ right - theta).max() <= tol * s_ref[0]
            if not graded:
                cond = s_ref[0] / s_ref[k - 1]
                if part == "V":
                    assert np.abs(left.conj().T @ left - np.eye(k)).max() <= tol * cond
                    assert np.abs(np.linalg.norm(right, axis=1) - s_ref[:k]).max() <= tol * s_ref[0]
                if part == "U":
                    assert np.abs(right @ right.conj().T - np.eye(k)).max() <= tol * cond
                    assert np.abs(np.linalg.norm(left, axis=0) - s_ref[:k]).max() <= tol * s_ref[0]
            results[(lq_rows, part)] = eng.stats["svd_sweeps"]
    assert results[(96, "UV")] <= results[(0, "UV")] + 2       # the preconditioner never costs sweeps


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("lmr", [(8, 24, 8), (40, 64, 20), (20, 64, 40), (64, 100, 64), (3, 5, 9)])
def test_qr_split_is_a_qr(lmr, dtype):
    """``svd_method: False`` (cuQuantum's default algorithm): exact split, isometric left factor and upper
    triangular right factor for blocks with p >= q, L.Q (lower triangular, isometric rows) for p < q."""
    l, m, r = lmr
    rng = np.random.default_rng(l + 10 * m + 100 * r)
    a, b = _rand(rng, l, 2, m).astype(dtype), _rand(rng, m, 2, r).astype(dtype)
    gate = np.linalg.qr(_rand(rng, 4, 4))[0].reshape(2, 2, 2, 2)
    for algo in ({"qr_method": {}, "svd_method": False}, {}, None):
        eng = gmps.MPSEngine(2, dtype, algo)
        eng.sites = [torch.from_numpy(a).cuda().contiguous(), torch.from_numpy(b).cuda().contiguous()]
        eng.apply_gate(gate, (0, 1))
        eng.flush()
        p, q = 2 * l, 2 * r
        left = eng.sites[0].reshape(p, -1).cpu().numpy().astype(np.complex128)
        right = eng.sites[1].reshape(-1, q).cpu().numpy().astype(np.complex128)
        k = min(p, q)
        assert left.shape == (p, k) and right.shape == (k, q)
        theta = np.einsum("ipj,jqk,rspq->irsk", a.astype(np.complex128), b.astype(np.complex128), gate).reshape(p, q)
        tol = 2e-5 if dtype == "complex64" else 1e-11
        scale = np.abs(theta).max()
        assert np.abs(left @ right - theta).max() <= tol * scale
        if p >= q:
            assert np.abs(left.conj().T @ left - np.eye(k)).max() <= tol * 10
            assert np.abs(np.tril(right, -1)).max() == 0.0
        else:
            assert np.abs(right @ right.conj().T - np.eye(k)).max() <= tol * 10
            assert np.abs(np.triu(left, 1)).max() == 0.0


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_qr_assisted_svd_gives_the_same_factors(dtype):
    """``qr_method: {}`` next to ``svd_method``: every block goes through the triangular reduction before the
    Jacobi sweeps (cuTensorNet's QR-assisted SVD); singular values and the product are unchanged."""
    rng = np.random.default_rng(3)
    a, b = _rand(rng, 12, 2, 30).astype(dtype), _rand(rng, 30, 2, 20).astype(dtype)
    out = {}
    for qr in (False, {}):
        eng = gmps.MPSEngine(2, dtype, {"qr_method": qr, "svd_method": {"partition": "V", "max_extent": 16}})
        assert eng.qr_assist == (qr is not False)
        eng.sites = [torch.from_numpy(a).cuda().contiguous(), torch.from_numpy(b).cuda().contiguous()]
        eng.apply_gate(np.eye(4).reshape(2, 2, 2, 2), (0, 1))
        eng.flush()
        left = eng.sites[0].reshape(24, -1).cpu().numpy().astype(np.complex128)
        right = eng.sites[1].reshape(-1, 40).cpu().numpy().astype(np.complex128)
        out[bool(qr is not False)] = (left @ right, np.linalg.norm(right, axis=1))
    tol = 3e-5 if dtype == "complex64" else 1e-11
    assert np.abs(out[True][0] - out[False][0]).max() <= tol * np.abs(out[False][0]).max() * 10
    assert np.abs(out[True][1] - out[False][1]).max() <= tol * out[False][1][0]


def _mps_state(circ, dtype, algorithm):
    conv = gmps.QiboCircuitToMPS(circ, algorithm, dtype=dtype)
    dense = gmps.MPSContractionHelper(circ.nqubits).contract_state_vector(conv.mps_tensors)
    return dense.reshape(-1).cpu().numpy(), conv


DEFAULT_ALGO = {"qr_method": False, "svd_method": {"partition": "UV", "abs_cutoff": 1e-12}}


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("nqubits", [2, 5, 10])
def test_mps_qft_uniform(nqubits, dtype):
    """reference tests/test_cuquantum_cutensor_backend.py:88-142 (bond dimension stays 1)."""
    b = MetaBackend.load("cutensornet")
    b.set_precision("single" if dtype == "complex64" else "double")
    want = np.full(1 << nqubits, 2.0 ** (-nqubits / 2))
    for runcard_value in (True, DEFAULT_ALGO):
        b.configure_tn_simulation({"MPI_enabled": False, "MPS_enabled": runcard_value, "NCCL_enabled": False,
                                   "expectation_enabled": False})
        got = b.execute_circuit(circuit=circuits.qft(nqubits, True)).statevector.flatten().get()
        assert np.allclose(got, want, rtol=1e-5, atol=1e-8 if dtype == "complex128" else 1e-6)


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("builder", ["ring", "brick", "tfim", "distant"])
def test_mps_state_vs_dense_simulator(builder, dtype):
    if builder == "ring":
        circ = circuits.ry_rz_cnot_ring(8, 3, seed=42)          # has the (7,0) CNOT: swaps + reversed pair
    elif builder == "brick":
        circ = circuits.brickwork(10, 6, seed=3)
    elif builder == "tfim":
        circ = circuits.tfim_trotter(12, 4)
    else:
        circ = circuits.Circuit(7)
        rng = np.random.default_rng(5)
        for q in range(7):
            circ.add(circuits.gates.U3(q, *rng.uniform(-3, 3, 3)))
        for a, b in [(6, 2), (0, 5), (3, 1), (2, 6), (4, 0)]:
            circ.add(circuits.gates.CNOT(a, b))
            circ.add(circuits.gates.RY(b, theta=float(rng.uniform(-3, 3))))
    want = osv.simulate(circ)
    got, conv = _mps_state(circ, dtype, DEFAULT_ALGO)
    assert np.abs(got - want).max() < TOL[dtype] * 10
    # the oracle MPS (numpy.linalg.svd, same truncation rule) has the same bond profile
    ref = omps.evolve(circ, DEFAULT_ALGO)
    if dtype == "complex128":
        assert [t.shape[2] for t in ref[:-1]] == [int(t.shape[2]) for t in conv.mps_tensors[:-1]]


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_mps_truncated_matches_oracle(dtype):
    circ = circuits.brickwork(10, 8, seed=11)
    algo = {"qr_method": False, "svd_method": {"partition": "UV", "abs_cutoff": 1e-12, "max_extent": 6}}
    ref = omps.to_dense(omps.evolve(circ, algo)).reshape(-1)
    got, conv = _mps_state(circ, dtype, algo)
    assert max(int(t.shape[2]) for t in conv.mps_tensors[:-1]) == 6
    # same local truncation => same (unnormalised) state up to rounding
    assert np.abs(got - ref).max() < (2e-4 if dtype == "complex64" else 1e-8)
    exact = osv.simulate(circ)
    assert np.abs(got - exact).max() > 1e-4            # the truncation really happened


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_norm_and_expectation_sweeps(dtype):
    circ = circuits.tfim_trotter(10, 3)
    psi = osv.simulate(circ)
    conv = gmps.QiboCircuitToMPS(circ, DEFAULT_ALGO, dtype=dtype)
    helper = gmps.MPSContractionHelper(10)
    tol = TOL[dtype] * 10
    assert abs(float(helper.contract_norm(conv.mps_tensors).cpu()) - 1.0) < tol
    zz = np.kron(np.diag([1.0, -1.0]), np.diag([1.0, -1.0]))
    for qubits, op in (((3,), np.array([[0, 1], [1, 0]])), ((2, 3), zz), ((1, 7), zz),
                       ((4, 5), np.array([[0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0], [1, 0, 0, 0]]))):
        got = complex(helper.contract_expectation(conv.mps_tensors, op, qubits).cpu())
        phi = osv.apply_matrix(psi.copy(), np.asarray(op, dtype=complex), list(qubits), 10)
        assert abs(got - np.vdot(psi, phi)) < tol
    ref = omps.evolve(circ, DEFAULT_ALGO)
    assert abs(omps.norm(ref) - 1.0) < 1e-10


def test_mps_rejects_three_qubit_gates_and_pure_qr_keeps_the_state():
    circ = circuits.ry_rz_cnot_ring(6, 2, seed=1)
    want = osv.simulate(circ)
    got, _ = _mps_state(circ, "complex128", {"qr_method": {}, "svd_method": False})
    assert np.abs(got - want).max() < 1e-10
    eng = gmps.MPSEngine(4, "complex128", DEFAULT_ALGO)
    with pytest.raises(NotImplementedError):
        eng.apply_gate(np.eye(8).reshape((2,) * 6), (0, 1, 2))
    gts, _ = gate_tensors(circ, np.complex128)
    assert len(gts) > 0


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_mps_sampling_matches_oracle_draw_for_draw(dtype):
    """Perfect sampling on the GPU (qtn_mps_sample_step) against the numpy restatement with the SAME
    uniforms: identical bit strings, |amplitude|^2 per shot equal to the dense simulator's."""
    n, nshots = 8, 1500
    circ = circuits.ry_rz_cnot_ring(n, 3, seed=5)
    psi = osv.simulate(circ)
    conv = gmps.QiboCircuitToMPS(circ, DEFAULT_ALGO, dtype=dtype)
    u = np.random.default_rng(11).random((n, nshots))
    bits, probs = gmps.MPSContractionHelper(n).sample(conv.mps_tensors, nshots, uniforms=u)
    ref_bits, ref_probs = omps.sample(omps.evolve(circ, DEFAULT_ALGO), u)
    differing = int((bits != ref_bits).any(axis=1).sum())
    # a uniform within rounding of a threshold may fall on the other side in single precision
    assert differing <= (0 if dtype == "complex128" else 3)
    idx = (bits.astype(np.int64) << np.arange(n - 1, -1, -1)).sum(axis=1)
    assert np.abs(probs - np.abs(psi[idx]) ** 2).max() < TOL[dtype] * 10
    same = ~(bits != ref_bits).any(axis=1)
    assert np.abs(probs[same] - ref_probs[same]).max() < TOL[dtype] * 10
    emp = np.bincount(idx, minlength=1 << n) / nshots
    assert np.abs(emp - np.abs(psi) ** 2).max() < 0.03


def test_mps_sampling_at_48_qubits_through_the_quimb_platform():
    """GHZ state on 48 qubits: no dense vector can exist, every shot is all zeros or all ones, and the
    reported probabilities are |amplitude|^2 = 1/2 (reference backends/quimb.py:172-186)."""
    n = 48
    circ = circuits.Circuit(n)
    circ.add(circuits.gates.H(0))
    for q in range(n - 1):
        circ.add(circuits.gates.CNOT(q, q + 1))
    b = MetaBackend.load("quimb")
    b.configure_tn_simulation(ansatz="mps", max_bond_dimension=8, svd_cutoff=1e-12)
    b.sampling_seed = 7
    res = b.execute_circuit(circ, nshots=400)
    freq = res.frequencies()
    assert set(freq) == {"0" * n, "1" * n} and sum(freq.values()) == 400
    assert abs(freq["0" * n] - 200) < 60
    for state, p in res.probabilities().items():
        assert abs(p - 0.5) < 1e-10



def test_c3_full_size_against_the_cpu_oracle_fixture():
    """BASELINE config 3 at FULL size (64-site TFIM Trotter, 20 steps, bond <= 256, abs_cutoff 1e-12,
    complex128) against tests/golden/c3_mps.json, produced by scripts/make_golden_c3.py from the CPU oracle
    (numpy.linalg.svd gate loop, minutes of CPU): identical bond profile, <psi|psi> (0.914: 8.6 % of the weight
    is truncated away, so the truncation itself is under test) and <Z_i Z_{i+1}> at three bonds."""
    import json
    import os

    path = os.path.join(os.path.dirname(__file__), "golden", "c3_mps.json")
    if not os.path.exists(path):
        pytest.skip("tests/golden/c3_mps.json not generated yet (scripts/make_golden_c3.py)")
    ref = json.load(open(path))
    circ = circuits.tfim_c3(ref["n"], ref["steps"])
    algo = {"qr_method": False, "svd_method": {"partition": "UV", "abs_cutoff": 1e-12, "max_extent": ref["chi"]}}
    conv = gmps.QiboCircuitToMPS(circ, algo, dtype="complex128")
    helper = gmps.MPSContractionHelper(ref["n"])
    bonds = [int(t.shape[2]) for t in conv.mps_tensors[:-1]]
    assert bonds == ref["bonds"]
    norm = float(helper.contract_norm(conv.mps_tensors).cpu())
    assert abs(norm - ref["norm"]) <= 1e-10 * ref["norm"], (norm, ref["norm"])
    for i, (re, im) in ref["zz"].items():
        i = int(i)
        got = complex(helper.pauli_expectation(conv.mps_tensors, {i: "Z", i + 1: "Z"}).cpu())
        assert abs(got - complex(re, im)) <= 1e-10, (i, got, re, im)


# ----------------------------------------------------------------------------- dense initial state (config 1)
def _random_state(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    return v / np.linalg.norm(v)


def _qft_qasm(n):
    """QFT with swaps as qibo's ``QFT(n).to_qasm()`` writes it (h / cu1 / swap)."""
    import math
    lines = ['OPENQASM 2.0;', 'include "qelib1.inc";', f"qreg q[{n}];"]
    for i in range(n):
        lines.append(f"h q[{i}];")
        for j in range(i + 1, n):
            lines.append(f"cu1({math.pi / 2 ** (j - i)!r}) q[{j}],q[{i}];")
    for i in range(n // 2):
        lines.append(f"swap q[{i}],q[{n - 1 - i}];")
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 10, 13])
def test_from_dense_round_trip(n, dtype):
    """reference eval_qu.py:5-18: the MPS of a dense vector contracts back to the vector; shapes (l,2,r)."""
    psi = _random_state(n, 100 + n)
    sites = gmps.from_dense(psi, dtype)
    assert len(sites) == n and sites[0].shape[0] == 1 and sites[-1].shape[2] == 1
    for a, b in zip(sites[:-1], sites[1:]):
        assert a.shape[1] == 2 and a.shape[2] == b.shape[0] <= 1 << (n // 2)
    got = gmps.MPSContractionHelper(n).contract_state_vector(sites).reshape(-1).cpu().numpy()
    assert np.abs(got - psi).max() < TOL[dtype]


def test_from_dense_honours_the_bond_cap_and_cutoff():
    """A state of Schmidt rank <= 3 across every cut (sum of three product states) plus nothing else: the cutoff
    finds rank <= 3 on the decomposed half and the cap sweep brings the identity half down to it as well."""
    n = 12
    rng = np.random.default_rng(5)
    psi = np.zeros(1 << n, dtype=np.complex128)
    for _ in range(3):
        v = np.ones(1, dtype=np.complex128)
        for _q in range(n):
            v = np.kron(v, rng.standard_normal(2) + 1j * rng.standard_normal(2))
        psi += v
    psi /= np.linalg.norm(psi)
    algo = {"qr_method": False, "svd_method": {"partition": "UV", "abs_cutoff": 1e-9, "max_extent": 4}}
    sites = gmps.from_dense(psi, "complex128", algo)
    bonds = [int(t.shape[2]) for t in sites[:-1]]
    assert max(bonds) <= 4, bonds
    assert max(bonds[n // 2:]) <= 3, bonds
    got = gmps.MPSContractionHelper(n).contract_state_vector(sites).reshape(-1).cpu().numpy()
    assert np.abs(got - psi).max() < 1e-9
    with pytest.raises(ValueError):
        gmps.from_dense(psi[:-1], "complex128")
    with pytest.raises(NotImplementedError):
        gmps.MPSEngine(23, "complex128").load_dense(np.zeros(1 << 23, dtype=np.complex64))   # rows too long to split

@pytest.mark.parametrize("nqubits,tolerance,is_mps", [(1, 1e-6, True), (2, 1e-6, False), (5, 1e-3, True),
                                                      (10, 1e-3, False), (10, 1e-5, True), (6, 1e-6, False)])
def test_eval_qu_dense_vector_with_initial_state(nqubits, tolerance, is_mps):
    """The reference's own quimb test (tests/test_quimb_backend.py:23-66): QFT with swaps on a random normalised
    initial state, as QASM text, MPS options {"method": "svd", "cutoff": 1e-6, "cutoff_mode": "abs"} or None."""
    from qibotn_b200 import eval_qu
    from qibotn_b200.qasm import from_openqasm2_str

    import os
    closed = np.load(os.path.join(os.path.dirname(__file__), "golden", "closed_forms.npz"))
    # the committed closed form (QFT = normalised inverse FFT) at the reference's own four sizes
    init = closed[f"qft_in_{nqubits}"] if f"qft_in_{nqubits}" in closed else _random_state(nqubits, 7 * nqubits)
    text = _qft_qasm(nqubits)
    want = osv.simulate(from_openqasm2_str(text), initial_state=init)
    assert np.abs(want - osv.simulate(circuits.qft(nqubits, True), initial_state=init)).max() < 1e-12
    opts = {"method": "svd", "cutoff": 1e-6, "cutoff_mode": "abs"} if is_mps else None
    got = eval_qu.dense_vector_tn_qu(text, init.copy(), opts, backend="numpy").flatten()
    assert isinstance(got, np.ndarray) and got.shape == (1 << nqubits,)
    assert np.allclose(want, got, atol=min(tolerance, 1e-5 if is_mps else 1e-10))
    if f"qft_out_{nqubits}" in closed:
        assert np.allclose(closed[f"qft_out_{nqubits}"], got, atol=min(tolerance, 1e-5 if is_mps else 1e-10))
    none = eval_qu.dense_vector_tn_qu(text, None, opts, backend="torch")
    assert none.is_cuda and np.allclose(none.cpu().numpy(), osv.simulate(circuits.qft(nqubits, True)), atol=1e-5)
    sites = eval_qu.init_state_tn(nqubits, init)
    assert len(sites) == nqubits


def test_quimb_platform_executes_from_a_dense_initial_state():
    """reference backends/quimb.py:162-165: the MPS ansatz accepts a dense initial state; others refuse it."""
    n = 7
    init = _random_state(n, 3)
    circ = circuits.qft(n, True)
    b = MetaBackend.load("quimb")
    res = b.execute_circuit(circ, initial_state=init, return_array=True)
    got = np.asarray(res.statevector.get()).reshape(-1)
    assert np.abs(got - osv.simulate(circ, initial_state=init)).max() < 1e-8
    res = b.execute_circuit(circ, initial_state=init, nshots=200)
    assert sum(res.frequencies().values()) == 200
    with pytest.raises(ValueError):
        b.execute_circuit(circ, initial_state=init[:-2])
    b.configure_tn_simulation(ansatz=None)
    with pytest.raises(ValueError):
        b.execute_circuit(circ, initial_state=init)
