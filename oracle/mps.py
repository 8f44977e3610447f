"""MPS gate loop restatement (oracle; test infrastructure only).

Follows ``/root/reference/src/qibotn/mps_utils.py`` and
``mps_contraction_helper.py`` with ``numpy.linalg.svd`` in place of cuQuantum's
``contract_decompose`` (cuquantum-python 25.11.1, not installable here).  The
truncation rules are cuTensorNet's published ``SVDMethod`` semantics
(SURVEY.md Appendix B): drop singular values below ``abs_cutoff`` or below
``rel_cutoff * s_max``, cap the bond at ``max_extent``, optionally renormalise,
then absorb S according to ``partition`` ("UV": sqrt(S) into both sides).
No canonical form is kept: truncation is local to the two-site block, as in the
reference.
"""

import numpy as np


def initial(n, dtype=np.complex128):
    """``initial`` (mps_utils.py:6-18): n copies of |0> shaped (1,2,1)."""
    return [np.array([1, 0], dtype=dtype).reshape(1, 2, 1) for _ in range(n)]


def truncation_rank(s, svd_method):
    """How many singular values survive (s is sorted, descending)."""
    keep = len(s)
    if svd_method:
        ac = svd_method.get("abs_cutoff", 0.0) or 0.0
        rc = svd_method.get("rel_cutoff", 0.0) or 0.0
        thr = max(ac, rc * (s[0] if len(s) else 0.0))
        if thr > 0:
            keep = int(np.count_nonzero(s > thr))
        dw = svd_method.get("discarded_weight_cutoff", 0.0) or 0.0
        if dw > 0 and len(s):
            w = s[:keep] ** 2
            tail = np.cumsum(w[::-1])[::-1] / np.sum(s ** 2)     # weight of s[i:]
            while keep > 1 and tail[keep - 1] < dw:
                keep -= 1
        me = svd_method.get("max_extent")
        if me:
            keep = min(keep, int(me))
    return max(keep, 1)


def split(theta, algorithm):
    """SVD-split a matrix ``theta`` (rows = left legs, cols = right legs) the way
    ``contract_decompose(..., algorithm=algorithm)`` does; returns ``(a, s, b)``
    with ``s`` None when it was absorbed (mps_utils.py:32,78 unpack it so)."""
    svd_method = (algorithm or {}).get("svd_method", {})
    if svd_method is False:
        q, r = np.linalg.qr(theta)
        return q, None, r
    u, s, vh = np.linalg.svd(theta, full_matrices=False)
    k = truncation_rank(s, svd_method)
    u, s, vh = u[:, :k], s[:k], vh[:k, :]
    norm = svd_method.get("normalization") if svd_method else None
    if norm == "L1":
        s = s / np.sum(s)
    elif norm == "L2":
        s = s / np.sqrt(np.sum(s ** 2))
    elif norm == "LInf":
        s = s / s[0]
    part = svd_method.get("partition") if svd_method else None
    if part == "U":
        return u * s, None, vh
    if part == "V":
        return u, None, s[:, None] * vh
    if part == "UV":
        r = np.sqrt(s)
        return u * r, None, r[:, None] * vh
    return u, s, vh


def two_site(a, b, gate, algorithm):
    """``contract_decompose("ipj,jqk,rspq->irj,jsk", A, B, G)`` (mps_utils.py:78-84)."""
    theta = np.einsum("ipj,jqk,rspq->irsk", a, b, gate)
    l, _, _, r = theta.shape
    x, _, y = split(theta.reshape(2 * l, 2 * r), algorithm)
    return x.reshape(l, 2, -1), y.reshape(-1, 2, r)


def right_swap(mps, i, algorithm):
    """``mps_site_right_swap`` (mps_utils.py:21-39): "ipj,jqk->iqj,jpk"."""
    theta = np.einsum("ipj,jqk->iqpk", mps[i], mps[i + 1])
    l, _, _, r = theta.shape
    x, _, y = split(theta.reshape(2 * l, 2 * r), algorithm)
    mps[i], mps[i + 1] = x.reshape(l, 2, -1), y.reshape(-1, 2, r)


def apply_gate(mps, gate, qubits, algorithm):
    """``apply_gate`` (mps_utils.py:42-95)."""
    if len(qubits) == 1:
        i = qubits[0]
        mps[i] = np.einsum("ipj,qp->iqj", mps[i], gate)
    elif len(qubits) == 2:
        i, j = qubits
        if i > j:
            return apply_gate(mps, gate.transpose(1, 0, 3, 2), (j, i), algorithm)
        if i + 1 == j:
            mps[i], mps[j] = two_site(mps[i], mps[j], gate, algorithm)
        else:
            right_swap(mps, i, algorithm)
            apply_gate(mps, gate, (i + 1, j), algorithm)
            right_swap(mps, i, algorithm)
    else:
        raise NotImplementedError("Only one- and two-qubit gates supported")
    return mps


def evolve(circuit, algorithm, dtype=np.complex128):
    """``QiboCircuitToMPS.__init__`` (circuit_to_mps.py:19-44)."""
    from .network import gate_tensors

    mps = initial(circuit.nqubits, dtype)
    gts, _ = gate_tensors(circuit, dtype)
    for g, qs in gts:
        apply_gate(mps, g, tuple(qs), algorithm)
    return mps


def to_dense(mps):
    """``contract_state_vector`` (mps_contraction_helper.py:53-71): (2,)*n array."""
    psi = mps[0]
    for t in mps[1:]:
        psi = np.tensordot(psi, t, axes=([psi.ndim - 1], [0]))
    return psi.reshape(psi.shape[1:-1])


def norm(mps):
    """``contract_norm`` (mps_contraction_helper.py:32-51): <psi|psi>, real."""
    env = np.ones((1, 1), dtype=mps[0].dtype)
    for t in mps:
        env = np.einsum("ab,apc,bpd->cd", env, t, t.conj())
    return env.reshape(()).real


def expectation(mps, operator, qubits, normalize=False):
    """``contract_expectation`` (mps_contraction_helper.py:73-113).  ``operator``
    has modes (outputs..., inputs...) on ``qubits`` (1 or 2 adjacent or not)."""
    qubits = list(qubits)
    k = len(qubits)
    op = np.asarray(operator).reshape((2,) * (2 * k))
    n = len(mps)
    # dense fallback via transfer matrices with the operator legs kept open
    # (small k only): env carries (bra bond, ket bond, pending operator legs)
    if k == 1:
        q = qubits[0]
        env = np.ones((1, 1), dtype=mps[0].dtype)
        for i, t in enumerate(mps):
            if i == q:
                env = np.einsum("ab,apc,bqd,qp->cd", env, t, t.conj(), op)
            else:
                env = np.einsum("ab,apc,bpd->cd", env, t, t.conj())
        val = env.reshape(())
    else:
        psi = to_dense(mps).reshape(-1)
        from .statevector import apply_matrix
        phi = apply_matrix(psi, op.reshape(1 << k, 1 << k), qubits, n)
        # reference contracts bra = tensors, ket = conj(tensors): value is <phi|psi>^*-style;
        # for Hermitian operators both orders agree.
        val = np.vdot(psi, phi)
    return val / (norm(mps) if normalize else 1)


def right_environments(mps):
    """R[i] = sum_p A_i^p R[i+1] (A_i^p)^H, R[n] = [[1]] (test infrastructure, numpy)."""
    n = len(mps)
    envs = [None] * (n + 1)
    envs[n] = np.ones((1, 1), dtype=np.complex128)
    for i in range(n - 1, -1, -1):
        a = mps[i].astype(np.complex128)
        envs[i] = np.einsum("ipj,jk,lpk->il", a, envs[i + 1], a.conj())
    return envs


def sample(mps, uniforms):
    """Perfect sampling of an (unnormalised) MPS, the CPU restatement of what quimb's
    ``Circuit.sample`` delivers for the reference's ``execute_circuit(nshots=...)``
    (backends/quimb.py:172-186): qubit by qubit from the left, bit = 1 iff u >= P(0 | earlier bits).
    ``uniforms``: (n, nshots) in [0, 1).  Returns (bits nshots x n uint8, |amplitude|^2 per shot)."""
    n = len(mps)
    uniforms = np.asarray(uniforms, dtype=np.float64)
    nshots = uniforms.shape[1]
    envs = right_environments(mps)
    norm2 = envs[0][0, 0].real
    bits = np.zeros((nshots, n), dtype=np.uint8)
    probs = np.ones(nshots)
    for s in range(nshots):
        v = np.ones((1,), dtype=np.complex128) / np.sqrt(norm2)
        for i in range(n):
            a = mps[i].astype(np.complex128)
            w = [v @ a[:, p, :] for p in range(2)]
            q = [max(float(np.real(np.vdot(wp, wp @ envs[i + 1]).conjugate())), 0.0) for wp in w]
            tot = q[0] + q[1]
            bit = int(uniforms[i, s] * tot >= q[0]) if tot > 0 else 0
            bits[s, i] = bit
            probs[s] *= q[bit] / tot if tot > 0 else 0.0
            v = w[bit] / np.sqrt(q[bit]) if q[bit] > 0 else w[bit] * 0
    return bits, probs * norm2

