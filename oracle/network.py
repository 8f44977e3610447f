"""Circuit -> einsum-operand restatement (oracle; test infrastructure only).

Follows ``/root/reference/src/qibotn/circuit_convertor.py`` and the observable
helpers of ``/root/reference/src/qibotn/eval.py`` in numpy.  Operand lists use
the same *interleaved* convention the reference hands to
``cuquantum.tensornet.contract``: ``[T0, modes0, T1, modes1, ..., out_modes]``.
"""

import numpy as np

_KET0 = np.array([1.0, 0.0], dtype=np.complex128)
_KET1 = np.array([0.0, 1.0], dtype=np.complex128)
_SIGMA = {
    "I": np.array([[1, 0], [0, 1]], dtype=np.complex128),
    "X": np.array([[0, 1], [1, 0]], dtype=np.complex128),
    "Y": np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
    "Z": np.array([[1, 0], [0, -1]], dtype=np.complex128),
}


def gate_tensors(circuit, dtype=np.complex128):
    """``init_intermediate_circuit`` (circuit_convertor.py:98-125).

    Per gate: ``(matrix reshaped (2,2)*k, control_qubits + target_qubits)``;
    also the sorted list of qubits touched by at least one gate.  Measurement
    markers carry no matrix and are skipped (the reference would call
    ``matrix()`` on them).
    """
    tensors, touched = [], set()
    for g in circuit.queue:
        qs = tuple(g.control_qubits) + tuple(g.target_qubits)
        try:
            m = np.asarray(g.matrix(), dtype=dtype)
        except NotImplementedError:
            continue
        tensors.append((m.reshape((2, 2) * len(qs)), qs))
        touched.update(qs)
    return tensors, sorted(touched)


def _thread(gates, frontier, next_label):
    """``_parse_gates_to_mode_labels_operands`` (circuit_convertor.py:69-85):
    gate modes = fresh output labels followed by the current frontier labels."""
    ops, modes = [], []
    for t, qs in gates:
        ins = [frontier[q] for q in qs]
        outs = list(range(next_label, next_label + len(qs)))
        next_label += len(qs)
        for q, o in zip(qs, outs):
            frontier[q] = o
        ops.append(t)
        modes.append(outs + ins)
    return ops, modes, next_label


def state_vector_operands(circuit, dtype=np.complex128, bitstring=None):
    """``state_vector_operands`` (circuit_convertor.py:28-58).

    One |0> per *active* qubit (labels 0..n_active-1), the gates threaded on the
    frontier, output = frontier labels in ascending qubit order.  With
    ``bitstring`` (one character per active qubit) the network is closed with
    <bit| vectors instead and the output is a scalar: the amplitude close-out
    the reference lacks but ``_get_bitstring_tensors`` (``:66-67``) allows.
    """
    gts, active = gate_tensors(circuit, dtype)
    frontier = {q: i for i, q in enumerate(active)}
    ops = [_KET0.astype(dtype) for _ in active]
    modes = [[i] for i in range(len(active))]
    g_ops, g_modes, _ = _thread(gts, frontier, len(active))
    ops += g_ops
    modes += g_modes
    out = [frontier[q] for q in active]
    if bitstring is not None:
        assert len(bitstring) == len(active)
        for q, b in zip(active, bitstring):
            ops.append((_KET1 if b in ("1", 1) else _KET0).astype(dtype))
            modes.append([frontier[q]])
        out = []
    inter = [x for pair in zip(ops, modes) for x in pair]
    inter.append(out)
    return inter


def ham_gates(term, dtype=np.complex128):
    """``get_ham_gates`` (eval.py:86-111): ``[(coeff * sigma, (qubit,))]``."""
    gates = []
    for qubit, letter, coeff in term:
        if letter not in _SIGMA:
            raise ValueError("pauli string character must be one of I/X/Y/Z")
        gates.append(((coeff * _SIGMA[letter]).astype(dtype), (qubit,)))
    return gates


def expectation_operands(circuit, term_gates, dtype=np.complex128):
    """``expectation_operands`` (circuit_convertor.py:198-246).

    |0> on *all* ``nqubits``, forward gates, then the term's Pauli tensors and
    the inverse circuit threaded on the same frontier, closed by |0> again.
    Output is a scalar.  (Uniform ``dtype`` throughout; the reference leaves the
    inverse gates and Paulis in complex128)
    """
    n = circuit.nqubits
    gts, _ = gate_tensors(circuit, dtype)
    inv, _ = gate_tensors(circuit.invert(), dtype)
    frontier = {q: q for q in range(n)}
    ops = [_KET0.astype(dtype) for _ in range(n)]
    modes = [[q] for q in range(n)]
    g_ops, g_modes, nxt = _thread(gts, frontier, n)
    ops += g_ops
    modes += g_modes
    i_ops, i_modes, _ = _thread(list(term_gates) + inv, frontier, nxt)
    ops += i_ops
    modes += i_modes
    ops += [_KET0.astype(dtype) for _ in range(n)]
    modes += [[frontier[q]] for q in range(n)]
    inter = [x for pair in zip(ops, modes) for x in pair]
    inter.append([])
    return inter


# ----------------------------------------------------------------- observables
def default_observable_terms(n):
    """``build_observable`` (eval.py:28-35): sum_i 0.5 * X_i * Z_{(i+1) mod n}."""
    return [[(i % n, "X", 0.5), ((i + 1) % n, "Z", 1.0)] for i in range(n)]


def terms_from_dict(data, n):
    """``create_hamiltonian_from_dict`` + ``extract_gates_and_qubits``
    (eval.py:38-83,114-150): every term is padded with I on the qubits it does
    not name, factors ordered by qubit, coefficient carried by the first one.

    Also accepts the older ``{"pauli_string_pattern": "XZ"}`` spelling the
    README documents: the pattern is tiled cyclically over the n qubits, one
    term, coefficient 1.
    """
    if "pauli_string_pattern" in data and "terms" not in data:
        pat = data["pauli_string_pattern"]
        if not pat or any(c not in _SIGMA for c in pat):
            raise ValueError("pauli string character must be one of I/X/Y/Z")
        return [[(q, pat[q % len(pat)], 1.0) for q in range(n)]]
    out = []
    for term in data["terms"]:
        named = {int(q): g for g, q in term["operators"]}
        for g in named.values():
            if g not in ("X", "Y", "Z"):
                raise KeyError(g)
        facs = [(q, named.get(q, "I")) for q in range(n)]
        out.append([(q, g, term["coefficient"] if i == 0 else 1.0)
                    for i, (q, g) in enumerate(facs)])
    if not out:
        raise ValueError("No valid Hamiltonian terms were added.")
    return out


def terms_from_symbolic(ham):
    """``extract_gates_and_qubits`` (eval.py:114-150) for a SymbolicHamiltonian-like object."""
    out = []
    for term in ham.terms:
        coeff, facs = term.coefficient, []
        for f in term.factors:
            s = str(f)
            facs.append((int(s[1:]), s[0], coeff))
            coeff = 1.0
        out.append(facs)
    return out


def observable_terms(observable, n):
    """``check_observable`` (eval.py:15-25)."""
    if observable is None:
        return default_observable_terms(n)
    if isinstance(observable, dict):
        return terms_from_dict(observable, n)
    if hasattr(observable, "terms"):
        return terms_from_symbolic(observable)
    raise TypeError("Invalid observable type.")


def slice_range(num_slices, rank, size):
    """``compute_slices`` (eval.py:197-205): contiguous ranges, remainder to the low ranks."""
    chunk, extra = divmod(num_slices, size)
    lo = rank * chunk + min(rank, extra)
    hi = num_slices if rank == size - 1 else (rank + 1) * chunk + min(rank + 1, extra)
    return range(lo, hi)
