"""Tiny stand-ins for ``qibo.symbols.{I,X,Y,Z}`` and ``qibo.hamiltonians.SymbolicHamiltonian``.

The reference only reads ``hamiltonian.terms[*].coefficient`` and
``str(factor)`` (``"X12"``) from these objects
(``/root/reference/src/qibotn/eval.py:128-141``), so that is all that is
modelled: a sum of coefficient-times-Pauli-product terms.  No sympy.
"""

from __future__ import annotations

from numbers import Number


class _Factor:
    def __init__(self, letter: str, qubit: int):
        self.letter, self.target_qubit = letter, int(qubit)

    def __str__(self):
        return f"{self.letter}{self.target_qubit}"

    __repr__ = __str__


class Term:
    def __init__(self, coefficient, factors):
        self.coefficient = coefficient
        self.factors = list(factors)

    def __repr__(self):
        return f"{self.coefficient}*" + "*".join(map(str, self.factors))


class Form:
    """A sum of terms; supports ``+``, ``*`` and scalar multiplication."""

    def __init__(self, terms=()):
        self.terms = list(terms)

    @staticmethod
    def _lift(x):
        if isinstance(x, Form):
            return x
        if isinstance(x, Number):
            return Form([Term(x, [])])
        return NotImplemented

    def __add__(self, other):
        other = Form._lift(other)
        if other is NotImplemented:
            return other
        # ``0 + form`` (what ``sum()`` and ``form = 0; form += ...`` produce) adds nothing
        mine = [t for t in self.terms if t.factors or t.coefficient != 0]
        theirs = [t for t in other.terms if t.factors or t.coefficient != 0]
        return Form(mine + theirs)

    __radd__ = lambda self, other: Form._lift(other).__add__(self)

    def __neg__(self):
        return self * -1

    def __sub__(self, other):
        return self + (-Form._lift(other))

    def __mul__(self, other):
        other = Form._lift(other)
        if other is NotImplemented:
            return other
        return Form([Term(a.coefficient * b.coefficient, a.factors + b.factors)
                     for a in self.terms for b in other.terms])

    __rmul__ = lambda self, other: Form._lift(other).__mul__(self)


def _symbol(letter):
    def make(q):
        return Form([Term(1.0, [_Factor(letter, q)])])
    make.__name__ = letter
    return make


I, X, Y, Z = (_symbol(c) for c in "IXYZ")


class SymbolicHamiltonian:
    def __init__(self, form, nqubits=None):
        form = Form._lift(form)
        self.form = form
        self.terms = list(form.terms)
        qs = [f.target_qubit for t in self.terms for f in t.factors]
        self.nqubits = nqubits if nqubits is not None else (max(qs) + 1 if qs else 0)
