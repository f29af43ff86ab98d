"""``from qiskit.opflow import I, Z`` (QCMRF.py:6) executes at import time; the
operators are only used by the reference's Hamiltonian helpers (QCMRF.py:159-197),
which are outside the simulator hot path.  A small Pauli-sum algebra keeps those
helpers working: ^ tensor, + - * / scalars, ~ adjoint."""
import numpy as np


class PauliSum:
    def __init__(self, terms):
        self.terms = {k: complex(v) for k, v in terms.items() if abs(v) > 0}

    @property
    def num_qubits(self):
        return len(next(iter(self.terms))) if self.terms else 0

    def _lift(self, other):
        if isinstance(other, PauliSum):
            return other
        if other == 0:
            return PauliSum({})
        raise TypeError('cannot combine PauliSum with %r' % (other,))

    def __add__(self, other):
        other = self._lift(other)
        out = dict(self.terms)
        for k, v in other.terms.items():
            out[k] = out.get(k, 0) + v
        return PauliSum(out)

    __radd__ = __add__

    def __neg__(self):
        return PauliSum({k: -v for k, v in self.terms.items()})

    def __sub__(self, other):
        return self + (-self._lift(other))

    def __rsub__(self, other):
        return self._lift(other) + (-self)

    def __mul__(self, s):
        return PauliSum({k: v * s for k, v in self.terms.items()})

    __rmul__ = __mul__

    def __truediv__(self, s):
        return PauliSum({k: v / s for k, v in self.terms.items()})

    def __xor__(self, other):                       # tensor product, self on the left
        if not isinstance(other, PauliSum):
            if other == 1:
                return self
            raise TypeError('tensor with %r' % (other,))
        return PauliSum({a + b: va * vb for a, va in self.terms.items() for b, vb in other.terms.items()})

    def __rxor__(self, other):
        if other == 1:
            return self
        raise TypeError('tensor with %r' % (other,))

    def __invert__(self):                           # adjoint (Paulis are Hermitian)
        return PauliSum({k: np.conj(v) for k, v in self.terms.items()})

    adjoint = __invert__

    def to_matrix(self):
        P = {'I': np.eye(2), 'X': np.array([[0, 1], [1, 0]]), 'Y': np.array([[0, -1j], [1j, 0]]),
             'Z': np.diag([1.0, -1.0])}
        out = 0
        for k, v in self.terms.items():
            m = np.ones((1, 1))
            for ch in k:
                m = np.kron(m, P[ch])
            out = out + v * m
        return out

    def __repr__(self):
        return ' + '.join('%s*%s' % (v, k) for k, v in sorted(self.terms.items())) or '0'


I = PauliSum({'I': 1})
X = PauliSum({'X': 1})
Y = PauliSum({'Y': 1})
Z = PauliSum({'Z': 1})
