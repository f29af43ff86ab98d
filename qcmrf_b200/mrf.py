"""Host-side mirror of the reference's circuit constructor and result helpers.

The GPU box has no /root/reference, so the product carries its own constructor with
the reference's interface -- same class name, argument meaning, properties and
error conditions as ``QCMRF`` in /root/reference/QCMRF.py:13-245 -- written against
this package's circuit layer.  The gate program it records is the one
``QCMRF._build`` emits (QCMRF.py:199-243); tests pin that equality against programs
captured from the reference module itself (tests/golden/ref_programs.json).

``extract_probs``, ``fidelity`` and ``KL`` mirror QCMRF.py:247-284 for callers that
post-select sampled counts on the host; the exact, on-GPU version of that
post-selection is ``Result.postselected_probabilities``.
"""
import itertools
import math

import numpy as np

from .circuit import AND, QuantumCircuit

__all__ = ['QCMRF', 'extract_probs', 'fidelity', 'KL']


def _is_clique_list(cliques):
    return (type(cliques) is list and len(cliques) > 0 and type(cliques[0]) is list and
            len(cliques[0]) > 0 and type(cliques[0][0]) is int)


_BITREV = {}


def _bit_reverse(m):
    """Index array r with r[i] = i with its m bits reversed."""
    r = _BITREV.get(m)
    if r is None:
        i = np.arange(1 << m)
        r = np.zeros(1 << m, dtype=np.int64)
        for j in range(m):
            r |= ((i >> j) & 1) << (m - 1 - j)
        _BITREV[m] = r
    return r


def _bit_reverse_inv(m):
    """r with table[t] = values[r[t]]: the theta index whose bits are t reversed (its own inverse)."""
    return _bit_reverse(m)


_H0 = np.array([1.0, 1.0], dtype=np.complex128) / math.sqrt(2.0)
_H0.setflags(write=False)


class QCMRF(QuantumCircuit):
    """Quantum-circuit Markov random field over binary variables.

    Register: n variable qubits (variable v on qubit n-1-v), one scratch qubit n for
    the clique-state AND, one ancilla per clique (qubit n+1+ii); as many clbits.
    Measuring all ancillas as 0 leaves the variables distributed as
    p(x) ~ exp(beta * sum_C theta[C, x_C]).
    """

    def __init__(self, cliques=None, theta=None, gamma=None, beta: float = 1, name: str = 'QCMRF',
                 with_measurements=True, with_barriers=False,
                 basis_gates=('cx', 'id', 'rz', 'sx', 'x')):
        if not _is_clique_list(cliques):
            raise ValueError('cliques must be a non-empty list of lists of int')
        self._mrf_cliques = cliques
        self._mrf_beta = beta
        self._mrf_measure = with_measurements
        self._mrf_barriers = with_barriers
        self.basis_gates = list(basis_gates)
        self._mrf_n = max(max(C) for C in cliques) + 1
        self._mrf_dim = sum(2 ** len(C) for C in cliques)
        self._mrf_cmax = max(len(C) for C in cliques)
        for label, vec in (('theta', theta), ('gamma', gamma)):
            if vec is not None and len(vec) != self._mrf_dim:
                raise ValueError('%s has %d entries, the clique structure needs %d' % (label, len(vec), self._mrf_dim))
        self._mrf_theta = None if theta is None else [float(t) for t in theta]
        self._mrf_gamma = None if gamma is None else [float(g) for g in gamma]
        width = self._mrf_n + len(cliques) + 1
        super().__init__(width, width, name=name)
        if self._mrf_theta is None and self._mrf_gamma is None:
            # the reference draws theta ~ U(-5, 0) from numpy's global state while building
            self._mrf_theta = [float(np.random.uniform(low=-5.0, high=0)) for _ in range(self._mrf_dim)]
        # The instruction list (nested cU_C / AND objects, what `.data`, `transpile` and a generic
        # backend walk) is materialised on first access; the engine's own lowering takes the flat
        # gate program of `_lower_program` -- the same gates, without building ~700 objects.
        self.__dict__['_mrf_data'] = None

    @property
    def _qc_data(self):
        d = self.__dict__.get('_mrf_data')
        if d is None:
            d = self.__dict__['_mrf_data'] = []
            self._emit_program()
        return d

    @_qc_data.setter
    def _qc_data(self, value):
        self.__dict__['_mrf_data'] = value

    def _lower_program(self):
        """Flat primitive gate program of this circuit (ir.Program), or None once the instruction
        list has been materialised (it may have been edited; the generic walk is used then).
        Emits exactly what ir.lower finds in `.data` (tests/test_circuit_api.py pins the equality)."""
        if self.__dict__.get('_mrf_data') is not None:
            return None
        from .ir import LazyProgram
        n = self._mrf_n
        width = self.num_qubits
        prog = LazyProgram(width, width, name=str(self.name), build=self._emit_gates)
        if self._mrf_measure:
            for ii in range(len(self._mrf_cliques)):
                prog.measures[n + 1 + ii] = n + 1 + ii
            for q in range(n):
                prog.measures[q] = q
        prog.metadata['num_vertices'] = n
        prog.fused_hint = self._fused_circuit
        return prog

    def _emit_gates(self, prog):
        """The gate list of `_lower_program`'s Program (built on first access)."""
        from .ir import Gate
        n = self._mrf_n
        gates = prog.gates
        for q in range(n):
            gates.append(Gate('h', (q,)))
        gam = self.gamma
        if not all(math.isfinite(g) for g in gam):
            raise ValueError('QCMRF: theta must be <= 0 (gamma is not finite)')
        offset = 0
        for ii, C in enumerate(self._mrf_cliques):
            anc = n + 1 + ii
            m = len(C)
            wires = tuple(n - 1 - v for v in C) + (n,)
            fwd = []
            for y, g in zip(itertools.product((0, 1), repeat=m), gam[offset:offset + 2 ** m]):
                if abs(g) <= 1e-8:                              # np.isclose(g, 0)
                    continue
                mark = Gate('mcx', wires, (), tuple(y))
                fwd.append((mark, 2.0 * g))
            offset += 2 ** m
            gates.append(Gate('h', (anc,)))
            for mark, lam in fwd:
                gates.append(mark)
                gates.append(Gate('cp', (n, anc), (lam,), (1,)))
                gates.append(mark)
            gates.append(Gate('x', (anc,)))
            for mark, lam in reversed(fwd):
                gates.append(mark)
                gates.append(Gate('cp', (n, anc), (-lam,), (1,)))
                gates.append(mark)
            gates.append(Gate('x', (anc,)))
            gates.append(Gate('h', (anc,)))

    def _fused_circuit(self):
        """What the gate-fusion pass makes of this circuit, written down directly:
        H|0> on the variable qubits, then per clique ONE uniformly-controlled RX(4 gamma) on its
        ancilla -- [[cos 2g, -i sin 2g], [-i sin 2g, cos 2g]] selected by the clique's variable qubits;
        terms the constructor skips (gamma ~ 0, QCMRF.py:223) are the identity; the scratch qubit is
        back on |0> and disappears.  Saves the numeric classification of ~12 gates per clique on the
        host; tests/test_host_fusion.py pins it against fusion.fuse on the gate list.  Returns None
        where the shortcut does not apply (edited circuit, repeated vertices, an all-skipped clique)."""
        if self.__dict__.get('_mrf_data') is not None:
            return None
        from .fusion import FusedCircuit, FusedOp
        n = self._mrf_n
        if self._mrf_gamma is not None:
            gam = np.asarray(self._mrf_gamma, dtype=np.float64)
        else:                                          # the `gamma` property, vectorised (same ufuncs, whole array)
            with np.errstate(invalid='ignore'):
                gam = 0.5 * np.arccos(np.exp(self._mrf_beta * 0.5 * np.asarray(self._mrf_theta, dtype=np.float64)))
        if not np.isfinite(gam).all():
            # theta > 0 has no circuit angle (arccos of a value > 1, QCMRF.py:154): the reference emits cp(nan)
            raise ValueError('QCMRF: theta must be <= 0 (gamma is not finite for %d parameter(s))'
                             % int((~np.isfinite(gam)).sum()))
        keep_all = np.abs(gam) > 1e-8
        c_all = np.where(keep_all, np.cos(2.0 * gam), 1.0)
        s_all = np.where(keep_all, np.sin(2.0 * gam), 0.0) * -1j
        st = self.__dict__.get('_mrf_struct')
        if st is None:
            # per clique size m: gather indices into gamma, already in table order (theta index
            # i = sum_j y_j 2^(m-1-j); table index = sum_j y_j 2^j, bit j <-> j-th listed vertex)
            by_m, offset = {}, 0
            for ii, C in enumerate(self._mrf_cliques):
                m = len(C)
                if len(set(C)) != m:
                    st = False
                    break
                by_m.setdefault(m, ([], []))
                by_m[m][0].append(ii)
                by_m[m][1].append(offset + _bit_reverse_inv(m))
                offset += 1 << m
            if st is None:
                st = [(m, ids, np.stack(idx)) for m, (ids, idx) in by_m.items()]
            self.__dict__['_mrf_struct'] = st
        if st is False:
            return None
        ops = [None] * len(self._mrf_cliques)
        n_gates = n
        for m, ids, idx in st:
            kept = keep_all[idx].sum(axis=1)
            if not kept.all():
                return None
            T = np.empty((len(ids), 1 << m, 2, 2), dtype=np.complex128)
            T[:, :, 0, 0] = T[:, :, 1, 1] = c_all[idx]
            T[:, :, 0, 1] = T[:, :, 1, 0] = s_all[idx]
            for r, ii in enumerate(ids):
                k = int(kept[r])
                ops[ii] = FusedOp('mux', n + 1 + ii, tuple(n - 1 - v for v in self._mrf_cliques[ii]), T[r], zero_in=True,
                                  n_gates=4 + 2 * k)
                n_gates += 4 + 6 * k
        h0 = _H0
        return FusedCircuit(self.num_qubits, {q: h0 for q in range(n)}, ops, 0.0, n_gates)

    # -- the reference's read-only surface (QCMRF.py:82-157) ------------------------------
    @property
    def dimension(self):
        return self._mrf_dim

    @property
    def cliques(self):
        return self._mrf_cliques

    @property
    def num_vertices(self):
        return self._mrf_n

    num_nodes = num_vertices

    @property
    def num_cliques(self):
        return len(self._mrf_cliques)

    @property
    def max_clique(self):
        return self._mrf_cmax

    @property
    def beta(self):
        return self._mrf_beta

    @property
    def theta(self):
        """theta_i = 2 ln(cos 2 gamma_i) / beta when only gamma was given."""
        if self._mrf_theta is None:
            self._mrf_theta = [2.0 * math.log(math.cos(2.0 * g)) / self._mrf_beta for g in self._mrf_gamma]
        return self._mrf_theta

    @property
    def gamma(self):
        """gamma_i = arccos(exp(beta theta_i / 2)) / 2; theta > 0 has no circuit angle (NaN,
        as in the reference, which only ever draws theta <= 0)."""
        if self._mrf_gamma is None:
            with np.errstate(invalid='ignore'):
                self._mrf_gamma = [float(0.5 * np.arccos(np.exp(self._mrf_beta * 0.5 * t))) for t in self._mrf_theta]
        return self._mrf_gamma

    # -- program ---------------------------------------------------------------------------
    def _clique_unitary(self, index, clique, angles):
        """cU_C: on (variables..., scratch, ancilla) -- for every clique state y add the
        phase 2*gamma_{C,y} to ancilla=1 where x_C == y, via AND / CP / AND."""
        n = self._mrf_n
        block = QuantumCircuit(n + 2, name='cU_C%d' % index)
        wires = [n - 1 - v for v in clique] + [n]
        for y, g in zip(itertools.product((0, 1), repeat=len(clique)), angles):
            if abs(g) <= 1e-8:                                  # np.isclose(g, 0) (QCMRF.py:223)
                continue
            marker = AND(len(clique), [2 * b - 1 for b in y])
            block.append(marker, wires)
            block.cp(2 * g, n, n + 1)
            block.append(marker, wires)
        return block

    def _emit_program(self):
        n = self._mrf_n
        for q in range(n):
            self.h(q)
        if self._mrf_barriers:
            self.barrier()
        gam = self.gamma
        offset = 0
        main = list(range(n + 1))
        for ii, C in enumerate(self._mrf_cliques):
            anc = n + 1 + ii
            block = self._clique_unitary(ii, C, gam[offset:offset + 2 ** len(C)])
            offset += 2 ** len(C)
            self.h(anc)
            self.append(block, main + [anc])
            self.x([anc])
            self.append(block.inverse(), main + [anc])
            self.x([anc])
            self.h(anc)
            if self._mrf_measure:
                self.measure(anc, anc)
            if self._mrf_barriers:
                self.barrier()
        if self._mrf_measure:
            self.measure(range(n), range(n))


def fidelity(P, Q):
    """(sum_i sqrt(P_i Q_i))^2 over entries where both are positive."""
    P = np.asarray(P, dtype=np.float64)
    Q = np.asarray(Q, dtype=np.float64)
    both = (P > 0) & (Q > 0)
    return float(np.sqrt(P[both] * Q[both]).sum() ** 2)


def KL(P, Q):
    """sum_i P_i ln(P_i / Q_i) over entries where both are positive."""
    P = np.asarray(P, dtype=np.float64)
    Q = np.asarray(Q, dtype=np.float64)
    both = (P > 0) & (Q > 0)
    return float((P[both] * np.log(P[both] / Q[both])).sum())


def extract_probs(R, n, a):
    """Post-select a counts dict: keep keys '0'*a + x_0..x_{n-1}.  Returns
    (pmf over x with x_0 as MSB, kept fraction), or (zeros, 0) when nothing survives."""
    P = np.zeros(2 ** n)
    prefix = '0' * a
    for key, val in R.items():
        if len(key) == a + n and key.startswith(prefix):
            P[int(key[a:], 2)] += val
    kept = P.sum()
    if kept == 0:
        return P, 0
    return P / kept, kept / sum(R.values())
