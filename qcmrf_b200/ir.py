"""Flat gate-program IR: what the backend lowers a circuit to before fusion.

A circuit object (the compat shim's ``QuantumCircuit``, the product's ``QCMRF``, or
a real Qiskit circuit when Qiskit is importable) is flattened -- nested
instructions expanded, ``.inverse()`` already applied by the circuit layer -- into
a ``Program``: primitive gates on integer qubits, a clbit->qubit measurement map
and a global phase.  The reference hands Aer exactly such a flat list after
``transpile`` (/root/reference/run_experiment.py:52-56).

Every primitive gate is either a (multi-)controlled single-qubit unitary or a
diagonal phase; ``Gate.matrix()`` gives its dense unitary on ``Gate.qubits``
(little-endian: bit j of the matrix index <-> qubits[j]).
"""
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

_SQ2 = 1.0 / np.sqrt(2.0)
_ONEQ = {
    'id': np.eye(2, dtype=np.complex128),
    'h': np.array([[_SQ2, _SQ2], [_SQ2, -_SQ2]], dtype=np.complex128),
    'x': np.array([[0, 1], [1, 0]], dtype=np.complex128),
    'y': np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
    'z': np.array([[1, 0], [0, -1]], dtype=np.complex128),
    's': np.array([[1, 0], [0, 1j]], dtype=np.complex128),
    'sdg': np.array([[1, 0], [0, -1j]], dtype=np.complex128),
    't': np.array([[1, 0], [0, np.exp(0.25j * np.pi)]], dtype=np.complex128),
    'tdg': np.array([[1, 0], [0, np.exp(-0.25j * np.pi)]], dtype=np.complex128),
    'sx': 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]], dtype=np.complex128),
    'sxdg': 0.5 * np.array([[1 - 1j, 1 + 1j], [1 + 1j, 1 - 1j]], dtype=np.complex128),
}
_ALIASES = {'i': 'id', 'u1': 'p', 'cnot': 'cx', 'toffoli': 'ccx', 'cu1': 'cp', 'mcu1': 'mcp',
            'mcphase': 'mcp', 'mcx_gray': 'mcx', 'c3x': 'mcx', 'c4x': 'mcx', 'ccx': 'mcx'}
_PARAM1Q = ('rz', 'rx', 'ry', 'p')
#: base single-qubit gate applied by each controlled primitive
_CTRL_BASE = {'cx': 'x', 'cy': 'y', 'cz': 'z', 'ch': 'h', 'cp': 'p', 'crz': 'rz', 'crx': 'rx',
              'cry': 'ry', 'mcx': 'x', 'mcp': 'p', 'csx': 'sx'}


_PARAM_CACHE = {}


def one_qubit_matrix(name, params=()):
    """2x2 matrix of a primitive 1-qubit gate.  Returned arrays are shared: treat them as read-only."""
    if name in _ONEQ:
        return _ONEQ[name]
    key = (name, params)
    try:
        hit = _PARAM_CACHE.get(key)
    except TypeError:                               # unhashable params (a list): no caching
        return _one_qubit_matrix(name, params)
    if hit is None:
        hit = _one_qubit_matrix(name, params)
        hit.setflags(write=False)
        if len(_PARAM_CACHE) > 65536:
            _PARAM_CACHE.clear()
        _PARAM_CACHE[key] = hit
    return hit


def _one_qubit_matrix(name, params=()):
    lam = float(params[0]) if params else 0.0
    if name == 'rz':
        return np.array([[np.exp(-0.5j * lam), 0], [0, np.exp(0.5j * lam)]])
    if name == 'p':
        return np.array([[1, 0], [0, np.exp(1j * lam)]])
    if name == 'rx':
        c, s = np.cos(lam / 2), np.sin(lam / 2)
        return np.array([[c, -1j * s], [-1j * s, c]])
    if name == 'ry':
        c, s = np.cos(lam / 2), np.sin(lam / 2)
        return np.array([[c, -s], [s, c]], dtype=np.complex128)
    if name == 'u':
        th, ph, lm = (float(p) for p in params)
        c, s = np.cos(th / 2), np.sin(th / 2)
        return np.array([[c, -np.exp(1j * lm) * s],
                         [np.exp(1j * ph) * s, np.exp(1j * (ph + lm)) * c]])
    raise ValueError('qcmrf_b200: unsupported gate %r' % name)


class Gate:
    """``base`` single-qubit gate on ``qubits[-1]``, applied where every control
    ``qubits[j]`` equals ``ctrl_values[j]`` (no controls => plain 1-qubit gate).
    Immutable by convention; a slotted plain class because programs hold hundreds of them
    and are rebuilt for every circuit."""
    __slots__ = ('name', 'qubits', 'params', 'ctrl_values')

    def __init__(self, name, qubits, params=(), ctrl_values=()):
        self.name = name                # canonical primitive name
        self.qubits = qubits            # controls..., target
        self.params = params
        self.ctrl_values = ctrl_values

    def _key(self):
        return (self.name, self.qubits, self.params, self.ctrl_values)

    def __eq__(self, other):
        return isinstance(other, Gate) and self._key() == other._key()

    def __hash__(self):
        return hash(self._key())

    def __repr__(self):
        return 'Gate(%r, %r, %r, %r)' % self._key()

    @property
    def target(self):
        return self.qubits[-1]

    @property
    def controls(self):
        return self.qubits[:-1]

    def base_matrix(self):
        base = _CTRL_BASE.get(self.name, self.name) if self.controls else self.name
        return one_qubit_matrix(base, self.params)

    def matrix(self):
        """Dense unitary on self.qubits, index bit j <-> qubits[j]."""
        k = len(self.qubits)
        U = np.eye(1 << k, dtype=np.complex128)
        B = self.base_matrix()
        i0 = sum(v << j for j, v in enumerate(self.ctrl_values))
        i1 = i0 | (1 << (k - 1))
        U[i0, i0], U[i0, i1], U[i1, i0], U[i1, i1] = B[0, 0], B[0, 1], B[1, 0], B[1, 1]
        return U

    def is_diagonal(self):
        B = self.base_matrix()
        return B[0, 1] == 0 and B[1, 0] == 0


@dataclass
class Program:
    n_qubits: int
    n_clbits: int
    gates: List[Gate] = field(default_factory=list)
    measures: Dict[int, int] = field(default_factory=dict)     # clbit -> qubit
    global_phase: float = 0.0
    name: str = ''
    metadata: dict = field(default_factory=dict)
    #: optional shortcut of the producing circuit class: a callable returning the FusedCircuit that
    #: fusion.fuse(self, 'clique') would compute from `gates`, or None (then the gates are fused)
    fused_hint: object = field(default=None, compare=False, repr=False)


class LazyProgram(Program):
    """A Program whose gate list is produced on first access (``build(self)`` appends to it): a circuit
    class that also supplies `fused_hint` usually never needs the gates at all."""

    def __init__(self, n_qubits, n_clbits, name='', build=None):
        Program.__init__(self, n_qubits, n_clbits, name=name)
        self._build = build
        self._gates = None

    @property
    def gates(self):
        if self._gates is None:
            self._gates = []
            if self._build is not None:
                self._build(self)
        return self._gates

    @gates.setter
    def gates(self, value):
        self._gates = value


BASIS_KINDS = ('rz', 'sx', 'x', 'id', 'cx')          # kind codes of a basis-gate program, in this order


class BasisProgram(Program):
    """A program in the reference's target basis (cx, id, rz, sx, x -- run_experiment.py:52) held as flat arrays
    instead of one Gate object per gate: ``bk`` kind code (index into BASIS_KINDS), ``bq`` target qubit, ``bc``
    control qubit (cx) or -1, ``bp`` angle (rz).  A transpiled fixture circuit has up to ~15 000 gates; the fusion pass
    walks the arrays in C (fusion._fuse_basis).  ``gates`` materialises the Gate list on first access for every
    other consumer."""

    def __init__(self, n_qubits, n_clbits, bk, bq, bc, bp, name='', global_phase=0.0):
        Program.__init__(self, n_qubits, n_clbits, name=name, global_phase=global_phase)
        self.bk = np.ascontiguousarray(bk, dtype=np.int8)
        self.bq = np.ascontiguousarray(bq, dtype=np.int32)
        self.bc = np.ascontiguousarray(bc, dtype=np.int32)
        self.bp = np.ascontiguousarray(bp, dtype=np.float64)
        self._gates = None

    def gate_at(self, i):
        k, q = int(self.bk[i]), int(self.bq[i])
        if k == 0:
            return Gate('rz', (q,), (float(self.bp[i]),))
        if k == 4:
            return Gate('cx', (int(self.bc[i]), q), (), (1,))
        return Gate(BASIS_KINDS[k], (q,))

    @property
    def gates(self):
        if self._gates is None:
            self._gates = [self.gate_at(i) for i in range(len(self.bk))]
        return self._gates

    @gates.setter
    def gates(self, value):
        self._gates = value


def _qindex(circ, q):
    if isinstance(q, (int, np.integer)):
        return int(q)
    return circ.find_bit(q).index               # real Qiskit Bit objects


def _emit(prog, name, qubits, params, ctrl_values=None):
    name = name.lower()
    canon = _ALIASES.get(name, name)
    qubits = tuple(int(q) for q in qubits)
    if canon in ('barrier', 'delay'):
        return
    if len(set(qubits)) != len(qubits):
        raise ValueError('duplicate qubit arguments in %s%r' % (name, qubits))
    if prog.measures:
        # measurements are deferred to the end of the program (sampled from the final state, as Aer does
        # for this circuit class); that is exact only while nothing touches a qubit after its measurement
        mq = getattr(prog, '_measured_q', None)
        if mq is None or len(mq[1]) != len(prog.measures):
            mq = prog._measured_q = (set(prog.measures.values()), dict(prog.measures))
        if not mq[0].isdisjoint(qubits):
            raise ValueError('qcmrf_b200: gate %s%r follows a measurement of one of its qubits; mid-circuit '
                             'measurement with later use of the qubit is not supported' % (name, qubits))
    if params and not all(np.isfinite(float(p)) for p in params):
        raise ValueError('qcmrf_b200: gate %s%r has a non-finite parameter (a QCMRF with theta > 0 has no '
                         'circuit angle: theta must be <= 0)' % (name, qubits))
    if canon == 'swap':
        a, b = qubits
        for c, t in ((a, b), (b, a), (a, b)):
            prog.gates.append(Gate('cx', (c, t), (), (1,)))
        return
    if len(qubits) == 1:
        one_qubit_matrix(canon, params)                        # validates the name
        prog.gates.append(Gate(canon, qubits, tuple(float(p) for p in params)))
        return
    if canon not in _CTRL_BASE:
        raise ValueError('qcmrf_b200: unsupported gate %r on %d qubits' % (name, len(qubits)))
    nc = len(qubits) - 1
    cv = tuple(int(v) for v in ctrl_values) if ctrl_values is not None else (1,) * nc
    if len(cv) != nc:
        raise ValueError('control-state length mismatch for %s' % name)
    prog.gates.append(Gate(canon, qubits, tuple(float(p) for p in params), cv))


def _ctrl_values_of(op, n_ctrl):
    cs = getattr(op, 'ctrl_values', None)
    if cs is not None:
        return tuple(cs)
    cs = getattr(op, 'ctrl_state', None)        # real Qiskit: int, bit j <-> control j
    if cs is None:
        return None
    return tuple((int(cs) >> j) & 1 for j in range(n_ctrl))


def _walk(prog, circ, qmap, cmap):
    for inst in circ.data:
        op = getattr(inst, 'operation', None)
        if op is None:                            # legacy (op, qargs, cargs) tuples
            op, qargs, cargs = inst
        else:
            qargs, cargs = inst.qubits, inst.clbits
        qs = [qmap[_qindex(circ, q)] for q in qargs]
        name = op.name.lower()
        if name == 'measure':
            cs = [cmap[_qindex_c(circ, c)] for c in cargs]
            for q, c in zip(qs, cs):
                prog.measures[c] = q
            continue
        if name in ('barrier', 'delay'):
            continue
        if name == 'reset' or getattr(op, 'condition', None) is not None:
            raise ValueError('qcmrf_b200: %s is not supported (measurements are deferred to the end of the '
                             'program)' % ('reset' if name == 'reset' else 'a classically conditioned gate'))
        canon = _ALIASES.get(name, name)
        primitive = canon in _ONEQ or canon in _PARAM1Q or canon in _CTRL_BASE or canon in ('swap', 'u')
        if primitive:
            _emit(prog, name, qs, list(getattr(op, 'params', ()) or ()),
                  _ctrl_values_of(op, len(qs) - 1))
            continue
        definition = getattr(op, 'definition', None)
        if definition is None:
            raise ValueError('qcmrf_b200: cannot lower instruction %r (no definition)' % op.name)
        prog.global_phase += float(getattr(definition, 'global_phase', 0.0) or 0.0)
        _walk(prog, definition, qs, [cmap[_qindex_c(circ, c)] for c in cargs])


def _qindex_c(circ, c):
    if isinstance(c, (int, np.integer)):
        return int(c)
    return circ.find_bit(c).index


def lower(circuit) -> Program:
    """Flatten a circuit object into a Program."""
    if isinstance(circuit, Program):
        return circuit
    fast = getattr(circuit, '_lower_program', None)
    if fast is not None:
        prog = fast()
        if prog is not None:
            return prog
    nq, nc = int(circuit.num_qubits), int(circuit.num_clbits)
    prog = Program(nq, nc, name=str(getattr(circuit, 'name', '') or ''))
    prog.global_phase = float(getattr(circuit, 'global_phase', 0.0) or 0.0)
    _walk(prog, circuit, list(range(nq)), list(range(nc)))
    nv = getattr(circuit, 'num_vertices', None)
    if nv is None:
        nv = (getattr(circuit, 'metadata', None) or {}).get('num_vertices')
    if nv is not None:
        prog.metadata['num_vertices'] = int(nv)
    return prog


def to_jsonable(prog: Program):
    return {'n_qubits': prog.n_qubits, 'n_clbits': prog.n_clbits,
            'gates': [[g.name, list(g.qubits), list(g.params), list(g.ctrl_values)] for g in prog.gates],
            'measures': {str(c): q for c, q in sorted(prog.measures.items())},
            'global_phase': prog.global_phase}


def from_jsonable(d) -> Program:
    p = Program(d['n_qubits'], d['n_clbits'], global_phase=d.get('global_phase', 0.0))
    p.gates = [Gate(n, tuple(q), tuple(pa), tuple(cv)) for n, q, pa, cv in d['gates']]
    p.measures = {int(c): int(q) for c, q in d['measures'].items()}
    return p


def to_oracle_ops(prog: Program):
    """The same program in the tuple format oracle.statevector executes (tests only)."""
    ops = []
    for g in prog.gates:
        if not g.controls:
            if g.params:
                ops.append((g.name, g.params[0], g.qubits[0]))
            else:
                ops.append((g.name, g.qubits[0]))
        elif g.name in ('cx', 'mcx'):
            ops.append(('mcx', g.controls, g.ctrl_values, g.target))
        elif g.name in ('cp', 'mcp'):
            ops.append(('mcp', g.params[0], g.controls, g.ctrl_values, g.target))
        elif g.name == 'cz':
            ops.append(('mcp', float(np.pi), g.controls, g.ctrl_values, g.target))
        else:
            raise ValueError('no oracle form for %s' % g.name)
    for c, q in sorted(prog.measures.items()):
        ops.append(('measure', q, c))
    if prog.global_phase:
        ops.append(('gphase', prog.global_phase))
    return ops
