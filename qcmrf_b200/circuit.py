"""Minimal circuit-construction layer with the Qiskit surface the reference touches.

Qiskit is not installed in this image (and cannot be), so the reference's scripts
need *something* that answers ``from qiskit import QuantumCircuit, transpile, Aer``.
This module supplies exactly the API surface /root/reference uses (SURVEY.md
App. D): ``QuantumCircuit(nq[, nc], name=)`` with ``h x cp append inverse measure
barrier num_qubits`` (QCMRF.py:78,205,207,218,225-227,231-236,239,243), plus the
basis gates ``transpile`` emits.  It is a recorder, not a simulator: all arithmetic
happens in the CUDA engine after ``ir.lower``.

Attribute names are prefixed ``_qc_`` on purpose: the reference's ``QCMRF`` stores
``_name _theta _gamma _beta _cliques _num_cliques _n _dim _c_max
_with_measurements _with_barriers basis_gates`` on ``self`` *before* calling
``super().__init__`` (QCMRF.py:36-43,50-65,78) and this class must not clobber them.
"""
from typing import Iterable, List, Sequence

import numpy as np

__all__ = ['Instruction', 'Gate', 'CircuitInstruction', 'QuantumCircuit', 'AND']

_SELF_INVERSE = {'h', 'x', 'y', 'z', 'id', 'cx', 'cy', 'cz', 'ch', 'swap', 'ccx', 'mcx', 'barrier'}
_INVERSE_NAME = {'s': 'sdg', 'sdg': 's', 't': 'tdg', 'tdg': 't', 'sx': 'sxdg', 'sxdg': 'sx'}
_NEGATE_PARAMS = {'rz', 'rx', 'ry', 'p', 'cp', 'crz', 'crx', 'cry', 'mcp', 'u1', 'cu1'}


class Instruction:
    """A named operation; composite ones carry a ``definition`` circuit."""

    def __init__(self, name, num_qubits, num_clbits=0, params=(), definition=None, ctrl_values=None):
        self.name = name
        self.num_qubits = int(num_qubits)
        self.num_clbits = int(num_clbits)
        self.params = list(params)
        self.definition = definition
        self.ctrl_values = None if ctrl_values is None else tuple(int(v) for v in ctrl_values)

    def inverse(self):
        n = self.name
        if self.definition is not None:
            return Instruction(n + '_dg', self.num_qubits, self.num_clbits, [], self.definition.inverse(),
                               self.ctrl_values)
        if n == 'measure':
            raise ValueError('measure is not invertible')
        if n in _SELF_INVERSE:
            return Instruction(n, self.num_qubits, 0, list(self.params), None, self.ctrl_values)
        if n in _INVERSE_NAME:
            return Instruction(_INVERSE_NAME[n], self.num_qubits, 0, [], None, self.ctrl_values)
        if n in _NEGATE_PARAMS:
            return Instruction(n, self.num_qubits, 0, [-p for p in self.params], None, self.ctrl_values)
        if n == 'u':
            th, ph, lm = self.params
            return Instruction('u', 1, 0, [-th, -lm, -ph])
        raise ValueError('cannot invert instruction %r' % n)

    def __repr__(self):
        return 'Instruction(%s, %d qubits, params=%r)' % (self.name, self.num_qubits, self.params)


Gate = Instruction


class CircuitInstruction:
    __slots__ = ('operation', 'qubits', 'clbits')

    def __init__(self, operation, qubits, clbits=()):
        self.operation = operation
        self.qubits = tuple(qubits)
        self.clbits = tuple(clbits)

    def __iter__(self):                       # legacy (op, qargs, cargs) unpacking
        return iter((self.operation, list(self.qubits), list(self.clbits)))


def _as_list(x) -> List[int]:
    if isinstance(x, (int, np.integer)):
        return [int(x)]
    if isinstance(x, (range, list, tuple, np.ndarray)):
        return [int(v) for v in x]
    if isinstance(x, Iterable):
        return [int(v) for v in x]
    raise TypeError('bad qubit specifier %r' % (x,))


class QuantumCircuit:
    def __init__(self, *regs, name=None, global_phase=0.0, metadata=None):
        sizes = [int(r) for r in regs]
        if len(sizes) == 0:
            sizes = [0, 0]
        elif len(sizes) == 1:
            sizes = [sizes[0], 0]
        elif len(sizes) > 2:
            raise TypeError('QuantumCircuit(num_qubits[, num_clbits]) expected')
        if sizes[0] < 0 or sizes[1] < 0:
            raise ValueError('register sizes must be non-negative')
        self._qc_nq, self._qc_nc = sizes
        self._qc_data: List[CircuitInstruction] = []
        self.name = name if name is not None else 'circuit'
        self.global_phase = float(global_phase)
        self.metadata = dict(metadata or {})

    # -- introspection ---------------------------------------------------------------
    @property
    def num_qubits(self):
        return self._qc_nq

    @property
    def num_clbits(self):
        return self._qc_nc

    @property
    def data(self):
        return self._qc_data

    @property
    def qubits(self):
        return list(range(self._qc_nq))

    @property
    def clbits(self):
        return list(range(self._qc_nc))

    def size(self):
        return sum(1 for i in self._qc_data if i.operation.name != 'barrier')

    def __len__(self):
        return len(self._qc_data)

    def count_ops(self):
        out = {}
        for inst in self._qc_data:
            out[inst.operation.name] = out.get(inst.operation.name, 0) + 1
        return out

    # -- building ----------------------------------------------------------------------
    def _qc_check(self, qs, cs=()):
        for q in qs:
            if not 0 <= q < self._qc_nq:
                raise IndexError('qubit index %d out of range for a %d-qubit circuit' % (q, self._qc_nq))
        for c in cs:
            if not 0 <= c < self._qc_nc:
                raise IndexError('clbit index %d out of range for %d clbits' % (c, self._qc_nc))
        if len(set(qs)) != len(qs):
            raise ValueError('duplicate qubit arguments %r' % (qs,))

    def _qc_add(self, op, qs, cs=()):
        self._qc_check(qs, cs)
        self._qc_data.append(CircuitInstruction(op, qs, cs))
        return self

    def _qc_1q(self, name, qubit, params=()):
        for q in _as_list(qubit):
            self._qc_add(Instruction(name, 1, 0, params), [q])
        return self

    def append(self, instruction, qargs=None, cargs=None):
        if isinstance(instruction, QuantumCircuit):
            instruction = instruction.to_instruction()
        qs = _as_list(qargs if qargs is not None else [])
        cs = _as_list(cargs if cargs is not None else [])
        if len(qs) != instruction.num_qubits:
            raise ValueError('instruction %s expects %d qubits, got %d' % (instruction.name, instruction.num_qubits, len(qs)))
        if len(cs) != instruction.num_clbits:
            raise ValueError('instruction %s expects %d clbits, got %d' % (instruction.name, instruction.num_clbits, len(cs)))
        return self._qc_add(instruction, qs, cs)

    def to_instruction(self, label=None):
        return Instruction(label or self.name, self._qc_nq, self._qc_nc, [], self)

    to_gate = to_instruction

    def compose(self, other, qubits=None, clbits=None, inplace=False):
        dst = self if inplace else self.copy()
        qmap = _as_list(qubits) if qubits is not None else list(range(other.num_qubits))
        cmap = _as_list(clbits) if clbits is not None else list(range(other.num_clbits))
        for inst in other.data:
            dst._qc_add(inst.operation, [qmap[q] for q in inst.qubits], [cmap[c] for c in inst.clbits])
        dst.global_phase += other.global_phase
        return None if inplace else dst

    def copy(self, name=None):
        out = QuantumCircuit(self._qc_nq, self._qc_nc, name=name or self.name, global_phase=self.global_phase,
                             metadata=self.metadata)
        out._qc_data = list(self._qc_data)
        return out

    def inverse(self):
        out = QuantumCircuit(self._qc_nq, self._qc_nc, name=str(self.name) + '_dg', global_phase=-self.global_phase)
        for inst in reversed(self._qc_data):
            out._qc_add(inst.operation.inverse(), inst.qubits, inst.clbits)
        return out

    # -- gates -------------------------------------------------------------------------
    def h(self, q): return self._qc_1q('h', q)
    def x(self, q): return self._qc_1q('x', q)
    def y(self, q): return self._qc_1q('y', q)
    def z(self, q): return self._qc_1q('z', q)
    def s(self, q): return self._qc_1q('s', q)
    def sdg(self, q): return self._qc_1q('sdg', q)
    def t(self, q): return self._qc_1q('t', q)
    def tdg(self, q): return self._qc_1q('tdg', q)
    def sx(self, q): return self._qc_1q('sx', q)
    def sxdg(self, q): return self._qc_1q('sxdg', q)
    def id(self, q): return self._qc_1q('id', q)
    i = id
    def rz(self, phi, q): return self._qc_1q('rz', q, [float(phi)])
    def rx(self, theta, q): return self._qc_1q('rx', q, [float(theta)])
    def ry(self, theta, q): return self._qc_1q('ry', q, [float(theta)])
    def p(self, lam, q): return self._qc_1q('p', q, [float(lam)])
    def u(self, theta, phi, lam, q): return self._qc_1q('u', q, [float(theta), float(phi), float(lam)])

    def cx(self, c, t): return self._qc_add(Instruction('cx', 2), [int(c), int(t)])
    cnot = cx
    def cz(self, c, t): return self._qc_add(Instruction('cz', 2), [int(c), int(t)])
    def cp(self, lam, c, t): return self._qc_add(Instruction('cp', 2, 0, [float(lam)]), [int(c), int(t)])
    def crz(self, lam, c, t): return self._qc_add(Instruction('crz', 2, 0, [float(lam)]), [int(c), int(t)])
    def swap(self, a, b): return self._qc_add(Instruction('swap', 2), [int(a), int(b)])
    def ccx(self, c1, c2, t): return self._qc_add(Instruction('mcx', 3, ctrl_values=(1, 1)), [int(c1), int(c2), int(t)])
    toffoli = ccx

    def mcx(self, control_qubits, target_qubit, ctrl_state=None, **_):
        cs = _as_list(control_qubits)
        if ctrl_state is None:
            vals = (1,) * len(cs)
        elif isinstance(ctrl_state, str):
            vals = tuple(int(ch) for ch in reversed(ctrl_state))
        else:
            vals = tuple((int(ctrl_state) >> j) & 1 for j in range(len(cs)))
        return self._qc_add(Instruction('mcx', len(cs) + 1, ctrl_values=vals), cs + [int(target_qubit)])

    def mcp(self, lam, control_qubits, target_qubit):
        cs = _as_list(control_qubits)
        return self._qc_add(Instruction('mcp', len(cs) + 1, 0, [float(lam)], ctrl_values=(1,) * len(cs)),
                            cs + [int(target_qubit)])

    def barrier(self, *qargs):
        qs = [q for a in qargs for q in _as_list(a)] if qargs else list(range(self._qc_nq))
        return self._qc_add(Instruction('barrier', len(qs)), qs)

    def measure(self, qubit, cbit):
        qs, cs = _as_list(qubit), _as_list(cbit)
        if len(qs) != len(cs):
            raise ValueError('measure: %d qubits but %d clbits' % (len(qs), len(cs)))
        for q, c in zip(qs, cs):
            self._qc_add(Instruction('measure', 1, 1), [q], [c])
        return self

    def measure_all(self):
        if self._qc_nc < self._qc_nq:
            self._qc_nc = self._qc_nq
        return self.measure(range(self._qc_nq), range(self._qc_nq))


class AND(QuantumCircuit):
    """``qiskit.circuit.library.AND(num_variable_qubits, flags)`` (call sites
    QCMRF.py:225,227): qubit ``num_variable_qubits`` is flipped iff every variable j
    with flags[j] > 0 is 1 and every variable with flags[j] < 0 is 0; flags[j] == 0
    leaves variable j out."""

    def __init__(self, num_variable_qubits, flags=None, mcx_mode='noancilla'):
        n = int(num_variable_qubits)
        super().__init__(n + 1, name='and')
        flags = list(flags) if flags is not None else [1] * n
        if len(flags) != n:
            raise ValueError('flags must have one entry per variable qubit')
        ctrls = [j for j in range(n) if flags[j] != 0]
        vals = tuple(1 if flags[j] > 0 else 0 for j in ctrls)
        if ctrls:
            self._qc_add(Instruction('mcx', len(ctrls) + 1, ctrl_values=vals), ctrls + [n])
        else:
            self.x(n)
