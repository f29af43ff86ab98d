"""Basis translation to the reference's target gate set.

The reference runs ``transpile(CIRCS, basis_gates=['cx','id','rz','sx','x'])`` before
handing circuits to Aer (/root/reference/run_experiment.py:52).  Qiskit's transpiler
is not available here; this is an independent, exact (up to a tracked global phase)
translation of the instructions QCMRF programs contain -- h, x, cp, flagged
multi-controlled X -- plus the usual single-qubit gates, into that basis:

  1q gate   -> rz . sx . rz . sx . rz   (ZSXZ Euler form; shorter when diagonal / X / SX)
  cp(l)     -> rz, cx, rz, cx, rz
  mcx/AND   -> X-conjugated open controls; cx (1 control), the 6-cx Toffoli (2), and
               H . mcp(pi) . H with the ancilla-free recursive multi-controlled phase
               (>= 3 controls)

The result is what the backend's fusion pass must collapse back into one sweep per
clique, so tests run every fixture model through this path as well.
"""
import cmath
import math
from typing import Iterable, List

import numpy as np

from . import ir
from .circuit import Instruction, QuantumCircuit

DEFAULT_BASIS = ('cx', 'id', 'rz', 'sx', 'x')
_PI = math.pi


class _Sink:
    """Collects the translated gates as flat lists (kind code, target, control, angle): see ir.BasisProgram."""

    def __init__(self, global_phase=0.0):
        self.k, self.q, self.c, self.p = [], [], [], []
        self.global_phase = global_phase

    def rz(self, lam, q):
        self.k.append(0); self.q.append(q); self.c.append(-1); self.p.append(lam)

    def sx(self, q):
        self.k.append(1); self.q.append(q); self.c.append(-1); self.p.append(0.0)

    def x(self, q):
        self.k.append(2); self.q.append(q); self.c.append(-1); self.p.append(0.0)

    def id(self, q):
        self.k.append(3); self.q.append(q); self.c.append(-1); self.p.append(0.0)

    def cx(self, a, b):
        self.k.append(4); self.q.append(b); self.c.append(a); self.p.append(0.0)


class BasisCircuit(QuantumCircuit):
    """What ``transpile`` returns: a QuantumCircuit in the cx/id/rz/sx/x basis whose instruction list is materialised
    on first access -- the engine's own lowering takes the flat gate arrays (ir.BasisProgram) instead of walking
    ~15 000 instruction objects; anything that touches ``.data`` (or edits the circuit) gets the ordinary list."""

    def __init__(self, nq, nc, name, sink, measures):
        super().__init__(nq, nc, name=name, global_phase=sink.global_phase)
        self.__dict__['_bc_sink'] = sink
        self.__dict__['_bc_measures'] = list(measures)
        self.__dict__['_bc_data'] = None

    @property
    def _qc_data(self):
        d = self.__dict__.get('_bc_data')
        if d is None:
            d = self.__dict__['_bc_data'] = []
            snk = self.__dict__.get('_bc_sink')
            if snk is not None:
                emit = (self.rz, self.sx, self.x, self.id)
                for k, q, c, p in zip(snk.k, snk.q, snk.c, snk.p):
                    if k == 0:
                        self.rz(p, q)
                    elif k == 4:
                        self.cx(c, q)
                    else:
                        emit[k](q)
                for q, c in self.__dict__['_bc_measures']:
                    self.measure(q, c)
        return d

    @_qc_data.setter
    def _qc_data(self, value):
        self.__dict__['_bc_data'] = value

    def _lower_program(self):
        """ir.BasisProgram over the gate arrays, or None once the instruction list exists (it may have been edited)."""
        if self.__dict__.get('_bc_data') is not None or self.__dict__.get('_bc_sink') is None:
            return None
        snk = self.__dict__['_bc_sink']
        prog = ir.BasisProgram(self.num_qubits, self.num_clbits, snk.k, snk.q, snk.c, snk.p, name=str(self.name),
                               global_phase=float(self.global_phase))
        for q, c in self.__dict__['_bc_measures']:
            prog.measures[c] = q
        if 'num_vertices' in self.metadata:
            prog.metadata['num_vertices'] = int(self.metadata['num_vertices'])
        return prog

    def copy(self, name=None):
        out = QuantumCircuit.copy(self, name)              # materialises: a copy is an ordinary circuit
        return out


class _Out:
    def __init__(self, sink):
        self.c = sink

    def rz(self, lam, q):
        lam = math.remainder(lam, 4 * _PI)
        if abs(lam) > 1e-15:
            self.c.rz(lam, q)

    def p(self, lam, q):                       # p(l) = e^{il/2} rz(l)
        self.rz(lam, q)
        self.c.global_phase += lam / 2

    def sx(self, q):
        self.c.sx(q)

    def x(self, q):
        self.c.x(q)

    def cx(self, a, b):
        self.c.cx(a, b)

    def h(self, q):                            # H = e^{i pi/4} rz(pi/2) sx rz(pi/2)
        self.rz(_PI / 2, q)
        self.sx(q)
        self.rz(_PI / 2, q)
        self.c.global_phase += _PI / 4

    def u1q(self, U, q):
        """Arbitrary 2x2 unitary in the ZSX basis."""
        U = np.asarray(U, dtype=np.complex128)
        if abs(U[0, 1]) < 1e-14 and abs(U[1, 0]) < 1e-14:      # diagonal
            a, b = np.angle(U[0, 0]), np.angle(U[1, 1])
            self.rz(b - a, q)
            self.c.global_phase += (a + b) / 2
            return
        # U = e^{ia} rz(phi) ry(theta) rz(lam);  ry(theta) = e^{..} via two sx:
        # rz(phi + pi) sx rz(theta + pi) sx rz(lam) = e^{-i pi/2}... -> fix the phase numerically
        theta = 2 * math.atan2(abs(U[1, 0]), abs(U[0, 0]))
        phi_plus_lam = np.angle(U[1, 1]) - np.angle(U[0, 0]) if abs(U[0, 0]) > 1e-14 else 0.0
        phi_minus_lam = np.angle(U[1, 0]) - np.angle(-U[0, 1]) if abs(U[1, 0]) > 1e-14 else 0.0
        if abs(U[0, 0]) <= 1e-14:
            phi_plus_lam = 0.0
            phi_minus_lam = np.angle(U[1, 0]) - np.angle(-U[0, 1])
        phi = 0.5 * (phi_plus_lam + phi_minus_lam)
        lam = 0.5 * (phi_plus_lam - phi_minus_lam)
        # the two angle differences are known modulo 2 pi, their halves modulo pi: (phi, lam) or (phi + pi, lam + pi) --
        # the second flips the sign of the off-diagonal entries against the diagonal, so exactly one of them is U
        for shift in (0.0, _PI):
            seq = [('rz', lam + shift), ('sx',), ('rz', theta + _PI), ('sx',), ('rz', phi + shift + _PI)]
            M = np.eye(2, dtype=np.complex128)
            for g in seq:
                G = ir.one_qubit_matrix(g[0], g[1:])
                M = G @ M
            k = np.argmax(np.abs(U))
            ph = np.angle(U.flat[k] / M.flat[k])
            if np.abs(M * np.exp(1j * ph) - U).max() <= 1e-9:
                break
        else:
            raise AssertionError('ZSX decomposition failed')
        for g in seq:
            if g[0] == 'rz':
                self.rz(g[1], q)
            else:
                self.sx(q)
        self.c.global_phase += ph

    def cp(self, lam, a, b):
        self.p(lam / 2, a)
        self.cx(a, b)
        self.p(-lam / 2, b)
        self.cx(a, b)
        self.p(lam / 2, b)

    def ccx(self, a, b, t):
        self.h(t)
        self.cx(b, t); self.p(-_PI / 4, t)
        self.cx(a, t); self.p(_PI / 4, t)
        self.cx(b, t); self.p(-_PI / 4, t)
        self.cx(a, t); self.p(_PI / 4, b); self.p(_PI / 4, t)
        self.h(t)
        self.cx(a, b); self.p(_PI / 4, a); self.p(-_PI / 4, b)
        self.cx(a, b)

    def mcx(self, ctrls: List[int], t):
        if len(ctrls) == 0:
            self.x(t)
        elif len(ctrls) == 1:
            self.cx(ctrls[0], t)
        elif len(ctrls) == 2:
            self.ccx(ctrls[0], ctrls[1], t)
        else:
            self.h(t)
            self.mcp(_PI, ctrls, t)
            self.h(t)

    def mcp(self, lam, ctrls: List[int], t):
        if len(ctrls) == 0:
            self.p(lam, t)
        elif len(ctrls) == 1:
            self.cp(lam, ctrls[0], t)
        else:
            last, rest = ctrls[-1], ctrls[:-1]
            self.cp(lam / 2, last, t)
            self.mcx(rest, last)
            self.cp(-lam / 2, last, t)
            self.mcx(rest, last)
            self.mcp(lam / 2, rest, t)


def _controlled_unitary(o, U, ctrls: List[int], t):
    """(Multi-)controlled arbitrary 1-qubit gate (cy, ch, crz, crx, cry, csx): U = e^{i alpha} Rz(beta) Ry(gamma) Rz(delta),
    A = Rz(beta) Ry(gamma/2), B = Ry(-gamma/2) Rz(-(delta+beta)/2), C = Rz((delta-beta)/2): A B C = 1 and A X B X C = U e^{-i
    alpha}, so C^n(U) = C(t), C^nX, B(t), C^nX, A(t) and the phase alpha on the controls alone."""
    U = np.asarray(U, dtype=np.complex128)
    alpha = 0.5 * cmath.phase(U[0, 0] * U[1, 1] - U[0, 1] * U[1, 0])
    V = U * cmath.exp(-1j * alpha)
    gamma = 2.0 * math.atan2(abs(V[1, 0]), abs(V[0, 0]))
    plus = 2.0 * cmath.phase(V[1, 1]) if abs(V[1, 1]) > 1e-12 else 0.0          # beta + delta
    minus = 2.0 * cmath.phase(V[1, 0]) if abs(V[1, 0]) > 1e-12 else 0.0         # beta - delta
    beta, delta = 0.5 * (plus + minus), 0.5 * (plus - minus)

    def rz(a):
        return np.array([[cmath.exp(-0.5j * a), 0], [0, cmath.exp(0.5j * a)]])

    def ry(a):
        c, s_ = math.cos(a / 2), math.sin(a / 2)
        return np.array([[c, -s_], [s_, c]], dtype=np.complex128)
    A, B, C = rz(beta) @ ry(gamma / 2), ry(-gamma / 2) @ rz(-(delta + beta) / 2), rz((delta - beta) / 2)
    o.u1q(C, t)
    o.mcx(ctrls, t)
    o.u1q(B, t)
    o.mcx(ctrls, t)
    o.u1q(A, t)
    if abs(math.remainder(alpha, 2 * _PI)) > 1e-15:
        o.mcp(alpha, ctrls[:-1], ctrls[-1])


def _translate(prog: ir.Program, name) -> QuantumCircuit:
    out = _Sink(prog.global_phase)
    o = _Out(out)
    for g in prog.gates:
        if not g.controls:
            if g.name == 'x':
                o.x(g.target)
            elif g.name == 'sx':
                o.sx(g.target)
            elif g.name == 'id':
                out.id(g.target)
            elif g.name == 'rz':
                o.rz(g.params[0], g.target)
            elif g.name == 'h':
                o.h(g.target)
            else:
                o.u1q(g.base_matrix(), g.target)
            continue
        opens = [c for c, v in zip(g.controls, g.ctrl_values) if v == 0]
        for c in opens:
            o.x(c)
        base = ir._CTRL_BASE[g.name]
        if base == 'x':
            o.mcx(list(g.controls), g.target)
        elif base in ('p', 'z'):
            o.mcp(g.params[0] if base == 'p' else _PI, list(g.controls), g.target)
        else:
            _controlled_unitary(o, g.base_matrix(), list(g.controls), g.target)
        for c in opens:
            o.x(c)
    out.global_phase = math.remainder(out.global_phase, 2 * _PI)
    measures = [(q, c) for c, q in sorted(prog.measures.items(), key=lambda cq: (cq[1], cq[0]))]
    return BasisCircuit(prog.n_qubits, prog.n_clbits, name, out, measures)


def transpile(circuits, backend=None, basis_gates=None, optimization_level=None, **_ignored):
    """``qiskit.transpile`` stand-in: basis translation only (no layout/routing: the
    simulator is all-to-all)."""
    basis = tuple(basis_gates) if basis_gates is not None else DEFAULT_BASIS
    missing = {'cx', 'rz', 'sx', 'x'} - set(basis)
    if missing:
        raise ValueError('transpile: basis must contain cx, rz, sx, x (missing %s)' % sorted(missing))
    single = not isinstance(circuits, (list, tuple))
    out = []
    for c in ([circuits] if single else circuits):
        prog = ir.lower(c)
        t = _translate(prog, getattr(c, 'name', 'circuit'))
        if 'num_vertices' in prog.metadata:
            t.metadata['num_vertices'] = prog.metadata['num_vertices']
        out.append(t)
    return out[0] if single else out
