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
import math
from typing import Iterable, List

import numpy as np

from . import ir
from .circuit import Instruction, QuantumCircuit

DEFAULT_BASIS = ('cx', 'id', 'rz', 'sx', 'x')
_PI = math.pi


class _Out:
    def __init__(self, circ):
        self.c = circ

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
        seq = [('rz', lam), ('sx',), ('rz', theta + _PI), ('sx',), ('rz', phi + _PI)]
        M = np.eye(2, dtype=np.complex128)
        for g in seq:
            G = ir.one_qubit_matrix(g[0], g[1:])
            M = G @ M
        k = np.argmax(np.abs(U))
        ph = np.angle(U.flat[k] / M.flat[k])
        if np.abs(M * np.exp(1j * ph) - U).max() > 1e-9:
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


def _translate(prog: ir.Program, name) -> QuantumCircuit:
    out = QuantumCircuit(prog.n_qubits, prog.n_clbits, name=name, global_phase=prog.global_phase)
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
            raise ValueError('transpile: controlled-%s is not supported' % base)
        for c in opens:
            o.x(c)
    for c, q in sorted(prog.measures.items(), key=lambda cq: (cq[1], cq[0])):
        out.measure(q, c)
    out.global_phase = math.remainder(out.global_phase, 2 * _PI)
    return out


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
