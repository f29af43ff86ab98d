"""Oracle restatement of the gate program the reference's circuit constructor emits.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows /root/reference/QCMRF.py:

* register sizing                      QCMRF.py:52-65,78
* theta -> gamma                       QCMRF.py:144-157
* program order (H layer, per-clique
  H . CUC . X . CUC^-1 . X . H, measures)  QCMRF.py:199-243
* variable v lives on qubit n-1-v      QCMRF.py:219
* y enumerated in itertools.product
  order, first listed vertex slowest   QCMRF.py:221
* terms with gamma ~ 0 are skipped     QCMRF.py:223

A program is a flat list of tuples:
    ('h', q) ('x', q) ('mcx', ctrls, values, target) ('cp', lam, c, t)
    ('measure', q, c) ('barrier',)
``mcx`` flips ``target`` iff every ctrls[j] equals values[j]: that is what
``qiskit.circuit.library.AND(m, flags)`` does on its result qubit with
flags = 2*y-1 (flag>0 -> control must be 1, flag<0 -> must be 0).
"""
import itertools

import numpy as np


def sizes(cliques):
    """(n, k, N, dim): variables, cliques, total qubits n+k+1, parameter count."""
    n = max(v for C in cliques for v in C) + 1
    k = len(cliques)
    dim = sum(2 ** len(C) for C in cliques)
    return n, k, n + k + 1, dim


def theta_to_gamma(theta, beta=1.0):
    """gamma_i = 0.5*arccos(exp(beta*theta_i/2))  (QCMRF.py:154)."""
    theta = np.asarray(theta, dtype=np.float64)
    return 0.5 * np.arccos(np.exp(beta * 0.5 * theta))


def gamma_to_theta(gamma, beta=1.0):
    """theta_i = 2*ln(cos 2 gamma_i)/beta  (QCMRF.py:139)."""
    gamma = np.asarray(gamma, dtype=np.float64)
    return 2.0 * np.log(np.cos(2.0 * gamma)) / beta


def qcmrf_program(cliques, theta=None, gamma=None, beta=1.0,
                  with_measurements=True, with_barriers=False):
    """Flat primitive program for QCMRF(cliques, theta|gamma, beta)."""
    n, k, N, dim = sizes(cliques)
    if gamma is None:
        if theta is None:
            raise ValueError("oracle needs explicit theta or gamma")
        gamma = theta_to_gamma(theta, beta)
    gamma = np.asarray(gamma, dtype=np.float64)
    if len(gamma) != dim:
        raise ValueError("parameter vector has wrong length, expected %d" % dim)
    scratch = n                     # AND result qubit   (QCMRF.py:219)
    ops = [('h', q) for q in range(n)]
    if with_barriers:
        ops.append(('barrier',))
    i = 0
    for ii, C in enumerate(cliques):
        anc = n + 1 + ii            # QCMRF.py:231
        ctrls = tuple((n - 1) - v for v in C)
        cuc = []
        for y in itertools.product([0, 1], repeat=len(C)):
            if not np.isclose(gamma[i], 0):
                cuc.append(('mcx', ctrls, tuple(y), scratch))
                cuc.append(('cp', 2.0 * float(gamma[i]), scratch, anc))
                cuc.append(('mcx', ctrls, tuple(y), scratch))
            i += 1
        cuc_inv = [(g[0], -g[1], g[2], g[3]) if g[0] == 'cp' else g
                   for g in reversed(cuc)]
        ops.append(('h', anc))
        ops.extend(cuc)
        ops.append(('x', anc))
        ops.extend(cuc_inv)
        ops.append(('x', anc))
        ops.append(('h', anc))
        if with_measurements:
            ops.append(('measure', anc, anc))
        if with_barriers:
            ops.append(('barrier',))
    if with_measurements:
        for q in range(n):
            ops.append(('measure', q, q))
    return ops, N


def rx_tables(cliques, theta=None, gamma=None, beta=1.0):
    """Closed form of one clique block: a uniformly-controlled RX(4*gamma) on the
    clique's ancilla.  Returns per clique (ctrl_qubits,
    cos 2g[2^m], sin 2g[2^m]) with the table index bit j = value of ctrl_qubits[j],
    ctrl_qubits listed lowest table bit first."""
    n, k, N, dim = sizes(cliques)
    if gamma is None:
        gamma = theta_to_gamma(theta, beta)
    gamma = np.asarray(gamma, dtype=np.float64)
    out, off = [], 0
    for C in cliques:
        m = len(C)
        ctrl = [(n - 1) - v for v in C]          # y[j] <-> qubit ctrl[j]
        c = np.empty(2 ** m)
        s = np.empty(2 ** m)
        for yi, y in enumerate(itertools.product([0, 1], repeat=m)):
            t = sum(y[j] << j for j in range(m))  # table bit j = y[j]
            c[t] = np.cos(2 * gamma[off + yi])
            s[t] = np.sin(2 * gamma[off + yi])
        out.append((ctrl, c, s))
        off += 2 ** m
    return out
