"""ctypes bridge to the oracle's C executor (oracle/csrc/qcm_oracle.c).

TEST INFRASTRUCTURE (see oracle/__init__.py): used by tests as a second,
independent executor and by bench.py as the timed CPU baseline ("port").
"""
import ctypes
import os
import subprocess

import numpy as np

from .statevector import gate_matrix

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libqcm_oracle.so')

ORC_U1, ORC_PHASE, ORC_MUX = 0, 1, 2


class OrcOp(ctypes.Structure):
    _fields_ = [('kind', ctypes.c_int32), ('target', ctypes.c_int32),
                ('cmask', ctypes.c_uint64), ('cval', ctypes.c_uint64),
                ('m', ctypes.c_double * 8),
                ('nctrl', ctypes.c_int32), ('ctrls', ctypes.c_int32 * 8),
                ('tab_off', ctypes.c_int64)]


def build(force=False):
    if force or not os.path.exists(_SO):
        subprocess.check_call(['make', '-C', _HERE], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        assert _lib.orc_sizeof_op() == ctypes.sizeof(OrcOp)
        _lib.orc_run.argtypes = [ctypes.c_int, ctypes.POINTER(OrcOp), ctypes.c_int,
                                 ctypes.c_void_p, ctypes.c_void_p]
        _lib.orc_init_zero_state.argtypes = [ctypes.c_int, ctypes.c_void_p]
        _lib.orc_postselect.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64,
                                        ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p,
                                        ctypes.POINTER(ctypes.c_double)]
        _lib.orc_sample.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64,
                                    ctypes.c_uint64, ctypes.c_void_p]
    return _lib


def _u1(mat, target, cmask=0, cval=0):
    op = OrcOp()
    op.kind, op.target, op.cmask, op.cval = ORC_U1, target, cmask, cval
    flat = np.asarray(mat, dtype=np.complex128).reshape(4)
    for i in range(4):
        op.m[2 * i], op.m[2 * i + 1] = flat[i].real, flat[i].imag
    return op


def _phase(lam, cmask):
    op = OrcOp()
    op.kind, op.cmask, op.cval = ORC_PHASE, cmask, cmask
    op.m[0] = lam
    return op


def compile_unfused(ops):
    """Primitive program (oracle.program tuples) -> (OrcOp array, measure map):
    one sweep per logical gate, the 'unfused, Aer-like' B1 baseline."""
    out, meas = [], {}
    for g in ops:
        name = g[0]
        if name == 'barrier':
            continue
        if name == 'measure':
            meas[g[2]] = g[1]
        elif name in ('h', 'x', 'y', 'z', 's', 'sdg', 't', 'tdg', 'sx', 'sxdg', 'id'):
            out.append(_u1(gate_matrix(name), g[1]))
        elif name in ('rz', 'rx', 'ry', 'p'):
            out.append(_u1(gate_matrix(name, g[1]), g[2]))
        elif name == 'cx':
            out.append(_u1(gate_matrix('x'), g[2], 1 << g[1], 1 << g[1]))
        elif name == 'cp':
            out.append(_phase(g[1], (1 << g[2]) | (1 << g[3])))
        elif name == 'mcx':
            cm = sum(1 << c for c in g[1])
            cv = sum(v << c for c, v in zip(g[1], g[2]))
            out.append(_u1(gate_matrix('x'), g[3], cm, cv))
        else:
            raise ValueError('oracle C executor: unsupported op %r' % (g,))
    arr = (OrcOp * max(len(out), 1))(*out)
    return arr, len(out), meas


def compile_fused(cliques, theta=None, gamma=None, beta=1.0):
    """B2 baseline: H layer + one uniformly-controlled RX(4 gamma) sweep per clique."""
    from .program import rx_tables, sizes
    n, k, N, dim = sizes(cliques)
    out, tabs = [_u1(gate_matrix('h'), q) for q in range(n)], []
    off = 0
    for ii, (ctrl, c, s) in enumerate(rx_tables(cliques, theta=theta, gamma=gamma, beta=beta)):
        op = OrcOp()
        op.kind, op.target, op.nctrl, op.tab_off = ORC_MUX, n + 1 + ii, len(ctrl), off
        for j, q in enumerate(ctrl):
            op.ctrls[j] = q
        t = np.zeros((len(c), 8))
        t[:, 0] = c; t[:, 3] = -s; t[:, 5] = -s; t[:, 6] = c   # [[c, -is], [-is, c]]
        tabs.append(t.reshape(-1))
        off += t.size
        out.append(op)
    arr = (OrcOp * len(out))(*out)
    return arr, len(out), np.concatenate(tabs), N


def run(N, arr, n_ops, tables=None, psi=None):
    if psi is None:
        psi = np.empty(1 << N, dtype=np.complex128)
        lib().orc_init_zero_state(N, psi.ctypes.data)
    tp = tables.ctypes.data if tables is not None else None
    rc = lib().orc_run(N, arr, n_ops, tp, psi.ctypes.data)
    if rc:
        raise RuntimeError('orc_run failed %d' % rc)
    return psi


def sample(N, psi, shots, seed):
    out = np.empty(shots, dtype=np.uint64)
    lib().orc_sample(N, psi.ctypes.data, shots, seed, out.ctypes.data)
    return out


def postselect(N, psi, mask, value, n_out):
    probs = np.empty(1 << n_out)
    kept = ctypes.c_double()
    lib().orc_postselect(N, psi.ctypes.data, mask, value, n_out, probs.ctypes.data,
                         ctypes.byref(kept))
    return probs, kept.value
