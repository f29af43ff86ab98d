"""Stand-in for the proprietary ``kiopto_native`` ("px") exact-inference module the
reference uses as its classical ground truth (run_experiment.py:26-27,
eval.py:33-34,84-107).  Brute-force enumeration over binary variables -- enough for
every call the two scripts make.  Conventions (the only ones consistent with the
reference's own results): weights are clique-major, itertools.product
order inside a clique; state id xid has x_0 as its most significant bit.

This is host-side evaluation glue for eval.py's table, not part of the simulator path.

``infer`` is memoised per weight vector and ``logpot`` evaluates ONE state in O(|cliques|), so eval.py's loop
over all 2^n states costs O(2^n) instead of O(4^n).  With a CUDA device present ln Z comes from the GPU
enumeration (the engine's qcm_mrf_exact, see qcmrf_b200/exact.py); the numpy enumeration below only serves
GPU-less boxes (this stand-in for a third-party module is evaluation glue, not the simulator).
"""
import itertools

import numpy as np


_GPU = None


def _gpu_present():
    global _GPU
    if _GPU is None:
        try:
            from qcmrf_b200 import _native
            _GPU = _native.device_count() > 0
        except (RuntimeError, OSError):
            _GPU = False
    return _GPU


class _Model:
    def __init__(self, cliques, states, inference=None):
        self.cliques = [list(C) for C in cliques]
        self.states = np.asarray(states, dtype=np.int64)
        if np.any(self.states != 2):
            raise ValueError('kiopto_native shim: binary variables only')
        self.n = len(self.states)
        self.w = np.zeros(sum(2 ** len(C) for C in self.cliques))
        self.inference = inference
        self._memo = None                       # (weights bytes, ln Z)

    def logpot(self, xid):
        xid, n, e, off = int(xid), self.n, 0.0, 0
        for C in self.cliques:
            y = 0
            for v in C:
                y = (y << 1) | ((xid >> (n - 1 - v)) & 1)
            e += float(self.w[off + y])
            off += 1 << len(C)
        return e

    def log_partition(self):
        key = self.w.tobytes()
        if self._memo is None or self._memo[0] != key:
            lz = None
            if _gpu_present():
                from qcmrf_b200 import _native                 # a GPU is there: its errors are errors
                lz = _native.mrf_exact(self.cliques, self.w, self.n, want_pmf=False)[0]
            if lz is None:                                     # GPU-less box (the build container's tests)
                e = self.energies()
                m = e.max()
                lz = float(m + np.log(np.exp(e - m).sum()))
            self._memo = (key, lz)
        return self._memo[1]

    def energies(self):
        n = self.n
        xid = np.arange(1 << n, dtype=np.int64)
        e = np.zeros(1 << n)
        off = 0
        for C in self.cliques:
            m = len(C)
            y = np.zeros(1 << n, dtype=np.int64)
            for j, v in enumerate(C):
                y |= ((xid >> (n - 1 - v)) & 1) << (m - 1 - j)
            e += self.w[off + y]
            off += 1 << m
        return e


def backend(cliques, states, inference=None):
    return _Model(cliques, states, inference)


def weights(b):
    return b.w


def infer(b, task='partition'):
    if task == 'partition':
        return b.log_partition()
    raise ValueError('kiopto_native shim: unsupported task %r' % task)


def logpot(b, xid):
    return b.logpot(xid)


def sample(b, pam=False, num=10000, burn=10):
    """Exact sampling stands in for the Gibbs / perturb-and-MAP samplers of
    eval.py:95-113 (classical baselines of the paper, out of the simulator's scope)."""
    e = b.energies()
    p = np.exp(e - e.max())
    p /= p.sum()
    total = num if pam else num * burn + burn
    ids = np.random.choice(len(p), size=total, p=p)
    return [[(int(i) >> (b.n - 1 - v)) & 1 for v in range(b.n)] for i in ids]
