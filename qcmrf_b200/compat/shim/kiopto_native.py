"""Stand-in for the proprietary ``kiopto_native`` ("px") exact-inference module the
reference uses as its classical ground truth (run_experiment.py:26-27,
eval.py:33-34,84-107).  Brute-force enumeration over binary variables -- enough for
every call the two scripts make.  Conventions (the only ones consistent with the
reference's own results): weights are clique-major, itertools.product
order inside a clique; state id xid has x_0 as its most significant bit.

This is host-side evaluation glue for eval.py's table, not part of the simulator path.
"""
import itertools

import numpy as np


class _Model:
    def __init__(self, cliques, states, inference=None):
        self.cliques = [list(C) for C in cliques]
        self.states = np.asarray(states, dtype=np.int64)
        if np.any(self.states != 2):
            raise ValueError('kiopto_native shim: binary variables only')
        self.n = len(self.states)
        self.w = np.zeros(sum(2 ** len(C) for C in self.cliques))
        self.inference = inference

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
    e = b.energies()
    if task == 'partition':
        m = e.max()
        return float(m + np.log(np.exp(e - m).sum()))
    raise ValueError('kiopto_native shim: unsupported task %r' % task)


def logpot(b, xid):
    return float(b.energies()[int(xid)])


def sample(b, pam=False, num=10000, burn=10):
    """Exact sampling stands in for the Gibbs / perturb-and-MAP samplers of
    eval.py:95-113 (classical baselines of the paper, out of the simulator's scope)."""
    e = b.energies()
    p = np.exp(e - e.max())
    p /= p.sum()
    total = num if pam else num * burn + burn
    ids = np.random.choice(len(p), size=total, p=p)
    return [[(int(i) >> (b.n - 1 - v)) & 1 for v in range(b.n)] for i in ids]
