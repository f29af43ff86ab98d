"""Exact MRF inference on the GPU: the classical ground truth of the reference's evaluation.

/root/reference/eval.py:84-93 obtains the exact distribution of a model from the proprietary
``kiopto_native`` ("px") module -- ``lnZ = px.infer(b, task='partition')`` and, for every state id,
``p[xid] = exp(px.logpot(b, xid) - lnZ)``.  For binary variables that is one enumeration of 2^n states
(SURVEY.md App. E.3 (iii)); ``ExactMRF`` runs it through the engine's C ABI (``qcm_mrf_exact``, one CUDA
kernel per pass: max, sum-exp, pmf), so the ground truth scales with the simulator (n ~ 26-30) instead of
stopping where a Python loop over 2^n states does.  Conventions as in the reference: weights
clique-major, ``itertools.product`` order inside a clique; state id with x_0 as its most significant bit.

There is no CPU fallback here: without the CUDA library / a GPU the calls raise (the compat shim keeps its own
numpy enumeration for the fixture-sized models of eval.py when no GPU is present).
"""
import numpy as np

from . import _native

__all__ = ['ExactMRF']


class ExactMRF:
    def __init__(self, cliques, weights=None, device=0):
        self.cliques = [list(int(v) for v in c) for c in cliques]
        self.n = max(max(c) for c in self.cliques) + 1
        self.dim = sum(1 << len(c) for c in self.cliques)
        self.weights = np.zeros(self.dim) if weights is None else np.array(weights, dtype=np.float64)
        if self.weights.size != self.dim:
            raise ValueError('weights has %d entries, the clique structure needs %d' % (self.weights.size, self.dim))
        self.device = device
        self._memo = None
        self.device_ms = None

    def _run(self, want_pmf):
        key = (self.weights.tobytes(), want_pmf)
        if self._memo is None or self._memo[0][0] != key[0] or (want_pmf and not self._memo[0][1]):
            lz, pmf, _e, ms = _native.mrf_exact(self.cliques, self.weights, self.n, want_pmf=want_pmf, device=self.device)
            self._memo = (key, lz, pmf)
            self.device_ms = ms
        return self._memo[1], self._memo[2]

    def log_partition(self):
        """ln Z = ln sum_x exp(sum_C w[C, x_C])   (px.infer(b, task='partition'))."""
        return self._run(False)[0]

    def pmf(self):
        """p[xid] = exp(energy(xid) - ln Z) for all 2^n states, x_0 = MSB of xid."""
        return self._run(True)[1]

    def logpot(self, xid):
        """Energy of one state (px.logpot(b, xid)): sum over the cliques, O(|cliques|), on the host."""
        xid = int(xid)
        e, off, n = 0.0, 0, self.n
        for c in self.cliques:
            y = 0
            for v in c:
                y = (y << 1) | ((xid >> (n - 1 - v)) & 1)
            e += float(self.weights[off + y])
            off += 1 << len(c)
        return e

    def success_probability(self):
        """delta = Z / 2^n of the QCMRF circuit with theta = these weights."""
        return float(np.exp(self.log_partition() - self.n * np.log(2.0)))
