"""Oracle: brute-force MRF enumeration and the reference's post-selection metrics.

TEST INFRASTRUCTURE (see oracle/__init__.py).

* ``brute_force_pmf`` stands in for the proprietary ``kiopto_native`` calls in
  /root/reference/eval.py:84-93 (``px.logpot(b, xid) - px.infer(b,'partition')``).
  State id convention: x_0 is the MSB of xid (eval.py:100-101, QCMRF.py:264);
  theta index: clique-major, then ``itertools.product`` order inside the clique
  (QCMRF.py:215-228).
* ``extract_probs`` restates QCMRF.py:263-284; ``postselect_counts`` restates
  the inline twin eval.py:115-123; ``fidelity`` / ``kl`` restate
  QCMRF.py:247-261.
* ``regenerate_thetas`` restates run_experiment.py:3,23-33 (d = sum 2^|C| takes
  the place of ``len(px.weights(b))``).
"""
import itertools

import numpy as np

GRAPHS = [[[0]], [[0, 1]], [[0, 1], [1, 2], [2, 3]], [[0, 1], [1, 2], [2, 3], [3, 4]],
          [[0, 1, 2]], [[0, 1, 2], [2, 3, 4]], [[0, 1, 2, 3]]]   # run_experiment.py:20


def brute_force_pmf(cliques, theta, beta=1.0):
    """(p[2^n], delta, lnZ): p(x) = exp(beta*sum_C theta_{C,x_C}) / Z, delta = Z/2^n."""
    n = max(v for C in cliques for v in C) + 1
    theta = np.asarray(theta, dtype=np.float64)
    xid = np.arange(1 << n, dtype=np.int64)
    bits = [(xid >> (n - 1 - v)) & 1 for v in range(n)]      # x_v, x_0 = MSB
    e = np.zeros(1 << n)
    off = 0
    for C in cliques:
        m = len(C)
        y = np.zeros(1 << n, dtype=np.int64)
        for j, v in enumerate(C):
            y |= bits[v] << (m - 1 - j)                      # product order: y[0] slowest
        e += theta[off + y]
        off += 1 << m
    w = np.exp(beta * e)
    Z = float(w.sum())
    return w / Z, Z / (1 << n), float(np.log(Z))


def regenerate_thetas(scale, graphs=GRAPHS, reps=10, seed=1984):
    """The reference's model draw (run_experiment.py:3,23-33)."""
    from scipy.stats import halfnorm
    np.random.seed(seed)
    out = {}
    for j, C in enumerate(graphs):
        d = sum(2 ** len(c) for c in C)
        out[j] = [(-halfnorm.rvs(loc=0, scale=scale, size=d)).tolist() for _ in range(reps)]
    return out


def extract_probs(R, n, a):
    """QCMRF.py:263-284: keep keys '0'*a + x_0..x_{n-1}; returns (P/z, z/z0) or (P, 0)."""
    P = np.zeros(2 ** n)
    for i, y in enumerate(itertools.product([0, 1], repeat=n)):
        s0 = '0' * a + ''.join(str(b) for b in y)
        if s0 in R:
            P[i] += R[s0]
    z = P.sum()
    z0 = sum(R.values())
    if z == 0:
        return P, 0
    return P / z, z / z0


def postselect_counts(Q, n):
    """eval.py:115-123: q[kid] = Q[k] for kid=int(k,2) < 2^n; returns (q/Z, Z)."""
    q = np.zeros(1 << n)
    Z = 0
    for k, val in Q.items():
        kid = int(k, 2)
        if kid < (1 << n):
            q[kid] = val
            Z += val
    return q / Z, Z


def fidelity(P, Q):
    """QCMRF.py:247-253."""
    P = np.asarray(P, dtype=np.float64)
    Q = np.asarray(Q, dtype=np.float64)
    m = (P > 0) & (Q > 0)
    return float(np.sum(np.sqrt(P[m] * Q[m])) ** 2)


def kl(P, Q):
    """QCMRF.py:255-261."""
    P = np.asarray(P, dtype=np.float64)
    Q = np.asarray(Q, dtype=np.float64)
    m = (P > 0) & (Q > 0)
    return float(np.sum(P[m] * np.log(P[m] / Q[m])))


def closed_form_state(cliques, theta=None, gamma=None, beta=1.0):
    """Product-form final state: amplitude of (x, a) is
    2^{-n/2} prod_C (a_C=0 ? cos 2g : -i sin 2g).  Used to cross-check the
    gate-by-gate executor at sizes where it is slow."""
    from .program import rx_tables, sizes
    n, k, N, dim = sizes(cliques)
    tabs = rx_tables(cliques, theta=theta, gamma=gamma, beta=beta)
    idx = np.arange(1 << N, dtype=np.int64)
    psi = np.full(1 << N, 2.0 ** (-n / 2.0), dtype=np.complex128)
    psi[((idx >> n) & 1) == 1] = 0.0
    for ii, (ctrl, c, s) in enumerate(tabs):
        t = np.zeros(1 << N, dtype=np.int64)
        for j, q in enumerate(ctrl):
            t |= ((idx >> q) & 1) << j
        a = (idx >> (n + 1 + ii)) & 1
        psi *= np.where(a == 0, c[t], -1j * s[t])
    return psi
