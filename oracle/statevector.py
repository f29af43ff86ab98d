"""Oracle: gate-by-gate complex128 statevector execution + shot sampling (numpy).

TEST INFRASTRUCTURE (see oracle/__init__.py).  This restates what the
reference obtains from ``Aer.get_backend('qasm_simulator').run(T, shots=...)``
(/root/reference/run_experiment.py:54-57): a dense little-endian statevector
(index bit q <-> qubit q), one sweep per gate, measurements deferred to one
multinomial draw from |psi|^2 (legal because no gate follows a measured qubit,
QCMRF.py:238-243), keys printed with clbit N-1 leftmost.

Accepted ops (tuples):
    ('h',q) ('x',q) ('y',q) ('z',q) ('s',q) ('sdg',q) ('t',q) ('tdg',q)
    ('sx',q) ('sxdg',q) ('id',q)
    ('rz',lam,q) ('rx',lam,q) ('ry',lam,q) ('p',lam,q)
    ('cx',c,t) ('cz',c,t) ('cp',lam,c,t) ('swap',a,b)
    ('mcx',ctrls,values,target) ('mcp',lam,ctrls,values,target)
    ('measure',q,c) ('barrier',) ('gphase',lam)
"""
import numpy as np

_SQ2 = 1.0 / np.sqrt(2.0)
_FIXED = {
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
    'id': np.eye(2, dtype=np.complex128),
}


def gate_matrix(name, lam=None):
    if name in _FIXED:
        return _FIXED[name]
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
    raise ValueError("unknown gate " + name)


def _apply_1q(psi, N, q, U, sel=None):
    """psi viewed as [hi, 2, lo]; apply U on axis 1 (optionally only where the
    boolean index mask ``sel`` over the full index space is true)."""
    v = psi.reshape(1 << (N - q - 1), 2, 1 << q)
    a0 = v[:, 0, :].copy()
    a1 = v[:, 1, :].copy()
    n0 = U[0, 0] * a0 + U[0, 1] * a1
    n1 = U[1, 0] * a0 + U[1, 1] * a1
    if sel is None:
        v[:, 0, :] = n0
        v[:, 1, :] = n1
    else:
        m = sel.reshape(1 << (N - q - 1), 2, 1 << q)[:, 0, :]
        v[:, 0, :] = np.where(m, n0, a0)
        v[:, 1, :] = np.where(m, n1, a1)


def _ctrl_mask(N, ctrls, values, target):
    idx = np.arange(1 << N, dtype=np.uint64)
    ok = np.ones(1 << N, dtype=bool)
    for c, v in zip(ctrls, values):
        ok &= ((idx >> np.uint64(c)) & np.uint64(1)) == np.uint64(v)
    return ok


def run_program(ops, N):
    """Execute the program from |0...0>; returns (psi complex128[2^N], measure_map
    {clbit: qubit})."""
    psi = np.zeros(1 << N, dtype=np.complex128)
    psi[0] = 1.0
    meas = {}
    idx = None
    for g in ops:
        name = g[0]
        if name == 'barrier':
            continue
        if name == 'measure':
            meas[g[2]] = g[1]
            continue
        if name == 'gphase':
            psi *= np.exp(1j * g[1])
            continue
        if name in _FIXED:
            _apply_1q(psi, N, g[1], _FIXED[name])
        elif name in ('rz', 'rx', 'ry', 'p'):
            _apply_1q(psi, N, g[2], gate_matrix(name, g[1]))
        elif name == 'cx':
            _apply_1q(psi, N, g[2], _FIXED['x'], _ctrl_mask(N, (g[1],), (1,), g[2]))
        elif name == 'cz':
            _apply_1q(psi, N, g[2], _FIXED['z'], _ctrl_mask(N, (g[1],), (1,), g[2]))
        elif name == 'cp':
            _apply_1q(psi, N, g[3], gate_matrix('p', g[1]),
                      _ctrl_mask(N, (g[2],), (1,), g[3]))
        elif name == 'swap':
            a, b = g[1], g[2]
            for c, t in ((a, b), (b, a), (a, b)):
                _apply_1q(psi, N, t, _FIXED['x'], _ctrl_mask(N, (c,), (1,), t))
        elif name == 'mcx':
            _apply_1q(psi, N, g[3], _FIXED['x'], _ctrl_mask(N, g[1], g[2], g[3]))
        elif name == 'mcp':
            _apply_1q(psi, N, g[4], gate_matrix('p', g[1]),
                      _ctrl_mask(N, g[2], g[3], g[4]))
        else:
            raise ValueError("unknown op %r" % (g,))
    return psi, meas


def key_probabilities(psi, N, meas, n_clbits=None):
    """Exact distribution over classical-register integers (clbit c = bit c).
    Unmeasured clbits read 0 (QCMRF.py:238-243 never writes clbit n)."""
    if n_clbits is None:
        n_clbits = N
    p = np.abs(psi) ** 2
    idx = np.arange(1 << N, dtype=np.uint64)
    key = np.zeros(1 << N, dtype=np.uint64)
    for c, q in meas.items():
        key |= ((idx >> np.uint64(q)) & np.uint64(1)) << np.uint64(c)
    out = np.zeros(1 << n_clbits)
    np.add.at(out, key.astype(np.int64), p)
    return out


def format_key(k, width):
    """Counts key: clbit width-1 leftmost, single register => no spaces."""
    return format(int(k), '0%db' % width)


def sample_counts(key_probs, shots, rng, width):
    """Multinomial draw -> {bitstring: count} as get_counts() returns it
    (run_experiment.py:57)."""
    p = np.asarray(key_probs, dtype=np.float64)
    draws = rng.multinomial(shots, p / p.sum())
    return {format_key(k, width): int(c) for k, c in enumerate(draws) if c}


def postselected(psi, n):
    """Exact post-selected pmf and success probability: keep the all-ancillas-0,
    scratch-0 subspace = the first 2^n amplitudes."""
    w = np.abs(psi[: 1 << n]) ** 2
    delta = float(w.sum())
    return w / delta, delta
