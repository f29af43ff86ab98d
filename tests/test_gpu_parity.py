"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle on
the same seeded inputs.  Bars: bit-exact keys / masks / orderings; post-selected
probabilities within 1e-10 (complex128) or 1e-5 (complex64) absolute of brute-force MRF
enumeration and of the oracle statevector; sampled histograms within shot-noise bounds.
"""
import json
import os

import numpy as np
import pytest
from scipy import stats

import engine_emulator as em
from conftest import all_models, has_cuda
from oracle import mrf, program, statevector as sv
from qcmrf_b200 import QCMRF, B200Simulator, _native, fusion, ir, transpile

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason='needs a CUDA device')]

TOL_P = {'double': 1e-10, 'single': 1e-5}
TOL_AMP = {'double': 1e-12, 'single': 2e-6}
#: relative bound on every pmf entry: the absolute 1e-5 of the north star says nothing once p(x) ~ 2^-17
#: (a 33/34-qubit circuit) -- an all-zero vector would pass it
TOL_REL = {'double': 1e-9, 'single': 2e-4}


def assert_pmf(p, pb, delta, db, precision, what=''):
    """post-selected pmf + success probability against brute force: absolute (north star) AND relative."""
    err = np.abs(p - pb).max()
    rel = (np.abs(p - pb) / pb).max()
    assert err < TOL_P[precision] and abs(delta - db) < TOL_P[precision], (what, err, abs(delta - db))
    assert rel < TOL_REL[precision] and abs(delta - db) / db < TOL_REL[precision], (what, rel, abs(delta - db) / db)
    return err, rel


def weissman_tv_bound(K, S, alpha=1e-6):
    """P(TV > bound) <= alpha for S draws over K outcomes."""
    return 0.5 * np.sqrt(2.0 * (K * np.log(2.0) + np.log(1.0 / alpha)) / S)


def counts_to_vec(counts, N):
    v = np.zeros(1 << N)
    for k, c in counts.items():
        assert len(k) == N
        v[int(k, 2)] = c
    return v


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('precision', ['double', 'single'])
def test_all_fixture_models_batched(models, aer_counts, precision):
    """BASELINE config 2: all 210 res_* models in one batch -- exact post-selected pmf +
    8192 shots each, checked against the brute-force MRF distribution."""
    sim = B200Simulator(precision=precision, seed=1984)
    items = list(all_models(models))
    circs = [QCMRF(C, th) for _, _, _, C, th in items]
    res = sim.run(circs, shots=8192).result()
    counts = res.get_counts()
    assert len(counts) == 210
    worst_p = worst_d = 0.0
    tv_ok = 0
    for e, (scale, j, i, C, th) in enumerate(items):
        n, k, N, _ = program.sizes(C)
        p, delta = res.postselected_probabilities(e)
        pb, db, _ = mrf.brute_force_pmf(C, th)
        worst_p = max(worst_p, np.abs(p - pb).max())
        worst_d = max(worst_d, abs(delta - db))
        assert (np.abs(p - pb) / pb).max() < TOL_REL[precision], (e, (np.abs(p - pb) / pb).max())
        psi, meas = sv.run_program(program.qcmrf_program(C, th)[0], N)
        kp = sv.key_probabilities(psi, N, meas)
        obs = counts_to_vec(counts[e], N)
        assert obs.sum() == 8192
        assert obs[kp == 0].sum() == 0                              # support incl. clbit n == '0'
        K = int((kp > 0).sum())
        assert 0.5 * np.abs(obs / 8192 - kp).sum() < weissman_tv_bound(K, 8192)
        tv_ok += 1
    assert worst_p < TOL_P[precision] and worst_d < TOL_P[precision]
    assert tv_ok == 210
    # same seed -> same histograms; the stream is the experiment index, so a sub-batch agrees too
    again = sim.run(circs[:5], shots=8192).result().get_counts()
    assert again == counts[:5]


@pytest.mark.parametrize('precision', ['double', 'single'])
@pytest.mark.parametrize('fusion_mode,block_max,expand_max', [('off', 1, 1), ('clique', 1, 1), ('blocked', 1, 1),
                                                              ('blocked', 2, 2), ('blocked', 4, 4), ('blocked', 5, 5),
                                                              ('blocked', 4, 8), ('blocked', 2, 6)])
def test_statevector_path_all_modes(models, precision, fusion_mode, block_max, expand_max):
    """The large-state kernels (init, blocked multiplexer pass, diag, lazy materialisation) on
    the fixture models: full statevector vs the oracle's gate-by-gate execution."""
    sim = B200Simulator(precision=precision, fusion=fusion_mode, block_max=block_max, expand_max=expand_max,
                        small_batch=False, seed=7)
    worst = 0.0
    for scale, j, i, C, th in all_models(models):
        if i not in (0, 7):
            continue
        n, k, N, _ = program.sizes(C)
        psi, _ = sv.run_program(program.qcmrf_program(C, th)[0], N)
        got = sim.statevector(QCMRF(C, th), precision=precision)
        worst = max(worst, np.abs(got - psi).max())
        p, delta = sim.exact(QCMRF(C, th))
        pb, db, _ = mrf.brute_force_pmf(C, th)
        assert np.abs(p - pb).max() < TOL_P[precision] and abs(delta - db) < TOL_P[precision]
    assert worst < TOL_AMP[precision]


def test_large_path_counts_and_keys(models):
    sim = B200Simulator(precision='double', small_batch=False, seed=99)
    for j in (1, 3, 5):
        C = models['0.5']['GRAPHS'][j]
        th = models['0.5']['THETAS'][str(j)][2]
        n, k, N, _ = program.sizes(C)
        res = sim.run(QCMRF(C, th), shots=20000).result()
        counts = res.get_counts()
        assert isinstance(counts, dict) and sum(counts.values()) == 20000
        psi, meas = sv.run_program(program.qcmrf_program(C, th)[0], N)
        kp = sv.key_probabilities(psi, N, meas)
        obs = counts_to_vec(counts, N)
        assert obs[kp == 0].sum() == 0
        assert 0.5 * np.abs(obs / 2e4 - kp).sum() < weissman_tv_bound(int((kp > 0).sum()), 20000)
        # host-side post-selection of the sampled counts, as the reference does it
        from qcmrf_b200 import extract_probs
        q, succ = extract_probs(counts, n, N - n)
        _, delta = res.postselected_probabilities()
        assert abs(succ - delta) < 0.02
        assert res.get_counts(0) == counts
        json.dumps(res.get_counts())                                # run_experiment.py:60


@pytest.mark.parametrize('graph', [0, 1, 2, 4])
def test_transpiled_circuits(models, graph):
    """run_experiment.py:52-56: transpile to cx/id/rz/sx/x, then run."""
    C = models['0.25']['GRAPHS'][graph]
    th = models['0.25']['THETAS'][str(graph)][5]
    n, k, N, _ = program.sizes(C)
    T = transpile([QCMRF(C, th)], basis_gates=['cx', 'id', 'rz', 'sx', 'x'])
    pb, db, _ = mrf.brute_force_pmf(C, th)
    for sim in (B200Simulator(seed=3), B200Simulator(seed=3, small_batch=False)):
        res = sim.run(T, shots=4096, n_vars=n).result()
        p, delta = res.postselected_probabilities(0)
        assert np.abs(p - pb).max() < 1e-10 and abs(delta - db) < 1e-10
        assert isinstance(res.get_counts(), list) and len(res.get_counts()) == 1
        psi, meas = sv.run_program(program.qcmrf_program(C, th)[0], N)
        kp = sv.key_probabilities(psi, N, meas)
        obs = counts_to_vec(res.get_counts(0), N)
        assert obs[kp == 0].sum() == 0


def test_aer_histograms_vs_gpu_exact_distribution(models, aer_counts):
    """The stored Aer outputs are statistically consistent with the GPU's exact |psi|^2."""
    sim = B200Simulator(precision='double', small_batch=False)
    pvals = []
    for scale, j, i, C, th in all_models(models):
        if i % 3:
            continue
        n, k, N, _ = program.sizes(C)
        psi = sim.statevector(QCMRF(C, th))
        prog = ir.lower(QCMRF(C, th))
        kp = sv.key_probabilities(psi, N, prog.measures)
        Q = aer_counts[scale][10 * j + i]
        obs = counts_to_vec(Q, N)
        assert obs[kp < 1e-20].sum() == 0
        m = kp > 1e-20
        exp = 1e4 * kp[m]
        big = exp >= 5
        o = np.append(obs[m][big], obs[m][~big].sum())
        e = np.append(exp[big], exp[~big].sum())
        if e[-1] == 0:
            o, e = o[:-1], e[:-1]
        pvals.append(stats.chi2.sf(((o - e) ** 2 / e).sum(), len(e) - 1))
    assert min(pvals) > 1e-5


# ------------------------------------------------------------------------------------------
def _random_program(rng, N, n_ops, kinds=('mux', 'diag', 'block', 'swap')):
    """Fully materialised random engine program over N qubits."""
    e = fusion._Emitter()
    qv = np.zeros((N, 4))
    v = rng.randn(N, 2) + 1j * rng.randn(N, 2)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    qv[:, 0], qv[:, 1], qv[:, 2], qv[:, 3] = v[:, 0].real, v[:, 0].imag, v[:, 1].real, v[:, 1].imag
    e.op(fusion.QCM_OP_INIT_PRODUCT, n_in=0, n_out=N, table_off=e.table(qv))

    def rand_u(m):
        a = rng.randn(1 << m, 2, 2) + 1j * rng.randn(1 << m, 2, 2)
        q, _ = np.linalg.qr(a)
        return q

    for _ in range(n_ops):
        kind = kinds[rng.randint(len(kinds))]
        if kind == 'mux':
            t = int(rng.randint(N))
            m = int(rng.randint(0, min(N - 1, 4) + 1))
            ctrl = [int(c) for c in rng.permutation([q for q in range(N) if q != t])[:m]]
            e.op(fusion.QCM_OP_MUX1Q, target=t, ctrl=ctrl, n_in=N, n_out=N,
                 table_off=e.table(fusion._mux_table_f64(rand_u(m))))
        elif kind == 'diag':
            m = int(rng.randint(1, min(N, 5) + 1))
            ctrl = [int(c) for c in rng.permutation(N)[:m]]
            d = np.exp(1j * rng.uniform(0, 2 * np.pi, 1 << m))
            e.op(fusion.QCM_OP_DIAG, ctrl=ctrl, n_in=N, n_out=N, table_off=e.table(fusion._diag_table_f64(d)))
        elif kind == 'swap' and N >= 2:
            a, b = (int(x) for x in rng.permutation(N)[:2])
            e.op(fusion.QCM_OP_SWAP, target=a, ctrl=[b], n_in=N, n_out=N)
        elif kind == 'block' and N >= 3:
            M = int(rng.randint(1, min(5, N - 1) + 1))
            tq = sorted(int(c) for c in rng.permutation(N)[:M])
            others = [q for q in range(N) if q not in tq]
            n_mem = int(rng.randint(1, 7))
            e.op(fusion.QCM_OP_BLOCK, target=M, ctrl=tq, n_in=N, n_out=N, n_ctrl=n_mem)
            for _g in range(n_mem):
                t = tq[rng.randint(M)]
                m = int(rng.randint(0, min(len(others), 3) + 1))
                ctrl = [int(c) for c in rng.permutation(others)[:m]]
                e.op(fusion.QCM_OP_MUX1Q, target=t, ctrl=ctrl, n_in=N, n_out=N,
                     table_off=e.table(fusion._mux_table_f64(rand_u(m))))
    ops, tabs = e.finish()
    return ops, tabs


class _P:
    pass


@pytest.mark.parametrize('precision', ['double', 'single'])
@pytest.mark.parametrize('N', [1, 2, 3, 5, 8, 11, 14])
def test_random_engine_programs_vs_op_semantics(precision, N):
    """Raw C-ABI: random MUX1Q (every target incl. qubit 0 and the shuffle/low range), DIAG,
    SWAP and BLOCK ops on random product states vs the numpy statement of the op semantics."""
    rng = np.random.RandomState(1000 + N)
    with _native.Handle(N, precision) as h:
        for trial in range(6):
            ops, tabs = _random_program(rng, N, 12)
            pl = _P()
            pl.ops, pl.tables, pl.n_phys = ops, tabs, N
            want, act = em.run_plan(pl)
            h.run_program(ops, tabs)
            got = h.get_amplitudes().astype(np.complex128)
            assert np.abs(got - want).max() < (1e-12 if precision == 'double' else 3e-6)
            assert abs(np.sum(np.abs(got) ** 2) - 1.0) < (1e-12 if precision == 'double' else 1e-5)


@pytest.mark.parametrize('precision', ['double', 'single'])
@pytest.mark.parametrize('N', [9, 12, 17])
def test_low_order_target_pass(precision, N):
    """North star (ii): gates whose target lies inside a warp's coalesced access -- the vector slot (qubit 0),
    the lane bits (warp shuffles) and the lane's further vectors (registers) -- run in k_lowq; single gates on
    every low target with index qubits all over the place, blocked passes mixing low and high targets, and
    diagonal members, against the numpy statement of the op semantics."""
    rng = np.random.RandomState(77 + N)
    LB = 9 if precision == 'single' else 8

    def rand_u(m):
        q, _ = np.linalg.qr(rng.randn(1 << m, 2, 2) + 1j * rng.randn(1 << m, 2, 2))
        return q
    tol = 1e-12 if precision == 'double' else 3e-6
    with _native.Handle(N, precision) as h:
        def check(e, expect_lowq):
            ops, tabs = e.finish()
            pl = _P(); pl.ops, pl.tables, pl.n_phys = ops, tabs, N
            want, _ = em.run_plan(pl)
            h.run_program(ops, tabs)
            got = h.get_amplitudes().astype(np.complex128)
            assert np.abs(got - want).max() < tol
            names = h.op_kernels()
            if expect_lowq is not None:
                assert names[-1].startswith('k_lowq') == expect_lowq, names

        def init(e):
            qv = rng.randn(N, 4)
            qv /= np.sqrt((qv ** 2).sum(axis=1, keepdims=True))
            e.op(fusion.QCM_OP_INIT_PRODUCT, n_in=0, n_out=N, table_off=e.table(qv))
        # (1) one gate per low target, 0..3 index qubits drawn from ALL other qubits (lane, vector, high)
        for t in range(min(N, LB + 1)):
            for m in (0, 1, 3, min(5, N - 1)):
                e = fusion._Emitter()
                init(e)
                others = [q for q in range(N) if q != t]
                ctrl = [int(c) for c in rng.permutation(others)[:m]]
                e.op(fusion.QCM_OP_MUX1Q, target=t, ctrl=ctrl, n_in=N, n_out=N, table_off=e.table(fusion._mux_table_f64(rand_u(len(ctrl)))))
                check(e, t < 6 and N >= LB)
        # (2) blocked passes: all-low, one high, two high targets; repeated targets; diagonal members
        if N >= 12:
            for tq in ([0, 1, 2, 3, 4], [0, 5, 8], [1, 3, N - 1], [0, 2, N - 3, N - 1], [2, 6, 7, N - 2], [4, 9, 10]):
                tq = sorted(set(t for t in tq if t < N))
                others = [q for q in range(N) if q not in tq]
                e = fusion._Emitter()
                init(e)
                n_mem = 6
                e.op(fusion.QCM_OP_BLOCK, target=len(tq), ctrl=tq, n_in=N, n_out=N, n_ctrl=n_mem)
                for g in range(n_mem):
                    m = int(rng.randint(0, 4))
                    ctrl = [int(c) for c in rng.permutation(others)[:m]]
                    if g == 2:
                        d = np.exp(1j * rng.uniform(0, 2 * np.pi, 1 << m))
                        e.op(fusion.QCM_OP_DIAG, ctrl=ctrl, n_in=N, n_out=N, table_off=e.table(fusion._diag_table_f64(d)))
                    else:
                        e.op(fusion.QCM_OP_MUX1Q, target=tq[g % len(tq)], ctrl=ctrl, n_in=N, n_out=N,
                             table_off=e.table(fusion._mux_table_f64(rand_u(m))))
                check(e, None)


@pytest.mark.parametrize('precision', ['double', 'single'])
@pytest.mark.parametrize('N', [3, 9, 12, 15])
def test_batched_handle_random_programs(precision, N):
    """qcm_create_batched: B states in one allocation, every kernel launched once with the sweep point as the
    second grid dimension, shared ops and per-point coefficient tables -- every point must equal the numpy
    statement of the op semantics run on its own tables (tables need not be unitary for that)."""
    rng = np.random.RandomState(4000 + N)
    B = 5
    with _native.Handle(N, precision, batch=B) as h:
        for trial in range(4):
            ops, tabs = _random_program(rng, N, 10)
            rows = [tabs] + [tabs * (1.0 + 0.3 * rng.standard_normal(tabs.shape)) for _ in range(B - 1)]
            h.run_program(ops, np.stack(rows))
            for y in range(B):
                pl = _P()
                pl.ops, pl.tables, pl.n_phys = ops, rows[y], N
                want, act = em.run_plan(pl)
                h.batch_select(y)
                got = h.get_amplitudes().astype(np.complex128)
                scale = max(np.abs(want).max(), 1e-30)
                assert np.abs(got - want).max() / scale < (1e-11 if precision == 'double' else 2e-5), (trial, y)
            # post-selection and sampling per point: mask on the two highest qubits, raw indices
            if N >= 3:
                mask = 0b11 << (N - 2)
                probs, kept = h.postselect(mask, 0, N - 2)
                keys = h.sample_batched(512, 11, np.arange(B) + 3)
                for y in range(B):
                    pl = _P()
                    pl.ops, pl.tables, pl.n_phys = ops, rows[y], N
                    want, act = em.run_plan(pl)
                    w = np.abs(want) ** 2
                    assert np.allclose(probs[y], w[: 1 << (N - 2)], rtol=1e-9 if precision == 'double' else 2e-4, atol=1e-300)
                    assert abs(kept[y] - w[: 1 << (N - 2)].sum()) <= (1e-9 if precision == 'double' else 2e-4) * w.sum()
                    assert (w[keys[y].astype(np.int64)] > 0).all()
                kept2 = h.postselect_resident(mask, 0, N - 2)
                assert np.array_equal(kept2, kept)
                assert np.array_equal(h.fetch_probs(B - 1, N - 2), probs[B - 1])


@pytest.mark.parametrize('width', ['release', 'full'])
@pytest.mark.parametrize('precision', ['double', 'single'])
def test_beta_sweep_through_one_batched_handle(width, precision):
    """BASELINE config 3 in small: a chain MRF swept over beta (QCMRF.py:21,154) -- the list goes through ONE batched
    handle (O(10) launches for the sweep); every point against brute force, shots through the per-clique marginals,
    and (release width) bit-identical keys to the one-circuit-at-a-time path."""
    from qcmrf_b200 import workloads
    n = 12 if width == 'release' else 8                          # N = 24 / 16 qubits at Aer width
    C = workloads.chain(n)
    th = workloads.theta_for(C, seed=9)
    betas = [(j + 1) / 4.0 for j in range(7)]
    shots = 20000
    sim = B200Simulator(precision=precision, width=width, seed=21, small_batch=False)
    one = B200Simulator(precision=precision, width=width, seed=21, small_batch=False, sweep_batch=False)
    l0 = sim.kernel_launches()
    res = sim.run([QCMRF(C, th, beta=b) for b in betas], shots=shots).result()
    launches = sim.kernel_launches() - l0
    assert all(res.metadata(i)['path'] == 'sweep' for i in range(len(betas)))
    assert launches <= 24, launches                              # not O(points)
    for i, b in enumerate(betas):
        pb, db, _ = mrf.brute_force_pmf(C, th, beta=b)
        p, d = res.postselected_probabilities(i)
        assert_pmf(p, pb, d, db, precision, (width, b))
        counts = res.get_counts(i)
        assert sum(counts.values()) == shots
        # per-clique (x_C, ancilla) marginals of the sampled keys vs the exact law (bench.shot_marginal_tv's statistic)
        keys = np.array([int(k, 2) for k in counts], dtype=np.uint64)
        cnt = np.array(list(counts.values()), dtype=np.float64)
        assert not ((keys >> np.uint64(n)) & np.uint64(1)).any()
        for ii, cl in enumerate(C):
            y = np.zeros(len(keys), dtype=np.int64)
            for v in cl:
                y = (y << 1) | ((keys >> np.uint64(n - 1 - v)) & np.uint64(1)).astype(np.int64)
            a = ((keys >> np.uint64(n + 1 + ii)) & np.uint64(1)).astype(np.int64)
            obs = np.bincount(y + (a << 2), weights=cnt, minlength=8) / shots
            w = np.exp(b * np.asarray(th[4 * ii:4 * ii + 4]))
            exact = np.concatenate([w, 1.0 - w]) / 4
            assert 0.5 * np.abs(obs - exact).sum() < weissman_tv_bound(8, shots, 1e-6 / (len(C) * len(betas)))
        if width == 'release' and i in (0, 3):
            r1 = one.run(QCMRF(C, th, beta=b), shots=shots, stream_ids=[i]).result()
            assert r1.metadata(0)['path'] == 'statevector'
            assert r1.get_counts(0) == counts
            assert np.allclose(r1.postselected_probabilities(0)[0], p, rtol=1e-13 if precision == 'double' else 1e-6, atol=0)
    sim.close()
    one.close()


@pytest.mark.parametrize('precision', ['double', 'single'])
def test_diagonal_block_is_one_sweep(precision):
    """A BLOCK header without targets: all of its DIAG members in one sweep (the projection passes of a
    release-width circuit), index qubits anywhere incl. qubit 0."""
    N = 13
    rng = np.random.RandomState(8)
    with _native.Handle(N, precision) as h:
        e = fusion._Emitter()
        qv = rng.randn(N, 4)
        qv /= np.sqrt((qv ** 2).sum(axis=1, keepdims=True))
        e.op(fusion.QCM_OP_INIT_PRODUCT, n_in=0, n_out=N, table_off=e.table(qv))
        ctrls = [[0, 3, 12], [1, 2, 4, 5, 6, 7, 8, 9, 10, 11], [5], []]
        e.op(fusion.QCM_OP_BLOCK, target=0, ctrl=(), n_in=N, n_out=N, n_ctrl=len(ctrls))
        for c in ctrls:
            d = rng.uniform(0.2, 1.0, 1 << len(c)) * np.exp(1j * rng.uniform(0, 6, 1 << len(c)))
            e.op(fusion.QCM_OP_DIAG, ctrl=c, n_in=N, n_out=N, table_off=e.table(fusion._diag_table_f64(d)))
        ops, tabs = e.finish()
        pl = _P(); pl.ops, pl.tables, pl.n_phys = ops, tabs, N
        want, _ = em.run_plan(pl)
        h.run_program(ops, tabs)
        got = h.get_amplitudes().astype(np.complex128)
        assert np.abs(got - want).max() < (1e-12 if precision == 'double' else 3e-6)
        assert len(h.op_profile()) == 2 and h.op_kernels()[-1].startswith('k_diag_multi')
        # index qubits in descending order: served by one pass per member, same result
        e = fusion._Emitter()
        e.op(fusion.QCM_OP_INIT_PRODUCT, n_in=0, n_out=N, table_off=e.table(qv))
        e.op(fusion.QCM_OP_BLOCK, target=0, ctrl=(), n_in=N, n_out=N, n_ctrl=2)
        for c in ([7, 3, 0], [2, 1]):
            d = np.exp(1j * rng.uniform(0, 6, 1 << len(c)))
            e.op(fusion.QCM_OP_DIAG, ctrl=c, n_in=N, n_out=N, table_off=e.table(fusion._diag_table_f64(d)))
        ops, tabs = e.finish()
        pl = _P(); pl.ops, pl.tables, pl.n_phys = ops, tabs, N
        want, _ = em.run_plan(pl)
        h.run_program(ops, tabs)
        assert np.abs(h.get_amplitudes().astype(np.complex128) - want).max() < (1e-12 if precision == 'double' else 3e-6)


@pytest.mark.parametrize('batch', [1, 3])
def test_product_state_is_sampled_qubit_by_qubit(batch):
    """A program that is INIT_PRODUCT and nothing else leaves a product state: shots are one Bernoulli draw per qubit
    (k_sample_product: no sum tree, no pass over the state) -- histogram against the exact product law, clbit map,
    seed determinism; any further op switches back to the tree sampler."""
    N, shots = 8, 200000
    rng = np.random.RandomState(31)
    qvs = []
    for _ in range(batch):
        qv = rng.randn(N, 4)
        qv /= np.sqrt((qv ** 2).sum(axis=1, keepdims=True))
        qvs.append(qv)
    with _native.Handle(N, 'double', batch=batch) as h:
        e = fusion._Emitter()
        e.op(fusion.QCM_OP_INIT_PRODUCT, n_in=0, n_out=N, table_off=e.table(qvs[0]))
        ops, tabs = e.finish()
        rows = np.stack([np.concatenate([q.reshape(-1), np.zeros(tabs.size - q.size)]) for q in qvs])
        l0 = h.timing()['kernel_launches']
        if batch == 1:
            h.run_program(ops, rows[0])
            keys = [h.sample(shots, 5, 2, None)]
            again = [h.sample(shots, 5, 2, None)]
            cl = h.sample(1000, 5, 2, np.array([3, -1, 0], dtype=np.int32))
            raw = h.sample(1000, 5, 2, None)
            assert np.array_equal(cl, ((raw >> 3) & 1) | (((raw >> 0) & 1) << 2))
        else:
            h.run_program(ops, rows)
            keys = list(h.sample_batched(shots, 5, np.array([2, 9, 11]), None))
            again = list(h.sample_batched(shots, 5, np.array([2, 9, 11]), None))
        # init tables + init + one sampling kernel per call: no tree kernels were launched
        assert h.timing()['kernel_launches'] - l0 == 2 + (4 if batch == 1 else 2)
        for y in range(batch):
            assert np.array_equal(keys[y], again[y])
            p1 = (qvs[y][:, 2] ** 2 + qvs[y][:, 3] ** 2)
            idx = np.arange(1 << N)
            exact = np.ones(1 << N)
            for q in range(N):
                exact *= np.where((idx >> q) & 1, p1[q], 1 - p1[q])
            obs = np.bincount(keys[y].astype(np.int64), minlength=1 << N) / shots
            assert 0.5 * np.abs(obs - exact).sum() < weissman_tv_bound(1 << N, shots)
        if batch == 1:
            assert np.array_equal(keys[0], _native_sample_again_after_noop(h, ops, rows[0], shots)) is False


def _native_sample_again_after_noop(h, ops, tabs, shots):
    """The same state after one more (identity) sweep is no longer known to be a product state: the tree sampler
    serves it -- other random numbers, same distribution."""
    e = fusion._Emitter()
    e.op(fusion.QCM_OP_INIT_PRODUCT, n_in=0, n_out=8, table_off=e.table(tabs[:32].reshape(8, 4)))
    e.op(fusion.QCM_OP_DIAG, ctrl=[1], n_in=8, n_out=8, table_off=e.table(np.array([[1.0, 0.0], [1.0, 0.0]])))
    ops2, tabs2 = e.finish()
    h.run_program(ops2, tabs2)
    return h.sample(shots, 5, 2, None)


@pytest.mark.parametrize('precision', ['double', 'single'])
def test_lazy_materialisation_ops(precision):
    """Ops that materialise qubits: zero-input targets are never read (the buffer holds NaN
    there), EXTEND zero-fills, get_amplitudes reports implicit zeros."""
    N = 12
    rng = np.random.RandomState(5)
    with _native.Handle(N, precision) as h:
        h.set_amplitudes(np.full(1 << N, np.nan + 1j * np.nan), 0, n_active=0)
        e = fusion._Emitter()
        qv = np.zeros((4, 4)); qv[:, 0] = qv[:, 2] = np.sqrt(0.5)
        e.op(fusion.QCM_OP_INIT_PRODUCT, n_in=0, n_out=4, table_off=e.table(qv))

        def rand_u(m):
            q, _ = np.linalg.qr(rng.randn(1 << m, 2, 2) + 1j * rng.randn(1 << m, 2, 2))
            return q
        e.op(fusion.QCM_OP_MUX1Q, target=4, ctrl=[0, 2], n_in=4, n_out=5, table_off=e.table(fusion._mux_table_f64(rand_u(2))))
        e.op(fusion.QCM_OP_BLOCK, target=4, ctrl=[3, 5, 6, 7], n_in=5, n_out=8, n_ctrl=5)
        for t, c in ((5, [0, 1]), (3, [4]), (6, [2]), (7, [1, 4]), (5, [])):
            e.op(fusion.QCM_OP_MUX1Q, target=t, ctrl=c, n_in=5, n_out=8, table_off=e.table(fusion._mux_table_f64(rand_u(len(c)))))
        e.op(fusion.QCM_OP_EXTEND, n_in=8, n_out=10)
        e.op(fusion.QCM_OP_DIAG, ctrl=[9, 0, 5], n_in=10, n_out=10,
             table_off=e.table(fusion._diag_table_f64(np.exp(1j * rng.uniform(0, 6, 8)))))
        ops, tabs = e.finish()
        pl = _P(); pl.ops, pl.tables, pl.n_phys = ops, tabs, N
        want, act = em.run_plan(pl)
        h.run_program(ops, tabs)
        assert h.get_active() == act == 10
        got = h.get_amplitudes().astype(np.complex128)
        assert not np.isnan(got).any()
        assert np.abs(got - want).max() < (1e-12 if precision == 'double' else 3e-6)
        assert np.all(got[1 << 10:] == 0)
        t = h.timing()
        assert t['kernel_launches'] >= 6 and t['bytes_written'] > 0


def test_invalid_programs_are_rejected():
    with _native.Handle(6, 'double') as h:
        e = fusion._Emitter()
        qv = np.zeros((6, 4)); qv[:, 0] = 1
        e.op(fusion.QCM_OP_INIT_PRODUCT, n_in=0, n_out=6, table_off=e.table(qv))
        e.op(fusion.QCM_OP_BLOCK, target=2, ctrl=[1, 2], n_in=6, n_out=6, n_ctrl=1)
        e.op(fusion.QCM_OP_MUX1Q, target=1, ctrl=[2], n_in=6, n_out=6, table_off=e.table(np.zeros(16)))  # ctrl is a block target
        ops, tabs = e.finish()
        with pytest.raises(_native.NativeError) as ei:
            h.run_program(ops, tabs)
        assert ei.value.code == -1
        e = fusion._Emitter()
        e.op(fusion.QCM_OP_MUX1Q, target=9, ctrl=[], n_in=0, n_out=6, table_off=e.table(np.zeros(8)))
        ops, tabs = e.finish()
        with pytest.raises(_native.NativeError):
            h.run_program(ops, tabs)
        e = fusion._Emitter()
        e.op(99, n_in=0, n_out=0)
        ops, tabs = e.finish()
        with pytest.raises(_native.NativeError):
            h.run_program(ops, np.zeros(1))


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('precision', ['double', 'single'])
@pytest.mark.parametrize('N', [3, 10, 13, 21])
def test_sampler(precision, N):
    """Sum tree + Philox + warp search: deterministic, never returns a zero-probability state,
    matches the exact distribution, honours the clbit map (N=21 exercises a 2-level tree)."""
    rng = np.random.RandomState(N)
    dim = 1 << N
    if N <= 13:
        psi = rng.randn(dim) + 1j * rng.randn(dim)
        psi[rng.rand(dim) < 0.3] = 0
    else:                                            # sparse-ish support so TV is testable
        psi = np.zeros(dim, dtype=np.complex128)
        support = rng.choice(dim, 3000, replace=False)
        psi[support] = rng.randn(3000) + 1j * rng.randn(3000)
    psi /= np.linalg.norm(psi)
    with _native.Handle(N, precision) as h:
        h.set_amplitudes(psi)
        S = 200000
        a = h.sample(S, seed=1984, stream_id=3)
        b = h.sample(S, seed=1984, stream_id=3)
        c = h.sample(S, seed=1984, stream_id=4)
        assert np.array_equal(a, b) and not np.array_equal(a, c)
        assert a.max() < dim
        pr = np.abs(h.get_amplitudes().astype(np.complex128)) ** 2
        assert np.all(pr[a.astype(np.int64)] > 0)
        vals, cnt = np.unique(a, return_counts=True)
        K = int((pr > 0).sum())
        emp = np.zeros(dim); emp[vals.astype(np.int64)] = cnt / S
        assert 0.5 * np.abs(emp - pr / pr.sum()).sum() < weissman_tv_bound(K, S)
        # clbit map: reverse the low 3 qubits, leave clbit 3 unmeasured, clbit 4 <- top qubit
        cmap = np.array([2, 1, 0, -1, N - 1], dtype=np.int32)
        k = h.sample(S, seed=1984, stream_id=3, clbit_qubit=cmap)
        ai = a.astype(np.int64)
        want = ((ai >> 2) & 1) | (((ai >> 1) & 1) << 1) | ((ai & 1) << 2) | (((ai >> (N - 1)) & 1) << 4)
        assert np.array_equal(k.astype(np.int64), want)
        # the first uniforms are Philox4x32-10 of (seed, stream, shot): spot-check shot 0 via the CDF
        small = h.sample(1, seed=1984, stream_id=3)
        assert small[0] == a[0]


def test_sharded_sampling_equals_single_gpu():
    """Two 'ranks' (two handles on this GPU, run one after the other) draw the same Philox
    stream and split the shots exactly as the single-state sampler does."""
    N = 21
    rng = np.random.RandomState(42)
    psi = rng.randn(1 << N) + 1j * rng.randn(1 << N)
    psi /= np.linalg.norm(psi)
    S = 50000
    with _native.Handle(N, 'double') as h:
        h.set_amplitudes(psi)
        ref = h.sample(S, seed=5, stream_id=1)
    half = 1 << (N - 1)
    keys, mines, masses = [], [], []
    hs = []
    for r in range(2):
        h = _native.Handle(N - 1, 'double')
        h.set_shard(1, r)
        h.set_amplitudes(psi[r * half:(r + 1) * half])
        masses.append(h.sample_prepare())
        hs.append(h)
    for r, h in enumerate(hs):
        k, m = h.sample_sharded(S, 5, 1, masses)
        keys.append(k); mines.append(m)
        h.close()
    assert np.all(mines[0] ^ mines[1])                   # every shot resolved by exactly one rank
    merged = np.where(mines[0], keys[0], keys[1])
    assert np.array_equal(merged, ref)


def test_postselect_general_mask():
    N = 9
    rng = np.random.RandomState(3)
    psi = rng.randn(1 << N) + 1j * rng.randn(1 << N)
    psi /= np.linalg.norm(psi)
    pr = np.abs(psi) ** 2
    idx = np.arange(1 << N)
    with _native.Handle(N, 'double') as h:
        h.set_amplitudes(psi)
        for mask, value, nb in ((0b110000000, 0, 7), (0b101000100, 0b001000100, 4), (0, 0, 3), (0b111111111, 5, 2)):
            probs, kept = h.postselect(mask, value, nb)
            sel = (idx & mask) == value
            want = np.zeros(1 << nb)
            np.add.at(want, idx[sel] & ((1 << nb) - 1), pr[sel])
            assert abs(kept - pr[sel].sum()) < 1e-14
            assert np.abs(probs - want).max() < 1e-14


def test_mid_size_tree_mrf_properties():
    """A 13-variable random-tree MRF (N = 26 total qubits, complex64): blocked execution;
    post-selected pmf vs brute force at the fp32 tolerance, unit norm, success probability."""
    rng = np.random.RandomState(1984)
    n = 13
    cliques = [[int(rng.randint(0, v)), v] for v in range(1, n)]
    th = -np.abs(rng.randn(4 * len(cliques))) * 0.5
    circ = QCMRF(cliques, list(th))
    assert circ.num_qubits == 26
    pb, db, _ = mrf.brute_force_pmf(cliques, th)
    for bm in (4, 5):
        sim = B200Simulator(precision='single', fusion='blocked', block_max=bm, expand_max=bm, seed=1)
        res = sim.run(circ, shots=10000).result()
        p, delta = res.postselected_probabilities(0)
        assert_pmf(p, pb, delta, db, 'single')
        meta = res.metadata(0)
        assert meta['path'] == 'statevector' and meta['n_phys'] == 25
        assert meta['passes'] == 1 + -(-12 // bm)
        counts = res.get_counts(0)
        kept = sum(v for k, v in counts.items() if int(k, 2) < (1 << n))
        assert abs(kept / 1e4 - db) < 5 * np.sqrt(db * (1 - db) / 1e4) + 1e-3
        for k in counts:
            assert len(k) == 26 and k[26 - 1 - n] == '0'
        sim.close()
    # dense in-place gate passes (one per clique, full 2^26 width) give the same answer
    sim = B200Simulator(precision='single', fusion='clique', seed=1)
    p2, d2 = sim.exact(circ)
    assert_pmf(p2, pb, d2, db, 'single')
    sim.close()


# ------------------------------------------------------------------------------------------
def _expansion_program(rng, n0, Ms, with_diag, flags_last):
    """INIT over n0 qubits, then pure expansion passes (every block qubit new, one MUX1Q each,
    optional diagonal members): the fast path of the lazily materialised schedule."""
    e = fusion._Emitter()
    v = rng.randn(n0, 2) + 1j * rng.randn(n0, 2)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    qv = np.stack([v[:, 0].real, v[:, 0].imag, v[:, 1].real, v[:, 1].imag], axis=1)
    e.op(fusion.QCM_OP_INIT_PRODUCT, n_in=0, n_out=n0, table_off=e.table(qv))

    def rand_u(m):
        q, _ = np.linalg.qr(rng.randn(1 << m, 2, 2) + 1j * rng.randn(1 << m, 2, 2))
        return q
    act = n0
    for M in Ms:
        tq = list(range(act, act + M))
        members = []
        for t in rng.permutation(tq):
            m = int(rng.randint(0, 3))
            ctrl = [int(c) for c in rng.permutation(act)[:m]]
            if act >= 1 and rng.rand() < 0.3 and 0 not in ctrl and m < 2:
                ctrl.append(0)                                       # qubit 0 splits the 2-amplitude vector
            members.append((fusion.QCM_OP_MUX1Q, int(t), ctrl, fusion._mux_table_f64(rand_u(len(ctrl)))))
        if with_diag:
            ctrl = [int(c) for c in rng.permutation(act)[:2]]
            d = np.exp(1j * rng.uniform(0, 6, 4))
            members.insert(int(rng.randint(len(members) + 1)), (fusion.QCM_OP_DIAG, 0, ctrl, fusion._diag_table_f64(d)))
        if len(members) == 1:
            k, t, c, tab = members[0]
            e.op(k, target=t, ctrl=c, n_in=act, n_out=act + M, table_off=e.table(tab))
        else:
            e.op(fusion.QCM_OP_BLOCK, target=M, ctrl=tq, n_in=act, n_out=act + M, n_ctrl=len(members))
            for k, t, c, tab in members:
                e.op(k, target=t, ctrl=c, n_in=act, n_out=act + M, table_off=e.table(tab))
        act += M
    ops, tabs = e.finish()
    if flags_last:
        fusion._flag_last_pass(ops)
    return ops, tabs, act


@pytest.mark.parametrize('precision', ['double', 'single'])
@pytest.mark.parametrize('with_diag', [False, True])
def test_expansion_fast_path(precision, with_diag):
    rng = np.random.RandomState(77)
    for n0, Ms in ((1, [1, 2, 3]), (3, [4, 5]), (2, [5, 1, 4]), (6, [3, 3, 2]), (0, [2, 2]),
                   (3, [6, 8]), (2, [7, 4]), (5, [8, 5]), (0, [8]), (1, [8, 8])):
        ops, tabs, act = _expansion_program(rng, n0, Ms, with_diag, False)
        with _native.Handle(act, precision) as h:
            h.set_amplitudes(np.full(1 << act, np.nan + 1j * np.nan), 0, n_active=0)
            pl = _P(); pl.ops, pl.tables, pl.n_phys = ops, tabs, act
            want, a2 = em.run_plan(pl)
            h.run_program(ops, tabs)
            got = h.get_amplitudes().astype(np.complex128)
            assert h.get_active() == a2 == act
            assert np.abs(got - want).max() < (1e-12 if precision == 'double' else 3e-6)


@pytest.mark.parametrize('precision', ['double', 'single'])
@pytest.mark.parametrize('Ms', [[3, 4], [2, 8], [7]])
def test_sampling_checkpoint_matches_full_tree(precision, Ms):
    """Shots drawn through the checkpoint tree (built before the final expansion pass, new
    qubits sampled conditionally) follow the same distribution as shots drawn from a tree
    over the whole result."""
    rng = np.random.RandomState(5)
    n0 = 11
    ops, tabs, act = _expansion_program(rng, n0, Ms, False, True)
    assert ops['flags'].sum() == 3            # checkpoint + rotated-output permission on the last pass
    S = 400000
    with _native.Handle(act, precision) as h:
        h.run_program(ops, tabs)
        pr = np.abs(h.get_amplitudes().astype(np.complex128)) ** 2
        a = h.sample(S, seed=9, stream_id=0)
        a2 = h.sample(S, seed=9, stream_id=0)
        plain = ops.copy(); plain['flags'] = 0
        h.run_program(plain, tabs)
        b = h.sample(S, seed=9, stream_id=0)
    assert np.array_equal(a, a2)
    assert np.all(pr[a.astype(np.int64)] > 0)
    # the two trees order the outcomes differently (x-major vs index order), so the same uniforms
    # give different shots: compare both histograms with the exact distribution
    assert not np.array_equal(a, b)
    idx = np.arange(1 << act)
    M = Ms[-1]
    for keys in (a, b):
        k = keys.astype(np.int64)
        emp = np.bincount(k, minlength=1 << act) / S
        assert 0.5 * np.abs(emp - pr / pr.sum()).sum() < weissman_tv_bound(int((pr > 0).sum()), S)
        # sharper: marginals over the qubits of the final pass (drawn conditionally) and over the low 8 qubits
        for shift, bits in ((act - M, M), (0, 8), (act - M - 3, 6)):
            want = np.bincount((idx >> shift) & ((1 << bits) - 1), weights=pr / pr.sum(), minlength=1 << bits)
            got = np.bincount((k >> shift) & ((1 << bits) - 1), minlength=1 << bits) / S
            assert 0.5 * np.abs(got - want).sum() < weissman_tv_bound(1 << bits, S)


@pytest.mark.parametrize('precision', ['double', 'single'])
def test_release_mode_chain(precision):
    """Measure-and-release (BASELINE config 2's enabler): a 12-variable chain (N = 24 qubits at Aer
    width) with only the 12 variable qubits stored; pmf/delta vs brute force, full-width keys vs the
    product-form key distribution, agreement with the full-width engine run."""
    n = 12
    C = [[i, i + 1] for i in range(n - 1)]
    rng = np.random.RandomState(4)
    th = list(-np.abs(rng.randn(4 * len(C))) * 0.5)
    pb, db, _ = mrf.brute_force_pmf(C, th)
    sim = B200Simulator(precision=precision, width='release', seed=8)
    res = sim.run(QCMRF(C, th, beta=0.7), shots=0).result()
    pbeta, dbeta, _ = mrf.brute_force_pmf(C, th, beta=0.7)
    p, d = res.postselected_probabilities()
    assert np.abs(p - pbeta).max() < TOL_P[precision] and abs(d - dbeta) < TOL_P[precision]
    res = sim.run(QCMRF(C, th), shots=50000).result()
    meta = res.metadata(0)
    assert meta['width'] == 'release' and meta['n_phys'] == n and meta['released_qubits'] == n - 1
    p, d = res.postselected_probabilities()
    assert np.abs(p - pb).max() < TOL_P[precision] and abs(d - db) < TOL_P[precision]
    counts = res.get_counts()
    N = 2 * n
    assert sum(counts.values()) == 50000 and all(len(k) == N and k[N - 1 - n] == '0' for k in counts)
    # post-selected shots reproduce the pmf; success fraction reproduces delta
    kept = {k: v for k, v in counts.items() if int(k, 2) < (1 << n)}
    succ = sum(kept.values()) / 5e4
    assert abs(succ - db) < 5 * np.sqrt(db * (1 - db) / 5e4) + 1e-3
    q = np.zeros(1 << n)
    for k, v in kept.items():
        q[int(k, 2)] = v
    q /= q.sum()
    assert 0.5 * np.abs(q - pb).sum() < weissman_tv_bound(1 << n, sum(kept.values()))
    # each ancilla's marginal failure rate matches the full-width engine's exact state
    full = B200Simulator(precision='double', small_batch=False)
    psi = full.statevector(QCMRF(C, th))
    pr = np.abs(psi) ** 2
    idx = np.arange(1 << N)
    for ii in (0, 5, n - 2):
        qb = n + 1 + ii
        exact = pr[((idx >> qb) & 1) == 1].sum()
        got = sum(v for k, v in counts.items() if k[N - 1 - qb] == '1') / 5e4
        assert abs(got - exact) < 5 * np.sqrt(exact * (1 - exact) / 5e4) + 1e-3
    sim.close(); full.close()


def _free_gpu_bytes():
    try:
        import torch
        return torch.cuda.mem_get_info()[0]
    except Exception:
        return 0


@pytest.mark.skipif(has_cuda() and _free_gpu_bytes() < 72 * 2 ** 30, reason='needs ~66 GiB of free device memory')
def test_full_size_33_qubit_state():
    """BASELINE config 3 at full size (33 total qubits, complex64, 32 stored qubits = 32 GiB): the
    post-selected pmf and delta against brute-force enumeration (2^16 states), shot bookkeeping, and
    the dense one-pass-per-clique schedule (33 stored qubits, 64 GiB, identity layout) against the
    lazily materialised one."""
    from qcmrf_b200 import workloads
    C, N = workloads.named('q33')
    assert N == 33
    th = workloads.theta_for(C)
    n = 16
    pb, db, _ = mrf.brute_force_pmf(C, th)
    sim = B200Simulator(precision='single', seed=2024, small_batch=False)
    res = sim.run(QCMRF(C, th), shots=100000).result()
    p, delta = res.postselected_probabilities()
    assert abs(p.sum() - 1.0) < 1e-6
    assert_pmf(p, pb, delta, db, 'single', 'q33 lazy')
    assert np.argmax(p) == np.argmax(pb)
    meta = res.metadata()
    assert meta['n_phys'] == 32 and meta['passes'] == 3             # init + two 8-qubit expansion passes
    counts = res.get_counts()
    assert sum(counts.values()) == 100000
    kept = 0
    for k, v in counts.items():
        assert len(k) == 33 and k[33 - 1 - n] == '0'                  # clbit n (scratch qubit) never written
        if int(k, 2) < (1 << n):
            kept += v
    assert abs(kept / 1e5 - db) < 5 * np.sqrt(db * (1 - db) / 1e5) + 1e-4
    # each ancilla's failure marginal: P(a_c = 1) = 2^-n sum_x sin^2(2 gamma_{c, x_c})
    gam = program.theta_to_gamma(np.asarray(th))
    for ii in (0, 7, 15):
        want = float(np.mean(np.sin(2 * gam[4 * ii:4 * ii + 4]) ** 2))   # pair clique: 4 states, uniform x
        got = sum(v for k, v in counts.items() if k[33 - 1 - (n + 1 + ii)] == '1') / 1e5
        assert abs(got - want) < 5 * np.sqrt(want * (1 - want) / 1e5) + 1e-4
    sim.close()
    dense = B200Simulator(precision='single', fusion='clique', seed=2024, small_batch=False)
    p2, d2 = dense.exact(QCMRF(C, th))
    assert_pmf(p2, pb, d2, db, 'single', 'q33 dense')
    assert (np.abs(p2 - p) / pb).max() < 2e-4
    dense.close()


@pytest.mark.parametrize('precision', ['double', 'single'])
@pytest.mark.parametrize('s', [1, 2, 3])
@pytest.mark.parametrize('variant', [('tma', 1, 32), ('tma', 2, 4), ('ldg', 2, 0), ('ldg', 4, 0)])
def test_gather_block_on_virtual_peers(precision, s, variant, monkeypatch):
    """qcm_run_gather_block (fused qubit swap + blocked pass that reads the peers' shards): all 2^s
    'ranks' live on this GPU, so the peer pointers are ordinary device pointers; expected = numpy qubit
    swap of the full state followed by the op semantics (engine emulator) per rank.  Variants: the
    TMA-ring kernel (deep enough that the ring wraps) and the plain 128-bit-load kernel."""
    import torch
    monkeypatch.setenv('QCM_GATHER', variant[0])
    monkeypatch.setenv('QCM_GATHER_U', str(variant[1]))
    monkeypatch.setenv('QCM_GATHER_K', str(variant[2] or 16))
    rng = np.random.RandomState(50 + s)
    nl = 15
    N = nl + s
    world = 1 << s
    cdt = np.complex128 if precision == 'double' else np.complex64
    tdt = torch.float64 if precision == 'double' else torch.float32
    psi = (rng.randn(1 << N) + 1j * rng.randn(1 << N))
    psi /= np.linalg.norm(psi)
    # the block: targets = the s highest local qubits after the swap; controls among low local and global qubits
    e = fusion._Emitter()
    tq = list(range(nl - s, nl))

    def rand_u(m):
        q, _ = np.linalg.qr(rng.randn(1 << m, 2, 2) + 1j * rng.randn(1 << m, 2, 2))
        return q
    members = []
    for t in tq + [tq[0]]:
        ctrl = [int(c) for c in rng.permutation(nl - s)[:2]] + ([int(nl + rng.randint(s))] if rng.rand() < 0.6 else [])
        if rng.rand() < 0.4 and 0 not in ctrl:
            ctrl[0] = 0
        members.append((fusion.QCM_OP_MUX1Q, t, ctrl, fusion._mux_table_f64(rand_u(len(ctrl)))))
    members.append((fusion.QCM_OP_DIAG, 0, [1, nl - s - 1], fusion._diag_table_f64(np.exp(1j * rng.uniform(0, 6, 4)))))
    if s == 1:
        members = members[:1]
        k, t, c, tab = members[0]
        e.op(k, target=t, ctrl=c, n_in=nl, n_out=nl, table_off=e.table(tab))
    else:
        e.op(fusion.QCM_OP_BLOCK, target=s, ctrl=tq, n_in=nl, n_out=nl, n_ctrl=len(members))
        for k, t, c, tab in members:
            e.op(k, target=t, ctrl=c, n_in=nl, n_out=nl, table_off=e.table(tab))
    ops, tabs = e.finish()
    # expected: swap global qubits (nl .. nl+s-1) with local qubits (nl-s .. nl-1), then the block per rank
    idx = np.arange(1 << N)
    lo_mask = (1 << (nl - s)) - 1
    a = (idx >> (nl - s)) & (world - 1)
    b = idx >> nl
    swapped = psi[(idx & lo_mask) | (b << (nl - s)) | (a << nl)]
    old = [torch.from_numpy(np.ascontiguousarray(psi[r << nl:(r + 1) << nl].astype(cdt)).view(np.float64 if precision == 'double' else np.float32)).cuda()
           for r in range(world)]
    new = [torch.empty_like(x) for x in old]
    slab_bytes = (1 << (nl - s)) * np.dtype(cdt).itemsize
    worst = 0.0
    for r in range(world):
        with _native.Handle(nl, precision, ext_state_ptr=old[r].data_ptr()) as h:
            h.set_shard(s, r)
            h.set_active(nl)
            src = [old[j].data_ptr() + r * slab_bytes for j in range(world)]
            h.run_gather_block(ops, tabs, src, new[r].data_ptr())
            got = h.get_amplitudes().astype(np.complex128)
        pl = _P(); pl.ops, pl.tables, pl.n_phys = ops, tabs, nl
        init = swapped[r << nl:(r + 1) << nl].astype(cdt).astype(np.complex128)
        want, _ = em.run_plan(pl, n_global=s, rank=r, n_local=nl, psi0=init, active0=nl)
        worst = max(worst, np.abs(got - want).max())
    assert worst < (1e-12 if precision == 'double' else 3e-6)


@pytest.mark.parametrize('precision', ['double', 'single'])
@pytest.mark.parametrize('s', [1, 2, 3])
def test_gather_block_in_place_on_virtual_peers(precision, s, monkeypatch):
    """qcm_run_gather_block_inplace: the fused qubit swap + pass writing over each rank's own slabs, ordered by
    per-tile flags.  The 2^s 'ranks' are handles on this GPU, each on its own stream and launched from its own thread
    so that their kernels run concurrently and signal each other (a rank that never hears from a peer gives up after
    QCM_GATHER_SPIN_S seconds and raises).  Expected = numpy qubit swap + the op semantics per rank; run twice (epochs)."""
    import threading
    import torch
    monkeypatch.setenv('QCM_GATHER_SPIN_S', '3')
    monkeypatch.setenv('QCM_GATHER_K', '4')
    rng = np.random.RandomState(90 + s)
    nl = 16
    N = nl + s
    world = 1 << s
    cdt = np.complex128 if precision == 'double' else np.complex64
    e = fusion._Emitter()
    tq = list(range(nl - s, nl))

    def rand_u(m):
        q, _ = np.linalg.qr(rng.randn(1 << m, 2, 2) + 1j * rng.randn(1 << m, 2, 2))
        return q
    members = []
    for t in tq + [tq[0]]:
        ctrl = [int(c) for c in rng.permutation(nl - s)[:2]] + ([int(nl + rng.randint(s))] if rng.rand() < 0.6 else [])
        members.append((fusion.QCM_OP_MUX1Q, t, ctrl, fusion._mux_table_f64(rand_u(len(ctrl)))))
    if s == 1:
        k, t, c, tab = members[0]
        e.op(k, target=t, ctrl=c, n_in=nl, n_out=nl, table_off=e.table(tab))
    else:
        e.op(fusion.QCM_OP_BLOCK, target=s, ctrl=tq, n_in=nl, n_out=nl, n_ctrl=len(members))
        for k, t, c, tab in members:
            e.op(k, target=t, ctrl=c, n_in=nl, n_out=nl, table_off=e.table(tab))
    ops, tabs = e.finish()
    idx = np.arange(1 << N)
    lo_mask = (1 << (nl - s)) - 1

    def swap(psi):
        a = (idx >> (nl - s)) & (world - 1)
        b = idx >> nl
        return psi[(idx & lo_mask) | (b << (nl - s)) | (a << nl)]
    psi = (rng.randn(1 << N) + 1j * rng.randn(1 << N))
    psi = (psi / np.linalg.norm(psi)).astype(cdt).astype(np.complex128)
    states = [torch.from_numpy(np.ascontiguousarray(psi[r << nl:(r + 1) << nl].astype(cdt)).view(np.float64 if precision == 'double' else np.float32)).cuda()
              for r in range(world)]
    words = _native.gather_flag_words(nl, s, precision)
    flags = [torch.zeros(words, dtype=torch.int32, device='cuda') for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    handles = [_native.Handle(nl, precision, ext_state_ptr=states[r].data_ptr(), ext_stream=streams[r].cuda_stream) for r in range(world)]
    slab_bytes = (1 << (nl - s)) * np.dtype(cdt).itemsize
    for h in handles:
        # the handles' table buffers are allocated here: a cudaMalloc while another virtual rank's kernel is already
        # waiting for signals would hold this rank's launch back (one GPU only; real ranks own their GPU)
        h.run_program(ops[:0], tabs)
    torch.cuda.synchronize()
    pl = _P(); pl.ops, pl.tables, pl.n_phys = ops, tabs, nl
    try:
        for epoch in (1, 2):
            errs = [None] * world

            walls = [None] * world

            def work(r):
                import time as _t
                t0 = _t.perf_counter()
                try:
                    h = handles[r]
                    h.set_shard(s, r)
                    h.set_active(nl)
                    src = [states[j].data_ptr() + r * slab_bytes for j in range(world)]
                    h.run_gather_block_inplace(ops, tabs, src, [f.data_ptr() for f in flags], words, epoch)
                except Exception as ex:                       # noqa: BLE001
                    errs[r] = ex
                walls[r] = (round(t0, 3), round(_t.perf_counter(), 3))
            ths = [threading.Thread(target=work, args=(r,)) for r in range(world)]
            for th in ths:
                th.start()
            for th in ths:
                th.join()
            assert not any(errs), (errs, walls, [h.timing()['program_ms'] for h in handles])
            torch.cuda.synchronize()
            swapped = swap(psi)
            new = np.empty_like(psi)
            for r in range(world):
                want, _ = em.run_plan(pl, n_global=s, rank=r, n_local=nl, psi0=swapped[r << nl:(r + 1) << nl].copy(), active0=nl)
                got = handles[r].get_amplitudes().astype(np.complex128)
                assert np.abs(got - want).max() < (1e-12 if precision == 'double' else 3e-6) * (1 if epoch == 1 else 4), (epoch, r)
                new[r << nl:(r + 1) << nl] = got
            psi = new                                         # second epoch: the same pass again on the new state
    finally:
        for h in handles:
            h.close()


@pytest.mark.parametrize('s', [1, 2, 3])
def test_gather_tma_ring_is_bit_identical_to_plain_loads(s, monkeypatch):
    """The TMA-ring variant of the fused gather pass against the plain-load variant on a state large
    enough for several resident CTAs per SM and many ring wraps, repeated: bit-identical every time
    (guards the ring's stage release, see k_block_gather_tma)."""
    import torch
    nl = 24
    world = 1 << s
    rng = np.random.RandomState(7)
    old = [torch.randn(2 << nl, dtype=torch.float32, device='cuda') for _ in range(world)]
    out = torch.empty(2 << nl, dtype=torch.float32, device='cuda')
    e = fusion._Emitter()
    tq = list(range(nl - s, nl))
    q, _ = np.linalg.qr(rng.randn(4, 2, 2) + 1j * rng.randn(4, 2, 2))
    if s == 1:
        e.op(fusion.QCM_OP_MUX1Q, target=tq[0], ctrl=[3, 7], n_in=nl, n_out=nl, table_off=e.table(fusion._mux_table_f64(q)))
    else:
        e.op(fusion.QCM_OP_BLOCK, target=s, ctrl=tq, n_in=nl, n_out=nl, n_ctrl=s)
        for t in tq:
            e.op(fusion.QCM_OP_MUX1Q, target=t, ctrl=[3, 7], n_in=nl, n_out=nl, table_off=e.table(fusion._mux_table_f64(q)))
    ops, tabs = e.finish()
    src = [x.data_ptr() for x in old]
    with _native.Handle(nl, 'single', ext_state_ptr=old[0].data_ptr()) as h:
        h.set_shard(s, 0)
        h.set_active(nl)
        monkeypatch.setenv('QCM_GATHER', 'ldg')
        h.run_gather_block(ops, tabs, src, out.data_ptr())
        torch.cuda.synchronize()
        ref = out.clone()
        monkeypatch.setenv('QCM_GATHER', 'tma')
        for U, K in ((1, 16), (1, 64), (2, 16)):
            monkeypatch.setenv('QCM_GATHER_U', str(U))
            monkeypatch.setenv('QCM_GATHER_K', str(K))
            for _ in range(10):
                out.zero_()
                torch.cuda.synchronize()
                h.run_gather_block(ops, tabs, src, out.data_ptr())
                torch.cuda.synchronize()
                assert torch.equal(out, ref), (U, K)


@pytest.mark.parametrize('precision', ['double', 'single'])
@pytest.mark.parametrize('with_diag', [False, True])
@pytest.mark.parametrize('flags', [2, 3])
def test_rotated_expansion_is_transparent(precision, with_diag, flags):
    """QCM_FLAG_ROTATED_OUTPUT_OK: the last wide expansion pass stores its result with the new qubits as
    the low address bits (k_expand_low: one sequential write stream).  Every reader undoes the rotation:
    amplitudes, post-selection (prefix and general masks) and shots are those of the logical state."""
    rng = np.random.RandomState(91 + flags)
    tol = 1e-12 if precision == 'double' else 3e-6
    for n0, Ms in ((10, [8]), (11, [6]), (10, [2, 7]), (12, [5]), (13, [8])):
        ops, tabs, act = _expansion_program(rng, n0, Ms, with_diag, False)
        fusion._flag_last_pass(ops)
        ops['flags'][ops['flags'] != 0] = flags
        pl = _P(); pl.ops, pl.tables, pl.n_phys = ops, tabs, act
        want, _ = em.run_plan(pl)
        pw = np.abs(want) ** 2
        with _native.Handle(act, precision) as h:
            h.run_program(ops, tabs)
            got = h.get_amplitudes().astype(np.complex128)
            assert np.abs(got - want).max() < tol, (n0, Ms)
            part = h.get_amplitudes(5, 1000).astype(np.complex128)
            assert np.abs(part - want[5:1005]).max() < tol
            # prefix post-selection: every qubit >= nb on 0
            for nb in (n0, n0 - 3, act - Ms[-1]):
                mask = ((1 << act) - 1) & ~((1 << nb) - 1)
                p, kept = h.postselect(mask, 0, nb)
                assert np.abs(p - pw[:1 << nb]).max() < 10 * tol and abs(kept - pw[:1 << nb].sum()) < 10 * tol
            # general mask: a new qubit on 1, an old one on 0; marginal over the low 6 qubits
            mask, value = (1 << (act - 1)) | (1 << 7), 1 << (act - 1)
            p, kept = h.postselect(mask, value, 6)
            idx = np.arange(1 << act)
            sel = (idx & mask) == value
            wantp = np.bincount(idx[sel] & 63, weights=pw[sel], minlength=64)
            assert np.abs(p - wantp).max() < 10 * tol and abs(kept - pw[sel].sum()) < 10 * tol
            S = 200000
            k = h.sample(S, seed=3, stream_id=1).astype(np.int64)
            assert np.all(pw[k] > 0)
            emp = np.bincount(k, minlength=1 << act) / S
            assert 0.5 * np.abs(emp - pw / pw.sum()).sum() < weissman_tv_bound(int((pw > 0).sum()), S)
            M = Ms[-1]
            for shift, bits in ((act - M, M), (0, 8)):
                wm = np.bincount((idx >> shift) & ((1 << bits) - 1), weights=pw / pw.sum(), minlength=1 << bits)
                gm = np.bincount((k >> shift) & ((1 << bits) - 1), minlength=1 << bits) / S
                assert 0.5 * np.abs(gm - wm).sum() < weissman_tv_bound(1 << bits, S)
            # the next program starts from INIT: plain again
            plain = ops.copy(); plain['flags'] = 0
            h.run_program(plain, tabs)
            assert np.abs(h.get_amplitudes().astype(np.complex128) - want).max() < tol


def test_default_schedule_ends_in_the_rotated_expansion_kernel():
    """The default (lazily materialised) schedule of a QCMRF circuit ends in a wide expansion pass; with
    shots or without, that pass must be served by the rotated sequential-write kernel (k_expand_low), not
    by a silent fallback -- and the results are those of brute-force enumeration."""
    from qcmrf_b200 import workloads
    n = 11
    C = workloads.random_tree(n, 0, seed=3)                      # k = 10 ancillas -> passes of 2 and 8 new qubits
    th = workloads.theta_for(C, seed=4)
    pb, db, _ = mrf.brute_force_pmf(C, th)
    for precision, tol in (('single', 1e-5), ('double', 1e-10)):
        sim = B200Simulator(precision=precision, fusion='blocked', seed=5, small_batch=False)
        for shots in (0, 4096):
            res = sim.run(QCMRF(C, th), shots=shots).result()
            names = sim.op_kernels()
            assert names and names[-1].startswith('k_expand_low'), names
            p, d = res.postselected_probabilities(0)
            assert np.abs(p - pb).max() < tol and abs(d - db) < tol
            assert_pmf(p, pb, d, db, precision)
            if shots:
                assert sum(res.get_counts().values()) == shots
        sim.close()


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('n', [1, 5, 13, 20, 24])
def test_exact_mrf_inference_on_the_gpu(n):
    """px's `infer('partition')` / `logpot` / exact pmf (eval.py:84-93) by enumeration on the GPU
    against the oracle's brute force -- cliques of size 1..4, state ids with x_0 as the MSB."""
    from qcmrf_b200 import ExactMRF
    rng = np.random.RandomState(300 + n)
    C = []
    for v in range(n):                                          # every vertex appears; sizes 1..4
        m = int(rng.randint(1, min(4, n) + 1))
        others = [int(x) for x in rng.permutation([u for u in range(n) if u != v])[:m - 1]]
        c = [v] + others
        rng.shuffle(c)
        C.append([int(x) for x in c])
    th = -np.abs(rng.randn(sum(2 ** len(c) for c in C)))
    pb, db, _ = mrf.brute_force_pmf(C, th)
    ex = ExactMRF(C, th)
    p = ex.pmf()
    lz = ex.log_partition()
    assert np.abs(p - pb).max() < 1e-12 and (np.abs(p - pb) / pb).max() < 1e-9
    assert abs(ex.success_probability() - db) < 1e-12 * max(db, 1e-3) + 1e-15
    for xid in (0, (1 << n) - 1, int(rng.randint(0, 1 << n))):
        assert abs(np.exp(ex.logpot(xid) - lz) - pb[xid]) < 1e-12
    # the compat shim's px takes ln Z from the same kernel when a GPU is present
    from qcmrf_b200.compat.shim import kiopto_native as px
    b = px.backend(C, np.array([2] * n))
    px.weights(b)[:] = th
    assert abs(px.infer(b, task='partition') - lz) < 1e-12
    assert abs(px.logpot(b, 1) - ex.logpot(1)) < 1e-15


# ------------------------------------------------------------------------------------------
def test_pipelined_list_equals_blocking_single_calls(models):
    """run(list) on the large-state path is a pipeline (execute_deferred: program, shots and post-selection enqueued
    into page-locked ring buffers, circuit i collected while circuit i+1 runs; the tree total is summed on the
    device).  Counts, pmf and delta must be IDENTICAL to one blocking run() per circuit with the same Philox
    stream -- also across a change of state size in the middle of the list and over more circuits than ring slots."""
    from qcmrf_b200 import workloads
    items = []
    for j, t in ((1, 0), (1, 1), (1, 2), (3, 0), (3, 1), (1, 3), (5, 0), (5, 1), (5, 2), (5, 3)):
        items.append((models['0.5']['GRAPHS'][j], models['0.5']['THETAS'][str(j)][t]))
    C = workloads.random_tree(11, 0, seed=3)                    # ends in k_expand_low (rotated storage)
    items += [(C, workloads.theta_for(C, seed=s)) for s in (4, 5, 6, 7)]
    for precision in ('double', 'single'):
        sim = B200Simulator(precision=precision, small_batch=False, sweep_batch=False, seed=77)
        res = sim.run([QCMRF(c, th) for c, th in items], shots=5000).result()
        assert all(res.metadata(i)['path'] == 'statevector' for i in range(len(items)))
        one = B200Simulator(precision=precision, small_batch=False, sweep_batch=False, seed=77)
        for i, (c, th) in enumerate(items):
            r1 = one.run(QCMRF(c, th), shots=5000, stream_ids=[i]).result()
            assert r1.get_counts() == res.get_counts(i), (precision, i)
            p1, d1 = r1.postselected_probabilities(0)
            p, d = res.postselected_probabilities(i)
            assert np.array_equal(p, p1) and d == d1, (precision, i)
            pb, db, _ = mrf.brute_force_pmf(c, th)
            assert_pmf(p, pb, d, db, precision, (precision, i))
        # the callable returned by execute_deferred is single-use, and a fourth pending execution is refused
        pr = sim.prepare(QCMRF(*items[0]))
        fins = [sim.execute_deferred(pr, 100, seed=1, stream=s) for s in range(3)]
        with pytest.raises(RuntimeError):
            sim.execute_deferred(pr, 100, seed=1, stream=3)
        outs = [f() for f in fins]
        with pytest.raises(RuntimeError):
            fins[0]()
        k0, p0, m0 = sim.execute(pr, 100, seed=1, stream=0)
        assert np.array_equal(outs[0][0], k0) and np.array_equal(outs[0][1], p0) and outs[0][2] == m0
        assert not np.array_equal(outs[0][0], outs[1][0])
        sim.close(); one.close()
