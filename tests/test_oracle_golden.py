"""Pins the oracle against everything the reference ships for this path
(SURVEY.md 8c): the model files, the Aer histograms, the key conventions, and the
known answers of SURVEY.md App. F.  CPU only."""
import numpy as np
import pytest
from scipy import stats

from conftest import all_models
from oracle import cbridge, mrf, program, statevector as sv

# SURVEY.md App. F (brute force, fp64), rep 0 of each graph
KNOWN = {
    '0.1': dict(delta=[0.965353932039, 0.926776278731, 0.772781056829, 0.782162013777, 0.909321667715,
                       0.864688665774, 0.947717561020],
                p0=[0.516553234840, 0.259731685003, 0.068454935448, 0.028519950778, 0.128032998145,
                    0.029273076206, 0.062400999301],
                argmax=[0, 0, 2, 26, 5, 7, 3]),
    '0.5': dict(delta=[0.847558271933, 0.691408206118, 0.312395261422, 0.312432520698, 0.656054421907,
                       0.502842375882, 0.773571643742],
                p0=[0.582047855122, 0.299229665466, 0.086912382987, 0.018538352696, 0.133541840390,
                    0.021667494346, 0.061281833757],
                p1=[0.417952144878, 0.270044481711, 0.108062881579, 0.024249106964, 0.185883290971,
                    0.025115873924, 0.048782223277],
                argmax=[0, 0, 2, 26, 5, 7, 3]),
}


def test_theta_regeneration_is_bit_exact(models):
    """run_experiment.py:3,23-33 with d = sum 2^|C| reproduces models*.json exactly."""
    for scale in models:
        T = mrf.regenerate_thetas(float(scale))
        for j in range(7):
            for i in range(10):
                assert T[j][i] == models[scale]['THETAS'][str(j)][i]
        assert models[scale]['GRAPHS'] == mrf.GRAPHS


def test_known_answers(models):
    for scale, K in KNOWN.items():
        for j, C in enumerate(models[scale]['GRAPHS']):
            p, delta, _ = mrf.brute_force_pmf(C, models[scale]['THETAS'][str(j)][0])
            assert abs(delta - K['delta'][j]) < 5e-12
            assert abs(p[0] - K['p0'][j]) < 5e-12
            assert int(np.argmax(p)) == K['argmax'][j]
            if 'p1' in K:
                assert abs(p[1] - K['p1'][j]) < 5e-12


def test_statevector_matches_bruteforce_and_closed_form(models):
    worst = 0.0
    for scale, j, i, C, th in all_models(models):
        ops, N = program.qcmrf_program(C, th)
        psi, meas = sv.run_program(ops, N)
        n = program.sizes(C)[0]
        p, d = sv.postselected(psi, n)
        pb, db, _ = mrf.brute_force_pmf(C, th)
        worst = max(worst, np.abs(p - pb).max(), abs(d - db))
        assert np.abs(mrf.closed_form_state(C, th) - psi).max() < 1e-14
        # AND scratch qubit n carries no amplitude
        idx = np.arange(1 << N)
        assert np.abs(psi[((idx >> n) & 1) == 1]).max() == 0.0
    assert worst < 1e-13


def test_aer_histograms_are_consistent_with_oracle(models, aer_counts):
    """The only record of the third-party hot path: 210 unseeded Aer histograms.
    Statistical pin: support, success rate, TV and chi^2 p-values."""
    pvals, tvs = [], []
    for scale, j, i, C, th in all_models(models):
        ops, N = program.qcmrf_program(C, th)
        psi, meas = sv.run_program(ops, N)
        n = program.sizes(C)[0]
        kp = sv.key_probabilities(psi, N, meas)
        Q = aer_counts[scale][10 * j + i]
        assert sum(Q.values()) == 10000
        obs = np.zeros(1 << N)
        for k, v in Q.items():
            assert len(k) == N and k[N - 1 - n] == '0'          # clbit n is never written
            obs[int(k, 2)] = v
        assert obs[kp == 0].sum() == 0                            # support
        q, Z = mrf.postselect_counts(Q, n)
        _, delta = sv.postselected(psi, n)
        assert abs(Z / 1e4 - delta) < 0.02                        # success rate
        tvs.append(0.5 * np.abs(obs / 1e4 - kp).sum())
        m = kp > 0
        exp = 1e4 * kp[m]
        big = exp >= 5
        o = np.append(obs[m][big], obs[m][~big].sum())
        e = np.append(exp[big], exp[~big].sum())
        if e[-1] == 0:
            o, e = o[:-1], e[:-1]
        chi = ((o - e) ** 2 / e).sum()
        pvals.append(stats.chi2.sf(chi, len(e) - 1))
    assert max(tvs) < 0.09
    pvals = np.array(pvals)
    assert pvals.min() > 1e-5                                     # 210 draws: min ~ 1/210
    assert stats.kstest(pvals, 'uniform').pvalue > 1e-3          # p-values look uniform


def test_reference_helpers(models, aer_counts):
    C = models['0.5']['GRAPHS'][1]
    Q = aer_counts['0.5'][10]
    assert Q['0000'] == 2069 and Q['1011'] == 1166                # SURVEY.md section 4 worked example
    p1, s1 = mrf.extract_probs(Q, 2, 2)
    q2, Z = mrf.postselect_counts(Q, 2)
    assert np.allclose(p1, q2) and abs(s1 - Z / 1e4) < 1e-15
    pb, _, _ = mrf.brute_force_pmf(C, models['0.5']['THETAS']['1'][0])
    assert mrf.fidelity(pb, p1) > 0.999
    assert mrf.kl(pb, pb) == 0.0
    P, s = mrf.extract_probs({'1000': 5}, 2, 2)
    assert s == 0 and P.sum() == 0


def test_c_executor_matches_numpy(models):
    for scale, j, i, C, th in all_models(models):
        if i > 1:
            continue
        ops, N = program.qcmrf_program(C, th)
        psi, _ = sv.run_program(ops, N)
        arr, n_ops, _ = cbridge.compile_unfused(ops)
        assert np.abs(cbridge.run(N, arr, n_ops) - psi).max() < 1e-14
        arr, n_ops, tabs, N2 = cbridge.compile_fused(C, th)
        assert N2 == N and np.abs(cbridge.run(N, arr, n_ops, tabs) - psi).max() < 1e-14


def test_c_sampler_tv():
    C, th = [[0, 1], [1, 2]], [-0.3, -0.1, -0.7, -0.2, -0.5, -0.9, -0.05, -0.4]
    arr, n_ops, tabs, N = cbridge.compile_fused(C, th)
    psi = cbridge.run(N, arr, n_ops, tabs)
    idx = cbridge.sample(N, psi, 200000, 11)
    h = np.bincount(idx.astype(np.int64), minlength=1 << N) / 2e5
    assert 0.5 * np.abs(h - np.abs(psi) ** 2).sum() < 0.01
    assert np.all(np.abs(psi[np.unique(idx).astype(np.int64)]) > 0)


def test_gamma_theta_roundtrip():
    th = -np.abs(np.random.RandomState(0).randn(16))
    g = program.theta_to_gamma(th, 0.7)
    assert np.allclose(program.gamma_to_theta(g, 0.7), th)
    with np.errstate(invalid='ignore'):
        assert np.isnan(program.theta_to_gamma([0.5])[0])         # theta > 0 is unguarded (QCMRF.py:154)
