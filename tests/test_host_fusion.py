"""Host logic: lowering, fusion, planning.  Every emitted engine program is executed
by the numpy op-semantics emulator and compared with the oracle's gate-by-gate
statevector.  CPU only."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import engine_emulator as em
from conftest import all_models
from oracle import mrf, program, statevector as sv
from qcmrf_b200 import QCMRF, fusion, ir, transpile

CONFIGS = [('off', False, 1), ('clique', False, 1), ('clique', True, 1), ('clique', True, 2), ('clique', True, 4),
           ('clique', True, 5)]


def _logical(prog, mode, lazy, bm, elide=None):
    fc = fusion.fuse(prog, mode, use_hint=False)          # the numeric gate-fusion pass itself (QCMRF's shortcut: below)
    pl = fusion.plan(fc, lazy=lazy, block_max=bm, elide=elide)
    phys, act = em.run_plan(pl)
    return em.logical_state(pl, phys), fc, pl


def test_product_program_equals_oracle_program(models):
    """The product's QCMRF emits gate-for-gate what the oracle's restatement of
    QCMRF._build (QCMRF.py:199-243) emits."""
    for scale, j, i, C, th in all_models(models):
        if i > 2:
            continue
        for wm in (True, False):
            prog = ir.lower(QCMRF(C, th, with_measurements=wm))
            ops, N = program.qcmrf_program(C, th, with_measurements=wm)
            mine = [g for g in ir.to_oracle_ops(prog)]
            ref = []
            for g in ops:
                if g[0] == 'cp':
                    ref.append(('mcp', g[1], (g[2],), (1,), g[3]))
                elif g[0] == 'measure':
                    continue
                else:
                    ref.append(g)
            mine_g = [g for g in mine if g[0] != 'measure']
            assert len(mine_g) == len(ref)
            for a, b in zip(mine_g, ref):
                assert a[0] == b[0]
                if a[0] == 'mcp':
                    assert a[2:] == b[2:] and abs(a[1] - b[1]) < 1e-15
                else:
                    assert a == b
            assert prog.measures == {g[2]: g[1] for g in ops if g[0] == 'measure'}
            assert prog.n_qubits == N == prog.n_clbits


def test_fused_plans_match_oracle_on_all_fixture_models(models):
    worst = 0.0
    for scale, j, i, C, th in all_models(models):
        ops, N = program.qcmrf_program(C, th)
        psi, _ = sv.run_program(ops, N)
        prog = ir.lower(QCMRF(C, th))
        cfgs = CONFIGS if i == 0 else [('clique', True, 4)]
        for mode, lazy, bm in cfgs:
            lg, fc, pl = _logical(prog, mode, lazy, bm)
            worst = max(worst, np.abs(lg - psi).max())
            if mode == 'clique':
                # one multiplexer sweep per clique, on that clique's ancilla, scratch qubit gone
                n = program.sizes(C)[0]
                assert sorted(fc.init) == list(range(n))
                assert [o.kind for o in fc.ops] == ['mux'] * len(C)
                assert [o.target for o in fc.ops] == [n + 1 + ii for ii in range(len(C))]
                for o, cl in zip(fc.ops, C):
                    assert sorted(o.ctrls) == sorted(n - 1 - v for v in cl) and o.zero_in
                if lazy:
                    assert pl.n_phys == N - 1 and pl.layout[n] >= pl.n_phys
    assert worst < 1e-14


def test_fused_tables_are_the_closed_form_rx(models):
    """App. A: the block is RX(4 gamma) selected by the clique state."""
    for scale, j, i, C, th in all_models(models):
        if i:
            continue
        fc = fusion.fuse(ir.lower(QCMRF(C, th)), 'clique', use_hint=False)
        for op, (ctrl, c, s) in zip(fc.ops, program.rx_tables(C, th)):
            assert sorted(op.ctrls) == sorted(ctrl)
            perm = [ctrl.index(q) for q in op.ctrls]               # my bit j <-> oracle bit perm[j]
            for t in range(len(c)):
                to = sum(((t >> jj) & 1) << perm[jj] for jj in range(len(perm)))
                assert abs(op.table[t, 0, 0] - c[to]) < 1e-14
                assert abs(op.table[t, 1, 0] - (-1j * s[to])) < 1e-14


@pytest.mark.parametrize('graph', [0, 1, 2, 4, 5])
def test_transpiled_circuits_fuse_back(models, graph):
    """cx/id/rz/sx/x form (run_experiment.py:52) collapses to the same sweeps."""
    C = models['0.25']['GRAPHS'][graph]
    th = models['0.25']['THETAS'][str(graph)][3]
    ops, N = program.qcmrf_program(C, th)
    psi, _ = sv.run_program(ops, N)
    t = transpile(QCMRF(C, th), basis_gates=['cx', 'id', 'rz', 'sx', 'x'])
    assert set(t.count_ops()) <= {'cx', 'id', 'rz', 'sx', 'x', 'measure'}
    prog = ir.lower(t)
    # the translation itself is exact including the tracked global phase
    psi_t, _ = sv.run_program(ir.to_oracle_ops(prog), N)
    assert np.abs(psi_t - psi).max() < 1e-11
    for mode, lazy, bm in [('off', False, 1), ('clique', True, 4)]:
        lg, fc, pl = _logical(prog, mode, lazy, bm)
        assert np.abs(lg - psi).max() < 1e-11
    assert sum(1 for o in fc.ops if o.kind == 'mux' and o.zero_in) == len(C)
    assert len(fc.ops) <= len(C) + 2


def test_gamma_zero_terms_and_beta(models):
    C = [[0, 1], [1, 2]]
    th = [0.0, -0.3, 0.0, -1.2, -0.4, 0.0, -0.1, -2.0]           # theta = 0 -> gamma = 0 -> term skipped
    for beta in (1.0, 0.25, 3.0):
        ops, N = program.qcmrf_program(C, th, beta=beta)
        psi, _ = sv.run_program(ops, N)
        prog = ir.lower(QCMRF(C, th, beta=beta))
        assert len(prog.gates) == len([g for g in ops if g[0] not in ('measure', 'barrier')])
        lg, fc, pl = _logical(prog, 'clique', True, 4)
        assert np.abs(lg - psi).max() < 1e-14
        p, d = sv.postselected(lg, 3)
        pb, db, _ = mrf.brute_force_pmf(C, th, beta)
        assert np.abs(p - pb).max() < 1e-13 and abs(d - db) < 1e-13


def test_gamma_parametrisation_and_no_measurements():
    C = [[0], [0, 1, 2]]
    g = list(np.linspace(0.05, 0.7, 10))
    ops, N = program.qcmrf_program(C, gamma=g, with_measurements=False)
    psi, _ = sv.run_program(ops, N)
    circ = QCMRF(C, gamma=g, with_measurements=False, with_barriers=True)
    prog = ir.lower(circ)
    assert prog.measures == {}
    lg, _, _ = _logical(prog, 'clique', True, 3)
    assert np.abs(lg - psi).max() < 1e-14
    assert np.allclose(circ.theta, program.gamma_to_theta(g))


def test_no_elision_keeps_full_width():
    C, th = [[0, 1], [1, 2]], [-0.3, -0.1, -0.7, -0.2, -0.5, -0.9, -0.05, -0.4]
    prog = ir.lower(QCMRF(C, th))
    psi, _ = sv.run_program(program.qcmrf_program(C, th)[0], 6)
    lg, fc, pl = _logical(prog, 'clique', True, 4, elide=False)
    assert pl.n_phys == 6 and pl.final_active == 6
    assert np.abs(lg - psi).max() < 1e-14


@st.composite
def clique_sets(draw):
    n = draw(st.integers(1, 6))
    k = draw(st.integers(1, 4))
    cliques = []
    for _ in range(k):
        m = draw(st.integers(1, min(4, n)))
        cliques.append(draw(st.permutations(list(range(n))))[:m])
    cliques[0] = sorted(set(cliques[0]) | {n - 1})[:4] if n - 1 not in cliques[0] and len(cliques[0]) < 4 else cliques[0]
    if not any(n - 1 in c for c in cliques):
        cliques.append([n - 1])
    return [list(map(int, c)) for c in cliques]


@settings(max_examples=25, deadline=None)
@given(clique_sets(), st.integers(0, 2 ** 31 - 1), st.sampled_from([1, 2, 4, 5]))
def test_random_clique_sets(cliques, seed, bm):
    """T6: random clique sets (|C| 1..4, unordered vertex lists, repeated cliques): fused vs
    unfused vs brute force."""
    rng = np.random.RandomState(seed)
    dim = sum(2 ** len(c) for c in cliques)
    th = -np.abs(rng.randn(dim)) * 0.7
    n, k, N, _ = program.sizes(cliques)
    if N > 11:
        return
    ops, _ = program.qcmrf_program(cliques, th)
    psi, _ = sv.run_program(ops, N)
    prog = ir.lower(QCMRF(cliques, list(th)))
    lg, fc, pl = _logical(prog, 'clique', True, bm)
    assert np.abs(lg - psi).max() < 1e-13
    lg2, _, _ = _logical(prog, 'off', False, 1)
    assert np.abs(lg2 - psi).max() < 1e-13
    p, d = sv.postselected(lg, n)
    pb, db, _ = mrf.brute_force_pmf(cliques, th)
    assert np.abs(p - pb).max() < 1e-12 and abs(d - db) < 1e-12


def test_foreign_circuit_generic_gates():
    """Circuits that are not QCMRF: every primitive still runs (as its own sweep when no
    block structure is found)."""
    from qcmrf_b200 import QuantumCircuit
    qc = QuantumCircuit(4, 4)
    qc.h(0); qc.cx(0, 1); qc.ry(0.3, 2); qc.cz(1, 2); qc.t(0); qc.swap(0, 3); qc.rx(1.1, 1)
    qc.mcx([0, 1], 2); qc.cp(0.7, 3, 0); qc.sdg(3); qc.u(0.1, 0.2, 0.3, 2); qc.crz(0.4, 2, 3)
    qc.measure_all()
    prog = ir.lower(qc)
    psi, meas = sv.run_program([g for g in _oracle_ops(prog)], 4)
    for mode, lazy, bm in CONFIGS:
        lg, fc, pl = _logical(prog, mode, lazy, bm)
        assert np.abs(lg - psi).max() < 1e-14, (mode, lazy, bm)


def _oracle_ops(prog):
    """ir.to_oracle_ops lacks a few gate names the oracle executor has no tuple for:
    expand them through their matrices."""
    out = []
    for g in prog.gates:
        if g.name == 'u':
            # u(theta, phi, lam) = rz(phi) ry(theta) rz(lam) up to phase e^{i(phi+lam)/2}
            th, ph, lm = g.params
            out += [('rz', lm, g.target), ('ry', th, g.target), ('rz', ph, g.target), ('gphase', (ph + lm) / 2)]
        elif g.name == 'crz':
            c, t = g.qubits
            out += [('rz', g.params[0] / 2, t), ('mcx', (c,), (1,), t), ('rz', -g.params[0] / 2, t), ('mcx', (c,), (1,), t)]
        else:
            out += ir.to_oracle_ops(ir.Program(prog.n_qubits, prog.n_clbits, [g]))
    return out


def test_plan_limits_are_respected(models):
    C = models['0.1']['GRAPHS'][3]
    th = models['0.1']['THETAS']['3'][0]
    fc = fusion.fuse(ir.lower(QCMRF(C, th)), 'clique')
    for bm in (1, 2, 3, 4, 5):
        pl = fusion.plan(fc, lazy=True, block_max=bm, expand_max=bm)
        for op in pl.ops:
            if op['kind'] == fusion.QCM_OP_BLOCK:
                assert op['target'] <= bm and op['n_ctrl'] <= fusion.QCM_MAX_MEMBERS
        assert pl.n_passes == 1 + -(-len(C) // bm)
        wide = fusion.plan(fc, lazy=True, block_max=bm)              # expansion passes may be wider
        for op in wide.ops:
            if op['kind'] == fusion.QCM_OP_BLOCK:
                assert op['target'] <= fusion.QCM_MAX_EXPAND and op['n_ctrl'] <= fusion.QCM_MAX_MEMBERS
        assert wide.n_passes == 2                                   # init + one pass for all 4 ancillas
    # remainder-first sizing: 18 expansion sweeps -> passes of 2, 8, 8 targets
    from qcmrf_b200 import workloads
    C37, _ = workloads.named('q37')
    pl = fusion.plan(fusion.fuse(ir.lower(QCMRF(C37, workloads.theta_for(C37))), 'clique'), lazy=True, block_max=4)
    assert [int(op['target']) for op in pl.ops if op['kind'] == fusion.QCM_OP_BLOCK] == [2, 8, 8]


def test_release_mode_on_the_emulator(monkeypatch, models):
    """width='release': clique ancillas are never stored; pmf/delta by projection, shots from the
    sweep coefficients.  Host logic on the numpy engine emulator vs brute force and the oracle."""
    import fake_native
    from oracle import mrf, program, statevector as sv
    from qcmrf_b200 import QCMRF, B200Simulator
    fake_native.install(monkeypatch)
    for j in (2, 3, 5):
        C = models['0.5']['GRAPHS'][j]
        th = models['0.5']['THETAS'][str(j)][4]
        n, k, N, _ = program.sizes(C)
        pb, db, _ = mrf.brute_force_pmf(C, th)
        sim = B200Simulator(precision='double', width='release', small_batch=False, seed=3)
        pr = sim.prepare(QCMRF(C, th))
        assert pr.plan.n_phys == n and len(pr.virtual) == k           # only the variables are stored
        res = sim.run(QCMRF(C, th), shots=30000).result()
        p, delta = res.postselected_probabilities()
        assert np.abs(p - pb).max() < 1e-12 and abs(delta - db) < 1e-12
        assert res.metadata(0)['width'] == 'release'
        psi, meas = sv.run_program(program.qcmrf_program(C, th)[0], N)
        kp = sv.key_probabilities(psi, N, meas)
        obs = np.zeros(1 << N)
        for key, v in res.get_counts().items():
            assert len(key) == N
            obs[int(key, 2)] = v
        assert obs.sum() == 30000 and obs[kp < 1e-15].sum() == 0
        assert 0.5 * np.abs(obs / 3e4 - kp).sum() < 0.5 * np.sqrt(2 * ((kp > 0).sum() * np.log(2) + np.log(1e6)) / 3e4)


def test_qcmrf_fused_shortcut_equals_the_fusion_pass(models):
    """QCMRF._fused_circuit (the closed form of SURVEY.md App. A, used by the backend instead of fusing
    ~12 gates per clique numerically) is exactly what fusion.fuse makes of the emitted gate list: same
    sweeps, same index-qubit order, tables equal to rounding; skipped terms (gamma ~ 0, QCMRF.py:223) are
    identities; an edited circuit falls back to the gate list."""
    rng = np.random.RandomState(12)
    cases = [(C, th) for _s, _j, i, C, th in all_models(models) if i < 2]
    for _ in range(12):
        n = int(rng.randint(2, 7))
        C = []
        for _k in range(int(rng.randint(1, 5))):
            m = int(rng.randint(1, min(4, n) + 1))
            C.append([int(v) for v in rng.permutation(n)[:m]])
        C.append([n - 1])
        th = -np.abs(rng.randn(sum(2 ** len(c) for c in C)))
        th[rng.rand(len(th)) < 0.2] = 0.0                      # gamma == 0: the constructor skips the term
        cases.append((C, list(th)))
    for C, th in cases:
        for beta in (1.0, 0.5):
            prog = ir.lower(QCMRF(C, th, beta=beta))
            fast = fusion.fuse(prog, 'clique')
            slow = fusion.fuse(prog, 'clique', use_hint=False)
            if prog.fused_hint() is None:                      # an all-skipped clique: no shortcut
                continue
            assert fast.n_qubits == slow.n_qubits and sorted(fast.init) == sorted(slow.init)
            for q in fast.init:
                assert np.abs(fast.init[q] - slow.init[q]).max() < 1e-15
            assert abs(fast.global_phase - slow.global_phase) < 1e-12
            assert len(fast.ops) == len(slow.ops)
            for a, b in zip(fast.ops, slow.ops):
                assert (a.kind, a.target, tuple(a.ctrls), a.zero_in) == (b.kind, b.target, tuple(b.ctrls), b.zero_in)
                assert np.abs(a.table - b.table).max() < 1e-14
    # editing the circuit materialises the instruction list: no shortcut any more
    c = QCMRF([[0, 1]], [-0.1, -0.2, -0.3, -0.4], with_measurements=False)
    c.x(0)
    prog = ir.lower(c)
    assert getattr(prog, 'fused_hint', None) is None or prog.fused_hint() is None


def test_plan_cache_refreshes_tables_only(models):
    """qcmrf_b200/plancache.py: a later circuit with the structure of an earlier one gets the cached plan
    with gathered tables -- bit-identical to planning it in full, for every schedule; a structure whose
    planning does arithmetic on the coefficients is recognised and never served from the cache."""
    from qcmrf_b200 import plancache, backend
    rng = np.random.RandomState(3)
    graphs = [C for _s, _j, i, C, _th in all_models(models) if i == 0][:7] + [[[0, 1], [1, 2], [2, 3], [3, 4], [0, 4]]]
    for C in graphs:
        dim = sum(2 ** len(c) for c in C)
        for lazy, bm in ((True, 4), (True, 1), (False, 1)):
            pc = plancache.PlanCache()

            def build(f):
                p = fusion.plan(f, lazy=lazy, block_max=bm)
                return p, [p.tables]
            for rep in range(4):
                th = -np.abs(rng.randn(dim))
                if rep == 2:
                    th[rng.randint(dim)] = 0.0
                fc = fusion.fuse(ir.lower(QCMRF(C, list(th), beta=1.0 - 0.2 * rep)), 'clique')
                got = pc.get(fc, (lazy, bm), build, backend._clone_plan)
                want = fusion.plan(fc, lazy=lazy, block_max=bm)
                assert np.array_equal(got.ops, want.ops) and np.array_equal(got.tables, want.tables)
                assert got.layout == want.layout and got.n_phys == want.n_phys and got.global_phase == want.global_phase
            assert pc.misses == 1 and pc.hits == 3
    # a planner that multiplies coefficients (here: squares the tables) must be found out
    fc = fusion.fuse(ir.lower(QCMRF([[0, 1]], [-0.1, -0.2, -0.3, -0.4])), 'clique')
    pc = plancache.PlanCache()

    def bad_build(f):
        p = fusion.plan(f, lazy=True, block_max=4)
        return p, [p.tables * p.tables]
    pc.get(fc, (), bad_build, backend._clone_plan)
    pc.get(fc, (), bad_build, backend._clone_plan)
    assert pc.hits == 0 and pc.uncacheable == 1


def test_native_block_apply_equals_numpy(monkeypatch):
    """csrc/qcm_host.c::qcm_block_apply (the fusion pass's row mixer) == the numpy version, gate by gate:
    random controlled / uncontrolled gates incl. diagonal and X-type fast paths, qubits joining mid-way."""
    from qcmrf_b200 import build
    from qcmrf_b200.ir import Gate
    build.build_host()
    assert fusion._load_host_apply() is not None
    rng = np.random.RandomState(8)
    names1 = ['h', 'x', 'sx', 't', 'z']
    for trial in range(20):
        nq = int(rng.randint(2, 7))
        zero = set(int(q) for q in rng.permutation(nq)[:int(rng.randint(0, nq))])
        gates = []
        for _ in range(60):
            r = rng.rand()
            q = [int(x) for x in rng.permutation(nq)]
            if r < 0.3:
                gates.append(Gate(names1[rng.randint(len(names1))], (q[0],)))
            elif r < 0.5:
                gates.append(Gate('rz', (q[0],), (float(rng.uniform(-3, 3)),)))
            elif r < 0.7:
                gates.append(Gate('cx', (q[0], q[1]), (), (1,)))
            elif r < 0.85 and nq >= 3:
                gates.append(Gate('mcx', (q[0], q[1], q[2]), (), (int(rng.randint(2)), int(rng.randint(2)))))
            else:
                gates.append(Gate('cp', (q[0], q[1]), (float(rng.uniform(-3, 3)),), (1,)))
        a = fusion._Block(zero)
        for g in gates:
            a.apply(g)
        monkeypatch.setattr(fusion, '_HOST_APPLY', None)
        b = fusion._Block(zero)
        for g in gates:
            b.apply(g)
        monkeypatch.undo()
        assert a.qubits == b.qubits and a.U.shape == b.U.shape
        assert np.abs(a.U - b.U).max() < 1e-13


def test_transpiled_circuits_fuse_to_the_same_program_structure(models):
    """The reference submits TRANSPILED circuits (run_experiment.py:52, basis cx/id/rz/sx/x).  Their fused
    form must have the structure of the untranspiled circuit's: the transpiled H layer folds into the
    initial product state completely, every clique block is ONE multiplexer on its still-|0> ancilla --
    so the planner emits the same passes (tables differ by per-qubit phases only)."""
    from qcmrf_b200 import workloads
    cases = [(C, models['0.5']['THETAS'][str(j)][0]) for j, C in enumerate(models['0.5']['GRAPHS'])]
    C16 = workloads.random_tree(8, 0, seed=2)                    # 16 qubits: large-state (lazy, blocked) planning
    cases.append((C16, workloads.theta_for(C16, seed=2)))
    for C, th in cases:
        plain = fusion.fuse(ir.lower(QCMRF(C, th)), 'clique', use_hint=False)
        T = transpile([QCMRF(C, th)], basis_gates=['cx', 'id', 'rz', 'sx', 'x'])[0]
        fused = fusion.fuse(ir.lower(T), 'clique')
        assert sorted(fused.init) == sorted(plain.init)
        assert [(o.kind, o.target, tuple(sorted(o.ctrls)), o.zero_in) for o in fused.ops] == \
               [(o.kind, o.target, tuple(sorted(o.ctrls)), o.zero_in) for o in plain.ops]
        pa = fusion.plan(plain, lazy=True, block_max=4)
        pb = fusion.plan(fused, lazy=True, block_max=4)
        assert pa.n_phys == pb.n_phys and pa.n_passes == pb.n_passes and pa.layout == pb.layout
        assert [int(k) for k in pa.ops['kind']] == [int(k) for k in pb.ops['kind']]


def test_merge_diagonals_equals_brute_force():
    """fusion.merge_diagonals (gather structure cached per index-qubit list): the product of the merged
    tables over any basis state equals the product of the original factors, also when the union of index
    qubits overflows QCM_MAX_CTRL and a second table is started; a repeated call with other values reuses
    the cached structure."""
    rng = np.random.RandomState(21)
    nq = 14
    for trial in range(6):
        ctrls = [tuple(int(c) for c in rng.permutation(nq)[:int(rng.randint(1, 4))]) for _ in range(int(rng.randint(3, 12)))]
        for rep in range(2):
            members = [(c, np.exp(1j * rng.uniform(0, 6, 1 << len(c)))) for c in ctrls]
            merged = fusion.merge_diagonals(members)
            assert all(len(c) <= fusion.QCM_MAX_CTRL for c, _t in merged)
            x = rng.randint(0, 1 << nq, size=200)

            def value(ctrl, tab):
                idx = np.zeros_like(x)
                for j, c in enumerate(ctrl):
                    idx |= ((x >> c) & 1) << j
                return np.asarray(tab)[idx]
            want = np.prod([value(c, t) for c, t in members], axis=0)
            got = np.prod([value(c, t) for c, t in merged], axis=0)
            assert np.abs(got - want).max() < 1e-12


# ---- regression tests for the round-1 advisor findings ---------------------------------------------
def _gate_by_gate(circ):
    prog = ir.lower(circ)
    psi, _ = sv.run_program(ir.to_oracle_ops(prog), prog.n_qubits)
    return prog, psi


def test_ghz_ending_in_cx_onto_clean_qubits():
    """fold_clean_scratch looked one gate past the end when the second-to-last gate was a cx onto a
    clean qubit (IndexError on the 3-qubit GHZ)."""
    from qcmrf_b200.circuit import QuantumCircuit
    c = QuantumCircuit(3, 3)
    c.h(0); c.cx(0, 1); c.cx(0, 2)
    c.measure(range(3), range(3))
    prog, psi = _gate_by_gate(c)
    for mode, lazy, bm in CONFIGS:
        lg, _, _ = _logical(prog, mode, lazy, bm)
        assert np.abs(lg - psi).max() < 1e-14, (mode, lazy, bm)


def test_controls_on_untouched_zero_qubits():
    """A control on a qubit that is still |0> never fires (closed) / always fires (open): the sweep must not
    be indexed by a qubit the lazy layout never materialises (`cx(2,1); cx(0,3)` on |0000>)."""
    from qcmrf_b200.circuit import QuantumCircuit
    c = QuantumCircuit(4)
    c.cx(2, 1); c.cx(0, 3)
    prog, psi = _gate_by_gate(c)
    for mode, lazy, bm in CONFIGS:
        lg, _, _ = _logical(prog, mode, lazy, bm)
        assert np.abs(lg - psi).max() < 1e-14
    c = QuantumCircuit(4)
    c.h(1)
    c.mcx([0, 1], 2, ctrl_state='10')          # qubit 0 open (always fires), qubit 1 closed
    c.mcx([3, 1], 0)                           # qubit 3 still |0>: identity
    c.cx(2, 3)
    prog, psi = _gate_by_gate(c)
    pruned = fusion.prune_zero_controls(prog.gates, 4)
    assert [g.qubits for g in pruned] == [(1,), (1, 2), (2, 3)]
    for mode, lazy, bm in CONFIGS:
        lg, _, _ = _logical(prog, mode, lazy, bm)
        assert np.abs(lg - psi).max() < 1e-14


def test_release_keeps_sweeps_whose_index_qubits_change_later():
    """width='release' may drop a sweep only if its index qubits keep their value to the end
    (`h(0); cx(0,1); h(0)`: the cx must stay in the core program)."""
    from qcmrf_b200.circuit import QuantumCircuit
    c = QuantumCircuit(2)
    c.h(0); c.cx(0, 1); c.h(0)
    fc = fusion.fuse(ir.lower(c), 'clique')
    core, virtual = fusion.split_releasable(fc, keep_below=1)
    assert not virtual and len(core.ops) == len(fc.ops)
    c = QuantumCircuit(2)
    c.h(0); c.cx(0, 1)
    core, virtual = fusion.split_releasable(fusion.fuse(ir.lower(c), 'clique'), keep_below=1)
    assert len(virtual) == 1 and virtual[0].target == 1


def test_gates_after_a_measurement_are_rejected():
    from qcmrf_b200.circuit import QuantumCircuit
    c = QuantumCircuit(2, 2)
    c.measure(0, 0); c.x(0)
    with pytest.raises(ValueError, match='follows a measurement'):
        ir.lower(c)
    c = QuantumCircuit(2, 2)
    c.h(0); c.measure(0, 0); c.cx(0, 1)                  # measured qubit as a control
    with pytest.raises(ValueError, match='follows a measurement'):
        ir.lower(c)
    c = QuantumCircuit(2, 2)
    c.h(0); c.measure(0, 0); c.h(1); c.measure(1, 1)     # QCMRF's pattern: measured qubits left alone
    assert ir.lower(c).measures == {0: 0, 1: 1}
    with pytest.raises(ValueError, match='follows a measurement'):
        transpile(_edit(c))


def _edit(c):
    c = c.copy()
    c.x(0)
    return c


def test_positive_theta_is_an_error_on_both_paths():
    """theta > 0 has no circuit angle (gamma = NaN, QCMRF.py:154): never a silently plausible pmf."""
    q = QCMRF([[0]], [0.5, -0.3])
    with pytest.raises(ValueError, match='theta must be <= 0'):
        fusion.fuse(ir.lower(q), 'clique')
    with pytest.raises(ValueError, match='theta must be <= 0'):
        fusion.fuse(ir.lower(q), 'off')
    q = QCMRF([[0]], [0.5, -0.3])
    q.data                                               # materialise the instruction list: generic walk
    with pytest.raises(ValueError, match='theta must be <= 0'):
        ir.lower(q)


def test_basis_array_fusion_equals_the_object_loop(models):
    """transpile() hands the engine flat gate arrays (ir.BasisProgram); fusion._fuse_basis walks them in C.  It must
    produce exactly what the object-per-gate loop produces from the same gates: same sweeps, same index-qubit order,
    tables to rounding (the C side accumulates rz phases as angles) -- on every fixture graph and on random circuits."""
    from qcmrf_b200.transpile import BasisCircuit
    if fusion._HOST_BASIS is None:
        pytest.skip('_qcm_host.so not built')
    cases = [(C, models['0.5']['THETAS'][str(j)][i]) for j, C in enumerate(models['0.5']['GRAPHS']) for i in (0, 7)]
    cases.append(([[0, 1], [1, 2]], [0.0, -0.3, 0.0, -1.0, -0.2, 0.0, -0.5, -0.1]))      # skipped (gamma ~ 0) terms
    for C, th in cases:
        T = transpile(QCMRF(C, th))
        assert isinstance(T, BasisCircuit) and T.__dict__['_bc_data'] is None
        prog = ir.lower(T)
        assert isinstance(prog, ir.BasisProgram) and prog.metadata['num_vertices'] == max(max(c) for c in C) + 1
        fast = fusion.fuse(prog, 'clique')
        slow = fusion.fuse(prog, 'clique', use_hint=False)
        assert sorted(fast.init) == sorted(slow.init) and len(fast.ops) == len(slow.ops)
        assert fast.n_gates_in == slow.n_gates_in == len(prog.bk)
        for q in fast.init:
            assert np.abs(fast.init[q] - slow.init[q]).max() < 1e-12
        for a, b in zip(fast.ops, slow.ops):
            assert (a.kind, a.target, tuple(a.ctrls), a.zero_in) == (b.kind, b.target, tuple(b.ctrls), b.zero_in)
            assert np.abs(a.table - b.table).max() < 1e-12
        assert abs(fast.global_phase - slow.global_phase) < 1e-12
        # the lazily materialised instruction list spells the same gates; touching it switches to the generic walk
        names = {0: 'rz', 1: 'sx', 2: 'x', 3: 'id', 4: 'cx'}
        data = [i for i in T.data if i.operation.name != 'measure']
        assert [i.operation.name for i in data] == [names[int(k)] for k in prog.bk]
        assert [i.qubits[-1] for i in data] == [int(q) for q in prog.bq]
        assert T._lower_program() is None
        walked = ir.lower(T)
        assert not hasattr(walked, 'bk') and len(walked.gates) == len(prog.bk) and walked.measures == prog.measures
    # random basis-gate circuits (not QCMRF-shaped): cx onto clean qubits, long 1-qubit runs, every qubit used
    rng = np.random.RandomState(4)
    for trial in range(25):
        n = int(rng.randint(2, 7))
        ng = int(rng.randint(5, 60))
        bk = rng.choice([0, 0, 1, 2, 3, 4, 4], size=ng).astype(np.int8)
        bq = rng.randint(0, n, size=ng).astype(np.int32)
        bc = np.full(ng, -1, dtype=np.int32)
        for g in range(ng):
            if bk[g] == 4:
                bc[g] = int(rng.choice([q for q in range(n) if q != bq[g]]))
        bp = rng.uniform(-3, 3, size=ng)
        prog = ir.BasisProgram(n, n, bk, bq, bc, bp)
        fast = fusion.fuse(prog, 'clique')
        slow = fusion.fuse(prog, 'clique', use_hint=False)
        pl_f, pl_s = fusion.plan(fast, lazy=True, block_max=4), fusion.plan(slow, lazy=True, block_max=4)
        sf = em.logical_state(pl_f, em.run_plan(pl_f)[0])
        ss = em.logical_state(pl_s, em.run_plan(pl_s)[0])
        psi, _ = sv.run_program(ir.to_oracle_ops(prog), n)
        assert np.abs(sf - psi).max() < 1e-12 and np.abs(ss - psi).max() < 1e-12, trial


def _random_circuit(rng, nq, n_gates):
    from qcmrf_b200.circuit import QuantumCircuit
    c = QuantumCircuit(nq)
    one = ['h', 'x', 'y', 'z', 's', 'sdg', 't', 'tdg', 'sx', 'sxdg', 'id']
    for _ in range(n_gates):
        r = rng.rand()
        q = [int(v) for v in rng.permutation(nq)]
        if r < 0.3:
            getattr(c, one[rng.randint(len(one))])(q[0])
        elif r < 0.45:
            getattr(c, ['rz', 'rx', 'ry', 'p'][rng.randint(4)])(float(rng.uniform(-3, 3)), q[0])
        elif r < 0.5:
            c.u(*[float(v) for v in rng.uniform(-3, 3, 3)], q[0])
        elif r < 0.7 and nq >= 2:
            getattr(c, ['cx', 'cz', 'swap'][rng.randint(3)])(q[0], q[1])
        elif r < 0.8 and nq >= 2:
            getattr(c, ['cp', 'crz'][rng.randint(2)])(float(rng.uniform(-3, 3)), q[0], q[1])
        elif r < 0.9 and nq >= 3:
            m = int(rng.randint(2, min(nq, 4)))
            c.mcx(q[:m], q[m], ctrl_state=''.join(rng.choice(['0', '1'], m)))
        elif nq >= 3:
            m = int(rng.randint(2, min(nq, 4)))
            c.mcp(float(rng.uniform(-3, 3)), q[:m], q[m])
        else:
            c.h(q[0])
    return c


def _textbook_state(circ):
    """Dense reference for the fuzz test below, independent of the product's gate tables: textbook 2x2 matrices
    (Qiskit's conventions) applied gate by gate to |0...0>, little-endian."""
    sq, i = np.sqrt(0.5), 1j
    fixed = {'h': [[sq, sq], [sq, -sq]], 'x': [[0, 1], [1, 0]], 'y': [[0, -i], [i, 0]], 'z': [[1, 0], [0, -1]],
             's': [[1, 0], [0, i]], 'sdg': [[1, 0], [0, -i]], 't': [[1, 0], [0, np.exp(i * np.pi / 4)]],
             'tdg': [[1, 0], [0, np.exp(-i * np.pi / 4)]], 'sx': [[(1 + i) / 2, (1 - i) / 2], [(1 - i) / 2, (1 + i) / 2]],
             'sxdg': [[(1 - i) / 2, (1 + i) / 2], [(1 + i) / 2, (1 - i) / 2]], 'id': [[1, 0], [0, 1]]}

    def mat(name, p):
        if name in fixed:
            return np.array(fixed[name], dtype=complex)
        if name in ('rz', 'crz'):
            return np.diag([np.exp(-i * p[0] / 2), np.exp(i * p[0] / 2)])
        if name == 'rx':
            c, s_ = np.cos(p[0] / 2), np.sin(p[0] / 2)
            return np.array([[c, -i * s_], [-i * s_, c]])
        if name == 'ry':
            c, s_ = np.cos(p[0] / 2), np.sin(p[0] / 2)
            return np.array([[c, -s_], [s_, c]], dtype=complex)
        if name in ('p', 'cp', 'mcp'):
            return np.diag([1, np.exp(i * p[0])])
        if name == 'u':
            c, s_ = np.cos(p[0] / 2), np.sin(p[0] / 2)
            return np.array([[c, -np.exp(i * p[2]) * s_], [np.exp(i * p[1]) * s_, np.exp(i * (p[1] + p[2])) * c]])
        if name in ('cx', 'mcx'):
            return np.array(fixed['x'], dtype=complex)
        if name == 'cz':
            return np.array(fixed['z'], dtype=complex)
        raise ValueError(name)

    n = circ.num_qubits
    psi = np.zeros(1 << n, dtype=complex)
    psi[0] = 1.0
    idx = np.arange(1 << n)

    def apply(u, ctrls, vals, t):
        sel = np.ones(1 << n, dtype=bool)
        for c, v in zip(ctrls, vals):
            sel &= ((idx >> c) & 1) == v
        lo = idx[sel & (((idx >> t) & 1) == 0)]
        hi = lo | (1 << t)
        a, b = psi[lo].copy(), psi[hi].copy()
        psi[lo] = u[0, 0] * a + u[0, 1] * b
        psi[hi] = u[1, 0] * a + u[1, 1] * b

    for ins in circ.data:
        op, qs = ins.operation, tuple(int(q) for q in ins.qubits)
        if op.name == 'swap':
            for c, t in ((qs[0], qs[1]), (qs[1], qs[0]), (qs[0], qs[1])):
                apply(mat('x', ()), (c,), (1,), t)
            continue
        vals = tuple(op.ctrl_values) if getattr(op, 'ctrl_values', None) is not None else (1,) * (len(qs) - 1)
        apply(mat(op.name, op.params), qs[:-1], vals, qs[-1])
    return psi


@pytest.mark.parametrize('seed', range(6))
def test_random_generic_circuits_through_every_fusion_mode(seed):
    """Seeded fuzz of the generic front end: random circuits over the whole gate surface -- many of them
    start with controls on untouched |0> qubits, end in gates onto clean qubits, or leave qubits untouched -- lowered,
    fused, planned (eager and lazy, blocks of 1..5) and run on the engine emulator must give the gate-by-gate
    state of an independent textbook simulation exactly (no global-phase slack), and the transpiled circuit the same state
    up to a global phase."""
    rng = np.random.RandomState(4000 + seed)
    for trial in range(25):
        nq = int(rng.randint(1, 7))
        c = _random_circuit(rng, nq, int(rng.randint(1, 28)))
        prog, psi = ir.lower(c), _textbook_state(c)
        for mode, lazy, bm in CONFIGS:
            lg, _, _ = _logical(prog, mode, lazy, bm)
            assert np.abs(lg - psi).max() < 1e-12, (seed, trial, mode, lazy, bm)
        if trial % 5 == 0:
            tc = transpile(c, basis_gates=['cx', 'id', 'rz', 'sx', 'x'])
            lg, _, _ = _logical(ir.lower(tc), 'clique', True, 4)
            assert abs(abs(np.vdot(lg, psi)) - 1.0) < 1e-9, (seed, trial, 'transpiled')
