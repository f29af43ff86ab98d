"""shard_plan on 2/4/8 virtual ranks vs the oracle's gate-by-gate statevector (CPU only)."""
import numpy as np
import pytest

import engine_emulator as em
import virtual_cluster as vc
from oracle import program, statevector as sv
from qcmrf_b200 import QCMRF, fusion, ir, transpile
from qcmrf_b200.circuit import QuantumCircuit

CASES = [
    ([[0, 1], [1, 2], [2, 3]], 11),
    ([[0, 1, 2], [2, 3]], 12),
    ([[0], [0, 1], [1, 2], [0, 2]], 13),
]


def _theta(C, seed):
    rng = np.random.RandomState(seed)
    return list(-np.abs(rng.randn(sum(2 ** len(c) for c in C))) * 0.7)


@pytest.mark.parametrize('g', [1, 2, 3])
@pytest.mark.parametrize('mode', ['lazy-canonical', 'lazy-auto', 'dense', 'dense-fused'])
def test_qcmrf_sharded_matches_oracle(g, mode):
    for C, seed in CASES:
        th = _theta(C, seed)
        n, k, N, _ = program.sizes(C)
        want, _ = sv.run_program(program.qcmrf_program(C, th)[0], N)
        fc = fusion.fuse(ir.lower(QCMRF(C, th)), 'clique')
        if mode.startswith('dense'):
            pl = fusion.plan(fc, lazy=False)
        elif mode == 'lazy-auto':
            if len(fusion.control_only_qubits(fc)) < g:
                continue
            pl = fusion.plan(fc, lazy=True, block_max=3, n_global=g)
        else:
            pl = fusion.plan(fc, lazy=True, block_max=3)
        if pl.n_phys - g < 1:
            continue
        psi, sps = vc.run_virtual(pl, g, fuse_exchange=(mode == 'dense-fused'))
        got = vc.logical_state(pl, psi, sps)
        assert np.abs(got - want).max() < 1e-12, (C, mode, g)
        if mode == 'dense-fused':
            kinds = [seg[0] for seg in sps[0].segments]
            assert kinds.count('xblock') == 1 and 'exchange' not in kinds     # swap + the last g clique sweeps: one kernel
            assert kinds[-1] == 'xblock'
        if mode.startswith('dense'):
            assert sps[0].n_exchanges == 1               # one all-to-all moves every global ancilla on-GPU
        else:
            assert sps[0].n_exchanges == 0               # lazily materialised: communication-free
        if mode == 'lazy-auto':
            assert all(sp.mat_mask == (1 << g) - 1 for sp in sps)


@pytest.mark.parametrize('g', [1, 2])
def test_generic_circuit_with_repeated_global_targets(g):
    """A foreign circuit whose high qubits are targeted again and again: exchanges with
    lookahead, local swaps, forced materialisation."""
    rng = np.random.RandomState(3 + g)
    N = 6
    qc = QuantumCircuit(N, N)
    for q in range(N):
        qc.h(q)
    for rep in range(4):
        for q in rng.permutation(N):
            q = int(q)
            c = int(rng.choice([x for x in range(N) if x != q]))
            qc.cp(float(rng.uniform(0, 3)), c, q)
            qc.h(q)
            qc.cx(c, q)
    prog = ir.lower(qc)
    want, _ = sv.run_program(ir.to_oracle_ops(prog), N)
    for lazy, mode in ((False, 'off'), (True, 'off'), (True, 'clique'), (False, 'clique')):
        fc = fusion.fuse(prog, mode, use_hint=False)
        pl = fusion.plan(fc, lazy=lazy, block_max=2)
        for fused in (False, True):
            psi, sps = vc.run_virtual(pl, g, fuse_exchange=fused)
            got = vc.logical_state(pl, psi, sps)
            assert np.abs(got - want).max() < 1e-12, (lazy, mode, fused)
    assert sps[0].n_exchanges >= 1


@pytest.mark.parametrize('seed', range(3))
def test_random_generic_circuits_on_virtual_ranks(seed):
    """Seeded fuzz of the sharded planner (sharded.plan_for_shards, what ShardedSimulator.prepare calls) on 2, 4 and 8
    virtual ranks: random circuits over the whole gate surface, eager and lazy plans, NCCL-style and fused exchanges.  A
    lazily materialised plan that would need an exchange is replaced by the dense plan on every rank alike (lazy plans
    with exchanges had ranks disagreeing on the segment list and exchanges of partially materialised shards); the state
    must be the textbook one exactly, and no circuit may be refused."""
    from test_host_fusion import _random_circuit, _textbook_state
    from qcmrf_b200 import sharded
    rng = np.random.RandomState(8100 + seed)
    fallbacks = lazy_kept = 0
    for trial in range(8):
        nq = int(rng.randint(4, 8))
        c = _random_circuit(rng, nq, int(rng.randint(4, 30)))
        psi = _textbook_state(c)
        prog = ir.lower(c)
        for lazy, mode in ((False, 'off'), (True, 'off'), (True, 'clique'), (False, 'clique')):
            fc = fusion.fuse(prog, mode, use_hint=False)
            for g in (1, 2, 3):
                for fused in (False, True):
                    pairs = [sharded.plan_for_shards(fc, g, r, lazy, 2, 0, 8, fused) for r in range(1 << g)]
                    pl = pairs[0][0]
                    assert all(np.array_equal(p.ops, pl.ops) and np.array_equal(p.tables, pl.tables) for p, _ in pairs)
                    kinds = [[s[0] for s in q.segments] for _, q in pairs]
                    assert all(k == kinds[0] for k in kinds), (seed, trial, lazy, mode, g, fused)    # one collective schedule
                    st, sps = vc.run_virtual(pl, g, fuse_exchange=fused)
                    got = vc.logical_state(pl, st, sps)
                    assert np.abs(got - psi).max() < 1e-12, (seed, trial, lazy, mode, g, fused)
                    if lazy:
                        if any(sp.n_exchanges for sp in sps):
                            fallbacks += 1
                        else:
                            lazy_kept += 1
    assert fallbacks and lazy_kept
