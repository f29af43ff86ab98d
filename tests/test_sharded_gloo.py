"""ShardedSimulator's host logic and collectives on 2 CPU ranks (gloo): the engine is replaced by
the numpy emulator (tests/fake_native.py), the state lives in torch CPU tensors, and the
qubit-swap exchange, pmf gather and shot merge run through torch.distributed for real."""
import os
import socket
import sys
import traceback

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, HERE)
        sys.path.insert(0, os.path.dirname(HERE))
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        import torch
        import torch.distributed as dist
        dist.init_process_group('gloo', rank=rank, world_size=world)
        import fake_native
        from oracle import mrf, program, statevector as sv
        from qcmrf_b200 import QCMRF, _native, sharded

        _native.Handle = fake_native.Handle

        class CpuSharded(sharded.ShardedSimulator):
            def _tensor_device(self):
                return torch.device('cpu')

        rng = np.random.RandomState(7)
        out = {}
        for C in ([[0, 1], [1, 2], [2, 3]], [[0, 1, 2], [1, 3]]):
            th = list(-np.abs(rng.randn(sum(2 ** len(c) for c in C))) * 0.6)
            n, k, N, _ = program.sizes(C)
            pb, db, _ = mrf.brute_force_pmf(C, th)
            psi, meas = sv.run_program(program.qcmrf_program(C, th)[0], N)
            kp = sv.key_probabilities(psi, N, meas)
            for fus, layout, prec, xch in (('blocked', 'auto', 'double', 'nccl'), ('blocked', 'canonical', 'double', 'nccl'),
                                           ('clique', 'canonical', 'double', 'nccl'), ('clique', 'canonical', 'single', 'nccl'),
                                           ('clique', 'canonical', 'double', 'p2p')):   # p2p on CPU: fused plan, fallback path
                sim = CpuSharded(precision=prec, fusion=fus, layout=layout, block_max=2, seed=5,
                                 staging_bytes=1 << 9, exchange=xch)   # tiny staging: several chunks per exchange
                res = sim.run(QCMRF(C, th), shots=20000).result()
                p, delta = res.postselected_probabilities(0)
                tol = 1e-10 if prec == 'double' else 1e-5
                assert np.abs(p - pb).max() < tol and abs(delta - db) < tol, (fus, layout, np.abs(p - pb).max())
                counts = res.get_counts()
                assert sum(counts.values()) == 20000
                obs = np.zeros(1 << N)
                for key, v in counts.items():
                    assert len(key) == N
                    obs[int(key, 2)] = v
                assert obs[kp < 1e-14].sum() == 0, (fus, layout)
                assert 0.5 * np.abs(obs / 2e4 - kp).sum() < 0.06
                meta = res.metadata(0)
                assert meta['exchanges'] == (1 if fus == 'clique' else 0)
                out[(tuple(map(tuple, C)), fus, layout, prec, xch)] = (dict(counts), float(delta))
                # a second and third parameter vector on the same graph: the plan cache serves them (same structure,
                # refreshed tables) -- results must be those of the new parameters
                for rep in range(2):
                    th2 = list(-np.abs(rng.randn(len(th))) * 0.6)
                    if rep == 1:
                        th2[0] = 0.0                           # gamma == 0: a skipped term (identity entry)
                    pb2, db2, _ = mrf.brute_force_pmf(C, th2)
                    p2, d2 = sim.run(QCMRF(C, th2), shots=0).result().postselected_probabilities(0)
                    assert np.abs(p2 - pb2).max() < tol and abs(d2 - db2) < tol, (fus, layout, 'cached plan')
                assert sim._plan_cache.hits >= 2 and sim._plan_cache.misses == 1, (sim._plan_cache.hits, sim._plan_cache.misses)
                sim.close()
        q.put((rank, 'ok', out))
        dist.destroy_process_group()
    except Exception:
        q.put((rank, 'fail', traceback.format_exc()))


def test_two_rank_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = {}
    for _ in procs:
        rank, status, payload = q.get(timeout=300)
        assert status == 'ok', payload
        results[rank] = payload
    for p in procs:
        p.join(timeout=60)
    assert results[0] == results[1]                   # every rank returns the same counts and delta


def _fuzz_worker(rank, world, port, q):
    try:
        sys.path.insert(0, HERE)
        sys.path.insert(0, os.path.dirname(HERE))
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        import torch
        import torch.distributed as dist
        dist.init_process_group('gloo', rank=rank, world_size=world)
        import fake_native
        from test_host_fusion import _random_circuit, _textbook_state
        from qcmrf_b200 import _native, sharded
        from qcmrf_b200.circuit import QuantumCircuit

        _native.Handle = fake_native.Handle

        class CpuSharded(sharded.ShardedSimulator):
            def _tensor_device(self):
                return torch.device('cpu')

        rng = np.random.RandomState(9811)
        out, general = [], 0
        for trial in range(12):
            nq = int(rng.randint(3, 7))
            base = _random_circuit(rng, nq, int(rng.randint(3, 24)))
            pw = np.abs(_textbook_state(base)) ** 2
            nv = int(rng.randint(1, nq + 1))
            c = QuantumCircuit(nq, nq)
            for ins in base.data:
                c._qc_add(ins.operation, list(ins.qubits))
            c.measure(range(nq), range(nq))
            kept = pw[:1 << nv].sum()
            for fus, layout, xch in (('blocked', 'auto', 'nccl'), ('blocked', 'canonical', 'nccl'), ('clique', 'canonical', 'nccl'),
                                     ('clique', 'canonical', 'p2p'), ('off', 'canonical', 'nccl')):
                sim = CpuSharded(precision='double', fusion=fus, layout=layout, block_max=2, seed=5, staging_bytes=1 << 9, exchange=xch)
                res = sim.run(c, shots=1000, n_vars=nv).result()
                p, d = res.postselected_probabilities(0)
                assert abs(d - kept) < 1e-10, (trial, fus, layout, xch)
                if kept > 1e-9:
                    assert np.abs(p - pw[:1 << nv] / kept).max() < 1e-10, (trial, fus, layout, xch)
                cnt = res.get_counts()
                assert sum(cnt.values()) == 1000 and all(pw[int(k, 2)] > 1e-14 for k in cnt), (trial, fus, layout, xch)
                out.append((dict(cnt), float(d)))
                sim.close()
        q.put((rank, 'ok', out))
        dist.destroy_process_group()
    except Exception:
        q.put((rank, 'fail', traceback.format_exc()))


def test_two_rank_gloo_random_generic_circuits():
    """Seeded fuzz of ShardedSimulator.run on 2 gloo ranks: random foreign circuits (variables anywhere in the first-use
    layout: the general post-selection path; targets on the global qubit: dense re-planning), every layout / fusion /
    exchange setting -- pmf and success probability against a textbook simulation, keys only where the exact
    distribution has mass, both ranks the same counts."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fuzz_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = {}
    for _ in procs:
        rank, status, payload = q.get(timeout=600)
        assert status == 'ok', payload
        results[rank] = payload
    for p in procs:
        p.join(timeout=60)
    assert results[0] == results[1] and len(results[0]) == 60
