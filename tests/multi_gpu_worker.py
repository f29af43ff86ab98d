"""Worker of tests/test_multi_gpu.py: one rank per GPU under torchrun (NCCL).  Checks the sharded
engine against brute force, against the single-GPU engine, and rank-to-rank equality."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    import torch
    import torch.distributed as dist
    rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(lr)
    dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
    from oracle import mrf
    from qcmrf_b200 import QCMRF, B200Simulator, workloads
    from qcmrf_b200.sharded import ShardedSimulator
    n = 11
    C = workloads.random_tree(n, 1, seed=5)                      # k = 11 -> N = 23 qubits
    th = workloads.theta_for(C, seed=6)
    pb, db, _ = mrf.brute_force_pmf(C, th)
    report = {}
    ref_counts = None
    if rank == 0:
        single = B200Simulator(precision='double', device=lr, seed=77, small_batch=False)
        psi = single.statevector(QCMRF(C, th))
        single.close()
        kp_full = np.abs(psi) ** 2
    quick = os.environ.get('QCM_MGW_QUICK') == '1'           # smoke(): the default layout + the NCCL exchange, complex64
    layouts = [('blocked', 'auto', 'nccl'), ('blocked', 'canonical', 'nccl'), ('clique', 'canonical', 'nccl'),
               ('clique', 'canonical', 'p2p'), ('clique', 'canonical', 'p2p-inplace')]
    for prec, tol in ((('single', 1e-5),) if quick else (('double', 1e-10), ('single', 1e-5))):
        for fus, layout, xch in ([layouts[0], layouts[2]] if quick else layouts):
            sim = ShardedSimulator(precision=prec, fusion=fus, layout=layout, device=lr, seed=77, block_max=3,
                                   staging_bytes=1 << 22, exchange=xch)
            res = sim.run(QCMRF(C, th), shots=200000).result()
            p, delta = res.postselected_probabilities(0)
            err = float(np.abs(p - pb).max())
            rel = float((np.abs(p - pb) / pb).max())
            assert err < tol and abs(delta - db) < tol, (prec, fus, layout, err)
            assert rel < (1e-9 if prec == 'double' else 2e-4), (prec, fus, layout, rel)
            counts = res.get_counts()
            assert sum(counts.values()) == 200000
            meta = res.metadata(0)
            assert meta['exchanges'] == (1 if fus == 'clique' else 0), meta
            if xch.startswith('p2p'):
                assert meta['exchange_path'] == {'p2p': 'p2p-fused', 'p2p-inplace': 'p2p-fused-inplace'}[xch], \
                    (meta, getattr(sim, 'p2p_error', None))
            # every rank must hold identical results
            blob = json.dumps(sorted(counts.items())) + repr(float(delta))
            gathered = [None] * world
            dist.all_gather_object(gathered, blob)
            assert all(g == gathered[0] for g in gathered)
            if rank == 0:
                N = 2 * n + 1
                obs = np.zeros(1 << N)
                for k, v in counts.items():
                    obs[int(k, 2)] = v
                # clbit c == qubit c here (every qubit but the scratch one is measured into its own clbit)
                assert obs[kp_full < 1e-18].sum() == 0
                # full-width support is 2^22 outcomes: test the variable marginal (2^n) and the success rate
                idx = np.arange(1 << N)
                marg = np.bincount(idx & ((1 << n) - 1), weights=kp_full, minlength=1 << n)
                omarg = np.bincount(idx & ((1 << n) - 1), weights=obs, minlength=1 << n) / 2e5
                tv = 0.5 * np.abs(omarg - marg).sum()
                bound = 0.5 * np.sqrt(2.0 * ((1 << n) * np.log(2.0) + np.log(1e6)) / 2e5)
                assert tv < bound, (tv, bound)
                succ = obs[: 1 << n].sum() / 2e5
                assert abs(succ - db) < 5 * np.sqrt(db * (1 - db) / 2e5) + 1e-4, (succ, db)
                report['%s/%s/%s/%s' % (prec, fus, layout, xch)] = {'max_p_err': err, 'tv': float(tv), 'exchanges': meta['exchanges']}
            sim.close()
    # a LIST of circuits is pipelined (deferred result handling + a worker thread): same results as one at a time
    sim = ShardedSimulator(precision='double', device=lr, seed=77)
    ths = [workloads.theta_for(C, seed=60 + k) for k in range(4)]
    res = sim.run([QCMRF(C, t_) for t_ in ths], shots=20000).result()
    one = sim.run(QCMRF(C, ths[0]), shots=20000).result()
    assert res.get_counts(0) == one.get_counts()
    for k, t_ in enumerate(ths):
        pk, dk = res.postselected_probabilities(k)
        pbk, dbk, _ = mrf.brute_force_pmf(C, t_)
        assert np.abs(pk - pbk).max() < 1e-10 and abs(dk - dbk) < 1e-10, k
        assert sum(res.get_counts(k).values()) == 20000
    blob = json.dumps([sorted(c.items()) for c in res.get_counts()])
    gathered = [None] * world
    dist.all_gather_object(gathered, blob)
    assert all(g == gathered[0] for g in gathered)
    sim.close()
    if rank == 0:
        report['pipelined_list'] = 'ok'
        print('MULTI_GPU_OK ' + json.dumps(report), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
