#!/usr/bin/env python
"""Low-order-target gate passes (north star (ii)): H on qubit t of an n-qubit state, one in-place pass per
gate through the C ABI, timed by the engine's per-op CUDA events.  Prints one JSON line:

    python tools/lowq_bench.py [--qubits 30] [--precision single] [--targets 0,1,2,4,8,12]
    QCM_LOWQ=0 python tools/lowq_bench.py ...      # the same gates through k_block (A/B)
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--qubits', type=int, default=30)
    ap.add_argument('--precision', default='single')
    ap.add_argument('--targets', default='0,1,2,3,4,5,6,8,12,20')
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--ctrl', type=int, default=0, help='index qubits per gate (taken right above the target)')
    args = ap.parse_args()
    from qcmrf_b200 import _native, fusion
    N = args.qubits
    targets = [int(t) for t in args.targets.split(',')]
    H = np.array([[1, 1], [1, -1]], dtype=np.complex128) / np.sqrt(2)
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        peak = 6650.0
    rows = []
    with _native.Handle(N, args.precision) as h:
        e = fusion._Emitter()
        qv = np.zeros((N, 4)); qv[:, 0] = 0.6; qv[:, 3] = 0.8
        e.op(fusion.QCM_OP_INIT_PRODUCT, n_in=0, n_out=N, table_off=e.table(qv))
        for _ in range(args.reps):
            for t in targets:
                ctrl = [(t + 1 + j) % N for j in range(args.ctrl)]
                tab = np.tile(H, (1 << len(ctrl), 1, 1))
                e.op(fusion.QCM_OP_MUX1Q, target=t, ctrl=ctrl, n_in=N, n_out=N, table_off=e.table(fusion._mux_table_f64(tab)))
        ops, tabs = e.finish()
        h.run_program(ops, tabs)                     # warm-up
        h.run_program(ops, tabs)
        prof = h.op_profile()[1:]
        names = h.op_kernels()[1:]
        norm = None
        if N <= 30:
            p, kept = h.postselect(0, 0, 0)
            norm = float(kept)
    for k, t in enumerate(targets):
        ms = [prof[r * len(targets) + k][1] for r in range(args.reps)]
        by = prof[k][2] + prof[k][3]
        med = float(np.median(ms))
        rows.append({'target': t, 'kernel': names[k], 'ms': med, 'gbs': by / med / 1e6, 'frac_of_measured_peak': by / med / 1e6 / peak})
    print(json.dumps({'qubits': N, 'precision': args.precision, 'index_qubits': args.ctrl, 'lowq': os.environ.get('QCM_LOWQ', '1'),
                      'bytes_per_pass': int(prof[0][2] + prof[0][3]), 'peak_gbs': peak, 'norm_after': norm, 'passes': rows}))


if __name__ == '__main__':
    main()
