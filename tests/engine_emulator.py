"""Executable specification of the qcm_op semantics (include/qcmrf_b200.h) in numpy.

TEST INFRASTRUCTURE ONLY: lets the CPU-side tests check that the fusion pass and the
pass planner emit programs that mean the same thing as the oracle's gate-by-gate
execution, without a GPU.  The product never imports this.
"""
import numpy as np

from qcmrf_b200 import fusion as F


def run_plan(plan, n_global=0, rank=0, n_local=None, psi0=None, active0=0):
    """Execute plan.ops/plan.tables; returns the physical state (2^n_local complex128, implicit
    zeros beyond the materialised part) and the final n_active.  psi0/active0 continue from an
    earlier call (segments of a sharded plan)."""
    nl = plan.n_phys if n_local is None else n_local
    if psi0 is None:
        psi = np.full(1 << nl, np.nan + 0j, dtype=np.complex128)      # NaN = never written
    else:
        psi = psi0.copy()
    active = active0
    tabs = plan.tables
    ops = plan.ops
    rank_bits = rank << nl
    i = 0

    def mux(op, n_in, n_out):
        t = int(op['target'])
        m = int(op['n_ctrl'])
        ctrl = [int(c) for c in op['ctrl'][:m]]
        tab = tabs[int(op['table_off']):int(op['table_off']) + (8 << m)].reshape(-1, 4, 2)
        tab = (tab[..., 0] + 1j * tab[..., 1]).reshape(-1, 2, 2)
        idx = np.arange(1 << n_out, dtype=np.int64)
        i0 = idx[((idx >> t) & 1) == 0]
        gi = i0 | rank_bits
        ti = np.zeros(len(i0), dtype=np.int64)
        for j, c in enumerate(ctrl):
            ti |= ((gi >> c) & 1) << j
        a0 = psi[i0].copy()
        a1 = psi[i0 | (1 << t)].copy()
        M = tab[ti]
        psi[i0] = M[:, 0, 0] * a0 + M[:, 0, 1] * a1
        psi[i0 | (1 << t)] = M[:, 1, 0] * a0 + M[:, 1, 1] * a1

    def diag(op, n_out):
        m = int(op['n_ctrl'])
        tab = tabs[int(op['table_off']):int(op['table_off']) + (2 << m)].reshape(-1, 2)
        tab = tab[:, 0] + 1j * tab[:, 1]
        idx = np.arange(1 << n_out, dtype=np.int64) | rank_bits
        ti = np.zeros(1 << n_out, dtype=np.int64)
        for j, c in enumerate(op['ctrl'][:m]):
            ti |= ((idx >> int(c)) & 1) << j
        psi[: 1 << n_out] *= tab[ti]

    while i < len(ops):
        op = ops[i]
        kind = int(op['kind'])
        n_in, n_out = int(op['n_active_in']), int(op['n_active_out'])
        if kind != F.QCM_OP_INIT_PRODUCT:
            assert n_in == active, (i, n_in, active)
        assert n_out <= nl
        if kind == F.QCM_OP_INIT_PRODUCT:
            qv = tabs[int(op['table_off']):int(op['table_off']) + 4 * max(n_out, 1)].reshape(-1, 4)
            v = np.ones(1, dtype=np.complex128)
            for q in range(n_out):
                v = np.kron(np.array([qv[q, 0] + 1j * qv[q, 1], qv[q, 2] + 1j * qv[q, 3]]), v)
            psi[: 1 << n_out] = v
        elif kind in (F.QCM_OP_MUX1Q, F.QCM_OP_BLOCK):
            if kind == F.QCM_OP_BLOCK:
                M, n_mem = int(op['target']), int(op['n_ctrl'])
                tq = [int(x) for x in op['ctrl'][:M]]
                members = [ops[i + 1 + g] for g in range(n_mem)]
                wide_ok = tq == list(range(n_in, n_out))            # every block qubit new: up to QCM_MAX_EXPAND
                assert M <= (F.QCM_MAX_EXPAND if wide_ok else F.QCM_MAX_BLOCK) and n_mem <= F.QCM_MAX_MEMBERS
                if M > F.QCM_MAX_BLOCK:
                    per = [sum(1 for mb in members if int(mb['kind']) == F.QCM_OP_MUX1Q and int(mb['target']) == q) for q in tq]
                    assert per == [1] * M, 'a wide block needs exactly one MUX1Q per qubit'
                assert tq == sorted(set(tq))
            else:
                tq, members = [int(op['target'])], [op]
            for q in range(n_in, n_out):
                assert q in tq, 'materialised qubit %d is not a block target' % q
            # the op may only read the materialised input
            assert not np.isnan(psi[: 1 << n_in]).any()
            psi[1 << n_in: 1 << n_out] = 0.0
            for mb in members:
                for c in mb['ctrl'][:int(mb['n_ctrl'])]:
                    assert int(c) not in tq
                if int(mb['kind']) == F.QCM_OP_DIAG and kind == F.QCM_OP_BLOCK:
                    diag(mb, n_out)                    # diagonal factor riding in the same sweep
                    continue
                assert int(mb['kind']) == F.QCM_OP_MUX1Q and int(mb['target']) in tq
                mux(mb, n_in, n_out)
            if kind == F.QCM_OP_BLOCK:
                i += len(members)
        elif kind == F.QCM_OP_DIAG:
            diag(op, n_out)
        elif kind == F.QCM_OP_EXTEND:
            psi[1 << n_in: 1 << n_out] = 0.0
        elif kind == F.QCM_OP_SWAP:
            a, b = sorted((int(op['target']), int(op['ctrl'][0])))
            idx = np.arange(1 << n_out, dtype=np.int64)
            sel = idx[(((idx >> a) & 1) == 1) & (((idx >> b) & 1) == 0)]
            other = sel ^ (1 << a) ^ (1 << b)
            tmp = psi[sel].copy()
            psi[sel] = psi[other]
            psi[other] = tmp
        else:
            raise AssertionError('unknown op kind %d' % kind)
        active = n_out
        i += 1
    assert not np.isnan(psi[: 1 << active]).any()
    out = psi.copy()
    out[1 << active:] = 0.0
    return out, active


def logical_state(plan, phys):
    """Physical state -> logical little-endian statevector (never-stored qubits are |0>)."""
    N = plan.n_logical
    idx = np.arange(1 << N, dtype=np.int64)
    pidx = np.zeros(1 << N, dtype=np.int64)
    dead = np.zeros(1 << N, dtype=bool)
    for q in range(N):
        b = (idx >> q) & 1
        if plan.layout[q] < plan.n_phys:
            pidx |= b << plan.layout[q]
        else:
            dead |= b == 1
    out = phys[pidx]
    out[dead] = 0.0
    return out * np.exp(1j * plan.global_phase)
