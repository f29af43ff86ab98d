"""All ranks of a sharded plan in one process (TEST INFRASTRUCTURE): every rank's segments run on
the numpy engine emulator, exchanges are done with numpy slab swaps.  Checks the index
arithmetic of qcmrf_b200.sharded.shard_plan for 2, 4 and 8 ranks without any GPU."""
import numpy as np

import engine_emulator as em
from qcmrf_b200 import sharded


class _P:
    pass


def run_virtual(pl, g, fuse_exchange=False):
    """Returns (list of per-rank local states, list of ShardedPlans)."""
    world = 1 << g
    sps = [sharded.shard_plan(pl, g, r, fuse_exchange=fuse_exchange) for r in range(world)]
    nl = sps[0].n_local
    psi = [None] * world
    act = [0] * world
    nseg = len(sps[0].segments)
    assert all(len(sp.segments) == nseg for sp in sps), 'ranks disagree on the segment structure'
    for si in range(nseg):
        kind = sps[0].segments[si][0]
        assert all(sp.segments[si][0] == kind for sp in sps)
        if kind == 'run':
            for r, sp in enumerate(sps):
                _, ops, tabs, mask = sp.segments[si]
                p = _P()
                p.ops, p.tables, p.n_phys = ops, tabs, nl
                psi[r], act[r] = em.run_plan(p, n_global=g, rank=r & mask, n_local=nl, psi0=psi[r], active0=act[r])
        else:
            betas = sps[0].segments[si][1]
            assert all(sp.segments[si][1] == betas for sp in sps)
            s = len(betas)
            slab = 1 << (nl - s)
            assert all(a == nl for a in act), 'exchange on a partially materialised shard'
            old = [x.reshape(1 << s, slab).copy() for x in psi]
            for r in range(world):
                c = sum(((r >> b) & 1) << i for i, b in enumerate(betas))
                base = r
                for b in betas:
                    base &= ~(1 << b)
                new = old[r].copy()
                for j in range(1 << s):
                    peer = base
                    for i, b in enumerate(betas):
                        peer |= ((j >> i) & 1) << b
                    new[j] = old[peer][c]
                psi[r] = new.reshape(-1)
            if kind == 'xblock':                       # fused: the sweeps on the swapped-in qubits follow at once
                for r, sp in enumerate(sps):
                    _, _, ops, tabs, mask = sp.segments[si]
                    p = _P()
                    p.ops, p.tables, p.n_phys = ops, tabs, nl
                    psi[r], act[r] = em.run_plan(p, n_global=g, rank=r & mask, n_local=nl, psi0=psi[r], active0=act[r])
    return psi, sps


def logical_state(pl, psi, sps):
    """Assemble the logical little-endian statevector from the ranks' shards."""
    g = sps[0].g
    nl = sps[0].n_local
    pos = sps[0].pos
    assert all(sp.pos == pos for sp in sps)
    N = pl.n_logical
    full = np.zeros(1 << (nl + g), dtype=np.complex128)
    for r, sp in enumerate(sps):
        if r & ~sp.mat_mask:
            continue                                   # replica of another rank
        full[r << nl:(r + 1) << nl] = psi[r]
    idx = np.arange(1 << N, dtype=np.int64)
    pidx = np.zeros(1 << N, dtype=np.int64)
    dead = np.zeros(1 << N, dtype=bool)
    for q in range(N):
        b = (idx >> q) & 1
        p = pl.layout[q]
        if p < pl.n_phys:
            pidx |= b << pos[p]
        else:
            dead |= b == 1
    out = full[pidx]
    out[dead] = 0.0
    return out * np.exp(1j * pl.global_phase)
