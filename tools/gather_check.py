"""Diagnostic for the fused gather kernels on ONE GPU (virtual peers): every variant must produce
bit-identical output; prints the mismatch pattern otherwise.  python tools/gather_check.py [n_local] [s]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcmrf_b200 import _native, fusion          # noqa: E402

nl = int(sys.argv[1]) if len(sys.argv) > 1 else 26
s = int(sys.argv[2]) if len(sys.argv) > 2 else 1
world = 1 << s
old = [torch.randn(2 << nl, dtype=torch.float32, device='cuda') for _ in range(world)]
out = torch.zeros(2 << nl, dtype=torch.float32, device='cuda')
rng = np.random.RandomState(3)
e = fusion._Emitter()
tq = list(range(nl - s, nl))


def rand_u(m):
    q, _ = np.linalg.qr(rng.randn(1 << m, 2, 2) + 1j * rng.randn(1 << m, 2, 2))
    return q


members = [(fusion.QCM_OP_MUX1Q, t, [3, 7], fusion._mux_table_f64(rand_u(2))) for t in tq]
if s == 1:
    k, t, c, tab = members[0]
    e.op(k, target=t, ctrl=c, n_in=nl, n_out=nl, table_off=e.table(tab))
else:
    e.op(fusion.QCM_OP_BLOCK, target=s, ctrl=tq, n_in=nl, n_out=nl, n_ctrl=len(members))
    for k, t, c, tab in members:
        e.op(k, target=t, ctrl=c, n_in=nl, n_out=nl, table_off=e.table(tab))
ops, tabs = e.finish()
slab_bytes = 8 << (nl - s)
h = _native.Handle(nl, 'single', 0, ext_state_ptr=old[0].data_ptr())
h.set_shard(s, 0)
h.set_active(nl)
src = [old[j].data_ptr() for j in range(world)]
ref = None
S = {1: 8, 2: 6, 3: 3}[s]
for mode, U, K in [('ldg', 2, 0), ('tma', 1, 16), ('tma', 1, 32), ('tma', 1, 64), ('tma', 2, 16), ('tma', 2, 64)]:
    os.environ['QCM_GATHER'] = mode
    os.environ['QCM_GATHER_U'] = str(U)
    os.environ['QCM_GATHER_K'] = str(K or 16)
    nbad = 0
    for it in range(int(os.environ.get('ITERS', '30'))):
        out.zero_()
        torch.cuda.synchronize()
        h.run_gather_block(ops, tabs, src, out.data_ptr())
        torch.cuda.synchronize()
        if ref is None:
            ref = out.clone()
        bad = (out != ref).nonzero().flatten()
        if bad.numel():
            nbad += 1
            b = bad.cpu().numpy()
            tile_f = 1024 * U                                  # floats per tile per source
            slab_f = 2 << (nl - s)
            i0 = int(b[0])
            r, off = i0 // slab_f, i0 % slab_f
            tile = off // tile_f
            k = tile % K
            o = out[i0:i0 + 4].cpu().numpy()
            w = ref[i0:i0 + 4].cpu().numpy()
            alt = {}
            for d in (-S, S, -1, 1):
                j = i0 + d * tile_f
                if 0 <= j < out.numel():
                    alt[d] = bool((ref[j:j + 4].cpu().numpy() == o).all())
            print('  %s U=%d K=%d it=%d: %d bad floats, r=%d tile=%d k=%d warp=%d lanefloat=%d got=%s want=%s zero=%s matches ref at tile offset: %s'
                  % (mode, U, K, it, bad.numel(), r, tile, k, (off % tile_f) // 128, off % 128, o, w, bool((o == 0).all()), alt), flush=True)
    print('%s U=%d K=%-3d  %.3f ms  iterations with mismatches: %d' % (mode, U, K, h.timing()['program_ms'], nbad), flush=True)
h.close()
