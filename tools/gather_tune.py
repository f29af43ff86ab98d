"""Tuning aid for the fused qubit-swap + gate pass (k_block_gather / k_block_gather_tma), not product code.
    torchrun --nproc-per-node 2 tools/gather_tune.py [n_local]
Every rank owns a 2^n_local complex64 shard (+ the output buffer); the kernel variants are selected
through the library's QCM_GATHER* environment knobs and timed with CUDA events around the call."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcmrf_b200 import _native, fusion          # noqa: E402


def main():
    rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(lr)
    dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
    s = world.bit_length() - 1
    nl = int(sys.argv[1]) if len(sys.argv) > 1 else 31
    bufs = [torch.zeros(2 << nl, dtype=torch.float32, device='cuda') for _ in range(2)]
    bufs[0].normal_()
    meta = [_native.ipc_export(lr, b.data_ptr()) for b in bufs]
    gathered = [None] * world
    dist.all_gather_object(gathered, meta)
    opened = {}
    ptrs = {}
    for r in range(world):
        if r == rank:
            ptrs[r] = [b.data_ptr() for b in bufs]
            continue
        ptrs[r] = []
        for hd, off in gathered[r]:
            if hd not in opened:
                opened[hd] = _native.ipc_open(lr, hd)
            ptrs[r].append(opened[hd] + off)
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
    shard_bytes = 8 << nl
    h = _native.Handle(nl, 'single', lr, ext_state_ptr=bufs[0].data_ptr())
    h.set_shard(s, rank)
    h.set_active(nl)
    variants = [('ldg', 1, 0), ('ldg', 2, 0), ('ldg', 4, 0), ('tma', 1, 4), ('tma', 1, 16), ('tma', 1, 64), ('tma', 1, 256),
                ('tma', 2, 4), ('tma', 2, 16), ('tma', 2, 64)]
    ref = None
    for mode, U, K in variants:
        os.environ['QCM_GATHER'] = mode
        os.environ['QCM_GATHER_U'] = str(U)
        os.environ['QCM_GATHER_K'] = str(K or 16)
        times = []
        for it in range(3):
            src = [ptrs[j][0] + rank * slab_bytes for j in range(world)]
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            h.run_gather_block(ops, tabs, src, bufs[1].data_ptr())
            e1.record()
            torch.cuda.synchronize(); dist.barrier()
            times.append(e0.elapsed_time(e1))
        chk = float(bufs[1][::4097].double().sum())
        if ref is None:
            ref = chk
        ms = min(times)
        t = torch.tensor([ms], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            remote = shard_bytes * (world - 1) // world
            print('%s U=%d K=%-3d  %.3f ms (max over ranks %.3f)  nvlink %.0f GB/s  local hbm %.0f GB/s  same=%s'
                  % (mode, U, K, ms, float(t), remote / float(t) / 1e6, (shard_bytes // world + shard_bytes) / float(t) / 1e6,
                     abs(chk - ref) < 1e-6 * max(1.0, abs(ref))), flush=True)
    h.close()
    for b in opened.values():
        _native.ipc_close(lr, b)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
