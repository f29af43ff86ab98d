"""Qubit-swap all-to-all strategies over NCCL on the GPUs of one box (tuning aid, not product code).
    torchrun --nproc-per-node 8 tools/exchange_bench.py [slab_mib]
Every rank owns W slabs of `slab_mib` MiB; slab j goes to rank j, which stores it as slab <sender>."""
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
    slab_mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    slab = slab_mib * (1 << 20) // 4
    st = torch.empty((world, slab), dtype=torch.float32, device='cuda')
    st.copy_(torch.arange(world, device='cuda', dtype=torch.float32)[:, None] + 100 * rank)
    peers = [j for j in range(world) if j != rank]
    sent = (world - 1) * slab * 4

    def check():
        want = torch.tensor([rank + 100 * j for j in range(world)], device='cuda', dtype=torch.float32)
        ok = bool((st[:, 0] == want).all() and (st[:, -1] == want).all())
        return ok

    def p2p(chunk_mib, nbuf=2):
        chunk = min(slab, chunk_mib * (1 << 20) // 4)
        stage = torch.empty((nbuf, len(peers), chunk), dtype=torch.float32, device='cuda')
        n_chunks = (slab + chunk - 1) // chunk
        pending = None
        for ci in range(n_chunks + 1):
            works = None
            if ci < n_chunks:
                off = ci * chunk
                n = min(chunk, slab - off)
                ops = []
                for k, j in enumerate(peers):
                    ops.append(dist.P2POp(dist.isend, st[j, off:off + n], j))
                    ops.append(dist.P2POp(dist.irecv, stage[ci % nbuf, k, :n], j))
                works = (dist.batch_isend_irecv(ops), off, n, ci % nbuf)
            if pending is not None:
                ws, poff, pn, pb = pending
                for w in ws:
                    w.wait()
                for k, j in enumerate(peers):
                    st[j, poff:poff + pn].copy_(stage[pb, k, :pn])
            pending = works

    def a2a_list(chunk_mib):
        chunk = min(slab, chunk_mib * (1 << 20) // 4)
        stage = torch.empty((world, chunk), dtype=torch.float32, device='cuda')
        for off in range(0, slab, chunk):
            n = min(chunk, slab - off)
            dist.all_to_all([stage[j, :n] for j in range(world)], [st[j, off:off + n] for j in range(world)])
            st[:, off:off + n].copy_(stage[:, :n])

    def a2a_single(chunk_mib):
        chunk = min(slab, chunk_mib * (1 << 20) // 4)
        sin = torch.empty((world, chunk), dtype=torch.float32, device='cuda')
        sout = torch.empty((world, chunk), dtype=torch.float32, device='cuda')
        for off in range(0, slab, chunk):
            n = min(chunk, slab - off)
            sin[:, :n].copy_(st[:, off:off + n])
            dist.all_to_all_single(sout[:, :n].contiguous() if n != chunk else sout, sin[:, :n].contiguous() if n != chunk else sin)
            st[:, off:off + n].copy_(sout[:, :n])

    def timed(name, f, *a):
        f(*a)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        f(*a)                                    # two exchanges restore the original layout
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f(*a)
        e1.record()
        torch.cuda.synchronize()
        ok = check()
        f(*a)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device='cuda')
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            print('%-34s %8.2f ms  %7.1f GB/s per direction per GPU  ok=%s' % (name + str(a), float(ms), sent / float(ms) / 1e6, ok), flush=True)

    if rank == 0:
        print('world %d, slab %d MiB, %0.2f GB sent per GPU' % (world, slab_mib, sent / 1e9), flush=True)
    for c in (32, 128, 512):
        timed('p2p batch_isend_irecv', p2p, c)
    timed('p2p 3 buffers', p2p, 128, 3)
    for c in (128, 512):
        timed('all_to_all (lists)', a2a_list, c)
    for c in (128, 512):
        timed('all_to_all_single', a2a_single, c)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
