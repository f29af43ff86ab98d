"""Sharded execution on real GPUs (2 ranks, NCCL).  Skipped on boxes with fewer than 2 GPUs; the host
logic and the collectives are also covered on the CPU by test_sharded_plan.py / test_sharded_gloo.py."""
import os
import subprocess
import sys

import pytest

from conftest import has_cuda

HERE = os.path.dirname(os.path.abspath(__file__))


def _n_gpus():
    try:
        from qcmrf_b200 import _native
        return _native.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda() or _n_gpus() < 2, reason='needs 2 CUDA devices')
def test_sharded_two_gpus():
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29533', os.path.join(HERE, 'multi_gpu_worker.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert 'MULTI_GPU_OK' in out.stdout
