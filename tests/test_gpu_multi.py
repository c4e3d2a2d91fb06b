"""Multi-GPU tests (-m gpu; skipped on a box with fewer than two GPUs -- run with `gpurun --gpus 2`): the numerical
check of the data-parallel train step (tools/gpu_ddp_check.py) under torchrun with NCCL."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_data_parallel_gradient_is_the_mean_and_replicas_stay_identical():
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', '29517', os.path.join(ROOT, 'tools', 'gpu_ddp_check.py')]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith('{')]
    assert p.returncode == 0 and lines, p.stdout[-2000:] + p.stderr[-2000:]
    rec = json.loads(lines[-1])
    assert rec['ok'] and rec['world'] == 2, rec
