"""Multi-rank parity on real GPUs (NCCL): the sharded paths -- delay-grid chains of one spectrum, parameter sets of
``qevolve_batch``, members of an ensemble -- return on every rank what a single GPU returns.  Spawns
``torch.distributed.run`` on ``tools/dist_check.py`` with as many ranks as GPUs are visible (2 at least, 4 at most);
skipped on a single-GPU box (the gloo tests in test_dist_gloo.py / test_response_tree_cpu.py cover the logic there)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_paths_equal_single_gpu_under_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ranks = min(n, 4)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(ranks),
           '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.join(REPO, 'tools', 'dist_check.py')]
    done = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=REPO)
    assert done.returncode == 0, done.stdout[-2000:] + done.stderr[-4000:]
    assert 'max |sharded - single| over ranks' in done.stdout
