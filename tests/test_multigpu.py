"""The only collective on the path on real GPUs: skipped unless the box has >= 2 CUDA devices.  Runs
tools/check_stop_nccl.py under torchrun on 2 GPUs: a batch sharded over the ranks, the asynchronous residual stopping rule
(per-sample sums reduced on the device, NCCL all-reduce on a side stream, decision consumed one check later); every rank must
stop at the same iteration and the gathered result must equal a single-process run of the whole batch."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_residual_stop_allreduce_over_nccl_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "check_stop_nccl.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "on every rank: True" in r.stdout
