"""-m gpu, needs >= 2 GPUs: data-parallel step over NCCL == single-GPU step on the concatenated
batch.  Exercises the product's flat CUDA path: per-rank backward into
the flat gradient buffer, NCCL allreduce of that buffer (tools.GradSync.flat), fused clip + Adam."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_dp_flat_adam_step_equals_global_batch_step(pkg, device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29677", os.path.join(HERE, "dp_nccl_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("DP_RESULT ")]
    assert line, out.stdout[-2000:]
    r = json.loads(line[-1][len("DP_RESULT "):])
    assert r["replicas_identical"], r
    for name, lr in (("model", 1e-4), ("actor", 3e-5), ("value", 3e-5)):
        assert r[f"{name}_flat_grad_rel"] < 2e-5, r       # same gradient up to summation order
        assert r[f"{name}_grad_norm_rel"] < 2e-5, r
        assert r[f"{name}_param_maxabs"] <= 2 * lr + 1e-6, r   # Adam sign noise on ~zero gradients
