"""GPU, 2+ devices: the data-parallel learner on hardware.  bench.py's dp_check under torchrun: one step on
a global batch by rank 0 alone vs the same batch column-sharded over the ranks + the (two-bucket, overlapped) NCCL
SUM all-reduce -> the flat gradient agrees to fp32 summation order; after the optimizer step the replicas' parameters are
bit-identical.  Skipped on single-GPU boxes (the CPU / gloo version of the host logic is tests/test_dist_cpu.py)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("scaling", ["weak", "strong"])
def test_two_rank_step_matches_single_rank(scaling):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "4", "--warmup", "3",
           "--no_cpu_baseline", "--no_profile", "--scaling", scaling]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    dp = line["dp_check"]
    assert dp["replicas_bit_identical"] is True
    assert dp["grad_rel_l2"] < 1e-5 and dp["grad_max_over_max"] < 1e-5 and dp["loss_rel"] < 1e-6, dp
    assert line["n_gpus"] == 2 and line["scaling"] == scaling
    assert line["config"]["global_batch"] == (64 if scaling == "weak" else 32)
