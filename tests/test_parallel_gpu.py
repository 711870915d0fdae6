"""The search loop's collectives on hardware (/root/reference pix2latent/optimizer/base_cma_optimizer.py:82-87
ask -> broadcast, :140 tell after the gather): BasinCMAOptimizer with the native BigGAN / loss under torchrun with
backend "nccl", one rank per GPU, against the same run in ONE process. Candidates are independent and the step is
bitwise reproducible, so the sharded run must equal the unsharded one BIT FOR BIT (losses, latents, CMA mean).
Needs >= 2 GPUs (gpurun --gpus 2); skipped on a single-GPU box."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(nproc, out, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", port, os.path.join(HERE, "_dist_worker.py"), out, "native"]
    # same host thread count in both runs: the synthetic generator's BN statistics are calibrated by a CPU forward pass
    # (oracle/biggan.py), whose fp32 summation order follows the OpenMP thread count — torchrun sets 1 only for nproc > 1
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, OMP_NUM_THREADS="4"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return np.load(out)


def test_basincma_nccl_two_ranks_equals_single_process_bitwise(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (NCCL)")
    one = _run(1, str(tmp_path / "one.npz"), "29631")
    two = _run(2, str(tmp_path / "two.npz"), "29632")
    assert int(one["world"]) == 1 and int(two["world"]) == 2
    assert int(two["fused_calls"]) == 3          # the device-resident loop ran on every rank's shard
    for k in ("asked", "told", "loss", "z", "c", "mean"):   # in the order things happen: the first mismatch names the culprit
        assert np.array_equal(one[k], two[k]), "%s differs between the sharded and the unsharded run (max |d| %.3e)" % (
            k, np.abs(one[k] - two[k]).max())
