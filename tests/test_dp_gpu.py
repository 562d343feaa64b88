"""Hardware evidence that the N-GPU data-parallel step == the 1-GPU step at N x batch: two
ranks over NCCL, launched the way the driver launches bench.py.  Skipped on a box with fewer than two GPUs (run it
with `gpurun --gpus 2`)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_step_equals_one_gpu_step_at_twice_the_batch(tmp_path):
    out = tmp_path / "dp.json"
    env = dict(os.environ, EMO_DP_OUT=str(out))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "dp_gpu_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.load(open(out))
    for tag, tol in (("fp32", 1e-3), ("bf16", 3e-2)):
        d = res[tag]
        assert d["count"][0] == d["count"][1]                                   # global non-pad count
        assert abs(d["loss_sum"][0] - d["loss_sum"][1]) < tol * abs(d["loss_sum"][1])
        assert d["grad_rms_rel"] < tol, (tag, d)                                # all-reduced gradient == 1-GPU gradient
        assert d["replica_checksums"][0] == d["replica_checksums"][1]           # replicas stay identical
    # the first Adam step moves every weight by ~lr * sign(g): only elements whose gradient is below the summation
    # noise may differ, so in fp32 the two updates agree to a fraction of a per cent of the update's own size
    assert res["fp32"]["param_diff_rel_to_update"] < 5e-2, res["fp32"]
