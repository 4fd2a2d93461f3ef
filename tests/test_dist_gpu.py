"""Data-parallel correctness on real GPUs over NCCL: needs >= 2 visible GPUs, skipped otherwise
(`gpurun --gpus 2 -- python -m pytest tests/test_dist_gpu.py -m gpu`).  Runs bench.py's --check leg under torchrun:
  * the exchanged gradient == mean of the per-shard single-GPU gradients,
  * after optimizer steps (CUDA-graph replayed, NCCL exchanges inside the graph) every rank holds bit-identical weights and
    BatchNormalization statistics."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("graph_nccl", ["1", "0"])
def test_two_gpu_gradient_exchange_and_weight_identity(graph_nccl):
    env = dict(os.environ, UNET_B200_GRAPH_NCCL=graph_nccl)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + int(graph_nccl)), os.path.join(ROOT, "bench.py"), "--gpus", "2", "--workload", "train256",
           "--steps", "3", "--warmup", "3", "--check", "--no-e2e", "--no-infer", "--no-cpu-baseline"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    chk = line["check"]
    assert chk["grad_ok"], chk
    assert chk["weights_identical_across_ranks_after_4_steps"] and chk["weights_moved"], chk
    assert line["n_gpus"] == 2 and line["value"] > 0
    assert line["config"]["cuda_graph"] == (graph_nccl == "1")
