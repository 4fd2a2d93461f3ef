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


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_train_cli_with_ragged_last_batch(tmp_path):
    """scripts/train.py under torchrun on a directory dataset whose size is not a multiple of the batch: the short last
    batch of each pass is skipped on every rank alike, so the ranks stay in lock-step and the run ends (no rank left
    waiting in an all-reduce), rank 0 alone writes the checkpoint and the logs."""
    import cv2
    import numpy as np
    rng = np.random.default_rng(0)
    for split, n in (("train", 11), ("val", 6)):                    # batch 4: 2 full batches + 3 / 1 full batch + 2
        fd = tmp_path / "dataset" / "train" / f"{split}_frames" / "image"
        md = tmp_path / "dataset" / "train" / f"{split}_masks" / "image"
        fd.mkdir(parents=True); md.mkdir(parents=True)
        for i in range(n):
            img = rng.integers(0, 255, (80, 100, 3), dtype=np.uint8)
            mask = np.zeros((80, 100), np.uint8); mask[20:60, 30:80] = 255
            img[20:60, 30:80] //= 4
            cv2.imwrite(str(fd / f"{i}.jpg"), img); cv2.imwrite(str(md / f"{i}.png"), mask)
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "scripts", "train.py"), "--epochs", "3", "--batch-size", "4",
           "--model-out", "./models/m.h5", "--image-size", "64"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "Steps per epoch: 2, Validation steps: 1" in r.stdout
    assert r.stdout.count("--- Training complete ---") == 2          # both ranks reached the end
    assert os.path.isfile(tmp_path / "models" / "m.h5")
    assert len(os.listdir(tmp_path / "logs")) == 1                    # rank 0 only
    # a batch size the ranks cannot share equally is refused up front
    cmd[cmd.index("--batch-size") + 1] = "3"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=str(tmp_path))
    assert r.returncode != 0 and "must be a positive multiple of the number of GPUs" in r.stdout
