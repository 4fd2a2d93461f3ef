"""The N>1 path on CPU: two gloo ranks exchange the flat gradient buffer region by region (dist.GradSync), exactly as
the NCCL ranks do on GPUs; the result must equal the sum of the per-rank buffers, and BN statistics must average."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from unet_b200 import dist as D
    from unet_b200.spec import UNetSpec
    r, lr, w = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    spec = UNetSpec((32, 32, 3))
    regions = D.grad_regions(spec)
    g = torch.full((spec.n_trainable_flat,), float(rank + 1))
    g[:: 7] = rank * 10.0
    moving = torch.tensor([float(rank), 4.0])               # rides with the first region, averaged (BN moving statistics)
    sync = D.GradSync(g, regions, mean_with_first=moving)
    order = []
    for name in ("decoder", "bottleneck", "encoder"):        # the order backward completes them
        sync.ready(name); order.append(name)
    sync.finish()
    assert moving.tolist() == [0.5, 4.0]
    stats = torch.tensor([float(rank), 2.0 * rank])
    D.average_(stats)
    lo, hi = D.shard_range(10, rank, world)
    out.put((rank, g[:16].tolist(), float(g.sum()), stats.tolist(), sync.bytes_exchanged, (lo, hi)))
    dist.destroy_process_group()


def test_two_rank_gradient_exchange():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from unet_b200.spec import UNetSpec
    n = UNetSpec((32, 32, 3)).n_trainable_flat
    expect = np.full(n, 3.0); expect[::7] = 10.0
    for rank, head, total, stats, nbytes, span in res:
        np.testing.assert_allclose(head, expect[:16])
        assert total == pytest.approx(expect.sum())
        assert stats == [0.5, 1.0]
        assert nbytes == n * 4 + 8                          # every parameter (and the 2 statistics) crossed the wire exactly once
    assert [r[5] for r in res] == [(0, 5), (5, 10)]


def test_regions_tile_the_flat_buffer():
    from unet_b200.dist import grad_regions
    from unet_b200.spec import UNetSpec
    spec = UNetSpec((64, 64, 3), num_classes=8)
    reg = sorted((lo, hi, name) for name, lo, hi in grad_regions(spec))
    assert reg[0][0] == 0 and reg[-1][1] == spec.n_trainable_flat
    assert all(reg[i][1] == reg[i + 1][0] for i in range(len(reg) - 1))
    assert [r[2] for r in reg] == ["encoder", "bottleneck", "decoder"]
    p = spec.params
    assert reg[2][0] <= p["output_mask/bias"].offset < reg[2][1] and reg[1][0] <= p["bneck_block2_bn/beta"].offset < reg[1][1]
