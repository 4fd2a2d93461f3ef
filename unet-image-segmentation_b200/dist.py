"""Data-parallel plumbing: one process per GPU, `torch.distributed` (NCCL over NVLink 5 / NVSwitch) for the exchange.

The reference is single-device (scripts/train.py:119-130 only sets memory growth); this is the partitioning the
north star adds: the batch is split across ranks, every rank runs the same engine on its shard, and ONE exchange step
per iteration sums the flat gradient buffer.  The buffer is exchanged in three contiguous regions, each launched on a
side stream as soon as backward has finished writing it (decoder+head first, then bottleneck, then encoder), so
the all-reduce of the 14 MB decoder region overlaps the rest of backward.  BatchNormalization is per replica, as in
Keras under tf.distribute; moving statistics are averaged.  Inference needs no communication.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when launched plainly."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device(f"cuda:{local_rank}"))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def shutdown(engine=None, timeout_s: float = 20.0) -> bool:
    """Tear the process group down at the end of a run.  CUDA graphs that captured NCCL kernels must be released first
    (communicator teardown otherwise waits on them forever), and the teardown itself is bounded: returns False when
    destroy_process_group() did not come back within `timeout_s` (the caller should then leave with os._exit)."""
    if engine is not None:
        engine.release_plans()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if not dist.is_initialized():
        return True
    import threading
    t = threading.Thread(target=dist.destroy_process_group, daemon=True)
    t.start()
    t.join(timeout_s)
    return not t.is_alive()


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of `n_items` units for `rank`; earlier ranks take the remainder."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def grad_regions(spec) -> List[Tuple[str, int, int]]:
    """Contiguous regions of the flat trainable buffer in the order backward completes them."""
    dec0 = spec.params["dec4_upsample/kernel"].offset
    bn0 = spec.params["bneck_block1_sepconv/depthwise_kernel"].offset
    return [("decoder", dec0, spec.n_trainable_flat), ("bottleneck", bn0, dec0), ("encoder", 0, bn0)]


class GradSync:
    """Sums `flat` across ranks region by region.  `ready(name)` is called by the engine when backward has finished
    a region; `finish()` makes the compute stream wait for all exchanges.  `mean_with_first`: a tensor averaged across ranks
    together with the first region (the BatchNormalization moving statistics: final once the forward pass is over, so their
    exchange hides behind backward instead of being a second, serialised collective at the end of the step).
    Fork/join structure only (events and stream waits, no host blocking), so a whole training step including its
    collectives can be captured into a CUDA graph.  Works on CPU tensors (gloo) for tests."""

    def __init__(self, flat: torch.Tensor, regions: List[Tuple[str, int, int]], group=None,
                 mean_with_first: Optional[torch.Tensor] = None):
        self.flat, self.group = flat, group
        self.regions = {name: (lo, hi) for name, lo, hi in regions}
        self.first = regions[0][0]
        self.mean_with_first = mean_with_first
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.cuda = flat.is_cuda
        self.comm_stream = torch.cuda.Stream(device=flat.device) if self.cuda and self.world > 1 else None
        self.bytes_exchanged = 0

    def _exchange(self, name: str) -> None:
        lo, hi = self.regions[name]
        view = self.flat[lo:hi]
        self.bytes_exchanged += view.numel() * view.element_size()
        dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
        t = self.mean_with_first
        if t is not None and name == self.first:
            self.bytes_exchanged += t.numel() * t.element_size()
            if self.cuda:
                dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)        # ncclAvg: no separate scaling kernel
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
                t.div_(self.world)

    def ready(self, name: str) -> None:
        if self.world == 1:
            return
        if not self.cuda:
            self._exchange(name)
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ev)          # fork: the collective starts when backward has written the region
            self._exchange(name)                     # NCCL runs on its own stream; comm_stream waits for it (no host block)

    def finish(self) -> None:
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)              # join


def average_(t: torch.Tensor, group=None) -> None:
    """In-place mean across ranks (BatchNormalization moving statistics; scalar metrics)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.div_(dist.get_world_size(group))
