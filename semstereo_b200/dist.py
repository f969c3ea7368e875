"""Multi-GPU execution of the hot path: one process per GPU, stereo pairs sharded across ranks, outputs gathered.

The path has no cross-sample operation (eval BatchNorm is affine), so batch sharding is exact:
rank r owns samples [r*B/N, (r+1)*B/N) and the N-GPU result equals the 1-GPU result bit for bit.  The only collective is
the gather of the disparity maps — what nn.DataParallel's implicit gather does in the reference (main_us3d.py:100,
test_us3d.py:58) — issued as one NCCL all_gather on the compute stream.  Works on the gloo backend too (CPU tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int):
    """Contiguous, balanced split: the first `batch % world` ranks get one extra sample (torch.chunk-like scatter of
    nn.DataParallel differs only for batches not divisible by world; divisible batches are identical)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi].contiguous()


def init_from_env(backend: str | None = None):
    """Initialises torch.distributed from the torchrun environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def gather_batch(local_out: torch.Tensor, batch: int, group=None) -> torch.Tensor:
    """All ranks receive the full (batch, ...) tensor assembled from per-rank shards made by `shard_bounds`."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_out
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(batch, r, world)[1] - shard_bounds(batch, r, world)[0] for r in range(world)]
    if len(set(sizes)) == 1:
        out = local_out.new_empty((batch,) + tuple(local_out.shape[1:]))
        dist.all_gather_into_tensor(out, local_out.contiguous(), group=group)
        return out
    # ragged split: pad every shard to the largest one, gather, trim (collectives need equal sizes)
    smax = max(sizes)
    padded = local_out.new_zeros((smax,) + tuple(local_out.shape[1:]))
    padded[: sizes[rank]] = local_out
    out = local_out.new_empty((world * smax,) + tuple(local_out.shape[1:]))
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * smax: r * smax + sizes[r]] for r in range(world)], 0)


class ShardedHotPath:
    """Runs `path` (a DisparityHotPath on this rank's device) on this rank's slice of a global batch and gathers the
    full-resolution disparity.  `inputs` may be the global batch (it is sliced here) or already the local shard."""

    def __init__(self, path, group=None):
        self.path, self.group = path, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def __call__(self, inputs: dict, global_batch: int | None = None, presharded: bool = False):
        order = ("f8_l", "f8_r", "f4_l", "f4_r", "cf_l", "cf_r", "spx_pred", "pred_label")
        if not presharded:
            global_batch = inputs[order[0]].shape[0]
            inputs = {k: shard(inputs[k], self.rank, self.world) for k in order if inputs.get(k) is not None}
        out = self.path(*[inputs.get(k) for k in order])
        key = "pred_att_up" if self.path.att_weights_only else "pred_up"
        return gather_batch(out[key], global_batch, self.group)
