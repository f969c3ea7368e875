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
            dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def gather_batch(local_out: torch.Tensor, batch: int, group=None, dst: int | None = None):
    """Assembles the full (batch, ...) tensor from the per-rank shards made by `shard_bounds`.  dst=None: every rank receives it
    (all-gather); dst=r: only rank r does (what nn.DataParallel's gather to device 0 does, main_us3d.py:100) and the other
    ranks return None -- 1/N of the all-gather's traffic per link.  A rank whose shard is empty (batch < world) passes a
    zero-length `local_out` of the right trailing shape."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_out
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(batch, r, world)[1] - shard_bounds(batch, r, world)[0] for r in range(world)]
    tail = tuple(local_out.shape[1:])
    if local_out.shape[0] != sizes[rank]:
        raise ValueError(f"gather_batch: rank {rank} holds {local_out.shape[0]} samples, its shard of a batch of {batch} has {sizes[rank]}")
    even = len(set(sizes)) == 1
    smax = max(sizes)
    if even:
        send = local_out.contiguous()
    else:       # ragged split: pad every shard to the largest one, gather, trim (collectives need equal sizes)
        send = local_out.new_zeros((smax,) + tail)
        send[: sizes[rank]] = local_out
    if dst is None:
        out = local_out.new_empty((world * smax,) + tail)
        dist.all_gather_into_tensor(out, send, group=group)
    else:
        out = local_out.new_empty((world * smax,) + tail) if rank == dst else None
        dist.gather(send, list(out.split(smax)) if rank == dst else None, dst=dist.get_global_rank(group, dst) if group is not None else dst,
                    group=group)
        if rank != dst:
            return None
    return out if even else torch.cat([out[r * smax: r * smax + sizes[r]] for r in range(world)], 0)


class ShardedHotPath:
    """Runs `path` (a DisparityHotPath on this rank's device) on this rank's slice of a global batch and gathers the
    full-resolution disparity.  `inputs` may be the global batch (it is sliced here) or already the local shard."""

    def __init__(self, path, group=None, dst: int | None = None):
        """dst: None = all-gather (every rank gets the result); r = gather to rank r only (DataParallel semantics)."""
        self.path, self.group, self.dst = path, group, dst
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def __call__(self, inputs: dict, global_batch: int | None = None, presharded: bool = False):
        order = ("f8_l", "f8_r", "f4_l", "f4_r", "cf_l", "cf_r", "spx_pred", "pred_label")
        if not presharded:
            global_batch = inputs[order[0]].shape[0]
            inputs = {k: shard(inputs[k], self.rank, self.world) for k in order if inputs.get(k) is not None}
        elif global_batch is None:
            raise ValueError("ShardedHotPath: presharded inputs need global_batch")
        spx = inputs["spx_pred"]
        if spx.shape[0] == 0:       # batch < world: this rank has nothing to compute but still takes part in the collective
            local = spx.new_empty((0,) + tuple(spx.shape[-2:]))
        else:
            key = "pred_att_up" if self.path.att_weights_only else "pred_up"
            local = self.path(*[inputs.get(k) for k in order])[key]
        return gather_batch(local, global_batch, self.group, self.dst)


# ---------------------------------------------------------------------------------------------------------------------
# Large single images: row tiles with halos, tiles spread over the ranks ("large single images")
# ---------------------------------------------------------------------------------------------------------------------
TILE_UNIT = 128          # tiles are whole 128-row units with origins on the window-attention grid (128 px at 1/32 resolution), so a
                         # tile sees the same (unpadded) windows as the whole image
EXACT_HALO = 384         # >= the vertical receptive field of the whole path (~380 px)


def plan_row_tiles(H: int, n_tiles: int, halo: int = EXACT_HALO):
    """Splits H rows into `n_tiles` bands of whole 128-row units (as even as possible) and extends each band by `halo` rows on
    the sides that are not image borders.  Epipolar lines are rows, so bands need no disparity halo, only the receptive-field
    one.  Returns [(ext_lo, ext_hi, keep_lo, keep_hi)], all multiples of 128."""
    if H % TILE_UNIT or halo % TILE_UNIT or halo < 0:
        raise ValueError("plan_row_tiles: H and halo must be multiples of 128")
    units = H // TILE_UNIT
    if not (1 <= n_tiles <= units):
        raise ValueError(f"plan_row_tiles: need 1 <= n_tiles <= H/128 = {units}")
    base, extra = divmod(units, n_tiles)
    tiles, lo = [], 0
    for t in range(n_tiles):
        hi = lo + (base + (1 if t < extra else 0)) * TILE_UNIT
        tiles.append((max(0, lo - halo), min(H, hi + halo), lo, hi))
        lo = hi
    return tiles


class TiledHotPath:
    """Runs `path` on row (or column) bands of one large image batch and stitches the kept rows.  With a process group the bands are dealt
    round-robin to the ranks and the stitched disparity is summed across ranks (every row is written by exactly one rank, the
    others contribute zeros, so the sum is exact).  With halo >= EXACT_HALO no kept row can see a cut; what remains is the fp32
    rounding of the reference's own grid normalisation (SpatialTransformer_grid divides by (H-1)/2 and multiplies back,
    submodule.py:279-280, so its sampling rows depend on the tile height at the 1e-6 level): <= 1e-4 px in fp32 mode
    (measured, tests/test_gpu_hotpath.py).  A smaller halo trades seam accuracy for less recomputation.  The reference has no
    tiling at all (main_us3d.py feeds whole crops)."""

    def __init__(self, path, n_tiles: int, halo: int | None = None, group=None, axis: str = "h"):
        """axis "h": row bands (halo = receptive field only).  axis "w": column bands; the halo must additionally cover the
        disparity range on the side(s) the right image is read from (signed models: both), so the default is the receptive field
        plus maxdisp rounded up to the 128-px grid (the same halo is used on both sides)."""
        if axis not in ("h", "w"):
            raise ValueError("TiledHotPath: axis must be 'h' or 'w'")
        if halo is None:
            halo = EXACT_HALO if axis == "h" else -(-(EXACT_HALO + int(path.maxdisp)) // TILE_UNIT) * TILE_UNIT
        self.path, self.n_tiles, self.halo, self.group, self.axis = path, n_tiles, halo, group, axis
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def __call__(self, inputs: dict):
        order = ("f8_l", "f8_r", "f4_l", "f4_r", "cf_l", "cf_r", "spx_pred", "pred_label")
        dim = -2 if self.axis == "h" else -1
        L = inputs["spx_pred"].shape[dim]
        key = "pred_att_up" if self.path.att_weights_only else "pred_up"
        out = None
        for t, (e0, e1, k0, k1) in enumerate(plan_row_tiles(L, self.n_tiles, self.halo)):
            if t % self.world != self.rank:
                continue
            args = []
            for k in order:
                x = inputs.get(k)
                if x is not None:
                    s = L // x.shape[dim]                # 8, 4 or 1
                    x = x.narrow(dim, e0 // s, (e1 - e0) // s).contiguous()
                args.append(x)
            o = self.path(*args)[key]
            if out is None:
                full = list(o.shape)
                full[dim] = L
                out = o.new_zeros(full)
            out.narrow(dim, k0, k1 - k0).copy_(o.narrow(dim, k0 - e0, k1 - k0))
        if out is None:                                    # more ranks than tiles
            ref = inputs["spx_pred"]
            out = ref.new_zeros((ref.shape[0], ref.shape[-2], ref.shape[-1]))
        if self.world > 1:
            dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
        return out


def bind_to_gpu_numa(local_rank: int):
    """Pins this process to the CPU cores NVML reports as local to GPU `local_rank` (best effort; returns the core list or None).
    Call it BEFORE allocating pinned host buffers: page-locked staging memory is then first-touched on the GPU's own NUMA node,
    so the H2D copies of the ranks of one box do not all cross the same inter-socket link (e2e with 8 ranks is host-memory bound)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[local_rank])
                                              if os.environ.get("CUDA_VISIBLE_DEVICES") else local_rank)
        ncpu = os.cpu_count() or 1
        words = (ncpu + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [i * 64 + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1 and i * 64 + b < ncpu]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:          # no NVML / no permission / exotic topology: staging still works, just not NUMA-local
        return None
