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


# ---------------------------------------------------------------------------------------------------------------------
# Large single images: row tiles with halos, tiles spread over the ranks ("large single images")
# ---------------------------------------------------------------------------------------------------------------------
TILE_UNIT = 128          # the path needs H % 128 == 0 and tile origins on the window-attention grid (128 px at 1/32 resolution)
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
