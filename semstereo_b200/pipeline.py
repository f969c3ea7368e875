"""HostPipeline — the end-to-end entry point for callers whose stereo-pair features live in HOST memory.

Every batch is copied host->device on a dedicated copy stream into one of `depth` device staging sets, the hot path runs on
the compute stream, and the disparity is copied device->host into a pinned result buffer.  Stage i+1's H2D overlaps stage i's
compute (CUDA events order the two streams; nothing synchronises the host until a result is consumed).  This is what replaces
the reference's `.cuda()` copies + forward + `.cpu()` in `test_us3d.py:95-102`.
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator

import torch

ORDER = ("f8_l", "f8_r", "f4_l", "f4_r", "cf_l", "cf_r", "spx_pred", "pred_label")


class HostPipeline:
    def __init__(self, path, depth: int = 2, post=None, keys=ORDER, call=None):
        """path: a DisparityHotPath (or, with `keys` / `call`, any module of this package, e.g. decoder.StereoHead) on a CUDA
        device.  keys: the batch-dict entries staged on the device; call(staged_dict) -> device result tensor (default: the hot
        path's full-resolution disparity).  post(device_result) -> device tensor to return (e.g. an all-gather).
        Host tensors may be fp32 or bf16 (per entry): bf16 entries cross PCIe at half the bytes and are widened to fp32 on the
        device by ss_widen_bf16 -- lossless for features that are bf16-valued to begin with (e.g. produced by a bf16 decoder), a
        rounding of the INPUTS otherwise (the caller's choice; the path itself computes the same either way)."""
        self.path, self.depth, self.post, self.keys = path, depth, post, tuple(keys)
        self._wide = [None] * depth           # fp32 copies of bf16-staged entries
        self.dev = next(path.parameters()).device
        self.copy_stream = torch.cuda.Stream(self.dev)
        if call is None:
            key = "pred_att_up" if path.att_weights_only else "pred_up"
            call = lambda st: path(*[st.get(k) for k in ORDER])[key]          # noqa: E731
        self.call = call
        self._stage = [None] * depth          # device staging sets
        self._host_out = [None] * depth       # pinned result buffers
        self._copied = [torch.cuda.Event() for _ in range(depth)]
        self._consumed = [torch.cuda.Event() for _ in range(depth)]
        self._done = [torch.cuda.Event() for _ in range(depth)]
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _staging(self, i, batch):
        st = self._stage[i]
        keys = [k for k in self.keys if batch.get(k) is not None]      # cf_l / cf_r are optional (computed on the device if absent)
        if st is None or set(st) != set(keys) or any(st[k].shape != batch[k].shape or st[k].dtype != batch[k].dtype for k in keys):
            for k in keys:
                if batch[k].dtype not in (torch.float32, torch.bfloat16):
                    raise TypeError(f"HostPipeline: entry '{k}' must be float32 or bfloat16, got {batch[k].dtype}")
            st = {k: torch.empty(batch[k].shape, dtype=batch[k].dtype, device=self.dev) for k in keys}
            self._stage[i] = st
            self._wide[i] = {k: torch.empty(batch[k].shape, dtype=torch.float32, device=self.dev) for k in keys if batch[k].dtype == torch.bfloat16}
            # the caching allocator may hand out blocks that kernels already queued on the compute stream still use
            # (stream-ordered reuse): order the first copy into a fresh staging set after that work, once
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.dev))
        return st

    def run(self, batches: Iterable[Dict[str, torch.Tensor]]) -> Iterator[torch.Tensor]:
        """Yields one pinned host tensor (B,H,W) per input batch, in order.  The yielded tensor is reused `depth` batches
        later: copy it if it must outlive that."""
        compute = torch.cuda.current_stream(self.dev)
        pending = []                           # slots whose result has not been yielded yet
        n = 0
        for batch in batches:
            i = n % self.depth
            if len(pending) == self.depth:     # slot i is about to be reused: hand out its result first
                j = pending.pop(0)
                self._done[j].synchronize()
                yield self._host_out[j]
            st = self._staging(i, batch)
            with torch.cuda.stream(self.copy_stream):
                if n >= self.depth:
                    self.copy_stream.wait_event(self._consumed[i])     # compute of the previous tenant has read the staging set
                for k in st:
                    st[k].copy_(batch[k], non_blocking=True)
                self._copied[i].record(self.copy_stream)
            self.h2d_bytes += sum(batch[k].numel() * batch[k].element_size() for k in st)
            compute.wait_event(self._copied[i])
            if self._wide[i]:
                from . import ops_tc
                st = dict(st)
                for k, w in self._wide[i].items():
                    st[k] = ops_tc.widen_bf16(st[k], w)
            out = self.call(st)
            self._consumed[i].record(compute)
            if self.post is not None:
                out = self.post(out)
            if self._host_out[i] is None or self._host_out[i].shape != out.shape:
                self._host_out[i] = torch.empty(out.shape, dtype=out.dtype).pin_memory()
            self._host_out[i].copy_(out, non_blocking=True)
            self.d2h_bytes += out.numel() * out.element_size()
            self._done[i].record(compute)
            pending.append(i)
            n += 1
        for j in pending:
            self._done[j].synchronize()
            yield self._host_out[j]
