"""GraphedCall — replay a whole forward (DisparityHotPath / StereoHead) as ONE CUDA graph.

Every C-ABI entry point only enqueues work on the stream it is given (no synchronisation, no allocation, descriptors passed by
value), so a forward can be stream-captured.  At batch 1 (the reference's own inference pattern, test_us3d.py) a pass is ~45 small
launches and the Python/ctypes launch path costs about as much as the kernels of the 1/8-resolution branch; a graph replay
removes that.  Outputs live in the graph's private memory pool and are overwritten by the next replay: copy what must persist.
"""
from __future__ import annotations

from typing import Callable, Dict

import torch

from . import ops


class GraphedCall:
    def __init__(self, fn: Callable[[Dict[str, torch.Tensor]], object], example_inputs: Dict[str, torch.Tensor], warmup: int = 2):
        """fn(inputs_dict) -> tensor or dict of tensors; example_inputs: CUDA tensors fixing shapes/dtypes (None entries allowed)."""
        tensors = [v for v in example_inputs.values() if v is not None]
        if not tensors or not all(t.is_cuda for t in tensors):
            raise RuntimeError("GraphedCall needs CUDA tensors: there is no CPU fallback on this path")
        self.static_in = {k: (None if v is None else v.clone()) for k, v in example_inputs.items()}
        dev = tensors[0].device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # warm-up off the capture: weight packing, kernel attribute setup
            for _ in range(max(warmup, 1)):
                fn(self.static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        prev = ops.record_launches(None)                   # event timing cannot be captured
        try:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.static_out = fn(self.static_in)
        finally:
            ops.record_launches(prev)

    def replay(self):
        """Runs the graph on the inputs currently in `static_in` (write them in place, e.g. as HostPipeline staging buffers)."""
        self.graph.replay()
        return self.static_out

    def __call__(self, inputs: Dict[str, torch.Tensor]):
        for k, dst in self.static_in.items():
            if dst is not None:
                dst.copy_(inputs[k], non_blocking=True)
        return self.replay()
