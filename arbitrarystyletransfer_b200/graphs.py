"""CUDA-graph capture of a whole step.  At training batch sizes (8 x 256 x 256) a step is ~250
kernel launches of tens of microseconds each: launch latency, not the GPU, bounds steps/s when the
step is driven call by call from Python.  Every kernel of this package is enqueued on the current
stream with caller-owned memory and never synchronises, so an entire step -- forward, backward,
gradient clipping, optimiser update -- can be captured once and replayed with one launch.
"""
from __future__ import annotations

from typing import Callable

import torch


class GraphedStep:
    """``step = GraphedStep(fn, static_inputs)``; ``step(*new_inputs)`` copies the new inputs into the
    static tensors, replays the captured graph and returns ``fn``'s (static) outputs.

    ``fn`` must be shape-static and must not synchronise or touch the host (no ``.item()``).
    Optimisers must be built with ``capturable=True``.  Build the GraphedStep BEFORE running ``fn``'s
    backward eagerly on the default stream: autograd ties each parameter's gradient accumulation to
    the stream of its first backward, and the legacy default stream cannot join a capture.  (The
    warm-up steps here run on a side stream for that reason.)"""

    def __init__(self, fn: Callable, static_inputs, warmup: int = 3):
        self.static_inputs = list(static_inputs)
        dev = self.static_inputs[0].device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):          # warm-up off the default stream, as capture requires
            for _ in range(warmup):
                fn(*self.static_inputs)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = fn(*self.static_inputs)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_inputs, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.outputs
