"""CUDA-graph execution of the forward / backward plans.

A batch-4 training step issues ~900 kernel launches; driven launch-by-launch from Python it is host-bound
("CUDA Graphs for small-batch latency").  `GraphedPlan` captures a plan function once per input
shape on static buffers and replays it; weight re-packing is captured too (it reads the parameters at their fixed
addresses, which the optimizer updates in place), so one replay per step stays correct while training.
"""
from __future__ import annotations

import os
from typing import Callable, List, Sequence

import torch

from . import ops


def enabled() -> bool:
    return os.environ.get("FNST_CUDA_GRAPH", "1") != "0"


class GraphedPlan:
    """fn(*static_inputs) -> (tensor | tuple of tensors | dict...) captured as one CUDA graph."""

    def __init__(self, fn: Callable, example_inputs: Sequence[torch.Tensor], warmup: int = 2, pool=None, device=None):
        """example_inputs may be empty (a plan that reads and writes fixed addresses only); `device` is then required."""
        dev = example_inputs[0].device if example_inputs else torch.device(device)
        self.static_inputs: List[torch.Tensor] = [t.detach().clone() for t in example_inputs]
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                fn(*self.static_inputs)
        cur.wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.launch_count
        with torch.no_grad(), torch.cuda.graph(self.graph, pool=pool, capture_error_mode="thread_local"):
            self.outputs = fn(*self.static_inputs)
        self.launches = ops.launch_count - n0

    def pool(self):
        return self.graph.pool()

    def __call__(self, *inputs: torch.Tensor):
        for s, x in zip(self.static_inputs, inputs):
            if s.data_ptr() != x.data_ptr():
                s.copy_(x)
        self.graph.replay()
        ops.launch_count += self.launches
        return self.outputs
