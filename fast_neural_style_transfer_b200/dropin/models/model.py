"""Drop-in for the reference's models/model.py: same class names, constructor signatures,
child-module names and 58 state-dict keys (models/model.py:7-90), forward executed by libfnst
(hand-written sm_100a kernels) instead of ATen/cuDNN.

The nn.Conv2d / nn.ConvTranspose2d / nn.InstanceNorm2d children are parameter containers only
(identical default initialisation and construction order as the reference, so the same seed gives
the same weights); their own forward is never called.  There is no CPU or library fallback:
non-CUDA inputs raise.

Precision: `net.precision` in {"fp16" (default; tcgen05 tensor cores), "bf16", "fp32" (CUDA
cores, 1e-4 parity path)}, or environment variable FNST_PRECISION.
"""
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

_PKG_PARENT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _PKG_PARENT not in sys.path:
    sys.path.append(_PKG_PARENT)

from fast_neural_style_transfer_b200 import engine, graphs, ops   # noqa: E402
from fast_neural_style_transfer_b200 import autograd_fns      # noqa: E402


def _standalone(name):
    raise RuntimeError(f"{name}.forward is not a separate operator in the B200 path; call StyleTransferNet.forward "
                       "(conv + InstanceNorm + ReLU are fused across these module boundaries)")


class UpsampleConv(nn.Module):
    def __init__(self, in_ch: int, out_ch: int, kernel: int, scale: int = 2):
        super().__init__()
        self.scale = scale
        self.upsample_conv = nn.ConvTranspose2d(in_ch, out_ch, kernel_size=kernel, stride=scale, padding=kernel // 2,
                                                output_padding=scale - 1)

    def forward(self, x):
        _standalone("UpsampleConv")


class ConvLayer(nn.Module):
    def __init__(self, in_ch: int, out_ch: int, kernel: int, stride=1):
        super().__init__()
        self.reflection_pad = nn.ReflectionPad2d(kernel // 2)
        self.conv = nn.Conv2d(in_ch, out_ch, kernel_size=kernel, stride=stride)

    def forward(self, x):
        _standalone("ConvLayer")


class ResidualBlock(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv1 = ConvLayer(channels, channels, kernel=3)
        self.in1 = nn.InstanceNorm2d(channels, affine=True)
        self.conv2 = ConvLayer(channels, channels, kernel=3)
        self.in2 = nn.InstanceNorm2d(channels, affine=True)
        self.dropout = nn.Dropout2d(0.1)

    def forward(self, x):
        _standalone("ResidualBlock")


class StyleTransferNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = ConvLayer(3, 64, kernel=9, stride=2)
        self.norm1 = nn.InstanceNorm2d(64, affine=True)
        self.conv2 = ConvLayer(64, 256, kernel=3, stride=2)
        self.norm2 = nn.InstanceNorm2d(256, affine=True)
        self.res_blocks = nn.ModuleList([ResidualBlock(256) for _ in range(5)])
        self.up1 = UpsampleConv(256, 64, kernel=3, scale=2)
        self.norm3 = nn.InstanceNorm2d(64, affine=True)
        self.up2 = UpsampleConv(64, 32, kernel=3, scale=2)
        self.norm4 = nn.InstanceNorm2d(32, affine=True)
        self.final_conv = ConvLayer(32, 3, kernel=9, stride=1)
        self.precision = os.environ.get("FNST_PRECISION", "fp16")

    # plan cache (packed weights) is derived state: never pickled, rebuilt when parameters change
    def __getstate__(self):
        state = self.__dict__.copy()
        for k in ("_plan_cache", "_graphs", "_train_graphs"):
            state.pop(k, None)
        return state

    def _plan(self) -> "engine.StyleNetPlan":
        params = dict(self.named_parameters())
        key = (self.precision, tuple((p.data_ptr(), p._version) for p in params.values()))
        cache = self.__dict__.get("_plan_cache")
        if cache is None or cache[0] != key:
            cache = (key, engine.StyleNetPlan(self.precision).pack(params))
            self.__dict__["_plan_cache"] = cache
        return cache[1]

    def _dropout_scales(self, x, out=None):
        """One (B,256) Dropout2d scale per residual block as a (5,B,256) tensor, drawn from torch's RNG in block order with
        the same draw the reference makes per block (models/model.py:84,88; SURVEY 8c stochasticity).  `out`: draw into
        this buffer (the captured training graph's own input: no copy)."""
        if not self.training:
            return None
        # Dropout2d on a (B,C,H,W) input draws noise = empty(B,C,1,1).bernoulli_(1-p).div_(1-p) (ATen feature_dropout): the same
        # five draws, in block order, written straight into one (5,B,256) buffer -- 5 Philox launches + 1 scale instead of 5 x
        # (ones, bernoulli, div, mul) + 5 copies into the captured graph's inputs
        keep = [1.0 - blk.dropout.p for blk in self.res_blocks]
        buf = out if out is not None else torch.empty((len(keep), x.shape[0], 256), dtype=torch.float32, device=x.device)
        for i, q in enumerate(keep):
            buf[i].view(x.shape[0], 256, 1, 1).bernoulli_(q)
        if len(set(keep)) == 1:
            buf.div_(keep[0])
        else:
            for i, q in enumerate(keep):
                buf[i].div_(q)
        return buf

    # Small no-grad forwards are launch-bound (59 launches): replay them as one CUDA graph per input shape.
    GRAPH_MAX_PIXELS = 4 * 1080 * 1920

    def _graph_forward(self, plan, x):
        graphs = self.__dict__.setdefault("_graphs", {})
        key = (id(plan), tuple(x.shape), x.device.index)
        entry = graphs.get(key)
        if entry is None:
            for k in [k for k in graphs if k[0] != id(plan)] if len(graphs) < 8 else list(graphs):
                del graphs[k]                                      # weights changed (or too many shapes): drop captures
            static_x = x.detach().clone().float().contiguous()
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            with torch.cuda.stream(side):
                for _ in range(2):
                    plan.forward(static_x)
            torch.cuda.current_stream(x.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            n0 = ops.launch_count
            with torch.cuda.graph(graph):
                static_y = plan.forward(static_x)
            entry = graphs[key] = (graph, static_x, static_y, ops.launch_count - n0)
        graph, static_x, static_y, n_launches = entry
        static_x.copy_(x)
        graph.replay()
        ops.launch_count += n_launches           # kernels replayed from the captured graph
        return static_y.clone()

    IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)

    @torch.no_grad()
    def stylize_uint8(self, img_u8: torch.Tensor, normalize_input: bool = False) -> torch.Tensor:
        """Extension beyond the reference API: uint8 HWC in, uint8 HWC out, pre/post-processing on the GPU.
        Equivalent to inference.py:44-60: ToTensor (optionally ImageNet-normalised, as the training monitor does),
        forward, de-normalise, clamp to [0,1], x255 -- with 4x fewer PCIe bytes than float tensors."""
        if not img_u8.is_cuda or img_u8.dtype != torch.uint8 or img_u8.dim() != 4 or img_u8.shape[-1] != 3:
            raise RuntimeError("stylize_uint8 expects a CUDA uint8 tensor of shape (n, h, w, 3)")
        mean, std = (self.IMAGENET_MEAN, self.IMAGENET_STD) if normalize_input else ((0.0,) * 3, (1.0,) * 3)
        x = ops.u8_to_nchw(img_u8.contiguous(), mean, std)
        y = self.forward(x)
        return ops.nchw_to_u8(y.contiguous(), self.IMAGENET_MEAN, self.IMAGENET_STD)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("StyleTransferNet (B200 drop-in) needs CUDA tensors: there is no CPU fallback")
        params = list(self.parameters())
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            if graphs.enabled() and not torch.cuda.is_current_stream_capturing():
                # training step as two CUDA-graph replays (forward incl. weight re-pack, backward) per input shape
                cache = self.__dict__.setdefault("_train_graphs", {})
                # (every parameter address is part of the key: the captured graphs read the weights in place, so a parameter
                #  whose storage was replaced -- `.to()`, `p.data = ...` -- must not hit a stale capture)
                key = (self.precision, tuple(x.shape), x.device.index, self.training, tuple(p.data_ptr() for p in params))
                state = cache.get(key)
                busy = state is not None and state.in_flight.busy()
                # the Dropout2d scales are drawn straight into the captured graph's input buffer (unless that graph is in flight)
                drops = self._dropout_scales(x, out=state.drop_input() if state is not None and not busy else None)
                if state is None:
                    if len(cache) >= 4:
                        cache.clear()
                    state = cache[key] = autograd_fns.StyleNetTrainGraph(dict(self.named_parameters()), self.precision, x, drops)
                if not busy:
                    return autograd_fns.stylenet_graphed_apply(state, x, drops, params)
                # an earlier forward of this graph still waits for its backward (gradient accumulation, two losses on two
                # inputs): the captured tape holds ONE forward, so this call takes the eager per-call-tape path below
            else:
                drops = self._dropout_scales(x)
            names = [n for n, _ in self.named_parameters()]
            return autograd_fns.stylenet_apply(self._plan(), names, x, drops, params)
        plan = self._plan()
        use_graph = (not self.training and os.environ.get("FNST_CUDA_GRAPH", "1") != "0"
                     and x.shape[0] * x.shape[2] * x.shape[3] <= self.GRAPH_MAX_PIXELS
                     and not torch.cuda.is_current_stream_capturing())
        if use_graph:
            return self._graph_forward(plan, x)
        return plan.forward(x, self._dropout_scales(x))
