"""Drop-in for the reference's models/model.py: same class names, constructor signatures,
child-module names and 58 state-dict keys (models/model.py:7-90), forward executed by libfnst
(hand-written sm_100a kernels) instead of ATen/cuDNN.

The nn.Conv2d / nn.ConvTranspose2d / nn.InstanceNorm2d children hold the parameters (identical
default initialisation and construction order as the reference, so the same seed gives the same
weights); their own stock forwards run only while the module is being exported (torch.jit.trace /
torch.onnx.export cannot trace ctypes calls).  There is no CPU or library fallback at run time:
non-CUDA inputs raise.

Precision: `net.precision`, or environment variable FNST_PRECISION:
  "auto" (default)  inference: "fp16x3"; training: "fp16"
  "fp16"            tcgen05 tensor cores, fp16 activations (bf16 gradients in training): 2e-3 on outputs
  "fp16x3"          tcgen05 tensor cores, error-compensated fp16 (hi, lo) pairs: 1e-5 on outputs (the
                    reference's fp32 class); in training an fp32-class forward + bf16 backward
  "bf16"            tcgen05 tensor cores, bf16 activations (1.5e-2 on outputs at random init)
  "fp32"            CUDA cores end to end (4e-6)
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
from fast_neural_style_transfer_b200._lib import PAD_REFLECT    # noqa: E402


def _exporting(x=None) -> bool:
    """torch.jit.trace / torch.onnx.export in progress (model_scripting/torchscript_model.py:25, onnx_version/onnx_model.py:24):
    ctypes calls into libfnst cannot be traced, so during an export -- and only then -- the modules evaluate themselves with the
    stock torch operators of their own parameter-holding children (the traced graph is the reference's graph, 58 shared tensors)."""
    if torch.jit.is_tracing() or torch.onnx.is_in_onnx_export():
        return True
    if x is not None and x.is_cuda:
        return False                                        # run time: nothing below is on the hot path
    # torch.jit.trace(check_trace=True, the default the reference script uses) re-runs the Python module once, eagerly, to
    # compare it with the trace: that call belongs to the export too
    f = sys._getframe(1)
    while f is not None:
        if f.f_code.co_name == "_check_trace" and f.f_globals.get("__name__", "").startswith("torch.jit"):
            return True
        f = f.f_back
    return False


def _need_cuda_nograd(x, params, what):
    if not x.is_cuda:
        raise RuntimeError(f"{what} (B200 drop-in) needs CUDA tensors: there is no CPU fallback")
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)):
        raise RuntimeError(f"{what}.forward on its own is an inference-time operator here; gradients flow through "
                           "StyleTransferNet.forward (whole-network autograd node).  Wrap the call in torch.no_grad().")


def _ceil16(v: int) -> int:
    return (v + 15) // 16 * 16


def _reflect_halo(x_nhwc: torch.Tensor, pad: int) -> torch.Tensor:
    """ReflectionPad2d(pad) on an NHWC tensor (index gather: data movement only)."""
    if pad == 0:
        return x_nhwc
    n, h, w, c = x_nhwc.shape
    if h <= pad or w <= pad:
        raise RuntimeError(f"ReflectionPad2d({pad}) needs H, W > {pad}")
    def idx(extent):
        i = torch.arange(-pad, extent + pad, device=x_nhwc.device).abs()
        return torch.where(i >= extent, 2 * (extent - 1) - i, i)
    return x_nhwc.index_select(1, idx(h)).index_select(2, idx(w)).contiguous()


def _conv_nhwc(xp: torch.Tensor, weight: torch.Tensor, bias, stride: int, stats=None) -> torch.Tensor:
    """'valid' Conv2d of an already padded fp32 NHWC tensor (channels padded to a multiple of 16) as a libfnst gather-GEMM on
    the fp32 path: stride 1 directly, stride 2 through the space-to-depth view (tap (kh,kw) -> spatial offset (kh>>1, kw>>1)
    of phase (kh&1, kw&1))."""
    o, c, k, _ = weight.shape
    n, hp, wp, cp = xp.shape
    ng = _ceil16(o)
    wk = torch.zeros((ng, k, k, cp), dtype=torch.float32, device=xp.device)
    wk[:o, :, :, :c] = weight.detach().float().permute(0, 2, 3, 1)
    b = None
    if bias is not None:
        b = torch.zeros(ng, dtype=torch.float32, device=xp.device)
        b[:o] = bias.detach().float()
    ho, wo = (hp - k) // stride + 1, (wp - k) // stride + 1
    out = torch.empty((n, ho, wo, o), dtype=torch.float32, device=xp.device)
    if stride == 1:
        taps = engine.taps_kxk(k)
        a, a_dims = xp, (n, hp, wp, cp)
    elif stride == 2:
        he, we = (hp + 1) // 2 * 2, (wp + 1) // 2 * 2
        xe = torch.zeros((n, he, we, cp), dtype=torch.float32, device=xp.device)
        xe[:, :hp, :wp] = xp
        a = xe.view(n, he // 2, 2, we // 2, 2, cp).permute(0, 1, 3, 2, 4, 5).reshape(n, he // 2, we // 2, 4 * cp).contiguous()
        a_dims = tuple(a.shape)
        taps = [(kh >> 1, kw >> 1, ((kh & 1) * 2 + (kw & 1)) * cp) for kh in range(k) for kw in range(k)]
    else:
        raise RuntimeError("ConvLayer (B200 drop-in): stride must be 1 or 2")
    spec = ops.ConvSpec(taps, cp, wk.reshape(ng, k * k * cp).contiguous(), ng, o, bias=b)
    ops.conv_gather(spec, a, a_dims, engine._nhwc_strides(a), out, (ho, wo), stats, False)
    return out


class UpsampleConv(nn.Module):
    def __init__(self, in_ch: int, out_ch: int, kernel: int, scale: int = 2):
        super().__init__()
        self.scale = scale
        self.upsample_conv = nn.ConvTranspose2d(in_ch, out_ch, kernel_size=kernel, stride=scale, padding=kernel // 2,
                                                output_padding=scale - 1)

    def forward(self, x):
        """ConvTranspose2d(k=3, s=2, p=1, op=1) (models/model.py:21-22) as the sub-pixel 2x2-tap gather-GEMM with the
        depth-to-space epilogue -- the same operator the fused network forward uses, on the fp32 path, unfused."""
        if _exporting(x):
            return self.upsample_conv(x)
        ct = self.upsample_conv
        _need_cuda_nograd(x, list(self.parameters()), "UpsampleConv")
        if ct.kernel_size != (3, 3) or self.scale != 2:
            raise RuntimeError("UpsampleConv (B200 drop-in): only the reference's kernel=3, scale=2 configuration is implemented")
        cin, cout = ct.weight.shape[0], ct.weight.shape[1]
        if cin % 16 or cout % 4:
            raise RuntimeError("UpsampleConv (B200 drop-in): needs in_ch % 16 == 0 and out_ch % 4 == 0")
        n, _, h, w = x.shape
        a = ops.nchw_to_nhwc(x.detach().float().contiguous(), torch.float32)
        bias4 = ct.bias.detach().float().repeat(4).contiguous() if ct.bias is not None else None
        spec = ops.ConvSpec(engine.TAPS_2X2, cin, engine.pack_conv_transpose(ct.weight.detach().float(), torch.float32), 4 * cout, cout,
                            epilogue=engine.EPI_D2S, bias=bias4)
        out = torch.empty((n, 2 * h, 2 * w, cout), dtype=torch.float32, device=x.device)
        ops.conv_gather(spec, a, (n, h, w, cin), engine._nhwc_strides(a), out, (h, w), None, False)
        return ops.nhwc_to_nchw(out)


class ConvLayer(nn.Module):
    def __init__(self, in_ch: int, out_ch: int, kernel: int, stride=1):
        super().__init__()
        self.reflection_pad = nn.ReflectionPad2d(kernel // 2)
        self.conv = nn.Conv2d(in_ch, out_ch, kernel_size=kernel, stride=stride)

    def forward(self, x):
        """conv(reflection_pad(x)) (models/model.py:74-75) as one libfnst gather-GEMM on the fp32 path, unfused."""
        if _exporting(x):
            return self.conv(self.reflection_pad(x))
        _need_cuda_nograd(x, list(self.parameters()), "ConvLayer")
        conv = self.conv
        k, stride = conv.kernel_size[0], conv.stride[0]
        if k * k > 81 or conv.kernel_size[0] != conv.kernel_size[1]:
            raise RuntimeError("ConvLayer (B200 drop-in): square kernels up to 9x9")
        a = ops.nchw_to_nhwc(x.detach().float().contiguous(), torch.float32, c_pad=_ceil16(x.shape[1]))
        out = _conv_nhwc(_reflect_halo(a, k // 2), conv.weight, conv.bias, stride)
        return ops.nhwc_to_nchw(out)


class ResidualBlock(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv1 = ConvLayer(channels, channels, kernel=3)
        self.in1 = nn.InstanceNorm2d(channels, affine=True)
        self.conv2 = ConvLayer(channels, channels, kernel=3)
        self.in2 = nn.InstanceNorm2d(channels, affine=True)
        self.dropout = nn.Dropout2d(0.1)

    def forward(self, x):
        """x + in2(conv2(dropout(relu(in1(conv1(x)))))) (models/model.py:86-90) on the libfnst operators of the fused forward
        (gather-GEMM with InstanceNorm statistics in its epilogue, inorm_apply writing the next reflect-halo buffer), fp32 path."""
        if _exporting(x):
            y = F.relu(self.in1(self.conv1(x)))
            y = self.dropout(y)
            return x + self.in2(self.conv2(y))
        _need_cuda_nograd(x, list(self.parameters()), "ResidualBlock")
        n, c, h, w = x.shape
        if c % 16 or 256 % (c // 8) or h < 2 or w < 2:
            raise RuntimeError("ResidualBlock (B200 drop-in): channel count must be 16, 32, 64, 128, 256, ... (a divisor pattern of the norm kernel)")
        f32 = torch.float32
        xp = _reflect_halo(ops.nchw_to_nhwc(x.detach().float().contiguous(), f32), 1)
        affine = lambda m: (m.weight.detach().float().contiguous(), m.bias.detach().float().contiguous())
        st = torch.zeros((n, c, 2), dtype=f32, device=x.device)
        raw = _conv_nhwc(xp, self.conv1.conv.weight, None, 1, stats=st)            # the conv bias is cancelled by the InstanceNorm mean
        drop = None
        if self.training:
            drop = torch.empty((n, c, 1, 1), dtype=f32, device=x.device).bernoulli_(1.0 - self.dropout.p).div_(1.0 - self.dropout.p).view(n, c)
        mid = torch.empty((n, h + 2, w + 2, c), dtype=f32, device=x.device)
        g, b = affine(self.in1)
        ops.inorm_apply(raw, st, g, b, mid, relu=True, pad=1, pad_mode=PAD_REFLECT, drop=drop)
        st2 = torch.zeros((n, c, 2), dtype=f32, device=x.device)
        raw2 = _conv_nhwc(mid, self.conv2.conv.weight, None, 1, stats=st2)
        out = torch.empty((n, h, w, c), dtype=f32, device=x.device)
        g, b = affine(self.in2)
        ops.inorm_apply(raw2, st2, g, b, out, relu=False, res=xp, res_pad=1)
        return ops.nhwc_to_nchw(out)


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
        self.precision = os.environ.get("FNST_PRECISION", "auto")

    # plan cache (packed weights) is derived state: never pickled, rebuilt when parameters change
    def __getstate__(self):
        state = self.__dict__.copy()
        for k in ("_plan_cache", "_graphs", "_train_graphs", "_named_cache", "_host_streams", "_fnst_bucket_hook"):
            state.pop(k, None)
        return state

    def _named(self):
        """dict(self.named_parameters()), cached: the module-tree walk costs ~0.1 ms per call and the forward needs it several
        times per step.  Parameter OBJECTS are stable under .to() / load_state_dict() (nn.Module swaps `.data`); anything
        that replaces them (register_parameter, parametrizations) changes `_parameters` of some child, which the cheap check
        below -- object identity of every cached parameter in its owner's dict -- catches."""
        cache = self.__dict__.get("_named_cache")
        if cache is not None:
            owners, named = cache
            if all(owner.get(attr) is p for owner, attr, p in owners):
                return named
        named = dict(self.named_parameters())
        owners = []
        for name, p in named.items():
            mod_path, _, attr = name.rpartition(".")
            owners.append((self.get_submodule(mod_path)._parameters, attr, p))
        self.__dict__["_named_cache"] = (owners, named)
        return named

    def _resolved_precision(self, need_grad: bool) -> str:
        """"auto" (default): inference in the reference's own fp32 class (fp16x3: error-compensated fp16 pairs on tensor cores,
        1e-5 relative L2), training on the fast tensor-core path (fp16 forward / bf16 backward: the class of PyTorch's default
        TF32 convolutions, which is what the reference itself trains in on a GPU)."""
        if self.precision != "auto":
            return self.precision
        return "fp16" if need_grad else "fp16x3"

    def _plan(self, need_grad: bool = False) -> "engine.StyleNetPlan":
        params = self._named()
        precision = self._resolved_precision(need_grad)
        key = (precision, tuple((p.data_ptr(), p._version) for p in params.values()))
        cache = self.__dict__.get("_plan_cache")
        if cache is None or cache[0] != key:
            cache = (key, engine.StyleNetPlan(precision).pack(params))
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

    def to_reference_module(self) -> "StyleTransferNet":
        """The module itself: its children are stock nn.Conv2d / ConvTranspose2d / InstanceNorm2d / ReflectionPad2d objects
        holding the 58 tensors, and under torch.jit.trace / torch.onnx.export every forward in this file evaluates through
        them (see _exporting) -- so the reference's export scripts work on the drop-in unchanged."""
        return self

    def _stock_forward(self, x):
        # models/model.py:49-65 on the children's own stock forwards (export only)
        h = F.relu(self.norm1(self.conv1(x)))
        h = F.relu(self.norm2(self.conv2(h)))
        for blk in self.res_blocks:
            h = blk(h)
        h = F.relu(self.norm3(self.up1(h)))
        h = F.relu(self.norm4(self.up2(h)))
        return self.final_conv(h)

    HOST_CHUNK_PIXELS = 32 * 256 * 256          # images per pipeline chunk = this many pixels (32 images at 256x256, 1 at 1080p)

    @torch.no_grad()
    def _forward_pinned_host(self, x_host: torch.Tensor) -> torch.Tensor:
        """Extension beyond the reference API (where a CPU input to a CUDA module is an error): a PINNED host batch is stylised
        on the module's GPU as a three-stage pipeline -- host->device copy of chunk i+1 (copy stream), network forward of chunk
        i (current stream), device->host copy of chunk i-1 (second copy stream) -- and comes back as a pinned host tensor.
        The PCIe transfers of a large batch (2 x 100 MB for 256 images) then hide under the forward instead of bracketing it."""
        dev = self.conv1.conv.weight.device
        B, _, H, W = x_host.shape
        per = max(1, min(B, self.HOST_CHUNK_PIXELS // (H * W)))
        if per >= B:                             # one chunk: nothing to overlap -- the ordinary (CUDA-graph) forward between two copies
            y = self.forward(x_host.to(dev, non_blocking=True))
            out = torch.empty(y.shape, dtype=y.dtype, pin_memory=True)
            out.copy_(y, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            return out
        plan = self._plan()
        cur = torch.cuda.current_stream(dev)
        s_in, s_out = (self.__dict__.get("_host_streams") or (None, None))
        if s_in is None or s_in.device != dev:
            s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
            self.__dict__["_host_streams"] = (s_in, s_out)
        x32 = x_host if x_host.dtype == torch.float32 else x_host.float().pin_memory()
        xd = [torch.empty((per, 3, H, W), dtype=torch.float32, device=dev) for _ in range(2)]
        free = [None, None]                      # event: the forward that read xd[j] has finished
        out_host, keep = None, []
        s_in.wait_stream(cur)
        bounds = [(lo, min(B, lo + per)) for lo in range(0, B, per)]
        for i, (lo, hi) in enumerate(bounds):
            j = i & 1
            with torch.cuda.stream(s_in):
                if free[j] is not None:
                    s_in.wait_event(free[j])
                xd[j][:hi - lo].copy_(x32[lo:hi], non_blocking=True)
                ready = torch.cuda.Event(); ready.record(s_in)
            cur.wait_event(ready)
            y = plan.forward(xd[j][:hi - lo])
            done = torch.cuda.Event(); done.record(cur)
            free[j] = done
            if out_host is None:
                out_host = torch.empty((B,) + tuple(y.shape[1:]), dtype=y.dtype, pin_memory=True)      # cached pinned block
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                out_host[lo:hi].copy_(y, non_blocking=True)
            keep.append(y)                        # keep device outputs alive until their copies have drained
        s_out.synchronize()                       # the caller receives a host tensor: its bytes must have landed
        cur.wait_stream(s_in)
        return out_host

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if _exporting(x):
            return self._stock_forward(x)
        if not x.is_cuda:
            if x.is_pinned() and x.dim() == 4 and x.shape[1] == 3 and not (torch.is_grad_enabled() and self.training) \
                    and self.conv1.conv.weight.is_cuda:
                return self._forward_pinned_host(x)
            raise RuntimeError("StyleTransferNet (B200 drop-in) needs CUDA tensors: there is no CPU fallback "
                               "(pinned host batches are accepted for inference and are pipelined through the GPU)")
        named = self._named()
        params = list(named.values())
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            if graphs.enabled() and not torch.cuda.is_current_stream_capturing():
                # training step as two CUDA-graph replays (forward incl. weight re-pack, backward) per input shape
                cache = self.__dict__.setdefault("_train_graphs", {})
                # (every parameter address is part of the key: the captured graphs read the weights in place, so a parameter
                #  whose storage was replaced -- `.to()`, `p.data = ...` -- must not hit a stale capture)
                precision = self._resolved_precision(True)
                key = (precision, tuple(x.shape), x.device.index, self.training, tuple(p.data_ptr() for p in params))
                state = cache.get(key)
                busy = state is not None and state.in_flight.busy()
                # the Dropout2d scales are drawn straight into the captured graph's input buffer (unless that graph is in flight)
                drops = self._dropout_scales(x, out=state.drop_input() if state is not None and not busy else None)
                if state is None:
                    if len(cache) >= 4:
                        cache.clear()
                    state = cache[key] = autograd_fns.StyleNetTrainGraph(dict(named), precision, x, drops)
                if not busy:
                    state.bucket_hook = self.__dict__.get("_fnst_bucket_hook")       # set by parallel.GradientAllReduce
                    return autograd_fns.stylenet_graphed_apply(state, x, drops, params)
                # an earlier forward of this graph still waits for its backward (gradient accumulation, two losses on two
                # inputs): the captured tape holds ONE forward, so this call takes the eager per-call-tape path below
            else:
                drops = self._dropout_scales(x)
            names = list(named)
            return autograd_fns.stylenet_apply(self._plan(need_grad=True), names, x, drops, params)
        plan = self._plan()
        use_graph = (not self.training and os.environ.get("FNST_CUDA_GRAPH", "1") != "0"
                     and x.shape[0] * x.shape[2] * x.shape[3] <= self.GRAPH_MAX_PIXELS
                     and not torch.cuda.is_current_stream_capturing())
        if use_graph:
            return self._graph_forward(plan, x)
        return plan.forward(x, self._dropout_scales(x))
