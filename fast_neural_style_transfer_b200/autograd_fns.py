"""torch.autograd.Function wrappers that connect the libfnst operators to PyTorch autograd:
whole-network functions for StyleTransferNet and the VGG-19 feature stack, and the loss
reductions (Gram, squared error, total variation)."""
from __future__ import annotations

import weakref
from typing import List, Optional, Sequence

import torch

from . import ops


def _as_nhwc(feat: torch.Tensor, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Logical (B,C,H,W) tensor -> contiguous NHWC tensor (zero-copy when the memory already is NHWC)."""
    v = feat.permute(0, 2, 3, 1)
    if v.is_contiguous() and (dtype is None or v.dtype == dtype):
        return v
    if feat.dtype == torch.float32 and feat.is_contiguous():
        return ops.nchw_to_nhwc(feat, dtype or torch.float32)
    return ops.nchw_to_nhwc(feat.float().contiguous(), dtype or torch.float32)


def _gram_tc_ok(f: torch.Tensor) -> bool:
    n, h, w, c = f.shape
    return f.dtype != torch.float32 and c % 64 == 0 and (h * w) % 8 == 0


def _need_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: CUDA tensor required (the B200 path has no CPU fallback)")


# ---- Gram matrix ----------------------------------------------------------------------------------

class _Gram(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat):
        f = _as_nhwc(feat.detach())
        ctx.save_for_backward(f)
        return ops.gram(f, use_tc=_gram_tc_ok(f))

    @staticmethod
    def backward(ctx, dg):
        from . import backward
        (f,) = ctx.saved_tensors
        return backward.gram_backward(f, dg).permute(0, 3, 1, 2)


def gram(feat: torch.Tensor) -> torch.Tensor:
    _need_cuda(feat, "gram_matrix")
    if torch.is_grad_enabled() and feat.requires_grad:
        return _Gram.apply(feat)
    f = _as_nhwc(feat.detach())
    return ops.gram(f, use_tc=_gram_tc_ok(f))


# ---- zero-copy gradient hand-off to a graphed VGG backward -------------------------------------------------------
#
# The captured VGG backward reads the feature-map gradients from its own static input buffers.  A loss node whose input IS a
# feature map of such a graph (recognised by the address of the graph's static output) claims the matching input buffer when
# it runs forward and writes its gradient straight into it in backward: autograd then hands the graph a tensor that already
# lives at the static address, and the 63 MB of device copies per step (4 x 256 x 256) disappear.  A buffer can be claimed
# once per forward of the graph -- a second consumer of the same feature map gets None and allocates as before, so autograd's
# gradient accumulation never sees two aliases of one buffer.

GRAD_SLOTS: dict = {}        # data_ptr of a graphed feature map -> (weakref to its VGGGraph, feature index)


def _claim_grad_slot(f_nhwc: torch.Tensor):
    ent = GRAD_SLOTS.get(f_nhwc.data_ptr())
    if ent is None:
        return None
    state = ent[0]()
    return None if state is None else state.claim_grad_slot(ent[1], f_nhwc)


# ---- sum of squared differences ----------------------------------------------------------------------

def _sse_forward(a: torch.Tensor, b: torch.Tensor, scale: float) -> torch.Tensor:
    out = torch.empty((), dtype=torch.float32, device=a.device)
    ops.sse_scaled(a, b, scale, out)                      # one launch: reduction, normalisation and the fp32 scalar
    return out


def _match_layout(a: torch.Tensor, b: torch.Tensor):
    """Bring a (B,C,H,W)-logical pair to identical contiguous memory order; Gram-like (B,C,C)/(C,C) pass through."""
    if a.dim() == 4:
        an = _as_nhwc(a)
        bn = _as_nhwc(b) if b.dim() == 4 else b                  # element types may differ (the reduction kernels take both)
        return an, bn.contiguous()
    return a.contiguous(), b.contiguous()


class _SSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, scale):
        an, bn = _match_layout(a.detach(), b.detach())
        ctx.save_for_backward(an, bn)
        ctx.is_feat, ctx.scale = a.dim() == 4, scale
        ctx.slot = _claim_grad_slot(an) if a.dim() == 4 else None
        return _sse_forward(an, bn, scale)

    @staticmethod
    def backward(ctx, g):
        from . import backward
        an, bn = ctx.saved_tensors
        da = backward.sse_backward(an, bn, g, ctx.scale, out=ctx.slot)
        return (da.permute(0, 3, 1, 2) if ctx.is_feat else da), None, None


def sse(a: torch.Tensor, b: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """scale * sum((a - b)^2) with b broadcast over the leading (batch) dimension; differentiable in a."""
    _need_cuda(a, "mse/sse")
    if torch.is_grad_enabled() and a.requires_grad:
        return _SSE.apply(a, b, float(scale))
    an, bn = _match_layout(a.detach(), b.detach())
    return _sse_forward(an, bn, float(scale))


# ---- fused style loss ----------------------------------------------------------------------------------------

def _style_forward(fs, targets, coefs):
    """Gram matrices of the layers (one zero fill for all of them) and sum_l coef_l * SSE(G_l, T_l) as one fp32 scalar."""
    dev = fs[0].device
    arena = ops.ZeroArena(sum(f.shape[0] * f.shape[3] * f.shape[3] for f in fs) + 4 * len(fs), dev)
    grams = [ops.gram(f, use_tc=_gram_tc_ok(f), out=arena.take(f.shape[0], f.shape[3], f.shape[3])) for f in fs]
    out = torch.empty((), dtype=torch.float32, device=dev)
    for i, (g, t, coef) in enumerate(zip(grams, targets, coefs)):
        ops.sse_scaled(g, t, coef, out, accumulate=i > 0)
    return grams, out


class _StyleLossGraph:
    """Forward and backward of the fused style loss as two CUDA graphs for ONE set of operand addresses (feature maps, targets,
    gradient slots).  The style backward runs right after the host-synchronising NaN check of the training loop
    (train.py:193-200) with the GPU idle: launched kernel by kernel it costs ~0.2 ms of host time (three Gram-difference
    kernels, three gather-GEMMs with their descriptors and tensor maps); as a captured graph it is one 4-byte copy of the
    incoming gradient scalar plus one replay.  The cache key holds every address the captured kernels read, so a replay always
    reads the tensors of the current call."""

    def __init__(self, fs, targets, coefs):
        from . import graphs
        # (no reference to the feature maps is kept: they are views of a VGG graph's output buffer, whose "nobody holds my
        #  outputs" test must not see this cache; the key re-establishes on every call that the same addresses are current)
        self.targets, self.coefs = list(targets), list(coefs)
        self.device = fs[0].device
        self.g_static = torch.zeros(1, dtype=torch.float32, device=self.device)
        fs = list(fs)
        self.fwd = graphs.GraphedPlan(lambda: _style_forward(fs, self.targets, self.coefs), [], device=self.device)
        self.grams, self.out = self.fwd.outputs
        self.bwd, self.bwd_slots = None, None
        del fs

    def forward(self) -> torch.Tensor:
        self.fwd()
        return self.out.clone()

    def backward(self, g: torch.Tensor, slots, fs) -> list:
        from . import backward, graphs
        key = tuple(None if t is None else t.data_ptr() for t in slots)
        if self.bwd is None or self.bwd_slots != key:
            def bwd():
                outs = []
                for f, gram_, t, coef, slot in zip(fs, self.grams, self.targets, self.coefs, slots):
                    s = ops.gram_diff_sym(gram_, t, self.g_static, 2.0 * coef, backward.grad_dtype_of(f.dtype))
                    outs.append(backward.gram_apply(f, s, out=slot))
                return outs
            self.g_static.copy_(g.reshape(1))
            self.bwd, self.bwd_slots = graphs.GraphedPlan(bwd, [], device=self.device), key
        self.g_static.copy_(g.reshape(1))
        return self.bwd()


_STYLE_GRAPHS: dict = {}


def _style_graph(fs, targets, coefs):
    """Cached _StyleLossGraph for exactly these operand addresses, or None when graphs are off / a capture is in progress."""
    from . import graphs
    if not graphs.enabled() or torch.cuda.is_current_stream_capturing():
        return None
    key = (tuple((f.data_ptr(), tuple(f.shape), f.dtype) for f in fs), tuple((t.data_ptr(), tuple(t.shape)) for t in targets),
           tuple(coefs), fs[0].device.index)
    state = _STYLE_GRAPHS.get(key)
    if state is None:
        if len(_STYLE_GRAPHS) >= 8:
            _STYLE_GRAPHS.clear()
        # (the graph keeps the tensors it was captured on alive, so their addresses cannot be re-used by other tensors)
        state = _STYLE_GRAPHS[key] = _StyleLossGraph(fs, targets, coefs)
    return state


class _StyleLoss(torch.autograd.Function):
    """sum_l w_l * SSE(gram(F_l), T_l) / c_l^2 over the given layers as ONE autograd node (losses/losses.py:15-44).
    Backward per layer: one kernel builds S = g*w/c^2 * 2*((G-T) + (G-T)^T)/... in the gradient dtype, one batched 1x1
    gather-GEMM computes dF = F S -- instead of ~10 autograd nodes per layer on the critical path after the NaN check.
    Feature maps at stable addresses (views of a captured VGG graph's output) run both passes as CUDA graphs."""

    @staticmethod
    def forward(ctx, weights, targets, cs, *feats):
        fs = [_as_nhwc(f.detach()) for f in feats]
        coefs = [w / (c * c) for w, c in zip(weights, cs)]
        ctx.slots = [_claim_grad_slot(f) for f in fs]
        ctx.graph = _style_graph(fs, targets, coefs) if all(f.data_ptr() in GRAD_SLOTS for f in fs) else None
        ctx.fs = fs
        if ctx.graph is not None:
            return ctx.graph.forward()
        grams, out = _style_forward(fs, targets, coefs)
        ctx.grams, ctx.targets, ctx.coefs = grams, targets, coefs
        return out

    @staticmethod
    def backward(ctx, g):
        from . import backward
        if ctx.graph is not None:
            outs = ctx.graph.backward(g, ctx.slots, ctx.fs)
            ctx.fs = None
            return (None, None, None) + tuple(o.permute(0, 3, 1, 2) for o in outs)
        scale = g.reshape(1).float()
        outs = []
        for f, gram_, t, coef, slot in zip(ctx.fs, ctx.grams, ctx.targets, ctx.coefs, ctx.slots):
            s = ops.gram_diff_sym(gram_, t, scale, 2.0 * coef, backward.grad_dtype_of(f.dtype))     # d/dG of coef*sum(G-T)^2 is 2*coef*(G-T); dF = F (dG + dG^T)
            outs.append(backward.gram_apply(f, s, out=slot).permute(0, 3, 1, 2))
        ctx.fs = ctx.grams = None
        return (None, None, None) + tuple(outs)


def style_loss_fused(feats: Sequence[torch.Tensor], targets: Sequence[torch.Tensor], weights: Sequence[float],
                     cs: Sequence[int]) -> torch.Tensor:
    for f in feats:
        _need_cuda(f, "style_loss")
    targets = [t.detach().float().contiguous() for t in targets]
    if torch.is_grad_enabled() and any(f.requires_grad for f in feats):
        return _StyleLoss.apply(list(weights), targets, list(cs), *feats)
    fs = [_as_nhwc(f.detach()) for f in feats]
    return _style_forward(fs, targets, [w / (c * c) for w, c in zip(weights, cs)])[1]


# ---- total variation ---------------------------------------------------------------------------------

def _tv_forward(x: torch.Tensor, scale: float) -> torch.Tensor:
    out = torch.empty((), dtype=torch.float32, device=x.device)
    ops.tv_scaled(x, scale, out)
    return out


class _TV(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, scale):
        x = img.detach().float().contiguous()
        ctx.save_for_backward(x)
        ctx.scale = scale
        return _tv_forward(x, scale)

    @staticmethod
    def backward(ctx, g):
        from . import backward
        (x,) = ctx.saved_tensors
        return backward.tv_backward(x, g, ctx.scale), None


def tv(img: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """scale * (sum dh^2 + sum dw^2); differentiable in img."""
    _need_cuda(img, "total_variation_loss")
    if torch.is_grad_enabled() and img.requires_grad:
        return _TV.apply(img, float(scale))
    return _tv_forward(img.detach().float().contiguous(), float(scale))


# ---- whole-network functions -----------------------------------------------------------------------------

class _StyleNet(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, names, x, drop, *params):
        tape: dict = {}
        y = plan.forward(x.detach(), drop, tape)
        ctx.plan, ctx.names, ctx.tape, ctx.drop = plan, names, tape, drop
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import backward
        grads = backward.stylenet_backward(ctx.plan, ctx.tape, dy.contiguous())
        ctx.tape = None
        return (None, None, None, None) + tuple(grads[n] for n in ctx.names)


def stylenet_apply(plan, names: Sequence[str], x: torch.Tensor, drop, params: Sequence[torch.Tensor]) -> torch.Tensor:
    return _StyleNet.apply(plan, list(names), x, drop, *params)


def _grad_interface(feats):
    """Feature maps that will receive gradients are handed to autograd in the gradient element type: autograd casts an
    incoming gradient to the dtype of the forward output, and fp16 cannot hold the un-normalised Gram / style gradients
    (SURVEY 7.2).  fp16 features are therefore returned as bfloat16 copies (the fp16 originals stay on the tape: they are the
    ReLU / max-pool masks and the operands of the next layer); bf16 / fp32 features pass through."""
    return tuple(ops.cast(f, torch.bfloat16) if f.dtype == torch.float16 else f for f in feats)


class _VGG(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, x):
        tape: dict = {}
        feats = plan.forward(x.detach(), tape)
        ctx.plan, ctx.tape = plan, tape
        return _grad_interface(feats)

    @staticmethod
    def backward(ctx, *dfeats):
        from . import backward
        dx = backward.vgg_backward(ctx.plan, ctx.tape, dfeats)
        ctx.tape = None
        return None, dx


def vgg_apply(plan, x: torch.Tensor) -> List[torch.Tensor]:
    return list(_VGG.apply(plan, x))


# ---- CUDA-graph variants (training): one replay for the forward, one for the backward ----------------------------
#
# A captured forward keeps its saved-for-backward tape in ONE set of static buffers, so only one forward of a given
# graph can be "in flight" (forward done, backward still to come) at a time.  Every graphed forward therefore hands a
# token to its autograd node and bumps the graph's generation: `busy()` tells the module that an earlier forward of
# this graph is still waiting for its backward (gradient accumulation over micro-batches, vgg(a) + vgg(b), ...) -- the
# module then runs the new forward through the eager per-call-tape Functions above -- and backward() refuses to run on
# a tape that a later forward has overwritten.

class _Token:
    pass


class _InFlight:
    def __init__(self):
        self.generation = 0
        self._live = None

    def busy(self) -> bool:
        return self._live is not None and self._live() is not None

    def begin(self, ctx) -> None:
        self.generation += 1
        ctx.generation = self.generation
        ctx.token = _Token()
        self._live = weakref.ref(ctx.token)

    def check(self, ctx, what: str) -> None:
        if ctx.generation != self.generation:
            raise RuntimeError(f"{what}: the saved activations of this forward were overwritten by a later forward of the same "
                               "CUDA graph before backward() ran (set FNST_CUDA_GRAPH=0 for unrestricted autograd semantics)")
        ctx.token = None


class StyleNetTrainGraph:
    """Forward (incl. weight re-pack) and backward of StyleTransferNet captured as two CUDA graphs for one input shape."""

    def __init__(self, params: "dict[str, torch.nn.Parameter]", precision: str, x: torch.Tensor, drops):
        from . import engine, graphs
        self.names = list(params)
        self.params = params
        self.numels = [params[n].numel() for n in self.names]
        self.tape: dict = {}
        self.plan = None
        self.has_drop = drops is not None
        self.in_flight = _InFlight()

        def fwd(x_, *drops_):
            self.tape.clear()
            self.plan = engine.StyleNetPlan(precision).pack(self.params, for_backward=True)
            return self.plan.forward(x_, drops_[0] if self.has_drop else None, self.tape)

        # drops: one (5,B,256) tensor (a single graph input; the module draws the Bernoulli scales straight into it)
        self.fwd = graphs.GraphedPlan(fwd, [x.float().contiguous()] + ([drops.float().contiguous()] if self.has_drop else []))
        self.bwd = None
        self.bucket_hook = None      # data-parallel training: callable(flat, lo, hi, last) invoked as each gradient bucket is final
        self.staged = None

    def drop_input(self):
        return self.fwd.static_inputs[1] if self.has_drop else None

    def forward(self, x, drops):
        return self.fwd(x, *([drops] if self.has_drop else []))

    def backward(self, dy):
        """Captured: the whole backward up to the packed weight gradients / InstanceNorm sums (static buffers).  Eager: the
        two assembly launches, which write the 58 gradients into a FRESH flat buffer (nothing the caller receives aliases
        graph memory; no concatenation, no clone)."""
        from . import backward, graphs
        if self.bucket_hook is not None:
            return self._backward_staged(dy)
        if self.bwd is None:
            def bwd(dy_):
                self.core = backward.stylenet_backward_core(self.plan, self.tape, dy_)
                return self.core["staging"]
            self.bwd = graphs.GraphedPlan(bwd, [dy.float().contiguous()])
        self.bwd(dy)
        flat = backward.assemble_gradients(self.core, self.names, self.params)
        return [t.view_as(self.params[n]) for t, n in zip(torch.split(flat, self.numels), self.names)]

    def _backward_staged(self, dy):
        """Data-parallel form: the backward is captured as THREE graphs cut after residual blocks 2 and 0
        (backward.STAGE_CUTS, one memory pool).  After each replay the gradients that stage completed are assembled into
        their slice of the flat buffer and handed to `bucket_hook` (an asynchronous all-reduce on NCCL's stream), so the
        exchange of 97 % of the bytes runs under the remaining stages instead of after the backward."""
        from . import backward
        dev = dy.device
        if self.staged is None:
            static_dy = dy.float().contiguous().clone()
            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side), torch.no_grad():
                backward.stylenet_backward_core(self.plan, self.tape, static_dy)          # warm-up (lazy initialisations)
            cur.wait_stream(side)
            gen = backward.stylenet_backward_stages(self.plan, self.tape, static_dy)
            stages, pool, done = [], None, False
            while not done:
                g = torch.cuda.CUDAGraph()
                n0 = ops.launch_count
                with torch.no_grad(), torch.cuda.graph(g, pool=pool, capture_error_mode="thread_local"):
                    try:
                        core = next(gen)
                    except StopIteration as end:
                        core, done = end.value, True
                pool = g.pool()
                stages.append((g, core, ops.launch_count - n0))
            self.staged = (static_dy, stages)
        static_dy, stages = self.staged
        if static_dy.data_ptr() != dy.data_ptr():
            static_dy.copy_(dy)
        flat = None
        for k, (g, core, launches) in enumerate(stages):
            g.replay()
            ops.launch_count += launches
            first, last = backward.bucket_bounds(k)
            flat = backward.assemble_gradients(core, self.names, self.params, flat, first, last)
            asm = backward._assembly(self.names, [self.params[n].shape for n in self.names], core["tc"], dev)
            lo = 0 if first is None else asm["offsets"][first]
            hi = asm["total"] if last is None else asm["offsets"][last]
            self.bucket_hook(flat, lo, hi, k == len(stages) - 1)
        return [t.view_as(self.params[n]) for t, n in zip(torch.split(flat, self.numels), self.names)]


class _StyleNetGraphed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, state, x, drops, *params):
        ctx.state = state
        state.in_flight.begin(ctx)
        return state.forward(x.detach(), drops).clone()

    @staticmethod
    def backward(ctx, dy):
        ctx.state.in_flight.check(ctx, "StyleTransferNet backward")
        return (None, None, None) + tuple(ctx.state.backward(dy.contiguous()))


def stylenet_graphed_apply(state: StyleNetTrainGraph, x, drops, params):
    return _StyleNetGraphed.apply(state, x, drops, *params)


def _storage_refs(t: torch.Tensor) -> int:
    """Number of tensors (views included) that share t's storage, plus one for the temporary handle of this call."""
    return torch._C._storage_Use_Count(t.untyped_storage()._cdata)


class VGGGraph:
    """Graphed VGG feature stack for one input shape: forward (with or without tape) and data-gradient backward.

    forward(x, alias=True) returns VIEWS of the graph's static output buffer (no 67 MB device copy per call at 4 x 256 x 256);
    the module only does so while `outputs_free()` -- no tensor handed out by an earlier call is still alive -- and otherwise
    replays another instance of the graph (the reference loop keeps last step's features in its local variables until the
    new ones are assigned, so two instances alternate) or, with all instances taken, falls back to a copy."""

    def __init__(self, plan, x: torch.Tensor, with_tape: bool):
        from . import graphs
        self.plan = plan
        self.tape: dict = {}
        self.with_tape = with_tape
        self.in_flight = _InFlight()
        self.grad_slots: dict = {}          # feature index -> static input of the most recently captured backward graph
        self.claimed: set = set()

        def fwd(x_):
            # the five feature maps live in ONE flat static buffer, so handing them to the caller is one device copy, not five
            self.tape.clear()
            B, _, H, W = x_.shape
            self.shapes = plan.feature_shapes(B, H, W)
            self.numels = [s[0] * s[1] * s[2] * s[3] for s in self.shapes]
            iface = torch.bfloat16 if (with_tape and plan.dtype == torch.float16) else plan.dtype       # see _grad_interface
            flat = torch.empty(sum(self.numels), dtype=iface, device=x_.device)
            views = [v.view(s) for v, s in zip(torch.split(flat, self.numels), self.shapes)]
            if iface == plan.dtype:
                plan.forward(x_, self.tape if with_tape else None, out_buffers=dict(zip(plan.FEATURE_LAYERS, views)))
            else:
                for f, v in zip(plan.forward(x_, self.tape), views):
                    ops.cast(f, iface, out=v)
            return flat

        self.fwd = graphs.GraphedPlan(fwd, [x.float().contiguous()])
        self.bwd = {}
        plan._out_buffers = None            # the plan must not keep views of this graph's output alive (see outputs_free)
        self._refs0 = _storage_refs(self.fwd.outputs)
        for k in [k for k, (ref, _) in GRAD_SLOTS.items() if ref() is None]:
            del GRAD_SLOTS[k]                # graphs that no longer exist
        offs = 0
        for i, n in enumerate(self.numels):
            GRAD_SLOTS[self.fwd.outputs.data_ptr() + offs * self.fwd.outputs.element_size()] = (weakref.ref(self), i)
            offs += n

    def outputs_free(self) -> bool:
        """True when nothing outside this object references the static output buffer any more."""
        return _storage_refs(self.fwd.outputs) <= self._refs0

    def claim_grad_slot(self, i: int, f_nhwc: torch.Tensor):
        """Static gradient-input buffer of feature i of the captured backward, once per forward (None: no capture yet,
        already claimed, or a different geometry)."""
        slot = self.grad_slots.get(i)
        if slot is None or i in self.claimed or tuple(slot.shape) != tuple(f_nhwc.shape):
            return None
        self.claimed.add(i)
        return slot

    def forward(self, x, alias: bool = False):
        self.claimed.clear()
        flat = self.fwd(x)
        if not alias:
            flat = flat.clone()
        return tuple(v.view(s) for v, s in zip(torch.split(flat, self.numels), self.shapes))

    def backward(self, dfeats):
        from . import backward, graphs
        pattern = tuple(g is not None for g in dfeats)
        live = [g.contiguous() for g in dfeats if g is not None]
        if pattern not in self.bwd:
            def bwd(*gs):
                it = iter(gs)
                full = [next(it) if used else None for used in pattern]
                return backward.vgg_backward(self.plan, self.tape, full)
            self.bwd[pattern] = graphs.GraphedPlan(bwd, live)
            it = iter(self.bwd[pattern].static_inputs)
            self.grad_slots = {i: next(it) for i, used in enumerate(pattern) if used}
        return self.bwd[pattern](*live).clone()


class _VGGGraphed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, state, x, alias):
        ctx.state = state
        state.in_flight.begin(ctx)
        return state.forward(x.detach(), alias)

    @staticmethod
    def backward(ctx, *dfeats):
        ctx.state.in_flight.check(ctx, "VGG19 backward")
        return None, ctx.state.backward(dfeats), None


def vgg_graphed_apply(state: VGGGraph, x, alias: bool = False) -> List[torch.Tensor]:
    return list(_VGGGraphed.apply(state, x, alias))
