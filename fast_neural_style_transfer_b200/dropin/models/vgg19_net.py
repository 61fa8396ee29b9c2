"""Drop-in for the reference's models/vgg19_net.py: class VGG19 whose forward returns the five
feature maps [relu1_2, relu2_2, relu3_3, relu4_2, relu4_3] (models/vgg19_net.py:56-65), computed by
libfnst.  The returned tensors have the reference's logical (B,C,H,W) shape; their memory is the
path's NHWC activation buffer (a permuted view, no copy) in the path's element type.

Weights follow the reference (models/vgg19_net.py:27 builds `vgg19(weights='DEFAULT')`, the pretrained ImageNet
weights): in order, FNST_VGG19_WEIGHTS=/path/vgg19.pth (a torchvision state dict, the offline override), then
FNST_VGG19_RANDOM_INIT=1 (explicit opt-in to torchvision's random init -- tests and the benchmark, which have no
network), then torchvision's own 'DEFAULT' download.  If none of them yields weights the constructor RAISES, as the
reference does without a network: a perceptual loss against a silently random VGG would be meaningless.
The reference constructor's undefined `slice5` (:51) is created here.

Precision: `vgg.precision` in {"bf16" (default), "fp16", "fp32"}; FNST_VGG_PRECISION overrides, FNST_PRECISION=fp32
selects fp32 for both networks.  Gradients are bf16 on both tensor-core paths (fp16 cannot hold the un-normalised
Gram / style gradients).  Because autograd casts an incoming gradient to the dtype of the forward output,
"fp16" keeps fp16 activations INSIDE the stack (11-bit masks and operands) but returns feature maps that need a gradient
as bfloat16 copies.
"""
import os
import sys

import torch
import torch.nn as nn

_PKG_PARENT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _PKG_PARENT not in sys.path:
    sys.path.append(_PKG_PARENT)

from fast_neural_style_transfer_b200 import engine, graphs    # noqa: E402
from fast_neural_style_transfer_b200 import autograd_fns      # noqa: E402

_SLICES = (("slice1", 0, 4), ("slice2", 4, 9), ("slice3", 9, 16), ("slice4", 16, 22), ("slice5", 22, 25))


class VGG19(nn.Module):
    def __init__(self):
        super().__init__()
        from torchvision.models import vgg19
        path = os.environ.get("FNST_VGG19_WEIGHTS")
        if path:
            net = vgg19(weights=None)
            net.load_state_dict(torch.load(path, map_location="cpu"))
        elif os.environ.get("FNST_VGG19_RANDOM_INIT", "0") not in ("", "0"):
            net = vgg19(weights=None)
        else:
            try:
                net = vgg19(weights="DEFAULT")                       # models/vgg19_net.py:27
            except Exception as exc:                                   # no network / no cached file
                raise RuntimeError(
                    "VGG19: the pretrained torchvision weights (vgg19(weights='DEFAULT'), as in the reference) could not be "
                    f"loaded ({type(exc).__name__}: {exc}).  Set FNST_VGG19_WEIGHTS=/path/to/vgg19.pth to load a torchvision "
                    "state dict from disk, or FNST_VGG19_RANDOM_INIT=1 to opt in to random weights (tests / benchmarks only)."
                ) from exc
        feats = net.features
        for name, lo, hi in _SLICES:
            seq = nn.Sequential()
            for i in range(lo, hi):
                seq.add_module(str(i), feats[i])
            setattr(self, name, seq)
        for p in self.parameters():
            p.requires_grad = False
        self.precision = os.environ.get("FNST_VGG_PRECISION") or ("fp32" if os.environ.get("FNST_PRECISION") == "fp32" else "bf16")

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_plan_cache", None)
        state.pop("_graphs", None)
        state.pop("_named_cache", None)
        return state

    def _named(self):
        """dict(self.named_parameters()), cached (see StyleTransferNet._named)."""
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

    def _plan(self) -> "engine.VGGPlan":
        params = self._named()
        key = (self.precision, tuple((p.data_ptr(), p._version) for p in params.values()))
        cache = self.__dict__.get("_plan_cache")
        if cache is None or cache[0] != key:
            cache = (key, engine.VGGPlan(self.precision).pack(params))
            self.__dict__["_plan_cache"] = cache
        return cache[1]

    MAX_GRAPH_INSTANCES = 3          # captured copies of one (shape, tape) graph that may hand out views of their outputs

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("VGG19 (B200 drop-in) needs CUDA tensors: there is no CPU fallback")
        plan = self._plan()
        need_grad = torch.is_grad_enabled() and x.requires_grad
        small = x.shape[0] * x.shape[2] * x.shape[3] <= 16 * 256 * 256
        if graphs.enabled() and small and not torch.cuda.is_current_stream_capturing():
            cache = self.__dict__.setdefault("_graphs", {})
            key = (id(plan), tuple(x.shape), x.device.index, need_grad)
            states = cache.get(key)
            if states is None:
                for k in [k for k in cache if k[0] != id(plan)] if len(cache) < 8 else list(cache):
                    del cache[k]
                states = cache[key] = []
            # An instance whose earlier outputs are all dead (and whose tape is not awaiting a backward) hands out views of its
            # static output buffer: no copy.  The reference loop still holds last step's features while it calls vgg() again,
            # so a second instance is captured and the two alternate; beyond MAX_GRAPH_INSTANCES the output is copied out.
            state = next((s for s in states if not s.in_flight.busy() and s.outputs_free()), None)
            alias = True
            if state is None and len(states) < self.MAX_GRAPH_INSTANCES:
                state = autograd_fns.VGGGraph(plan, x, with_tape=need_grad)
                states.append(state)
            if state is None:
                alias = False
                state = next((s for s in states if not s.in_flight.busy()), None)
            if state is None:
                feats = autograd_fns.vgg_apply(plan, x)      # every instance awaits its backward (vgg(a), vgg(b), ...): per-call tape
            elif not need_grad:
                feats = state.forward(x.detach(), alias)
            else:
                feats = autograd_fns.vgg_graphed_apply(state, x, alias)
        elif need_grad:
            feats = autograd_fns.vgg_apply(plan, x)
        else:
            feats = plan.forward(x)
        return [f.permute(0, 3, 1, 2) for f in feats]
