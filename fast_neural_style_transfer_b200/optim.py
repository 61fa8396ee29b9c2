"""Optimizer tail of the reference's training step on libfnst kernels.

The reference's step ends with (train.py:203-205)

    torch.nn.utils.clip_grad_norm_(style_net.parameters(), max_norm=1.0)
    optimizer.step()                      # optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5), train.py:135-139

which PyTorch executes as ~60 foreach launches over the 58 parameter tensors plus a few hundred microseconds of
Python.  `clip_grad_norm_` and `Adam` below keep those names, argument meanings, return values and the
`state_dict()` layout of `torch.optim.Adam` (so `load_model_from_checkpoint`, train.py:39-66, can resume either
way), and run as three multi-tensor kernels: squared norm (+ clip coefficient, no host sync), in-place scale, Adam.
`Adam.step(grad_scale=...)` can take the clip coefficient directly and skip the scale pass.

No fallback: CPU tensors, non-fp32 or non-contiguous tensors, amsgrad / maximize / capturable / decoupled weight
decay raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional, Tuple, Union

import torch

from . import ops
from ._lib import check, lib

_WS: Dict[Tuple[int, int], torch.Tensor] = {}


def _workspace(device: torch.device) -> torch.Tensor:
    """Zeroed once; fnst_grad_norm leaves it zeroed after every call.  One per (device, stream)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(idx).cuda_stream)
    ws = _WS.get(key)
    if ws is None:
        ws = _WS[key] = torch.zeros(int(lib.fnst_grad_norm_workspace_bytes()) // 8, dtype=torch.float64, device=device)
    return ws


class _PtrList:
    """ctypes view of a list of device tensors (host array of pointers + element counts), rebuilt (and re-validated) whenever
    an address, element count, dtype or contiguity changes -- `zero_grad(set_to_none=True)` frees and re-allocates the
    gradients every step, so another tensor set can reappear at the same addresses."""

    def __init__(self):
        self.key: Optional[tuple] = None
        self.ptrs = None
        self.numels = None
        self.n = 0

    def update(self, tensors: List[torch.Tensor], what: str, stable: bool = False):
        """stable: the tensors are long-lived objects whose geometry cannot change behind their identity (parameters, Adam
        moments): the key is the tuple of object ids plus addresses (`.data = ...` swaps storage under a stable object) --
        two cheap calls per tensor instead of four.  Gradients are re-created every step, so they are keyed in full."""
        ptrs = tuple([t.data_ptr() for t in tensors])
        if stable:
            key = (ptrs, tuple([id(t) for t in tensors]))
        else:
            key = (ptrs, tuple([t.numel() for t in tensors]), tuple([t.dtype for t in tensors]),
                   all([t.is_contiguous() for t in tensors]), tensors[0].device)
        if key != self.key:
            ops._ctx(tensors[0])                     # raises for CPU tensors: there is no CPU fallback
            for t in tensors:
                if t.device != tensors[0].device:
                    raise RuntimeError(f"{what}: all tensors must live on one CUDA device (one process per GPU)")
                if t.dtype != torch.float32 or not t.is_contiguous():
                    raise RuntimeError(f"{what}: contiguous float32 tensors only (got {t.dtype}, contiguous={t.is_contiguous()})")
                if t.numel() == 0:
                    raise RuntimeError(f"{what}: empty tensors are not supported")
            self.n = len(tensors)
            self.ptrs = (C.c_void_p * self.n)(*ptrs)
            self.numels = (C.c_int64 * self.n)(*[t.numel() for t in tensors])
            self.key = key
        return self


_CLIP_LISTS: Dict[int, _PtrList] = {}


def clip_grad_norm_(parameters: Union[torch.Tensor, Iterable[torch.Tensor]], max_norm: float, norm_type: float = 2.0,
                    error_if_nonfinite: bool = False, foreach: Optional[bool] = None) -> torch.Tensor:
    """`torch.nn.utils.clip_grad_norm_` (train.py:203): scales every `.grad` in place by
    min(1, max_norm / (total_norm + 1e-6)) and returns the total 2-norm as a 0-d tensor.  Two launches, no host sync
    (unless error_if_nonfinite, which has to look at the norm)."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    if float(norm_type) != 2.0:
        raise RuntimeError("clip_grad_norm_ (B200 path): only norm_type=2.0 is implemented")
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return torch.tensor(0.0)
    norm_coef = compute_grad_norm(grads, float(max_norm))
    if error_if_nonfinite and not bool(torch.isfinite(norm_coef[0])):
        raise RuntimeError("The total norm of order 2.0 for gradients from `parameters` is non-finite, so it cannot be clipped.")
    pl = _CLIP_LISTS[grads[0].device.index]
    dev, st = ops._ctx(grads[0])
    check(lib.fnst_grad_scale(pl.ptrs, pl.numels, pl.n, C.c_void_p(norm_coef.data_ptr() + 4), dev, st), "grad_scale")
    ops._count((pl.n + 63) // 64)
    return norm_coef[0]


def compute_grad_norm(grads: List[torch.Tensor], max_norm: float) -> torch.Tensor:
    """float32[2] on the device: (total 2-norm of `grads`, clip coefficient min(1, max_norm/(norm+1e-6)))."""
    g0 = grads[0]
    pl = _CLIP_LISTS.setdefault(g0.device.index, _PtrList()).update(grads, "clip_grad_norm_")
    out = torch.empty(2, dtype=torch.float32, device=g0.device)
    dev, st = ops._ctx(g0)
    check(lib.fnst_grad_norm(pl.ptrs, pl.numels, pl.n, C.c_void_p(_workspace(g0.device).data_ptr()), max_norm,
                             C.c_void_p(out.data_ptr()), dev, st), "grad_norm")
    ops._count((pl.n + 63) // 64)
    return out


class Adam(torch.optim.Optimizer):
    """`torch.optim.Adam` for the options the reference uses (train.py:135-139): lr, betas, eps, coupled L2
    weight_decay.  Same per-parameter state (`step`, `exp_avg`, `exp_avg_sq`) and `state_dict()` layout; `lr` is read
    from the param group at every step, so `CosineAnnealingLR` (train.py:141-145) drives it unchanged."""

    def __init__(self, params, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0, amsgrad: bool = False, *, foreach: Optional[bool] = None, maximize: bool = False,
                 capturable: bool = False, differentiable: bool = False, fused: Optional[bool] = None,
                 decoupled_weight_decay: bool = False):
        if isinstance(lr, torch.Tensor):
            raise ValueError("Adam (B200 path): lr must be a Python float")
        if amsgrad or maximize or capturable or differentiable or decoupled_weight_decay:
            raise ValueError("Adam (B200 path): amsgrad / maximize / capturable / differentiable / decoupled_weight_decay "
                             "are not implemented (the reference uses none of them)")
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.5 < betas[0] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 0: {betas[0]} (0.5 < beta1 < 1 supported)")
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 1: {betas[1]}")
        if not 0.0 <= weight_decay:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False, foreach=None,
                        capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False)
        super().__init__(params, defaults)
        self._lists: Dict[int, Tuple[_PtrList, _PtrList, _PtrList, _PtrList]] = {}

    def _share_steps(self) -> None:
        """All parameters of a group advance together: they share ONE 0-d CPU `step` tensor (torch keeps 58 of them)."""
        for group in self.param_groups:
            shared = None
            for p in group["params"]:
                st = self.state.get(p)
                if st is None or "step" not in st:
                    continue
                if shared is None:
                    shared = st["step"].detach().to("cpu", torch.float32).reshape(())
                st["step"] = shared

    def state_dict(self):
        """torch.optim.Adam's layout; every parameter gets its OWN copy of the step counter (torch.optim.Adam, loading this
        dict, increments each `step` tensor once per parameter -- a shared tensor would advance 58 times per step)."""
        sd = super().state_dict()
        sd["state"] = {k: {name: (v.clone() if name == "step" and isinstance(v, torch.Tensor) else v) for name, v in st.items()}
                       for k, st in sd["state"].items()}
        return sd

    def load_state_dict(self, state_dict) -> None:
        super().load_state_dict(state_dict)
        self._share_steps()
        self._lists.clear()

    def zero_grad(self, set_to_none: bool = True) -> None:
        """`Optimizer.zero_grad` (train.py:199) without the generic per-dtype/device grouping (a few microseconds instead of
        ~0.1 ms of host time for 58 parameters; the step is host-paced right after the NaN check)."""
        if not set_to_none:
            return super().zero_grad(set_to_none=False)
        for group in self.param_groups:
            for p in group["params"]:
                p.grad = None

    @torch.no_grad()
    def step(self, closure=None, *, grad_scale: Optional[torch.Tensor] = None):
        """One update.  `grad_scale`: optional 1-element float32 device tensor multiplied into every gradient inside the
        kernel (e.g. `compute_grad_norm(...)[1:]`, the clip coefficient) -- the gradients themselves are not modified."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            shared_step = None
            for p in params:
                if p.grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["step"] = None
                if shared_step is None:
                    shared_step = st["step"] if st["step"] is not None else torch.zeros((), dtype=torch.float32)
                if st["step"] is not shared_step:
                    if st["step"] is not None and float(st["step"]) != float(shared_step):
                        raise RuntimeError("Adam (B200 path): parameters of one group must share their step count")
                    st["step"] = shared_step
            shared_step += 1
            lists = self._lists.get(gi)
            if lists is None:
                lists = self._lists[gi] = (_PtrList(), _PtrList(), _PtrList(), _PtrList())
            pl = lists[0].update(params, "Adam(params)", stable=True)
            gl = lists[1].update([p.grad for p in params], "Adam(grads)")
            ml = lists[2].update([self.state[p]["exp_avg"] for p in params], "Adam(exp_avg)", stable=True)
            vl = lists[3].update([self.state[p]["exp_avg_sq"] for p in params], "Adam(exp_avg_sq)", stable=True)
            if any(a != b for a, b in zip(pl.numels, gl.numels)):
                raise RuntimeError("Adam (B200 path): gradient shapes do not match their parameters")
            if grad_scale is not None and (grad_scale.device != params[0].device or grad_scale.dtype != torch.float32
                                           or grad_scale.numel() != 1):
                raise RuntimeError("Adam.step: grad_scale must be a 1-element float32 tensor on the parameters' device")
            beta1, beta2 = group["betas"]
            dev, stream = ops._ctx(params[0])
            check(lib.fnst_adam_step(pl.ptrs, gl.ptrs, ml.ptrs, vl.ptrs, pl.numels, pl.n, float(group["lr"]), float(beta1),
                                     float(beta2), float(group["eps"]), float(group["weight_decay"]), int(shared_step),
                                     None if grad_scale is None else C.c_void_p(grad_scale.data_ptr()), dev, stream), "adam_step")
            ops._count((pl.n + 63) // 64)
            # the kernel wrote through raw pointers: tell autograd (and the packed-weight cache keyed on _version)
            torch.autograd.graph.increment_version(params)
        return loss
