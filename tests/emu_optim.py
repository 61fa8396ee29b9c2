"""numpy emulation of the optimizer-tail entry points of include/fnst.h (fnst_grad_norm, fnst_grad_scale,
fnst_adam_step) operating on HOST pointers.

TEST INFRASTRUCTURE: lets the CPU suite drive the host logic of fast_neural_style_transfer_b200.optim (pointer lists,
shared step counter, lr read from the param group, state_dict round trips) against torch.optim.Adam without a GPU.
Never imported by the product.
"""
import ctypes as C
import math

import numpy as np


def _arr(ptr, n):
    return np.ctypeslib.as_array((C.c_float * n).from_address(int(ptr)))


def _val(p):
    return p.value if hasattr(p, "value") else p


class EmuLib:
    def __init__(self):
        self.calls = []

    def fnst_grad_norm_workspace_bytes(self):
        return 16

    def fnst_last_error(self):
        return b"emulated"

    def fnst_grad_norm(self, grads, numels, n, ws, max_norm, out, dev, stream):
        self.calls.append("grad_norm")
        sq = 0.0
        for i in range(n):
            g = _arr(grads[i], numels[i]).astype(np.float64)
            sq += float((g * g).sum())
        norm = np.float32(math.sqrt(sq))
        coef = min(np.float32(1.0), np.float32(max_norm) / (norm + np.float32(1e-6)))
        o = _arr(_val(out), 2)
        o[0], o[1] = norm, coef
        return 0

    def fnst_grad_scale(self, grads, numels, n, coef, dev, stream):
        self.calls.append("grad_scale")
        c = _arr(_val(coef), 1)[0]
        for i in range(n):
            g = _arr(grads[i], numels[i])
            g *= c
        return 0

    def fnst_adam_step(self, params, grads, ms, vs, numels, n, lr, b1, b2, eps, wd, step, gscale, dev, stream):
        self.calls.append("adam_step")
        f = np.float32
        scale = f(1.0) if gscale is None else _arr(_val(gscale), 1)[0]
        bc1 = 1.0 - b1 ** step
        bc2 = 1.0 - b2 ** step
        for i in range(n):
            p, g = _arr(params[i], numels[i]), _arr(grads[i], numels[i])
            m, v = _arr(ms[i], numels[i]), _arr(vs[i], numels[i])
            gg = g * scale + f(wd) * p
            m += f(1.0 - b1) * (gg - m)
            v *= f(b2)
            v += f(1.0 - b2) * gg * gg
            denom = np.sqrt(v) / f(math.sqrt(bc2)) + f(eps)
            p -= f(lr / bc1) * (m / denom)
        return 0
