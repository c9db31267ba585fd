"""Distortion-function base class (mirror of pymde/functions/function.py:9-30).

A `Function` is a torch.nn.Module holding per-edge parameter tensors as buffers.  Unlike the
reference, evaluation does not run torch elementwise ops: `forward` calls the CUDA kernel
`mde_function_eval` (value) and its closed-form derivative feeds autograd, and the fused MDE
path reads the table form (`_table()`) directly."""
import ctypes as C

import torch

from .. import _lib
from .. import util


class _Eval(torch.autograd.Function):
    @staticmethod
    def forward(ctx, distances, fn):
        f, fp = fn._eval(distances, want_grad=distances.requires_grad)
        if fp is not None:
            ctx.save_for_backward(fp)
        return f

    @staticmethod
    def backward(ctx, grad_output):
        (fp,) = ctx.saved_tensors
        return grad_output * fp, None


class Function(torch.nn.Module):
    """Vector distortion function: distances (p,) -> distortions (p,)."""

    # subclasses set these
    _fn_id = None

    def __init__(self):
        super(Function, self).__init__()

    def __setattr__(self, name, value):
        if isinstance(value, torch.Tensor):
            self.register_buffer(name, value)
        else:
            super(Function, self).__setattr__(name, value)

    @property
    def device(self):
        bufs = list(self.buffers())
        if not bufs:
            return None
        dev = str(bufs[0].device)
        return dev if all(str(b.device) == dev for b in bufs) else None

    # ---- table form consumed by the C ABI ------------------------------------------------
    def _scalars(self):
        return (0.0, 0.0, 0.0)

    def _par0(self):
        raise NotImplementedError

    def _par1(self):
        return None

    def _table(self):
        """(mde_fn_t, par0 tensor, par1 tensor or None)"""
        t = _lib.mde_fn_t()
        t.fn_att = t.fn_rep = int(self._fn_id)
        sc = self._scalars()
        for i in range(3):
            t.att[i] = t.rep[i] = float(sc[i])
        t.push_pull = 0
        return t, self._par0(), self._par1()

    def _supported(self):
        """True when the fused CUDA path can consume this function (fp32 parameters)."""
        p0 = self._par0()
        p1 = self._par1()
        ok = p0.dtype == torch.float32 and (p1 is None or p1.dtype == torch.float32)
        return bool(ok)

    # ---- evaluation -----------------------------------------------------------------------
    def _eval(self, distances, want_grad):
        if distances.device.type != "cuda":
            raise ValueError("pymde_b200 distortion functions evaluate CUDA tensors only")
        lib = _lib.load()
        dev = distances.device
        d = distances.detach().to(torch.float32).contiguous()
        table, par0, par1 = self._table()
        par0 = util.as_f32_cuda(par0, dev).reshape(-1)
        if par0.numel() not in (1, d.numel()):
            raise ValueError("parameter length %d does not match %d distances" % (par0.numel(), d.numel()))
        par1 = None if par1 is None else util.as_f32_cuda(par1, dev).reshape(-1)
        f = torch.empty_like(d)
        fp = torch.empty_like(d) if want_grad else None
        _lib.check(lib.mde_function_eval(C.byref(table), par0.data_ptr(), par0.numel(),
                                         None if par1 is None else par1.data_ptr(), d.data_ptr(), d.numel(),
                                         f.data_ptr(), None if fp is None else fp.data_ptr(),
                                         util.stream_ptr(dev)))
        return f, fp

    def forward(self, distances):
        return _Eval.apply(distances, self)
