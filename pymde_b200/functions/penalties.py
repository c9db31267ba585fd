"""Penalties: distortion functions derived from weights, f_k(d) = w_k p(d).

Class names and constructor signatures follow pymde/functions/penalties.py:112-400; the bodies
are table entries for the CUDA kernels (pymde_b200/csrc/mde_common.cuh::eval_fn)."""
import torch

from .. import _lib
from .. import util
from .function import Function


def _exponent_tensor(exponent, device=None):
    if not isinstance(exponent, torch.Tensor):
        exponent = torch.tensor(exponent, device=device)
    return exponent


class _Weighted(Function):
    def __init__(self, weights):
        super(_Weighted, self).__init__()
        self.weights = util.to_tensor(weights)

    def _par0(self):
        return self.weights


class Linear(_Weighted):
    """p(d) = d"""
    _fn_id = 1


class Quadratic(_Weighted):
    """p(d) = d^2"""
    _fn_id = 2


class Cubic(_Weighted):
    """p(d) = d^3"""
    _fn_id = 3


class _WithExponent(_Weighted):
    _default_exponent = None

    def __init__(self, weights, exponent=None):
        super(_WithExponent, self).__init__(weights)
        if exponent is None:
            exponent = self._default_exponent
        self.exponent = _exponent_tensor(exponent, self.weights.device)

    def _scalars(self):
        return (float(self.exponent), 0.0, 0.0)


class Power(_WithExponent):
    """p(d) = d^exponent"""
    _fn_id = 4

    def __init__(self, weights, exponent):
        super(Power, self).__init__(weights, exponent)


class Huber(_Weighted):
    """p(d) = 0.5 d^2 for d < threshold, threshold (d - 0.5 threshold) otherwise"""
    _fn_id = 5

    def __init__(self, weights, threshold=0.5):
        if threshold < 0:
            raise ValueError("Threshold must be nonnegative, received ", threshold)
        super(Huber, self).__init__(weights)
        self.threshold = threshold

    def _scalars(self):
        return (float(self.threshold), 0.0, 0.0)


class Logistic(_Weighted):
    """p(d) = log(1 + exp(alpha (d - threshold)))"""
    _fn_id = 6

    def __init__(self, weights, threshold=0.0, alpha=3.0):
        if threshold < 0:
            raise ValueError("Threshold must be nonnegative, received ", threshold)
        super(Logistic, self).__init__(weights)
        self.threshold = threshold
        self.alpha = alpha

    def _scalars(self):
        return (float(self.threshold), float(self.alpha), 0.0)


class _TorchPenalty(_Weighted):
    """Penalties the reference ships without documentation (pymde/functions/penalties.py:269-307).  They are not
    in the kernels' function table: an MDE using them evaluates f and f' with torch on the per-edge distances the
    CUDA path produces and scatters the result with `mde_scatter_external`, like any user-defined callable."""

    def _supported(self):
        return False


class Sigmoid(_TorchPenalty):
    """f(d) = w * sigmoid(alpha (d - threshold))"""

    def __init__(self, weights, threshold, alpha=1.0):
        if threshold < 0:
            raise ValueError("Threshold must be nonnegative, received ", threshold)
        super(Sigmoid, self).__init__(weights)
        self.threshold = threshold
        self.alpha = alpha

    def forward(self, distances):
        return self.weights * torch.sigmoid(self.alpha * (distances - self.threshold))


class Hinge(_TorchPenalty):
    """f(d) = max(0, w * (d - (threshold - sign(w) * sigma))), sigma defaults to threshold / 2"""

    def __init__(self, weights, threshold, sigma=None):
        if threshold < 0:
            raise ValueError("Threshold must be nonnegative, received ", threshold)
        super(Hinge, self).__init__(weights)
        self.threshold = threshold
        self.sigma = threshold / 2 if sigma is None else sigma

    def forward(self, distances):
        knee = self.threshold - torch.sign(self.weights) * self.sigma
        return torch.clamp(self.weights * (distances - knee), min=0.0)


class Log1p(_WithExponent):
    """p(d) = log(1 + d^exponent)"""
    _fn_id = 7
    _default_exponent = 1.5

    def __init__(self, weights, exponent=1.5):
        super(Log1p, self).__init__(weights, exponent)


class Log(_WithExponent):
    """p(d) = log(1 - exp(-d^exponent))"""
    _fn_id = 8
    _default_exponent = 1.0

    def __init__(self, weights, exponent=1.0):
        super(Log, self).__init__(weights, exponent)


class InvPower(_WithExponent):
    """p(d) = 1 / d^exponent (weights must be nonpositive)"""
    _fn_id = 9

    def __init__(self, weights, exponent=1):
        if not bool((util.to_tensor(weights) <= 0).all()):
            raise ValueError("Weights must be negative.")
        super(InvPower, self).__init__(weights, exponent)


class LogRatio(_WithExponent):
    """p(d) = log(d^exponent / (1 + d^exponent))"""
    _fn_id = 10

    def __init__(self, weights, exponent=2):
        super(LogRatio, self).__init__(weights, exponent)


class PushAndPull(Function):
    """Attractive penalty for weights >= 0, repulsive penalty for weights < 0
    (pymde/functions/penalties.py:372-400)."""

    def __init__(self, weights, attractive_penalty=Log1p, repulsive_penalty=LogRatio):
        super(PushAndPull, self).__init__()
        self.weights = util.to_tensor(weights)
        if self.weights.nelement() == 1:
            raise ValueError("`PushAndPull` requires at least two weights.")
        self.pos_idx = self.weights >= 0
        self.attractive_penalty = attractive_penalty(self.weights[self.pos_idx])
        self.repulsive_penalty = repulsive_penalty(self.weights[~self.pos_idx])

    def _par0(self):
        return self.weights

    def _table(self):
        a, r = self.attractive_penalty, self.repulsive_penalty
        t = _lib.mde_fn_t()
        t.fn_att, t.fn_rep = int(a._fn_id), int(r._fn_id)
        sa, sr = a._scalars(), r._scalars()
        for i in range(3):
            t.att[i], t.rep[i] = float(sa[i]), float(sr[i])
        t.push_pull = 1
        return t, self.weights, None

    def forward(self, distances):
        # kernel path when both sub-penalties are table functions; otherwise the reference's masked evaluation
        # (pymde/functions/penalties.py:394-400): any callable class is a legal attractive / repulsive penalty
        if self._supported():
            return super(PushAndPull, self).forward(distances)
        out = torch.zeros_like(distances)
        pos = self.pos_idx
        out[pos] = self.attractive_penalty(distances[pos])
        out[~pos] = self.repulsive_penalty(distances[~pos])
        return out

    def _supported(self):
        a, r = self.attractive_penalty, self.repulsive_penalty
        return (isinstance(a, Function) and isinstance(r, Function) and a._fn_id is not None
                and r._fn_id is not None and a._par1() is None and r._par1() is None
                and self.weights.dtype == torch.float32)
