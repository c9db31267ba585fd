"""Losses: distortion functions derived from original deviations, f_k(d) = l(d, delta_k).

Class names and constructor signatures follow pymde/functions/losses.py:61-239; the bodies are
table entries for the CUDA kernels (pymde_b200/csrc/mde_common.cuh::eval_fn)."""
from .. import util
from .function import Function


class _Deviated(Function):
    def __init__(self, deviations):
        super(_Deviated, self).__init__()
        self.deviations = util.to_tensor(deviations)

    def _par0(self):
        return self.deviations


class Absolute(_Deviated):
    """l(d, delta) = |d - delta|"""
    _fn_id = 20


class Quadratic(_Deviated):
    """l(d, delta) = (d - delta)^2"""
    _fn_id = 21


class WeightedQuadratic(_Deviated):
    """l(d, delta) = w (d - delta)^2, w = 1/delta^2 unless given"""
    _fn_id = 22

    def __init__(self, deviations, weights=None):
        super(WeightedQuadratic, self).__init__(deviations)
        if weights is None:
            weights = 1.0 / self.deviations.pow(2)
        self.weights = util.to_tensor(weights, device=self.deviations.device)

    def _par1(self):
        return self.weights


class Huber(_Deviated):
    """l = r^2 for r = |d - delta| < threshold, threshold (2 r - threshold) otherwise"""
    _fn_id = 23

    def __init__(self, deviations, threshold):
        super(Huber, self).__init__(deviations)
        self.threshold = threshold

    def _scalars(self):
        return (float(self.threshold), 0.0, 0.0)


class Cubic(_Deviated):
    """l(d, delta) = |d - delta|^3"""
    _fn_id = 24


class Power(_Deviated):
    """l(d, delta) = |d - delta|^exponent"""
    _fn_id = 25

    def __init__(self, deviations, exponent):
        super(Power, self).__init__(deviations)
        self.exponent = util.to_tensor(exponent, device=self.deviations.device)

    def _scalars(self):
        return (float(self.exponent), 0.0, 0.0)


class Logistic(_Deviated):
    """l(d, delta) = log(1 + exp(|d - delta|))"""
    _fn_id = 26


class Fractional(_Deviated):
    """l(d, delta) = max(delta / d, d / delta) - 1"""
    _fn_id = 27


class SoftFractional(_Deviated):
    """soft maximum (parameter gamma) of delta/d and d/delta"""
    _fn_id = 28

    def __init__(self, deviations, gamma=10.0):
        super(SoftFractional, self).__init__(deviations)
        self.gamma = util.to_tensor(gamma, device=self.deviations.device)
        if gamma <= 0.0:
            raise ValueError("gamma must be positive, received ", float(gamma))

    def _scalars(self):
        return (float(self.gamma), 0.0, 0.0)
