"""Shared description of the golden function cases (mirrors tests/golden/make_golden.py)."""
from oracle import mde_oracle as O

# name -> (fn_att, att scalars, fn_rep, rep scalars)
CASES = {
    "pen_linear": (O.P_LINEAR, (0, 0, 0), None, None),
    "pen_quadratic": (O.P_QUADRATIC, (0, 0, 0), None, None),
    "pen_cubic": (O.P_CUBIC, (0, 0, 0), None, None),
    "pen_power_2.5": (O.P_POWER, (2.5, 0, 0), None, None),
    "pen_huber_0.5": (O.P_HUBER, (0.5, 0, 0), None, None),
    "pen_logistic_0.3_3": (O.P_LOGISTIC, (0.3, 3.0, 0), None, None),
    "pen_log1p_1.5": (O.P_LOG1P, (1.5, 0, 0), None, None),
    "pen_log_1": (O.P_LOG, (1.0, 0, 0), None, None),
    "pen_invpower_1": (O.P_INVPOWER, (1.0, 0, 0), None, None),
    "pen_logratio_2": (O.P_LOGRATIO, (2.0, 0, 0), None, None),
    "pen_pushpull_log1p_log": (O.P_LOG1P, (1.5, 0, 0), O.P_LOG, (1.0, 0, 0)),
    "pen_pushpull_default": (O.P_LOG1P, (1.5, 0, 0), O.P_LOGRATIO, (2.0, 0, 0)),
    "pen_pushpull_quad_invpower": (O.P_QUADRATIC, (0, 0, 0), O.P_INVPOWER, (1.0, 0, 0)),
    "loss_absolute": (O.L_ABSOLUTE, (0, 0, 0), None, None),
    "loss_quadratic": (O.L_QUADRATIC, (0, 0, 0), None, None),
    "loss_weighted_quadratic": (O.L_WEIGHTED_QUADRATIC, (0, 0, 0), None, None),
    "loss_weighted_quadratic_w": (O.L_WEIGHTED_QUADRATIC, (0, 0, 0), None, None),
    "loss_huber_0.7": (O.L_HUBER, (0.7, 0, 0), None, None),
    "loss_cubic": (O.L_CUBIC, (0, 0, 0), None, None),
    "loss_power_1.5": (O.L_POWER, (1.5, 0, 0), None, None),
    "loss_logistic": (O.L_LOGISTIC, (0, 0, 0), None, None),
    "loss_fractional": (O.L_FRACTIONAL, (0, 0, 0), None, None),
    "loss_soft_fractional_10": (O.L_SOFT_FRACTIONAL, (10.0, 0, 0), None, None),
}


def spec_for(name, fn_golden, tag):
    """Build the oracle FnSpec for golden case `name` from functions.npz arrays."""
    fa, sa, fr, sr = CASES[name]
    par0 = fn_golden["%s/%s/par0" % (name, tag)]
    par1 = fn_golden.get("%s/%s/par1" % (name, tag))
    if fa == O.L_WEIGHTED_QUADRATIC and par1 is None:
        par1 = 1.0 / par0 ** 2  # losses.py:80-85 default weights
    return O.FnSpec(fa, par0, sa, fn_rep=fr, rep=sr, par1=par1)
