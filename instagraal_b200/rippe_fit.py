"""Host-side p(s) model fitting used by the sampler facade.

These are the reference's own scipy calls (out of scope for native work),
restated so the facade is self-contained on machines that do not have the reference package:
``optim_rippe_curve_update.py`` lines 9 (d = 2), 21-31 (peval), 34-48 (log_residuals),
64-106 (estimate_param_rippe), 109-149 (residual_4_max_dist, estimate_max_dist_intra[_nuis]).
"""
from __future__ import annotations

import warnings

import numpy as np
from scipy.optimize import fsolve, leastsq

d = 2


def peval(x, param):
    return param[3] * (0.53 * (param[0] ** -3.0) * np.power((param[1] * x / param[0]), (param[2]))
                       * np.exp((d - 2) / (np.power((param[1] * x / param[0]), 2) + d)))


def log_residuals(p, y, x):
    kuhn, lm, slope, A = p
    with np.errstate(invalid="ignore", divide="ignore"):
        rippe = (np.log(A) + np.log(0.53) - 3 * np.log(kuhn) + slope * (np.log(lm * x / kuhn))
                 + (d - 2) / (np.power((lm * x / kuhn), 2) + d))
    return y - rippe


def estimate_param_rippe(y_meas, x_bins):
    kuhn, lm, slope = 50, 9.6, -1.5
    A = np.max(y_meas)
    p0 = [kuhn, lm, slope, A]
    plsq = leastsq(log_residuals, p0, args=(np.log(y_meas / 7.0), x_bins))
    y_estim = peval(x_bins, plsq[0])
    kuhn_x, lm_x, slope_x, A_x = plsq[0]
    plsq_out = [kuhn_x, lm_x, slope_x, d, A_x]
    if np.any(np.isnan(np.array(plsq_out))) or slope_x >= 0:
        A = np.max(y_meas)
        test = peval(x_bins, [kuhn, lm, slope, A])
        new_A = y_meas[0] * A / test.max()
        plsq_out = [kuhn, lm, slope, d, A * new_A]
        y_estim = peval(x_bins, [kuhn, lm, slope, new_A])
    return plsq_out, y_estim


def residual_4_max_dist(x, p):
    kuhn, lm, slope, dd, A, y = p
    x[np.isnan(x)] = 0
    x = np.abs(x)
    rippe = A * (0.53 * (kuhn ** -3.0) * np.power((lm * x / kuhn), slope)
                 * np.exp((dd - 2) / (np.power((lm * x / kuhn), 2) + dd)))
    return np.abs(y - rippe)


def estimate_max_dist_intra(p, val_inter):
    kuhn, lm, slope, dd, A = p
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        x = fsolve(residual_4_max_dist, 500, args=([kuhn, lm, slope, dd, A, val_inter]))
    return np.abs(x[0])


def estimate_max_dist_intra_nuis(p, val_inter, old_s):
    kuhn, lm, slope, dd, A = p
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        x = fsolve(residual_4_max_dist, old_s, args=([kuhn, lm, slope, dd, A, val_inter]))
    return x[0]
