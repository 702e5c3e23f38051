"""Model selection (reference notebook section 5): the three sequential searches over e_drive,
(p_min, p_max) and e_dest, restated as host control logic on top of the device pipeline.

The reference has these loops only in its notebook (README.md:956-1316, 1435-1681, 1831-2278), each
with the whole sampling loop pasted inline.  Here the update rules are small pure functions (pinned
by the notebook's printed iterates, tests/test_oracle_kat.py) and every trial -- rebuild one table,
restart from the fixed `initial_state`, resample one day, histogram, A_drive -- is one call of an
`evaluate(params) -> Trial` callable, normally `DeviceEvaluator` (cpm_sweep on the GPU).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .api import correctparameters


@dataclass
class Params:
    e_drive: float = 0.5
    e_dest: float = 2.0
    p_min: float = 0.1
    p_max: float = 0.9

    def row(self):
        return [self.e_drive, self.e_dest, self.p_min, self.p_max]


@dataclass
class Trial:
    traffic_activity: np.ndarray            # [T] min-max normalised hourly driving totals (README.md:1106-1113)
    A_drive: float                          # share of lifetime spent driving (README.md:1649-1655)
    parking: np.ndarray | None = None       # [N,T] densities (README.md:1102)


# ------------------------------------------------------------------------------------------------
# error measures
# ------------------------------------------------------------------------------------------------
def traffic_error(measured, sampled):
    """README.md:1115-1119: mean squared error over the 24 hours."""
    measured, sampled = np.asarray(measured, float), np.asarray(sampled, float)
    err = 0.0
    for t in range(len(measured)):
        err = err + (measured[t] - sampled[t]) ** 2
    return err / len(measured)


def parking_error(measured, parking):
    """README.md:2021-2046: per-zone min-max normalisation over the day, MSE over hours for zones that are
    measured (row sum != 0) and not constant, averaged over the validated zones."""
    measured = np.asarray(measured, float)
    pk = np.array(parking, dtype=float, copy=True)
    mn, mx = pk.min(axis=1), pk.max(axis=1)
    ok = mx != mn
    pk[ok] = (pk[ok] - mn[ok, None]) / (mx[ok] - mn[ok])[:, None]
    valid = ok & (measured.sum(axis=1) != 0)
    if not valid.any():
        return math.nan
    per_zone = ((pk[valid] - measured[valid]) ** 2).sum(axis=1) * (1.0 / measured.shape[1])
    return per_zone.sum() / int(valid.sum())


# ------------------------------------------------------------------------------------------------
# update rules (pure functions; pinned by the notebook's printed iterates)
# ------------------------------------------------------------------------------------------------
def first_step(param, step_size, best_error):
    """README.md:1139 (e_drive, step 10) and README.md:2066 (e_dest, step 2): a step in an arbitrary direction."""
    return param + step_size * best_error


def secant_step(param, best_param, error, best_error, step_size):
    """README.md:1286-1299 / 2247-2260.  Returns (next_param, best_param, best_error, error_gradient)."""
    error_gradient = error - best_error
    param_gradient = param - best_param
    if error_gradient < 0:
        best_error, best_param = error, param
    else:
        param = best_param
    with np.errstate(all="ignore"):   # Julia: x/0 -> Inf/NaN, no exception
        param = float(np.float64(param) - np.float64(step_size) * np.float64(error_gradient) / np.float64(param_gradient))
    return param, best_param, best_error, error_gradient


def pminmax_step(n, p_min_prev, p_max_prev, dA_prev, A_drive_prev):
    """README.md:1466-1510: choose which bound to move, apply the proportional update, then clamp
    (the clamp is correctparameters, src/correctparameters.jl:3-23).  `n` is the notebook's 1-based
    iteration counter (starts at 2).  Returns (p_min_n, p_max_n) or None when p_min == p_max at a bound."""
    upd = (p_max_prev + p_min_prev) * dA_prev / (n * A_drive_prev)
    if p_max_prev == 1:
        if p_min_prev == p_max_prev:
            return None
        p_min_n, p_max_n = p_min_prev + upd, p_max_prev
    elif p_min_prev == 0:
        if p_min_prev == p_max_prev:
            return None
        p_min_n, p_max_n = p_min_prev, p_max_prev + upd
    elif p_min_prev == p_max_prev:
        p_min_n, p_max_n = p_min_prev + upd, p_max_prev
    else:
        p_min_n, p_max_n = p_min_prev, p_max_prev + upd
    return correctparameters(p_min_n, p_max_n, p_min_prev, p_max_prev)


# ------------------------------------------------------------------------------------------------
# the three searches
# ------------------------------------------------------------------------------------------------
@dataclass
class SearchLog:
    params: list = field(default_factory=list)
    errors: list = field(default_factory=list)
    messages: list = field(default_factory=list)


def _search_exponent(evaluate, base: Params, which: str, error_of, initial_error, alternatives, step_size, max_iter=3,
                     epsilon=0.001, good=0.02):
    """Common body of README.md:956-1316 (which='e_drive', step 10, traffic error) and
    README.md:1831-2278 (which='e_dest', step 2, parking error)."""
    log = SearchLog()
    best_param, best_error = getattr(base, which), initial_error
    param = best_param
    for alt in alternatives:                                    # Step 2: alternative starting values
        param = alt
        tr = evaluate(Params(**{**base.__dict__, which: alt}))
        err = error_of(tr)
        log.params.append(alt); log.errors.append(err)
        if err < best_error:
            best_error, best_param = err, alt
    # Step 3.1 -- the notebook steps from the LAST alternative tried, not from the best one (README.md:1139)
    param = first_step(param, step_size, best_error)
    n = 1
    while n <= max_iter:                                        # Step 3.2
        tr = evaluate(Params(**{**base.__dict__, which: param}))
        err = error_of(tr)
        log.params.append(param); log.errors.append(err)
        param, best_param, best_error, grad = secant_step(param, best_param, err, best_error, step_size)
        if best_error < good:
            log.messages.append("good solution"); break
        if abs(grad) < epsilon:
            log.messages.append("converged"); break
        n += 1
    return best_param, best_error, log


def search_e_drive(evaluate, base: Params, traffic_measured, initial_trial: Trial, alternatives=(1, 0.5)):
    """README.md:956-1316."""
    err = lambda tr: traffic_error(traffic_measured, tr.traffic_activity)
    return _search_exponent(evaluate, base, "e_drive", err, err(initial_trial), alternatives, step_size=10)


def search_e_dest(evaluate, base: Params, parking_measured, initial_trial: Trial, alternatives=(1, 0.5)):
    """README.md:1831-2278."""
    err = lambda tr: parking_error(parking_measured, tr.parking)
    return _search_exponent(evaluate, base, "e_dest", err, err(initial_trial), alternatives, step_size=2)


def search_pminmax(evaluate, base: Params, A_drive_initial, A_set=0.07, max_iter=5):
    """README.md:1435-1681.  Returns (p_min, p_max, history) with history rows (p_min, p_max, A_drive, dA)."""
    p_min, p_max = [base.p_min], [base.p_max]
    A, dA = [A_drive_initial], [A_set - A_drive_initial]
    n = 2
    while n <= max_iter:
        nxt = pminmax_step(n, p_min[-1], p_max[-1], dA[-1], A[-1])
        if nxt is None:
            break
        p_min.append(nxt[0]); p_max.append(nxt[1])
        tr = evaluate(Params(base.e_drive, base.e_dest, nxt[0], nxt[1]))
        A.append(tr.A_drive); dA.append(A_set - tr.A_drive)
        if int(A_set * 100) == int(tr.A_drive * 100):             # README.md:1661
            break
        if n > 2 and abs(dA[-1]) > abs(dA[-2]):                   # README.md:1665-1671: revert
            p_min[-1], p_max[-1] = p_min[-2], p_max[-2]
            dA[-1], A[-1] = A_set - A[-2], A[-2]
        n += 1
    return p_min[-1], p_max[-1], list(zip(p_min, p_max, A, dA))


class DeviceEvaluator:
    """evaluate(Params) -> Trial on the GPU: one cpm_sweep call per trial (or per batch via `many`)."""

    def __init__(self, city, initial_zones, *, seed=0, common_random_numbers=False, want_parking=True):
        self.city, self.initial_zones = city, np.ascontiguousarray(initial_zones, dtype=np.int64)
        self.seed, self.crn, self.want_parking = seed, common_random_numbers, want_parking
        self.calls = 0

    def many(self, params_list):
        from .engine import sweep
        tc, A, parking = sweep(self.city, self.initial_zones, [p.row() for p in params_list], seed=self.seed + self.calls,
                               common_random_numbers=self.crn, want_parking=self.want_parking)
        self.calls += 1
        out = []
        for k in range(len(params_list)):
            col = tc[:, k].astype(np.float64)
            with np.errstate(all="ignore"):
                act = (col - col.min()) / (col.max() - col.min())
            out.append(Trial(act, float(A[k]), None if parking is None else parking[:, :, k]))
        return out

    def __call__(self, params: Params) -> Trial:
        return self.many([params])[0]
