"""Second, independent restatement of the reference hot path in pure Python (loops), used ONLY by
tests to cross-check the C oracle on tiny cases.  TEST INFRASTRUCTURE, not the product.

Written directly from the Julia sources (cited per function), not from cpm_oracle.c: arrays are numpy
Fortran-ordered and indexed with Julia's 1-based indices through the `J` helper so each line can be
read against the Julia line it restates.  Variates come from a `rand()` callable in the reference's
draw order; the truncated normal consumes one uniform by inverse CDF (scipy.special), because
Distributions.jl is not available (parity unpinned, see oracle/cpm_oracle.c header).
"""
from __future__ import annotations

import math

import numpy as np
from scipy import special


class J:
    """1-based view of a numpy array."""

    def __init__(self, a):
        self.a = a

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        return self.a[tuple(i - 1 for i in idx)]

    def __setitem__(self, idx, v):
        if not isinstance(idx, tuple):
            idx = (idx,)
        self.a[tuple(i - 1 for i in idx)] = v


def _pow(x, e):
    # Julia: Float64^Int special-cases 2 -> x*x (base/math.jl); otherwise libm pow.
    if isinstance(e, int) or float(e).is_integer():
        e = int(e)
        if e == 2:
            return x * x
        if e == 1:
            return x
        if e == 3:
            return x * x * x
        if e == 0:
            return 1.0
    if math.isnan(x):
        return math.nan
    return math.pow(x, e)


def _maximum(v):
    return math.nan if any(math.isnan(x) for x in v) else max(v)


def _minimum(v):
    return math.nan if any(math.isnan(x) for x in v) else min(v)


def _div(a, b):
    try:
        return a / b
    except ZeroDivisionError:
        if a == 0 or math.isnan(a):
            return math.nan
        return math.copysign(math.inf, a)


def createpdrive(datamatrix, distance_matrix_km, number_zones, T, p_min, p_max, e_drive):
    """src/createpdrive.jl:3-38"""
    N = number_zones
    dm, dist = J(datamatrix), J(distance_matrix_km)
    p_drive = np.zeros((N, T), order="F")
    mean_sum = np.zeros((N, T), order="F")
    pd, ms = J(p_drive), J(mean_sum)
    min_t, max_t = [0.0] * (N + 1), [0.0] * (N + 1)
    for i in range(1, N + 1):
        for t in range(1, T + 1):
            acc, counter = 0.0, 0
            for j in range(1, N + 1):
                if dm[i, j, t, 1] != 0:
                    acc = acc + float(dm[i, j, t, 1]) / float(dist[i, j])
                    counter += 1
            ms[i, t] = _div(acc, counter)
        row = [float(ms[i, t]) for t in range(1, T + 1)]
        max_t[i], min_t[i] = _maximum(row), _minimum(row)
    for i in range(1, N + 1):
        if max_t[i] > 0:
            for t in range(1, T + 1):
                pd[i, t] = p_min + (p_max - p_min) * _pow(_div(float(ms[i, t]) - min_t[i], max_t[i] - min_t[i]), e_drive)
    return p_drive


def createpdestin(datamatrix, number_zones, T, e_dest):
    """src/createpdestin.jl:3-50"""
    N = number_zones
    dm = J(datamatrix)
    p_dest = np.zeros((N, N, T), order="F")
    p = J(p_dest)
    for i in range(1, N + 1):
        for j in range(1, N + 1):
            col = [float(dm[i, j, t, 1]) for t in range(1, T + 1)]
            mx, mn = _maximum(col), _minimum(col)
            if mx > 0:
                for t in range(1, T + 1):
                    p[i, j, t] = _pow(_div(col[t - 1] - mn, mx - mn), e_dest)
    for i in range(1, N + 1):
        for t in range(1, T + 1):
            nf = 0.0
            for j in range(1, N + 1):
                nf += float(p[i, j, t])
            if nf > 0:
                for j in range(1, N + 1):
                    p[i, j, t] = float(p[i, j, t]) / nf
    return p_dest


def initializestates(C, T, cars_per_zone):
    """src/initializestates.jl:4-22"""
    state = np.zeros((C, T), dtype=np.int64, order="F")
    trans = np.zeros((C, T, 4), dtype=np.float64, order="F")
    zone = 0
    for i in range(0, C - cars_per_zone + 1, cars_per_zone):
        zone += 1
        for j in range(1, cars_per_zone + 1):
            state[i + j - 1, 0] = zone
    return state, trans


def _phases12(state, trans, p_drive, p_dest, C, N, t, rand):
    """src/resampling.jl:11-49 (identical to src/solveinitialvalueproblem.jl:11-49)"""
    S, TR, PD, PX = J(state), J(trans), J(p_drive), J(p_dest)
    for i in range(1, C + 1):
        origin = int(S[i, t])
        if origin < 1:  # documented extension: lost car
            TR[i, t, 1] = 0
            TR[i, t, 2] = origin
            continue
        u = rand()
        if u <= PD[origin, t]:
            drive = 1
        else:
            drive = 0
            TR[i, t, 2] = origin
        TR[i, t, 1] = drive
    for i in range(1, C + 1):
        if TR[i, t, 1] == 1:
            u = rand()
            range_up, range_low, destination = 0.0, 0.0, 0
            origin = int(S[i, t])
            distribution = [float(PX[origin, j, t]) for j in range(1, N + 1)]
            if sum(distribution) == 0:
                destination = origin
            else:
                for j in range(1, N + 1):
                    range_up = range_up + distribution[j - 1]
                    if range_low < u <= range_up:
                        destination = j
                        break
                    range_low = range_up
            TR[i, t, 2] = destination


def truncnormal(mu, sigma, u):
    if not (sigma > 0) or not math.isfinite(sigma):
        return mu
    b = 0.1 * mu / sigma
    z = math.sqrt(2.0) * float(special.erfinv((2.0 * u - 1.0) * float(special.erf(b / math.sqrt(2.0)))))
    return min(max(mu + sigma * z, 0.9 * mu), 1.1 * mu)


def solveinitialvalueproblem(state, trans, p_drive, p_dest, C, N, T, rand):
    """src/solveinitialvalueproblem.jl:4-62"""
    for t in range(1, T):
        _phases12(state, trans, p_drive, p_dest, C, N, t, rand)
        state[:, t] = trans[:, t - 1, 1].astype(np.int64)
    return state[:, T - 1].copy()


def resampling(state, trans, C, N, T, p_drive, p_dest, datamatrix, distance_matrix_km, rand):
    """src/resampling.jl:3-89"""
    S, TR, DM, DK = J(state), J(trans), J(datamatrix), J(distance_matrix_km)
    for t in range(1, T + 1):
        _phases12(state, trans, p_drive, p_dest, C, N, t, rand)
        for i in range(1, C + 1):
            if TR[i, t, 1] == 1:
                origin, destination = int(S[i, t]), int(TR[i, t, 2])
                if origin == destination:
                    TR[i, t, 3] = 5 * 60
                    TR[i, t, 4] = 1
                elif destination >= 1:
                    mean, sd = float(DM[origin, destination, t, 1]), float(DM[origin, destination, t, 2])
                    if sd == 0:
                        sd = 0.1 * mean
                    TR[i, t, 3] = truncnormal(mean, sd, rand())
                    mean = float(DK[origin, destination])
                    TR[i, t, 4] = truncnormal(mean, 0.1 * mean, rand())
        if t < T:
            state[:, t] = trans[:, t - 1, 1].astype(np.int64)
    return state, trans


def averagedrivingtime(C, T, A_drive, trans):
    """src/averagedrivingtime.jl:3-12"""
    s = 0.0
    for t in range(T):
        s = s + float(np.sum(trans[:, t, 2]))
    return A_drive + s / (C * T * 60 * 60)


def correctparameters(p_min_next, p_max_next, p_min, p_max):
    """src/correctparameters.jl:3-23"""
    if p_min_next < 0:
        p_min_next = 0
    elif p_min_next > p_max:
        p_min_next = p_max
    if p_max_next > 1:
        p_max_next = 1
    elif p_max_next < p_min:
        p_max_next = p_min
    return p_min_next, p_max_next


def zonehistogram(N, T, C, state, trans):
    """src/saveresults.jl:6-28"""
    driving = np.zeros((N, T), order="F")
    parking = np.zeros((N, T), order="F")
    for i in range(C):
        for t in range(T):
            index = int(state[i, t])
            if index < 1:
                continue
            parking[index - 1, t] += 1
            if trans[i, t, 0] == 1:
                driving[index - 1, t] += 1
    parking = parking / C
    traffic = driving.sum(axis=0)
    mn, mx = traffic.min(), traffic.max()
    with np.errstate(all="ignore"):
        traffic = (traffic - mn) / (mx - mn)
    return parking, driving, traffic


def progress_percent(t):
    """trunc(Int, t/0.23): src/resampling.jl:84, src/solveinitialvalueproblem.jl:54"""
    return int(t / 0.23)
