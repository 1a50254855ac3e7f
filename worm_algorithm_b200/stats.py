"""Error bars and estimators on the reduced sums the kernels return (float64 on the host: a few
numbers per temperature).

`block_resampling` / `jackknife_err` are the reference's error model (utils.py:90-104, 123-135):
delete-one-block samples in sklearn-KFold order and sqrt((nb-1)/nb * sum (y_i - y)^2).  The
fluctuation estimators are SURVEY App. D (the reference itself only has the finite-difference C_V
of specific_heat.py:115-169, mirrored in specific_heat.py here).
"""
from __future__ import annotations

import numpy as np


def block_resampling(data, num_blocks: int):
    """The `num_blocks` delete-one-block copies of `data` (first axis = samples), in the order
    KFold(n_splits=num_blocks) without shuffling produces its training folds: the held-out blocks
    are contiguous and the first `n % num_blocks` of them are one sample longer."""
    data = np.asarray(data)
    n = data.shape[0]
    if n < 1:
        raise ValueError("Data must have at least one sample.")
    if num_blocks < 1:
        raise ValueError("Number of resampled blocks must be greater than or equal to 1.")
    if num_blocks < 2 or num_blocks > n:
        raise ValueError("num_blocks must be in [2, number of samples]")
    base, extra = divmod(n, num_blocks)
    out, lo = [], 0
    for b in range(num_blocks):
        hi = lo + base + (1 if b < extra else 0)
        out.append(np.concatenate([data[:lo], data[hi:]], axis=0))
        lo = hi
    return out


def jackknife_err(y_i, y_full, num_blocks: int):
    y_i = np.asarray(y_i, dtype=np.float64)
    y_full = np.asarray(y_full, dtype=np.float64)
    return np.sqrt((num_blocks - 1) * np.sum((y_i - y_full) ** 2) / num_blocks)


# ---------------------------------------------------------------------------------------------
# estimators from the WORM_ACC_* sums (rows: Z, sum Nb, sum Nb^2, sum n, sum n^2, iterations)
# ---------------------------------------------------------------------------------------------
def taylor_observables(acc, T, N):
    """Taylor chain: E/N = -T<Nb>/N (worm_ising_2d.cpp:185), C/N = (<Nb^2> - <Nb>^2 - <Nb>)/N,
    Z_av = Z / iterations = 1/chi (:183)."""
    a = np.asarray(acc, dtype=np.float64)
    Z = np.maximum(a[..., 0], 1.0)
    nb, nb2 = a[..., 1] / Z, a[..., 2] / Z
    T = np.asarray(T, dtype=np.float64)
    return {"E": -T * nb / N, "C": (nb2 - nb * nb - nb) / N, "Z_av": a[..., 0] / np.maximum(a[..., 5], 1.0),
            "Nb": nb, "n": a[..., 3] / Z, "var_n": a[..., 4] / Z - (a[..., 3] / Z) ** 2}


def binary_observables(acc, T, N):
    """Binary (tanh) chain, n = occupied bonds, t = tanh K, s = (1 - t^2)/t:
    E/N = -(2t + <n> s / N),  C/N = K^2/N [2N(1 - t^2) + var(n) s^2 - <n>(1 - t^2)(1 + t^2)/t^2]."""
    a = np.asarray(acc, dtype=np.float64)
    Z = np.maximum(a[..., 0], 1.0)
    n, n2 = a[..., 3] / Z, a[..., 4] / Z
    K = 1.0 / np.asarray(T, dtype=np.float64)
    t = np.tanh(K)
    s = (1.0 - t * t) / t
    var = n2 - n * n
    C = K * K / N * (2.0 * N * (1.0 - t * t) + var * s * s - n * (1.0 - t * t) * (1.0 + t * t) / (t * t))
    return {"E": -(2.0 * t + n * s / N), "C": C, "Z_av": a[..., 0] / np.maximum(a[..., 5], 1.0), "n": n, "var_n": var}


def replica_mean_err(values, axis=0):
    """Mean and standard error over independent replicas (the valid error model for the 3-sigma
    tests: seeds are independent chains)."""
    v = np.asarray(values, dtype=np.float64)
    r = v.shape[axis]
    return v.mean(axis=axis), v.std(axis=axis, ddof=1) / np.sqrt(r)


def pooled_jackknife(acc, T, N, algo=0):
    """Estimators from the sums POOLED over replicas, with delete-one-replica jackknife errors and
    bias correction.  acc: [replicas, ..., 8] per-chain WORM_ACC_* sums, T broadcastable to acc[0, ..., 0].
    Pooling matters for the fluctuation estimators: <Nb^2> - <Nb>^2 evaluated per chain is biased low
    by O(tau_int / chain length), evaluated on the pooled sums by 1/replicas of that.
    Returns {name: (value, error)}."""
    a = np.asarray(acc, dtype=np.float64)
    R = a.shape[0]
    if R < 2:
        raise ValueError("need at least two replicas")
    f = taylor_observables if algo == 0 else binary_observables
    tot = a.sum(axis=0)
    full = f(tot, T, N)
    loo = f(tot[None] - a, T, N)                     # delete-one-replica sums
    out = {}
    for k, v in full.items():
        th = loo[k]
        mean = th.mean(axis=0)
        err = np.sqrt((R - 1) / R * ((th - mean) ** 2).sum(axis=0))
        out[k] = (R * v - (R - 1) * mean, err)
    return out
