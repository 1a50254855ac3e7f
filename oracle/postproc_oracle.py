"""TEST INFRASTRUCTURE ONLY -- numpy/pure-Python restatement of the reference's post-processing.

Restates, from their behaviour, the Python steps that follow the worm loop:

  image_from_bonds      bonds.py:199-239 (parity filter :212) + bonds.py:241-319 (rasterise)
  block_config          block_configs.py:6-71 and iterated_blocking.py:6-59 (one RG step; recursion :56-57)
  block_config_T        bonds.py:333-410 (the Bonds-class variant with block_val)
  count_bonds           count_bonds.py:107-139
  block_resampling      utils.py:90-104 (KFold training folds = delete-one-block samples)
  jackknife_err         utils.py:123-135
  count_bonds_with_err  count_bonds.py:141-172
  avg_energies / specific_heat_fd   specific_heat.py:75-113 / :115-169

Parity status: PINNED for blocking and counting -- tests/test_oracle.py compares these against
tests/golden/post_*.npz, which oracle/make_golden.py produced by importing the reference's own
block_configs._block_config, iterated_blocking._block_config, CountBonds._count_bonds,
utils.block_resampling/jackknife_err and Bonds._set_config_data from /root/reference (with the
import shims of SURVEY App. F; no source edits).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

import numpy as np


# ---------------------------------------------------------------------------------------------
# imaging
# ---------------------------------------------------------------------------------------------
def image_from_bonds(bonds, L: int) -> np.ndarray:
    """[2N] occupations -> [2L, 2L] 0/1 image, first index = x pixel (bonds.py:241-319).

    Deliberately a per-bond loop (like the reference's) so it is an independent statement of the
    rule; bond b belongs to site i = b // 2 (x = i % L, y = i // L), even b = +x bond, odd b = +y
    bond (worm_ising_2d.cpp:214-244).  A wrapping bond lands on pixel 2L-1 (bonds.py:271-274,292-295)."""
    bonds = np.asarray(bonds).reshape(-1)
    assert bonds.size == 2 * L * L
    img = np.zeros((2 * L, 2 * L), dtype=np.int64)
    for b in range(2 * L * L):
        if bonds[b] % 2 != 1:                    # bonds.py:212
            continue
        i = b // 2
        x, y = i % L, i // L
        if b % 2 == 0:
            x2 = (x + 1) % L
            img[2 * x, 2 * y] = 1
            img[2 * x2, 2 * y] = 1
            # sorted endpoints: a wrap bond is ((0,y),(L-1,y)) -> pixel 2*(L-1)+1   (bonds.py:270-278)
            img[(2 * max(x, x2) + 1) if abs(x2 - x) == L - 1 else (x + x2), 2 * y] = 1
        else:
            y2 = (y + 1) % L
            img[2 * x, 2 * y] = 1
            img[2 * x, 2 * y2] = 1
            img[2 * x, (2 * max(y, y2) + 1) if abs(y2 - y) == L - 1 else (y + y2)] = 1    # bonds.py:291-299
    return img


# ---------------------------------------------------------------------------------------------
# blocking
# ---------------------------------------------------------------------------------------------
def block_config(config, num_block_steps: int = 1, double_bonds_value=0) -> np.ndarray:
    """One (or k) RG blocking step(s), flattened output (block_configs.py:6-71).

    `double_bonds_value` 0/None = the "1+1=0" scheme of both modules; a non-zero value follows
    block_configs.py:41-56 (iterated_blocking.py's non-None branch reads unassigned variables,
    SURVEY App. G, and is not restated)."""
    a = np.asarray(config)
    L = int(np.sqrt(a.size) / 2)
    a = a.reshape(2 * L, 2 * L)
    v = 0 if double_bonds_value is None else double_bonds_value
    out = np.zeros((L, L), dtype=np.int64)
    h = L // 2
    for p in range(h):
        for q in range(h):
            ex = (a[4 * p, 4 * q + 3], a[4 * p + 2, 4 * q + 3])       # block_configs.py:35
            ey = (a[4 * p + 3, 4 * q], a[4 * p + 3, 4 * q + 2])       # :36
            if v == 0:
                xa = int(ex[0]) ^ int(ex[1])                            # :38
                ya = int(ey[0]) ^ int(ey[1])                            # :39
            else:
                sx, sy = int(ex[0]) + int(ex[1]), int(ey[0]) + int(ey[1])
                xa = v if sx == 2 else (1 if sx == 1 else 0)            # :46-50
                ya = v if sy == 2 else (1 if sy == 1 else 0)            # :51-55
            out[2 * p, 2 * q] = xa or ya                                # :56-57
            out[2 * p, 2 * q + 1] = xa                                  # :58
            out[2 * p + 1, 2 * q] = ya                                  # :59
    for p in range(h):                                                  # :61-65 (final pass decides)
        for q in range(h):
            if out[2 * p, 2 * q - 1] or out[2 * p - 1, 2 * q]:
                out[2 * p, 2 * q] = 1
    if v == 0 and num_block_steps > 1:                                  # :66-69
        return block_config(out.flatten(), num_block_steps - 1, 0)
    return out.flatten()


def block_config_T(config, block_val: int = 0) -> np.ndarray:
    """bonds.py:333-410 -- the Bonds-class blocking variant, [2L,2L] -> [L,L]."""
    a = np.asarray(config)
    L = int(np.sqrt(a.size) / 2)
    a = a.reshape(2 * L, 2 * L)
    out = np.zeros((L, L), dtype=np.int64)
    h = L // 2
    for p in range(h):
        for q in range(h):
            ex = [int(a[4 * p, 4 * q + 3]), int(a[4 * p + 2, 4 * q + 3])]
            ey = [int(a[4 * p + 3, 4 * q]), int(a[4 * p + 3, 4 * q + 2])]
            xa, ya = ex[0] ^ ex[1], ey[0] ^ ey[1]
            out[2 * p, 2 * q] = xa or ya
            out[2 * p, 2 * q + 1] = xa
            out[2 * p + 1, 2 * q] = ya
            if block_val != 0:
                if ex == [1, 1]:
                    out[2 * p, 2 * q] = block_val
                    out[2 * p, 2 * q + 1] = block_val
                if ey == [1, 1]:
                    out[2 * p, 2 * q] = block_val
                    out[2 * p + 1, 2 * q] = block_val
    for p in range(h):
        for q in range(h):
            xs, ys = out[2 * p, 2 * q - 1], out[2 * p - 1, 2 * q]
            if xs == 1 or ys == 1:
                out[2 * p, 2 * q] = 1
            if block_val != 0 and (xs == block_val or ys == block_val):
                out[2 * p, 2 * q] = block_val
    return out


# ---------------------------------------------------------------------------------------------
# counting and statistics
# ---------------------------------------------------------------------------------------------
def bond_counts(data, width: int) -> np.ndarray:
    """per-config n = sum of pixels with (i + j) odd (count_bonds.py:110-117)."""
    d = np.asarray(data).reshape(-1, width, width)
    out = []
    for cfg in d:
        s = 0
        for i in range(width):
            for j in range(width):
                if (i + j) % 2 == 1:
                    s += int(cfg[i, j])
        out.append(s)
    return np.array(out, dtype=np.int64)


def count_bonds(data, width: int):
    """(<n>, <n^2> - <n>^2) (count_bonds.py:119-123,139)."""
    bc = bond_counts(data, width)
    nb_avg = np.mean(bc)
    nb2_avg = np.mean(bc ** 2)
    return nb_avg, nb2_avg - nb_avg ** 2


def block_resampling(data, num_blocks: int):
    """utils.py:90-104: sklearn KFold(n_splits) training folds, i.e. delete-one-block samples.
    KFold without shuffling makes the first n % k folds one longer."""
    data = np.asarray(data)
    n = data.shape[0]
    sizes = np.full(num_blocks, n // num_blocks, dtype=int)
    sizes[: n % num_blocks] += 1
    out, start = [], 0
    for s in sizes:
        keep = np.ones(n, dtype=bool)
        keep[start:start + s] = False
        out.append(data[keep])
        start += s
    return out


def jackknife_err(y_i, y_full, num_blocks: int):
    """utils.py:123-135."""
    y_i = np.asarray(y_i)
    return np.sqrt((num_blocks - 1) * np.sum((y_i - y_full) ** 2) / num_blocks)


def count_bonds_with_err(data, width: int, num_blocks: int = 10):
    """count_bonds.py:141-172 -> ((<n>, var), [err_n, err_var])."""
    full = count_bonds(data, width)
    rs = np.array([count_bonds(blk, width) for blk in block_resampling(np.asarray(data), num_blocks)])
    err = [jackknife_err(rs[:, k], full[k], num_blocks) for k in range(2)]
    return full, err


def avg_energy_with_err(values, num_blocks: int):
    """specific_heat.py:93-107: mean and jackknife error of the rows for one temperature."""
    values = np.asarray(values, dtype=np.float64)
    m = np.mean(values)
    rs = [np.mean(b) for b in block_resampling(values, num_blocks)]
    return m, jackknife_err(rs, m, num_blocks)


def specific_heat_fd(T, E):
    """specific_heat.py:115-169: mean of six |dE|/dT stencils; first and last two T dropped."""
    T = list(T)
    E = list(E)
    ts, cs = [], []
    for i in range(2, len(E)):
        if i in (len(E) - 1, len(E) - 2):
            continue
        d1 = abs(E[i + 1] - E[i]) / (T[i + 1] - T[i])
        d2 = abs(E[i] - E[i - 1]) / (T[i] - T[i - 1])
        d3 = abs(E[i + 1] - E[i - 1]) / (T[i + 1] - T[i - 1])
        d4 = abs(E[i + 2] - E[i - 2]) / (T[i + 2] - T[i - 2])
        d5 = abs(E[i + 2] - E[i]) / (T[i + 2] - T[i])
        d6 = abs(E[i + 2] - E[i - 1]) / (T[i + 2] - T[i - 1])
        ts.append(T[i])
        cs.append((d1 + d2 + d3 + d4 + d5 + d6) / 6)
    return np.array(ts), np.array(cs)
