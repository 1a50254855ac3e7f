"""TEST INFRASTRUCTURE ONLY -- Kaufman's exact solution of the 2D Ising model on an m x n torus.

The reference has no implementation of it (it reads a data file that is not in the repository,
kauffman_specific_heat.ipynb:89), so the published formula is restated here:

    Z = 1/2 (2 sinh 2K)^(mn/2) [ prod_r 2cosh(m g_{2r+1}/2) + prod_r 2sinh(m g_{2r+1}/2)
                                + prod_r 2cosh(m g_{2r}/2)   + prod_r 2sinh(m g_{2r}/2) ],  r = 0..n-1
    cosh g_l = cosh 2K coth 2K - cos(l pi / n)  (l >= 1),   g_0 = 2K + ln tanh K  (signed)

(B. Kaufman, Phys. Rev. 76, 1232 (1949)).  E/N = -d lnZ/dK / N, C/N = K^2 d^2 lnZ/dK^2 / N, J = k_B = 1.
Derivatives are taken with float64 autograd, so E and C are exact to rounding.  Pinned in
tests/test_oracle.py against brute-force enumeration (L = 2, 3, 4) and the SURVEY App. D table.

Only tests/, __graft_entry__.smoke() and bench.py may import this.
"""
from __future__ import annotations

import itertools

import numpy as np
import torch


def _log_z(K: torch.Tensor, m: int, n: int) -> torch.Tensor:
    two_k = 2.0 * K
    c = torch.cosh(two_k) * torch.cosh(two_k) / torch.sinh(two_k)          # cosh 2K coth 2K
    l = torch.arange(0, 2 * n, dtype=torch.float64)
    ch = c - torch.cos(l * np.pi / n)
    g = torch.acosh(torch.clamp(ch, min=1.0))
    g0 = two_k + torch.log(torch.tanh(K))
    g = torch.cat([g0.reshape(1), g[1:]])
    odd, even = g[1::2], g[0::2]
    half_m = 0.5 * m

    def log2cosh(x):
        ax = torch.abs(x)
        return ax + torch.log1p(torch.exp(-2.0 * ax))

    def log2sinh_abs(x):
        ax = torch.abs(x)
        return ax + torch.log(-torch.expm1(-2.0 * ax))

    t1, t2 = log2cosh(half_m * odd).sum(), log2sinh_abs(half_m * odd).sum()
    t3 = log2cosh(half_m * even).sum()
    r4 = log2sinh_abs(half_m * even[1:]).sum()
    x0 = half_m * even[0]
    # g_0 goes through zero at T_c: its factor 2 sinh(m g_0 / 2) is kept in linear space there (in log space the
    # value is right but the derivatives are not: ln|sinh| is singular at 0), in signed log space elsewhere
    if abs(float(x0.detach())) < 1.0:
        mx = max(float(t1.detach()), float(t2.detach()), float(t3.detach()), float(r4.detach()))
        s = torch.exp(t1 - mx) + torch.exp(t2 - mx) + torch.exp(t3 - mx) + 2.0 * torch.sinh(x0) * torch.exp(r4 - mx)
    else:
        t4 = r4 + log2sinh_abs(x0)
        mx = max(float(t1.detach()), float(t2.detach()), float(t3.detach()), float(t4.detach()))
        s = torch.exp(t1 - mx) + torch.exp(t2 - mx) + torch.exp(t3 - mx) + float(torch.sign(x0.detach())) * torch.exp(t4 - mx)
    return -np.log(2.0) + 0.5 * m * n * torch.log(2.0 * torch.sinh(two_k)) + mx + torch.log(s)


def energy_and_specific_heat(T: float, L: int, m: int | None = None):
    """(E/N, C/N) of the L x L (or m x L) periodic lattice at temperature T."""
    m = L if m is None else m
    K = torch.tensor(1.0 / T, dtype=torch.float64, requires_grad=True)
    lnz = _log_z(K, m, L)
    (d1,) = torch.autograd.grad(lnz, K, create_graph=True)
    (d2,) = torch.autograd.grad(d1, K)
    N = m * L
    return float(-d1.detach() / N), float(K.detach() ** 2 * d2.detach() / N)


def brute_force(T: float, L: int):
    """(E/N, C/N) by enumeration of all 2^(L*L) states (L <= 4)."""
    N = L * L
    K = 1.0 / T
    idx = np.arange(N).reshape(L, L)
    right, down = np.roll(idx, -1, axis=1), np.roll(idx, -1, axis=0)
    if L == 2:   # both neighbours in a direction coincide: each pair is coupled twice, like the torus formula
        pass
    es = []
    for bits in itertools.product((-1, 1), repeat=N):
        s = np.array(bits)
        es.append(-(s[idx] * s[right]).sum() - (s[idx] * s[down]).sum())
    es = np.array(es, dtype=np.float64)
    w = np.exp(-K * (es - es.min()))
    Z = w.sum()
    e1 = (w * es).sum() / Z
    e2 = (w * es * es).sum() / Z
    return e1 / N, K * K * (e2 - e1 * e1) / N
