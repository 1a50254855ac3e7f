"""Kaufman's exact solution of the 2D Ising model on an m x n torus, and the finite-size-scaling driver.

The reference compares its C_V(T) with an exact curve read from `../data/Cvexact32.dat`
(kauffman_specific_heat.ipynb:89), a file that is not in the repository; this module
computes that curve for any lattice:

    Z = 1/2 (2 sinh 2K)^(mn/2) [ prod_r 2cosh(m g_{2r+1}/2) + prod_r 2sinh(m g_{2r+1}/2)
                                + prod_r 2cosh(m g_{2r}/2)   + prod_r 2sinh(m g_{2r}/2) ],  r = 0..n-1
    cosh g_l = cosh 2K coth 2K - cos(l pi / n)  (l >= 1),   g_0 = 2K + ln tanh K  (signed)

(B. Kaufman, Phys. Rev. 76, 1232 (1949)); E/N = -d lnZ/dK / N, C/N = K^2 d^2 lnZ/dK^2 / N, J = k_B = 1.
ln Z is evaluated on second-order jets (value, d/dK, d^2/dK^2 propagated through every elementary
operation in float64), so E and C are exact to rounding -- no finite differences.  Host arithmetic on a
handful of numbers; the Monte-Carlo side of the comparison is the CUDA path.
"""
from __future__ import annotations

import numpy as np


class _Jet:
    """f, f', f'' of a function of one variable (arrays broadcast like numpy)."""
    __slots__ = ("v", "d", "dd")

    def __init__(self, v, d=0.0, dd=0.0):
        self.v = np.asarray(v, dtype=np.float64)
        self.d = np.asarray(d, dtype=np.float64) + 0.0 * self.v
        self.dd = np.asarray(dd, dtype=np.float64) + 0.0 * self.v

    @staticmethod
    def lift(x):
        return x if isinstance(x, _Jet) else _Jet(x)

    def chain(self, f, f1, f2):
        """g = F(self) with F' = f1, F'' = f2 evaluated at self.v."""
        return _Jet(f, f1 * self.d, f2 * self.d * self.d + f1 * self.dd)

    def __add__(self, o):
        o = _Jet.lift(o)
        return _Jet(self.v + o.v, self.d + o.d, self.dd + o.dd)
    __radd__ = __add__

    def __neg__(self):
        return _Jet(-self.v, -self.d, -self.dd)

    def __sub__(self, o):
        return self + (-_Jet.lift(o))

    def __rsub__(self, o):
        return _Jet.lift(o) + (-self)

    def __mul__(self, o):
        o = _Jet.lift(o)
        return _Jet(self.v * o.v, self.d * o.v + self.v * o.d, self.dd * o.v + 2.0 * self.d * o.d + self.v * o.dd)
    __rmul__ = __mul__

    def __truediv__(self, o):
        o = _Jet.lift(o)
        return self * o.chain(1.0 / o.v, -1.0 / o.v ** 2, 2.0 / o.v ** 3)

    def sum(self):
        return _Jet(self.v.sum(), self.d.sum(), self.dd.sum())


def _cosh(x): return x.chain(np.cosh(x.v), np.sinh(x.v), np.cosh(x.v))
def _sinh(x): return x.chain(np.sinh(x.v), np.cosh(x.v), np.sinh(x.v))
def _log(x): return x.chain(np.log(x.v), 1.0 / x.v, -1.0 / x.v ** 2)
def _exp(x): return x.chain(np.exp(x.v), np.exp(x.v), np.exp(x.v))


def _tanh(x):
    t = np.tanh(x.v)
    return x.chain(t, 1.0 - t * t, -2.0 * t * (1.0 - t * t))


def _acosh(x):
    s = np.sqrt(x.v * x.v - 1.0)
    return x.chain(np.arccosh(x.v), 1.0 / s, -x.v / s ** 3)


def _log_2cosh(x):
    """ln(2 cosh x), stable for large |x|."""
    t = np.tanh(x.v)
    return x.chain(np.abs(x.v) + np.log1p(np.exp(-2.0 * np.abs(x.v))), t, 1.0 - t * t)


def _log_2sinh_abs(x):
    """ln(2 |sinh x|)."""
    c = 1.0 / np.tanh(x.v)
    return x.chain(np.abs(x.v) + np.log(-np.expm1(-2.0 * np.abs(x.v))), c, 1.0 - c * c)


def log_partition_jet(K: float, m: int, n: int):
    """(ln Z, d ln Z / dK, d^2 ln Z / dK^2) of the m x n torus."""
    k = _Jet(float(K), 1.0, 0.0)
    two_k = 2.0 * k
    c = _cosh(two_k) * _cosh(two_k) / _sinh(two_k)                    # cosh 2K coth 2K
    l = np.arange(1, 2 * n, dtype=np.float64)
    g = _acosh(c + _Jet(-np.cos(l * np.pi / n)))                      # g_1 .. g_{2n-1}
    g0 = two_k + _log(_tanh(k))                                        # signed
    half_m = 0.5 * m
    odd = _Jet(g.v[0::2], g.d[0::2], g.dd[0::2])                       # l = 1, 3, ...
    even_rest = _Jet(g.v[1::2], g.d[1::2], g.dd[1::2])                 # l = 2, 4, ...
    t1 = _log_2cosh(half_m * odd).sum()
    t2 = _log_2sinh_abs(half_m * odd).sum()
    t3 = _log_2cosh(half_m * even_rest).sum() + _log_2cosh(half_m * g0)
    r4 = _log_2sinh_abs(half_m * even_rest).sum()
    x0 = half_m * g0
    # g_0 changes sign at T_c (the term's factor 2 sinh(m g_0 / 2) goes through zero there, smoothly): near the
    # zero that factor is kept in linear space, away from it in log space with its sign
    near = abs(float(x0.v)) < 1.0
    if near:
        f0 = 2.0 * _sinh(x0)
        terms = [t1, t2, t3]
        mx = max(max(float(t.v) for t in terms), float(r4.v))
        tot = f0 * _exp(r4 - mx)
    else:
        t4 = r4 + _log_2sinh_abs(x0)
        terms = [t1, t2, t3]
        mx = max(max(float(t.v) for t in terms), float(t4.v))
        tot = float(np.sign(x0.v)) * _exp(t4 - mx)
    for t in terms:
        tot = tot + _exp(t - mx)
    lnz = (-np.log(2.0)) + (0.5 * m * n) * _log(2.0 * _sinh(two_k)) + mx + _log(tot)
    return float(lnz.v), float(lnz.d), float(lnz.dd)


def energy_and_specific_heat(T: float, L: int, m: int | None = None):
    """(E/N, C/N) of the L x L (or m x L) periodic lattice at temperature T, exact to rounding."""
    m = L if m is None else int(m)
    K = 1.0 / float(T)
    _, d1, d2 = log_partition_jet(K, m, int(L))
    N = m * L
    return -d1 / N, K * K * d2 / N


def exact_curve(T, L: int):
    """Arrays (E/N, C/N) over temperatures T -- what the reference reads from Cvexact{L}.dat."""
    T = np.asarray(T, dtype=np.float64).reshape(-1)
    ec = np.array([energy_and_specific_heat(float(t), L) for t in T])
    return ec[:, 0], ec[:, 1]


T_C = 2.0 / np.log(1.0 + np.sqrt(2.0))


def finite_size_scaling(Ls=(8, 16, 32, 64, 128, 256), T: float = 2.269, n_replicas: int = 64, steps_per_site: int = 6000,
                        algo: str = "taylor", device=None, seed0: int = 1000, lanes_per_chain: int = 0):
    """BASELINE configs[3]: E/N and C/N (fluctuation estimator) at temperature T for every lattice
    size in Ls from `n_replicas` independent production chains, beside Kaufman's exact values.

    Returns a list of dicts {L, steps, E, E_err, E_exact, C, C_err, C_exact, z_E, z_C}; errors are
    delete-one-replica jackknife errors of the pooled sums (stats.pooled_jackknife)."""
    from . import engine, stats
    from ._lib import ALGO_BINARY, ALGO_TAYLOR
    a = ALGO_TAYLOR if algo == "taylor" else ALGO_BINARY
    out = []
    for L in Ls:
        N = L * L
        steps = int(steps_per_site) * N
        Tc = np.full(n_replicas, float(T))
        seeds = (np.arange(n_replicas) + seed0).astype(np.uint64)
        st = engine.PhiloxState(L, Tc, seeds, algo=a, device=device)
        st.run(steps, therm_steps=steps // 10, lanes_per_chain=lanes_per_chain)
        acc = st.acc.cpu().numpy().reshape(n_replicas, 1, 8)
        est = stats.pooled_jackknife(acc, np.array([float(T)]), N, 0 if algo == "taylor" else 1)
        e_exact, c_exact = energy_and_specific_heat(float(T), L)
        (E, Ee), (Cv, Ce) = est["E"], est["C"]
        out.append(dict(L=L, steps=steps, E=float(E[0]), E_err=float(Ee[0]), E_exact=e_exact, C=float(Cv[0]),
                        C_err=float(Ce[0]), C_exact=c_exact, z_E=float((E[0] - e_exact) / Ee[0]),
                        z_C=float((Cv[0] - c_exact) / Ce[0])))
    return out
