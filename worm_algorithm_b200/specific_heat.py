"""`SpecificHeat` of the reference (specific_heat.py:7-169): per-temperature mean energy with
jackknife error from `observables_{L}.txt` and C_V as the mean of six |dE|/dT finite-difference
stencils.  Float64 on the host -- a few thousand numbers; this module is the *definition* of the
reference's C_V.  The fluctuation estimators the kernels make possible are in stats.py."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from .stats import block_resampling, jackknife_err


def finite_difference_cv(T, E):
    """specific_heat.py:137-156: for i = 2 .. len-3 the mean of
    |E[i+1]-E[i]|/dT, |E[i]-E[i-1]|/dT, |E[i+1]-E[i-1]|/dT, |E[i+2]-E[i-2]|/dT, |E[i+2]-E[i]|/dT,
    |E[i+2]-E[i-1]|/dT.  Returns (temperatures, values)."""
    T = [float(t) for t in T]
    E = [float(e) for e in E]
    pairs = ((1, 0), (0, -1), (1, -1), (2, -2), (2, 0), (2, -1))
    ts, cs = [], []
    for i in range(2, len(E) - 2):
        acc = 0.0
        for hi, lo in pairs:
            acc = acc + np.abs(E[i + hi] - E[i + lo]) / (T[i + hi] - T[i + lo])
        ts.append(T[i])
        cs.append(acc / 6)
    return ts, cs


class SpecificHeat(object):
    def __init__(self, L, num_blocks=5, data_file=None):
        self._L = L
        self._num_blocks = num_blocks
        self._data_file = ('../data/observables/lattice_{}/observables_{}.txt'.format(L, L)
                           if data_file is None else data_file)
        self._energy_dict = self._load_energies()
        self._energy_temps, self._avg_energy, self._energy_error = self._calc_avg_energies()
        tup = self._calc_specific_heat()
        self._specific_heat_dict, self._spec_heat_temps, self._spec_heat, self._spec_heat_err = tup
        self._summary = {}

    def _load_energies(self):
        """Column 1 of the observables rows keyed by `str(T).rstrip('0')`; like the reference
        (specific_heat.py:62-68) the sign is forced negative for every row but the first of a key."""
        try:
            obs = np.loadtxt(self._data_file, ndmin=2)
        except IOError:
            print("Unable to find {}".format(self._data_file))
            raise
        energy_dict = {}
        for row in obs:
            key = str(row[0]).rstrip('0')
            if key in energy_dict:
                e = row[1]
                energy_dict[key].append(-e if e > 0 else e)
            else:
                energy_dict[key] = [row[1]]
        return OrderedDict(sorted(energy_dict.items(), key=lambda t: t[0]))

    def _calc_avg_energies(self):
        self._err = {}
        temps, avg, errors = [], [], []
        for key, val in self._energy_dict.items():
            temps.append(float(key))
            m = np.mean(val)
            avg.append(m)
            rs = [np.mean(b) for b in block_resampling(np.array(val), self._num_blocks)]
            e = jackknife_err(y_i=rs, y_full=m, num_blocks=self._num_blocks)
            errors.append(e)
            self._err[key] = e
        return temps, avg, errors

    def _calc_specific_heat(self, E_avg=None):
        E = self._avg_energy if E_avg is None else E_avg
        ts, cs = finite_difference_cv(self._energy_temps, E)
        d = OrderedDict((str(t).rstrip('0'), [c]) for t, c in zip(ts, cs))
        return d, [float(k) for k in d], [float(v[0]) for v in d.values()], [self._err[k] for k in d]
