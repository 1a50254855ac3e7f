"""`PrincipalComponent` of the reference (pca.py:14-183) with the decomposition on the GPU.

Per temperature file of `{L}_config_{T}.txt` rows (`T p0 p1 ...`), the reference fits
`TruncatedSVD(n_components=1)` to the un-centred image matrix, keeps `explained_variance_`
(= the variance of the projection on the leading right singular vector) and estimates its error with a
delete-one-block jackknife over `num_blocks` KFold training sets.  Here the Gram matrices of the
blocks are built on the tensor cores (csrc/pca.cu, exact integers) and the leading eigenpairs of the
full and all delete-one sets come from one batched power iteration on the device (`worm_pca`).

sklearn's solver is a randomised approximation (nondeterministic unless numpy's global RNG is seeded);
this class returns the eigenpair it approximates.  On the reference's own output for the fixtures in
tests/golden/pca.npz the two agree to < 1e-8 relative.
"""
from __future__ import annotations

import os
import shutil

import numpy as np

from . import engine
from .stats import jackknife_err


def leading_eigenvalue(data, num_blocks=10, device=None, max_iters=2000, tol=1e-11):
    """(explained variance, jackknife error, component [1, d], per-set explained variances [1 + nb])
    of an [n, d] array / tensor of small integers (pca.py:97-147 for one file)."""
    import torch
    if isinstance(data, torch.Tensor):
        img = data.reshape(data.shape[0], -1).to(torch.int32)
    else:
        a = np.asarray(data)
        ai = np.ascontiguousarray(a.reshape(a.shape[0], -1).astype(np.int32))
        if not np.array_equal(ai, a.reshape(a.shape[0], -1)):
            raise ValueError("configuration data must be integers")
        img = torch.from_numpy(ai)
    img = img.to(engine._dev(device))
    res = engine.pca(img, num_blocks=num_blocks, max_iters=max_iters, tol=tol)
    ev = res.explained.cpu().numpy()
    err = jackknife_err(y_i=ev[1:], y_full=ev[0], num_blocks=num_blocks)
    return float(ev[0]), float(err), res.component.cpu().numpy()[None, :], ev


class PrincipalComponent(object):
    """Same constructor, attributes and `leading_eigenvalue_{L}.txt` format as pca.py:14-183.

    Attributes kept: `_temps`, `_eig_vals[key]` (a 1-tuple holding array([lambda_1]), the shape the
    reference's trailing comma at pca.py:113 produces), `_eig_vecs[key]` ([1, d]),
    `_leading_eig_val[key]`, `_err[key]`; keys are the file-name temperature with trailing zeros
    stripped (pca.py:137)."""

    def __init__(self, L, block_val=None, num_blocks=10, data_dir=None, save_dir=None, save=True, load=False,
                 verbose=False, device=None):
        self._L, self._block_val, self._num_blocks = L, block_val, num_blocks
        self._verbose, self._device = verbose, device
        self._temps = []
        self._err, self._eig_vals, self._eig_vecs, self._leading_eig_val = {}, {}, {}, {}
        # default directories of pca.py:50-71: raw configurations, or the blocked ones of a given scheme
        sub = '' if block_val is None else 'double_bonds_{}/'.format(block_val)
        if data_dir is not None:
            self._data_dir = data_dir
        elif block_val is None:
            self._data_dir = '../data/configs/{}_lattice/separated_data/'.format(L)
        else:
            self._data_dir = '../data/blocked_configs/{}_lattice/'.format(L) + sub
        self._save_dir = save_dir if save_dir is not None else '../data/pca/{}_lattice/'.format(L) + sub
        names = [f for f in os.listdir(self._data_dir)]
        self._config_files = sorted(os.path.join(self._data_dir, f) for f in names if f.endswith('.txt'))
        self._temp_strings = [f.split('_')[-1].rstrip('.txt').rstrip('0') for f in names]
        if not load:
            self._PCA()
            if save:
                self._save()
        else:
            self._load()

    def _get_data(self, _file=None):
        """pca.py:82-91: whitespace separated rows, first column is the temperature."""
        try:
            data = np.loadtxt(_file, ndmin=2)
        except IOError:
            raise IOError("Unable to read from: {}".format(_file))
        return data[:, 1:]

    def calc_PCA(self, data, num_components=1):
        """pca.py:97-115 for num_components = 1: ((array([lambda_1]),), components [1, d])."""
        if num_components != 1:
            raise NotImplementedError("only the leading component is computed on the device")
        lam, _, vec, _ = leading_eigenvalue(data, 1, self._device)
        return (np.array([lam]),), vec

    def _calc_leading_eigenvalue(self, data, num_components=1):
        return self.calc_PCA(data, num_components)[0][0]

    def _PCA(self, num_components=1):
        """pca.py:123-147, file by file; files with fewer samples than features are skipped (:132-133)."""
        for _file in self._config_files:
            if self._verbose:
                print("Reading data from: {}".format(_file))
            data = self._get_data(_file)
            num_samples, num_features = data.shape
            if num_samples < num_features:
                continue
            key = _file.split('_')[-1].rstrip('.txt').rstrip('0')
            self._temps.append(float(key))
            lam, err, vec, _ = leading_eigenvalue(data, self._num_blocks, self._device)
            self._eig_vals[key] = (np.array([lam]),)
            self._eig_vecs[key] = vec
            self._leading_eig_val[key] = self._eig_vals[key][0]
            self._err[key] = err

    def _save(self):
        """pca.py:149-166: `key lambda err` rows, previous file kept as .bak."""
        os.makedirs(self._save_dir, exist_ok=True)
        save_file = os.path.join(self._save_dir, 'leading_eigenvalue_{}.txt'.format(self._L))
        if os.path.exists(save_file):
            shutil.copy(save_file, save_file + '.bak')
        with open(save_file, 'w') as f:
            for key, val in self._eig_vals.items():
                f.write("{} {} {}\n".format(key, val[0][0], self._err[key]))

    def _load(self, data_dir=None):
        """pca.py:168-183."""
        data_file = os.path.join(self._save_dir if data_dir is None else data_dir,
                                 'leading_eigenvalue_{}.txt'.format(self._L))
        try:
            raw_data = np.loadtxt(data_file, ndmin=2, dtype=str)
        except Exception:
            print("Unable to read from {}".format(data_file))
            raise
        for row in raw_data:
            key = str(float(row[0]))           # the reference reads the key back as a float (pca.py:180)
            self._eig_vals[key] = [float(row[1]), float(row[2])]
            self._err[key] = float(row[2])
