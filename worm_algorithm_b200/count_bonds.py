"""`CountBonds` of the reference (count_bonds.py:11-235) with the counting on the GPU.

<n> and <n^2> - <n>^2 of the parity-bond count n = sum of image pixels with (i + j) odd, with the
reference's delete-one-block jackknife errors, per temperature file; save / load of
`bond_stats_{L}.txt`.  The per-configuration counts come from `worm_count` (csrc/postproc.cu); the
statistics are the reference's float64 expressions evaluated on those integers, so they agree with
it to the last bit.
"""
from __future__ import annotations

import os
import shutil
from collections import OrderedDict

import numpy as np

from . import engine
from .stats import block_resampling, jackknife_err


def bond_counts(data, width, device=None) -> np.ndarray:
    """int64 count per configuration of [n, width*width] (or [n, width, width]) pixel data."""
    import torch
    if isinstance(data, torch.Tensor):
        img = data.reshape(-1, width, width).to(torch.int32)
    else:
        img = torch.from_numpy(np.ascontiguousarray(np.asarray(data).reshape(-1, width, width).astype(np.int32)))
    img = img.to(engine._dev(device))
    return engine.count(img).cpu().numpy()


def count_stats(bc):
    """(<n>, <n^2> - <n>^2) with the reference's expressions (count_bonds.py:118-123)."""
    bc = np.asarray(bc)
    Nb_avg = np.mean(bc)
    Nb2_avg = np.mean(bc ** 2)
    return Nb_avg, Nb2_avg - Nb_avg ** 2


def count_stats_with_err(bc, num_blocks=10):
    """count_bonds.py:141-172 on precomputed per-configuration counts: blocks of configurations
    are deleted, so resampling the counts equals recounting the resampled configurations."""
    full = count_stats(bc)
    rs = np.array([count_stats(b) for b in block_resampling(np.asarray(bc), num_blocks)])
    err = [jackknife_err(y_i=rs[:, k], y_full=full[k], num_blocks=num_blocks) for k in range(len(full))]
    return full, err


class CountBonds(object):
    def __init__(self, L, block_val=None, num_blocks=10, data_dir=None, save_dir=None, data_file=None,
                 save=False, load=False, verbose=False, device=None):
        self._L = L
        self._block_val = block_val
        self._num_blocks = num_blocks
        self._verbose = verbose
        self._device = device
        self._width = 2 * L
        self._data_dir = '../data/configs/{}_lattice/separated_data/'.format(L) if data_dir is None else data_dir
        self._save_dir = '../data/bond_stats/{}_lattice/'.format(L) if save_dir is None else save_dir
        os.makedirs(self._save_dir, exist_ok=True)
        self._config_files = sorted(os.path.join(self._data_dir, i) for i in os.listdir(self._data_dir)
                                    if i.endswith('.txt'))
        temp_strings = [i.split('_')[-1].rstrip('.txt') for i in self._config_files]
        self._temp_strings = [i.rstrip('0') for i in temp_strings]
        self.bond_stats = {}
        if not load:
            self.count_bonds()
            self._save()
        else:
            self._load(data_file)

    def _get_configs(self, _file):
        try:
            return np.loadtxt(_file, ndmin=2)[:, 1:]          # first column (T) is the index column
        except IOError:
            print("Unable to read from: {}".format(_file))
            raise

    def _count_bonds(self, data):
        return count_stats(bond_counts(data, self._width, self._device))

    def _count_bonds_with_err(self, data, num_blocks=10):
        return count_stats_with_err(bond_counts(data, self._width, self._device), num_blocks)

    def count_bonds(self):
        for config_file in self._config_files:
            if self._verbose:
                print("Reading in from: {}\n".format(config_file))
            key = config_file.split('_')[-1].rstrip('.txt')
            data = self._get_configs(config_file)
            val, err = self._count_bonds_with_err(data, self._num_blocks)
            self.bond_stats[key] = [val[0], err[0], val[1], err[1]]

    def _save(self):
        save_file = os.path.join(self._save_dir, 'bond_stats_{}.txt'.format(self._L))
        save_file_copy = save_file + '.bak'
        orig_exists, copy_exists = os.path.exists(save_file), os.path.exists(save_file_copy)
        if orig_exists:
            shutil.copy(save_file, save_file_copy)
        if orig_exists and copy_exists:
            os.remove(save_file)
        ordered = OrderedDict(sorted(self.bond_stats.items(), key=lambda t: t[0]))
        with open(save_file, 'w') as f:
            for key, val in ordered.items():
                f.write('{} {} {} {} {}\n'.format(key, val[0], val[1], val[2], val[3]))

    def _load(self, data_file=None):
        if data_file is None:
            data_file = os.path.join(self._save_dir, 'bond_stats_{}.txt'.format(self._L))
        print("Reading from: {}".format(data_file))
        for row in np.loadtxt(data_file, ndmin=2):
            self.bond_stats[str(row[0])] = [row[1], row[2], row[3], row[4]]
