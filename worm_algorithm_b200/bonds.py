"""`Bonds` of the reference (bonds.py:14-503): bond dumps -> 2L x 2L images (-> blocked images),
written in the reference's `{L}_config_{T}.txt` format.

The reference walks dict-of-dict Python loops per bond (bonds.py:132-319); here every dump of a
temperature is rasterised in one `worm_image` launch and blocked in one `worm_block` launch
(csrc/postproc.cu, variant 1 = `Bonds._block_config_T`).  Dumps come straight from the simulation's
device tensors when `run=True`, or from `bonds_*.txt` files of a reference-format data tree.
The plotting half of the reference class (bonds.py:504-627) is out of scope.
"""
from __future__ import annotations

import os

import numpy as np

from . import engine
from .worm_simulation import WormSimulation, temperature_key


class Bonds(WormSimulation):
    def __init__(self, L, run=False, num_steps=1E7, decay_steps=False, verbose=True,
                 T_start=1., T_end=3.5, T_step=0.1, T_arr=None, block_configs=False, block_val=0,
                 write=True, write_blocked=False, data_dir=None, config_dir=None, **sim_kwargs):
        # (the reference passes decay_steps into WormSimulation's `verbose` slot, bonds.py:43-47 /
        #  SURVEY App. G; the arguments are passed by name here)
        sim_kwargs.setdefault("bond_flag", 1)
        WormSimulation.__init__(self, L=L, run=run, num_steps=num_steps, verbose=verbose, T_start=T_start,
                                T_end=T_end, T_step=T_step, T_arr=T_arr, data_dir=data_dir, **sim_kwargs)
        self._L = int(L)
        self._num_bonds = 2 * self._L * self._L
        self._data_dir = data_dir
        self._config_dir = config_dir
        self._bonds_dir = (os.path.join(data_dir, 'bonds', 'lattice_{}'.format(self._L)) + '/'
                           if data_dir is not None else '../data/bonds/lattice_{}/'.format(self._L))
        self._dumps = self._get_bonds()                      # {temperature key: uint8 [n_cfg, 2N]}
        self._active_bonds = {k: (v & 1) for k, v in self._dumps.items()}    # bonds.py:199-239 (odd parity)
        self._config_data = self.set_config_data()
        if block_configs:
            self._blocked_config_data = self.block_configs(block_val)
        if write:
            self.write_config_data(config_dir)
        if write_blocked:
            self.write_blocked_config_data(config_dir)

    # -- dumps -------------------------------------------------------------------------------------
    def _get_bonds(self):
        """Dumped occupations per temperature: from the run that just happened (device tensors), else
        from the `bonds_{T}.txt` files (bonds.py:132-174: chunks of 2L^2 rows `b occ`)."""
        res = getattr(self, "results", None)
        if res is not None and res.get("dumps") is not None:
            return self._dumps_from_results(res)
        if res is not None:
            raise ValueError("the run kept no dumps (bond_flag=0 or max_dumps=0): nothing to rasterise")
        out = {}
        if not os.path.isdir(self._bonds_dir):
            return out
        for f in sorted(os.listdir(self._bonds_dir)):
            if not f.endswith('.txt'):
                continue
            key = f.split('_')[-1].rstrip('.txt').rstrip('0')
            rows = np.loadtxt(os.path.join(self._bonds_dir, f), ndmin=2)
            if rows.shape[0] % self._num_bonds:
                raise ValueError("{}: {} rows is not a multiple of 2L^2".format(f, rows.shape[0]))
            occ = rows[:, 1].reshape(-1, self._num_bonds)
            if occ.max(initial=0) > 255:
                raise ValueError("bond occupation above 255")
            out[key] = occ.astype(np.uint8)
        return out

    def _dumps_from_results(self, res):
        out = {}
        T = np.asarray(res["T"])
        if res["mode"] == "mt19937":
            n_ev = res["n_events"].cpu().numpy()
            dumps = res["dumps"]
            final = res["bonds"]
            for c in range(T.size):
                k = min(int(n_ev[c]), dumps.shape[1])
                d = [dumps[c, :k], final[c:c + 1]]            # the executable dumps once more at the end (:160-167)
                out.setdefault(temperature_key(T[c]), []).extend(d)
        else:
            # the stored dumps of every chain, then its final configuration -- what the executable leaves in
            # bonds_{T}.txt (worm_ising_2d.cpp:120-124, :160-167)
            dumps, final = res["dumps"], res["bonds"]
            nd = res["n_dumps"].cpu().numpy()
            for c in range(T.size):
                k = min(int(nd[c]), dumps.shape[1])
                out.setdefault(temperature_key(float("%.12f" % T[c])), []).extend([dumps[c, :k], final[c:c + 1]])
        import torch
        return {k: torch.cat(v, dim=0) for k, v in out.items()}

    # -- images ------------------------------------------------------------------------------------
    def set_config_data(self):
        """{T: {config index: (2L, 2L) int image}} (bonds.py:321-331), one launch per temperature."""
        import torch
        self._images = {}
        out = {}
        for key, d in self._dumps.items():
            dt = d if isinstance(d, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(d))
            img = engine.image(dt.to(engine._dev(self._device)), self._L)
            self._images[key] = img
            host = img.cpu().numpy()
            out[key] = {i: host[i] for i in range(host.shape[0])}
        return out

    def _set_config_data(self, T, config_idx):
        return self._config_data[T][config_idx]

    def _block_config_T(self, T, config_idx, block_val):
        """bonds.py:333-410 for one configuration."""
        img = self._images[T][config_idx:config_idx + 1]
        return engine.block(img, int(block_val), variant=1)[0].cpu().numpy()

    def block_configs(self, block_val=0):
        """{block_val: {T: {config index: (L, L) image}}} (bonds.py:412-426)."""
        out = {block_val: {}}
        for key, img in self._images.items():
            host = engine.block(img, int(block_val), variant=1).cpu().numpy()
            out[block_val][key] = {i: host[i] for i in range(host.shape[0])}
        return out

    # -- files -------------------------------------------------------------------------------------
    def _write(self, configs, config_dir):
        os.makedirs(config_dir, exist_ok=True)
        for temp, cfgs in configs.items():
            tf = float(temp)
            fn = os.path.join(config_dir, '{}_config_{}.txt'.format(self._L, tf))
            print("Saving configs to: {}".format(fn))
            with open(fn, "a") as f:
                for config in cfgs.values():
                    f.write('{} {}\n'.format(tf, ' '.join(str(int(i)) for i in np.asarray(config).flatten())))

    def write_config_data(self, data_dir=None):
        """bonds.py:428-459: append `T p0 p1 ...` rows to {L}_config_{float(T)}.txt."""
        config_dir = data_dir if data_dir is not None else '../data/configs/{}_lattice/separated_data/'.format(self._L)
        self._write(self._config_data, config_dir)

    def write_blocked_config_data(self, data_dir=None, blocked_configs=None):
        """bonds.py:461-502."""
        configs = self._blocked_config_data if blocked_configs is None else blocked_configs
        block_val = list(configs.keys())[0]
        config_dir = (data_dir if data_dir is not None else
                      '../data/blocked_configs/{}_lattice/double_bonds_{}/'.format(self._L, block_val))
        self._write(configs[block_val], config_dir)
