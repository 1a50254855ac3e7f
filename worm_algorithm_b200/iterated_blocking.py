"""`iterated_blocking` of the reference (iterated_blocking.py:6-122) on the GPU: the same kernel
as block_configs, recursion for repeated blocking (64 -> 32 -> ... -> 2), its own output tree."""
from __future__ import annotations

import numpy as np

from . import block_configs as _bc
from .block_configs import save_blocked_configs  # noqa: F401  (same writer, iterated_blocking.py:93-122)


def _block_config(config, num_block_steps=1, double_bonds_value=None):
    """iterated_blocking.py:6-59.  Only double_bonds_value None/0 is meaningful there (any other
    value reads unassigned variables) and is rejected here."""
    if double_bonds_value not in (None, 0):
        raise ValueError("iterated_blocking._block_config only implements the '1+1=0' scheme")
    return _bc.block_many(config, num_block_steps, 0)[0]


def block_configs(file, out_dir=None):
    """iterated_blocking.py:61-90."""
    print("Reading from {}".format(file))
    temp = _bc.temperature_tag(file)
    configs = _bc.read_configs(file)
    L = int(np.sqrt(configs.shape[1]) / 2)
    blocked = _bc.block_many(configs, 1, 0)
    L_blocked = int(np.sqrt(blocked.shape[1]) / 2)
    if out_dir is None:
        out_dir = '../data/iterated_blocking/{}_lattice/blocked_{}/'.format(L, L_blocked)
    save_blocked_configs(blocked, temp, out_dir=out_dir)
    return blocked


def iterate(images, levels=None, device=None):
    """All levels of the chain W -> W/2 -> ... -> 4 (lattice 2) as device tensors, without leaving
    the GPU (iterated_blocking.ipynb cells 11-22 chain five file round trips for this)."""
    import torch
    from . import engine
    cur = images if isinstance(images, torch.Tensor) else torch.from_numpy(_bc._as_images(images)).to(engine._dev(device))
    out = []
    while cur.shape[-1] > 4 and (levels is None or len(out) < levels):
        cur = engine.block(cur, 0)
        out.append(cur)
    return out
