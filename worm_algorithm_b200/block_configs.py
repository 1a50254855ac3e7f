"""`block_configs` of the reference (block_configs.py:6-137) on the GPU.

Same function names and argument meaning; the blocking itself is `worm_block` (csrc/postproc.cu),
every configuration of a file in one launch.  Files are the reference's formats.
"""
from __future__ import annotations

import os

import numpy as np

from . import engine


def _as_images(configs) -> np.ndarray:
    c = np.asarray(configs)
    flat = c.reshape(1, -1) if c.ndim == 1 else c.reshape(c.shape[0], -1)
    W = int(round(np.sqrt(flat.shape[1])))
    if W * W != flat.shape[1] or W % 2:
        raise ValueError("a configuration must be a flattened (2L x 2L) image")
    return np.ascontiguousarray(flat.reshape(flat.shape[0], W, W).astype(np.int32))


def block_many(configs, num_block_steps=1, double_bonds_value=0, variant=0, device=None):
    """[n, 4L^2] (or [n, 2L, 2L]) images -> [n, L'^2] blocked images after num_block_steps RG steps.
    Like the reference, only the xor scheme iterates (block_configs.py:66-69)."""
    import torch
    img = torch.from_numpy(_as_images(configs)).to(engine._dev(device))
    v = 0 if double_bonds_value is None else int(double_bonds_value)
    steps = int(num_block_steps) if v == 0 else 1
    for _ in range(max(steps, 1)):
        img = engine.block(img, v, variant=variant)
    return img.reshape(img.shape[0], -1).cpu().numpy().astype(np.int64)


def _block_config(config, num_block_steps=1, double_bonds_value=0):
    """One flattened 2L x 2L image -> flattened L x L image (block_configs.py:6-71)."""
    return block_many(config, num_block_steps, double_bonds_value)[0]


def read_configs(config_file):
    """`T p0 p1 ...` rows -> pixel matrix (first column dropped, like read_csv(index_col=0))."""
    rows = np.loadtxt(config_file, ndmin=2)
    return rows[:, 1:]


def temperature_tag(config_file) -> str:
    return config_file.split('/')[-1].split('_')[-1].rstrip('.txt')        # block_configs.py:86


def block_configs(config_file, double_bonds_value=0, out_dir=None, apply_double_bonds_value=False):
    """Block every configuration of `config_file` and save them (block_configs.py:73-105).
    The reference names the output directory after `double_bonds_value` but blocks with the default
    scheme 0 (:93); that is kept unless apply_double_bonds_value=True."""
    print("Reading from {}".format(config_file))
    temp = temperature_tag(config_file)
    configs = read_configs(config_file)
    L = int(np.sqrt(configs.shape[1]) / 2)
    blocked = block_many(configs, 1, double_bonds_value if apply_double_bonds_value else 0)
    if out_dir is None:
        out_dir = "../data/blocked_configs/{}_lattice/double_bonds_{}/".format(L, double_bonds_value)
    save_blocked_configs(blocked, temp, out_dir=out_dir)
    return blocked


def save_blocked_configs(configs, temperature, out_dir):
    """`DataFrame(configs).to_csv(sep=' ', header=None)`: row index, then the pixels
    (block_configs.py:108-137); an existing file is renamed `.bak`."""
    configs = np.asarray(configs)
    L = int(np.sqrt(configs.shape[1]) / 2)
    os.makedirs(out_dir, exist_ok=True)
    sep = "" if out_dir.endswith('/') else "/"
    out_file = out_dir + sep + "{}_blocked_configs_{}.txt".format(L, str(temperature))
    if os.path.exists(out_file):
        os.rename(out_file, out_file + ".bak")
    print("Saving to: {}\n".format(out_file))
    with open(out_file, "w") as f:
        for i, row in enumerate(configs):
            f.write("{} {}\n".format(i, " ".join(str(int(v)) for v in row)))
    return out_file
