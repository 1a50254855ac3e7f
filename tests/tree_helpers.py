"""Shared by the CPU and GPU tests of the production-mode file layer: the parameters of one small sweep and
the tree the compat writer must produce for it, built on the CPU from the ORACLE's integers."""
import hashlib
import os

import numpy as np

L = 8
TEMPS = [1.5, 1.75, 2.0, 2.25, 2.5, 2.75, 3.0]
REPLICAS = 2
STEPS = 30000
THERM_FRAC = 0.1
WRITE_STEPS = 50
MAX_DUMPS = 40


def chain_table():
    """(T, seed, global chain id) per chain, in WormSimulation._chain_table order (replica-major)."""
    nT = len(TEMPS)
    return [(TEMPS[c % nT], c // nT, c) for c in range(nT * REPLICAS)]


def oracle_tree(root):
    """Run every chain through oracle/worm_oracle.c and write the tree with the product's compat writer.
    Returns {temperature: uint8 [n_cfg, 2N]} in file order (stored dumps, then the final configuration, per chain)."""
    from oracle import worm_oracle as wo
    from worm_algorithm_b200 import compat_io
    tree = compat_io.DataTree(root, L)
    tree.remove_old_bonds()
    therm = int(THERM_FRAC * STEPS)
    by_T = {}
    for T, seed, cid in chain_table():
        o = wo.run_philox(L, 0, cid, seed, T, 0, STEPS, therm, dump_every=WRITE_STEPS, max_dumps=MAX_DUMPS)
        k = min(o["n_dumps"], MAX_DUMPS)
        tree.write_chain_production(float("%.12f" % T), int(o["acc"][0]), int(o["acc"][1]), STEPS, therm,
                                    int(o["bonds"].sum()), o["events"][:k], o["dumps"][:k], o["bonds"])
        by_T.setdefault(T, []).extend([o["dumps"][:k], o["bonds"][None, :]])
    tree.write_headers()
    return {T: np.concatenate(v, axis=0) for T, v in by_T.items()}


def tree_digest(root):
    """{relative path: sha256} of every file under root."""
    out = {}
    for dirpath, _, files in os.walk(root):
        for f in files:
            p = os.path.join(dirpath, f)
            out[os.path.relpath(p, root)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    return out
