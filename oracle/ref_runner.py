"""TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference executable in a sandbox.

`oracle/_ref/worm_ising_2d` is the reference's own `worm_algorithm/src/worm_ising_2d.cpp`
compiled by `oracle/Makefile` (target `ref`).  The executable hard-codes its file names
relative to the current directory (`../data/setup/input.txt`, worm_ising_2d.cpp:39; the
five output files, worm_ising_2d.cpp:68-78,189) and does not create directories, so every
run gets its own private tree

    <sandbox>/worm_algorithm/         <- cwd of the process
    <sandbox>/data/{setup,bonds,observables,num_bonds,bond_map}/lattice_L/

which is what `WormSimulation.__init__` (worm_simulation.py:27-38) creates for the real thing.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm may import this.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import tempfile
import time
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "_ref", "worm_ising_2d")


def have_ref() -> bool:
    return os.path.isfile(REF_BIN) and os.access(REF_BIN, os.X_OK)


@dataclass
class RefResult:
    L: int
    T: float
    output_line: str                 # raw text of output.txt (worm_ising_2d.cpp:189-191)
    K: float
    Z_av: float
    E_av: float
    Nb_av: float                     # the garbage column
    step_num: int
    observables_text: str            # raw text appended to observables_L.txt
    observables: np.ndarray          # parsed rows [n_events, 5]
    bonds_text: str                  # raw text of bonds_{T}.txt ('' when bond_flag == 0)
    dumps: np.ndarray                # [n_dumps, 2N] int32 occupations (incl. the final dump)
    bond_map_text: str
    bond_map: np.ndarray             # [4N, 5]
    num_bonds_text: str
    Nb_final: int
    wall_s: float = 0.0
    files: dict = field(default_factory=dict)


def _mk_tree(root: str, L: int) -> str:
    cwd = os.path.join(root, "worm_algorithm")
    os.makedirs(cwd, exist_ok=True)
    os.makedirs(os.path.join(root, "data", "setup"), exist_ok=True)
    for d in ("bonds", "observables", "num_bonds", "bond_map"):
        os.makedirs(os.path.join(root, "data", d, "lattice_%d" % L), exist_ok=True)
    return cwd


def write_input(root: str, L: int, T: float, num_steps: int, seed: float, bond_flag: int) -> None:
    # same format string as worm_simulation.py:53-55
    with open(os.path.join(root, "data", "setup", "input.txt"), "w") as f:
        f.write("%i %.12f %i %.32f %i\n" % (L, T, num_steps, seed, bond_flag))


def run_ref(L: int, T: float, num_steps: int, seed: float, bond_flag: int = 1,
            keep_dir: str | None = None) -> RefResult:
    """One run of the reference executable; returns everything it wrote."""
    if not have_ref():
        raise RuntimeError("oracle/_ref/worm_ising_2d missing: run `make -C oracle ref`")
    root = keep_dir or tempfile.mkdtemp(prefix="wormref_")
    try:
        cwd = _mk_tree(root, L)
        write_input(root, L, T, int(num_steps), seed, bond_flag)
        t0 = time.perf_counter()
        rc = subprocess.run([REF_BIN], cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        wall = time.perf_counter() - t0
        if rc.returncode != 0:
            raise RuntimeError("reference exited %d: %s" % (rc.returncode, rc.stderr[-500:]))
        data = os.path.join(root, "data")
        with open(os.path.join(data, "setup", "output.txt")) as f:
            out_line = f.read()
        vals = out_line.split()
        obs_path = os.path.join(data, "observables", "lattice_%d" % L, "observables_%d.txt" % L)
        with open(obs_path) as f:
            obs_text = f.read()
        obs_rows = [[float(x) for x in ln.split()] for ln in obs_text.splitlines() if ln.strip()]
        obs = np.array(obs_rows, dtype=np.float64).reshape(-1, 5)
        bdir = os.path.join(data, "bonds", "lattice_%d" % L)
        bfiles = sorted(os.listdir(bdir))
        bonds_text = ""
        files = {"bonds_files": bfiles}
        for bf in bfiles:
            with open(os.path.join(bdir, bf)) as f:
                bonds_text += f.read()
        nb2 = 2 * L * L
        if bonds_text.strip():
            arr = np.array(bonds_text.split(), dtype=np.int64).reshape(-1, 2)
            assert arr.shape[0] % nb2 == 0
            assert (arr[:, 0].reshape(-1, nb2) == np.arange(nb2)).all()
            dumps = arr[:, 1].reshape(-1, nb2).astype(np.int32)
        else:
            dumps = np.zeros((0, nb2), dtype=np.int32)
        with open(os.path.join(data, "bond_map", "lattice_%d" % L, "bond_map_%d.txt" % L)) as f:
            bm_text = f.read()
        bm = np.array(bm_text.split(), dtype=np.int64).reshape(-1, 5)
        with open(os.path.join(data, "num_bonds", "lattice_%d" % L, "num_bonds_%d.txt" % L)) as f:
            nbt = f.read()
        return RefResult(
            L=int(vals[0]), T=float(vals[1]), output_line=out_line, K=float(vals[2]),
            Z_av=float(vals[3]), E_av=float(vals[4]), Nb_av=float(vals[5]), step_num=int(vals[6]),
            observables_text=obs_text, observables=obs, bonds_text=bonds_text, dumps=dumps,
            bond_map_text=bm_text, bond_map=bm, num_bonds_text=nbt,
            Nb_final=int(nbt.split()[-1]), wall_s=wall, files=files)
    finally:
        if keep_dir is None:
            shutil.rmtree(root, ignore_errors=True)


def time_ref_all_cores(L: int, temps, num_steps: int, n_procs: int, seed0: int = 1,
                       bond_flag: int = 0, budget_s: float | None = None):
    """CPU baseline: `n_procs` concurrent copies of the reference (it is single-threaded,
    one chain per process: static RNG state mt19937ar.c:64-65), each in its own sandbox,
    over the temperature list round-robin.  Returns (total_steps, wall_s, n_runs).

    All-core rate = sum(step_num) / wall."""
    if not have_ref():
        raise RuntimeError("oracle/_ref/worm_ising_2d missing")
    temps = list(temps)
    roots = [tempfile.mkdtemp(prefix="wormref_p%d_" % p) for p in range(n_procs)]
    try:
        cwds = [_mk_tree(r, L) for r in roots]
        total_steps = 0
        n_runs = 0
        t0 = time.perf_counter()
        pending = list(enumerate(temps))
        running = {}           # slot -> Popen
        while pending or running:
            for slot in range(n_procs):
                if slot not in running and pending:
                    if budget_s is not None and time.perf_counter() - t0 > budget_s:
                        pending = []
                        break
                    idx, T = pending.pop(0)
                    write_input(roots[slot], L, float(T), int(num_steps), float(seed0 + idx), bond_flag)
                    running[slot] = subprocess.Popen([REF_BIN], cwd=cwds[slot],
                                                     stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            done = [s for s, p in running.items() if p.poll() is not None]
            for s in done:
                p = running.pop(s)
                if p.returncode != 0:
                    raise RuntimeError("reference exited %d" % p.returncode)
                with open(os.path.join(roots[s], "data", "setup", "output.txt")) as f:
                    total_steps += int(f.read().split()[6])
                n_runs += 1
            if not done:
                time.sleep(0.002)
        wall = time.perf_counter() - t0
        return total_steps, wall, n_runs
    finally:
        for r in roots:
            shutil.rmtree(r, ignore_errors=True)
