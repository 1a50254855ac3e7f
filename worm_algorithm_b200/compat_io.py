"""Writers/readers for the reference's on-disk formats, so that the untouched
reference analysis classes (`SpecificHeat`, `CountBonds`, `Bonds`) can consume GPU output.

Everything here is host-side text formatting of integers the kernels produced; the float
expressions are the reference's own (worm_ising_2d.cpp:115-119,183-191), evaluated in double.
"""
from __future__ import annotations

import os

import numpy as np

_U64 = (1 << 64) - 1


def _wrap_i32(v: int) -> int:
    """The reference accumulates Nb_tot in a 32-bit `int` (worm_ising_2d.cpp:53,132)."""
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def averages(T: float, L: int, Z: int, Nb_tot: int, step: int):
    """(Z_av, E_av, Nb_av) exactly as worm_ising_2d.cpp:115-117 / :183-186 compute them,
    including the signed->unsigned garbage of Nb_av."""
    N = L * L
    nbt = _wrap_i32(int(Nb_tot))
    Z_av = Z / (step * 1.0)
    E_av = nbt * T / float(Z * N)
    Nb_av = float((nbt & _U64) // (Z * N))
    return Z_av, E_av, Nb_av


def production_averages(T: float, L: int, Z: int, SNb: int, step: int, therm: int):
    """(Z_av, E_av, Nb_av) of the production mode, where the sums start after the warm-up: Z_av = closed
    fraction of the measured iterations, E_av = -T <Nb> / N (the reference's estimator, :116,185, on sums that
    really exclude the warm-up), Nb_av = <Nb> / N (the reference's column is garbage)."""
    N = L * L
    measured = max(int(step) - int(therm), 1)
    Z_av = Z / float(measured)
    nb = SNb / float(Z * N) if Z else 0.0
    return Z_av, -T * nb, nb


def output_line(L: int, T: float, Z: int, Nb_tot: int, step: int) -> str:
    """setup/output.txt, worm_ising_2d.cpp:189-191."""
    Z_av, E_av, Nb_av = averages(T, L, Z, Nb_tot, step)
    return "%4i %.12f %.12f %.12f %.12f %.12f %u\n" % (L, T, 1.0 / T, Z_av, E_av, Nb_av, step)


def observables_rows(L: int, T: float, events) -> str:
    """observables_L.txt rows `T E_av Z_av Nb_av step` with ostream default formatting (%g, 6
    significant digits), one per event {step, Z, Nb_tot}; worm_ising_2d.cpp:115-119."""
    out = []
    for step, Z, nbt in np.asarray(events, dtype=np.int64).reshape(-1, 3).tolist():
        Z_av, E_av, Nb_av = averages(T, L, Z, nbt, step)
        out.append("%g %g %g %g %d\n" % (T, E_av, Z_av, Nb_av, step))
    return "".join(out)


_BONDS_TEMPLATE = {}


def _bonds_template(nb2: int):
    """One configuration of bonds_{T}.txt with every occupation 0, as bytes, and where the digits sit."""
    t = _BONDS_TEMPLATE.get(nb2)
    if t is None:
        text = "".join("%d 0\n" % b for b in range(nb2)).encode()
        buf = np.frombuffer(text, dtype=np.uint8).copy()
        pos = np.flatnonzero(buf == ord("\n")) - 1
        t = _BONDS_TEMPLATE[nb2] = (buf, pos)
    return t


def bonds_bytes(dumps) -> bytes:
    """bonds_{T}.txt: per configuration 2N lines `b occ` (worm_ising_2d.cpp:121-123,163-165).  Single-digit
    occupations (all the Taylor chain produces at T >= 1) are poked into a byte template of the file, all
    configurations at once; a configuration with an occupation >= 10 is formatted line by line."""
    d = np.asarray(dumps)
    d = d.reshape(-1, d.shape[-1])
    if d.shape[0] == 0:
        return b""
    buf, pos = _bonds_template(d.shape[1])
    if int(d.max()) <= 9:
        out = np.tile(buf, (d.shape[0], 1))
        out[:, pos] = d.astype(np.uint8) + ord("0")
        return out.tobytes()
    idx = range(d.shape[1])
    return "".join("".join("%d %d\n" % (b, o) for b, o in zip(idx, cfg.tolist())) for cfg in d).encode()


def bonds_text(dumps) -> str:
    return bonds_bytes(dumps).decode()


def production_observables_rows(L: int, T: float, events, therm: int) -> str:
    """observables_L.txt rows of the production mode, one per stored event {step, Z, sum Nb} (the values
    before that iteration's own update, like worm_ising_2d.cpp:115-119).  An event with Z = 0 has nothing to
    average yet and writes no row."""
    out = []
    for step, Z, snb in np.asarray(events, dtype=np.int64).reshape(-1, 3).tolist():
        if Z <= 0:
            continue
        Z_av, E_av, Nb_av = production_averages(T, L, Z, snb, step, therm)
        out.append("%g %g %g %g %d\n" % (T, E_av, Z_av, Nb_av, step))
    return "".join(out)


def temperature_file_tag(T: float) -> str:
    """First five characters of std::to_string(T) (worm_ising_2d.cpp:64-66)."""
    return ("%f" % T)[:5]


def bond_map_rows(L: int) -> np.ndarray:
    """bond_map_L.txt rows `bond x1 y1 x2 y2`, site-major, neighbour order +x +y -x -y
    (worm_ising_2d.cpp:169-181 with bond_number :214-244).  Needs L >= 3 (closed form)."""
    if L < 3:
        raise ValueError("bond map closed form needs L >= 3")
    N = L * L
    rows = np.zeros((4 * N, 5), dtype=np.int64)
    for i in range(N):
        x, y = i % L, i // L
        nbrs = ((y * L + (x + 1) % L, 2 * i), (((y + 1) % L) * L + x, 2 * i + 1),
                (y * L + (x - 1) % L, None), (((y - 1) % L) * L + x, None))
        for j, (n2, b) in enumerate(nbrs):
            if b is None:
                b = 2 * n2 + (0 if j == 2 else 1)
            rows[4 * i + j] = (b, x, y, n2 % L, n2 // L)
    return rows


def bond_map_text(L: int) -> str:
    return "".join("%d %d %d %d %d\n" % tuple(r) for r in bond_map_rows(L).tolist())


def num_bonds_line(L: int, T: float, Nb: int) -> str:
    """num_bonds_L.txt, worm_ising_2d.cpp:195-198 (`<<` formatting)."""
    return "%d %g %d\n" % (L, T, Nb)


class DataTree:
    """The `../data/` tree of the reference (worm_simulation.py:27-38), rooted anywhere."""

    def __init__(self, root: str, L: int):
        self.root = root
        self.L = L
        self.setup = os.path.join(root, "setup")
        self.obs_dir = os.path.join(root, "observables", "lattice_%d" % L)
        self.bonds_dir = os.path.join(root, "bonds", "lattice_%d" % L)
        self.num_bonds_dir = os.path.join(root, "num_bonds", "lattice_%d" % L)
        self.bond_map_dir = os.path.join(root, "bond_map", "lattice_%d" % L)
        for d in (self.setup, self.obs_dir, self.bonds_dir, self.num_bonds_dir, self.bond_map_dir):
            os.makedirs(d, exist_ok=True)

    def remove_old_bonds(self):
        """worm_simulation.py:59-67."""
        for f in os.listdir(self.bonds_dir):
            if f.endswith(".txt"):
                os.remove(os.path.join(self.bonds_dir, f))

    def write_chain(self, T: float, Z: int, Nb_tot: int, step: int, Nb_final: int, events,
                    dumps=None, final_bonds=None):
        """Everything one run of the reference executable leaves behind for one temperature."""
        L = self.L
        with open(os.path.join(self.obs_dir, "observables_%d.txt" % L), "a") as f:
            f.write(observables_rows(L, T, events))
        if dumps is not None or final_bonds is not None:
            with open(os.path.join(self.bonds_dir, "bonds_%s.txt" % temperature_file_tag(T)), "a") as f:
                if dumps is not None and len(dumps):
                    f.write(bonds_text(dumps))
                if final_bonds is not None:
                    f.write(bonds_text(np.asarray(final_bonds).reshape(1, -1)))
        self._write_tail(T, output_line(L, T, Z, Nb_tot, step), Nb_final)

    def write_chain_production(self, T: float, Z: int, SNb: int, step: int, therm: int, Nb_final: int, events,
                               dumps=None, final_bonds=None):
        """The same files for one chain of the production mode (fixed-length run, sums measured after the
        warm-up): observables rows from the stored events, the stored dumps plus the final configuration
        (worm_ising_2d.cpp:160-167), output.txt, num_bonds."""
        L = self.L
        with open(os.path.join(self.obs_dir, "observables_%d.txt" % L), "a") as f:
            f.write(production_observables_rows(L, T, events, therm))
        if dumps is not None or final_bonds is not None:
            with open(os.path.join(self.bonds_dir, "bonds_%s.txt" % temperature_file_tag(T)), "ab") as f:
                if dumps is not None and len(dumps):
                    f.write(bonds_bytes(dumps))
                if final_bonds is not None:
                    f.write(bonds_bytes(np.asarray(final_bonds).reshape(1, -1)))
        Z_av, E_av, Nb_av = production_averages(T, L, Z, SNb, step, therm)
        line = "%4i %.12f %.12f %.12f %.12f %.12f %u\n" % (L, T, 1.0 / T, Z_av, E_av, Nb_av, step)
        self._write_tail(T, line, Nb_final)

    def _write_tail(self, T: float, out_line: str, Nb_final: int):
        L = self.L
        bm = os.path.join(self.bond_map_dir, "bond_map_%d.txt" % L)
        if not os.path.exists(bm):
            with open(bm, "w") as f:
                f.write(bond_map_text(L))
        with open(os.path.join(self.setup, "output.txt"), "w") as f:
            f.write(out_line)
        with open(os.path.join(self.num_bonds_dir, "num_bonds_%d.txt" % L), "a") as f:
            f.write(num_bonds_line(L, T, Nb_final))

    def write_headers(self):
        """worm_simulation.py:135-155."""
        h = os.path.join(self.obs_dir, "observables_header.txt")
        if not os.path.isfile(h):
            with open(h, "w") as f:
                f.write("T E_avg Z_avg Nb_avg step_num")
        d = os.path.join(self.obs_dir, "observables_description.txt")
        if not os.path.isfile(d):
            with open(d, "w") as f:
                f.write("T: Temperature of simulation\n"
                        "E_avg: Averaged energy of the simulation.\n"
                        "Z_avg: Number of times head==tail / number ofsteps\n"
                        "Nb_avg: Average number of active bonds duringsimulation.\n"
                        "step_num: Number of steps performed. ")
