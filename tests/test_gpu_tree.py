"""Production mode writes the reference's file tree from the device buffers of a B200 run.
GPU half of the drop-in proof on the file boundary: the tree WormSimulation(rng="philox", data_dir=...) writes is
byte-for-byte the tree built on the CPU from the oracle's integers (tests/tree_helpers.py); the CPU half
(tests/test_reference_reads_our_files.py) feeds that same tree -- and the copy of the GPU-written one committed
under tests/golden/ -- to the UNMODIFIED reference classes."""
import os
import shutil

import numpy as np
import pytest

import tree_helpers as th

pytestmark = pytest.mark.gpu


def _run(tmp, **kw):
    from worm_algorithm_b200.worm_simulation import WormSimulation
    return WormSimulation(L=th.L, num_steps=th.STEPS, T_arr=th.TEMPS, n_replicas=th.REPLICAS, rng="philox", bond_flag=1,
                          therm_frac=th.THERM_FRAC, write_steps=th.WRITE_STEPS, max_dumps=th.MAX_DUMPS,
                          data_dir=str(tmp), verbose=False, **kw)


@pytest.mark.parametrize("lanes", [0, -2, 1])
def test_production_tree_equals_oracle_built_tree(tmp_path, lanes):
    sim = _run(tmp_path / "gpu", lanes_per_chain=lanes)
    th.oracle_tree(str(tmp_path / "cpu"))
    got, want = th.tree_digest(str(tmp_path / "gpu")), th.tree_digest(str(tmp_path / "cpu"))
    assert set(got) == set(want), set(got) ^ set(want)
    assert got == want, [k for k in want if got[k] != want[k]]
    assert {"observables/lattice_8/observables_8.txt", "num_bonds/lattice_8/num_bonds_8.txt", "setup/output.txt",
            "bond_map/lattice_8/bond_map_8.txt", "bonds/lattice_8/bonds_2.250.txt"} <= set(got)
    assert sorted(sim._E) == sorted(str(float(T)).rstrip("0") for T in th.TEMPS)
    # a copy for the CPU-side test with the unmodified reference (merged back by gpurun, committed as a fixture)
    out = os.environ.get("WORM_TREE_OUT")
    if out and lanes == 0:
        shutil.rmtree(out, ignore_errors=True)
        shutil.copytree(str(tmp_path / "gpu"), out)


def test_our_mirrors_read_the_production_tree(tmp_path):
    """SpecificHeat / Bonds(run=False) / CountBonds mirrors on the GPU-written tree against the oracle side."""
    from oracle import postproc_oracle as po
    from worm_algorithm_b200.bonds import Bonds
    from worm_algorithm_b200.count_bonds import count_stats_with_err
    from worm_algorithm_b200.specific_heat import SpecificHeat
    root = tmp_path / "gpu"
    _run(root)
    dumps_by_T = th.oracle_tree(str(tmp_path / "cpu"))
    sh = SpecificHeat(th.L, num_blocks=4, data_file=os.path.join(str(root), "observables", "lattice_8", "observables_8.txt"))
    ref = SpecificHeat(th.L, num_blocks=4, data_file=os.path.join(str(tmp_path / "cpu"), "observables", "lattice_8", "observables_8.txt"))
    assert sh._energy_temps == ref._energy_temps and list(sh._avg_energy) == list(ref._avg_energy)
    assert sh._spec_heat == ref._spec_heat and len(sh._spec_heat) == len(th.TEMPS) - 4
    b = Bonds(th.L, run=False, T_arr=th.TEMPS, data_dir=str(root), write=False, verbose=False)
    for T in th.TEMPS:
        key = str(float(T)).rstrip("0")
        imgs = np.array([b._config_data[key][i] for i in range(len(b._config_data[key]))])
        want = np.array([po.image_from_bonds(d, th.L) for d in dumps_by_T[T]])
        assert (imgs == want).all(), T
        n = (dumps_by_T[T] & 1).sum(axis=1)
        assert (imgs[:, ::2, 1::2].sum(axis=(1, 2)) + imgs[:, 1::2, ::2].sum(axis=(1, 2)) == n).all()
        assert count_stats_with_err(n, 3)[0][0] == float(np.mean(n))


def test_bonds_run_true_philox_uses_the_dumps_of_its_own_run(tmp_path):
    """Bonds(run=True, rng='philox') must rasterise the dumps of the run it just made."""
    from worm_algorithm_b200 import engine
    from worm_algorithm_b200.bonds import Bonds
    b = Bonds(th.L, run=True, num_steps=th.STEPS, T_arr=th.TEMPS, rng="philox", n_replicas=th.REPLICAS, bond_flag=1,
              max_dumps=th.MAX_DUMPS, write=False, verbose=False)
    res = b.results
    nd = res["n_dumps"].cpu().numpy()
    assert (nd > 0).all()
    T = np.asarray(res["T"])
    for T0 in th.TEMPS:
        key = str(float(T0)).rstrip("0")
        parts = []
        for c in np.flatnonzero(T == T0):
            k = min(int(nd[c]), th.MAX_DUMPS)
            parts += [res["dumps"][c, :k], res["bonds"][c:c + 1]]
        import torch
        want = engine.image(torch.cat(parts, dim=0), th.L).cpu().numpy()
        got = np.array([b._config_data[key][i] for i in range(len(b._config_data[key]))])
        assert got.shape == want.shape and (got == want).all()


def test_events_are_recorded_without_dumps(tmp_path):
    """bond_flag = 0 still writes the observables rows (worm_ising_2d.cpp:113-119 do not depend on bond_flag)."""
    from worm_algorithm_b200.worm_simulation import WormSimulation
    sim = WormSimulation(L=th.L, num_steps=th.STEPS, T_arr=th.TEMPS, rng="philox", bond_flag=0, data_dir=str(tmp_path),
                         verbose=False)
    rows = np.loadtxt(os.path.join(str(tmp_path), "observables", "lattice_8", "observables_8.txt"))
    assert rows.shape[1] == 5 and rows.shape[0] >= len(th.TEMPS)
    assert not os.listdir(os.path.join(str(tmp_path), "bonds", "lattice_8"))
    assert sim.results["dumps"] is None and int(sim.results["n_dumps"].min()) > 0
    assert (rows[:, 1] < 0).all() and set(np.round(rows[:, 0], 6)) == set(th.TEMPS)
