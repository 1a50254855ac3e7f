"""Drop-in proof on the file boundary (CPU; needs the reference checkout, skipped elsewhere): the
data tree written by OUR compat layer from the integers a run returns is consumed by the UNMODIFIED
reference analysis classes (`SpecificHeat`, `Bonds`, `CountBonds`, run in a subprocess with the
SURVEY App. F import shims), and what they compute equals what our mirrors / the oracle compute."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/worm_algorithm"

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")

_REF_SIDE = r'''
import json, os, sys
sys.path.insert(0, sys.argv[1])
from oracle.make_golden import import_reference
ref_bc, ref_ib, ref_utils, ref_cb, ref_sh, ref_bonds = import_reference()
import numpy as np
L, temps = int(sys.argv[2]), json.loads(sys.argv[3])
sh = ref_sh.SpecificHeat(L, num_blocks=4)
b = ref_bonds.Bonds(L, run=False, T_arr=temps, write=True)
cb = ref_cb.CountBonds(L, num_blocks=3)
out = dict(E_T=sh._energy_temps, E=[float(x) for x in sh._avg_energy], E_err=[float(x) for x in sh._energy_error],
           C_T=sh._spec_heat_temps, C=sh._spec_heat,
           images={k: [np.asarray(v[i]).astype(int).flatten().tolist() for i in range(len(v))] for k, v in b._config_data.items()},
           bond_stats={k: [float(x) for x in v] for k, v in cb.bond_stats.items()})
print("RESULT" + json.dumps(out))
'''


def test_unmodified_reference_classes_consume_our_tree(tmp_path):
    from oracle import postproc_oracle as po
    from oracle import worm_oracle as wo
    from worm_algorithm_b200 import compat_io
    from worm_algorithm_b200.specific_heat import SpecificHeat
    from worm_algorithm_b200.count_bonds import count_stats_with_err
    L, steps = 8, 30000
    temps = [1.5, 1.75, 2.0, 2.25, 2.5, 2.75, 3.0]
    root = tmp_path
    os.makedirs(root / "worm_algorithm")
    tree = compat_io.DataTree(str(root / "data"), L)
    dumps_by_T = {}
    for i, T in enumerate(temps):
        o = wo.run_mt(L, T, steps, 100.0 + i, 1, max_events=4096)        # the integers the GPU returns (test_gpu_parity)
        tree.write_chain(T, o["Z"], o["Nb_tot"], o["step_num"], o["Nb_final"], o["events"], o["dumps"], o["final_bonds"])
        dumps_by_T[T] = np.concatenate([o["dumps"], o["final_bonds"][None, :]], axis=0)
    tree.write_headers()
    script = root / "ref_side.py"
    script.write_text(_REF_SIDE)
    p = subprocess.run([sys.executable, str(script), ROOT, str(L), json.dumps(temps)], cwd=str(root / "worm_algorithm"),
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    res = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("RESULT")][0][6:])
    # SpecificHeat: the reference on our observables file == our mirror on the same file
    ours = SpecificHeat(L, num_blocks=4, data_file=os.path.join(tree.obs_dir, "observables_%d.txt" % L))
    assert res["E_T"] == ours._energy_temps and res["E"] == [float(x) for x in ours._avg_energy]
    assert res["E_err"] == [float(x) for x in ours._energy_error]
    assert res["C_T"] == ours._spec_heat_temps and res["C"] == ours._spec_heat
    # Bonds: the images the reference builds from our bonds_*.txt / bond_map == imaging of the dumps
    for T in temps:
        key = str(float(T)).rstrip('0')
        imgs = np.array(res["images"][key]).reshape(-1, 2 * L, 2 * L)
        want = np.array([po.image_from_bonds(d, L) for d in dumps_by_T[T]])
        assert (imgs == want).all(), T
    # CountBonds on the config files the reference Bonds wrote == our statistics of the same counts
    for T in temps:
        n = (dumps_by_T[T] & 1).sum(axis=1)
        full, err = count_stats_with_err(n, 3)
        got = res["bond_stats"][str(float(T))]
        assert got == [float(full[0]), float(err[0]), float(full[1]), float(err[1])], (T, got)


def _run_reference_on(root, L, temps):
    """The unmodified reference classes (subprocess, cwd = <root>/worm_algorithm so that `../data` is our tree)."""
    os.makedirs(os.path.join(root, "worm_algorithm"), exist_ok=True)
    script = os.path.join(root, "ref_side.py")
    with open(script, "w") as f:
        f.write(_REF_SIDE)
    p = subprocess.run([sys.executable, script, ROOT, str(L), json.dumps(temps)], cwd=os.path.join(root, "worm_algorithm"),
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("RESULT")][0][6:])


def _check_reference_result(res, obs_file, dumps_by_T, L, temps):
    from oracle import postproc_oracle as po
    from worm_algorithm_b200.count_bonds import count_stats_with_err
    from worm_algorithm_b200.specific_heat import SpecificHeat
    ours = SpecificHeat(L, num_blocks=4, data_file=obs_file)
    assert res["E_T"] == ours._energy_temps and res["E"] == [float(x) for x in ours._avg_energy]
    assert res["E_err"] == [float(x) for x in ours._energy_error]
    assert res["C_T"] == ours._spec_heat_temps and res["C"] == ours._spec_heat
    for T in temps:
        key = str(float(T)).rstrip('0')
        imgs = np.array(res["images"][key]).reshape(-1, 2 * L, 2 * L)
        want = np.array([po.image_from_bonds(d, L) for d in dumps_by_T[T]])
        assert (imgs == want).all(), T
        n = (dumps_by_T[T] & 1).sum(axis=1)
        full, err = count_stats_with_err(n, 3)
        assert res["bond_stats"][str(float(T))] == [float(full[0]), float(err[0]), float(full[1]), float(err[1])], T


def test_unmodified_reference_classes_consume_production_tree(tmp_path):
    """PRODUCTION mode (rng="philox"): the tree `WormSimulation._write_tree_philox` writes.  Built here from the
    oracle's integers with the product's writer; tests/test_gpu_tree.py proves on the B200 that the GPU run writes
    the same bytes."""
    import tree_helpers as th
    root = str(tmp_path)
    dumps_by_T = th.oracle_tree(os.path.join(root, "data"))
    res = _run_reference_on(root, th.L, th.TEMPS)
    _check_reference_result(res, os.path.join(root, "data", "observables", "lattice_8", "observables_8.txt"),
                            dumps_by_T, th.L, th.TEMPS)


def test_unmodified_reference_classes_consume_the_gpu_written_tree(tmp_path):
    """tests/golden/gpu_tree_L8.tar.gz was WRITTEN BY A B200 RUN (WormSimulation(rng="philox", data_dir=...) inside
    tests/test_gpu_tree.py, brought back through gpurun_out/ and committed unchanged).  It must be the oracle-built tree,
    file for file, and the unmodified reference classes must read it."""
    import tarfile
    import tree_helpers as th
    tar = os.path.join(ROOT, "tests", "golden", "gpu_tree_L8.tar.gz")
    if not os.path.exists(tar):
        pytest.skip("fixture not generated yet")
    root = str(tmp_path)
    with tarfile.open(tar) as t:
        t.extractall(os.path.join(root, "data"), filter="data")
    dumps_by_T = th.oracle_tree(os.path.join(root, "cpu"))
    assert th.tree_digest(os.path.join(root, "data")) == th.tree_digest(os.path.join(root, "cpu"))
    res = _run_reference_on(root, th.L, th.TEMPS)
    _check_reference_result(res, os.path.join(root, "data", "observables", "lattice_8", "observables_8.txt"),
                            dumps_by_T, th.L, th.TEMPS)
