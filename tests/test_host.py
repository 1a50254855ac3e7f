"""CPU tests (no GPU): the C-ABI library loads and exports every declared symbol, refuses to compute
without a device, host-side mirrors of the reference's statistics / file formats agree with the
oracle and the golden vectors, and the multi-rank plumbing (gloo, world_size 2) reduces exactly."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------------
# boundary
# ------------------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    from worm_algorithm_b200 import _lib
    lib = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "worm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(worm_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.worm_abi_version() == 1


def test_no_cpu_fallback():
    """Without a GPU every compute entry point fails loudly (status code + message)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from worm_algorithm_b200 import _lib, engine
    lib = _lib.lib()
    assert lib.worm_device_count() == 0
    img = np.zeros((1, 8, 8), dtype=np.int32)
    cnt = np.zeros(1, dtype=np.int64)
    rc = lib.worm_count_host(8, 1, img.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p))
    assert rc == -3 and "no CPU fallback" in _lib.last_error()
    with pytest.raises(_lib.WormError):
        engine.PhiloxState(8, [2.0], [1])
    from worm_algorithm_b200.worm_simulation import WormSimulation
    with pytest.raises(_lib.WormError):
        WormSimulation(L=8, num_steps=100, T_arr=[2.0], rng="philox", verbose=False)


def test_thresholds_without_gpu_match_oracle():
    from oracle import worm_oracle as wo
    from worm_algorithm_b200 import _lib
    for T in (0.5, 1.0, 2.269185314213022, 3.5):
        assert _lib.thresholds(T) == wo.thresholds(T)


def test_header_stream_contract_matches_oracle():
    """include/worm_b200.h states the production stream (version 2) in words; worm_philox_words() returns it.
    (1) the words equal the oracle's Philox block read the way the header says; (2) a worm written here from
    nothing but the header's paragraph (words, bit fields, integer acceptance tests) lands on the same state
    and sums as oracle/worm_oracle.c:worm_oracle_philox_run -- so a binder reading the header gets the
    contract the kernels implement."""
    from oracle import worm_oracle as wo
    from worm_algorithm_b200 import _lib
    for seed, chain, step in ((1, 0, 0), (1, 0, 1), (1, 0, 2), (1, 0, 3), (2 ** 63 + 5, 2 ** 40 + 3, 2 ** 35 + 7),
                              (0xDEADBEEF12345678, 77, 2 ** 33 - 1)):
        pair = step >> 1
        blk = wo.philox([pair & 0xFFFFFFFF, pair >> 32, chain & 0xFFFFFFFF, chain >> 32], [seed & 0xFFFFFFFF, seed >> 32])
        a, b = _lib.philox_words(seed, chain, step)
        assert (a, b) == ((int(blk[2]), int(blk[3])) if step & 1 else (int(blk[0]), int(blk[1])))
    # the known answer printed in the header: chain 0, seed 1, steps 0..3
    assert [_lib.philox_words(1, 0, s) for s in range(4)] == [(0xe3e80670, 0xe50a0ebc), (0x95f222c0, 0xb615aa27),
                                                              (0xac08141b, 0xdfc5ccbe), (0x79c07a47, 0xa7f66093)]
    for algo in (0, 1):
        L, T, seed, chain, steps = 4, 1.7, 12345, 9, 600
        N = L * L
        kfix, tfix, hfix = _lib.thresholds(T)
        bonds = [0] * (2 * N)
        head = tail = 0
        Z = SNb = 0
        for s in range(steps):
            a, b = _lib.philox_words(seed, chain, s)
            if head == tail:
                Z += 1
                SNb += sum(bonds)
                head = tail = (((b << 3) & 0xFFFFFFFF) * N) >> 32
            d = b >> 30
            x, y = head % L, head // L
            if d == 0:
                nh, ib = y * L + (x + 1) % L, 2 * head
            elif d == 1:
                nh, ib = ((y + 1) % L) * L + x, 2 * head + 1
            elif d == 2:
                nh = y * L + (x - 1) % L
                ib = 2 * nh
            else:
                nh = ((y - 1) % L) * L + x
                ib = 2 * nh + 1
            nb = bonds[ib]
            if algo == 0:
                if (b >> 29) & 1 == 0:
                    acc, delta = (nb < 255 and (nb + 1) * a < kfix), 1
                else:
                    acc, delta = (a < nb * tfix), -1
            else:
                acc, delta = (True, -1) if nb else (a < hfix, 1)
            if acc:
                bonds[ib] += delta
                head = nh
        o = wo.run_philox(L, algo, chain, seed, T, 0, steps, 0)
        assert list(o["bonds"]) == bonds and o["head"] == head and o["tail"] == tail
        assert int(o["acc"][0]) == Z and int(o["acc"][1]) == SNb


def test_product_kaufman_matches_enumeration_and_the_oracle():
    """worm_algorithm_b200/kaufman.py (second-order jets in numpy) against a brute-force enumeration written here
    (L = 2, 3, 4, T_c included: the formula's g_0 changes sign there) and against oracle/kaufman.py (torch
    autograd) on the lattices of BASELINE configs[3]."""
    import itertools
    from oracle import kaufman as ok
    from worm_algorithm_b200 import kaufman as pk
    tc = pk.T_C
    for L in (2, 3, 4):
        idx = np.arange(L * L).reshape(L, L)
        right, down = np.roll(idx, -1, axis=1), np.roll(idx, -1, axis=0)
        es = np.array([-(np.array(b)[idx] * np.array(b)[right]).sum() - (np.array(b)[idx] * np.array(b)[down]).sum()
                       for b in itertools.product((-1, 1), repeat=L * L)], dtype=np.float64)
        for T in (1.2, 2.269, tc, tc * (1 + 1e-9), 3.5):
            w = np.exp(-(es - es.min()) / T)
            e1, e2 = (w * es).sum() / w.sum(), (w * es * es).sum() / w.sum()
            e, c = pk.energy_and_specific_heat(T, L)
            assert abs(e - e1 / (L * L)) < 1e-12 and abs(c - (e2 - e1 * e1) / (T * T * L * L)) < 1e-10, (L, T)
    for L in (8, 16, 32, 64, 128, 256):
        for T in (1.0, 1.5, 2.269, tc, 3.5):
            a, b = pk.energy_and_specific_heat(T, L), ok.energy_and_specific_heat(T, L)
            assert abs(a[0] - b[0]) < 1e-11 and abs(a[1] - b[1]) < 1e-8, (L, T, a, b)
    E, Cv = pk.exact_curve([2.0, 2.269], 32)
    assert abs(E[1] + 1.43400062226) < 1e-9 and abs(Cv[1] - 1.84594475339) < 1e-9


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "worm_algorithm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libworm_oracle" not in src, f


# ------------------------------------------------------------------------------------------------
# statistics and file formats (host-side mirrors of the reference)
# ------------------------------------------------------------------------------------------------
def test_stats_match_oracle_and_reference_golden(post_golden):
    from oracle import postproc_oracle as po
    from worm_algorithm_b200 import stats
    from worm_algorithm_b200.count_bonds import count_stats_with_err
    from worm_algorithm_b200.specific_heat import finite_difference_cv
    rng = np.random.default_rng(3)
    for n, nb in ((17, 5), (40, 10), (9, 3), (20, 20)):
        x = rng.normal(size=(n, 3))
        a, b = stats.block_resampling(x, nb), po.block_resampling(x, nb)
        assert len(a) == len(b) == nb and all((u == v).all() for u, v in zip(a, b))
    y = rng.normal(size=7)
    assert stats.jackknife_err(y, 0.3, 7) == po.jackknife_err(y, 0.3, 7)
    g = post_golden
    # the reference's own block_resampling / jackknife_err outputs (utils.py, run unmodified)
    x = g["jk_x"]
    for nb in (2, 5, 10):
        means = np.array([np.mean(b) for b in stats.block_resampling(x, nb)])
        assert np.array_equal(means, g["jk_n37_b%d_means" % nb])
        assert stats.jackknife_err(means, np.mean(x), nb) == float(g["jk_n37_b%d_err" % nb])
    # counts -> (<n>, var) and jackknife errors: identical floats to the reference's CountBonds
    for L in (4, 8, 16):
        tag = "img_L%d" % L
        bc = po.bond_counts(g[tag + "_images"], 2 * L)
        ref = g[tag + "_count_err"]
        full, err = count_stats_with_err(bc, int(ref[4]))
        assert [full[0], err[0], full[1], err[1]] == list(ref[:4])
    # the reference's SpecificHeat._calc_specific_heat output
    ts, cs = finite_difference_cv(g["sh_T"], g["sh_E"])
    assert np.array_equal(np.array(ts), g["sh_out_T"]) and np.array_equal(np.array(cs), g["sh_out_C"])
    T = np.linspace(1.0, 3.4, 13)
    E = -2.0 + 0.3 * T + 0.05 * np.sin(5 * T)
    ts, cs = finite_difference_cv(T, E)
    t2, c2 = po.specific_heat_fd(T, E)
    assert np.array_equal(np.array(ts), t2) and np.array_equal(np.array(cs), c2)


def test_specific_heat_class_on_reference_format_file(tmp_path, ref_runs):
    """observables_{L}.txt written by our compat layer == the reference executable's text, and
    SpecificHeat on it reproduces the oracle's finite-difference numbers."""
    from oracle import postproc_oracle as po
    from worm_algorithm_b200.specific_heat import SpecificHeat
    rng = np.random.default_rng(5)
    f = tmp_path / "observables_8.txt"
    temps = [1.0 + 0.25 * i for i in range(9)]
    rows = []
    for T in temps:
        for k in range(12):
            rows.append((T, -2.0 + 0.3 * T + 0.01 * rng.normal(), 0.1, 1.0, 1000 + 50 * k))
    with open(f, "w") as fh:
        for r in rows:
            fh.write("%g %g %g %g %d\n" % r)
    sh = SpecificHeat(8, num_blocks=4, data_file=str(f))
    tab = np.loadtxt(f)
    E, err = [], []
    for T in temps:
        m, e = po.avg_energy_with_err(tab[tab[:, 0] == T][:, 1], 4)
        E.append(m); err.append(e)
    assert np.allclose(sh._avg_energy, E, rtol=0, atol=0)
    assert np.allclose(sh._energy_error, err, rtol=0, atol=0)
    t2, c2 = po.specific_heat_fd(temps, E)
    assert sh._spec_heat_temps == list(t2) and sh._spec_heat == list(c2)


def test_compat_files_equal_reference_executable_text(ref_runs):
    """output.txt / observables rows / num_bonds line / bonds text / bond map formatted from the
    integers a run returns (here from the CPU restatement; the GPU returns the same integers, see
    test_gpu_parity) == the text the reference executable wrote for that run (golden)."""
    import hashlib
    from oracle import worm_oracle as wo
    from worm_algorithm_b200 import compat_io
    checked = 0
    for r in ref_runs:
        if r["num_steps"] > 400000:
            continue
        L, T = r["L"], r["T"]
        o = wo.run_mt(L, T, r["num_steps"], r["seed"], r["bond_flag"], max_events=4096)
        assert compat_io.output_line(L, T, o["Z"], o["Nb_tot"], o["step_num"]) == r["output_line"]
        assert compat_io.num_bonds_line(L, T, o["Nb_final"]) == r["num_bonds_text"]
        obs = compat_io.observables_rows(L, T, o["events"])
        assert hashlib.sha256(obs.encode()).hexdigest() == r["observables_sha"]
        if L >= 3:
            bm = compat_io.bond_map_rows(L).astype(np.int64)
            assert hashlib.sha256(np.ascontiguousarray(bm).tobytes()).hexdigest() == r["bond_map_sha"]
        if r["bond_flag"] and r.get("dumps"):
            alld = np.concatenate([o["dumps"], o["final_bonds"][None, :]], axis=0)
            txt = compat_io.bonds_text(alld)
            want = "".join("".join("%d %s\n" % (b, ch) for b, ch in enumerate(d)) for d in r["dumps"])
            assert txt == want
        assert compat_io.temperature_file_tag(T) == r["bonds_file"][0][len("bonds_"):-len(".txt")] if r["bonds_file"] else True
        checked += 1
    assert checked >= 5


def test_blocked_config_file_format(tmp_path):
    from worm_algorithm_b200.block_configs import save_blocked_configs
    cfg = np.arange(32).reshape(2, 16) % 2
    out = save_blocked_configs(cfg, "2.0", str(tmp_path))
    assert out.endswith("2_blocked_configs_2.0.txt")
    assert open(out).read() == "0 " + " ".join(str(v) for v in cfg[0]) + "\n1 " + " ".join(str(v) for v in cfg[1]) + "\n"
    save_blocked_configs(cfg, "2.0", str(tmp_path))
    assert os.path.exists(out + ".bak")


# ------------------------------------------------------------------------------------------------
# multi-rank plumbing
# ------------------------------------------------------------------------------------------------
def test_shard_range_partitions_the_chain_range():
    from worm_algorithm_b200.dist import shard_range
    for n in (0, 1, 7, 64, 4096, 1000003):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as td
from worm_algorithm_b200 import dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
td.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=rank, world_size=world)
n, nT = 64 * 6, 6                                  # replica-major chain table, like WormSimulation
t_idx = np.tile(np.arange(nT), n // nT)
rng = np.random.default_rng(0)
acc = rng.integers(0, 2 ** 40, (n, 8)).astype(np.int64)      # stand-in for the kernel's per-chain sums
c0, c1 = dist.shard_range(n, rank, world)
per_T = torch.zeros((nT, 8), dtype=torch.int64)
per_T.index_add_(0, torch.from_numpy(t_idx[c0:c1].astype(np.int64)), torch.from_numpy(acc[c0:c1]))
per_T = dist.allreduce_sum(per_T, world)
full = np.zeros((nT, 8), dtype=np.int64)
np.add.at(full, t_idx, acc)
assert (per_T.numpy() == full).all(), "rank %d" % rank
# the optional per-chain gather (SURVEY 8e): every rank ends up with all rows, in global chain order
for n2 in (n, n - 5, 3):
    a0, a1 = dist.shard_range(n2, rank, world)
    got = dist.allgather_rows(torch.from_numpy(acc[a0:a1]), n2, rank, world)
    assert got.shape == (n2, 8) and (got.numpy() == acc[:n2]).all(), "gather rank %d n %d" % (rank, n2)
td.destroy_process_group()
print("ok", rank)
'''


def test_two_rank_gloo_reduce_equals_single_rank(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    port = str(29000 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT, port], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the driver's CPU arm) prints ONE JSON line with the contract's keys; a tiny
    sample here (the unmodified executable when oracle/_ref is built, else the C port)."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-steps", "20000", "--ref-runs-per-core", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "worm steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["gpu_launches"] == 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["config"]["workload"].startswith("L=32 temperature sweep")


def test_bench_reference_arm_other_ranks_are_silent():
    """Under torchrun only rank 0 runs the reference arm; the other ranks exit 0 without output."""
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
