"""GPU tests of the statistics and of the host-side mirrors of the reference's post-processing
classes (run on a B200: `pytest -m gpu`).

Production mode is a different random stream from the reference's, so parity there is statistical:
energy and specific heat against Kaufman's exact finite-size result within 3 sigma, sigma = standard
error over independent replicas (seeds fixed, so a pass is reproducible bit for bit)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TC = 2.0 / np.log(1.0 + np.sqrt(2.0))


@pytest.fixture(scope="module")
def eng():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a B200"
    from worm_algorithm_b200 import engine
    engine.lib()
    return engine


def _run(eng, L, temps, reps, steps, algo, hist=False, lanes=0, seed0=1000):
    temps = np.asarray(temps, dtype=np.float64)
    T = np.tile(temps, reps)                                  # replica-major
    n = T.size
    seeds = (np.arange(n) + seed0).astype(np.uint64)
    st = eng.PhiloxState(L, T, seeds, algo=algo, group_index=np.arange(n) if hist else None, hist=hist)
    st.run(steps, therm_steps=steps // 10, lanes_per_chain=lanes)
    return st, T.reshape(reps, temps.size)


def _family_rule(z, what):
    """The north star's "within 3 sigma" for a FAMILY of m checks with independent honest error bars: a single
    |z| > 3 has probability 0.0027, so among m = 128 checks one is expected every third run, two or more in 5 % of
    runs, and one beyond 4 sigma in 0.8 %.  Rule (written before looking at the data, seeds fixed): at most
    ceil(m / 128) values beyond 3 sigma, none beyond 4 sigma, and mean z^2 < 1 + 4 sqrt(2 / m) (the chi^2 of the
    family, 4 standard deviations of its own sampling noise)."""
    z = np.asarray(z, dtype=np.float64)
    m = z.size
    assert (np.abs(z) > 3.0).sum() <= -(-m // 128), (what, np.sort(np.abs(z))[-4:])
    assert (np.abs(z) < 4.0).all(), (what, np.abs(z).max())
    assert np.mean(z ** 2) < 1.0 + 4.0 * np.sqrt(2.0 / m), (what, np.mean(z ** 2))


def _check_vs_kaufman(acc, T2d, L, algo, nsig=3.0):
    """E/N and C/N from the sums pooled over replicas (delete-one-replica jackknife) vs Kaufman."""
    from oracle import kaufman as kf                  # (worm_algorithm_b200.kaufman is the product's own; tests/test_host.py
    from worm_algorithm_b200 import stats             #  checks the two against each other and against enumeration)
    N = L * L
    est = stats.pooled_jackknife(acc.reshape(T2d.shape + (8,)), T2d[0], N, algo)
    zs = []
    for name in ("E", "C"):
        m, se = est[name]
        for j, T in enumerate(T2d[0]):
            exact = kf.energy_and_specific_heat(float(T), L)[0 if name == "E" else 1]
            z = (m[j] - exact) / se[j]
            zs.append((name, float(T), m[j], exact, se[j], z))
            assert abs(z) < nsig, (L, algo, name, T, m[j], exact, se[j], z)
    return zs


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", [0, 1])
def test_energy_and_specific_heat_match_kaufman_L16(eng, algo):
    """BASELINE configs[0] lattice (L = 16) at T_c and around it; 64 replicas per temperature."""
    st, T2d = _run(eng, 16, [1.8, 2.269, TC, 2.6, 3.2], 64, 2_000_000, algo)
    _check_vs_kaufman(st.acc.cpu().numpy(), T2d, 16, algo)


def test_temperature_sweep_L32_matches_kaufman(eng):
    """BASELINE configs[1] shape: 64 temperatures in [1.5, 3.5] x 64 seed replicas at L = 32 (shorter
    chains than the bench), per-temperature sums straight from the kernel's group accumulators."""
    from oracle import kaufman as kf
    from worm_algorithm_b200.worm_simulation import WormSimulation
    temps = 1.5 + 2.0 * np.arange(64) / 63.0
    sim = WormSimulation(L=32, num_steps=2_000_000, verbose=False, T_arr=temps, bond_flag=0, n_replicas=64,
                         rng="philox")
    acc = sim.results["acc"].cpu().numpy().reshape(64, 64, 8)       # [replica, T, 8]
    T2d = np.tile(temps, (64, 1))
    zs = _check_vs_kaufman(acc, T2d, 32, 0, nsig=4.0)              # 128 checks: the family-wise 3-sigma rule
    _family_rule([r[5] for r in zs], "C2 sweep, Taylor chain")
    # the reference-style result dicts hold exactly the pooled estimate of the kernel's per-temperature sums
    tot = acc.sum(axis=0)
    for j in (0, 31, 63):
        key = str(float("%.12f" % temps[j])).rstrip('0')      # the reference's keys: T after the "%.12f" text round trip
        assert abs(sim._E[key] - (-temps[j] * tot[j, 1] / tot[j, 0] / 1024.0)) < 1e-11
        assert sim.results["per_T"][j, 0] == tot[j, 0] and sim.results["per_T"][j, 2] == tot[j, 2]


def test_binary_chain_temperature_sweep_L32_matches_kaufman(eng):
    """The north star's chain (binary bonds, tanh K acceptance, bit-packed kernel) on the C2 shape: 64 temperatures
    x 64 replicas at L = 32, E/N and C/N through the binary estimators of SURVEY App. D, family-wise 3-sigma rule."""
    temps = 1.5 + 2.0 * np.arange(64) / 63.0
    st, T2d = _run(eng, 32, temps, 64, 2_000_000, 1)
    zs = _check_vs_kaufman(st.acc.cpu().numpy(), T2d, 32, 1, nsig=4.0)
    _family_rule([r[5] for r in zs], "C2 sweep, binary chain")
    assert (st.acc[:, 6] == 0).all()


def test_binary_chain_L64_near_Tc_matches_kaufman(eng):
    """Binary chain on the C3 lattice and temperature grid (L = 64 around T_c)."""
    temps = [2.21, 2.22, 2.23, 2.24, 2.26, 2.27, 2.28, 2.269]
    st, T2d = _run(eng, 64, temps, 64, 24_000_000, 1)
    zs = _check_vs_kaufman(st.acc.cpu().numpy(), T2d, 64, 1, nsig=4.0)
    _family_rule([r[5] for r in zs], "L = 64 near T_c, binary chain")


def test_real_dumps_L64_L128_through_the_pipeline_match_the_oracle_elementwise(eng):
    """BASELINE configs[2] on REAL dumps (not random bytes): L = 64 and 128 near T_c, dumped configurations ->
    images -> every blocking level -> counts, compared element by element with oracle/postproc_oracle.py (pinned to
    block_configs.py:6-71 / iterated_blocking.py:6-59 / bonds.py:241-319), for the launch chain and the fused kernel."""
    import torch
    from oracle import postproc_oracle as po
    for L, steps in ((64, 400_000), (128, 1_200_000)):
        T = [2.21, 2.24, 2.269, 2.28]
        st = eng.PhiloxState(L, T, np.arange(4, dtype=np.uint64) + 11, algo=0, max_dumps=3)
        st.run(steps, therm_steps=steps // 10, dump_every=50)
        nd = st.n_dumps.cpu().numpy()
        assert (nd >= 1).all(), nd
        dumps = torch.stack([st.dumps[c, min(int(nd[c]), 3) - 1] for c in range(4)])      # last stored dump of every chain
        levels = L.bit_length() - 1
        counts, imgs = eng.pipeline(dumps, L, images=range(levels))
        chain = [eng.image(dumps, L)]
        while chain[-1].shape[-1] > 4:
            chain.append(eng.block(chain[-1], 0))
        d = dumps.cpu().numpy()
        for c in range(4):
            img = po.image_from_bonds(d[c], L)
            lv = L
            for k in range(levels):
                assert (imgs[k][c].cpu().numpy() == img).all() and (chain[k][c].cpu().numpy() == img).all(), (L, c, k)
                assert int(counts[c, k]) == int(po.bond_counts(img.reshape(1, -1), 2 * lv)[0])
                if k + 1 < levels:
                    img = po.block_config(img.reshape(-1), 1, 0).reshape(lv, lv)
                    lv //= 2
            assert int(counts[c, 0]) == int((d[c] & 1).sum()) and int(counts[c, 0]) > 0


@pytest.mark.parametrize("L,steps", [(8, 400_000), (16, 1_500_000), (32, 6_000_000), (64, 24_000_000), (128, 100_000_000),
                                     (256, 600_000_000)])
def test_finite_size_scaling_specific_heat(eng, L, steps):
    """BASELINE configs[3]: C/N at T = 2.269 for L = 8 .. 256 against Kaufman (fluctuation estimator)."""
    st, T2d = _run(eng, L, [2.269], 64, steps, 0)
    _check_vs_kaufman(st.acc.cpu().numpy(), T2d, L, 0)


def test_product_fss_driver(eng):
    """worm_algorithm_b200.kaufman.finite_size_scaling: the product-side C_V-vs-exact comparison the reference does
    in kauffman_specific_heat.ipynb (its exact curve is a missing data file)."""
    from worm_algorithm_b200 import kaufman as pk
    rows = pk.finite_size_scaling(Ls=(8, 16, 32), T=2.269, n_replicas=64, steps_per_site=6000)
    assert [r["L"] for r in rows] == [8, 16, 32]
    for r in rows:
        assert abs(r["z_E"]) < 3.0 and abs(r["z_C"]) < 3.0, r
    assert rows[0]["C_exact"] < rows[1]["C_exact"] < rows[2]["C_exact"]        # C(T_c) grows like ln L


def test_multi_gpu_run_equals_single_gpu_run(eng):
    """N-GPU result equality of the reduced accumulators.  Needs two visible GPUs (skipped on a
    one-GPU box; `gpurun --gpus 2 -- python -m pytest tests -m gpu -k multi_gpu`): two ranks under torchrun shard the
    chains of a sweep with G(r) histograms, all-reduce over NCCL, and rank 0 compares with its own single-rank run."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU visible")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(root, "tools", "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-1500:]
    assert "identical to the single-rank run: True; result dicts identical: True" in p.stdout


def test_correlation_function_G_r(eng):
    """G(r) histogram: bin 0 = Z, sum of bins = measured iterations (so sum G / G(0) = 1/Z_av = chi),
    G(1,0)/G(0) = <s0 s1> = -(E/N)/2 (Kaufman), and the x / y symmetry of the torus."""
    from oracle import kaufman as kf
    from worm_algorithm_b200 import stats
    L, T = 16, 2.269
    st, _ = _run(eng, L, [T], 64, 3_000_000, 0, hist=True)
    acc = st.acc.cpu().numpy()
    h = st.hist.cpu().numpy().astype(np.float64)                    # one row per chain
    assert (h[:, 0] == acc[:, 0]).all() and (h.sum(axis=1) == acc[:, 5]).all()
    g = h / h[:, :1]
    for r_bin in (1, L):                                            # (dx,dy) = (1,0) and (0,1)
        m, se = stats.replica_mean_err(g[:, r_bin])
        exact = -kf.energy_and_specific_heat(T, L)[0] / 2.0
        assert abs(m - exact) < 3 * se, (r_bin, m, exact, se)
    gx, sx = stats.replica_mean_err(g[:, 1] - g[:, L])
    assert abs(gx) < 3.5 * sx


def test_production_mode_matches_reference_executable_statistics(eng, golden_dir):
    """BASELINE north star: the Philox mode reproduces the REFERENCE's own estimates within 3 sigma.
    `tests/golden/ref_stats.json` holds Z_av and E_av of the unmodified executable for 24 seeds at 8
    temperatures of L = 16 (oracle/make_golden.py stats).  Same chain law, same start (empty lattice,
    head = tail = 0), same estimator: the reference never excludes its warm-up from the sums
    (worm_ising_2d.cpp:113 gates only the writing), so the GPU run measures from step 0 as well and both
    have the same expectation.  sigma = standard error over independent seeds on both sides; the
    reference's C_V (finite differences of E(T), specific_heat.py:134-150) is compared on the mean
    curves with a delete-one-seed jackknife on both sides."""
    import json
    from worm_algorithm_b200 import stats
    from worm_algorithm_b200.specific_heat import finite_difference_cv
    with open(os.path.join(golden_dir, "ref_stats.json")) as f:
        g = json.load(f)
    L, steps, temps = g["L"], g["num_steps"], np.array(g["T"])
    N = L * L
    refE, refZ = np.array(g["E_av"]), np.array(g["Z_av"])            # [T][seed]
    reps = 64
    T = np.tile(temps, reps)
    seeds = (np.arange(T.size) + 31337).astype(np.uint64)
    st = eng.PhiloxState(L, T, seeds, algo=0)
    st.run(steps, therm_steps=0)
    acc = st.acc.cpu().numpy().reshape(reps, temps.size, 8)
    obs = stats.taylor_observables(acc, temps, N)                    # per chain, like one output.txt each
    for name, ref in (("E", refE), ("Z_av", refZ)):
        m, se = stats.replica_mean_err(obs[name], axis=0)
        rm, rse = stats.replica_mean_err(ref, axis=1)
        z = (m - rm) / np.sqrt(se ** 2 + rse ** 2)
        assert (np.abs(z) < 3.0).all(), (name, z)

    def cv_with_err(E):                                              # E: [seed][T]
        R = E.shape[0]
        full = np.asarray(finite_difference_cv(temps, E.mean(axis=0))[1])
        loo = np.array([finite_difference_cv(temps, np.delete(E, k, axis=0).mean(axis=0))[1] for k in range(R)])
        return full, np.sqrt((R - 1) / R * ((loo - loo.mean(axis=0)) ** 2).sum(axis=0))
    c, ce = cv_with_err(obs["E"])
    rc, rce = cv_with_err(refE.T)
    zc = (c - rc) / np.sqrt(ce ** 2 + rce ** 2)
    assert c.size == temps.size - 4 and (np.abs(zc) < 3.0).all(), zc


def test_bond_flag_run_switches_the_dump_rule_off_after_the_cap(eng):
    """WormSimulation(bond_flag=1, rng='philox') keeps the first max_dumps dumps per chain and then runs the rest
    without the dump rule (a segment boundary every write_steps steps halves the step rate): dumps, events,
    sums and final configurations are those of one run with the rule on throughout."""
    from worm_algorithm_b200.worm_simulation import WormSimulation
    L, steps = 16, 300_000
    sim = WormSimulation(L=L, num_steps=steps, verbose=False, T_arr=[2.6, 3.0, 3.4], bond_flag=1, n_replicas=8,
                         rng="philox", max_dumps=8)        # (a paramagnetic worm closes often: the caps fill early)
    st = sim.results["state"]
    ref = eng.PhiloxState(L, sim.results["T"], sim.results["seeds"], algo=0, chain_id0=0,
                          group_index=sim.results["t_index"], n_groups=3, max_dumps=8)
    ref.run(steps, therm_steps=steps // 10, dump_every=50)
    assert (st.dumps == ref.dumps).all() and (st.events == ref.events).all()
    assert (st.acc == ref.acc).all() and (st.bonds == ref.bonds).all() and (st.group_acc == ref.group_acc).all()
    assert int(st.n_dumps.min()) >= 8
    assert int(st.n_dumps.max()) < int(ref.n_dumps.max())           # the rule really was switched off


def test_lat_and_wide_kernels_agree_at_scale(eng):
    """BASELINE configs[4] shape, scaled: 16384 independent L = 32 chains with G(r) histograms; the
    thread-per-chain kernel and the latency kernel give identical reduced sums and histograms."""
    L, nT, reps = 32, 16, 1024
    temps = np.linspace(1.5, 3.5, nT)
    T = np.tile(temps, reps)
    n = T.size
    gi = np.tile(np.arange(nT), reps)
    seeds = (np.arange(n) // nT).astype(np.uint64)
    a = eng.PhiloxState(L, T, seeds, algo=0, group_index=gi, n_groups=nT, hist=True).run(3000, therm_steps=300, lanes_per_chain=1)
    b = eng.PhiloxState(L, T, seeds, algo=0, group_index=gi, n_groups=nT, hist=True).run(3000, therm_steps=300, lanes_per_chain=32)
    assert (a.group_acc == b.group_acc).all() and (a.hist == b.hist).all() and (a.acc == b.acc).all()
    assert int(a.group_acc[:, 5].sum()) == n * 2700


@pytest.mark.parametrize("L", [4, 16, 32])
def test_solo_histogram_bins_survive_their_flushes(eng, L):
    """Solo latency mode keeps 16-bit G(r) bins per chain in shared memory and sweeps them into the
    int64 histogram every 32768 steps: a run across several sweeps (resumed once, so that the chains
    start open) gives the histogram of the thread-per-chain kernel bit for bit."""
    nT, reps = 5, 20
    T = np.tile(np.linspace(1.5, 3.5, nT), reps)
    n = T.size                                  # 100 chains: three full walkers and a ragged one
    gi = np.tile(np.arange(nT), reps)
    seeds = (np.arange(n) // nT + 7).astype(np.uint64)
    runs = []
    for lanes in (1, 2):
        st = eng.PhiloxState(L, T, seeds, algo=0, group_index=gi, n_groups=nT, hist=True)
        st.run(1001, therm_steps=100, lanes_per_chain=lanes)
        st.run(80000, therm_steps=100, lanes_per_chain=lanes)
        runs.append(st)
    a, b = runs
    assert (a.hist == b.hist).all() and (a.acc == b.acc).all() and (a.bonds == b.bonds).all()
    h = b.hist.cpu().numpy()
    assert h.sum() == n * 80901
    assert (h[:, 0] == b.group_acc[:, 0].cpu().numpy()).all()        # G(0) = Z


# ------------------------------------------------------------------------------------------------
# host-side mirrors of the reference's post-processing classes
# ------------------------------------------------------------------------------------------------
def test_block_and_count_mirrors_match_reference_golden(eng, post_golden):
    from worm_algorithm_b200 import block_configs as bc, iterated_blocking as ib
    from worm_algorithm_b200.count_bonds import bond_counts, count_stats_with_err
    g = post_golden
    tags = sorted({k[:-3] for k in g.files if k.startswith("blk_") and k.endswith("_in")})
    for tag in tags:
        img = g[tag + "_in"].astype(np.int64)
        assert (bc._block_config(img.flatten(), 1, 0) == g[tag + "_bc0"]).all()
        assert (bc._block_config(img.flatten(), 1, 2) == g[tag + "_bc2"]).all()
        if tag + "_ib" in g.files:
            assert (ib._block_config(img.flatten(), 1) == g[tag + "_ib"]).all()
        s = 2
        while tag + "_ib_s%d" % s in g.files:
            assert (ib._block_config(img.flatten(), s) == g[tag + "_ib_s%d" % s]).all()
            assert (bc._block_config(img.flatten(), s, 0) == g[tag + "_ib_s%d" % s]).all()
            s += 1
    for L in (4, 8, 16):
        tag = "img_L%d" % L
        cnt = bond_counts(g[tag + "_images"].astype(np.int32), 2 * L)
        ref = g[tag + "_count_err"]
        full, err = count_stats_with_err(cnt, int(ref[4]))
        assert [full[0], err[0], full[1], err[1]] == list(ref[:4])      # same floats as CountBonds._count_bonds_with_err


def test_drop_in_pipeline_on_reference_format_tree(eng, post_golden, tmp_path):
    """WormSimulation (deterministic mode) -> reference-format data tree -> Bonds -> config files ->
    block_configs / CountBonds, all through the mirrors; images equal the ones the reference's own
    Bonds class produced for the same run (golden)."""
    from worm_algorithm_b200.bonds import Bonds
    from worm_algorithm_b200.count_bonds import CountBonds
    from worm_algorithm_b200 import block_configs as bc, iterated_blocking as ib
    g = post_golden
    L, T, steps, seed = 8, 2.269, 60000, 9.0                      # the run behind img_L8_* (oracle/make_golden.py)
    data = str(tmp_path / "data")
    cfg_dir = str(tmp_path / "configs") + "/"
    b = Bonds(L, run=True, num_steps=steps, verbose=False, T_arr=[T], seeds=[seed], rng="mt19937",
              data_dir=data, config_dir=cfg_dir, block_configs=True, block_val=2, write=True)
    (key,) = list(b._config_data.keys())
    imgs = np.array([b._config_data[key][i] for i in range(12)])
    assert (imgs == g["img_L8_images"]).all()
    blk = np.array([b._blocked_config_data[2][key][i] for i in range(12)])
    assert (blk == g["img_L8_blockT2"]).all()
    # the same tree read back from the text files (what a reference user has on disk)
    b2 = Bonds(L, run=False, verbose=False, T_arr=[T], data_dir=data, write=False)
    (key2,) = list(b2._config_data.keys())
    assert len(b2._config_data[key2]) == len(b._config_data[key])
    assert all((b2._config_data[key2][i] == b._config_data[key][i]).all() for i in range(len(b._config_data[key])))
    # config file -> blocked file -> bond statistics
    (cfg_file,) = [cfg_dir + f for f in os.listdir(cfg_dir)]
    blocked = ib.block_configs(cfg_file, out_dir=str(tmp_path / "blocked") + "/")
    assert blocked.shape == (len(b._config_data[key]), (L) * (L))
    assert (blocked[:12] == np.array([bc._block_config(im.flatten()) for im in imgs])).all()
    cb = CountBonds(L, num_blocks=3, data_dir=cfg_dir, save_dir=str(tmp_path / "stats") + "/")
    (val,) = list(cb.bond_stats.values())
    dumps = np.concatenate([b.results["dumps"][0, :int(b.results["n_events"][0])].cpu().numpy(),
                            b.results["bonds"].cpu().numpy()], axis=0)
    n = (dumps & 1).sum(axis=1)
    assert val[0] == np.mean(n) and abs(val[2] - (np.mean(n ** 2) - np.mean(n) ** 2)) < 1e-9
    assert os.path.exists(str(tmp_path / "stats") + "/bond_stats_8.txt")


def test_iterated_blocking_chain_on_device(eng):
    """BASELINE configs[2]: L = 64 dumps near T_c -> images -> blocking chain down to lattice 2 ->
    counts per level, without leaving the GPU; closed loops stay closed at every level."""
    import torch
    from worm_algorithm_b200 import iterated_blocking as ib
    L = 64
    T = [2.21, 2.22, 2.23, 2.24, 2.26, 2.27, 2.28, 2.269]
    st = eng.PhiloxState(L, T, np.arange(8, dtype=np.uint64), algo=0, max_dumps=16)
    st.run(600_000, therm_steps=60_000, dump_every=50)
    k = int(st.n_dumps.min().clamp(max=16))
    assert k > 0
    imgs = eng.image(st.dumps[:, :k], L)
    levels = ib.iterate(imgs.reshape(-1, 2 * L, 2 * L))
    assert [lv.shape[-1] for lv in levels] == [64, 32, 16, 8, 4]
    for lv in levels:
        xb, yb = lv[..., 0::2, 1::2], lv[..., 1::2, 0::2]
        deg = xb + yb + torch.roll(xb, 1, dims=-1) + torch.roll(yb, 1, dims=-2)
        assert (deg % 2 == 0).all()
        assert (eng.count(lv) == (xb.sum((-1, -2)) + yb.sum((-1, -2)))).all()
