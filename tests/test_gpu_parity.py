"""GPU parity tests (run on a B200: `pytest -m gpu`).  Every check goes through the C ABI
(libworm_b200.so via worm_algorithm_b200.engine) and compares with the oracle / golden vectors.
Bar: bit-exact (integer state, accumulators, images)."""
import ctypes as C
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def eng():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a B200"
    from worm_algorithm_b200 import engine
    engine.lib()                      # fails loudly if the extension was not built
    return engine


# ------------------------------------------------------------------------------------------------
# deterministic mode
# ------------------------------------------------------------------------------------------------
def test_mt_mode_matches_reference_golden(eng, ref_runs):
    """All golden (L, T, steps, seed) cases in ONE batched launch: dumps, Z, Nb_tot, step_num, Nb."""
    recs = [r for r in ref_runs if r["num_steps"] <= 400000]
    by_L = {}
    for r in recs:
        by_L.setdefault(r["L"], []).append(r)
    for L, group in by_L.items():
        res = eng.run_mt(L, [r["T"] for r in group], [r["seed"] for r in group],
                         group[0]["num_steps"], bond_flag=1, max_events=1024) \
            if len({r["num_steps"] for r in group}) == 1 else None
        for i, r in enumerate(group):
            one = res if res is not None else eng.run_mt(L, [r["T"]], [r["seed"]], r["num_steps"], bond_flag=1, max_events=1024)
            k = i if res is not None else 0
            out = one.out[k].cpu().numpy()
            N = L * L
            T = r["T"]
            fields = r["output_line"].split()
            assert out[4] == 0
            assert int(out[2]) == r["step_num"]
            assert "%.12f" % (int(out[0]) / (int(out[2]) * 1.0)) == fields[3]
            assert "%.12f" % (int(out[1]) * T / (int(out[0]) * N)) == fields[4]
            assert int(out[3]) == r["Nb_final"]
            ne = int(one.n_events[k])
            assert ne == r["n_observables"]
            if r["bond_flag"]:
                dumps = one.dumps[k, :ne].cpu().numpy().astype(np.int32)
                alld = np.concatenate([dumps, one.bonds[k].cpu().numpy().astype(np.int32)[None, :]], axis=0)
                assert _sha(alld) == r["dumps_sha"]


def test_mt_mode_long_run_survey_vector(eng):
    # L=32, T=2.269, 1e7 steps, seed 12345
    res = eng.run_mt(32, [2.269], [12345.0], 10000000, bond_flag=0, max_events=0)
    out = res.out[0].cpu().numpy()
    assert int(out[2]) == 10023006
    assert "%.12f" % (int(out[0]) / int(out[2])) == "0.002276263229"
    assert "%.12f" % (int(out[1]) * 2.269 / (int(out[0]) * 1024)) == "-1.407139231916"


def test_mt_mode_baseline_config0(eng, ref_runs):
    """BASELINE.json configs[0]: L = 16 at T_c = 2.269, one chain, 1e5 sweeps (5.12e7 steps), against the
    reference executable's own output line (tests/golden/ref_runs.json, made by oracle/make_golden.py)."""
    rec = [r for r in ref_runs if r["num_steps"] == 51200000][0]
    res = eng.run_mt(rec["L"], [rec["T"]], [rec["seed"]], rec["num_steps"], bond_flag=0, max_events=0)
    out = res.out[0].cpu().numpy()
    fields = rec["output_line"].split()
    assert int(out[4]) == 0 and int(out[2]) == rec["step_num"] == int(fields[6])
    assert "%.12f" % (int(out[0]) / int(out[2])) == fields[3]
    assert "%.12f" % (int(out[1]) * rec["T"] / (int(out[0]) * rec["L"] ** 2)) == fields[4]
    assert int(out[3]) == rec["Nb_final"]


def test_mt_mode_batch_matches_oracle_with_trace(eng):
    from oracle import worm_oracle as wo
    rng = np.random.default_rng(11)
    for L in (2, 3, 8, 13):
        n = 40
        T = rng.uniform(0.7, 4.0, n)
        seeds = rng.integers(0, 2 ** 32, n).astype(np.float64) + 0.25
        steps = 3000
        res = eng.run_mt(L, T, seeds, steps, bond_flag=1, max_events=64, max_trace=128)
        out = res.out.cpu().numpy()
        for c in range(n):
            o = wo.run_mt(L, float(T[c]), steps, float(seeds[c]), 1, max_events=64, max_trace=128)
            assert [int(v) for v in out[c, :4]] == [o["Z"], o["Nb_tot"], o["step_num"], o["Nb_final"]]
            assert (res.bonds[c].cpu().numpy() == o["final_bonds"]).all()
            ne = min(o["n_events"], 64)
            assert int(res.n_events[c]) == o["n_events"]
            assert (res.events[c, :ne].cpu().numpy() == o["events"]).all()
            assert (res.dumps[c, :ne].cpu().numpy() == o["dumps"]).all()
            nt = min(o["n_trace"], 128)
            assert int(res.n_trace[c]) == o["n_trace"]
            assert (res.trace[c, :nt].cpu().numpy() == o["trace"]).all()


# ------------------------------------------------------------------------------------------------
# production mode: every kernel variant reproduces the CPU restatement bit for bit
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", [0, 1])
@pytest.mark.parametrize("lanes", [0, 1, 2, 4, 8, 32, -1, -2])
@pytest.mark.parametrize("L,tmin", [(3, 0.8), (8, 0.8), (16, 1.0), (32, 1.2)])
def test_philox_matches_oracle(eng, algo, lanes, L, tmin):
    """Every kernel variant (latency mode with 4/8/32 lanes per chain, thread per chain with bonds in
    shared or global memory) and both record builders (T >= 1 fast path, general T) against the CPU
    restatement: accumulators, final bonds, head/tail, dumps, events, G(r), per-group sums."""
    from oracle import worm_oracle as wo
    if lanes in (2, 4, 8, 32, -2) and (L & (L - 1)):
        pytest.skip("latency mode and packed mode need a power-of-two L")
    rng = np.random.default_rng(100 * L + 10 * algo + lanes + 1)
    n = 37                      # not a multiple of the chains per warp: exercises dead lanes
    T = rng.uniform(tmin, 3.5, n)
    T[:8] = 2.269
    seeds = rng.integers(0, 2 ** 63, n, dtype=np.uint64)
    hidx = np.arange(n) % 3
    steps, therm, cid0 = 2500, 400, 1000
    st = eng.PhiloxState(L, T, seeds, algo=algo, chain_id0=cid0, group_index=hidx, n_groups=3, hist=True, max_dumps=16)
    st.run(1000, therm_steps=therm, dump_every=50, lanes_per_chain=lanes)      # resumable: two launches
    st.run(steps - 1000, therm_steps=therm, dump_every=50, lanes_per_chain=lanes)
    acc = st.acc.cpu().numpy()
    bonds = st.bonds.cpu().numpy()
    head, tail = st.head.cpu().numpy(), st.tail.cpu().numpy()
    nd = st.n_dumps.cpu().numpy()
    ev = st.events.cpu().numpy()
    dm = st.dumps.cpu().numpy()
    hist = np.zeros((3, L * L), dtype=np.int64)
    for c in range(n):
        o = wo.run_philox(L, algo, cid0 + c, int(seeds[c]), float(T[c]), 0, steps, therm, hist=True,
                          dump_every=50, max_dumps=16)
        assert (acc[c, :6] == o["acc"][:6]).all(), (c, acc[c], o["acc"])
        assert acc[c, 6] == 0                      # WORM_ACC_STATUS: nothing the packed format could not hold
        assert (bonds[c] == o["bonds"]).all()
        assert head[c] == o["head"] and tail[c] == o["tail"]
        assert nd[c] == o["n_dumps"]
        k = min(o["n_dumps"], 16)
        assert (ev[c, :k] == o["events"]).all()
        assert (dm[c, :k] == o["dumps"]).all()
        hist[hidx[c]] += o["hist"]
    assert (st.hist.cpu().numpy() == hist).all()
    gacc = np.zeros((3, 8), dtype=np.int64)
    np.add.at(gacc, hidx, acc)
    assert (st.group_acc.cpu().numpy() == gacc).all()


def test_philox_nibble_solo_L64_matches_oracle(eng):
    """L = 64 runs 32 chains per SM in the latency kernel's nibble layout with a solo walker (what lanes_per_chain = 0
    picks for the Taylor chain): 70 chains = two full blocks and a ragged one, dumps and events, resumed once."""
    from oracle import worm_oracle as wo
    from worm_algorithm_b200 import _lib
    n, L = 70, 64
    assert _lib.philox_choice(L, 0, 4096, False) == 2
    rng = np.random.default_rng(64)
    T = rng.uniform(1.0, 3.5, n)
    T[:5] = 2.269
    seeds = rng.integers(0, 2 ** 63, n, dtype=np.uint64)
    st = eng.PhiloxState(L, T, seeds, algo=0, chain_id0=9, max_dumps=5)
    st.run(4001, therm_steps=500, dump_every=50, lanes_per_chain=2).run(9999, therm_steps=500, dump_every=50, lanes_per_chain=0)
    acc, bonds = st.acc.cpu().numpy(), st.bonds.cpu().numpy()
    for c in range(0, n, 3):
        o = wo.run_philox(L, 0, 9 + c, int(seeds[c]), float(T[c]), 0, 14000, 500, dump_every=50, max_dumps=5)
        assert (acc[c, :6] == o["acc"][:6]).all() and acc[c, 6] == 0 and (bonds[c] == o["bonds"]).all()
        k = min(o["n_dumps"], 5)
        assert int(st.n_dumps[c]) == o["n_dumps"] and (st.dumps[c, :k].cpu().numpy() == o["dumps"]).all()
        assert (st.events[c, :k].cpu().numpy() == o["events"]).all()


@pytest.mark.parametrize("algo", [0, 1])
@pytest.mark.parametrize("L,n", [(8, 6007), (16, 80000), (32, 33152 + 4000), (64, 9000)])
def test_philox_packed_big_batches(eng, algo, L, n):
    """Packed mode with real blocks: whole warps plus a ragged last warp, several waves, dead lanes in the last
    block.  Every chain against the thread-per-chain kernel (itself pinned to the oracle above), a sample of
    chains against the oracle directly."""
    from oracle import worm_oracle as wo
    rng = np.random.default_rng(5 * L + algo)
    nT = 16
    gi = (np.arange(n) * nT // n).astype(np.int32)
    T = (1.2 + 2.3 * np.arange(nT) / (nT - 1))[gi]
    seeds = rng.integers(0, 2 ** 63, n, dtype=np.uint64)
    steps, therm = 1500, 300
    kw = dict(algo=algo, chain_id0=7, group_index=gi, n_groups=nT, hist=True, max_dumps=3)
    a = eng.PhiloxState(L, T, seeds, **kw).run(700, therm_steps=therm, dump_every=50, lanes_per_chain=-2)
    a.run(steps - 700, therm_steps=therm, dump_every=50, lanes_per_chain=-2)
    b = eng.PhiloxState(L, T, seeds, **kw).run(steps, therm_steps=therm, dump_every=50, lanes_per_chain=1)
    assert (a.acc == b.acc).all() and (a.bonds == b.bonds).all()
    assert (a.head == b.head).all() and (a.tail == b.tail).all()
    assert (a.hist == b.hist).all() and (a.group_acc == b.group_acc).all()
    assert (a.n_dumps == b.n_dumps).all() and (a.events == b.events).all() and (a.dumps == b.dumps).all()
    acc, bonds = a.acc.cpu().numpy(), a.bonds.cpu().numpy()
    for c in (0, 31, 32, n // 2, n - 1):
        o = wo.run_philox(L, algo, 7 + c, int(seeds[c]), float(T[c]), 0, steps, therm, hist=True, dump_every=50, max_dumps=3)
        assert (acc[c, :6] == o["acc"][:6]).all() and acc[c, 6] == 0 and (bonds[c] == o["bonds"]).all()


@pytest.mark.parametrize("lanes", [2, 4, 8, 32])
@pytest.mark.parametrize("L,tmin", [(4, 1.0), (8, 0.8), (16, 1.0), (32, 1.2)])
def test_latency_kernel_nibble_layout_matches_oracle(eng, monkeypatch, L, tmin, lanes):
    """The latency kernel's nibble layout (one byte per site, 1 / 4 / 8 chains per walker) is what large lattices
    run on (L >= 64, tests below); pinned here on small ones so that every path -- both record builders, segment
    edges, dumps, events, G(r), resumed launches -- is compared with the oracle chain by chain."""
    from oracle import worm_oracle as wo
    monkeypatch.setenv("WORM_LAT_LAYOUT", "nibble")
    rng = np.random.default_rng(7 * L + lanes)
    n = 45
    T = rng.uniform(tmin, 3.5, n)
    T[:6] = 2.269
    seeds = rng.integers(0, 2 ** 63, n, dtype=np.uint64)
    hidx = np.arange(n) % 3
    steps, therm, cid0 = 2311, 390, 77
    st = eng.PhiloxState(L, T, seeds, algo=0, chain_id0=cid0, group_index=hidx, n_groups=3, hist=True, max_dumps=9)
    for ch in (1, 700, 33, steps - 734):
        st.run(ch, therm_steps=therm, dump_every=50, lanes_per_chain=lanes)
    acc, bonds = st.acc.cpu().numpy(), st.bonds.cpu().numpy()
    head, tail, nd = st.head.cpu().numpy(), st.tail.cpu().numpy(), st.n_dumps.cpu().numpy()
    ev, dm = st.events.cpu().numpy(), st.dumps.cpu().numpy()
    hist = np.zeros((3, L * L), dtype=np.int64)
    for c in range(n):
        o = wo.run_philox(L, 0, cid0 + c, int(seeds[c]), float(T[c]), 0, steps, therm, hist=True, dump_every=50, max_dumps=9)
        assert (acc[c, :6] == o["acc"][:6]).all() and acc[c, 6] == 0, (c, acc[c], o["acc"])
        assert (bonds[c] == o["bonds"]).all() and head[c] == o["head"] and tail[c] == o["tail"]
        assert nd[c] == o["n_dumps"]
        k = min(o["n_dumps"], 9)
        assert (ev[c, :k] == o["events"]).all() and (dm[c, :k] == o["dumps"]).all()
        hist[hidx[c]] += o["hist"]
    assert (st.hist.cpu().numpy() == hist).all()
    # an input the layout cannot hold: reported, left untouched, the neighbours unaffected
    st2 = eng.PhiloxState(L, T, seeds, algo=0, chain_id0=cid0)
    st2.bonds[3, 5] = 16
    st2.run(300, lanes_per_chain=lanes)
    ref = eng.PhiloxState(L, T, seeds, algo=0, chain_id0=cid0).run(300, lanes_per_chain=1)
    a2 = st2.acc.cpu().numpy()
    keep = np.arange(n) != 3
    assert a2[3, 6] == 1 and int(st2.bonds[3].sum()) == 16 and (a2[3, :6] == 0).all()
    assert (a2[keep] == ref.acc.cpu().numpy()[keep]).all() and (st2.bonds.cpu().numpy()[keep] == ref.bonds.cpu().numpy()[keep]).all()


def test_philox_many_distinct_temperatures_match_oracle(eng):
    """The per-chain fixed-point thresholds come from a small host-side table keyed by the bits of T (c_api.cu): more
    distinct temperatures than it holds, interleaved so that equal temperatures are far apart, and a few repeats --
    chains from the start, the middle (colliding keys) and beyond the table's capacity against the oracle."""
    from oracle import worm_oracle as wo
    L, n, steps = 4, 2100, 300
    rng = np.random.default_rng(5)
    T = rng.uniform(0.9, 3.6, n)
    T[1::7] = T[0]                                  # one temperature repeated all over the batch
    seeds = rng.integers(0, 2 ** 63, n, dtype=np.uint64)
    st = eng.PhiloxState(L, T, seeds, algo=0).run(steps, therm_steps=30, lanes_per_chain=1)
    acc, bonds = st.acc.cpu().numpy(), st.bonds.cpu().numpy()
    for c in list(range(0, 40)) + list(range(760, 800)) + list(range(n - 40, n)):
        o = wo.run_philox(L, 0, c, int(seeds[c]), float(T[c]), 0, steps, 30)
        assert (acc[c, :6] == o["acc"][:6]).all() and (bonds[c] == o["bonds"]).all(), c


@pytest.mark.parametrize("seed", [2024, 31337])
def test_philox_random_batches_all_variants_agree(eng, seed):
    """Property test over launch geometry: random lattice, batch size, chunking, warm-up, dump period, chain and
    histogram options; every kernel variant that can run the batch (and whatever lanes_per_chain = 0 picks) must give the
    thread-per-chain byte kernel's result bit for bit -- that kernel is pinned to the oracle chain by chain above."""
    rng = np.random.default_rng(seed)
    for case in range(24):
        L = int(rng.choice([4, 8, 16, 32, 64]))
        algo = int(rng.integers(0, 2))
        n = int(rng.choice([1, 3, 31, 33, 100, 257, 1000, 4100 if L <= 32 else 300]))
        nT = int(rng.integers(1, 6))
        gi = (np.arange(n) % nT).astype(np.int32)
        T = rng.uniform(0.9 if case % 5 == 0 else 1.0, 3.6, nT)[gi]
        seeds = rng.integers(0, 2 ** 63, n, dtype=np.uint64)
        hist = bool(rng.integers(0, 2))
        de = int(rng.choice([0, 1, 7, 50, 64] if seed == 2024 else [0, 2, 10, 48, 50, 100]))
        md = int(rng.integers(1, 6)) if de else 0
        chunks = [int(x) for x in rng.integers(1, 900, int(rng.integers(1, 4)))]
        therm = int(rng.integers(0, sum(chunks)))
        cid0 = int(rng.integers(0, 2 ** 40))

        def run(lanes):
            st = eng.PhiloxState(L, T, seeds, algo=algo, chain_id0=cid0, group_index=gi, n_groups=nT, hist=hist, max_dumps=md)
            for ch in chunks:
                st.run(ch, therm_steps=therm, dump_every=de, lanes_per_chain=lanes)
            return st
        ref = run(1)
        for lanes in (0, -2, 2, 4, 8, 32):
            if lanes == 2 and n < 64:
                continue
            try:
                st = run(lanes)
            except Exception as ex:                     # a variant that cannot hold this lattice says so
                assert "latency mode" in str(ex) or "packed mode" in str(ex), (case, lanes, ex)
                continue
            key = (case, L, algo, n, hist, de, md, chunks, therm, lanes)
            assert (st.acc == ref.acc).all() and (st.bonds == ref.bonds).all(), key
            assert (st.head == ref.head).all() and (st.tail == ref.tail).all() and (st.group_acc == ref.group_acc).all(), key
            if hist:
                assert (st.hist == ref.hist).all(), key
            if de:
                assert (st.n_dumps == ref.n_dumps).all() and (st.events == ref.events).all() and (st.dumps == ref.dumps).all(), key


def test_philox_packed_refuses_what_it_cannot_hold(eng):
    """An occupation above 15 cannot be packed in 4 bits: the chain is left untouched and says so in
    WORM_ACC_STATUS (bit 0); the other chains of the batch run normally."""
    L, n = 8, 40
    T = np.full(n, 2.0)
    seeds = np.arange(n, dtype=np.uint64)
    st = eng.PhiloxState(L, T, seeds, algo=0)
    ref = eng.PhiloxState(L, T, seeds, algo=0)
    st.bonds[5, 17] = 16
    st.run(500, lanes_per_chain=-2)
    ref.run(500, lanes_per_chain=1)
    acc = st.acc.cpu().numpy()
    assert acc[5, 6] == 1 and (np.delete(acc[:, 6], 5) == 0).all()
    assert int(st.bonds[5, 17]) == 16 and int(st.bonds[5].sum()) == 16 and (acc[5, :6] == 0).all()
    keep = np.arange(n) != 5
    assert (acc[keep] == ref.acc.cpu().numpy()[keep]).all()
    assert (st.bonds.cpu().numpy()[keep] == ref.bonds.cpu().numpy()[keep]).all()


def test_philox_thresholds_match_oracle(eng):
    from oracle import worm_oracle as wo
    from worm_algorithm_b200 import _lib
    for T in (0.5, 1.0, 1.5, 2.269185314213022, 3.5, 10.0):
        assert _lib.thresholds(T) == wo.thresholds(T)


def test_philox_geometry_independence(eng):
    """Same chains, different launch geometry / chunking / chain offset -> identical results
    (what makes 1/2/4/8-GPU runs reproduce each other)."""
    L, n = 16, 64
    T = np.repeat(np.linspace(1.5, 3.5, 8), 8)
    seeds = np.arange(n, dtype=np.uint64) + 7
    a = eng.PhiloxState(L, T, seeds, algo=0).run(5000, therm_steps=500, lanes_per_chain=1)
    b = eng.PhiloxState(L, T, seeds, algo=0).run(1233, therm_steps=500, lanes_per_chain=4).run(3767, therm_steps=500, lanes_per_chain=-1)
    assert (a.acc == b.acc).all() and (a.bonds == b.bonds).all()
    p = eng.PhiloxState(L, T, seeds, algo=0).run(1501, therm_steps=500, lanes_per_chain=-2).run(3499, therm_steps=500, lanes_per_chain=2)
    assert (a.acc == p.acc).all() and (a.bonds == p.bonds).all() and (a.head == p.head).all() and (a.tail == p.tail).all()
    d = eng.PhiloxState(L, T, seeds, algo=0).run(77, therm_steps=500, lanes_per_chain=32).run(2000, therm_steps=500, lanes_per_chain=8).run(2000, therm_steps=500, lanes_per_chain=2).run(923, therm_steps=500)
    assert (a.acc == d.acc).all() and (a.bonds == d.bonds).all() and (a.head == d.head).all() and (a.tail == d.tail).all()
    # second half of the chains run as their own batch with chain_id0 = 32
    c = eng.PhiloxState(L, T[32:], seeds[32:], algo=0, chain_id0=32).run(5000, therm_steps=500)
    assert (a.acc[32:] == c.acc).all() and (a.bonds[32:] == c.bonds).all()


# ------------------------------------------------------------------------------------------------
# post-processing
# ------------------------------------------------------------------------------------------------
def test_image_block_count_match_reference_golden(eng, post_golden):
    import torch
    g = post_golden
    for L in (4, 8, 16):
        tag = "img_L%d" % L
        dumps = torch.from_numpy(g[tag + "_dumps"]).cuda()
        imgs = eng.image(dumps, L)
        assert (imgs.cpu().numpy() == g[tag + "_images"]).all()
        assert (eng.block(imgs, 0, variant=1).cpu().numpy() == g[tag + "_blockT0"]).all()
        assert (eng.block(imgs, 2, variant=1).cpu().numpy() == g[tag + "_blockT2"]).all()
        cnt = eng.count(imgs).cpu().numpy()
        assert (cnt == (g[tag + "_dumps"] & 1).sum(axis=1)).all()
        assert np.mean(cnt) == g[tag + "_count"][0]
    tags = sorted({k[:-3] for k in g.files if k.startswith("blk_") and k.endswith("_in")})
    for tag in tags:
        img = torch.from_numpy(g[tag + "_in"].astype(np.int32)).cuda()
        Wo = img.shape[0] // 2
        assert (eng.block(img, 0).cpu().numpy().reshape(-1) == g[tag + "_bc0"]).all()
        if tag + "_bc1" in g.files:
            assert (eng.block(img, 1).cpu().numpy().reshape(-1) == g[tag + "_bc1"]).all()
        assert (eng.block(img, 2).cpu().numpy().reshape(-1) == g[tag + "_bc2"]).all()
        cur, s = eng.block(img, 0), 2
        while tag + "_ib_s%d" % s in g.files:
            cur = eng.block(cur, 0)
            assert (cur.cpu().numpy().reshape(-1) == g[tag + "_ib_s%d" % s]).all()
            s += 1
        assert Wo >= 1


@pytest.mark.parametrize("L", [3, 5, 6, 16, 64, 128])
def test_postproc_matches_oracle_random(eng, L):
    import torch
    from oracle import postproc_oracle as po
    rng = np.random.default_rng(L)
    n = 5 if L <= 16 else 2
    bonds = rng.integers(0, 4, (n, 2 * L * L)).astype(np.uint8)
    imgs = eng.image(torch.from_numpy(bonds).cuda(), L)
    ref = np.array([po.image_from_bonds(b, L) for b in bonds])
    assert (imgs.cpu().numpy() == ref).all()
    # blocking on random (non-loop) 0/1/2 images, every variant
    a = rng.integers(0, 2, (n, 2 * L, 2 * L)).astype(np.int32)
    at = torch.from_numpy(a).cuda()
    for v in (0, 1, 2):
        got = eng.block(at, v, variant=0).cpu().numpy()
        assert (got.reshape(n, -1) == np.array([po.block_config(x, 1, v) for x in a])).all()
        got = eng.block(at, v, variant=1).cpu().numpy()
        assert (got == np.array([po.block_config_T(x, v) for x in a])).all()
    cnt = eng.count(at).cpu().numpy()
    assert (cnt == po.bond_counts(a, 2 * L)).all()


@pytest.mark.parametrize("L", [4, 8, 16, 32, 64, 128, 256])
def test_fused_pipeline_equals_the_launch_chain_and_the_oracle(eng, L):
    """worm_pipeline (one launch, works on parity bits) == worm_count(worm_block^k(worm_image(.))) at every level,
    images included, and == oracle/postproc_oracle.py (pinned to block_configs.py:6-71 / iterated_blocking.py:6-59)."""
    import torch
    from oracle import postproc_oracle as po
    rng = np.random.default_rng(1000 + L)
    n = 37 if L <= 64 else 9
    bonds = rng.integers(0, 5, (n, 2 * L * L)).astype(np.uint8)
    bonds[0] = 0
    bonds[1] = 1
    bt = torch.from_numpy(bonds).cuda()
    levels = L.bit_length() - 1
    counts, imgs = eng.pipeline(bt, L, images=range(levels))
    assert counts.shape == (n, levels)
    cur = eng.image(bt, L)
    for k in range(levels):
        assert (imgs[k] == cur).all(), (L, k)
        assert (counts[:, k] == eng.count(cur)).all(), (L, k)
        if k + 1 < levels:
            cur = eng.block(cur, 0)
    assert (eng.pipeline(bt, L) == counts).all()                       # without images
    assert (eng.pipeline(bt, L, levels=1)[:, 0] == torch.from_numpy((bonds & 1).sum(axis=1).astype(np.int64)).cuda()).all()
    if L <= 64:
        c = counts.cpu().numpy()
        for i in (0, 1, 2, n - 1):
            img = po.image_from_bonds(bonds[i], L)
            lv = L
            for k in range(levels):
                assert c[i, k] == po.bond_counts(img.reshape(1, -1), 2 * lv)[0], (L, i, k)
                if k + 1 < levels:
                    img = po.block_config(img.reshape(-1), 1, 0).reshape(lv, lv)
                    lv //= 2


def test_full_size_pipeline_properties(eng):
    """BASELINE config 3 sizes (L = 64, 128): size-independent properties at full size."""
    import torch
    for L in (64, 128):
        T = [2.21, 2.22, 2.23, 2.24, 2.26, 2.27, 2.28, 2.269]
        st = eng.PhiloxState(L, T, np.arange(8, dtype=np.uint64), algo=0, max_dumps=32)
        st.run(400000, therm_steps=40000, dump_every=50)
        nd = st.n_dumps.cpu().numpy()
        assert (nd > 0).all()
        k = int(min(nd.min(), 32))
        dumps = st.dumps[:, :k]
        imgs = eng.image(dumps, L)
        par = (dumps & 1).to(torch.int64)
        # (1) bond pixels == odd bonds; (2) closed configs: every site has even degree
        assert (eng.count(imgs) == par.sum(-1)).all()
        p4 = par.reshape(8, k, L, L, 2)
        deg = p4[..., 0] + p4[..., 1] + torch.roll(p4[..., 0], 1, dims=3) + torch.roll(p4[..., 1], 1, dims=2)
        assert (deg % 2 == 0).all()
        # (3) "1+1=0" blocking keeps closed loops closed at every level down to 2x2 (even degree)
        cur, W = imgs, 2 * L
        while W > 4:
            cur = eng.block(cur, 0)
            W //= 2
            Lc = W // 2
            xb = cur[..., 0::2, 1::2]
            yb = cur[..., 1::2, 0::2]
            deg = xb + yb + torch.roll(xb, 1, dims=-1) + torch.roll(yb, 1, dims=-2)
            assert (deg % 2 == 0).all(), (L, W)
            site = cur[..., 0::2, 0::2]
            assert (site == (deg > 0).to(site.dtype)).all()
            assert cur.shape[-1] == 2 * Lc


# ------------------------------------------------------------------------------------------------
# host-buffer C entry points (the drop-in without torch)
# ------------------------------------------------------------------------------------------------
def test_host_entry_points(eng, ref_runs):
    from worm_algorithm_b200 import _lib
    lib = _lib.lib()
    rec = ref_runs[2]                                      # L=8, T=2.269, 20000 steps, seed 42
    L = rec["L"]
    T = np.array([rec["T"]], dtype=np.float64)
    sd = np.array([rec["seed"]], dtype=np.float64)
    out = np.zeros((1, 8), dtype=np.int64)
    bonds = np.zeros((1, 2 * L * L), dtype=np.uint8)
    _lib.check(lib.worm_sim_host(L, 1, T.ctypes.data_as(_lib._pd), sd.ctypes.data_as(_lib._pd), None,
                                 rec["num_steps"], 0, 0, 0, out.ctypes.data_as(C.c_void_p), bonds.ctypes.data_as(C.c_void_p)))
    assert int(out[0, 2]) == rec["step_num"]
    assert "".join(str(int(v)) for v in bonds[0]) == rec["final_bonds"]
    img = np.zeros((1, 2 * L, 2 * L), dtype=np.int32)
    _lib.check(lib.worm_image_host(L, 1, bonds.ctypes.data_as(C.c_void_p), img.ctypes.data_as(C.c_void_p)))
    from oracle import postproc_oracle as po
    assert (img[0] == po.image_from_bonds(bonds[0], L)).all()
    blk = np.zeros((1, L, L), dtype=np.int32)
    _lib.check(lib.worm_block_host(2 * L, 1, img.ctypes.data_as(C.c_void_p), blk.ctypes.data_as(C.c_void_p), 0, 0))
    assert (blk[0].reshape(-1) == po.block_config(img[0], 1, 0)).all()
    cnt = np.zeros(1, dtype=np.int64)
    _lib.check(lib.worm_count_host(2 * L, 1, img.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p)))
    assert cnt[0] == (bonds[0] & 1).sum()
    pc = np.zeros((1, 3), dtype=np.int64)
    _lib.check(lib.worm_pipeline_host(L, 1, bonds.ctypes.data_as(C.c_void_p), 3, pc.ctypes.data_as(C.c_void_p)))
    lvl = img[0]
    for k in range(3):
        assert pc[0, k] == po.bond_counts(lvl.reshape(1, -1), lvl.shape[0])[0]
        lvl = po.block_config(lvl.reshape(-1), 1, 0).reshape(lvl.shape[0] // 2, lvl.shape[0] // 2)
    # error path: bad argument -> status code + message, no exception from C
    rc = lib.worm_count_host(0, 1, img.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p))
    assert rc == -1 and "W" in _lib.last_error()


# ------------------------------------------------------------------------------------------------
# round / segment edges of the latency kernel: odd starts, runs shorter than a round, odd dump
# periods, warm-up ending inside a round, empty ranges
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("with_hist", [True, False])
@pytest.mark.parametrize("lanes", [2, 4, 8, 32, 1, -2])
def test_philox_segment_edges_match_oracle(eng, lanes, with_hist):
    from oracle import worm_oracle as wo
    L, n = 8, 11
    rng = np.random.default_rng(77 + lanes)
    T = rng.uniform(1.0, 3.5, n)
    seeds = rng.integers(0, 2 ** 63, n, dtype=np.uint64)
    cases = [  # (chunks, therm, dump_every)
        ([1, 1, 1, 2, 3], 0, 0), ([15, 17, 16, 1], 9, 7), ([33, 0, 64, 5], 40, 50), ([7, 250, 3], 123, 13),
        ([128], 128, 1), ([65], 1000, 3),
        # whole dump intervals (the latency kernel walks them in its steady-state loops: short round first, then the full
        # rounds): intervals shorter than / equal to / a multiple of / longer than a round, entered at and off a multiple
        ([700, 301], 50, 50), ([640], 0, 16), ([500, 500], 100, 48), ([333, 667], 10, 10), ([1000], 0, 2),
        ([900, 150], 64, 100), ([49, 51, 400], 0, 50), ([1200], 37, 64),
    ]
    for chunks, therm, dump_every in cases:
        st = eng.PhiloxState(L, T, seeds, algo=0, chain_id0=5, group_index=np.arange(n) % 2, n_groups=2, hist=with_hist,
                             max_dumps=300)
        for ch in chunks:
            st.run(ch, therm_steps=therm, dump_every=dump_every, lanes_per_chain=lanes)
        total = sum(chunks)
        acc, bonds = st.acc.cpu().numpy(), st.bonds.cpu().numpy()
        head, tail, nd = st.head.cpu().numpy(), st.tail.cpu().numpy(), st.n_dumps.cpu().numpy()
        ev, dm = st.events.cpu().numpy(), st.dumps.cpu().numpy()
        hist = np.zeros((2, L * L), dtype=np.int64)
        for c in range(n):
            o = wo.run_philox(L, 0, 5 + c, int(seeds[c]), float(T[c]), 0, total, therm, hist=True,
                              dump_every=dump_every, max_dumps=300)
            key = (lanes, chunks, therm, dump_every, c)
            assert (acc[c, :6] == o["acc"][:6]).all(), (key, acc[c], o["acc"])
            assert (bonds[c] == o["bonds"]).all(), key
            assert head[c] == o["head"] and tail[c] == o["tail"], key
            assert nd[c] == o["n_dumps"], key
            k = min(o["n_dumps"], 300)
            assert (ev[c, :k] == o["events"]).all(), key
            assert (dm[c, :k] == o["dumps"]).all(), key
            hist[c % 2] += o["hist"]
        if with_hist:
            assert (st.hist.cpu().numpy() == hist).all(), (lanes, chunks)


@pytest.mark.parametrize("L,lanes", [(64, 32), (128, 32), (256, 32), (64, 0), (128, 0), (256, 0), (64, 8), (64, 4), (64, 2),
                                     (128, 8), (128, 4), (64, -2), (128, -2), (256, -2)])
def test_philox_large_lattices_match_oracle(eng, L, lanes):
    """L = 64 (one chain per walker in the dual layout; three walkers of 4 chains or one of 8 for big
    batches), L = 128 (compact layout, four walkers of one chain) and L = 256 (compact, one chain per block).
    14 chains: more than a block of any variant, with ragged last blocks."""
    from oracle import worm_oracle as wo
    n = 14
    T = np.array([2.269, 1.2, 3.0, 2.0, 2.5, 2.269, 1.5, 3.5, 2.1, 2.4, 2.269, 1.0, 2.8, 2.2])
    seeds = np.arange(n, dtype=np.uint64) + 3
    gi = np.arange(n) % 3
    st = eng.PhiloxState(L, T, seeds, algo=0, max_dumps=4, group_index=gi, n_groups=3, hist=True).run(
        20000, therm_steps=1000, dump_every=50, lanes_per_chain=lanes)
    acc, bonds = st.acc.cpu().numpy(), st.bonds.cpu().numpy()
    hist = np.zeros((3, L * L), dtype=np.int64)
    for c in range(n):
        o = wo.run_philox(L, 0, c, int(seeds[c]), float(T[c]), 0, 20000, 1000, hist=True, dump_every=50, max_dumps=4)
        assert (acc[c, :6] == o["acc"][:6]).all() and (bonds[c] == o["bonds"]).all()
        assert (st.dumps[c, :min(o["n_dumps"], 4)].cpu().numpy() == o["dumps"]).all()
        hist[gi[c]] += o["hist"]
    assert (st.hist.cpu().numpy() == hist).all()
