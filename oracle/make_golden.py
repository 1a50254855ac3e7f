"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/* from the UNMODIFIED reference.

Run in the build container (needs /root/reference and oracle/_ref/worm_ising_2d):

    python oracle/make_golden.py

  tests/golden/ref_runs.json   outputs of the reference executable (output.txt line, observables
                               rows, bond dumps, bond map, num_bonds) for a set of (L, T, steps, seed)
  tests/golden/post.npz        outputs of the reference's own Python post-processing
                               (block_configs._block_config, iterated_blocking._block_config,
                               Bonds._set_config_data / _block_config_T, CountBonds._count_bonds[_with_err],
                               utils.block_resampling / jackknife_err, SpecificHeat._calc_specific_heat)
                               on seeded random inputs and on real dumps.

The reference modules are imported from /root/reference with the shims of SURVEY App. F
(stub matplotlib; pandas `delim_whitespace`; ragged np.array at bonds.py:185) -- the sources are
not edited or copied.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = "/root/reference/worm_algorithm"
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import ref_runner as rr  # noqa: E402

RUNS = [
    # (L, T, num_steps, seed, bond_flag)       SURVEY App. E cases first
    (4, 2.0, 1000, 7.0, 1),
    (4, 2.0, 1000, 7.9, 1),
    (8, 2.269, 20000, 42.0, 1),
    (16, 2.269, 100000, 2024.0, 1),
    (32, 2.269, 10000000, 12345.0, 0),
    # extra coverage: odd L, L = 2 (degenerate bond numbering), T < 1 (K > 1), high T, bond_flag 0
    (2, 2.5, 500, 3.0, 1),
    (3, 1.7, 3000, 11.0, 1),
    (5, 0.8, 3000, 5.0, 1),
    (6, 3.5, 5000, 99.0, 1),
    (8, 1.0, 20000, 1234.5678, 1),
    (12, 2.269185314213022, 50000, 77.0, 0),
    (32, 1.5, 200000, 1.0, 1),
    (64, 2.27, 300000, 31337.0, 1),
    # BASELINE.json configs[0]: L = 16 at T_c = 2.269, single chain, 1e5 sweeps = 5.12e7 steps
    (16, 2.269, 51200000, 1.0, 0),
]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def gen_runs():
    out = []
    for (L, T, steps, seed, bf) in RUNS:
        r = rr.run_ref(L, T, steps, seed, bf)
        rec = dict(L=L, T=T, num_steps=steps, seed=seed, bond_flag=bf,
                   output_line=r.output_line, step_num=r.step_num,
                   observables_text=r.observables_text if len(r.observables_text) < 20000 else None,
                   observables_sha=hashlib.sha256(r.observables_text.encode()).hexdigest(),
                   n_observables=int(r.observables.shape[0]),
                   num_bonds_text=r.num_bonds_text, Nb_final=r.Nb_final,
                   n_dumps=int(r.dumps.shape[0]),
                   dumps_sha=sha(r.dumps.astype(np.int32)),
                   bonds_file=r.files["bonds_files"],
                   final_bonds="".join(str(int(v)) for v in r.dumps[-1]) if bf and r.dumps[-1].max() < 10 else None,
                   bond_map_sha=sha(r.bond_map.astype(np.int64)))
        if L <= 8 and bf:
            rec["dumps"] = ["".join(str(int(v)) for v in d) for d in r.dumps]
        if L <= 4:
            rec["bond_map"] = r.bond_map.tolist()
        out.append(rec)
        print("ref run", L, T, steps, seed, "->", r.output_line.strip())
    with open(os.path.join(GOLD, "ref_runs.json"), "w") as f:
        json.dump(out, f, indent=1)


STAT_L, STAT_STEPS, STAT_SEEDS = 16, 2_000_000, 24
STAT_T = [2.0, 2.1, 2.2, 2.269, 2.3, 2.4, 2.5, 2.6]


def gen_stats():
    """Statistics of the UNMODIFIED reference executable for the production-mode (Philox) comparison:
    per (T, seed) the three numbers of output.txt that are estimators (Z_av, E_av, step_num;
    worm_ising_2d.cpp:183-191).  24 seeds x 8 temperatures x 2e6 steps of L = 16: ~30 s on one core."""
    E = np.zeros((len(STAT_T), STAT_SEEDS))
    Zav = np.zeros_like(E)
    steps = np.zeros(E.shape, dtype=np.int64)
    for i, T in enumerate(STAT_T):
        for k in range(STAT_SEEDS):
            r = rr.run_ref(STAT_L, T, STAT_STEPS, float(100 + k), 0)
            E[i, k], Zav[i, k], steps[i, k] = r.E_av, r.Z_av, r.step_num
        print("ref stats T=%.3f: E = %.6f +- %.6f" % (T, E[i].mean(), E[i].std(ddof=1) / np.sqrt(STAT_SEEDS)))
    with open(os.path.join(GOLD, "ref_stats.json"), "w") as f:
        json.dump(dict(L=STAT_L, num_steps=STAT_STEPS, seeds=[100 + k for k in range(STAT_SEEDS)], T=STAT_T,
                       E_av=E.tolist(), Z_av=Zav.tolist(), step_num=steps.tolist()), f, indent=1)


def import_reference():
    """Import the reference's Python modules unmodified, with the App. F shims."""
    import pandas as pd
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "matplotlib.cm",
                 "matplotlib.colors", "matplotlib.gridspec", "mpl_toolkits", "mpl_toolkits.axes_grid1",
                 "seaborn"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.setdefault("__path__", [])
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    for attr in ("MaxNLocator", "FormatStrFormatter", "MultipleLocator"):
        setattr(sys.modules["matplotlib.ticker"], attr, object)
    sys.modules["mpl_toolkits.axes_grid1"].make_axes_locatable = lambda *a, **k: None
    _read_csv = pd.read_csv

    def read_csv(*a, **k):
        if k.pop("delim_whitespace", False):
            k["sep"] = r"\s+"
        return _read_csv(*a, **k)
    pd.read_csv = read_csv
    sys.path.insert(0, REF)
    import block_configs as ref_bc          # noqa
    import iterated_blocking as ref_ib      # noqa
    import utils as ref_utils               # noqa
    import count_bonds as ref_cb            # noqa
    import specific_heat as ref_sh          # noqa
    import bonds as ref_bonds               # noqa

    class _NP:                              # bonds.py:185 builds a ragged array (numpy >= 1.24 refuses)
        def __getattr__(self, k):
            return getattr(np, k)

        @staticmethod
        def array(obj, *a, **k):
            try:
                return np.array(obj, *a, **k)
            except ValueError:
                return np.array(obj, dtype=object)
    ref_bonds.np = _NP()
    return ref_bc, ref_ib, ref_utils, ref_cb, ref_sh, ref_bonds


def gen_post():
    ref_bc, ref_ib, ref_utils, ref_cb, ref_sh, ref_bonds = import_reference()
    rng = np.random.default_rng(20261019)
    g = {}
    # ---- blocking on random 0/1 images, several sizes and densities -------------------------
    for L in (2, 4, 6, 8, 16, 32):
        for k, dens in enumerate((0.1, 0.5, 0.9)):
            img = (rng.random((2 * L, 2 * L)) < dens).astype(np.int64)
            tag = "blk_L%d_%d" % (L, k)
            g[tag + "_in"] = img.astype(np.int8)
            g[tag + "_bc0"] = ref_bc._block_config(img.flatten().copy(), 1, 0).astype(np.int8)
            g[tag + "_bc1"] = ref_bc._block_config(img.flatten().copy(), 1, 1).astype(np.int8)
            g[tag + "_bc2"] = ref_bc._block_config(img.flatten().copy(), 1, 2).astype(np.int8)
            g[tag + "_ib"] = ref_ib._block_config(img.flatten().copy(), 1).astype(np.int8)
            assert (g[tag + "_ib"] == g[tag + "_bc0"]).all()
            nlev = int(np.log2(L)) if L & (L - 1) == 0 else 1
            for s in range(2, nlev + 1):
                g[tag + "_ib_s%d" % s] = ref_ib._block_config(img.flatten().copy(), s).astype(np.int8)
                assert (ref_bc._block_config(img.flatten().copy(), s, 0) == g[tag + "_ib_s%d" % s]).all()
    # all-ones / all-zeros edge cases
    for L in (4, 8):
        for name, img in (("ones", np.ones((2 * L, 2 * L), dtype=np.int64)), ("zeros", np.zeros((2 * L, 2 * L), dtype=np.int64))):
            g["blk_%s_L%d_in" % (name, L)] = img.astype(np.int8)
            g["blk_%s_L%d_bc0" % (name, L)] = ref_bc._block_config(img.flatten().copy(), 1, 0).astype(np.int8)
            g["blk_%s_L%d_bc2" % (name, L)] = ref_bc._block_config(img.flatten().copy(), 1, 2).astype(np.int8)

    # ---- Bonds: real dumps -> images through the reference's own class ----------------------
    for (L, T, steps, seed) in ((4, 2.0, 20000, 5.0), (8, 2.269, 60000, 9.0), (16, 2.0, 400000, 21.0)):
        root = tempfile.mkdtemp(prefix="wormgold_")
        r = rr.run_ref(L, T, steps, seed, 1, keep_dir=root)
        cwd = os.getcwd()
        os.chdir(os.path.join(root, "worm_algorithm"))
        try:
            b = ref_bonds.Bonds(L, run=False, T_arr=[T], write=False)
            (tkey,) = list(b._config_data.keys())
            imgs = np.array([b._config_data[tkey][i] for i in range(len(b._config_data[tkey]))])
            blk0 = np.array([b._block_config_T(tkey, i, 0) for i in range(imgs.shape[0])])
            blk2 = np.array([b._block_config_T(tkey, i, 2) for i in range(imgs.shape[0])])
        finally:
            os.chdir(cwd)
        tag = "img_L%d" % L
        keep = slice(0, 12)
        g[tag + "_dumps"] = r.dumps[keep].astype(np.uint8)
        g[tag + "_images"] = imgs[keep].astype(np.int8)
        g[tag + "_blockT0"] = blk0[keep].astype(np.int8)
        g[tag + "_blockT2"] = blk2[keep].astype(np.int8)
        # count on the images with the reference's CountBonds._count_bonds
        cb = ref_cb.CountBonds.__new__(ref_cb.CountBonds)
        cb._width = 2 * L
        flat = imgs[keep].reshape(imgs[keep].shape[0], -1)
        g[tag + "_count"] = np.array(cb._count_bonds(flat), dtype=np.float64)
        nblk = 3
        val, err = cb._count_bonds_with_err(flat, nblk)
        g[tag + "_count_err"] = np.array([val[0], err[0], val[1], err[1], nblk], dtype=np.float64)
        import shutil
        shutil.rmtree(root, ignore_errors=True)

    # ---- jackknife helpers --------------------------------------------------------------------
    x = rng.normal(size=37)
    for nb in (2, 5, 10):
        rs = ref_utils.block_resampling(x, nb)
        g["jk_n37_b%d_means" % nb] = np.array([np.mean(b) for b in rs])
        g["jk_n37_b%d_err" % nb] = np.array(ref_utils.jackknife_err(np.array([np.mean(b) for b in rs]), np.mean(x), nb))
    g["jk_x"] = x
    # ---- specific heat finite difference ------------------------------------------------------
    sh = ref_sh.SpecificHeat.__new__(ref_sh.SpecificHeat)
    T = np.arange(1.0, 3.5, 0.1)
    E = -2.0 * np.tanh(1.0 / T) - 0.3 * np.exp(-(T - 2.269) ** 2 * 8) + 0.01 * rng.normal(size=T.size)
    sh._avg_energy = list(E)
    sh._energy_temps = list(T)
    sh._err = {str(t).rstrip('0'): 0.0 for t in T}
    _, sht, shv, _ = sh._calc_specific_heat()
    g["sh_T"] = T
    g["sh_E"] = E
    g["sh_out_T"] = np.array(sht)
    g["sh_out_C"] = np.array(shv)
    np.savez_compressed(os.path.join(GOLD, "post.npz"), **g)
    print("post.npz:", len(g), "arrays")


def gen_pca():
    """Run the reference's own PrincipalComponent (pca.py, unmodified, sklearn TruncatedSVD) on image files
    in the reference's `{L}_config_{T}.txt` format built from reference-executable dumps, and keep the
    images with its answers.  numpy's global RNG is seeded first: TruncatedSVD's default solver is
    randomised (random_state=None)."""
    import_reference()
    import pca as ref_pca                   # noqa  (resolved from REF on sys.path)
    from oracle import postproc_oracle as po
    from oracle import pca_oracle
    g = {}
    cases = ((4, 2.0, 400000, 3.0, 10), (4, 3.0, 400000, 4.0, 7), (8, 2.269, 1500000, 5.0, 10))
    root = tempfile.mkdtemp(prefix="wormpca_")
    by_L = {}
    for (L, T, steps, seed, nb) in cases:
        r = rr.run_ref(L, T, steps, seed, 1)
        imgs = np.stack([po.image_from_bonds(dmp, L) for dmp in r.dumps]).reshape(len(r.dumps), -1)
        assert imgs.shape[0] >= imgs.shape[1], (L, imgs.shape)
        imgs = imgs[: 3 * imgs.shape[1] + 5]                     # keep the fixture small, length not a multiple of nb
        by_L.setdefault((L, nb), []).append((T, imgs))
    for (L, nb), items in by_L.items():
        ddir = os.path.join(root, "L%d_nb%d" % (L, nb), "in") + "/"
        sdir = os.path.join(root, "L%d_nb%d" % (L, nb), "out") + "/"
        os.makedirs(ddir)
        for (T, imgs) in items:
            with open(ddir + "%d_config_%s.txt" % (L, float(T)), "w") as f:
                for row in imgs:
                    f.write("%s %s\n" % (T, " ".join(str(int(v)) for v in row)))
        np.random.seed(12345)
        pc = ref_pca.PrincipalComponent(L, num_blocks=nb, data_dir=ddir, save_dir=sdir, save=True)
        saved = open(sdir + "leading_eigenvalue_%d.txt" % L).read()
        for (T, imgs) in items:
            key = ("%d_config_%s.txt" % (L, float(T))).split("_")[-1].rstrip(".txt").rstrip("0")
            tag = "pca_L%d_T%s" % (L, key)
            lam = float(pc._eig_vals[key][0][0])
            err = float(pc._err[key])
            vec = np.asarray(pc._eig_vecs[key])[0]
            evs, s2, comp, oerr = pca_oracle.pca_leading(imgs, nb)
            print(tag, imgs.shape, "ref", lam, err, "exact", evs[0], oerr)
            assert abs(lam - evs[0]) < 1e-7 * abs(lam) and abs(err - oerr) < 1e-5 * abs(err) + 1e-12
            g[tag + "_images"] = imgs.astype(np.int8)
            g[tag + "_ref"] = np.array([lam, err, nb], dtype=np.float64)
            g[tag + "_vec"] = vec if vec.sum() > 0 else -vec
        g["pca_L%d_nb%d_file" % (L, nb)] = np.frombuffer(saved.encode(), dtype=np.uint8)
    import shutil
    shutil.rmtree(root, ignore_errors=True)
    np.savez_compressed(os.path.join(GOLD, "pca.npz"), **g)
    print("pca.npz:", len(g), "arrays")


def gen_boundaries():
    """image_analysis/image.py's own Image class (unmodified) on synthetic grey images.  scipy.misc.bytescale /
    imread no longer exist: the module is imported with a stub scipy.misc whose bytescale is the restated
    scipy 1.1 algorithm (oracle/image_oracle.py); _cutoff and get_boundaries are the reference's code."""
    import_reference()
    from oracle import image_oracle as io
    import scipy
    stub = types.ModuleType("scipy.misc")
    stub.bytescale = io.bytescale
    stub.imread = None
    sys.modules["scipy.misc"] = stub
    scipy.misc = stub
    sys.path.insert(0, os.path.join(os.path.dirname(REF.rstrip("/")), "image_analysis"))
    import image as ref_image               # noqa
    rng = np.random.default_rng(4242)
    g = {}
    yy, xx = np.mgrid[0:28, 0:28]
    blob = np.exp(-((xx - 9.0) ** 2 + (yy - 12.0) ** 2) / 30.0) + 0.7 * np.exp(-((xx - 20.0) ** 2 + (yy - 18.0) ** 2) / 12.0)
    cases = {"rand8": rng.random((8, 8)), "rand12x10": rng.random((12, 10)) * 300.0 - 20.0,
             "blob28": blob + 0.05 * rng.random((28, 28)), "bytes6": rng.integers(0, 256, (6, 6)).astype(np.uint8),
             "flat5": np.full((5, 5), 3.0)}
    for name, img in cases.items():
        ri = ref_image.Image(img)
        cutoffs = np.array([0.0, 0.1, 0.35, 0.5, 0.8, 1.0, 1.1])
        g["bdy_%s_in" % name] = img
        g["bdy_%s_prepared" % name] = np.asarray(ri._image, dtype=np.float64)
        g["bdy_%s_cutoffs" % name] = cutoffs
        outs, bws = [], []
        for c in cutoffs:
            outs.append(np.asarray(ri.get_boundaries(c)).astype(np.int8))
            bws.append(np.asarray(ri._image_bw).astype(np.int8))
            assert (io.get_boundaries(io.prepare(img), c) == outs[-1]).all(), (name, c)
        g["bdy_%s_links" % name] = np.stack(outs)
        g["bdy_%s_bw" % name] = np.stack(bws)
    # ---- the notebook's chain on a small image set: boundaries -> block_image -> CountBonds per cutoff ----
    # (image_analysis/utils is a package named `utils`, like worm_algorithm/utils.py: swap the module in and out)
    stub.toimage = None
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "utils" or k.startswith("utils.")}
    saved_path = list(sys.path)
    sys.path[:] = [q for q in sys.path if os.path.abspath(q) != os.path.abspath(REF)]   # a namespace package loses to utils.py
    try:
        from utils.block_images import block_image          # noqa
        from utils.count_bonds import CountBonds as IACountBonds   # noqa
        imgs = rng.random((7, 8, 8))
        cutoffs = np.array([0.3, 0.5, 0.7])
        bdy = np.array([[np.asarray(ref_image.Image(im).get_boundaries(c)) for c in cutoffs] for im in imgs])   # get_bdy_images
        blk = np.array([[block_image(b).reshape(8, 8) for b in row] for row in bdy])                            # block_images
        per_cut = np.transpose(bdy, (1, 0, 2, 3))
        per_cut_blk = np.transpose(blk, (1, 0, 2, 3))
        g["chain_in"] = imgs
        g["chain_cutoffs"] = cutoffs
        g["chain_bdy"] = bdy.astype(np.int8)
        g["chain_blocked"] = blk.astype(np.int8)
        g["chain_stats"] = np.array([IACountBonds(s_, 3).count_bonds() for s_ in per_cut], dtype=np.float64)
        g["chain_stats_blocked"] = np.array([IACountBonds(s_, 3).count_bonds() for s_ in per_cut_blk], dtype=np.float64)
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    np.savez_compressed(os.path.join(GOLD, "boundaries.npz"), **g)
    print("boundaries.npz:", len(g), "arrays")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "runs"):
        gen_runs()
    if what in ("all", "stats"):
        gen_stats()
    if what in ("all", "post"):
        gen_post()
    if what in ("all", "pca"):
        gen_pca()
    if what in ("all", "boundaries"):
        gen_boundaries()
