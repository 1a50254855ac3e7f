"""GPU parity of the PCA row: tensor-core Gram matrices (exact), leading component and
jackknife against the numpy oracle and against the reference class's own output (tests/golden/pca.npz,
made by oracle/make_golden.py gen_pca from worm_algorithm/pca.py + sklearn)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a B200"
    from worm_algorithm_b200 import engine
    engine.lib()
    return engine


@pytest.fixture(scope="module")
def gold(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "pca.npz")))


def _cases(gold):
    return sorted(k[:-7] for k in gold if k.endswith("_images"))


def _gram_ref(x, nb):
    from oracle.pca_oracle import kfold_bounds
    x = x.astype(np.int64)
    return np.stack([x[a:b].T @ x[a:b] for (a, b) in kfold_bounds(x.shape[0], nb)])


@pytest.mark.parametrize("d,n,nb,vmax", [(64, 197, 10, 1), (256, 660, 10, 1), (6, 50, 3, 2), (100, 333, 1, 1),
                                         (320, 700, 4, 2), (1024, 3000, 7, 1), (1028, 1100, 2, 1)])
def test_gram_matrices_are_exact(eng, d, n, nb, vmax):
    import torch
    rng = np.random.default_rng(d * 1000 + n)
    x = (rng.random((n, d)) < 0.3).astype(np.int32) * rng.integers(1, vmax + 1, (n, d)).astype(np.int32)
    g = eng.pca_gram(torch.from_numpy(x).cuda(), nb).cpu().numpy()
    ref = _gram_ref(x, nb)
    assert g.shape == ref.shape
    assert np.array_equal(g.astype(np.int64), ref)


def test_gram_on_reference_images(eng, gold):
    import torch
    for tag in _cases(gold):
        x = gold[tag + "_images"].astype(np.int32)
        nb = int(gold[tag + "_ref"][2])
        g = eng.pca_gram(torch.from_numpy(x).cuda(), nb).cpu().numpy()
        assert np.array_equal(g.astype(np.int64), _gram_ref(x, nb))


def test_pca_matches_oracle_and_reference_class(eng, gold):
    import torch
    from oracle import pca_oracle
    for tag in _cases(gold):
        x = gold[tag + "_images"].astype(np.int32)
        lam_ref, err_ref, nb = gold[tag + "_ref"]
        nb = int(nb)
        res = eng.pca(torch.from_numpy(x).cuda(), nb)
        ev = res.explained.cpu().numpy()
        evs, s2, comp, oerr = pca_oracle.pca_leading(x, nb)
        assert res.iters < 2000 and res.residual <= 1e-10
        np.testing.assert_allclose(ev, evs, rtol=1e-9)                       # float64 oracle
        np.testing.assert_allclose(res.sigma2.cpu().numpy(), s2, rtol=1e-12)
        np.testing.assert_allclose(res.component.cpu().numpy(), comp, atol=1e-9)
        # the reference class (sklearn's randomised solver): tolerance = its own approximation error
        assert abs(ev[0] - lam_ref) <= 1e-8 * lam_ref
        err = np.sqrt((nb - 1) * np.sum((ev[1:] - ev[0]) ** 2) / nb)
        assert abs(err - err_ref) <= 1e-6 * err_ref
        np.testing.assert_allclose(res.component.cpu().numpy(), gold[tag + "_vec"], atol=1e-6)


def test_pca_random_images_with_small_gap(eng):
    """Random images have a much smaller spectral gap below the mean direction than worm images."""
    import torch
    from oracle import pca_oracle
    rng = np.random.default_rng(5)
    x = (rng.random((1500, 512)) < 0.5).astype(np.int32)
    x[:, :40] *= (rng.random((1500, 1)) < 0.5).astype(np.int32)              # one strong extra direction
    res = eng.pca(torch.from_numpy(x).cuda(), 5)
    evs, _, comp, _ = pca_oracle.pca_leading(x, 5)
    np.testing.assert_allclose(res.explained.cpu().numpy(), evs, rtol=1e-8)
    np.testing.assert_allclose(res.component.cpu().numpy(), comp, atol=1e-8)


def test_principal_component_class_on_reference_format_files(eng, gold, tmp_path):
    """The host mirror on `{L}_config_{T}.txt` files: same keys, values and saved file as the reference class."""
    from worm_algorithm_b200.pca import PrincipalComponent
    groups = {}
    for tag in _cases(gold):
        L = int(tag.split("_")[1][1:])
        nb = int(gold[tag + "_ref"][2])
        groups.setdefault((L, nb), []).append(tag)
    for (L, nb), tags in groups.items():
        ddir = tmp_path / ("L%d_nb%d" % (L, nb)) / "in"
        sdir = tmp_path / ("L%d_nb%d" % (L, nb)) / "out"
        ddir.mkdir(parents=True)
        for tag in tags:
            key = tag.split("_T")[1]
            T = float(key)
            with open(ddir / ("%d_config_%s.txt" % (L, T)), "w") as f:
                for row in gold[tag + "_images"]:
                    f.write("%s %s\n" % (T, " ".join(str(int(v)) for v in row)))
        pc = PrincipalComponent(L, num_blocks=nb, data_dir=str(ddir) + "/", save_dir=str(sdir) + "/", save=True)
        ref_rows = bytes(gold["pca_L%d_nb%d_file" % (L, nb)]).decode().split("\n")
        ref_rows = [r.split() for r in ref_rows if r.strip()]
        got_rows = [r.split() for r in open(sdir / ("leading_eigenvalue_%d.txt" % L)).read().split("\n") if r.strip()]
        assert [r[0] for r in got_rows] == [r[0] for r in ref_rows]
        for g, r in zip(got_rows, ref_rows):
            assert abs(float(g[1]) - float(r[1])) <= 1e-8 * float(r[1])
            assert abs(float(g[2]) - float(r[2])) <= 1e-6 * float(r[2])
        assert sorted(pc._eig_vals) == sorted(r[0] for r in ref_rows)
        for key in pc._eig_vals:
            assert pc._eig_vecs[key].shape == (1, 4 * L * L)
            assert pc._leading_eig_val[key][0] == pc._eig_vals[key][0][0]
        again = PrincipalComponent(L, num_blocks=nb, data_dir=str(ddir) + "/", save_dir=str(sdir) + "/", load=True)
        assert sorted(again._err) == sorted(str(float(r[0])) for r in ref_rows)


def test_pca_refuses_values_it_cannot_represent_exactly(eng):
    import torch
    from worm_algorithm_b200._lib import WormError
    x = torch.zeros((64, 16), dtype=torch.int32, device="cuda")
    x[3, 5] = 1000
    with pytest.raises(WormError):
        eng.pca(x, 2)
    with pytest.raises(WormError):
        eng.pca(torch.ones((3, 16), dtype=torch.int32, device="cuda"), 10)   # fewer samples than blocks


def test_pca_host_entry_point(eng, gold):
    import ctypes as C
    from oracle import pca_oracle
    tag = _cases(gold)[0]
    x = np.ascontiguousarray(gold[tag + "_images"].astype(np.int32))
    nb = 4
    ev = np.zeros(nb + 1)
    comp = np.zeros(x.shape[1])
    it = C.c_int32(0)
    res = C.c_double(0.0)
    lib = eng.lib()
    rc = lib.worm_pca_host(x.shape[1], x.shape[0], x.ctypes.data_as(C.c_void_p), nb, 2000, 1e-11,
                           ev.ctypes.data_as(C.c_void_p), comp.ctypes.data_as(C.c_void_p), C.byref(it), C.byref(res))
    assert rc == 0
    evs, _, c, _ = pca_oracle.pca_leading(x, nb)
    np.testing.assert_allclose(ev, evs, rtol=1e-9)
    np.testing.assert_allclose(comp, c, atol=1e-9)


def test_pca_edge_cases(eng):
    """One block (no resampling), one configuration per block, constant and all-zero images, d not a multiple of 4."""
    import torch
    from oracle import pca_oracle
    rng = np.random.default_rng(77)
    x = (rng.random((90, 30)) < 0.4).astype(np.int32)
    res = eng.pca(torch.from_numpy(x).cuda(), 1)                      # num_blocks = 1: explained[1] is the empty set -> 0
    ev0, _, comp = pca_oracle.leading(x)
    assert abs(float(res.explained[0]) - ev0) <= 1e-9 * ev0 and float(res.explained[1]) == 0.0
    np.testing.assert_allclose(res.component.cpu().numpy(), comp, atol=1e-9)
    x = (rng.random((12, 9)) < 0.5).astype(np.int32)                  # n == num_blocks: every block is one configuration
    res = eng.pca(torch.from_numpy(x).cuda(), 12)
    np.testing.assert_allclose(res.explained.cpu().numpy(), pca_oracle.pca_leading(x, 12)[0], rtol=1e-8, atol=1e-12)
    zeros = torch.zeros((40, 16), dtype=torch.int32, device="cuda")   # nothing to decompose
    res = eng.pca(zeros, 4)
    assert np.all(res.explained.cpu().numpy() == 0.0) and res.iters <= 4
    ones = torch.ones((40, 16), dtype=torch.int32, device="cuda")     # rank one, zero variance along it
    res = eng.pca(ones, 4)
    np.testing.assert_allclose(res.explained.cpu().numpy(), 0.0, atol=1e-9)
    np.testing.assert_allclose(res.sigma2.cpu().numpy()[0], 40 * 16, rtol=1e-12)
    neg = rng.integers(-3, 4, (200, 21)).astype(np.int32)             # signed values, odd d
    res = eng.pca(torch.from_numpy(neg).cuda(), 5, max_iters=20000, tol=1e-10)
    evs = pca_oracle.pca_leading(neg, 5)[0]
    np.testing.assert_allclose(res.explained.cpu().numpy(), evs, rtol=1e-6)


def test_pca_many_blocks_and_skipped_files(eng, tmp_path):
    """More problems than one pass of the G v kernel holds (1 + 30 > 12), and the reference's rule that a
    file with fewer samples than features is skipped (pca.py:132-133)."""
    import torch
    from oracle import pca_oracle
    from worm_algorithm_b200.pca import PrincipalComponent
    rng = np.random.default_rng(31)
    x = (rng.random((700, 128)) < 0.3).astype(np.int32)
    x[:, :10] |= (rng.random((700, 1)) < 0.5).astype(np.int32)
    res = eng.pca(torch.from_numpy(x).cuda(), 30)
    evs, s2, comp, _ = pca_oracle.pca_leading(x, 30)
    np.testing.assert_allclose(res.explained.cpu().numpy(), evs, rtol=1e-8)
    np.testing.assert_allclose(res.sigma2.cpu().numpy(), s2, rtol=1e-11)
    ddir = tmp_path / "in"
    ddir.mkdir()
    for T, rows in ((2.0, x[:200, :16]), (2.5, x[:10, :16])):            # the second file: 10 samples < 16 features
        with open(ddir / ("2_config_%s.txt" % T), "w") as f:
            for row in rows:
                f.write("%s %s\n" % (T, " ".join(str(int(v)) for v in row)))
    pc = PrincipalComponent(2, num_blocks=5, data_dir=str(ddir) + "/", save_dir=str(tmp_path / "out") + "/", save=False)
    assert list(pc._eig_vals) == ["2."] and pc._temps == [2.0]
    lam = pca_oracle.pca_leading(x[:200, :16], 5)[0][0]
    assert abs(pc._eig_vals["2."][0][0] - lam) <= 1e-9 * lam
