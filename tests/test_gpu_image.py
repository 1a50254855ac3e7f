"""GPU parity of the boundary-extraction row (image_analysis/image.py:46-80)."""
import ctypes as C
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
    return dict(np.load(os.path.join(golden_dir, "boundaries.npz")))


def _names(gold):
    return sorted(k[4:-3] for k in gold if k.startswith("bdy_") and k.endswith("_in"))


def test_boundaries_match_reference_class_golden(eng, gold):
    import torch
    for name in _names(gold):
        prepared, cutoffs = gold["bdy_%s_prepared" % name], gold["bdy_%s_cutoffs" % name]
        out = eng.boundaries(torch.from_numpy(prepared[None]).cuda(), cutoffs).cpu().numpy()[0]
        assert np.array_equal(out, gold["bdy_%s_links" % name].astype(np.int32))


def test_image_class_matches_reference_class_golden(eng, gold):
    from worm_algorithm_b200.image import Image
    for name in _names(gold):
        img = Image(gold["bdy_%s_in" % name])
        assert np.array_equal(img._image, gold["bdy_%s_prepared" % name])
        for k, c in enumerate(gold["bdy_%s_cutoffs" % name]):
            links = img.get_boundaries(c)
            assert np.array_equal(links, gold["bdy_%s_links" % name][k])
            assert np.array_equal(img._image_bw, gold["bdy_%s_bw" % name][k])
            assert np.array_equal(img._image_links, links)
        assert img.crop(1, 1, 4, 3).shape == (3, 2)


def test_boundaries_batch_matches_oracle(eng):
    import torch
    from oracle import image_oracle as io
    rng = np.random.default_rng(9)
    for (n, Nx, Ny, m) in ((5, 28, 28, 9), (3, 7, 13, 4), (2, 1, 6, 3), (2, 64, 48, 5)):
        imgs = rng.random((n, Nx, Ny))
        cutoffs = np.sort(rng.random(m))
        out = eng.boundaries(torch.from_numpy(imgs).cuda(), cutoffs).cpu().numpy()
        for i in range(n):
            for k in range(m):
                assert np.array_equal(out[i, k], io.get_boundaries(imgs[i], cutoffs[k]))


def test_boundaries_feed_blocking_and_counting(eng):
    """The notebook's chain: boundaries -> iterated blocking -> bond counts, all on the device, against the oracle."""
    import torch
    from oracle import image_oracle as io
    from oracle import postproc_oracle as po
    rng = np.random.default_rng(10)
    imgs = rng.random((4, 32, 32))
    cutoffs = np.array([0.2, 0.5, 0.8])
    links = eng.boundaries(torch.from_numpy(imgs).cuda(), cutoffs)            # [4, 3, 64, 64]
    blocked = eng.block(links.reshape(-1, 64, 64), 0)
    counts = eng.count(blocked).cpu().numpy()
    k = 0
    for i in range(4):
        for c in cutoffs:
            ref = io.get_boundaries(imgs[i], c)
            b = po.block_config(ref.flatten(), 1, 0).reshape(32, 32)
            assert np.array_equal(blocked[k].cpu().numpy(), b)
            assert counts[k] == po.bond_counts(b.reshape(1, -1), 32)[0]
            k += 1


def test_boundaries_host_entry_point(eng, gold):
    name = "blob28"
    prepared = np.ascontiguousarray(gold["bdy_%s_prepared" % name])
    cutoffs = np.ascontiguousarray(gold["bdy_%s_cutoffs" % name])
    out = np.zeros((cutoffs.size, 56, 56), dtype=np.int32)
    rc = eng.lib().worm_boundaries_host(28, 28, 1, prepared.ctypes.data_as(C.c_void_p), cutoffs.size,
                                        cutoffs.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    assert np.array_equal(out, gold["bdy_%s_links" % name].astype(np.int32))


def test_notebook_chain_matches_reference_helpers(eng, gold):
    """get_bdy_images -> block_images -> count_bonds_helper of image_analysis.ipynb against the reference's own
    Image / block_image / CountBonds on the same images (oracle/make_golden.py gen_boundaries)."""
    from worm_algorithm_b200 import image as im
    imgs, cutoffs = gold["chain_in"], gold["chain_cutoffs"]
    bdy = im.get_bdy_images(cutoffs, imgs)
    assert np.array_equal(bdy, gold["chain_bdy"])
    blk = im.block_images(bdy)
    assert np.array_equal(blk, gold["chain_blocked"])
    np.testing.assert_allclose(im.count_bonds_helper(bdy, 3), gold["chain_stats"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(im.count_bonds_helper(blk, 3), gold["chain_stats_blocked"], rtol=1e-12, atol=1e-12)
