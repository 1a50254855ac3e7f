"""`Image` of the reference's image_analysis/image.py (boundaries between the black and white regions of a
thresholded grey image) with `get_boundaries` on the GPU (`worm_boundaries`, csrc/postproc.cu).

The reference calls `img.get_boundaries(cutoff)` in a Python double loop per cutoff (image_analysis.ipynb
cell 3); `boundaries_for_cutoffs` does all cutoffs (and `engine.boundaries` all images) in one launch.  The
output uses the bond-image convention of `Bonds` (sites on even-even pixels, links on odd-parity pixels), so
`block_configs`, `iterated_blocking` and `count_bonds` of this package apply to it unchanged.
"""
from __future__ import annotations

import numpy as np

from . import engine


def bytescale(data, cmin=None, cmax=None, high=255, low=0):
    """scipy.misc.bytescale (scipy <= 1.1), which image.py:24 calls and modern scipy no longer ships."""
    data = np.asarray(data)
    if data.dtype == np.uint8:
        return data
    if cmin is None:
        cmin = data.min()
    if cmax is None:
        cmax = data.max()
    cscale = cmax - cmin
    if cscale == 0:
        cscale = 1
    scale = float(high - low) / cscale
    bytedata = (data - cmin) * scale + low
    return (bytedata.clip(low, high) + 0.5).astype(np.uint8)


class Image(object):
    """Same constructor, attributes (`_image`, `_image_cropped`, `_image_bw`, `_image_links`) and methods as
    image.py:12-80."""

    def __init__(self, image, from_file=False, device=None):
        self._device = device
        if from_file:
            try:
                from PIL import Image as _PIL          # scipy.misc.imread(image, "mode=L") of image.py:19
            except ImportError as e:
                raise ImportError("from_file=True needs Pillow to decode the image file") from e
            try:
                self._image_orig = np.asarray(_PIL.open(image).convert("L"))
            except FileNotFoundError:
                raise FileNotFoundError("Unable to load from {}".format(image))
        else:
            self._image_orig = bytescale(image)
        orig = self._image_orig
        if orig.ndim == 1:                                   # flattened square image (image.py:26-29)
            self._Nx = self._Ny = int(np.sqrt(orig.shape[0]))
        else:
            self._Nx, self._Ny = orig.shape
        self._image_flat = orig.reshape(-1)
        self._image = (1 - self._image_flat / 255.0).reshape(self._Nx, self._Ny)
        self._image_cropped = self._image_bw = self._image_links = None

    def crop(self, x_start, y_start, x_end, y_end):
        """image.py:38-44."""
        cropped_image = self._image[x_start:x_end, y_start:y_end]
        self._image_cropped = cropped_image
        return cropped_image

    def _cutoff(self, cutoff, image=None):
        """image.py:46-57: 0.0 where pixel < cutoff, else 1.0."""
        if image is None:
            image = self._image
        image_bw = (np.asarray(image, dtype=np.float64) >= cutoff).astype(np.float64)
        self._image_bw = image_bw
        return image_bw

    def boundaries_for_cutoffs(self, cutoffs, image=None):
        """int array [len(cutoffs), 2Nx, 2Ny]: get_boundaries for every cutoff, one launch."""
        import torch
        if image is None:
            image = self._image
        img = torch.from_numpy(np.ascontiguousarray(np.asarray(image, dtype=np.float64))[None]).to(engine._dev(self._device))
        return engine.boundaries(img, cutoffs)[0].cpu().numpy().astype(int)

    def get_boundaries(self, cutoff, image=None):
        """image.py:59-80."""
        self._cutoff(cutoff, image)
        image_links = self.boundaries_for_cutoffs([cutoff], image)[0]
        self._image_links = image_links
        return image_links


# ---------------------------------------------------------------------------------------------
# The three helper functions of image_analysis.ipynb (cells 2-4), each one batched launch here.
# ---------------------------------------------------------------------------------------------
def get_bdy_images(cutoffs, images, device=None):
    """[n_images, n_cutoffs, 2Nx, 2Ny] boundary images (notebook cell 2: `Image(image).get_boundaries(cutoff)`
    for every image and cutoff)."""
    import torch
    prepared = np.stack([Image(im)._image for im in images])           # per-image bytescale, as Image.__init__ does
    dev = engine._dev(device)
    return engine.boundaries(torch.from_numpy(np.ascontiguousarray(prepared)).to(dev), cutoffs).cpu().numpy().astype(int)


def block_images(images, device=None):
    """One blocking step on every image of an [n_images, n_cutoffs, W, W] set (notebook cell 3;
    image_analysis/utils/block_images.py:4-53 is the "1+1=0" scheme of block_configs.py)."""
    import torch
    a = np.ascontiguousarray(np.asarray(images).astype(np.int32))
    out = engine.block(torch.from_numpy(a).to(engine._dev(device)), 0)
    return out.cpu().numpy().astype(int)


def count_bonds_helper(images, num_blocks=20, device=None):
    """[n_cutoffs, 4] rows (<Nb>, err, <Nb^2> - <Nb>^2, err) over the images of each cutoff (notebook cell 4 with
    image_analysis/utils/count_bonds.py:CountBonds.count_bonds)."""
    import torch
    from .count_bonds import count_stats_with_err
    a = np.asarray(images)
    if a.ndim != 4:
        raise ValueError("Images have the wrong shape.")
    t = torch.from_numpy(np.ascontiguousarray(np.transpose(a, (1, 0, 2, 3)).astype(np.int32))).to(engine._dev(device))
    counts = engine.count(t).cpu().numpy()                             # [n_cutoffs, n_images]
    rows = []
    for bc in counts:
        val, err = count_stats_with_err(bc, num_blocks)
        rows.append([val[0], err[0], val[1], err[1]])
    return np.array(rows)
