"""TEST INFRASTRUCTURE ONLY -- CPU restatement of image_analysis/image.py (numpy, float64).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.

  * bytescale: scipy.misc.bytescale, a third-party function the reference calls at image.py:24 and that
    no longer exists (removed in scipy 1.2; the reference's requirements.txt does not pin scipy).  Its
    published algorithm (scipy 1.1.0, scipy/misc/pilutil.py) is restated here: uint8 input is returned
    unchanged; otherwise (data - cmin) * (high - low) / (cmax - cmin or 1) + low, clipped, + 0.5, truncated.
  * prepare: Image.__init__ (image.py:13-35): _image = 1 - bytescale(image) / 255.
  * cutoff_bw: Image._cutoff (image.py:46-57): 0 where pixel < cutoff else 1.
  * get_boundaries: Image.get_boundaries (image.py:59-80), vectorised with np.roll; pinned against the
    reference's own loops in tests/golden/boundaries.npz (oracle/make_golden.py gen_boundaries).
"""
import numpy as np


def bytescale(data, cmin=None, cmax=None, high=255, low=0):
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


def prepare(image) -> np.ndarray:
    orig = bytescale(np.asarray(image))
    if orig.ndim == 1:
        n = int(np.sqrt(orig.shape[0]))
        orig = orig.reshape(n, n)
    return 1 - orig / 255.0


def cutoff_bw(image, cutoff) -> np.ndarray:
    return (np.asarray(image, dtype=np.float64) >= cutoff).astype(np.int64)


def get_boundaries(image, cutoff) -> np.ndarray:
    bw = cutoff_bw(image, cutoff)
    Nx, Ny = bw.shape
    link0 = bw ^ np.roll(bw, 1, axis=1)          # links_arr[x, y, 0] = bw[x, y-1] ^ bw[x, y]     (image.py:69-70)
    link1 = bw ^ np.roll(bw, 1, axis=0)          # links_arr[x, y, 1] = bw[x-1, y] ^ bw[x, y]     (image.py:67-68)
    out = np.zeros((2 * Nx, 2 * Ny), dtype=np.int64)
    out[1::2, 0::2] = link0                      # image.py:73
    out[0::2, 1::2] = link1                      # image.py:74
    site = link0 | link1 | np.roll(link0, 1, axis=0) | np.roll(link1, 1, axis=1)   # image.py:75-78
    out[0::2, 0::2] = site
    return out
