"""Torch-tensor front end of the C ABI: allocation, streams and result tensors only.

PyTorch is plumbing here (device memory, the current CUDA stream, `torch.distributed`); all
compute is in libworm_b200.so.  Every function raises if the library or a B200 is missing.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import ACC_STRIDE, ALGO_BINARY, ALGO_TAYLOR, MT_STRIDE, check, lib  # noqa: F401


def _dev(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.WormError(-3, "no CUDA device visible; worm_algorithm_b200 has no CPU fallback")
    d = torch.device("cuda" if device is None else device)
    if d.type != "cuda":
        raise _lib.WormError(-3, "device %s: worm_algorithm_b200 runs on a B200 (cuda) only" % d)
    return d


def _stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: torch.Tensor | None) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint64).reshape(-1))


# ---------------------------------------------------------------------------------------------
@dataclass
class MtResult:
    """Per-chain results of the deterministic mode (device tensors)."""
    L: int
    T: np.ndarray
    out: torch.Tensor          # int64 [n, 8]   WORM_MT_* record
    bonds: torch.Tensor        # uint8 [n, 2N]  final occupations
    n_events: torch.Tensor     # int32 [n]
    events: torch.Tensor       # int64 [n, max_events, 3]  {step, Z, Nb_tot}
    dumps: torch.Tensor | None  # uint8 [n, max_events, 2N]
    n_trace: torch.Tensor | None
    trace: torch.Tensor | None  # int64 [n, max_trace, 2]


def run_mt(L: int, T, seeds, num_steps: int, bond_flag: int = 0, max_events: int = 0,
           max_trace: int = 0, device=None) -> MtResult:
    """n independent copies of the reference loop on the reference's MT19937 stream
    (worm_ising_2d.cpp:111-157).  T, seeds: per-chain arrays (seeds are doubles, like input.txt)."""
    dev = _dev(device)
    T = _f64(T)
    seeds = _f64(seeds)
    n = T.size
    if seeds.size != n:
        raise ValueError("T and seeds must have the same length")
    nb2 = 2 * L * L
    with torch.cuda.device(dev):
        out = torch.zeros((n, MT_STRIDE), dtype=torch.int64, device=dev)
        bonds = torch.zeros((n, nb2), dtype=torch.uint8, device=dev)
        me = max(int(max_events), 0)
        n_events = torch.zeros((n,), dtype=torch.int32, device=dev)
        events = torch.zeros((n, max(me, 1), 3), dtype=torch.int64, device=dev)
        dumps = torch.zeros((n, max(me, 1), nb2), dtype=torch.uint8, device=dev) if (bond_flag and me) else None
        mt = max(int(max_trace), 0)
        n_trace = torch.zeros((n,), dtype=torch.int32, device=dev) if mt else None
        trace = torch.zeros((n, mt, 2), dtype=torch.int64, device=dev) if mt else None
        wsb = lib().worm_mt_workspace_bytes(n)
        ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
        check(lib().worm_mt_run(L, n, T.ctypes.data_as(_lib._pd), seeds.ctypes.data_as(_lib._pd),
                                int(num_steps), int(bond_flag), _ptr(bonds), _ptr(out),
                                me, _ptr(n_events), _ptr(events), _ptr(dumps),
                                mt, _ptr(n_trace), _ptr(trace), _ptr(ws), wsb, _stream_ptr(dev)))
        # the workspace must stay alive until the kernel has consumed it
        torch.cuda.current_stream(dev).synchronize()
    return MtResult(L, T, out, bonds, n_events, events[:, :me], dumps, n_trace, trace)


# ---------------------------------------------------------------------------------------------
class PhiloxState:
    """Resumable state + accumulators of a batch of production chains (all device tensors).

    group_index (one int per chain, e.g. its temperature index) switches on the per-group sums:
    `group_acc` int64 [n_groups, 8] (the WORM_ACC_* record summed over the group's chains by the
    kernel itself) and, with hist=True, the G(r) histograms `hist` int64 [n_groups, N]."""

    launches_per_run = 1          # kernels launched by one run() call

    def __init__(self, L: int, T, seeds, algo: int = ALGO_TAYLOR, chain_id0: int = 0,
                 group_index=None, n_groups: int = 0, hist: bool = False,
                 max_dumps: int = 0, store_dumps: bool = True, device=None):
        self.device = _dev(device)
        self.L = int(L)
        self.N = self.L * self.L
        self.algo = int(algo)
        self.T = _f64(T)
        self.seeds = _u64(seeds)
        self.n = self.T.size
        if self.seeds.size != self.n:
            raise ValueError("T and seeds must have the same length")
        self.chain_id0 = int(chain_id0)
        self.step = 0
        dev = self.device
        with torch.cuda.device(dev):
            self.bonds = torch.zeros((self.n, 2 * self.N), dtype=torch.uint8, device=dev)
            self.head = torch.zeros((self.n,), dtype=torch.int32, device=dev)
            self.tail = torch.zeros((self.n,), dtype=torch.int32, device=dev)
            self.acc = torch.zeros((self.n, ACC_STRIDE), dtype=torch.int64, device=dev)
            self.group_index = None
            self.group_acc = None
            self.hist = None
            self.n_groups = 0
            if group_index is not None:
                gi = np.ascontiguousarray(np.asarray(group_index, dtype=np.int32).reshape(-1))
                if gi.size != self.n:
                    raise ValueError("group_index must have one entry per chain")
                self.n_groups = int(n_groups) if n_groups else int(gi.max()) + 1
                if gi.min() < 0 or gi.max() >= self.n_groups:
                    raise ValueError("group_index out of range")
                self.group_index = torch.from_numpy(gi).to(dev)
                self.group_acc = torch.zeros((self.n_groups, ACC_STRIDE), dtype=torch.int64, device=dev)
                if hist:
                    self.hist = torch.zeros((self.n_groups, self.N), dtype=torch.int64, device=dev)
            elif hist:
                raise ValueError("hist=True needs group_index")
            self.max_dumps = int(max_dumps)
            self.n_dumps = torch.zeros((self.n,), dtype=torch.int32, device=dev)
            self.events = torch.zeros((self.n, max(self.max_dumps, 1), 3), dtype=torch.int64, device=dev)
            self.dumps = (torch.zeros((self.n, self.max_dumps, 2 * self.N), dtype=torch.uint8, device=dev)
                          if (self.max_dumps and store_dumps) else None)
            self._wsb = lib().worm_philox_workspace_bytes(self.n)
            self._ws = torch.empty((self._wsb,), dtype=torch.uint8, device=dev)

    def reset(self):
        """Back to step 0: empty lattices, head = tail = 0, zero sums (stream-ordered)."""
        for t in (self.bonds, self.head, self.tail, self.acc, self.group_acc, self.hist, self.n_dumps):
            if t is not None:
                t.zero_()
        self.step = 0
        return self

    def run(self, n_steps: int, therm_steps: int = 0, dump_every: int = 0, lanes_per_chain: int = 0):
        """Advance every chain by n_steps (stream-ordered, no host synchronisation)."""
        dev = self.device
        with torch.cuda.device(dev):
            check(lib().worm_philox_run(
                self.L, self.algo, self.n, self.chain_id0,
                self.T.ctypes.data_as(_lib._pd), self.seeds.ctypes.data_as(_lib._pu64),
                self.step, int(n_steps), int(therm_steps),
                _ptr(self.bonds), _ptr(self.head), _ptr(self.tail), _ptr(self.acc),
                _ptr(self.group_index), self.n_groups, _ptr(self.hist), _ptr(self.group_acc),
                int(dump_every) if self.max_dumps else 0, self.max_dumps, _ptr(self.n_dumps),
                _ptr(self.events), _ptr(self.dumps),
                int(lanes_per_chain), _ptr(self._ws), self._wsb, _stream_ptr(dev)))
        self.step += int(n_steps)
        return self


# ---------------------------------------------------------------------------------------------
def image(bonds: torch.Tensor, L: int) -> torch.Tensor:
    """uint8 occupations [..., 2N] -> int32 images [..., 2L, 2L]   (bonds.py:199-319)."""
    dev = _dev(bonds.device)
    b = bonds.contiguous()
    if b.dtype != torch.uint8:
        raise TypeError("bonds must be uint8")
    lead = b.shape[:-1]
    n = int(np.prod(lead)) if len(lead) else 1
    if b.shape[-1] != 2 * L * L:
        raise ValueError("last dimension must be 2*L*L")
    with torch.cuda.device(dev):
        out = torch.empty((n, 2 * L, 2 * L), dtype=torch.int32, device=dev)
        check(lib().worm_image(L, n, _ptr(b), _ptr(out), _stream_ptr(dev)))
    return out.reshape(*lead, 2 * L, 2 * L)


def block(images: torch.Tensor, double_bonds_value: int = 0, variant: int = 0) -> torch.Tensor:
    """One RG blocking step int32 [..., W, W] -> [..., W/2, W/2]   (block_configs.py:6-71)."""
    dev = _dev(images.device)
    a = images.contiguous()
    if a.dtype != torch.int32:
        raise TypeError("images must be int32")
    W = a.shape[-1]
    if a.shape[-2] != W:
        raise ValueError("images must be square")
    lead = a.shape[:-2]
    n = int(np.prod(lead)) if len(lead) else 1
    with torch.cuda.device(dev):
        out = torch.empty((n, W // 2, W // 2), dtype=torch.int32, device=dev)
        check(lib().worm_block(W, n, _ptr(a), _ptr(out), int(double_bonds_value), int(variant), _stream_ptr(dev)))
    return out.reshape(*lead, W // 2, W // 2)


def count(images: torch.Tensor) -> torch.Tensor:
    """n = sum of pixels with (i + j) odd, int32 [..., W, W] -> int64 [...]   (count_bonds.py:110-117)."""
    dev = _dev(images.device)
    a = images.contiguous()
    if a.dtype != torch.int32:
        raise TypeError("images must be int32")
    W = a.shape[-1]
    lead = a.shape[:-2]
    n = int(np.prod(lead)) if len(lead) else 1
    with torch.cuda.device(dev):
        out = torch.empty((n,), dtype=torch.int64, device=dev)
        check(lib().worm_count(W, n, _ptr(a), _ptr(out), _stream_ptr(dev)))
    return out.reshape(lead) if len(lead) else out.reshape(())


def pipeline(bonds: torch.Tensor, L: int, levels: int | None = None, images=()):
    """The C3 chain in one launch: uint8 occupations [..., 2N] -> int64 bond counts [..., levels] of the dump's image
    and of its 1+1=0 blockings (bonds.py:199-319 -> iterated_blocking.py:6-59 -> count_bonds.py:107-123).
    `images`: levels whose int32 images [..., 2(L>>k), 2(L>>k)] are also wanted -> (counts, {level: images})."""
    dev = _dev(bonds.device)
    b = bonds.contiguous()
    if b.dtype != torch.uint8:
        raise TypeError("bonds must be uint8")
    if b.shape[-1] != 2 * L * L:
        raise ValueError("last dimension must be 2*L*L")
    lead = b.shape[:-1]
    n = int(np.prod(lead)) if len(lead) else 1
    if levels is None:
        levels = max(int(L).bit_length() - 1, 1)            # down to the 2 x 2 lattice
    want = sorted(set(int(k) for k in images))
    with torch.cuda.device(dev):
        counts = torch.empty((n, levels), dtype=torch.int64, device=dev)
        imgs = {k: torch.empty((n, 2 * (L >> k), 2 * (L >> k)), dtype=torch.int32, device=dev) for k in want}
        ptrs = (C.c_void_p * levels)(*[imgs[k].data_ptr() if k in imgs else None for k in range(levels)])
        check(lib().worm_pipeline(L, n, _ptr(b), levels, _ptr(counts), ptrs if want else None, _stream_ptr(dev)))
    counts = counts.reshape(*lead, levels)
    if want:
        return counts, {k: v.reshape(*lead, 2 * (L >> k), 2 * (L >> k)) for k, v in imgs.items()}
    return counts


def boundaries(images: torch.Tensor, cutoffs) -> torch.Tensor:
    """Black / white boundaries of thresholded grey images (image_analysis/image.py:46-80):
    f64 [n, Nx, Ny] x cutoffs [m] -> int32 [n, m, 2Nx, 2Ny]."""
    dev = _dev(images.device)
    a = images.contiguous()
    if a.dtype != torch.float64 or a.dim() != 3:
        raise TypeError("images must be float64 [n, Nx, Ny]")
    n, Nx, Ny = a.shape
    with torch.cuda.device(dev):
        cut = torch.as_tensor(np.asarray(cutoffs, dtype=np.float64).reshape(-1)).to(dev)
        m = cut.numel()
        out = torch.empty((n, m, 2 * Nx, 2 * Ny), dtype=torch.int32, device=dev)
        check(lib().worm_boundaries(Nx, Ny, n, _ptr(a), m, _ptr(cut), _ptr(out), _stream_ptr(dev)))
        torch.cuda.current_stream(dev).synchronize()     # `cut` must outlive the kernel
    return out


# ---------------------------------------------------------------------------------------------
@dataclass
class PcaResult:
    """Leading principal component of an image stack (device tensors) -- pca.py:97-147."""
    explained: torch.Tensor     # f64 [1 + num_blocks]: full set, then each delete-one-block set
    sigma2: torch.Tensor        # f64 [1 + num_blocks]: leading eigenvalue of X^T X
    component: torch.Tensor     # f64 [d]: components_[0] of the full set
    iters: int
    residual: float


def _flat_images(images: torch.Tensor):
    a = images.contiguous()
    if a.dtype != torch.int32:
        raise TypeError("images must be int32")
    if a.dim() < 2:
        raise ValueError("images must be [n, d] or [n, W, W]")
    n = a.shape[0]
    return a.reshape(n, -1), n, int(np.prod(a.shape[1:]))


def pca(images: torch.Tensor, num_blocks: int = 10, max_iters: int = 2000, tol: float = 1e-11) -> PcaResult:
    """explained_variance_[0] of TruncatedSVD(1) on the un-centred [n, d] matrix, for the full set and
    for the `num_blocks` KFold training sets (block_resampling, utils.py:90-104)."""
    dev = _dev(images.device)
    a, n, d = _flat_images(images)
    nb = int(num_blocks)
    with torch.cuda.device(dev):
        wsb = lib().worm_pca_workspace_bytes(d, n, nb, 0)
        ws = torch.empty((max(wsb, 1),), dtype=torch.uint8, device=dev)
        ev = torch.empty((nb + 1,), dtype=torch.float64, device=dev)
        s2 = torch.empty((nb + 1,), dtype=torch.float64, device=dev)
        comp = torch.empty((d,), dtype=torch.float64, device=dev)
        it, res = C.c_int32(0), C.c_double(0.0)
        check(lib().worm_pca(d, n, _ptr(a), nb, int(max_iters), float(tol), _ptr(ev), _ptr(s2), _ptr(comp),
                             C.byref(it), C.byref(res), _ptr(ws), wsb, _stream_ptr(dev)))
    return PcaResult(ev, s2, comp, it.value, res.value)


def pca_gram(images: torch.Tensor, num_blocks: int = 1) -> torch.Tensor:
    """The tensor-core stage of `pca` alone: int32 [num_blocks, d, d], G_b = X_b^T X_b."""
    dev = _dev(images.device)
    a, n, d = _flat_images(images)
    nb = int(num_blocks)
    ldg = (d + 3) & ~3
    with torch.cuda.device(dev):
        wsb = lib().worm_pca_workspace_bytes(d, n, nb, 1)
        ws = torch.empty((max(wsb, 1),), dtype=torch.uint8, device=dev)
        g = torch.empty((nb, ldg, ldg), dtype=torch.int32, device=dev)
        check(lib().worm_pca_gram(d, n, _ptr(a), nb, _ptr(g), _ptr(ws), wsb, _stream_ptr(dev)))
        torch.cuda.current_stream(dev).synchronize()     # the workspace must outlive the kernels
    return g[:, :d, :d]
