"""Throughput of the post-processing kernels against the HBM roofline (development aid)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from worm_algorithm_b200 import engine

peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]


def timeit(f, reps=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


for L, n in ((32, 65536), (64, 16384), (128, 4096)):
    bonds = (torch.rand((n, 2 * L * L), device="cuda") < 0.3).to(torch.uint8)
    img = engine.image(bonds, L)
    t = timeit(lambda: engine.image(bonds, L)); b = n * (2 * L * L + 16 * L * L)
    print("image  L=%3d n=%6d: %.3f ms  %.0f GB/s algorithmic (%.2f of %g)" % (L, n, t * 1e3, b / t / 1e9, b / t / 1e9 / peak, peak))
    t = timeit(lambda: engine.block(img, 0)); b = n * 20 * L * L
    print("block  L=%3d n=%6d: %.3f ms  %.0f GB/s algorithmic (%.2f)" % (L, n, t * 1e3, b / t / 1e9, b / t / 1e9 / peak))
    t = timeit(lambda: engine.block(img, 2, variant=1)); 
    print("blockT L=%3d n=%6d: %.3f ms  %.0f GB/s algorithmic (%.2f)" % (L, n, t * 1e3, b / t / 1e9, b / t / 1e9 / peak))
    t = timeit(lambda: engine.count(img)); b = n * 16 * L * L
    print("count  L=%3d n=%6d: %.3f ms  %.0f GB/s algorithmic (%.2f)" % (L, n, t * 1e3, b / t / 1e9, b / t / 1e9 / peak))

# boundary extraction (image_analysis/image.py:46-80): all images x cutoffs in one launch
for (n, N, m) in ((4096, 64, 8), (16384, 28, 16)):
    imgs = torch.rand((n, N, N), device="cuda", dtype=torch.float64)
    cut = np.linspace(0.05, 0.95, m)
    t = timeit(lambda: engine.boundaries(imgs, cut), reps=5)
    b = n * N * N * 8 + n * m * N * N * 16
    print("bdry   %dx%d n=%6d m=%2d: %.3f ms  %.0f GB/s algorithmic (%.2f)" % (N, N, n, m, t * 1e3, b / t / 1e9, b / t / 1e9 / peak))
