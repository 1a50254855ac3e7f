"""Stage timings of the PCA path on the GPU (CUDA events inside the library, worm_pca_timing) and the
reference's CPU algorithm (sklearn TruncatedSVD, pca.py:97-115) on a bounded sample of the same data."""
import argparse
import ctypes as C
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from worm_algorithm_b200 import engine  # noqa: E402


def gpu_case(L, n, nb, reps, peak_tf, peak_gbs):
    d = 4 * L * L
    g = torch.Generator(device="cuda").manual_seed(L)
    x = (torch.rand((n, d), device="cuda", generator=g) < 0.2).to(torch.int32)
    lib = engine.lib()
    lib.worm_pca_timing(1)
    ms = (C.c_double * 4)()
    best = None
    for r in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = engine.pca(x, nb)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        lib.worm_pca_last_timing(ms)
        cur = dict(pack_ms=ms[0], gate_ms=ms[1], gram_ms=ms[2], iter_ms=ms[3], wall_ms=wall * 1e3, iters=res.iters)
        if best is None or cur["gram_ms"] < best["gram_ms"]:
            best = cur
    lib.worm_pca_timing(0)
    flops = float(n) * d * (d + 1)                      # 2 flops x n x d(d+1)/2 distinct entries
    ldg = (d + 3) & ~3
    best.update(L=L, d=d, n=n, num_blocks=nb,
                gram_tflops=flops / (best["gram_ms"] * 1e-3) / 1e12,
                gram_frac=flops / (best["gram_ms"] * 1e-3) / 1e12 / peak_tf,
                gram_out_gbs=ldg * ldg * 4 / (best["gram_ms"] * 1e-3) / 1e9,
                pack_gbs=(n * d * 4 + n * d * 2) / (best["pack_ms"] * 1e-3) / 1e9,
                pack_frac=(n * d * 4 + n * d * 2) / (best["pack_ms"] * 1e-3) / 1e9 / peak_gbs,
                iter_ms_per_iteration=best["iter_ms"] / max(best["iters"], 1),
                configs_per_s=n / (best["wall_ms"] * 1e-3), explained=float(res.explained[0]))
    return best, x


def cpu_case(x, nb, folds):
    from sklearn.decomposition import TruncatedSVD
    data = x.cpu().numpy().astype(np.float64)
    n = data.shape[0]
    t0 = time.perf_counter()
    svd = TruncatedSVD(n_components=1)
    svd.fit(data)
    lam = float(svd.explained_variance_[0])
    sizes = np.full(nb, n // nb)
    sizes[: n % nb] += 1
    ends = np.cumsum(sizes)
    for b in range(folds):
        TruncatedSVD(n_components=1).fit(np.concatenate([data[: ends[b] - sizes[b]], data[ends[b]:]]))
    wall = time.perf_counter() - t0
    est = wall * (1 + nb) / (1 + folds)
    return dict(explained=lam, wall_s=wall, fits=1 + folds, est_full_s=est, configs_per_s=n / est)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="32:20480:10,64:16384:10")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu-folds", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    try:
        peaks = json.load(open("MEASURED_PEAKS.json"))
    except Exception:
        peaks = {}
    peak_tf, peak_gbs = float(peaks.get("bf16_tflops", 1627.0)), float(peaks.get("hbm_gbs", 6536.0))
    for c in a.cases.split(","):
        L, n, nb = (int(v) for v in c.split(":"))
        g, x = gpu_case(L, n, nb, a.reps, peak_tf, peak_gbs)
        out = {"gpu": g}
        if not a.no_cpu and L <= 32:
            out["cpu_sklearn"] = cpu_case(x, nb, a.cpu_folds)
            out["speedup"] = out["cpu_sklearn"]["est_full_s"] / (g["wall_ms"] * 1e-3)
        print(json.dumps(out), flush=True)
        del x
        torch.cuda.empty_cache()
