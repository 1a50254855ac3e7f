"""Bring-up aid for the tcgen05 Gram kernel: run a few shapes, compare with an int64 matmul, and print
where mismatches sit (rows / columns / k-blocks) so that a descriptor or swizzle error can be read off."""
import sys
import numpy as np
import torch

sys.path.insert(0, ".")
from worm_algorithm_b200 import engine  # noqa: E402


def one(d, n, nb, seed=0, pattern="rand"):
    rng = np.random.default_rng(seed)
    if pattern == "rand":
        x = (rng.random((n, d)) < 0.3).astype(np.int32)
    else:   # one-hot rows: G[i][j] counts co-occurrence -> diagonal only
        x = np.zeros((n, d), dtype=np.int32)
        x[np.arange(n), np.arange(n) % d] = 1
    try:
        g = engine.pca_gram(torch.from_numpy(x).cuda(), nb).cpu().numpy().astype(np.int64)
    except Exception as e:   # noqa
        print("d=%d n=%d nb=%d %s: EXCEPTION %s" % (d, n, nb, pattern, e))
        return False
    sizes = np.full(nb, n // nb)
    sizes[: n % nb] += 1
    ends = np.cumsum(sizes)
    xs = x.astype(np.int64)
    ref = np.stack([xs[e - s:e].T @ xs[e - s:e] for s, e in zip(sizes, ends)])
    bad = np.argwhere(g != ref)
    print("d=%d n=%d nb=%d %s: mismatches %d / %d" % (d, n, nb, pattern, len(bad), ref.size))
    if len(bad):
        print("   first:", bad[:6].tolist(), "got", [int(g[tuple(b)]) for b in bad[:6]], "want", [int(ref[tuple(b)]) for b in bad[:6]])
        print("   bad rows range", bad[:, 1].min(), bad[:, 1].max(), "cols", bad[:, 2].min(), bad[:, 2].max(),
              "sum got", int(g.sum()), "sum want", int(ref.sum()))
    return len(bad) == 0


if __name__ == "__main__":
    ok = True
    for args in ((128, 64, 1, 0, "onehot"), (128, 64, 1), (128, 256, 1), (256, 64, 1), (256, 300, 2), (512, 1000, 3), (1024, 2000, 5)):
        ok &= one(*args)
    print("ALL OK" if ok else "FAILED")
