"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's PCA path (numpy, float64).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product
(worm_algorithm_b200/) never does.

Follows worm_algorithm/pca.py:97-147 and utils.py:90-104, 123-135:
  * calc_PCA: TruncatedSVD(n_components=1).fit(data) -> explained_variance_ = var(data @ v1) with v1 the
    leading right singular vector of the UN-CENTRED data (sklearn computes np.var of the transformed
    samples, ddof = 0).  sklearn's default solver is a randomised range finder; this restatement uses
    the exact leading eigenpair of data^T data, to which the randomised result converges (pinned
    against the reference class itself in tests/golden/pca.npz, see oracle/make_golden.py gen_pca).
  * block_resampling(data, num_blocks): the KFold(num_blocks) TRAINING sets, i.e. the data with one
    contiguous block deleted; fold sizes n // nb (+1 for the first n % nb folds).
  * jackknife_err(y_i, y_full, nb) = sqrt((nb - 1) / nb * sum((y_i - y_full)^2)).
"""
import numpy as np


def kfold_bounds(n: int, nb: int):
    """[start, end) of the nb contiguous folds of sklearn KFold(n_splits=nb) without shuffling."""
    sizes = np.full(nb, n // nb, dtype=np.int64)
    sizes[: n % nb] += 1
    ends = np.cumsum(sizes)
    return [(int(e - s), int(e)) for s, e in zip(sizes, ends)]


def leading(data: np.ndarray):
    """(explained variance of the leading component, sigma_1^2, component with positive sum)."""
    x = np.asarray(data, dtype=np.float64)
    g = x.T @ x
    w, v = np.linalg.eigh(g)
    v1 = v[:, -1]
    if v1.sum() < 0:
        v1 = -v1
    return float(np.var(x @ v1)), float(w[-1]), v1


def pca_leading(data: np.ndarray, num_blocks: int):
    """explained[1 + nb] (full set first), sigma2[1 + nb], component[d] of the full set, jackknife error."""
    x = np.asarray(data, dtype=np.float64)
    ev, s2, comp = leading(x)
    evs, s2s = [ev], [s2]
    for (a, b) in kfold_bounds(x.shape[0], num_blocks):
        e, s, _ = leading(np.concatenate([x[:a], x[b:]], axis=0))
        evs.append(e)
        s2s.append(s)
    evs = np.array(evs)
    err = float(np.sqrt((num_blocks - 1) * np.sum((evs[1:] - evs[0]) ** 2) / num_blocks))
    return evs, np.array(s2s), comp, err
