"""CPU restatement of the reference's scale initialiser.  TEST INFRASTRUCTURE ONLY (tier rule 3).

Restates SimpleKNN::knn / boxMeanDist / updateKBest (submodules/simple-knn/simple_knn.cu:139-221) by
definition: for every point the three smallest squared distances to the OTHER points (index != own index, so
coincident points count with distance 0), averaged as (d0 + d1 + d2) / 3 in float32; fewer than three other
points leave FLT_MAX entries, as the reference's initial `best` does.  Brute force, O(P^2): small P only.
Pinned against the real reference extension (tools/make_golden_knn.py -> tests/golden/knn_*.npz, produced on
a B200 from oracle/_ref/simple_knn).
"""
from __future__ import annotations

import numpy as np

FLT_MAX = np.float32(3.4028234663852886e38)


def mean_dist2(points: np.ndarray) -> np.ndarray:
    p = np.asarray(points, dtype=np.float32)
    P = p.shape[0]
    out = np.empty(P, np.float32)
    for i in range(P):
        d = p - p[i]
        # dx*dx + dy*dy + dz*dz in float32 (the kernel contracts it to FMAs: last-bit differences only)
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]).astype(np.float32)
        d2[i] = np.inf
        best = np.full(3, FLT_MAX, np.float32)
        k = min(3, P - 1)
        if k > 0:
            best[:k] = np.sort(np.partition(d2, k - 1)[:k])
        with np.errstate(over="ignore"):
            out[i] = (best[0] + best[1] + best[2]) / np.float32(3.0)
    return out
