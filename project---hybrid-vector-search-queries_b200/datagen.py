"""Seedable contest-shaped input generator.

Same binary layouts and value ranges as the reference's src/write_data.c / src/write_query.c
(D row = [C, T, x0..x99], Q row = [type, v, l, r, q0..q99]; T,l ~ U(-3,3), r ~ U(l,4),
vectors ~ U(-6,6)) but reproducible (numpy PCG64) and with INTEGER categories (optionally Zipf-skewed, and
optionally clustered vectors with queries drawn near the data), because the
reference generators draw C and v continuously in (-1,1) so `C == (int)v` never matches
(BASELINE.md section 1 caveat) and every type-1/3 query degenerates to the pad path.
"""
from __future__ import annotations

import numpy as np

DROW, QROW = 102, 104


def _zipf_cdf(ncat: int, s: float) -> np.ndarray:
    w = 1.0 / np.arange(1, ncat + 1, dtype=np.float64) ** s
    return np.cumsum(w / w.sum())


def gen_data(n: int, seed: int, ncat: int = 100, chunk: int = 1 << 20, zipf: float = 0.0, clusters: int = 0,
             cluster_sigma: float = 0.5) -> np.ndarray:
    """zipf > 0: category c drawn with probability ~ 1/(c+1)^zipf (a few huge categories, many tiny ones) instead of
    uniformly.  clusters > 0: vectors are `clusters` centres ~ U(-5,5)^100 plus N(0, cluster_sigma^2) noise instead of
    U(-6,6)^100 -- distances to a query then concentrate in bands, with many near-ties inside a cluster (SURVEY 8f-3:
    uniform data makes thresholds unrealistically easy)."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, DROW), np.float32)
    cdf = _zipf_cdf(ncat, zipf) if zipf > 0 else None
    centres = (rng.random((clusters, 100), dtype=np.float32) * np.float32(10.0) - np.float32(5.0)) if clusters > 0 else None
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        u = rng.random((e - s, DROW), dtype=np.float32)
        if cdf is None:
            out[s:e, 0] = np.floor(u[:, 0] * np.float32(ncat))
        else:
            out[s:e, 0] = np.minimum(np.searchsorted(cdf, u[:, 0].astype(np.float64)), ncat - 1).astype(np.float32)
        out[s:e, 1] = u[:, 1] * np.float32(6.0) - np.float32(3.0)
        if centres is None:
            out[s:e, 2:] = u[:, 2:] * np.float32(12.0) - np.float32(6.0)
        else:
            which = rng.integers(0, clusters, e - s)
            out[s:e, 2:] = centres[which] + rng.standard_normal((e - s, 100), dtype=np.float32) * np.float32(cluster_sigma)
    return out


def gen_queries(m: int, seed: int, ncat: int = 100, types=(0, 1, 2, 3), range_width: float | None = None,
                near: np.ndarray | None = None, near_sigma: float = 0.5) -> np.ndarray:
    """types: the query types to draw uniformly from.  range_width: if given, r = l + U(0,width)
    (selective ranges, config C5); else r ~ U(l, 4) like src/write_query.c:35.  near: data rows (n x 102); the query
    vectors are then random data vectors plus N(0, near_sigma^2) noise (queries that live where the data lives)."""
    rng = np.random.default_rng(seed)
    out = np.empty((m, QROW), np.float32)
    t = np.asarray(types, np.int64)[rng.integers(0, len(types), m)]
    v = rng.integers(0, ncat, m).astype(np.float32)
    l = (rng.random(m, dtype=np.float32) * np.float32(6.0) - np.float32(3.0)).astype(np.float32)
    if range_width is None:
        r = ((np.float32(4.0) - l) * rng.random(m, dtype=np.float32) + l).astype(np.float32)
    else:
        r = (l + np.float32(range_width) * rng.random(m, dtype=np.float32)).astype(np.float32)
    out[:, 0] = t
    out[:, 1] = np.where((t == 1) | (t == 3), v, np.float32(-1.0))
    out[:, 2] = np.where(t >= 2, l, np.float32(-1.0))
    out[:, 3] = np.where(t >= 2, r, np.float32(-1.0))
    if near is None:
        out[:, 4:] = rng.random((m, 100), dtype=np.float32) * np.float32(12.0) - np.float32(6.0)
    else:
        pick = rng.integers(0, near.shape[0], m)
        out[:, 4:] = near[pick, 2:] + rng.standard_normal((m, 100), dtype=np.float32) * np.float32(near_sigma)
    return out


def write_bin(path: str, rows: np.ndarray) -> None:
    """Reference file layout: uint32 row count, then rows (include/io.h:111-136 reads it back)."""
    with open(path, "wb") as f:
        np.uint32(rows.shape[0]).tofile(f)
        np.ascontiguousarray(rows, np.float32).tofile(f)


def read_bin(path: str, width: int) -> np.ndarray:
    with open(path, "rb") as f:
        n = int(np.fromfile(f, np.uint32, 1)[0])
        return np.fromfile(f, np.float32, n * width).reshape(n, width)
