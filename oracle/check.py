"""check.py -- tie-aware parity checker.  TEST INFRASTRUCTURE ONLY.

Implements the parity rule of SURVEY.md 8c on top of the reference's own acceptance test:
  * the reference compares DISTANCES, position-wise, abs tol 0.002 (src/compare_data.cpp:5,40-60),
    after re-scoring every returned id with the scalar sequential fp32 distance
    (include/io.h:38-78);
  * north_star tightens that to 1e-5 relative and asks for identical id lists up to
    reordering among equal-distance ties;
  * recall@100 is a MULTISET intersection, because the pad rule (include/baseline.hpp:138-147)
    can legitimately return the same id twice.
"""
from __future__ import annotations

from collections import Counter
from dataclasses import dataclass, field

import numpy as np

from . import oracle as _o

REF_ABS_TOL = 0.002   # src/compare_data.cpp:5
RTOL = 1e-5           # BASELINE.json north_star


@dataclass
class Parity:
    n_queries: int = 0
    not_ascending: int = 0        # rows whose re-scored distances are not non-decreasing
    pos_fail_rel: int = 0         # positions with |d_ref-d_got| > RTOL*d_ref
    pos_fail_abs: int = 0         # positions with |d_ref-d_got| >= 0.002 (the reference's own test)
    dist_bit_identical_rows: int = 0
    id_rows_differ: int = 0       # rows whose id multisets differ at all
    id_fail: int = 0              # ... and the difference is NOT explained by a (near-)tie
    out_of_range_ids: int = 0
    max_rel: float = 0.0
    recall_mean: float = 1.0
    recall_min: float = 1.0
    bad_rows: list = field(default_factory=list)

    @property
    def ok(self) -> bool:
        return (self.not_ascending == 0 and self.pos_fail_rel == 0 and self.pos_fail_abs == 0
                and self.id_fail == 0 and self.out_of_range_ids == 0)

    def summary(self) -> str:
        return (f"queries={self.n_queries} ok={self.ok} bit_identical_dist_rows={self.dist_bit_identical_rows} "
                f"id_rows_differ={self.id_rows_differ} id_fail={self.id_fail} pos_fail_rel={self.pos_fail_rel} "
                f"pos_fail_abs={self.pos_fail_abs} not_ascending={self.not_ascending} max_rel={self.max_rel:.3g} "
                f"recall_mean={self.recall_mean:.6f} recall_min={self.recall_min:.4f} bad_rows={self.bad_rows[:8]}")


def compare(nodes, queries, ids_ref, ids_got, rtol: float = RTOL) -> Parity:
    nodes = np.ascontiguousarray(nodes, np.float32)
    queries = np.ascontiguousarray(queries, np.float32)
    ids_ref = np.ascontiguousarray(ids_ref, np.uint32)
    ids_got = np.ascontiguousarray(ids_got, np.uint32)
    assert ids_ref.shape == ids_got.shape and ids_ref.shape[1] == _o.K
    p = Parity(n_queries=ids_ref.shape[0])
    n = nodes.shape[0]
    p.out_of_range_ids = int((ids_got >= n).sum())
    if p.out_of_range_ids:
        return p
    d_ref = _o.rescore(nodes, queries, ids_ref)
    d_got = _o.rescore(nodes, queries, ids_got)
    p.not_ascending = int((np.diff(d_got, axis=1) < 0).any(axis=1).sum())
    diff = np.abs(d_ref.astype(np.float64) - d_got.astype(np.float64))
    scale = np.maximum(np.abs(d_ref.astype(np.float64)), 1e-30)
    rel = diff / scale
    p.max_rel = float(rel.max()) if rel.size else 0.0
    p.pos_fail_rel = int((rel > rtol).sum())
    p.pos_fail_abs = int((diff >= REF_ABS_TOL).sum())
    p.dist_bit_identical_rows = int((d_ref.view(np.uint32) == d_got.view(np.uint32)).all(axis=1).sum())
    recalls = np.ones(p.n_queries)
    same_sorted = (np.sort(ids_ref, axis=1) == np.sort(ids_got, axis=1)).all(axis=1)
    for i in np.nonzero(~same_sorted)[0]:
        p.id_rows_differ += 1
        cr, cg = Counter(ids_ref[i].tolist()), Counter(ids_got[i].tolist())
        inter = sum((cr & cg).values())
        recalls[i] = inter / _o.K
        # ids present on one side only must sit within rtol of the rank-100 boundary distance
        bound = float(d_ref[i, -1])
        only = list((cr - cg).elements()) + list((cg - cr).elements())
        q = queries[i, 4:]
        ok = True
        for j in only:
            dj = float(_o.dist_seq(nodes[j, 2:], q))
            if abs(dj - bound) > rtol * max(abs(bound), 1e-30):
                ok = False
        if not ok:
            p.id_fail += 1
            if len(p.bad_rows) < 32:
                p.bad_rows.append(int(i))
    p.recall_mean = float(recalls.mean()) if recalls.size else 1.0
    p.recall_min = float(recalls.min()) if recalls.size else 1.0
    return p
