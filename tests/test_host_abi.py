"""CPU-side checks of the product's host layer: the C-ABI library loads, exports exactly the symbols
include/hvs.h declares, refuses to run without a GPU (no CPU fallback), and the host planner behaves."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def H(hvs):
    if not os.path.exists(hvs.LIB_PATH):
        hvs.build()
    hvs.lib()
    return hvs


def test_header_symbols_exported(H):
    hdr = open(os.path.join(ROOT, "include", "hvs.h")).read()
    declared = set(re.findall(r"HVS_API\s+[\w\s\*]+?\b(hvs_\w+)\s*\(", hdr))
    assert declared == set(H.ABI_SYMBOLS), declared ^ set(H.ABI_SYMBOLS)
    L = ctypes.CDLL(H.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert H.lib().hvs_abi_version() == 1


def test_struct_layouts_match_header(H):
    assert ctypes.sizeof(H.Config) == 32
    assert ctypes.sizeof(H.Stats) == 136


def test_no_cpu_fallback(H):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(H.HvsError) as ei:
        H.Engine()
    assert ei.value.code == H.HVS_ERR_NO_DEVICE
    with pytest.raises(H.HvsError):
        H.vec_query(np.zeros((100, 102), np.float32), np.zeros((1, 104), np.float32), 1.0, [])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "project---hybrid-vector-search-queries_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"#include[^\n]*oracle|^\s*(import|from)\s+oracle|liboracle|dlopen[^\n]*oracle", src, re.M), f


def test_planner_direct_vs_tile(H):
    # 300 queries share the whole T arena (type 0) -> tile path; 5 lone short slices -> direct
    m = 305
    arena = np.zeros(m, np.uint32); begin = np.zeros(m, np.uint32); end = np.full(m, 500_000, np.uint32)
    arena[300:] = 1
    begin[300:] = np.arange(5) * 1000
    end[300:] = begin[300:] + 900
    kind, items, pc = H.plan_dryrun(arena, begin, end, H.MODE_EXACT)
    assert kind[:300].all() and not kind[300:].any()
    assert len(items) > 0 and (items[:, 0] == 0).all()
    # every item: <= 128 queries, rows inside the arena, FFMA kind in exact mode
    assert ((items[:, 3] & 0xffff) <= 128).all() and (items[:, 3] >> 16 == 0).all()
    assert (items[:, 2] > items[:, 1]).all() and items[:, 2].max() <= 500_000
    # the items cover rows x queries exactly once: sum(rows * nq) == 300 * 500000
    assert int(((items[:, 2] - items[:, 1]).astype(np.int64) * (items[:, 3] & 0xffff)).sum()) == 300 * 500_000
    assert pc == 300 * 500_000 + 5 * 900
    kind2, items2, _ = H.plan_dryrun(arena, begin, end, H.MODE_DIRECT)
    assert not kind2.any() and len(items2) == 0


def test_planner_rejects_bad_slices(H):
    with pytest.raises(H.HvsError):
        H.plan_dryrun([2], [0], [1])
    with pytest.raises(H.HvsError):
        H.plan_dryrun([0], [5], [1])
    kind, items, pc = H.plan_dryrun(np.zeros(0), np.zeros(0), np.zeros(0))
    assert len(kind) == 0 and len(items) == 0 and pc == 0


def test_compare_dist_mirrors_reference_checker(H, tmp_path):
    """src/compare_data.cpp: |a-b| >= 0.002 is an error; identical tables are 'the same'; files must be whole."""
    import numpy as np
    import pytest
    a = np.linspace(1.0, 5000.0, 3 * 100, dtype=np.float32).reshape(3, 100)
    assert H.compare_dist(a, a)["verdict"] == "Datasets are the same!"
    b = a.copy(); b[1, 7] += np.float32(0.0009765625)
    v = H.compare_dist(a, b)
    assert v["ok"] and v["verdict"] == "Datasets are similar under error delta!" and 0 < v["max_error"] < 0.002
    c = a.copy(); c[2, 99] += np.float32(0.5); c[0, 0] = np.nan
    v = H.compare_dist(a, c)
    assert not v["ok"] and v["errors"] == 2 and v["verdict"].startswith("ERROR: Found a total of 2")
    assert not H.compare_dist(a, a[:2])["ok"]
    p = str(tmp_path / "x.dist")
    H.save_knn_dist(a, p)
    assert np.array_equal(H.read_knn_dist(p), a)
    with open(p, "r+b") as f:
        f.truncate(4 + 4 * 100 * 2)            # the reference's comparer would call two short files "the same"
    with pytest.raises(ValueError):
        H.read_knn_dist(p)


def test_planner_tiny_job_goes_direct(H):
    """BASELINE.json configs[0] sized job (100 queries x 10^4 rows = 10^6 pairs): a tile sweep's fixed cost would
    dominate, so every query takes the direct scan; the same shape 100x larger is swept in tiles."""
    m = 100
    arena = np.zeros(m, np.uint32); begin = np.zeros(m, np.uint32)
    for mode in (H.MODE_EXACT, H.MODE_AUTO):
        kind, items, pc = H.plan_dryrun(arena, begin, np.full(m, 10_000, np.uint32), mode)
        assert not kind.any() and len(items) == 0 and pc == m * 10_000
        kind, items, _ = H.plan_dryrun(arena, begin, np.full(m, 1_000_000, np.uint32), mode)
        assert kind.all() and len(items) > 0


def test_planner_items_cover_rows_once(H):
    """Host statement of the planner (hvs_plan_dryrun): category-sized slices cut at chunk boundaries -- the items cover
    every (query, row) of the tile queries (an item sweeps the union of its queries' rows inside its chunk, so where a
    chunk holds the end of one category and the start of the next a little is swept for nothing: < 3 % here)."""
    ncat, cs, per = 40, 100_000, 2000
    m = ncat * per
    arena = np.ones(m, np.uint32)
    begin = (np.repeat(np.arange(ncat), per) * cs).astype(np.uint32)
    end = (begin + cs).astype(np.uint32)
    kind, items, pc = H.plan_dryrun(arena, begin, end, H.MODE_AUTO)
    assert kind.all()
    rows = (items[:, 2] - items[:, 1]).astype(np.int64)
    nq = (items[:, 3] & 0xffff).astype(np.int64)
    assert int((rows * nq).sum()) == pc and m * cs <= pc <= 1.03 * m * cs, (int((rows * nq).sum()), m * cs, pc)
    assert nq.max() == 256 and rows.max() <= (1 << 22)
