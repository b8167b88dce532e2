"""Stress probe (run under gpurun): many seeded solves, AUTO (tcgen05 sweep) and EXACT (FFMA sweep) against DIRECT
(the direct scan is the reference's arithmetic by construction and is pinned to the oracle by tests/).  Catches rare
scheduling-dependent errors (dynamic item hand-out, synchronised compaction, overlapping launches) that one seeded test
would miss.  usage: python tools/stress_parity.py [iterations]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("HVS_MIN_TILE_PAIRS", "0")
hvs = importlib.import_module("project---hybrid-vector-search-queries_b200")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = np.random.default_rng(2026)
bad = 0
for it in range(iters):
    n = int(rng.integers(20_000, 400_000))
    m = int(rng.integers(64, 3000))
    ncat = int(rng.choice([1, 3, 10, 50, 300]))
    zipf = float(rng.choice([0.0, 0.0, 1.1]))
    clusters = int(rng.choice([0, 0, 0, 25]))
    d = hvs.gen_data(n, 1000 + it, ncat=ncat, zipf=zipf, clusters=clusters, cluster_sigma=1.5)
    q = hvs.gen_queries(m, 2000 + it, ncat=ncat, near=d if clusters else None, near_sigma=1.5)
    res = {}
    for name, mode in (("direct", hvs.MODE_DIRECT), ("auto", hvs.MODE_AUTO), ("exact", hvs.MODE_EXACT)):
        with hvs.Engine(mode=mode) as e:
            e.index_build(d)
            ids = e.solve(q)
            res[name] = (ids, e.rescore(q, ids), e.stats())
    ref_ids, ref_dist, _ = res["direct"]
    for name in ("auto", "exact"):
        ids, dist, st = res[name]
        same_dist = np.array_equal(dist.view(np.uint32), ref_dist.view(np.uint32))     # ties may permute ids, never distances
        ok = same_dist and all(np.array_equal(np.sort(a), np.sort(b)) or len(set(np.round(x, 6) for x in dd)) < 100
                               for a, b, dd in zip(ids, ref_ids, dist))
        if not ok:
            bad += 1
        print(f"it={it} n={n} m={m} ncat={ncat} zipf={zipf} clusters={clusters} {name}: ok={ok} tile_q={st['n_tile']} items={st['n_items_ffma']}+{st['n_items_tensor']} fallback={st['n_fallback']}", flush=True)
print("STRESS", "FAILED" if bad else "OK", f"({bad} bad of {2 * iters})")
sys.exit(1 if bad else 0)
