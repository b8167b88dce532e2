"""One small solve for profiling: python tools/prof_case.py N M NCAT MODE [types]  (run under ncu via gpurun)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hvs = importlib.import_module("project---hybrid-vector-search-queries_b200")
n, m, ncat, mode = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
types = tuple(int(c) for c in sys.argv[5]) if len(sys.argv) > 5 else (0, 1, 2, 3)
mode_id = {"auto": hvs.MODE_AUTO, "exact": hvs.MODE_EXACT, "direct": hvs.MODE_DIRECT, "tensor": hvs.MODE_TENSOR}[mode]
d = hvs.gen_data(n, 3, ncat=ncat)
q = hvs.gen_queries(m, 4, ncat=ncat, types=types)
with hvs.Engine(mode=mode_id) as e:
    e.index_build(d)
    for _ in range(2):
        ids = e.solve(q)
        st = e.stats()
    print({k: st[k] for k in ("ms_solve_device", "ms_plan", "ms_direct", "ms_tile_ffma", "ms_tile_tensor", "ms_finalize", "pairs", "n_items_ffma", "n_items_tensor", "n_direct", "n_fallback", "launches")})
