"""Every rank's share of the headline batch, one after the other on ONE GPU (hvs_solve_shard_device, world = W): per rank
the solve / K3 / K5 time, the query count and type mix, the pairs and tensor items -- what the balance of the cost model
(SHARD_QUERY_COST, stripes) looks like without an 8-GPU box.
usage: python tools/shard_all_ranks.py W"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
hvs = importlib.import_module("project---hybrid-vector-search-queries_b200")
W = int(sys.argv[1])
d = hvs.gen_data(10_000_000, 3, ncat=100)
q = hvs.gen_queries(40_000, 4, ncat=100)
qt = q[:, 0].astype(np.int64)
stream = torch.cuda.Stream()
keys = ("ms_solve_device", "ms_plan", "ms_tile", "ms_finalize", "n_items_tensor", "m", "pairs", "pairs_tile", "n_tile", "n_direct")
with hvs.Engine(stream=stream.cuda_stream) as e:
    e.index_build(d)
    with torch.cuda.stream(stream):
        qd = torch.from_numpy(q).cuda()
        out = torch.empty((40_000, 100), dtype=torch.int32, device="cuda")
        for r in range(W):
            for _ in range(3):
                order, counts = e.solve_shard_device(qd, r, W, out)
            ts = []
            for _ in range(5):
                e.solve_shard_device(qd, r, W, out)
                ts.append(e.stats())
            st = {k: float(np.mean([t[k] for t in ts])) for k in keys}
            off = int(counts[:r].sum())
            mine = order[off:off + int(counts[r])]
            mix = np.bincount(qt[mine], minlength=4)
            print(f"W={W} r={r} solve={st['ms_solve_device']:.3f} plan={st['ms_plan']:.3f} k3={st['ms_tile']:.3f} k5={st['ms_finalize']:.3f} "
                  f"m={int(st['m'])} types={mix.tolist()} pairs={st['pairs']:.4g} pairs_tile={st['pairs_tile']:.4g} items={int(st['n_items_tensor'])} "
                  f"n_tile={int(st['n_tile'])} n_direct={int(st['n_direct'])}", flush=True)
