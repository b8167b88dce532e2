"""One rank's share of the headline batch on one GPU (hvs_solve_shard_device with world = W, rank = r): how the
per-rank solve time depends on the planner's chunk size (HVS_ITEMS_PER_SM, read once per process: run one value per process).
usage: python tools/shard_rank_probe.py W [rank]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
hvs = importlib.import_module("project---hybrid-vector-search-queries_b200")
W = int(sys.argv[1]); r = int(sys.argv[2]) if len(sys.argv) > 2 else 0
d = hvs.gen_data(10_000_000, 3, ncat=100)
q = hvs.gen_queries(40_000, 4, ncat=100)
stream = torch.cuda.Stream()
with hvs.Engine(stream=stream.cuda_stream) as e:
    e.index_build(d)
    with torch.cuda.stream(stream):
        qd = torch.from_numpy(q).cuda()
        out = torch.empty((40_000, 100), dtype=torch.int32, device="cuda")
        for _ in range(3):
            e.solve_shard_device(qd, r, W, out)
        ts = []
        for _ in range(6):
            e.solve_shard_device(qd, r, W, out)
            ts.append(e.stats())
    st = {k: float(np.mean([t[k] for t in ts])) for k in ("ms_solve_device", "ms_plan", "ms_tile", "ms_finalize", "n_items_tensor", "m", "ms_solve_wall")}
    print(f"world {W} rank {r} ips={os.environ.get('HVS_ITEMS_PER_SM', '16')}: " + " ".join(f"{k}={v:.3f}" for k, v in st.items()), flush=True)
