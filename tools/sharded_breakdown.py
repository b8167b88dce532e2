"""Where a strong-scaling step goes outside the engine call (run under torchrun, N ranks)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
hvs = importlib.import_module("project---hybrid-vector-search-queries_b200")
sh = importlib.import_module("project---hybrid-vector-search-queries_b200.sharding")
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
d = hvs.gen_data(10_000_000, 3, ncat=100)
q = hvs.gen_queries(40_000, 4, ncat=100)
stream = torch.cuda.Stream()
eng = hvs.Engine(device=lr, stream=stream.cuda_stream)
eng.index_build(d)
with torch.cuda.stream(stream):
    qd = torch.from_numpy(q).cuda()
    own = torch.empty((40_000, 100), dtype=torch.int32, device="cuda")
    sc = {}
    for _ in range(3):
        sh.solve_sharded(eng, qd, rank, world, sc)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    T = {"engine": 0.0, "rest_enqueue": 0.0, "rest_sync": 0.0, "total": 0.0, "solve_dev": 0.0}
    steps = 10
    for _ in range(steps):
        t0 = time.perf_counter()
        order, counts = eng.solve_shard_device(qd, rank, world, own)
        t1 = time.perf_counter()
        T["solve_dev"] += eng.stats()["ms_solve_device"]
        sc["own"].copy_(own) if False else None
        out = sh.solve_sharded.__wrapped__(eng, qd, rank, world, sc) if hasattr(sh.solve_sharded, "__wrapped__") else None
        t2 = time.perf_counter()
        stream.synchronize()
        t3 = time.perf_counter()
        T["engine"] += (t1 - t0) * 1e3; T["total"] += (t3 - t0) * 1e3
    # second loop: the real call, synchronised per step
    T2 = 0.0
    for _ in range(steps):
        t0 = time.perf_counter()
        sh.solve_sharded(eng, qd, rank, world, sc)
        stream.synchronize()
        T2 += (time.perf_counter() - t0) * 1e3
    # third: back to back (as bench.py times it)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        sh.solve_sharded(eng, qd, rank, world, sc)
    stream.synchronize()
    T3 = (time.perf_counter() - t0) * 1e3
print(f"rank {rank}: engine call wall {T['engine']/steps:.3f} ms (device {T['solve_dev']/steps:.3f}); solve_sharded + sync {T2/steps:.3f} ms; back-to-back {T3/steps:.3f} ms per step", flush=True)
dist.barrier()
eng.close()
dist.destroy_process_group()
