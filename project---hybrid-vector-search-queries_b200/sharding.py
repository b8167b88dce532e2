"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

Two ways to use G GPUs for the solve step:

* query-sharded (primary): D is replicated; `solve_sharded` has every rank solve its share of ONE batch
  (hvs_solve_shard_device: shares balanced by the rows the queries sweep, queries that share rows kept on one rank)
  and combines the rows with a single all-gather -- the only exchange, 400 bytes per query.
* data-sharded (comparison; the reference's own strategy, include/optimized_parallel.hpp:100-157, at
  GPU scale): rank r indexes rows `data_shard(n, r, G)` with id_offset = lo and produces, per query,
  its local best <= 100 (distance, global id) pairs plus the local match count
  (hvs_solve_partial_device); `gather_partials` all-gathers them shard-major ([G][m][100]) and
  hvs_merge_partials_device folds them and applies the pad rule once, globally.

Works with the nccl backend (GPU tensors) and with gloo (CPU tensors; used by the CPU tests).
"""
from __future__ import annotations


def data_shard(n: int, rank: int, world: int) -> tuple[int, int]:
    """Rows [lo, hi) of D that rank `rank` indexes (same arithmetic on every rank)."""
    return rank * n // world, (rank + 1) * n // world


def gather_partials(dist, ids, cnt, world: int, out=None):
    """all-gather the per-shard partial results.  dist/ids: [m,100], cnt: [m] (any device).
    -> (g_dist[G,m,100], g_ids[G,m,100], g_cnt[G,m]), shard-major, as hvs_merge_partials_device expects.
    `out`: optional (g_dist, g_ids, g_cnt) buffers of shapes [G*m,100], [G*m,100], [G*m] to gather into.
    The collectives are enqueued relative to torch's CURRENT stream: call this inside `with torch.cuda.stream(s)`
    where s is the stream the engine was created on, so that the merge that follows is ordered after them."""
    import torch
    import torch.distributed as td
    m = dist.shape[0]
    # outputs are the concatenation along dim 0 (what both gloo and nccl accept), viewed shard-major afterwards
    if out is not None:
        g_dist, g_ids, g_cnt = out
    else:
        g_dist = torch.empty((world * m,) + tuple(dist.shape[1:]), dtype=dist.dtype, device=dist.device)
        g_ids = torch.empty((world * m,) + tuple(ids.shape[1:]), dtype=ids.dtype, device=ids.device)
        g_cnt = torch.empty((world * m,), dtype=cnt.dtype, device=cnt.device)
    td.all_gather_into_tensor(g_dist, dist.contiguous())
    td.all_gather_into_tensor(g_ids, ids.contiguous())
    td.all_gather_into_tensor(g_cnt, cnt.contiguous())
    return (g_dist.view((world, m) + tuple(dist.shape[1:])), g_ids.view((world, m) + tuple(ids.shape[1:])),
            g_cnt.view(world, m))


def solve_sharded(engine, queries_dev, rank: int, world: int, scratch=None):
    """One batch, `world` ranks (torch.distributed initialised, one process per GPU, D indexed on every rank):
    every rank solves its share (hvs_solve_shard_device) and ONE all_gather_into_tensor brings the rows together;
    returns the [m, 100] int32 ids in query order, on every rank.  `scratch` (a dict, optional) keeps the device
    buffers between calls."""
    import torch
    import torch.distributed as td
    m = queries_dev.shape[0]
    sc = scratch if scratch is not None else {}
    dev = queries_dev.device
    if sc.get("m") != m or sc.get("world") != world:
        sc.clear()
        sc.update(m=m, world=world, own=torch.empty((m, 100), dtype=torch.int32, device=dev),
                  out=torch.empty((m, 100), dtype=torch.int32, device=dev))
    _, counts = engine.solve_shard_device(queries_dev, rank, world, sc["own"], want_order=False)
    longest = int(counts.max()) if m else 0
    if sc.get("cap", -1) < longest or "gath" not in sc:    # rows every rank contributes to the gather (same on all ranks)
        sc["cap"] = max(1, min(m, longest + longest // 8 + 8))
        sc["gath"] = torch.empty((world * sc["cap"], 100), dtype=torch.int32, device=dev)
    cap = sc["cap"]
    if world == 1:
        sc["gath"][:cap].copy_(sc["own"][:cap])
    else:
        td.all_gather_into_tensor(sc["gath"], sc["own"][:cap])
    # every gathered row to its query's position: a kernel of the engine (torch's index_select + index_copy_ take 1.5 ms
    # for 4x10^4 rows of 400 bytes; this takes microseconds), with the assignment the engine kept on the device
    engine.shard_scatter_device(sc["gath"], cap, sc["out"])
    return sc["out"]
