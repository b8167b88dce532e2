"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

Two ways to use G GPUs for the solve step:

* query-sharded (primary): D is replicated, rank r solves queries [lo, hi) of `query_shard`; every
  rank writes its own rows of the result -- no collective on the data path.
* data-sharded (comparison; the reference's own strategy, include/optimized_parallel.hpp:100-157, at
  GPU scale): rank r indexes rows `data_shard(n, r, G)` with id_offset = lo and produces, per query,
  its local best <= 100 (distance, global id) pairs plus the local match count
  (hvs_solve_partial_device); `gather_partials` all-gathers them shard-major ([G][m][100]) and
  hvs_merge_partials_device folds them and applies the pad rule once, globally.

Works with the nccl backend (GPU tensors) and with gloo (CPU tensors; used by the CPU tests).
"""
from __future__ import annotations


def query_shard(m: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of m queries: ranks < m % world get one more."""
    base, rem = divmod(m, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def data_shard(n: int, rank: int, world: int) -> tuple[int, int]:
    """Rows [lo, hi) of D that rank `rank` indexes (same arithmetic on every rank)."""
    return rank * n // world, (rank + 1) * n // world


def gather_partials(dist, ids, cnt, world: int):
    """all-gather the per-shard partial results.  dist/ids: [m,100], cnt: [m] (any device).
    -> (g_dist[G,m,100], g_ids[G,m,100], g_cnt[G,m]), shard-major, as hvs_merge_partials_device expects."""
    import torch
    import torch.distributed as td
    m = dist.shape[0]
    # outputs are the concatenation along dim 0 (what both gloo and nccl accept), viewed shard-major afterwards
    g_dist = torch.empty((world * m,) + tuple(dist.shape[1:]), dtype=dist.dtype, device=dist.device)
    g_ids = torch.empty((world * m,) + tuple(ids.shape[1:]), dtype=ids.dtype, device=ids.device)
    g_cnt = torch.empty((world * m,), dtype=cnt.dtype, device=cnt.device)
    td.all_gather_into_tensor(g_dist, dist.contiguous())
    td.all_gather_into_tensor(g_ids, ids.contiguous())
    td.all_gather_into_tensor(g_cnt, cnt.contiguous())
    return (g_dist.view((world, m) + tuple(dist.shape[1:])), g_ids.view((world, m) + tuple(ids.shape[1:])),
            g_cnt.view(world, m))


def gather_query_results(ids_local, m: int, world: int):
    """Query-sharded runs that want the whole result on every rank (not on the timed path):
    all-gather variable-length row blocks by padding to the largest shard."""
    import torch
    import torch.distributed as td
    rank = td.get_rank()
    width = ids_local.shape[1]
    longest = max(query_shard(m, r, world)[1] - query_shard(m, r, world)[0] for r in range(world))
    pad = torch.zeros((longest, width), dtype=ids_local.dtype, device=ids_local.device)
    pad[: ids_local.shape[0]] = ids_local
    out = torch.empty((world * longest, width), dtype=ids_local.dtype, device=ids_local.device)
    td.all_gather_into_tensor(out, pad)
    out = out.view(world, longest, width)
    parts = []
    for r in range(world):
        lo, hi = query_shard(m, r, world)
        parts.append(out[r, : hi - lo])
    del rank
    return torch.cat(parts, 0)
