"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: shard arithmetic, the shard-major
all-gather layout that hvs_merge_partials_device consumes, id offsets, match counts and the
"pad once, globally" rule.  The per-shard partial results come from the oracle (test infrastructure),
the gather runs through the product's sharding.py over gloo, and the merge is restated in numpy here."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "project---hybrid-vector-search-queries_b200"
K = 100


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _partial_from_oracle(O, d_shard, lo, q, n_total):
    """What hvs_solve_partial_device returns for one shard: best <=100 (dist, global id) ascending,
    unused slots (+inf, 0xFFFFFFFF), and the match count -- WITHOUT the pad rule."""
    m = q.shape[0]
    dist = np.full((m, K), np.inf, np.float32)
    ids = np.full((m, K), 0xFFFFFFFF, np.uint32)
    cnt = np.zeros(m, np.uint32)
    for i in range(m):
        t, v, l, r = O.decode_query(q[i])
        mask = O.match_mask(d_shard, t, v, l, r)
        rows = np.nonzero(mask)[0]
        cnt[i] = rows.size
        if rows.size:
            dd = O.dist_seq_rows(d_shard[rows, 2:], q[i, 4:])
            order = np.lexsort((rows, dd))[:K]
            dist[i, : order.size] = dd[order]
            ids[i, : order.size] = rows[order] + lo
    return dist, ids, cnt


def _worker(rank, world, port, n, m, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as td
    from oracle import oracle as O
    sh = importlib.import_module(PKG + ".sharding")
    dg = importlib.import_module(PKG + ".datagen")
    td.init_process_group("gloo", rank=rank, world_size=world)
    d = dg.gen_data(n, 51, ncat=40)
    q = dg.gen_queries(m, 52, ncat=40)
    lo, hi = sh.data_shard(n, rank, world)
    dist, ids, cnt = _partial_from_oracle(O, d[lo:hi], lo, q, n)
    g_dist, g_ids, g_cnt = sh.gather_partials(torch.from_numpy(dist), torch.from_numpy(ids.view(np.int32)),
                                              torch.from_numpy(cnt.view(np.int32)), world)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), g_dist=g_dist.numpy(), g_ids=g_ids.numpy().view(np.uint32),
             g_cnt=g_cnt.numpy().view(np.uint32))
    td.destroy_process_group()


def _merge_numpy(O, d, q, g_dist, g_ids, g_cnt):
    """K5 merge restated: fold G lists, pad once from the global tail, order by (distance, id)."""
    n, m = d.shape[0], q.shape[0]
    out = np.empty((m, K), np.uint32)
    for i in range(m):
        dd = g_dist[:, i].reshape(-1)
        ii = g_ids[:, i].reshape(-1)
        ok = ii != 0xFFFFFFFF
        dd, ii = dd[ok], ii[ok]
        order = np.lexsort((ii, dd))[:K]
        dd, ii = dd[order], ii[order]
        total = int(g_cnt[:, i].astype(np.int64).sum())
        if total < K:
            pad = (n - np.arange(1, K - total + 1)).astype(np.uint32)
            pd = O.dist_seq_rows(d[pad, 2:], q[i, 4:])
            dd, ii = np.concatenate([dd, pd]), np.concatenate([ii, pad])
            order = np.lexsort((ii, dd))
            dd, ii = dd[order], ii[order]
        out[i] = ii[:K]
    return out


def test_shard_arithmetic(hvs):
    sh = importlib.import_module(PKG + ".sharding")
    for n, w in [(10_000_000, 8), (1001, 2), (100, 3)]:
        edges = [sh.data_shard(n, r, w) for r in range(w)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        assert all(edges[r][1] == edges[r + 1][0] for r in range(w - 1))


@pytest.mark.timeout(300)
def test_data_sharded_gather_and_merge_world2(tmp_path, oracle, check):
    import torch.multiprocessing as mp
    n, m, world = 6000, 48, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, m, str(tmp_path)), nprocs=world, join=True)
    dg = importlib.import_module(PKG + ".datagen")
    d = dg.gen_data(n, 51, ncat=40)
    q = dg.gen_queries(m, 52, ncat=40)
    ref, nmatch = oracle.vec_query(d, q, want_dist=False, want_nmatch=True)
    assert (nmatch < K).any() and (nmatch >= K).any()          # both the pad rule and the plain merge are exercised
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    for k in ("g_dist", "g_ids", "g_cnt"):                         # every rank holds the same gathered tensors
        assert np.array_equal(r0[k], r1[k]), k
    assert r0["g_dist"].shape == (world, m, K) and r0["g_cnt"].shape == (world, m)
    assert np.array_equal(r0["g_cnt"].astype(np.int64).sum(0), nmatch)
    half = n // 2                                                   # shard-major: slice s holds ids of shard s only
    v0, v1 = r0["g_ids"][0], r0["g_ids"][1]
    assert (v0[v0 != 0xFFFFFFFF] < half).all() and (v1[v1 != 0xFFFFFFFF] >= half).all()
    got = _merge_numpy(oracle, d, q, r0["g_dist"], r0["g_ids"], r0["g_cnt"])
    p = check.compare(d, q, ref, got)
    assert p.ok and p.dist_bit_identical_rows == m, p.summary()


# ---- query-sharded solve of ONE batch (strong scaling): assignment + the single all-gather ---------------------------
def _headline_like_slices(m, n, ncat, seed):
    """Slices shaped like BASELINE configs[2]: a quarter each of all-row sweeps, T ranges, categories, category ranges."""
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 4, m)
    arena = np.where((t == 1) | (t == 3), 1, 0).astype(np.uint32)
    begin = np.zeros(m, np.uint32)
    end = np.full(m, n, np.uint32)
    lo = rng.random(m)
    hi = lo + rng.random(m) * (1.0 - lo)
    r2 = t == 2
    begin[r2] = (lo[r2] * n).astype(np.uint32)
    end[r2] = np.maximum((hi[r2] * n).astype(np.uint32), begin[r2])
    cat = rng.integers(0, ncat, m)
    clen = n // ncat
    c1 = t == 1
    begin[c1] = cat[c1] * clen
    end[c1] = (cat[c1] + 1) * clen
    c3 = t == 3
    begin[c3] = (cat[c3] * clen + lo[c3] * clen).astype(np.uint32)
    end[c3] = np.maximum((cat[c3] * clen + hi[c3] * clen).astype(np.uint32), begin[c3])
    return t, arena, begin, end, cat


def test_shard_assign_balance_locality_determinism(hvs):
    n, m, ncat = 10_000_000, 40_000, 100
    t, arena, begin, end, cat = _headline_like_slices(m, n, ncat, 7)
    rows = np.maximum(end - begin, 100).astype(np.int64)
    cost = rows + np.minimum(14 * rows, 14_000_000)       # hvs_engine.h shard_cost_of: rows + a fixed part that grows with the slice
    for world in (1, 2, 4, 8):
        order, counts = hvs.shard_assign(arena, begin, end, world)
        order2, counts2 = hvs.shard_assign(arena, begin, end, world)
        assert np.array_equal(order, order2) and np.array_equal(counts, counts2)          # a pure function of the slices
        assert int(counts.sum()) == m and sorted(order.tolist()) == list(range(m))      # every query exactly once
        owner = np.empty(m, np.int64)
        off = 0
        for r in range(world):
            owner[order[off:off + counts[r]]] = r
            off += int(counts[r])
        per = np.array([cost[owner == r].sum() for r in range(world)], np.float64)
        assert per.max() / per.mean() < 1.03, (world, per / per.mean())                   # balanced by rows swept
        if world > 1:
            # every rank gets the same mix: all-row sweeps split evenly, and ranges, and category queries
            for typ in range(4):
                share = np.array([(t[owner == r] == typ).sum() for r in range(world)], np.float64)
                assert share.min() > 0 and share.max() / share.mean() < 1.6, (world, typ, share)
            # queries of one category stay on few ranks (they share rows: the tile kernels batch them)
            spread = [len(set(owner[(t == 1) & (cat == c)].tolist())) for c in range(ncat)]
            assert np.mean(spread) <= 2.0, np.mean(spread)
            # each rank's queries come in (arena, begin, end) order
            off = 0
            for r in range(world):
                o = order[off:off + counts[r]]
                key = (arena[o].astype(np.uint64) << np.uint64(63)) | (begin[o].astype(np.uint64) << np.uint64(31)) | (end[o].astype(np.uint64) >> np.uint64(1))
                assert (np.diff(key.astype(np.float64)) >= 0).all()
                off += int(counts[r])
    for bad in (0, 256):
        with pytest.raises(hvs.HvsError):
            hvs.shard_assign(arena[:10], begin[:10], end[:10], bad)
    order, counts = hvs.shard_assign(arena[:0], begin[:0], end[:0], 4)
    assert counts.tolist() == [0, 0, 0, 0]
    order, counts = hvs.shard_assign(arena[:3], begin[:3], end[:3], 8)                      # fewer queries than ranks
    assert int(counts.sum()) == 3


class _StubEngine:
    """Stands in for Engine.solve_shard_device on the CPU: the assignment is the product's own (hvs_shard_assign_host),
    the 'answer' of query i is the row [i, i+1, ..., i+99] so that any misplaced row shows."""

    def __init__(self, hvs, arena, begin, end):
        self.hvs, self.sl = hvs, (arena, begin, end)

    def solve_shard_device(self, q, rank, world, out, want_order=True):
        import torch
        order, counts = self.hvs.shard_assign(*self.sl, world)
        off = int(counts[:rank].sum())
        own = order[off:off + int(counts[rank])].astype(np.int64)
        out[: len(own)] = torch.from_numpy(own[:, None] + np.arange(100)[None, :]).to(torch.int32)
        out[len(own):] = -7                                 # stale rows beyond the rank's share must never be picked up
        self.last = (order, counts)
        return order, counts

    def shard_scatter_device(self, gathered, cap, out):
        """hvs_shard_scatter_device restated: row j of rank r (gathered[r * cap + j]) goes to out[order[off_r + j]]."""
        import torch
        order, counts = self.last
        off = 0
        for r, c in enumerate(counts.tolist()):
            out[torch.from_numpy(order[off:off + c].astype(np.int64))] = gathered[r * cap: r * cap + c]
            off += c


def _worker_sharded(rank, world, port, m, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as td
    hvs = importlib.import_module(PKG)
    sh = importlib.import_module(PKG + ".sharding")
    td.init_process_group("gloo", rank=rank, world_size=world)
    res = []
    scratch = {}
    for mm, seed in ((m, 1), (m, 2), (m // 3, 3), (2 * m, 4)):           # changing batch sizes re-size the gather block
        _, arena, begin, end, _ = _headline_like_slices(mm, 1_000_000, 10, seed)
        eng = _StubEngine(hvs, arena, begin, end)
        q = torch.zeros((mm, 104), dtype=torch.float32)
        out = sh.solve_sharded(eng, q, rank, world, scratch)
        res.append(out.numpy().copy())
    np.savez(os.path.join(out_dir, f"s{rank}.npz"), *res)
    td.destroy_process_group()


@pytest.mark.timeout(300)
def test_solve_sharded_gather_world2(tmp_path, hvs):
    import torch.multiprocessing as mp
    m, world = 900, 2
    port = _free_port()
    mp.spawn(_worker_sharded, args=(world, port, m, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "s0.npz"), np.load(tmp_path / "s1.npz")
    for k, mm in zip(r0.files, (m, m, m // 3, 2 * m)):
        want = np.arange(mm)[:, None] + np.arange(100)[None, :]
        assert np.array_equal(r0[k], want) and np.array_equal(r1[k], want), k      # query order, on every rank
