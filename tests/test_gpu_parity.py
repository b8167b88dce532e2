"""GPU parity tests: the CUDA engine, called through the C ABI (include/hvs.h), against
  * outputs of the UNMODIFIED reference committed under tests/golden/ (oracle/gen_golden.py),
  * the CPU oracle (oracle/hvs_oracle.c, itself pinned to the reference) on seeded inputs,
  * size-independent properties at sizes the oracle cannot reach.
Bar: distances bit-identical to the reference's sequential fp32 arithmetic; id lists identical up to
reordering among equal-distance ties (checked by oracle/check.py, rtol 1e-5 as BASELINE.json asks)."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-5   # BASELINE.json north_star: distances within 1e-5 relative


@pytest.fixture(scope="module")
def H(hvs):
    hvs.lib()
    return hvs


def modes(H):
    return [("direct", H.MODE_DIRECT), ("exact", H.MODE_EXACT), ("auto", H.MODE_AUTO)]


def solve(H, d, q, mode, sp=1.0):
    with H.Engine(mode=mode) as e:
        e.index_build(d, sp)
        ids = e.solve(q)
        return ids, e.stats(), e.rescore(q, ids)


def assert_parity(check, oracle, d, q, ids_ref, ids, dist_dev, tag):
    p = check.compare(d, q, ids_ref, ids, rtol=RTOL)
    assert p.ok, f"{tag}: {p.summary()}"
    # position-wise distances are not merely close: they are the same fp32 numbers
    assert p.dist_bit_identical_rows == len(q), f"{tag}: {p.summary()}"
    # the engine's own SaveKNNFull (include/io.h:50-78 on the device) equals the CPU re-score
    assert np.array_equal(dist_dev.view(np.uint32), oracle.rescore(d, q, ids).view(np.uint32)), tag


@pytest.mark.parametrize("name,sp", [("edge_small.npz", 1.0), ("sample_half.npz", 0.5), ("intcat_2k.npz", 1.0)])
def test_reference_fixtures(H, oracle, check, name, sp):
    g = load_golden(name)
    src = load_golden("edge_small.npz") if name == "sample_half.npz" else g
    d, q = src["d"], src["q"]
    for tag, mode in modes(H):
        ids, st, dd = solve(H, d, q, mode, sp)
        assert_parity(check, oracle, d, q, g["ids_baseline"], ids, dd, f"{name}/{tag}")


def test_c1_default_config(H, oracle, check):
    """BASELINE.json configs[0]: D=10^4, Q=10^2 from the reference's own generators (seed 1)."""
    import hashlib
    g = load_golden("c1_refgen_seed1.npz")
    d, q = oracle.refgen_data(1, 10000), oracle.refgen_query(1, 100)
    if hashlib.sha256(d.tobytes()).hexdigest() != str(g["sha_d"]):
        pytest.skip("glibc rand() stream differs from the fixture's")
    for tag, mode in modes(H):
        ids, st, dd = solve(H, d, q, mode)
        assert_parity(check, oracle, d, q, g["ids_baseline"], ids, dd, f"c1/{tag}")
        assert np.array_equal(ids, g["ids_baseline"]), tag      # no ties in this set: identical lists


@pytest.mark.parametrize("n,m,ncat,seed", [(50_000, 384, 8, 11), (120_000, 300, 3, 12), (4_000, 64, 50, 13)])
def test_seeded_vs_oracle(H, oracle, check, datagen, n, m, ncat, seed):
    d = datagen.gen_data(n, seed, ncat=ncat)
    q = datagen.gen_queries(m, seed + 100, ncat=ncat)
    ref = oracle.vec_query(d, q, want_dist=False)
    for tag, mode in modes(H):
        ids, st, dd = solve(H, d, q, mode)
        assert_parity(check, oracle, d, q, ref, ids, dd, f"n={n}/{tag}")
        if mode == H.MODE_DIRECT:
            assert st["n_tile"] == 0 and st["n_direct"] == m
        if mode == H.MODE_EXACT and n >= 50_000:
            assert st["n_tile"] > 0 and st["n_items_ffma"] > 0 and st["n_items_tensor"] == 0, st
        if mode == H.MODE_AUTO and n >= 50_000:         # the tcgen05 sweep ran, and it is exact too
            assert st["n_tile"] > 0 and st["n_items_tensor"] > 0 and st["n_items_ffma"] == 0, st


def test_selective_type3_pad_heavy(H, oracle, check, datagen):
    """configs[4] in miniature: ~1000 categories, narrow T ranges -> tiny slices, pad rule hot."""
    d = datagen.gen_data(200_000, 21, ncat=1000)
    q = datagen.gen_queries(500, 22, ncat=1000, types=(3,), range_width=0.06)
    ref, nmatch = oracle.vec_query(d, q, want_dist=False, want_nmatch=True)
    assert (nmatch < 100).mean() > 0.5
    for tag, mode in modes(H):
        ids, st, dd = solve(H, d, q, mode)
        assert_parity(check, oracle, d, q, ref, ids, dd, f"selective/{tag}")


def test_ties_and_duplicates(H, oracle, check):
    """Many exactly equal vectors: equal-distance ties at the rank-100 boundary and candidate lists
    that overflow their margin (the engine must fall back to the exact scan, not drop rows)."""
    rng = np.random.default_rng(5)
    n = 60_000
    d = np.empty((n, 102), np.float32)
    d[:, 0] = rng.integers(0, 2, n)
    d[:, 1] = rng.random(n) * 6 - 3
    base = (rng.random((40, 100), dtype=np.float32) * 12 - 6).astype(np.float32)
    d[:, 2:] = base[rng.integers(0, 40, n)]                     # 40 distinct vectors, ~1500 copies each
    q = np.zeros((200, 104), np.float32)
    q[:, 0] = rng.integers(0, 4, 200)
    q[:, 1] = rng.integers(0, 2, 200)
    q[:, 2] = -3
    q[:, 3] = 3
    q[:, 4:] = rng.random((200, 100), dtype=np.float32) * 12 - 6
    ref = oracle.vec_query(d, q, want_dist=False)
    for tag, mode in modes(H):
        ids, st, dd = solve(H, d, q, mode)
        p = check.compare(d, q, ref, ids, rtol=RTOL)
        assert p.pos_fail_rel == 0 and p.pos_fail_abs == 0 and p.not_ascending == 0 and p.id_fail == 0, f"{tag}: {p.summary()}"
        assert p.dist_bit_identical_rows == len(q)
        if mode == H.MODE_EXACT:
            assert st["n_fallback"] > 0, st


def test_modes_agree_and_properties_at_scale(H, datagen):
    """10^6 rows: too slow for the scalar oracle in full, so (i) the direct scan (reference arithmetic,
    no approximation anywhere) and the tile path must return identical id lists, (ii) re-scored
    distances are ascending, (iii) every id satisfies the query's predicate, (iv) a 24-query sample
    is checked against the oracle by the next test."""
    n, m = 1_000_000, 2048
    d = datagen.gen_data(n, 31, ncat=20)
    q = datagen.gen_queries(m, 32, ncat=20)
    with H.Engine(mode=H.MODE_DIRECT) as e:
        e.index_build(d)
        ids_direct = e.solve(q)
    with H.Engine(mode=H.MODE_EXACT) as e:
        e.index_build(d)
        ids = e.solve(q)
        st = e.stats()
        dist = e.rescore(q, ids)
    assert st["n_tile"] > m // 2, st
    assert np.array_equal(ids, ids_direct)
    assert (np.diff(dist, axis=1) >= 0).all()
    t = q[:, 0].astype(int)
    C, T = d[:, 0], d[:, 1]
    for i in range(0, m, 7):
        mask = np.ones(n, bool)
        if t[i] in (1, 3):
            mask &= C == q[i, 1]
        if t[i] in (2, 3):
            mask &= (T >= q[i, 2]) & (T <= q[i, 3])
        nmatch = int(mask.sum())
        hit = mask[ids[i]]
        if nmatch >= 100:
            assert hit.all(), i                        # every returned row satisfies the predicate
        else:                                          # pad rule: all matches + rows n-1, n-2, ... (baseline.hpp:138-147)
            want = sorted(np.nonzero(mask)[0].tolist() + list(range(n - (100 - nmatch), n)))
            assert sorted(ids[i].tolist()) == want, i


def test_sample_against_oracle_at_scale(H, oracle, check, datagen):
    n = 1_000_000
    d = datagen.gen_data(n, 31, ncat=20)
    q = datagen.gen_queries(2048, 32, ncat=20)
    with H.Engine(mode=H.MODE_AUTO) as e:
        e.index_build(d)
        ids = e.solve(q)
    pick = np.arange(0, 2048, 86)[:24]
    ref = oracle.vec_query(d, q[pick], want_dist=False)
    p = check.compare(d, q[pick], ref, ids[pick], rtol=RTOL)
    assert p.ok and p.dist_bit_identical_rows == len(pick), p.summary()


def test_data_sharded_partials_merge(H, oracle, check, datagen):
    """SURVEY 8e comparison variant on one GPU: two shard engines (id_offset) -> partial top-100 +
    match counts -> K5 merge applies the pad rule once, globally.  Must equal the single-engine answer."""
    import torch
    n, m = 40_000, 300
    d = datagen.gen_data(n, 41, ncat=200)
    q = datagen.gen_queries(m, 42, ncat=200)            # many type-1/3 slices shorter than 100: pad rule
    ref = oracle.vec_query(d, q, want_dist=False)
    half = n // 2
    qd = torch.from_numpy(q).cuda()
    dist = torch.empty((2, m, 100), dtype=torch.float32, device="cuda")
    ids = torch.empty((2, m, 100), dtype=torch.int32, device="cuda")
    cnt = torch.empty((2, m), dtype=torch.int32, device="cuda")
    engines = []
    for s, (lo, hi) in enumerate([(0, half), (half, n)]):
        e = H.Engine(mode=H.MODE_EXACT, id_offset=lo)
        e.index_build(d[lo:hi])
        e.solve_partial_device(qd, dist[s], ids[s], cnt[s])
        engines.append(e)
    tail = torch.from_numpy(d[n - 100:]).cuda()
    out = torch.empty((m, 100), dtype=torch.int32, device="cuda")
    engines[0].merge_partials_device(qd, 2, dist, ids, cnt, tail, n, out)
    torch.cuda.synchronize()
    got = out.cpu().numpy().view(np.uint32)
    for e in engines:
        e.close()
    p = check.compare(d, q, ref, got, rtol=RTOL)
    assert p.ok and p.dist_bit_identical_rows == m, p.summary()


def test_error_behaviour(H):
    with H.Engine() as e:
        with pytest.raises(H.HvsError) as ei:
            e.solve(np.zeros((1, 104), np.float32))
        assert ei.value.code == H.HVS_ERR_STATE
        with pytest.raises(H.HvsError) as ei:
            e.index_build(np.zeros((50, 102), np.float32))      # n < 100: reference reads nodes[n-s] out of bounds
        assert ei.value.code == H.HVS_ERR_INVALID
        e.index_build(np.zeros((100, 102), np.float32))
        assert e.solve(np.zeros((0, 104), np.float32)).shape == (0, 100)
        ids = e.solve(np.zeros((3, 104), np.float32))
        assert sorted(ids[0].tolist()) == list(range(100))


def test_vec_query_operator_contract(H, oracle, check):
    """Same call shape as src/test.cpp:80-85: results arrive in an empty list, one 100-id row per query."""
    g = load_golden("intcat_2k.npz")
    out = []
    H.vec_query(g["d"], g["q"], 1.0, out)
    assert len(out) == len(g["q"]) and all(len(r) == 100 for r in out)
    p = check.compare(g["d"], g["q"], g["ids_baseline"], np.asarray(out, np.uint32), rtol=RTOL)
    assert p.ok, p.summary()


def test_engine_reuse_and_batch_shapes(H, oracle, check, datagen):
    """One index, many solves (what the C++ shim does once, a server would do repeatedly): results must not
    depend on what was solved before, on the batch size, or on the mode of an earlier engine."""
    d = datagen.gen_data(150_000, 61, ncat=6)
    qa = datagen.gen_queries(700, 62, ncat=6)
    qb = datagen.gen_queries(33, 63, ncat=6, types=(2, 3))
    ref_a = oracle.vec_query(d, qa[:64], want_dist=False)
    ref_b = oracle.vec_query(d, qb, want_dist=False)
    with H.Engine(mode=H.MODE_AUTO) as e:
        e.index_build(d)
        a1 = e.solve(qa)
        b1 = e.solve(qb)
        one = e.solve(qa[5:6])
        a2 = e.solve(qa)
        e.index_build(d[:100_000])                       # re-index the same engine with less data
        small = e.solve(qb)
        e.index_build(d)
        a3 = e.solve(qa)
    assert np.array_equal(a1, a2) and np.array_equal(a1, a3)
    assert np.array_equal(one[0], a1[5])
    p = check.compare(d, qa[:64], ref_a, a1[:64], rtol=RTOL)
    assert p.ok and p.dist_bit_identical_rows == 64, p.summary()
    p = check.compare(d, qb, ref_b, b1, rtol=RTOL)
    assert p.ok and p.dist_bit_identical_rows == len(qb), p.summary()
    ref_small = oracle.vec_query(d[:100_000], qb, want_dist=False)
    p = check.compare(d[:100_000], qb, ref_small, small, rtol=RTOL)
    assert p.ok, p.summary()


def test_cpp_driver_same_files_same_bytes(H, oracle, check, datagen, tmp_path):
    """The drop-in driver (reference command line and file formats, include/hvs_vec_query.hpp underneath):
    output.bin must be M x 100 uint32 without header (include/io.h:23-36) and hold the reference's answer;
    output.bin.dist is what src/compare_data.cpp reads."""
    import os
    import subprocess
    drv = os.path.join(os.path.dirname(H.LIB_PATH), "driver", "hvs_test.out")
    if not os.path.exists(drv):
        pytest.skip("driver not built")
    d = datagen.gen_data(20_000, 71, ncat=12)
    q = datagen.gen_queries(120, 72, ncat=12)
    dp, qp, op = str(tmp_path / "d.bin"), str(tmp_path / "q.bin"), str(tmp_path / "out.bin")
    datagen.write_bin(dp, d)
    datagen.write_bin(qp, q)
    r = subprocess.run([drv, dp, qp, op], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Vector Search took" in r.stderr
    raw = np.fromfile(op, np.uint32)
    assert raw.size == len(q) * 100                       # headerless
    ids = raw.reshape(len(q), 100)
    ref = oracle.vec_query(d, q, want_dist=False)
    p = check.compare(d, q, ref, ids, rtol=RTOL)
    assert p.ok and p.dist_bit_identical_rows == len(q), p.summary()
    with open(op + ".dist", "rb") as f:
        m = int(np.fromfile(f, np.uint32, 1)[0])
        dist = np.fromfile(f, np.float32).reshape(m, 100)
    assert m == len(q)
    assert np.array_equal(dist.view(np.uint32), oracle.rescore(d, q, ids).view(np.uint32))


def test_index_build_from_file_and_rows(H, oracle, check, datagen, tmp_path):
    """The streaming ingest entry points must index exactly what hvs_index_build indexes:
    from the D file (reference layout) and -- through the C++ driver test above -- from N row pointers."""
    d = datagen.gen_data(70_000, 81, ncat=9)       # 28.6 MB: less than one 64 MB staging chunk ...
    q = datagen.gen_queries(150, 82, ncat=9)
    path = str(tmp_path / "d.bin")
    datagen.write_bin(path, d)
    with H.Engine(mode=H.MODE_AUTO) as e:
        e.index_build(d)
        want = e.solve(q)
    with H.Engine(mode=H.MODE_AUTO) as e:
        assert e.index_build_from_file(path) == len(d)
        got = e.solve(q)
        with pytest.raises(H.HvsError):
            e.index_build_from_file(str(tmp_path / "missing.bin"))
    assert np.array_equal(got, want)
    big = datagen.gen_data(400_000, 83, ncat=9)    # ... and 163 MB: three chunks, rows straddle chunk borders
    bpath = str(tmp_path / "big.bin")
    datagen.write_bin(bpath, big)
    with H.Engine(mode=H.MODE_EXACT) as e:
        e.index_build_from_file(bpath)
        got = e.solve(q)
    ref = oracle.vec_query(big, q[:40], want_dist=False)
    p = check.compare(big, q[:40], ref, got[:40], rtol=RTOL)
    assert p.ok and p.dist_bit_identical_rows == 40, p.summary()
    with open(bpath, "r+b") as f:                  # a file shorter than its header says is refused, not read past
        f.truncate(4 + 102 * 4 * 1000)
    with H.Engine() as e:
        with pytest.raises(H.HvsError):
            e.index_build_from_file(bpath)


def test_solve_full_dist_side_file(H, oracle, check, datagen, tmp_path):
    """hvs_solve_full = solve + SaveKNNFull (include/io.h:50-78) in one call: ids equal hvs_solve's, the distances
    are the reference's calc_dist numbers (pad rows included), and the `.dist` file compares as
    src/compare_data.cpp would against one written from the oracle's distances."""
    d = datagen.gen_data(30_000, 31, ncat=40)
    q = np.concatenate([datagen.gen_queries(200, 32, ncat=40), datagen.gen_queries(56, 33, ncat=40, types=(3,), range_width=0.02)])
    with H.Engine(mode=H.MODE_AUTO) as e:
        e.index_build(d)
        ids = e.solve(q)
        ids2, dist = e.solve_full(q)
    assert np.array_equal(ids, ids2)
    ref_dist = oracle.rescore(d, q, ids)
    assert np.array_equal(dist.view(np.uint32), ref_dist.view(np.uint32))
    a, b = str(tmp_path / "a.bin.dist"), str(tmp_path / "b.bin.dist")
    H.save_knn_dist(dist, a)
    H.save_knn_dist(ref_dist, b)
    v = H.compare_dist_files(a, b)
    assert v["ok"] and v["verdict"] == "Datasets are the same!" and v["max_error"] == 0.0
    assert H.read_knn_dist(a).shape == (len(q), 100)


@pytest.mark.parametrize("zipf,clusters,sigma", [(1.2, 0, 0.0), (0.0, 12, 0.25), (1.0, 40, 1.0)])
def test_skewed_categories_and_clustered_vectors(H, oracle, check, datagen, zipf, clusters, sigma):
    """SURVEY 8f-3 inputs: Zipf-skewed categories (huge and tiny slices side by side) and clustered vectors with
    queries drawn near the data (distances concentrate, many near-ties inside a cluster: the hard case for the
    candidate margins).  Same bar as everywhere: the reference's ids, bit-identical distances."""
    n, m = 60_000, 320
    d = datagen.gen_data(n, 41, ncat=30, zipf=zipf, clusters=clusters, cluster_sigma=sigma or 0.5)
    q = datagen.gen_queries(m, 42, ncat=30, near=d if clusters else None, near_sigma=sigma or 0.5)
    ref = oracle.vec_query(d, q, want_dist=False)
    for tag, mode in modes(H):
        ids, st, dd = solve(H, d, q, mode)
        assert_parity(check, oracle, d, q, ref, ids, dd, f"zipf={zipf} clusters={clusters}/{tag}")


# ---- the candidate margins under adversarial inputs ---------------------------------------------
def _audit_solve(H, d, q, mode):
    with H.Engine(mode=mode, flags=H.FLAG_MARGIN_AUDIT) as e:
        e.index_build(d)
        ids = e.solve(q)
        return ids, e.stats(), e.rescore(q, ids)


def test_margin_audit_regular_data(H, oracle, check, datagen):
    """The error bound behind the margins (hvs_margin.cuh), measured on the device: over every re-ranked survivor
    |approximate score + ||q||^2 - reference distance| must stay below the eps the margins assume."""
    d = datagen.gen_data(80_000, 91, ncat=4)
    q = datagen.gen_queries(512, 92, ncat=4)
    ref = oracle.vec_query(d, q[:96], want_dist=False)
    for tag, mode in (("exact", H.MODE_EXACT), ("auto", H.MODE_AUTO)):
        ids, st, dd = _audit_solve(H, d, q, mode)
        assert st["n_tile"] > 0 and 0.0 < st["margin_audit"] < 1.0, (tag, st)
        p = check.compare(d, q[:96], ref, ids[:96], rtol=RTOL)
        assert p.ok and p.dist_bit_identical_rows == 96, f"{tag}: {p.summary()}"


def test_norm_outlier_rows(H, oracle, check, datagen):
    """One row with 10^3 x the norm of the rest (and a few more, and a non-finite one) must neither wreck the fp16 scale
    nor push every query onto the exact fallback: the index sets such rows aside, K5 scores them exactly -- also when an
    outlier IS among a query's nearest rows."""
    n, m = 70_000, 300
    d = datagen.gen_data(n, 93, ncat=3)
    rng = np.random.default_rng(94)
    big = rng.choice(n - 200, 7, replace=False)
    d[big, 2:] *= 1000.0
    d[big[0], 2:] *= 1000.0                                   # 10^6 x: ||x||^2 ~ 10^15
    q = datagen.gen_queries(m, 95, ncat=3)
    q[:8, 4:] = d[big[:4].repeat(2), 2:] * np.float32(0.999)  # queries right next to outliers
    q[:8, 0] = [0, 0, 2, 2, 1, 1, 3, 3]
    q[:8, 1] = d[big[:4].repeat(2), 0]
    q[:8, 2], q[:8, 3] = -3.0, 3.0
    ref = oracle.vec_query(d, q, want_dist=False)
    assert any(big[0] in ref[i] for i in range(8))            # the outlier really is an answer
    for tag, mode in modes(H):
        ids, st, dd = _audit_solve(H, d, q, mode)
        assert_parity(check, oracle, d, q, ref, ids, dd, f"outliers/{tag}")
        if mode != H.MODE_DIRECT:
            assert st["n_outliers"] == 7 and st["n_tile"] > 0, st
            # the 8 queries with a huge norm may take the fallback (their own -2 sx q leaves fp16); nobody else does
            assert st["n_fallback"] <= 8 and st["margin_audit"] < 1.0, (tag, st)


def test_tiny_and_huge_scales(H, oracle, check, datagen):
    """All-tiny norms (fp16 would flush to subnormals without the power-of-two image scale), all-huge norms, and a
    query whose scaled norm leaves fp16's range (the `flags` path of K3: that query alone takes the exact scan)."""
    n, m = 50_000, 200
    base = datagen.gen_data(n, 96, ncat=3)
    qb = datagen.gen_queries(m, 97, ncat=3)
    for scale in (1e-4, 3e3):
        d, q = base.copy(), qb.copy()
        d[:, 2:] *= np.float32(scale)
        q[:, 4:] *= np.float32(scale)
        if scale > 1:
            q[5, 4:] *= np.float32(500.0)                     # 2 sx ||q|| > 60000
        ref = oracle.vec_query(d, q, want_dist=False)
        for tag, mode in modes(H):
            ids, st, dd = _audit_solve(H, d, q, mode)
            assert_parity(check, oracle, d, q, ref, ids, dd, f"scale={scale}/{tag}")
            if mode == H.MODE_AUTO:
                assert st["n_items_tensor"] > 0 and st["margin_audit"] < 1.0, st
                assert st["n_fallback"] <= (1 if scale > 1 else 0), st


def test_id_offset_with_pad_rows(H, oracle, check, datagen):
    """A shard engine (id_offset != 0) used through the non-partial entry points: pad ids carry the offset like real
    matches do, and hvs_rescore maps both back."""
    n, off = 20_000, 1_000_000
    d = datagen.gen_data(n, 98, ncat=500)
    q = datagen.gen_queries(120, 99, ncat=500, types=(1, 3))     # tiny categories: the pad rule on most queries
    ref = oracle.vec_query(d, q, want_dist=False)
    with H.Engine(mode=H.MODE_AUTO, id_offset=off) as e:
        e.index_build(d)
        ids = e.solve(q)
        dist = e.rescore(q, ids)
    assert ids.min() >= off
    local = (ids - np.uint32(off)).astype(np.uint32)
    p = check.compare(d, q, ref, local, rtol=RTOL)
    assert p.ok and p.dist_bit_identical_rows == len(q), p.summary()
    assert np.array_equal(dist.view(np.uint32), oracle.rescore(d, q, local).view(np.uint32))


def test_sharded_solve_covers_the_batch(H, oracle, check, datagen):
    """hvs_solve_shard_device, all ranks played by one engine: the shares are disjoint, cover the batch, and every
    share's rows are the rows hvs_solve gives for those queries (strong-scaling path of bench.py --gpus N)."""
    import torch
    n, m = 150_000, 1500
    d = datagen.gen_data(n, 101, ncat=12)
    q = datagen.gen_queries(m, 102, ncat=12)
    qd = torch.from_numpy(q).cuda()
    with H.Engine(mode=H.MODE_AUTO) as e:
        e.index_build(d)
        want = e.solve(q)
        for world in (1, 3, 8):
            got = np.full((m, 100), 0xFFFFFFFF, np.uint32)
            seen = np.zeros(m, np.int64)
            buf = torch.empty((m, 100), dtype=torch.int32, device="cuda")
            for rank in range(world):
                buf.fill_(-1)
                order, counts = e.solve_shard_device(qd, rank, world, buf)
                torch.cuda.synchronize()
                assert int(counts.sum()) == m
                off = int(counts[:rank].sum())
                own = order[off:off + int(counts[rank])].astype(np.int64)
                got[own] = buf[: len(own)].cpu().numpy().view(np.uint32)
                seen[own] += 1
                assert e.stats()["m"] == len(own)
            assert (seen == 1).all()
            assert np.array_equal(got, want), world
            # the engine's own placement kernel (what follows the NCCL all-gather in sharding.solve_sharded): rank r's rows at
            # gathered[r * cap ...], garbage in the slack rows
            cap = int(counts.max()) + 5
            gath = torch.full((world * cap, 100), -3, dtype=torch.int32, device="cuda")
            off = 0
            for r in range(world):
                own = order[off:off + int(counts[r])].astype(np.int64)
                gath[r * cap: r * cap + len(own)] = torch.from_numpy(want[own].view(np.int32)).cuda()
                off += int(counts[r])
            placed = torch.full((m, 100), -1, dtype=torch.int32, device="cuda")
            e.shard_scatter_device(gath, cap, placed)
            torch.cuda.synchronize()
            assert np.array_equal(placed.cpu().numpy().view(np.uint32), want), world
    ref = oracle.vec_query(d, q[:48], want_dist=False)
    p = check.compare(d, q[:48], ref, want[:48], rtol=RTOL)
    assert p.ok and p.dist_bit_identical_rows == 48, p.summary()


def test_many_chunks_per_slice_and_global_lists(H, oracle, check, datagen):
    """D = 4x10^6, enough queries that the planner cuts slices into 16k-row chunks swept by different CTAs (>= 128
    stages per item: the per-query global best-score lists of K3 and the shared thresholds are exercised), several
    chunks per slice.  AUTO and EXACT (independent kernel families) must agree on every query; a sample is checked
    against the oracle and every returned id against its predicate."""
    n, m = 4_000_000, 4096
    d = datagen.gen_data(n, 111, ncat=10)
    q = datagen.gen_queries(m, 112, ncat=10, types=(0, 2))
    with H.Engine(mode=H.MODE_AUTO, flags=H.FLAG_MARGIN_AUDIT) as e:
        e.index_build(d)
        ids = e.solve(q)
        st = e.stats()
        dist = e.rescore(q, ids)
        e.set_mode(H.MODE_EXACT)
        ids_exact = e.solve(q)
        st_exact = e.stats()
        dist_exact = e.rescore(q, ids_exact)
    assert st["n_items_tensor"] > 0 and st["pairs_tile"] > 9.9e9 and st["n_fallback"] == 0, st
    assert st["n_items_tensor"] * 256 * 16384 >= st["pairs_tile"] * 0.5          # items are >= 16k rows: ntiles >= 128
    assert 0.0 < st["margin_audit"] < 1.0 and 0.0 < st_exact["margin_audit"] < 1.0, (st, st_exact)
    assert st_exact["n_items_ffma"] > 0
    assert np.array_equal(dist.view(np.uint32), dist_exact.view(np.uint32))
    assert (ids == ids_exact).all(axis=1).mean() > 0.999
    assert (np.diff(dist, axis=1) >= 0).all()
    T = d[:, 1]
    for i in range(0, m, 97):
        if q[i, 0] == 2:
            assert ((T[ids[i]] >= q[i, 2]) & (T[ids[i]] <= q[i, 3])).all() or (T[(T >= q[i, 2]) & (T <= q[i, 3])].size < 100)
    pick = np.arange(5, m, m // 16)[:16]
    ref = oracle.vec_query(d, q[pick], want_dist=False)
    p = check.compare(d, q[pick], ref, ids[pick], rtol=RTOL)
    assert p.ok and p.dist_bit_identical_rows == len(pick), p.summary()


def test_reference_driver_impl4(H, oracle, check, datagen, tmp_path):
    """The reference's OWN driver (src/test.cpp: its argv handling, io.h ReadBin / SaveKNN / SaveKNNFull) compiled with the
    one-line `#elif IMPL == 4` ladder entry of INTEGRATION.md, include/hvs_vec_query.hpp behind vec_query and
    libhvs_b200.so linked (oracle/Makefile builds it into oracle/_ref/ when /root/reference is present).  Its output.bin
    must hold the reference's answer and its .dist file must pass the reference's own comparer against baseline.out's."""
    import os
    import subprocess
    ref_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    drv, base, cmp_ = (os.path.join(ref_dir, f) for f in ("hvs_impl4.out", "baseline.out", "compare.out"))
    if not all(os.path.exists(p) for p in (drv, base, cmp_)):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    d = datagen.gen_data(30_000, 121, ncat=15)
    q = datagen.gen_queries(90, 122, ncat=15)
    dp, qp = str(tmp_path / "d.bin"), str(tmp_path / "q.bin")
    datagen.write_bin(dp, d)
    datagen.write_bin(qp, q)
    outs = {}
    for name, exe in (("impl4", drv), ("baseline", base)):
        op = str(tmp_path / f"{name}.bin")
        r = subprocess.run([exe, dp, qp, op], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        assert "Vector Search took" in r.stderr
        outs[name] = op
    ids = np.fromfile(outs["impl4"], np.uint32).reshape(len(q), 100)           # SaveKNN: headerless M x 100 uint32
    ids_ref = np.fromfile(outs["baseline"], np.uint32).reshape(len(q), 100)
    p = check.compare(d, q, ids_ref, ids, rtol=RTOL)
    assert p.ok and p.dist_bit_identical_rows == len(q), p.summary()
    for o in outs.values():                                # compare.out with NDEBUG calls two missing files "the same": check first
        assert os.path.getsize(o) == 400 * len(q) and os.path.getsize(o + ".dist") == 4 + 400 * len(q)
    r = subprocess.run([cmp_, outs["impl4"], outs["baseline"]], capture_output=True, text=True, timeout=120)   # it appends ".dist" itself
    assert r.returncode == 0 and "Datasets are the same!" in (r.stdout + r.stderr), (r.stdout + r.stderr)[-500:]


def test_tiny_job_fast_path(H, oracle, check, datagen, tmp_path):
    """A job that cannot reach the planner's tiny-job bound (m x n < 4x10^6 pairs; BASELINE configs[0]) skips the planner
    altogether: slice search, K4s for the small slices, one (split) CTA scan for the rest.  The test-suite switches
    that rule off (HVS_MIN_TILE_PAIRS=0, read once per process), so this runs in a fresh interpreter without it."""
    import os
    import subprocess
    import sys
    d = datagen.gen_data(9_000, 131, ncat=10)
    q = datagen.gen_queries(120, 132, ncat=10)
    np.savez(tmp_path / "in.npz", d=d, q=q)
    code = (
        "import importlib, sys, numpy as np\n"
        f"sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})\n"
        "H = importlib.import_module('project---hybrid-vector-search-queries_b200')\n"
        f"z = np.load({str(tmp_path / 'in.npz')!r})\n"
        "with H.Engine(mode=H.MODE_AUTO) as e:\n"
        "    e.index_build(z['d']); ids = e.solve(z['q']); st = e.stats()\n"
        f"np.savez({str(tmp_path / 'out.npz')!r}, ids=ids, n_tile=st['n_tile'], n_direct=st['n_direct'], launches=st['launches'], pairs=st['pairs'])\n")
    env = {k: v for k, v in os.environ.items() if k != "HVS_MIN_TILE_PAIRS"}
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = np.load(tmp_path / "out.npz")
    assert int(out["n_tile"]) == 0 and int(out["n_direct"]) == len(q) and int(out["launches"]) <= 3, dict(out)
    ref, nmatch = oracle.vec_query(d, q, want_dist=False, want_nmatch=True)
    assert int(out["pairs"]) == int(np.maximum(nmatch, 100).sum())
    p = check.compare(d, q, ref, out["ids"], rtol=RTOL)
    assert p.ok and p.dist_bit_identical_rows == len(q), p.summary()
