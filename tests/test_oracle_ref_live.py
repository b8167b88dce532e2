"""Live differential test: oracle vs the reference compiled into oracle/_ref (skipped where
the prebuilt reference is absent)."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed,n,m,ncat", [(3, 3000, 40, 5), (4, 1000, 30, 3), (5, 5000, 24, 50)])
def test_oracle_vs_live_reference(oracle, check, datagen, seed, n, m, ncat):
    if not oracle.ref_available("baseline"):
        pytest.skip("oracle/_ref not built")
    d = datagen.gen_data(n, seed, ncat)
    q = datagen.gen_queries(m, seed + 100, ncat)
    ids = oracle.vec_query(d, q, want_dist=False)
    rid, _ = oracle.ref_vec_query("baseline", d, q)
    p = check.compare(d, q, rid, ids)
    assert p.ok and p.dist_bit_identical_rows == m, p.summary()
    assert np.array_equal(rid, ids)                       # no exact ties in continuous random data


def test_refgen_matches_reference_generators(oracle, tmp_path):
    import os, subprocess
    wd = os.path.join(oracle.REF_DIR, "write_data")
    shim = os.path.join(oracle.REF_DIR, "time_shim.so")
    if not (os.path.exists(wd) and os.path.exists(shim)):
        pytest.skip("oracle/_ref not built")
    env = dict(os.environ, HVS_SEED="5", LD_PRELOAD=shim)
    dp, qp = str(tmp_path / "d.bin"), str(tmp_path / "q.bin")
    subprocess.check_call([wd, dp, "2000"], env=env, stdout=subprocess.DEVNULL)
    subprocess.check_call([os.path.join(oracle.REF_DIR, "write_query"), qp, "50"], env=env, stdout=subprocess.DEVNULL)
    d = np.fromfile(dp, np.float32, offset=4).reshape(-1, 102)
    q = np.fromfile(qp, np.float32, offset=4).reshape(-1, 104)
    assert np.array_equal(d.view(np.uint32), oracle.refgen_data(5, 2000).view(np.uint32))
    assert np.array_equal(q.view(np.uint32), oracle.refgen_query(5, 50).view(np.uint32))


def test_compare_dist_mirror_against_reference_comparer(hvs, tmp_path):
    """The host mirror of src/compare_data.cpp (compare_dist_files) prints the verdict the reference's own compare.out
    prints for the same pairs of `.dist` files (the reference appends ".dist" to its arguments, src/compare_data.cpp:103)."""
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "compare.out")
    if not os.path.exists(exe):
        pytest.skip("reference comparer not built (oracle/_ref/compare.out)")
    rng = np.random.default_rng(5)
    a = (rng.random((7, 100), dtype=np.float32) * 3000).astype(np.float32)
    a.sort(axis=1)
    b = a.copy(); b[3, 10] += np.float32(0.0009765625)           # below the 0.002 threshold
    c = a.copy(); c[6, 99] += np.float32(1.0); c[0, 0] += np.float32(0.25)   # two errors
    paths = {}
    for name, t in (("a", a), ("a2", a), ("b", b), ("c", c)):
        paths[name] = str(tmp_path / f"{name}.bin")
        hvs.save_knn_dist(t, paths[name] + ".dist")
    for x, y in (("a", "a2"), ("a", "b"), ("a", "c")):
        r = subprocess.run([exe, paths[x], paths[y]], capture_output=True, text=True, timeout=60)
        mine = hvs.compare_dist_files(paths[x] + ".dist", paths[y] + ".dist")
        assert mine["verdict"] in r.stdout, (mine, r.stdout, r.stderr)
