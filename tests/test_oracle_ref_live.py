"""Live differential test: oracle vs the reference compiled into oracle/_ref (skipped where
the prebuilt reference is absent)."""
import numpy as np
import pytest


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
