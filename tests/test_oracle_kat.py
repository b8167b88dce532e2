"""Known-answer tests that pin the oracle's arithmetic to the reference's only published
vector, src/fp_inaccuracy_test.cpp:77-97."""
import numpy as np


def kat_vectors():
    vb = [np.float32(0.11232)]
    for i in range(1, 102):  # fp_inaccuracy_test.cpp:79-84: float *= double literal, rounded to float
        vb.append(np.float32(np.float64(vb[i - 1]) * (1.321431 if i % 2 == 0 else -0.87382)))
    a = np.array(vb, np.float32)
    return a, a[::-1].copy()


def test_scalar_sequential_kat(oracle):
    a, b = kat_vectors()
    assert float(oracle.dist_seq(a[2:], b[2:])) == 277762.34375


def test_avx2_order_kat(oracle):
    a, b = kat_vectors()
    assert float(oracle.dist_avx_order(a[2:], b[2:])) == 277762.28125


def test_fp64_reference_value():
    a, b = kat_vectors()
    vd = [0.11232]
    for i in range(1, 102):
        vd.append(vd[i - 1] * (1.321431 if i % 2 == 0 else -0.87382))
    vd = np.array(vd)
    s = 0.0
    for x, y in zip(vd[2:], vd[::-1][2:]):
        s += (x - y) * (x - y)
    assert abs(s - 277762.245000211) < 1e-6


def test_numpy_rows_match_c(oracle):
    rng = np.random.default_rng(0)
    x = (rng.random((64, 100), dtype=np.float32) * 12 - 6).astype(np.float32)
    q = (rng.random(100, dtype=np.float32) * 12 - 6).astype(np.float32)
    d = oracle.dist_seq_rows(x, q)
    for i in range(64):
        assert d[i] == oracle.dist_seq(x[i], q)


def test_query_decode_truncation(oracle):
    L = oracle.lib()
    assert L.hvs_oracle_query_cat(7.9) == 7            # SURVEY S2 [probed]
    assert L.hvs_oracle_query_cat(-7.9) == -7
    assert L.hvs_oracle_query_cat(-0.5) == 0
    assert L.hvs_oracle_query_type(2.7) == 2
    assert L.hvs_oracle_query_type(-0.5) == 0
    assert L.hvs_oracle_query_type(-1.0) == 0xFFFFFFFF
    assert L.hvs_oracle_query_type(4.0) == 4
