"""oracle.py -- Python face of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product (the CUDA engine behind include/hvs.h) never does.

Three layers, strongest first:
  * `ref_vec_query(impl, ...)`  -- the UNMODIFIED reference compiled from /root/reference
    into oracle/_ref/libref_*.so by oracle/Makefile (kind "reference").
  * `vec_query(...)`            -- our plain-C restatement oracle/hvs_oracle.c of
    include/baseline.hpp:68-190 (kind "port"), pinned against the reference's outputs
    (tests/golden/, tests/test_oracle_*.py).
  * `vec_query_numpy(...)`      -- an independent numpy restatement for small cases.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
K = 100
DIM = 100
DROW = 102
QROW = 104

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


def build(verbose: bool = False) -> None:
    """Compile liboracle.so and (when /root/reference exists) oracle/_ref/*."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", HERE, "-j8"], stdout=out)


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.hvs_oracle_dist_seq.restype = C.c_float
        L.hvs_oracle_dist_seq.argtypes = [_f32p, _f32p]
        L.hvs_oracle_dist_avx_order.restype = C.c_float
        L.hvs_oracle_dist_avx_order.argtypes = [_f32p, _f32p]
        L.hvs_oracle_query_type.restype = C.c_uint32
        L.hvs_oracle_query_type.argtypes = [C.c_float]
        L.hvs_oracle_query_cat.restype = C.c_int32
        L.hvs_oracle_query_cat.argtypes = [C.c_float]
        L.hvs_oracle_vec_query.restype = C.c_int
        L.hvs_oracle_vec_query.argtypes = [_f32p, C.c_uint32, _f32p, C.c_uint32, C.c_float,
                                           _u32p, C.c_void_p, C.c_void_p]
        L.hvs_oracle_vec_query_optimized.restype = C.c_int
        L.hvs_oracle_vec_query_optimized.argtypes = [_f32p, C.c_uint32, _f32p, C.c_uint32, C.c_float,
                                                     _u32p, C.c_void_p]
        L.hvs_oracle_rescore.restype = None
        L.hvs_oracle_rescore.argtypes = [_f32p, _f32p, C.c_uint32, _u32p, _f32p]
        L.hvs_oracle_refgen_data.restype = None
        L.hvs_oracle_refgen_data.argtypes = [C.c_uint, C.c_uint32, _f32p]
        L.hvs_oracle_refgen_query.restype = C.c_uint32
        L.hvs_oracle_refgen_query.argtypes = [C.c_uint, C.c_uint32, _f32p]
        _lib = L
    return _lib


def _chk(nodes, queries):
    nodes = np.ascontiguousarray(nodes, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    assert nodes.ndim == 2 and nodes.shape[1] == DROW, nodes.shape
    assert queries.ndim == 2 and queries.shape[1] == QROW, queries.shape
    return nodes, queries


def dist_seq(x, q) -> np.float32:
    return np.float32(lib().hvs_oracle_dist_seq(np.ascontiguousarray(x, np.float32),
                                                np.ascontiguousarray(q, np.float32)))


def dist_avx_order(x, q) -> np.float32:
    return np.float32(lib().hvs_oracle_dist_avx_order(np.ascontiguousarray(x, np.float32),
                                                      np.ascontiguousarray(q, np.float32)))


def vec_query(nodes, queries, sample_proportion: float = 1.0, want_dist=True, want_nmatch=False):
    """include/baseline.hpp:68-190 restated (C).  -> ids[M,100] (+ dist[M,100], nmatch[M])."""
    nodes, queries = _chk(nodes, queries)
    m = queries.shape[0]
    ids = np.empty((m, K), np.uint32)
    dist = np.empty((m, K), np.float32) if want_dist else None
    nmatch = np.empty(m, np.uint32) if want_nmatch else None
    rc = lib().hvs_oracle_vec_query(nodes, nodes.shape[0], queries, m, sample_proportion, ids,
                                    dist.ctypes.data if want_dist else None,
                                    nmatch.ctypes.data if want_nmatch else None)
    if rc != 0:
        raise ValueError("oracle: n < 100 (reference would read out of bounds, baseline.hpp:138-147)")
    out = [ids]
    if want_dist:
        out.append(dist)
    if want_nmatch:
        out.append(nmatch)
    return out[0] if len(out) == 1 else tuple(out)


def vec_query_optimized(nodes, queries, sample_proportion: float = 1.0):
    """include/optimized.hpp:54-146 + Knn (optimized_impl.h:179-438) restated (C)."""
    nodes, queries = _chk(nodes, queries)
    m = queries.shape[0]
    ids = np.empty((m, K), np.uint32)
    dist = np.empty((m, K), np.float32)
    rc = lib().hvs_oracle_vec_query_optimized(nodes, nodes.shape[0], queries, m, sample_proportion,
                                              ids, dist.ctypes.data)
    if rc != 0:
        raise ValueError("oracle: n < 100")
    return ids, dist


def rescore(nodes, queries, ids) -> np.ndarray:
    """include/io.h:50-78 (SaveKNNFull): sequential fp32 distance of each returned id."""
    nodes, queries = _chk(nodes, queries)
    ids = np.ascontiguousarray(ids, np.uint32)
    out = np.empty(ids.shape, np.float32)
    lib().hvs_oracle_rescore(nodes, queries, queries.shape[0], ids, out)
    return out


def refgen_data(seed: int, n: int) -> np.ndarray:
    """src/write_data.c restated with srand(seed) (glibc rand)."""
    out = np.empty((n, DROW), np.float32)
    lib().hvs_oracle_refgen_data(seed, n, out)
    return out


def refgen_query(seed: int, m: int) -> np.ndarray:
    """src/write_query.c restated with srand(seed) (glibc rand)."""
    out = np.empty((m, QROW), np.float32)
    got = lib().hvs_oracle_refgen_query(seed, m, out)
    return out[:got]


# ----------------------------------------------------------------------------------------------
# independent numpy restatement (small cases only)
def decode_query(qrow):
    """include/baseline.hpp:90-93."""
    L = lib()
    return (int(L.hvs_oracle_query_type(float(qrow[0]))), int(L.hvs_oracle_query_cat(float(qrow[1]))),
            np.float32(qrow[2]), np.float32(qrow[3]))


def match_mask(nodes, qtype, v, l, r, sn=None):
    """include/baseline.hpp:107-136 vectorised over rows."""
    n = nodes.shape[0]
    sn = n if sn is None else sn
    Cc, T = nodes[:sn, 0], nodes[:sn, 1]
    if qtype == 0:
        m = np.ones(sn, bool)
    elif qtype == 1:
        m = Cc == np.float32(v)
    elif qtype == 2:
        m = (T >= l) & (T <= r)
    elif qtype == 3:
        m = (Cc == np.float32(v)) & (T >= l) & (T <= r)
    else:
        m = np.zeros(sn, bool)
    return m


def dist_seq_rows(x_rows, q):
    """baseline.hpp:53-64 for many rows: per-row sequential fp32 sum (vectorised over rows only)."""
    s = np.zeros(x_rows.shape[0], np.float32)
    for i in range(DIM):
        d = x_rows[:, i] - q[i]
        s += d * d
    return s


def vec_query_numpy(nodes, queries, sample_proportion: float = 1.0):
    nodes, queries = _chk(nodes, queries)
    n = nodes.shape[0]
    sn = min(n, int(np.uint32(np.float32(sample_proportion) * np.float32(n))))
    ids_out = np.empty((queries.shape[0], K), np.uint32)
    dist_out = np.empty((queries.shape[0], K), np.float32)
    for i, qrow in enumerate(queries):
        t, v, l, r = decode_query(qrow)
        cand = np.nonzero(match_mask(nodes, t, v, l, r, sn))[0].astype(np.uint32)
        if cand.size < K:  # baseline.hpp:138-147
            pad = n - np.arange(1, K - cand.size + 1, dtype=np.uint32)
            cand = np.concatenate([cand, pad.astype(np.uint32)])
        d = dist_seq_rows(nodes[cand, 2:], qrow[4:])
        order = np.argsort(d, kind="stable")[:K]
        ids_out[i] = cand[order]
        dist_out[i] = d[order]
    return ids_out, dist_out


# ----------------------------------------------------------------------------------------------
# the real reference, in process
_REF_NAMES = {"baseline": "libref_baseline.so", "optimized": "libref_optimized.so",
              "parallel": "libref_parallel.so", "parallel_nodbg": "libref_parallel_nodbg.so"}
_ref_libs = {}


def ref_available(impl: str = "baseline") -> bool:
    return os.path.exists(os.path.join(REF_DIR, _REF_NAMES[impl]))


def ref_vec_query(impl, nodes, queries, sample_proportion: float = 1.0):
    """Run the unmodified reference vec_query (IMPL chosen at its compile time).
    -> (ids[M,100], seconds spent inside vec_query)."""
    nodes, queries = _chk(nodes, queries)
    if impl not in _ref_libs:
        L = C.CDLL(os.path.join(REF_DIR, _REF_NAMES[impl]))
        L.ref_vec_query.restype = C.c_double
        L.ref_vec_query.argtypes = [_f32p, C.c_uint32, _f32p, C.c_uint32, C.c_float, _u32p]
        _ref_libs[impl] = L
    m = queries.shape[0]
    ids = np.empty((m, K), np.uint32)
    secs = _ref_libs[impl].ref_vec_query(nodes, nodes.shape[0], queries, m, sample_proportion, ids)
    if secs < 0:
        raise RuntimeError(f"reference vec_query failed ({secs})")
    return ids, secs


def _ref_lib(impl):
    if impl not in _ref_libs:
        L = C.CDLL(os.path.join(REF_DIR, _REF_NAMES[impl]))
        L.ref_vec_query.restype = C.c_double
        L.ref_vec_query.argtypes = [_f32p, C.c_uint32, _f32p, C.c_uint32, C.c_float, _u32p]
        _ref_libs[impl] = L
    L = _ref_libs[impl]
    if hasattr(L, "ref_set_nodes") and not getattr(L, "_session_api", False):
        L.ref_set_nodes.restype = C.c_int
        L.ref_set_nodes.argtypes = [_f32p, C.c_uint32]
        L.ref_clear_nodes.restype = None
        L.ref_query_loaded.restype = C.c_double
        L.ref_query_loaded.argtypes = [_f32p, C.c_uint32, C.c_float, _u32p]
        L._session_api = True
    return L


class RefSession:
    """The reference with D converted to its nested vectors ONCE (10^7 heap rows, ~4.5 GB) for many vec_query calls.
    `query` may run on several host threads at once for the single-threaded variants (baseline, optimized)."""

    def __init__(self, impl, nodes):
        self.impl = impl
        self.L = _ref_lib(impl)
        self.nodes = np.ascontiguousarray(nodes, np.float32)
        self.loaded = hasattr(self.L, "ref_set_nodes")
        if self.loaded and self.L.ref_set_nodes(self.nodes, self.nodes.shape[0]) != 0:
            raise RuntimeError("ref_set_nodes failed")

    def query(self, queries, sample_proportion: float = 1.0):
        queries = np.ascontiguousarray(queries, np.float32)
        m = queries.shape[0]
        ids = np.empty((m, K), np.uint32)
        if self.loaded:
            secs = self.L.ref_query_loaded(queries, m, sample_proportion, ids)
        else:                                              # an older prebuilt library: the one-shot entry point
            secs = self.L.ref_vec_query(self.nodes, self.nodes.shape[0], queries, m, sample_proportion, ids)
        if secs < 0:
            raise RuntimeError(f"reference vec_query failed ({secs})")
        return ids, secs

    def close(self):
        if self.loaded:
            self.L.ref_clear_nodes()
            self.loaded = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
