"""hvs-b200: host-side mirror of the reference's solve-step interface over the C-ABI library.

The product is `libhvs_b200.so` (CUDA, sm_100a; ABI in include/hvs.h).  This module is the thin
Python face of it -- ctypes only, plain pointers and sizes -- and mirrors the reference's own
interface for the path:

  * `vec_query(nodes, queries, sample_proportion, knn_results)`  -- same name, argument meaning and
    result convention as the reference's one operator symbol (include/baseline.hpp:68-69,
    include/optimized.hpp:54-55, include/optimized_parallel.hpp:61-62; called at src/test.cpp:85):
    `knn_results` receives one 100-element list of uint32 row ids per query, in query order.
  * `read_bin` / `save_knn`  -- the io.h driver contract (include/io.h:111-136, :23-36).
  * `Engine`  -- index once, solve many (what the C++ shim include/hvs_vec_query.hpp does inside
    vec_query()).

There is NO CPU fallback: if the library is missing or no sm_100 device is present every entry
point raises.  Nothing here imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .datagen import gen_data, gen_queries, read_bin as _read_bin, write_bin  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhvs_b200.so")
K, DIM, DROW, QROW = 100, 100, 102, 104

MODE_AUTO, MODE_EXACT, MODE_DIRECT, MODE_TENSOR = 0, 1, 2, 3
FLAG_USE_GIVEN_STREAM, FLAG_MARGIN_AUDIT = 1, 2
HVS_OK, HVS_ERR_INVALID, HVS_ERR_NO_DEVICE, HVS_ERR_CUDA, HVS_ERR_STATE, HVS_ERR_NOMEM = 0, -1, -2, -3, -4, -5

# every symbol include/hvs.h declares (tests check the library exports exactly these)
ABI_SYMBOLS = (
    "hvs_abi_version", "hvs_last_error", "hvs_create", "hvs_destroy", "hvs_set_mode", "hvs_index_build", "hvs_index_build_device",
    "hvs_index_build_rows", "hvs_index_build_from_file",
    "hvs_solve", "hvs_solve_full", "hvs_solve_device", "hvs_solve_shard_device", "hvs_shard_scatter_device", "hvs_shard_assign_host",
    "hvs_solve_partial_device", "hvs_merge_partials_device", "hvs_rescore",
    "hvs_get_stats", "hvs_measure_ffma_peak", "hvs_plan_dryrun",
)


class HvsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"hvs error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("mode", C.c_uint32), ("flags", C.c_uint32),
                ("stream", C.c_void_p), ("id_offset", C.c_uint32), ("reserved", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("n", C.c_uint32), ("n_total", C.c_uint32), ("m", C.c_uint32),
                ("pairs", C.c_uint64), ("pairs_computed", C.c_uint64), ("rows_union", C.c_uint64),
                ("n_direct", C.c_uint32), ("n_tile", C.c_uint32), ("n_items_ffma", C.c_uint32),
                ("n_items_tensor", C.c_uint32), ("n_fallback", C.c_uint32), ("launches", C.c_uint32),
                ("ms_index_build", C.c_float), ("ms_h2d", C.c_float), ("ms_plan", C.c_float), ("ms_direct", C.c_float),
                ("ms_tile", C.c_float), ("ms_tile_ffma", C.c_float), ("ms_tile_tensor", C.c_float),
                ("ms_finalize", C.c_float), ("ms_d2h", C.c_float), ("ms_solve_device", C.c_float),
                ("ms_solve_wall", C.c_float), ("pairs_tile", C.c_uint64), ("pairs_direct", C.c_uint64),
                ("n_outliers", C.c_uint32), ("margin_audit", C.c_float)]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "struct_size"}


def build(verbose: bool = False) -> str:
    """Compile libhvs_b200.so in-tree with nvcc for sm_100a (csrc/Makefile)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(HERE, "csrc"), "-j8"], stdout=out)
    return LIB_PATH


_lib = None


def lib():
    """The C-ABI library.  Fails loudly when it has not been built: there is nothing to fall back to."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HvsError(HVS_ERR_STATE, f"{LIB_PATH} is not built (run __graft_entry__.build() or "
                                          f"make -C {os.path.join(HERE, 'csrc')}); this engine has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, u32, i32, f32 = C.c_void_p, C.c_uint32, C.c_int, C.c_float
        L.hvs_abi_version.restype = u32
        L.hvs_last_error.restype = C.c_char_p
        L.hvs_last_error.argtypes = [vp]
        L.hvs_create.restype = i32
        L.hvs_create.argtypes = [C.POINTER(vp), C.POINTER(Config)]
        L.hvs_destroy.restype = None
        L.hvs_destroy.argtypes = [vp]
        L.hvs_set_mode.restype = i32
        L.hvs_set_mode.argtypes = [vp, u32]
        for name in ("hvs_index_build", "hvs_index_build_device"):
            f = getattr(L, name)
            f.restype = i32
            f.argtypes = [vp, vp, u32, f32]
        L.hvs_index_build_rows.restype = i32
        L.hvs_index_build_rows.argtypes = [vp, vp, u32, f32]
        L.hvs_index_build_from_file.restype = i32
        L.hvs_index_build_from_file.argtypes = [vp, C.c_char_p, f32, C.POINTER(u32)]
        for name in ("hvs_solve", "hvs_solve_device"):
            f = getattr(L, name)
            f.restype = i32
            f.argtypes = [vp, vp, u32, vp]
        L.hvs_solve_shard_device.restype = i32
        L.hvs_solve_shard_device.argtypes = [vp, vp, u32, u32, u32, vp, vp, vp]
        L.hvs_shard_scatter_device.restype = i32
        L.hvs_shard_scatter_device.argtypes = [vp, vp, u32, vp]
        L.hvs_shard_assign_host.restype = i32
        L.hvs_shard_assign_host.argtypes = [vp, vp, vp, u32, u32, vp, vp]
        L.hvs_solve_partial_device.restype = i32
        L.hvs_solve_partial_device.argtypes = [vp, vp, u32, vp, vp, vp]
        L.hvs_merge_partials_device.restype = i32
        L.hvs_merge_partials_device.argtypes = [vp, vp, u32, u32, vp, vp, vp, vp, u32, vp]
        L.hvs_rescore.restype = i32
        L.hvs_rescore.argtypes = [vp, vp, u32, vp, vp]
        L.hvs_solve_full.restype = i32
        L.hvs_solve_full.argtypes = [vp, vp, u32, vp, vp]
        L.hvs_get_stats.restype = i32
        L.hvs_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.hvs_measure_ffma_peak.restype = i32
        L.hvs_measure_ffma_peak.argtypes = [vp, u32, C.POINTER(f32), C.POINTER(f32)]
        L.hvs_plan_dryrun.restype = i32
        L.hvs_plan_dryrun.argtypes = [vp, vp, vp, u32, u32, vp, vp, u32, C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def _host_f32(a, width: int) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] != width:
        raise ValueError(f"expected an [n, {width}] float32 array, got {a.shape}")
    return a


def _dev_ptr(t, dtype_name: str, numel: int) -> int:
    """Raw device pointer of a CUDA torch tensor (torch is plumbing: it only owns the memory)."""
    if not (hasattr(t, "is_cuda") and t.is_cuda and t.is_contiguous()):
        raise ValueError("expected a contiguous CUDA tensor")
    if str(t.dtype) != "torch." + dtype_name or t.numel() < numel:
        raise ValueError(f"expected {numel} x {dtype_name}, got {t.numel()} x {t.dtype}")
    return t.data_ptr()


class Engine:
    """One engine = one GPU: hvs_create / hvs_index_build / hvs_solve / hvs_destroy."""

    def __init__(self, device: int = -1, mode: int = MODE_AUTO, id_offset: int = 0, stream: int | None = None, flags: int = 0):
        self._h = C.c_void_p()
        cfg = Config(struct_size=C.sizeof(Config), device=device, mode=mode, flags=flags, stream=stream, id_offset=id_offset)
        rc = lib().hvs_create(C.byref(self._h), C.byref(cfg))
        if rc != HVS_OK:
            raise HvsError(rc, lib().hvs_last_error(None).decode())

    def _ck(self, rc: int) -> None:
        if rc != HVS_OK:
            raise HvsError(rc, lib().hvs_last_error(self._h).decode())

    def set_mode(self, mode: int) -> None:
        self._ck(lib().hvs_set_mode(self._h, mode))

    def close(self) -> None:
        if self._h:
            lib().hvs_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- indexing phase (never sees queries) ----
    def index_build(self, nodes, sample_proportion: float = 1.0) -> None:
        if hasattr(nodes, "is_cuda") and nodes.is_cuda:
            n = nodes.shape[0]
            self._ck(lib().hvs_index_build_device(self._h, _dev_ptr(nodes, "float32", n * DROW), n, sample_proportion))
        else:
            a = _host_f32(nodes, DROW)
            self._ck(lib().hvs_index_build(self._h, a.ctypes.data, a.shape[0], sample_proportion))

    def index_build_from_file(self, path: str, sample_proportion: float = 1.0) -> int:
        """Index the D file directly (reference layout, README.md:32-44); returns the row count."""
        n = C.c_uint32()
        self._ck(lib().hvs_index_build_from_file(self._h, os.fsencode(path), sample_proportion, C.byref(n)))
        return n.value

    # ---- solve step ----
    def solve(self, queries, out: np.ndarray | None = None) -> np.ndarray:
        """Host buffers in, host buffers out (H2D and D2H inside the call)."""
        q = _host_f32(queries, QROW)
        m = q.shape[0]
        if out is None:
            out = np.empty((m, K), np.uint32)
        assert out.dtype == np.uint32 and out.flags.c_contiguous and out.size >= m * K
        self._ck(lib().hvs_solve(self._h, q.ctypes.data, m, out.ctypes.data))
        return out

    def solve_full(self, queries) -> tuple[np.ndarray, np.ndarray]:
        """ids as solve(), plus the distances SaveKNNFull would write for them (include/io.h:50-78), one call."""
        q = _host_f32(queries, QROW)
        m = q.shape[0]
        ids = np.empty((m, K), np.uint32)
        dist = np.empty((m, K), np.float32)
        self._ck(lib().hvs_solve_full(self._h, q.ctypes.data, m, ids.ctypes.data, dist.ctypes.data))
        return ids, dist

    def solve_device(self, queries_dev, out_dev) -> None:
        m = queries_dev.shape[0]
        self._ck(lib().hvs_solve_device(self._h, _dev_ptr(queries_dev, "float32", m * QROW), m,
                                        _dev_ptr(out_dev, "int32", m * K)))

    def solve_shard_device(self, queries_dev, rank: int, world: int, out_dev, want_order: bool = True):
        """Query-sharded solve (hvs_solve_shard_device): this rank's share of the batch every rank passes.
        -> (order[m] rank-major query indices or None, counts[world]); out_dev[:counts[rank]] holds the answers of
        order[offset(rank):][:counts[rank]].  No communication: sharding.solve_sharded combines the ranks."""
        m = queries_dev.shape[0]
        order = np.empty(m, np.uint32) if want_order else None
        counts = np.zeros(world, np.uint32)
        self._ck(lib().hvs_solve_shard_device(self._h, _dev_ptr(queries_dev, "float32", m * QROW), m, rank, world,
                                              _dev_ptr(out_dev, "int32", m * K), order.ctypes.data if want_order else None,
                                              counts.ctypes.data))
        return order, counts

    def shard_scatter_device(self, gathered_dev, cap: int, out_dev) -> None:
        """After the all-gather of solve_shard_device's rows (rank r's at gathered[r * cap ...]): every row to its
        query's position in out_dev[m, 100] (hvs_shard_scatter_device; enqueued on the engine's stream)."""
        self._ck(lib().hvs_shard_scatter_device(self._h, _dev_ptr(gathered_dev, "int32", cap * K), cap,
                                                _dev_ptr(out_dev, "int32", K)))

    def solve_partial_device(self, queries_dev, out_dist_dev, out_ids_dev, out_count_dev) -> None:
        m = queries_dev.shape[0]
        self._ck(lib().hvs_solve_partial_device(self._h, _dev_ptr(queries_dev, "float32", m * QROW), m,
                                                _dev_ptr(out_dist_dev, "float32", m * K),
                                                _dev_ptr(out_ids_dev, "int32", m * K),
                                                _dev_ptr(out_count_dev, "int32", m)))

    def merge_partials_device(self, queries_dev, g: int, dist_dev, ids_dev, count_dev, tail_rows_dev, n_total: int,
                              out_ids_dev) -> None:
        m = queries_dev.shape[0]
        self._ck(lib().hvs_merge_partials_device(self._h, _dev_ptr(queries_dev, "float32", m * QROW), m, g,
                                                 _dev_ptr(dist_dev, "float32", g * m * K),
                                                 _dev_ptr(ids_dev, "int32", g * m * K),
                                                 _dev_ptr(count_dev, "int32", g * m),
                                                 _dev_ptr(tail_rows_dev, "float32", K * DROW), n_total,
                                                 _dev_ptr(out_ids_dev, "int32", m * K)))

    def rescore(self, queries, ids) -> np.ndarray:
        """include/io.h:50-78 (SaveKNNFull): sequential fp32 distance of every returned id."""
        q = _host_f32(queries, QROW)
        ids = np.ascontiguousarray(ids, np.uint32)
        out = np.empty(ids.shape, np.float32)
        self._ck(lib().hvs_rescore(self._h, q.ctypes.data, q.shape[0], ids.ctypes.data, out.ctypes.data))
        return out

    def stats(self) -> dict:
        st = Stats(struct_size=C.sizeof(Stats))
        self._ck(lib().hvs_get_stats(self._h, C.byref(st)))
        return st.as_dict()

    def measure_ffma_peak(self, iters: int = 5) -> tuple[float, float]:
        tf, mhz = C.c_float(), C.c_float()
        self._ck(lib().hvs_measure_ffma_peak(self._h, iters, C.byref(tf), C.byref(mhz)))
        return tf.value, mhz.value


def vec_query(nodes, queries, sample_proportion: float, knn_results: list, mode: int = MODE_AUTO, device: int = -1) -> None:
    """The reference's operator, same contract (include/baseline.hpp:68-69; src/test.cpp:80-85):
    `knn_results` arrives empty and receives len(queries) lists of 100 row ids, ascending by distance."""
    d = _host_f32(nodes, DROW)
    q = _host_f32(queries, QROW)
    with Engine(device=device, mode=mode) as e:
        e.index_build(d, sample_proportion)
        ids = e.solve(q)
    knn_results.extend(ids.tolist())


def read_bin(path: str, dims: int) -> np.ndarray:
    """include/io.h:111-136 ReadBin: uint32 row count, then rows of `dims` float32."""
    return _read_bin(path, dims)


def save_knn(knn_results, path: str) -> None:
    """include/io.h:23-36 SaveKNN: M x 100 uint32, no header, query order."""
    a = np.ascontiguousarray(knn_results, dtype=np.uint32)
    if a.ndim != 2 or a.shape[1] != K:
        raise ValueError("every result row must hold exactly 100 ids (include/io.h:29)")
    a.tofile(path)


def save_knn_dist(dist, path: str) -> None:
    """The `.dist` side file of include/io.h:50-78: uint32 M, then M x 100 float32."""
    a = np.ascontiguousarray(dist, dtype=np.float32)
    with open(path, "wb") as f:
        np.uint32(a.shape[0]).tofile(f)
        a.tofile(f)


def read_knn_dist(path: str) -> np.ndarray:
    """include/io.h ReadBinFull as src/compare_data.cpp uses it: uint32 M, then M x 100 float32.  Unlike the
    reference (whose comparer reports "the same" for two missing files) a short or missing file is an error."""
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        m = int(np.fromfile(f, np.uint32, 1)[0])
        if size != 4 + 4 * K * m:
            raise ValueError(f"{path}: header says {m} queries, file holds {size} bytes (expected {4 + 4 * K * m})")
        return np.fromfile(f, np.float32, m * K).reshape(m, K)


def compare_dist(a: np.ndarray, b: np.ndarray, error_delta: float = 0.002) -> dict:
    """src/compare_data.cpp:8-84 on two M x 100 distance tables: position-wise |a - b| in double; a difference
    >= error_delta (src/compare_data.cpp:5) is an error.  Returns the counts and the verdict line it would print."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    if a.shape != b.shape:
        return {"ok": False, "errors": -1, "max_error": float("nan"),
                "verdict": f"Datasets have different number of queries! {a.shape[0]}, {b.shape[0]}"}
    diff = np.abs(a.astype(np.float64) - b.astype(np.float64))
    errs = int(np.count_nonzero(~(diff < error_delta)))     # NaN counts as an error
    mx = float(diff.max()) if diff.size else 0.0
    if errs == 0 and mx == 0.0:
        verdict = "Datasets are the same!"
    elif errs == 0:
        verdict = "Datasets are similar under error delta!"
    else:
        verdict = f"ERROR: Found a total of {errs} differences!"
    return {"ok": errs == 0, "errors": errs, "max_error": mx, "verdict": verdict}


def compare_dist_files(a_path: str, b_path: str, error_delta: float = 0.002) -> dict:
    return compare_dist(read_knn_dist(a_path), read_knn_dist(b_path), error_delta)


def shard_assign(arena, begin, end, world: int):
    """hvs_shard_assign_host (CPU only): -> (order[m] rank-major query indices, counts[world])."""
    arena = np.ascontiguousarray(arena, np.uint32)
    begin = np.ascontiguousarray(begin, np.uint32)
    end = np.ascontiguousarray(end, np.uint32)
    m = arena.shape[0]
    order = np.empty(m, np.uint32)
    counts = np.zeros(world, np.uint32)
    rc = lib().hvs_shard_assign_host(arena.ctypes.data, begin.ctypes.data, end.ctypes.data, m, world, order.ctypes.data,
                                     counts.ctypes.data)
    if rc != HVS_OK:
        raise HvsError(rc, "hvs_shard_assign_host: invalid slices or world")
    return order, counts


def plan_dryrun(arena, begin, end, mode: int = MODE_EXACT, max_items: int = 1 << 20):
    """Host planner only (no device needed): -> (kind[m] 0=direct/1=tile, items[k,4], pairs_computed)."""
    arena = np.ascontiguousarray(arena, np.uint32)
    begin = np.ascontiguousarray(begin, np.uint32)
    end = np.ascontiguousarray(end, np.uint32)
    m = arena.shape[0]
    kind = np.zeros(m, np.uint8)
    items = np.zeros((max_items, 4), np.uint32)
    pc = C.c_uint64()
    n = lib().hvs_plan_dryrun(arena.ctypes.data, begin.ctypes.data, end.ctypes.data, m, mode, kind.ctypes.data,
                              items.ctypes.data, max_items, C.byref(pc))
    if n < 0:
        raise HvsError(n, "hvs_plan_dryrun: invalid slices")
    return kind, items[:min(n, max_items)].copy(), int(pc.value)
