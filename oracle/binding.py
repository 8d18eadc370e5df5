"""ctypes binding of the CPU oracle (oracle/libturdb_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
``--impl reference`` legs of bench.py.  The product package (turdb_b200/) never imports this.
Parity status: "parity unpinned" for HNSW search/insert/distance (see hnsw_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libturdb_oracle.so")

L2, COSINE, IP = 0, 1, 2
BUILD_VERBATIM, BUILD_INTENT = 0, 1
DIST_AVX2, DIST_AVX2_EMULATED, DIST_SCALAR = 0, 1, 2
INVALID = 0xFFFFFFFF
MAX_L0, MAX_UP = 32, 16


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (gcc only)."""
    src = os.path.join(_HERE, "hnsw_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class Stats(C.Structure):
    _fields_ = [("n_dist", C.c_uint32), ("n_dist_upper", C.c_uint32),
                ("n_expanded", C.c_uint32), ("n_upper_hops", C.c_uint32)]


STATS_DTYPE = np.dtype([("n_dist", "<u4"), ("n_dist_upper", "<u4"),
                        ("n_expanded", "<u4"), ("n_upper_hops", "<u4")])

_lib = None


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    vp = C.c_void_p
    L.tdo_distance.restype = C.c_float
    L.tdo_distance.argtypes = [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_uint32, C.c_int]
    L.tdo_select_level.restype = C.c_uint8
    L.tdo_select_level.argtypes = [C.c_double, C.c_uint16]
    L.tdo_graph_new.restype = vp
    L.tdo_graph_new.argtypes = [C.c_uint16, C.c_uint16, C.c_uint16, C.c_int]
    L.tdo_graph_free.argtypes = [vp]
    L.tdo_graph_insert.argtypes = [vp, C.c_uint64, C.POINTER(C.c_float), C.c_double]
    L.tdo_graph_insert_batch.argtypes = [vp, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_float),
                                         C.POINTER(C.c_double)]
    L.tdo_graph_from_arrays.restype = vp
    L.tdo_graph_from_arrays.argtypes = [C.c_uint16, C.c_uint64, C.POINTER(C.c_float), C.POINTER(C.c_uint64),
                                        C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8),
                                        C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8),
                                        C.c_uint64, C.c_uint32, C.c_uint8]
    for name, rt in [("tdo_graph_n", C.c_uint64), ("tdo_graph_n_up_slots", C.c_uint64),
                     ("tdo_graph_entry", C.c_uint32), ("tdo_graph_max_level", C.c_uint8),
                     ("tdo_graph_build_dist_evals", C.c_uint64)]:
        getattr(L, name).restype = rt
        getattr(L, name).argtypes = [vp]
    L.tdo_graph_export.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.POINTER(C.c_uint8),
                                   C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.POINTER(C.c_uint32),
                                   C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)]
    L.tdo_search_batch.argtypes = [vp, C.POINTER(C.c_float), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                   C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32),
                                   C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(Stats), C.c_int]
    L.tdo_sql_topk.argtypes = [C.POINTER(C.c_float), C.c_uint64, C.c_uint32, C.POINTER(C.c_float), C.c_uint32,
                               C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_double),
                               C.POINTER(C.c_uint32), C.c_int]
    L.tdo_sql_projection_distance.restype = C.c_float
    L.tdo_sql_projection_distance.argtypes = [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_uint32,
                                              C.POINTER(C.c_int)]
    L.tdo_node_write.restype = C.c_int64
    L.tdo_node_write.argtypes = [C.c_uint64, C.c_uint8, C.POINTER(C.c_uint32), C.POINTER(C.c_uint16), C.c_uint8,
                                 C.POINTER(C.c_uint32), C.POINTER(C.c_uint16), C.POINTER(C.c_uint8),
                                 C.POINTER(C.c_uint8), C.c_uint64]
    L.tdo_node_read.argtypes = [C.POINTER(C.c_uint8), C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint8),
                                C.POINTER(C.c_uint32), C.POINTER(C.c_uint16), C.POINTER(C.c_uint8),
                                C.POINTER(C.c_uint32), C.POINTER(C.c_uint16), C.POINTER(C.c_uint8)]
    _lib = L
    return L


def distance(metric: int, a, b, impl: int = DIST_AVX2) -> np.float32:
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    assert a.shape == b.shape and a.ndim == 1
    return np.float32(lib().tdo_distance(metric, _p(a, C.c_float), _p(b, C.c_float), a.size, impl))


def select_level(r: float, m: int = 16) -> int:
    return int(lib().tdo_select_level(float(r), m))


def level_randoms(n: int, seed: int) -> np.ndarray:
    """Seeded `random_value` stream in (0, 1] for insert (src/hnsw/mod.rs:1003)."""
    rng = np.random.default_rng(seed)
    return 1.0 - rng.random(n)  # (0, 1]


class OracleGraph:
    """Owns a tdo_graph.  Node id == insertion order."""

    def __init__(self, handle, dim):
        self._h = handle
        self.dim = dim

    @classmethod
    def new(cls, dim, m=16, ef_construction=100, mode=BUILD_INTENT):
        return cls(lib().tdo_graph_new(dim, m, ef_construction, mode), dim)

    @classmethod
    def build(cls, vectors, m=16, ef_construction=100, mode=BUILD_INTENT, seed=1234, row_ids=None):
        vectors = np.ascontiguousarray(vectors, dtype=np.float32)
        n, dim = vectors.shape
        g = cls.new(dim, m, ef_construction, mode)
        rid = np.arange(n, dtype=np.uint64) if row_ids is None else np.ascontiguousarray(row_ids, np.uint64)
        g.insert_batch(rid, vectors, level_randoms(n, seed))
        return g

    @classmethod
    def from_arrays(cls, arrays: dict):
        a = arrays
        n, dim = a["vectors"].shape
        nslots = int(a["up_cnt"].shape[0])
        up_adj = a["up_adj"] if nslots else np.zeros((1, MAX_UP), np.uint32)
        up_cnt = a["up_cnt"] if nslots else np.zeros(1, np.uint8)
        h = lib().tdo_graph_from_arrays(dim, n, _p(a["vectors"], C.c_float), _p(a["row_ids"], C.c_uint64),
                                        _p(a["levels"], C.c_uint8), _p(a["l0_adj"], C.c_uint32),
                                        _p(a["l0_cnt"], C.c_uint8), _p(a["up_base"], C.c_uint32),
                                        _p(up_adj, C.c_uint32), _p(up_cnt, C.c_uint8), nslots,
                                        int(a["entry"]), int(a["max_level"]))
        return cls(h, dim)

    def __del__(self):
        try:
            if getattr(self, "_h", None) and _lib is not None:
                _lib.tdo_graph_free(self._h)
                self._h = None
        except Exception:
            pass

    def insert(self, row_id, vec, random_value):
        vec = np.ascontiguousarray(vec, dtype=np.float32)
        assert vec.size == self.dim
        return lib().tdo_graph_insert(self._h, int(row_id), _p(vec, C.c_float), float(random_value))

    def insert_batch(self, row_ids, vectors, randoms):
        vectors = np.ascontiguousarray(vectors, dtype=np.float32)
        row_ids = np.ascontiguousarray(row_ids, dtype=np.uint64)
        randoms = np.ascontiguousarray(randoms, dtype=np.float64)
        assert vectors.shape[1] == self.dim
        return lib().tdo_graph_insert_batch(self._h, vectors.shape[0], _p(row_ids, C.c_uint64),
                                            _p(vectors, C.c_float), _p(randoms, C.c_double))

    @property
    def n(self):
        return int(lib().tdo_graph_n(self._h))

    @property
    def entry(self):
        return int(lib().tdo_graph_entry(self._h))

    @property
    def max_level(self):
        return int(lib().tdo_graph_max_level(self._h))

    @property
    def build_dist_evals(self):
        return int(lib().tdo_graph_build_dist_evals(self._h))

    def export(self) -> dict:
        """Flattened arrays — exactly what turdb_cuda_index_create takes."""
        L = lib()
        n = self.n
        ns = int(L.tdo_graph_n_up_slots(self._h))
        out = dict(
            vectors=np.zeros((n, self.dim), np.float32), row_ids=np.zeros(n, np.uint64),
            levels=np.zeros(n, np.uint8), l0_adj=np.full((n, MAX_L0), INVALID, np.uint32),
            l0_cnt=np.zeros(n, np.uint8), up_base=np.full(n, INVALID, np.uint32),
            up_adj=np.full((ns, MAX_UP), INVALID, np.uint32), up_cnt=np.zeros(ns, np.uint8))
        L.tdo_graph_export(self._h, _p(out["vectors"], C.c_float), _p(out["row_ids"], C.c_uint64),
                           _p(out["levels"], C.c_uint8), _p(out["l0_adj"], C.c_uint32),
                           _p(out["l0_cnt"], C.c_uint8), _p(out["up_base"], C.c_uint32),
                           _p(out["up_adj"], C.c_uint32) if ns else None,
                           _p(out["up_cnt"], C.c_uint8) if ns else None)
        out["entry"] = self.entry
        out["max_level"] = self.max_level
        return out

    def search(self, queries, k, ef, metric=L2, visible=None, n_threads=1):
        """Returns (row_ids[nq,k] u64, node_ids[nq,k] u32, dist[nq,k] f32, counts[nq] u32, stats[nq])."""
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        if queries.ndim == 1:
            queries = queries[None, :]
        nq, qd = queries.shape
        kk = max(k, 1)
        rows = np.full((nq, kk), 2**64 - 1, np.uint64)
        nodes = np.full((nq, kk), INVALID, np.uint32)
        dist = np.full((nq, kk), np.inf, np.float32)
        counts = np.zeros(nq, np.uint32)
        stats = np.zeros(nq, STATS_DTYPE)
        vis = None if visible is None else np.ascontiguousarray(visible, dtype=np.uint64)
        rc = lib().tdo_search_batch(self._h, _p(queries, C.c_float), qd, nq, k, ef, metric,
                                    _p(vis, C.c_uint64), _p(rows, C.c_uint64), _p(nodes, C.c_uint32),
                                    _p(dist, C.c_float), _p(counts, C.c_uint32),
                                    stats.ctypes.data_as(C.POINTER(Stats)), n_threads)
        if rc == 2:
            raise ValueError(f"query dimension {qd} does not match index dimension {self.dim}")
        return rows[:, :k], nodes[:, :k], dist[:, :k], counts, stats


def sql_topk(vectors, queries, limit, op=L2, offset=0, n_threads=1):
    """SQL `ORDER BY vec <op> q LIMIT limit OFFSET offset` (executor.rs:2239-2379). NULL == NaN."""
    vectors = np.ascontiguousarray(vectors, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    if queries.ndim == 1:
        queries = queries[None, :]
    n, dim = vectors.shape
    nq = queries.shape[0]
    lim = max(limit, 1)
    rows = np.full((nq, lim), 2**64 - 1, np.uint64)
    dist = np.full((nq, lim), np.inf, np.float64)
    counts = np.zeros(nq, np.uint32)
    lib().tdo_sql_topk(_p(vectors, C.c_float), n, dim, _p(queries, C.c_float), nq, limit, offset, op,
                       _p(rows, C.c_uint64), _p(dist, C.c_double), _p(counts, C.c_uint32), n_threads)
    return rows[:, :limit], dist[:, :limit], counts


def sql_projection_distance(op, a, b):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    isnull = C.c_int(0)
    v = lib().tdo_sql_projection_distance(op, _p(a, C.c_float), _p(b, C.c_float), a.size, C.byref(isnull))
    return None if isnull.value else np.float32(v)


def sq8_encode(vectors):
    """SQ8Vector::from_f32 per row -> (codes u8 [n, dim], min f32 [n], scale f32 [n])."""
    v = np.ascontiguousarray(vectors, dtype=np.float32)
    n, dim = v.shape
    L = lib()
    L.tdo_sq8_encode.restype = None
    L.tdo_sq8_encode.argtypes = [C.POINTER(C.c_float), C.c_uint32, C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    codes = np.zeros((n, dim), np.uint8)
    mn, sc = np.zeros(n, np.float32), np.zeros(n, np.float32)
    for i in range(n):
        a, b = C.c_float(0), C.c_float(0)
        L.tdo_sq8_encode(_p(v[i], C.c_float), dim, _p(codes[i], C.c_uint8), C.byref(a), C.byref(b))
        mn[i], sc[i] = a.value, b.value
    return codes, mn, sc


def sq8_decode(codes, mn, sc):
    """SQ8Vector::decode per row."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    n, dim = codes.shape
    L = lib()
    L.tdo_sq8_decode.restype = None
    L.tdo_sq8_decode.argtypes = [C.POINTER(C.c_uint8), C.c_uint32, C.c_float, C.c_float, C.POINTER(C.c_float)]
    out = np.zeros((n, dim), np.float32)
    for i in range(n):
        L.tdo_sq8_decode(_p(codes[i], C.c_uint8), dim, float(mn[i]), float(sc[i]), _p(out[i], C.c_float))
    return out


def hnsw_file_write(graph: "OracleGraph", index_id=1, table_id=1, ef_search=32, distance_fn=L2, quantization=0,
                    mode=1):
    """The bytes PersistentHnswIndex would leave on disk for this graph (see tdo_hnsw_file_write).
    mode 0 = the reference's page-fill rule verbatim (records overlap), 1 = no overlapping records.
    Returns (file bytes, pages u32[n], slots u16[n])."""
    L = lib()
    L.tdo_hnsw_file_write.restype = C.c_int64
    L.tdo_hnsw_file_write.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint16, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(C.c_uint8), C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint16)]
    n = graph.n
    need = L.tdo_hnsw_file_write(graph._h, index_id, table_id, ef_search, distance_fn, quantization, mode, None, 0, None, None)
    if need < 0:
        raise RuntimeError(f"tdo_hnsw_file_write: {need}")
    buf = np.zeros(need, np.uint8)
    pages, slots = np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.uint16)
    got = L.tdo_hnsw_file_write(graph._h, index_id, table_id, ef_search, distance_fn, quantization, mode,
                                _p(buf, C.c_uint8), need, _p(pages, C.c_uint32), _p(slots, C.c_uint16))
    if got != need:
        raise RuntimeError(f"tdo_hnsw_file_write: {got}")
    return buf.tobytes(), pages[:n], slots[:n]


def node_write(row_id, max_level, l0, upper) -> bytes:
    """l0: list[(page, slot)]; upper: list (len max_level) of list[(page, slot)]."""
    l0p = np.array([p for p, _ in l0] + [0], np.uint32)
    l0s = np.array([s for _, s in l0] + [0], np.uint16)
    upp = np.zeros((max(max_level, 1), MAX_UP), np.uint32)
    ups = np.zeros((max(max_level, 1), MAX_UP), np.uint16)
    upc = np.zeros(max(max_level, 1), np.uint8)
    for l, lst in enumerate(upper):
        upc[l] = len(lst)
        for i, (p, s) in enumerate(lst):
            upp[l, i], ups[l, i] = p, s
    buf = np.zeros(1024, np.uint8)
    w = lib().tdo_node_write(row_id, max_level, _p(l0p, C.c_uint32), _p(l0s, C.c_uint16), len(l0),
                             _p(upp, C.c_uint32), _p(ups, C.c_uint16), _p(upc, C.c_uint8),
                             _p(buf, C.c_uint8), buf.size)
    assert w >= 0
    return bytes(buf[:w])


def node_read(data: bytes):
    buf = np.frombuffer(data, np.uint8).copy()
    row_id = C.c_uint64(0)
    ml = C.c_uint8(0)
    l0c = C.c_uint8(0)
    l0p = np.zeros(MAX_L0, np.uint32)
    l0s = np.zeros(MAX_L0, np.uint16)
    upp = np.zeros((255, MAX_UP), np.uint32)
    ups = np.zeros((255, MAX_UP), np.uint16)
    upc = np.zeros(255, np.uint8)
    rc = lib().tdo_node_read(_p(buf, C.c_uint8), buf.size, C.byref(row_id), C.byref(ml),
                             _p(l0p, C.c_uint32), _p(l0s, C.c_uint16), C.byref(l0c),
                             _p(upp, C.c_uint32), _p(ups, C.c_uint16), _p(upc, C.c_uint8))
    if rc:
        raise ValueError(f"node decode error {rc}")
    l0 = [(int(l0p[i]), int(l0s[i])) for i in range(l0c.value)]
    upper = [[(int(upp[l, i]), int(ups[l, i])) for i in range(min(int(upc[l]), MAX_UP))]
             for l in range(ml.value)]
    return dict(row_id=row_id.value, max_level=ml.value, l0=l0, upper=upper)
