"""The SQL vector-scan operator: `SELECT ... ORDER BY vec <op> '[...]' LIMIT k OFFSET o` on the GPU.

Host-side mirror of the reference's TopK executor for that statement shape (kahflane/TurDB):
  - `PhysicalOperator::TopKExec{input, order_by, limit, offset}`  src/sql/planner/physical.rs:229
  - `DynamicExecutor::TopK`, pulled with open / next / close       src/sql/executor.rs:346-350, 2239-2392
  - sort-key arithmetic of `<->` and `<=>`                         src/sql/executor.rs:169-212
  - `<#>` is not a sort key there (every row evaluates to NULL)    src/sql/executor.rs:241
A planner replaces TopKExec(scan) by this operator when order_by[0] is `Column <op> literal`
(src/sql/ast.rs:907-909); many statements over one table go through ONE launch (`VectorScanBatch`,
BASELINE.json config 5's batch path).  All arithmetic is below the C ABI (csrc/sql_topk.inl).
"""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np

from . import _lib
from .hnsw import INVALID_ROW, CudaHnswIndex, _check, _ptr


class VectorOp(enum.IntEnum):
    L2Distance = 0       # `<->`  BinaryOperator::VectorL2Distance
    CosineDistance = 1   # `<=>`  BinaryOperator::VectorCosineDistance
    InnerProduct = 2     # `<#>`  BinaryOperator::VectorInnerProduct (NULL as a sort key in the reference)


def parse_vector_literal(text: str) -> np.ndarray:
    """'[0.1, 0.2]' -> f32 vector; value_to_vec_standalone, src/sql/executor.rs:245-261."""
    t = text.strip()
    if not (t.startswith("[") and t.endswith("]")):
        raise ValueError("vector literal must be bracketed")
    return np.array([np.float32(x.strip()) for x in t[1:-1].split(",")], dtype=np.float32)


class VectorScanBatch:
    """nq statements `SELECT [vec <proj_op> literal_i] ... ORDER BY vec <op> literal_i LIMIT limit OFFSET offset` against
    one table.  `project` = the operator whose value the statement projects (predicate.rs:1634-1688), or None."""

    def __init__(self, index: CudaHnswIndex, op: VectorOp, limit: int, offset: int = 0, use_index: bool = False,
                 ef_search: int = 0, project: VectorOp | None = None):
        self.index, self.op, self.limit, self.offset = index, VectorOp(op), int(limit), int(offset)
        self.use_index, self.ef_search = bool(use_index), int(ef_search)
        self.project = None if project is None else VectorOp(project)
        self.projected = None

    def execute(self, literals) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        """-> (row_ids u64 [nq, limit], keys f64 [nq, limit] (NaN = NULL), counts u32 [nq]); the projected values, when
        asked for, are left in `self.projected` (f64 [nq, limit], NaN = NULL)."""
        q = np.ascontiguousarray(literals, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        nq, qd = q.shape
        lim = max(self.limit, 1)
        rows = np.full((nq, lim), INVALID_ROW, np.uint64)
        keys = np.full((nq, lim), np.nan, np.float64)
        proj = np.full((nq, lim), np.nan, np.float64) if self.project is not None else None
        counts = np.zeros(nq, np.uint32)
        _check(_lib.load().turdb_cuda_sql_topk_batch(self.index._h, _ptr(q, C.c_float), qd, nq, self.limit, self.offset,
                                                     int(self.op), int(self.project or 0), 1 if self.use_index else 0,
                                                     self.ef_search, _ptr(rows, C.c_uint64), _ptr(keys, C.c_double),
                                                     _ptr(proj, C.c_double) if proj is not None else None,
                                                     _ptr(counts, C.c_uint32)))
        self.projected = None if proj is None else proj[:, :self.limit]
        return rows[:, :self.limit], keys[:, :self.limit], counts

    def execute_device(self, d_literals: int, nq: int, d_rows: int, d_keys: int, d_counts: int, stream: int = 0,
                       d_proj: int = 0):
        """Device-resident form: raw device addresses, enqueued on `stream` (no host synchronisation)."""
        _check(_lib.load().turdb_cuda_sql_topk_batch_device(self.index._h, d_literals, self.index.dim, nq, self.limit, self.offset,
                                                            int(self.op), int(self.project or 0), 1 if self.use_index else 0,
                                                            self.ef_search, d_rows, d_keys, d_proj or None, d_counts,
                                                            stream or None))


class VectorTopKExec:
    """One statement with the executor protocol of the reference: open() / next() -> (row_id, key) | None / close()."""

    def __init__(self, index: CudaHnswIndex, op: VectorOp, literal, limit: int, offset: int = 0, use_index: bool = False,
                 ef_search: int = 0):
        self._batch = VectorScanBatch(index, op, limit, offset, use_index, ef_search)
        self._literal = parse_vector_literal(literal) if isinstance(literal, str) else np.asarray(literal, np.float32)
        self._result = None
        self._iter = 0

    def open(self):
        self._result, self._iter = None, 0

    def next(self):
        if self._result is None:  # `computed` flag of TopKState, executor.rs:2240
            rows, keys, counts = self._batch.execute(self._literal)
            self._result = [(int(rows[0, i]), float(keys[0, i])) for i in range(int(counts[0]))]
        if self._iter < len(self._result):
            self._iter += 1
            return self._result[self._iter - 1]
        return None

    def close(self):
        self._result = None


# ---- the planner rule (mirror of where the reference turns Limit(Sort(..)) into TopKExec, -------------------------
# src/sql/planner/convert.rs:348-397): a statement whose ORDER BY is `column <op> vector-literal` followed by LIMIT
# [OFFSET] is rewritten to the vector-scan operator; anything else keeps the stock plan (None).
import re as _re

_VECTOR_TOPK = _re.compile(
    r"""^\s*SELECT\s+(?P<proj>.+?)\s+FROM\s+(?P<table>[A-Za-z_][\w.]*)\s+
        ORDER\s+BY\s+(?P<col>[A-Za-z_][\w.]*)\s*(?P<op><->|<=>|<\#>)\s*'(?P<lit>\[[^']*\])'\s*(?P<dir>ASC|DESC)?\s+
        LIMIT\s+(?P<limit>\d+)(?:\s+OFFSET\s+(?P<offset>\d+))?\s*;?\s*$""",
    _re.IGNORECASE | _re.VERBOSE | _re.DOTALL)

_OPS = {"<->": VectorOp.L2Distance, "<=>": VectorOp.CosineDistance, "<#>": VectorOp.InnerProduct}


def plan_vector_topk(sql: str):
    """`SELECT .. FROM t ORDER BY col <op> '[..]' LIMIT k [OFFSET o]` -> dict(table, column, op, literal, limit, offset,
    projection) for VectorTopKExec, or None when the rule does not apply: a WHERE / JOIN / GROUP BY (the scan below the
    sort is not the bare table), a descending order (the k FARTHEST rows), a second sort key, or `<#>` — a NULL sort key
    in the reference (src/sql/executor.rs:241), which must keep the stock TopK to stay result-compatible."""
    m = _VECTOR_TOPK.match(sql)
    if not m:
        return None
    if (m.group("dir") or "ASC").upper() == "DESC":
        return None
    op = _OPS[m.group("op")]
    if op == VectorOp.InnerProduct:
        return None
    try:
        lit = parse_vector_literal(m.group("lit"))
    except ValueError:
        return None
    return dict(table=m.group("table"), column=m.group("col"), op=op, literal=lit, limit=int(m.group("limit")),
                offset=int(m.group("offset") or 0), projection=[c.strip() for c in m.group("proj").split(",")])
