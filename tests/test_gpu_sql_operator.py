"""SQL vector-scan operator (csrc/sql_topk.inl) vs the oracle's restatement of TopKExec
(src/sql/executor.rs:2239-2392, :169-212), plus the reference's own three k-NN answers
(tests/hnsw_integration.rs:220-276) run through the GPU operator."""
import numpy as np
import pytest

from oracle import binding as ob
from turdb_b200 import datasets as ds
from turdb_b200.hnsw import CudaHnswIndex
from turdb_b200.sql_operator import VectorOp, VectorScanBatch, VectorTopKExec

pytestmark = pytest.mark.gpu


def table_index(x, row_ids=None):
    """A table without an HNSW graph: the exact scan only needs the arena (entry = none)."""
    n = x.shape[0]
    g = dict(vectors=x, row_ids=np.arange(n, dtype=np.uint64) if row_ids is None else row_ids,
             levels=np.zeros(n, np.uint8), l0_adj=np.full((n, 32), 0xFFFFFFFF, np.uint32), l0_cnt=np.zeros(n, np.uint8),
             up_base=np.full(n, 0xFFFFFFFF, np.uint32), up_adj=np.zeros((0, 16), np.uint32), up_cnt=np.zeros(0, np.uint8),
             entry=0 if n else 0xFFFFFFFF, max_level=0)
    return CudaHnswIndex.from_graph(g)


def test_reference_knn_answers(gpu_required):
    # knn_search_returns_nearest_neighbors, tests/hnsw_integration.rs:220-236
    ids = np.array([1, 2, 3], np.uint64)
    x = np.array([[0.1] * 4, [0.5] * 4, [0.9] * 4], np.float32)
    ex = VectorTopKExec(table_index(x, ids), VectorOp.L2Distance, "[0.1, 0.1, 0.1, 0.1]", limit=2)
    ex.open()
    got = [ex.next(), ex.next(), ex.next()]
    assert [g[0] for g in got[:2]] == [1, 2] and got[2] is None
    # knn_search_after_delete_excludes_deleted, :238-256 (the deleted row is gone from the table)
    ex = VectorTopKExec(table_index(x[1:], ids[1:]), VectorOp.L2Distance, "[0.1, 0.1, 0.1, 0.1]", limit=2)
    ex.open()
    assert [ex.next()[0], ex.next()[0]] == [2, 3]
    # insert_many_vectors_and_search, :258-276
    x20 = np.array([[i / 20.0] * 4 for i in range(20)], np.float32)
    ex = VectorTopKExec(table_index(x20), VectorOp.L2Distance, "[0.5, 0.5, 0.5, 0.5]", limit=3)
    ex.open()
    first = ex.next()
    assert 8 <= first[0] <= 12 and ex.next() is not None and ex.next() is not None and ex.next() is None


@pytest.mark.parametrize("op,oop", [(VectorOp.L2Distance, ob.L2), (VectorOp.CosineDistance, ob.COSINE)])
@pytest.mark.parametrize("limit,offset", [(10, 0), (5, 7), (1, 0), (64, 3)])
def test_matches_topk_executor(gpu_required, op, oop, limit, offset):
    x = ds.gaussian_latent(20_000, 96, seed=31)
    q = ds.gaussian_latent(64, 96, seed=32)
    rows, keys, counts = VectorScanBatch(table_index(x, np.arange(20_000, dtype=np.uint64) * 2 + 1), op, limit, offset).execute(q)
    o_rows, o_keys, o_counts = ob.sql_topk(x, q, limit, op=oop, offset=offset, n_threads=8)
    assert np.array_equal(counts, o_counts)
    assert np.array_equal(rows, o_rows.astype(np.uint64) * 2 + 1)
    assert np.array_equal(keys.view(np.uint64), o_keys.view(np.uint64))  # the f64 key bit for bit


def test_small_table_and_ties(gpu_required):
    # fewer rows than limit + offset; integer-valued vectors give exact key ties (order by scan position)
    x = np.array([[0, 0], [3, 4], [3, 4], [6, 8], [0, 5]], np.float32)
    rows, keys, counts = VectorScanBatch(table_index(x), VectorOp.L2Distance, limit=10).execute(np.zeros((1, 2), np.float32))
    assert counts[0] == 5 and keys[0, :5].tolist() == [0.0, 5.0, 5.0, 5.0, 10.0]
    assert rows[0, 0] == 0 and rows[0, 4] == 3 and sorted(rows[0, 1:4].tolist()) == [1, 2, 4]
    rows, keys, counts = VectorScanBatch(table_index(x), VectorOp.L2Distance, limit=2, offset=4).execute(np.zeros((1, 2), np.float32))
    assert counts[0] == 1 and rows[0, 0] == 3
    # cosine: a zero-norm row has a NULL key and sorts last (executor.rs:201-206)
    rows, keys, counts = VectorScanBatch(table_index(x), VectorOp.CosineDistance, limit=5).execute(np.array([[3, 4]], np.float32))
    assert counts[0] == 5 and rows[0, 4] == 0 and np.isnan(keys[0, 4]) and keys[0, 0] == 0.0


def test_inner_product_is_rejected_like_the_reference_evaluates_it(gpu_required):
    x = ds.gaussian_latent(100, 8, seed=1)
    with pytest.raises(Exception, match="NULL for every row"):
        VectorScanBatch(table_index(x), VectorOp.InnerProduct, limit=3).execute(x[:1])


def test_index_backed_scan_agrees_on_a_good_graph(gpu_required, small_graph):
    g, arrays = small_graph
    q = ds.gaussian_latent(200, 128, seed=2)
    idx = CudaHnswIndex.from_graph(arrays)
    exact = VectorScanBatch(idx, VectorOp.L2Distance, 10).execute(q)
    approx = VectorScanBatch(idx, VectorOp.L2Distance, 10, use_index=True, ef_search=128).execute(q)
    agree = np.mean([len(set(exact[0][i].tolist()) & set(approx[0][i].tolist())) / 10 for i in range(200)])
    assert agree >= 0.97
    same = exact[0] == approx[0]
    assert np.array_equal(exact[1][same].view(np.uint64), approx[1][same].view(np.uint64))


def test_tied_keys_come_back_in_the_reference_heap_order(gpu_required):
    """Integer-valued vectors: many exactly equal keys inside and across the LIMIT boundary.  Which of them survive and
    in which order follows from the reference's heap procedure (executor.rs:2248-2378) — replayed on the device."""
    rng = np.random.default_rng(3)
    x = rng.integers(0, 4, (5000, 6)).astype(np.float32)
    q = rng.integers(0, 4, (40, 6)).astype(np.float32)
    idx = table_index(x)
    for op, oop in ((VectorOp.L2Distance, ob.L2), (VectorOp.CosineDistance, ob.COSINE)):
        for limit, offset in ((10, 0), (7, 5), (1, 0), (100, 20), (300, 212)):
            rows, keys, counts = VectorScanBatch(idx, op, limit, offset).execute(q)
            o_rows, o_keys, o_counts = ob.sql_topk(x, q, limit, op=oop, offset=offset, n_threads=8)
            assert np.array_equal(counts, o_counts)
            ok = ~np.isnan(o_keys)
            assert np.array_equal(np.isnan(keys), np.isnan(o_keys))
            assert np.array_equal(keys[ok].view(np.uint64), o_keys[ok].view(np.uint64))
            nn = np.array([not np.isnan(o_keys[i]).any() for i in range(len(q))])  # NULL keys: order unspecified
            assert np.array_equal(rows[nn], o_rows[nn].astype(np.uint64)), (op, limit, offset)
    idx.close()


def test_thousands_of_rows_tying_with_the_limit_th_key(gpu_required):
    """4000 duplicates of the nearest row: the filter's buffers overflow and the statement is redone by the kernel that
    runs the reference's loop over every row — same rows, same order, no error."""
    x = ds.gaussian_latent(9000, 32, seed=41)
    x[500:4500] = x[499]
    q = np.concatenate([x[499:500] + 1e-3, ds.gaussian_latent(5, 32, seed=42)])
    idx = table_index(x)
    for limit, offset in ((10, 0), (25, 10)):
        rows, keys, counts = VectorScanBatch(idx, VectorOp.L2Distance, limit, offset).execute(q)
        o_rows, o_keys, o_counts = ob.sql_topk(x, q, limit, op=ob.L2, offset=offset, n_threads=8)
        assert np.array_equal(counts, o_counts) and np.array_equal(rows, o_rows.astype(np.uint64))
        assert np.array_equal(keys.view(np.uint64), o_keys.view(np.uint64))
    idx.close()


@pytest.mark.parametrize("op,proj", [(VectorOp.L2Distance, VectorOp.L2Distance), (VectorOp.CosineDistance, VectorOp.CosineDistance),
                                     (VectorOp.L2Distance, VectorOp.InnerProduct)])
def test_projected_distance_matches_predicate_eval(gpu_required, op, proj):
    """SELECT id, vec <proj> q AS d ... ORDER BY vec <op> q LIMIT 8: the projected value is the f32 sequential arithmetic of
    src/sql/predicate.rs:1634-1688 (sqrt L2, cosine with NULL on a zero norm, +dot for <#>), bit for bit."""
    x = ds.gaussian_latent(3000, 100, seed=51)
    x[7] = 0.0  # zero norm: cosine projects NULL
    q = np.concatenate([ds.gaussian_latent(30, 100, seed=52), x[7:8] + 1e-3])
    sb = VectorScanBatch(table_index(x), op, 8, project=proj)
    rows, keys, counts = sb.execute(q)
    for i in range(len(q)):
        for j in range(int(counts[i])):
            want = ob.sql_projection_distance(int(proj), x[int(rows[i, j])], q[i])
            got = sb.projected[i, j]
            if want is None:
                assert np.isnan(got)
            else:
                assert np.float64(want) == got, (i, j, want, got)
