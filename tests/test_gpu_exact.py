"""Exact path (tensor-core BF16 pass + FP32 rerank) vs the oracle's exact scan and vs the HNSW distance
contract.  Tolerance: the north_star's 1e-5 relative FP32 bound on distances; id sets must match except where
the boundary distance is tied within that tolerance."""
import numpy as np
import pytest

from oracle import binding as ob
from turdb_b200 import datasets as ds
from turdb_b200.hnsw import CudaHnswIndex, DistanceFunction

pytestmark = pytest.mark.gpu
REL_TOL = 1e-5


def exact_reference(x, q, k, metric):
    """k smallest oracle distances (AVX2 lane order), ties by node id."""
    out_ids = np.zeros((len(q), k), np.uint32)
    out_d = np.zeros((len(q), k), np.float32)
    for i, qq in enumerate(q):
        d = np.array([ob.distance(metric, qq, v) for v in x], np.float32)
        order = np.lexsort((np.arange(len(x)), d))[:k]
        out_ids[i], out_d[i] = order, d[order]
    return out_ids, out_d


def flat_graph(x):
    n = len(x)
    return dict(vectors=x, row_ids=np.arange(n, dtype=np.uint64) + 5, levels=np.zeros(n, np.uint8),
                l0_adj=np.full((n, 32), 0xFFFFFFFF, np.uint32), l0_cnt=np.zeros(n, np.uint8),
                up_base=np.full(n, 0xFFFFFFFF, np.uint32), up_adj=np.zeros((0, 16), np.uint32),
                up_cnt=np.zeros(0, np.uint8), entry=0, max_level=0)


@pytest.mark.parametrize("n,dim,nq,k", [(3000, 128, 40, 10), (5000, 384, 200, 10), (2500, 100, 17, 5), (700, 64, 130, 100),
                                        (2000, 768, 150, 10), (1500, 520, 33, 7)])  # > 512 dims: streamed query block
def test_exact_topk_matches_reference(gpu_required, n, dim, nq, k):
    x = ds.gaussian_latent(n, dim, seed=n, normalise=False)
    q = ds.gaussian_latent(nq, dim, seed=n + 1, normalise=False)
    idx = CudaHnswIndex.from_graph(flat_graph(x))
    for metric in (ob.L2, ob.COSINE, ob.IP):
        rows, nodes, dist, cnt = idx.bruteforce_topk(q, k, DistanceFunction(metric), rerank_factor=4)
        ref_ids, ref_d = exact_reference(x, q[:16], k, metric)
        assert (cnt == min(k, n)).all()
        assert np.array_equal(rows, nodes.astype(np.uint64) + 5)
        for i in range(16):
            # reranked distances are the reference's own arithmetic: equal bit for bit where ids agree
            same = nodes[i] == ref_ids[i]
            assert np.array_equal(dist[i][same].view(np.uint32), ref_d[i][same].view(np.uint32))
            if not same.all():  # any id difference must sit on a (near-)tie
                assert np.allclose(dist[i], ref_d[i], rtol=REL_TOL, atol=1e-6), (metric, i)
        assert (np.diff(dist, axis=1) >= 0).all()
    idx.close()


def test_exact_sql_known_answers(gpu_required):
    """The reference's SQL k-NN answers (tests/hnsw_integration.rs:220-276) through the exact path."""
    x = np.array([[.1] * 4, [.5] * 4, [.9] * 4], np.float32)
    idx = CudaHnswIndex.from_graph(flat_graph(x))
    rows, nodes, dist, cnt = idx.bruteforce_topk(np.array([[.1] * 4], np.float32), 2, DistanceFunction.L2)
    assert nodes[0].tolist() == [0, 1] and cnt[0] == 2
    idx.close()
    x20 = np.array([[i / 20.0] * 4 for i in range(20)], np.float32)
    idx = CudaHnswIndex.from_graph(flat_graph(x20))
    rows, nodes, dist, cnt = idx.bruteforce_topk(np.array([[.5] * 4], np.float32), 3, DistanceFunction.L2)
    assert 8 <= nodes[0][0] <= 12 and cnt[0] == 3
    # agrees with the oracle's SQL TopK scan on the id set (SQL reports sqrt of the same squared distance)
    srows, sdist, _ = ob.sql_topk(x20, np.array([.5] * 4, np.float32), 3)
    assert set(srows[0].tolist()) == set(nodes[0].tolist())
    assert np.allclose(np.sqrt(dist[0].astype(np.float64)), sdist[0], rtol=1e-6)
    idx.close()


def test_exact_recall_against_sql_scan(gpu_required):
    x = ds.gaussian_latent(20000, 128, seed=3, normalise=True)
    q = ds.gaussian_latent(64, 128, seed=4, normalise=True)
    idx = CudaHnswIndex.from_graph(flat_graph(x))
    rows, nodes, dist, cnt = idx.bruteforce_topk(q, 10, DistanceFunction.Cosine)
    srows, _, _ = ob.sql_topk(x, q, 10, op=ob.COSINE, n_threads=8)
    rec = np.mean([len(set(nodes[i].tolist()) & set(srows[i].tolist())) / 10 for i in range(len(q))])
    assert rec >= 0.999
    idx.close()
