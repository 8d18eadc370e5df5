"""The insert path on the device (turdb_cuda_index_build, csrc/graph_insert.inl) vs the oracle's restatement of
insert_with_callback (src/hnsw/mod.rs:999-1084, src/hnsw/operations.rs:76-233): with max_batch = 1 the two are the
same sequential procedure, so the GRAPHS must be equal array for array; larger steps are the batched construction,
checked for search quality against the sequential graph."""
import numpy as np
import pytest

from oracle import binding as ob
from turdb_b200 import datasets as ds
from turdb_b200.hnsw import CudaHnswIndex, DistanceFunction

pytestmark = pytest.mark.gpu


def graphs_equal(a, b):
    assert a["entry"] == b["entry"] and a["max_level"] == b["max_level"]
    assert np.array_equal(a["levels"], b["levels"]) and np.array_equal(a["up_base"], b["up_base"])
    assert np.array_equal(a["l0_cnt"], b["l0_cnt"]), np.where(a["l0_cnt"] != b["l0_cnt"])[0][:10]
    assert np.array_equal(a["up_cnt"], b["up_cnt"])
    for i in np.where((a["l0_adj"] != b["l0_adj"]).any(axis=1))[0][:5]:
        raise AssertionError(f"level-0 list of node {i}: {a['l0_adj'][i][:a['l0_cnt'][i]]} vs {b['l0_adj'][i][:b['l0_cnt'][i]]}")
    assert np.array_equal(a["up_adj"], b["up_adj"])
    assert np.array_equal(a["row_ids"], b["row_ids"])


@pytest.mark.parametrize("n,dim,mode", [(3000, 32, ob.BUILD_INTENT), (3000, 32, ob.BUILD_VERBATIM), (2500, 100, ob.BUILD_INTENT),
                                        (1200, 384, ob.BUILD_INTENT), (40, 8, ob.BUILD_INTENT)])
def test_sequential_build_equals_the_oracle_graph(gpu_required, n, dim, mode):
    x = ds.gaussian_latent(n, dim, seed=n + dim)
    rnd = ob.level_randoms(n, 77)
    rid = np.arange(n, dtype=np.uint64) * 3 + 1
    og = ob.OracleGraph.new(dim, 16, 100, mode)
    og.insert_batch(rid, x, rnd)
    idx = CudaHnswIndex.build(x, rid, rnd, m=16, ef_construction=100, mode=mode, max_batch=1)
    graphs_equal(idx.export_graph(), og.export())
    # and the built index searches like any uploaded one
    q = ds.gaussian_latent(50, dim, seed=5)
    gpu = idx.search_batch(q, 10, 64, DistanceFunction.L2)
    cpu = og.search(q, 10, 64, ob.L2)
    assert np.array_equal(gpu[1], cpu[1]) and np.array_equal(gpu[2].view(np.uint32), cpu[2].view(np.uint32))
    idx.close()


def test_level_above_max_level_links_one_way(gpu_required):
    """Appendix B.4: a level draw above the current max_level -> the new entry has a one-way upper link to the old one."""
    x = ds.gaussian_latent(30, 16, seed=3)
    rnd = np.full(30, 0.9)
    rnd[0] = 0.05   # level 1
    rnd[20] = 1e-4  # level 3 > max_level 1
    og = ob.OracleGraph.new(16, 16, 100, ob.BUILD_INTENT)
    og.insert_batch(np.arange(30, dtype=np.uint64), x, rnd)
    idx = CudaHnswIndex.build(x, None, rnd, max_batch=1)
    g, o = idx.export_graph(), og.export()
    graphs_equal(g, o)
    assert g["entry"] == 20 and g["max_level"] == 3
    idx.close()


@pytest.mark.parametrize("gen,kw,dim", [("gaussian_latent", dict(latent=16), 64), ("clustered", dict(sigma=0.1, corpus_n=20000), 64)])
def test_batched_build_matches_sequential_quality(gpu_required, gen, kw, dim):
    n = 20000
    x = ds.make(gen, n, dim, seed=1, **kw)
    q = ds.make(gen, 300, dim, seed=2, **kw)
    rnd = ob.level_randoms(n, 9)
    og = ob.OracleGraph.new(dim, 16, 100, ob.BUILD_INTENT)
    og.insert_batch(np.arange(n, dtype=np.uint64), x, rnd)
    d = ((q[:, None, :] - x[None, :, :]) ** 2).sum(-1)
    gt = np.argsort(d, axis=1, kind="stable")[:, :10]

    def recall(nodes):
        return float(np.mean([len(set(nodes[i].tolist()) & set(gt[i].tolist())) / 10 for i in range(len(q))]))
    seq = og.search(q, 10, 64, ob.L2, n_threads=8)
    idx = CudaHnswIndex.build(x, None, rnd, max_batch=1024)
    bat = idx.search_batch(q, 10, 64, DistanceFunction.L2)
    g = idx.export_graph()
    assert (g["l0_cnt"] > 0).all(), "every node is linked"
    assert recall(bat[1]) >= recall(seq[1]) - 0.02, (recall(bat[1]), recall(seq[1]))
    # the batched graph is still a graph the oracle traverses identically
    cpu = ob.OracleGraph.from_arrays(g).search(q, 10, 64, ob.L2, n_threads=8)
    assert np.array_equal(bat[1], cpu[1]) and np.array_equal(bat[2].view(np.uint32), cpu[2].view(np.uint32))
    idx.close()
