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


def lists_equal_up_to_ties(x, owner_rows, la, lb, what):
    """Two adjacency tables must agree entry for entry, except that two neighbours at EXACTLY the same distance from the
    owner may swap places: Candidate's order is distance-only (search.rs:94-115), so which of two equidistant results the
    reference's heaps emit first follows std BinaryHeap internals.  Returns the number of such lists."""
    swapped = 0
    for r in np.where((la != lb).any(axis=1))[0]:
        o = int(owner_rows[r])
        pos = np.where(la[r] != lb[r])[0]
        assert sorted(la[r].tolist()) == sorted(lb[r].tolist()), f"{what} of node {o}: {la[r]} vs {lb[r]}"
        for p_ in pos:
            da, db = ob.distance(ob.L2, x[o], x[int(la[r][p_])]), ob.distance(ob.L2, x[o], x[int(lb[r][p_])])
            assert np.float32(da).view(np.uint32) == np.float32(db).view(np.uint32), \
                f"{what} of node {o} differs at {p_} without a distance tie: {la[r]} vs {lb[r]}"
        swapped += 1
    return swapped


def graphs_equal(a, b, x):
    assert a["entry"] == b["entry"] and a["max_level"] == b["max_level"]
    assert np.array_equal(a["levels"], b["levels"]) and np.array_equal(a["up_base"], b["up_base"])
    assert np.array_equal(a["l0_cnt"], b["l0_cnt"]), np.where(a["l0_cnt"] != b["l0_cnt"])[0][:10]
    assert np.array_equal(a["up_cnt"], b["up_cnt"])
    n = len(a["levels"])
    tied = lists_equal_up_to_ties(x, np.arange(n), a["l0_adj"], b["l0_adj"], "level-0 list")
    owner = np.repeat(np.arange(n), a["levels"])  # upper slot -> owning node
    tied += lists_equal_up_to_ties(x, owner, a["up_adj"], b["up_adj"], "upper list")
    assert tied <= max(1, n // 500), f"{tied} lists differ by tie order"
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
    graphs_equal(idx.export_graph(), og.export(), x)
    # and the built index searches like any uploaded one
    q = ds.gaussian_latent(50, dim, seed=5)
    gpu = idx.search_batch(q, 10, 64, DistanceFunction.L2)
    cpu = ob.OracleGraph.from_arrays(idx.export_graph()).search(q, 10, 64, ob.L2)
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
    graphs_equal(g, o, x)
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


def test_capacity_caps_on_the_device(gpu_required):
    """The reference's node capacity tests (tests/hnsw_integration.rs:79-112: MAX + 5 additions leave MAX = 32 / 16) through
    the device insert path (verbatim mode, sequential) and through index_create's clamp of an over-long count."""
    n = 38
    x = np.zeros((n, 4), np.float32)
    x[:, 0] = np.arange(n) * 1e-3
    rnd = np.full(n, 0.05)  # level 1 for every node (select_level, M = 16)
    og = ob.OracleGraph.new(4, 16, 100, ob.BUILD_VERBATIM)
    og.insert_batch(np.arange(n, dtype=np.uint64), x, rnd)
    idx = CudaHnswIndex.build(x, None, rnd, mode=ob.BUILD_VERBATIM, max_batch=1)
    g = idx.export_graph()
    graphs_equal(g, og.export(), x)
    assert g["l0_cnt"][0] == 32 and g["l0_adj"][0].tolist() == list(range(1, 33))
    assert g["up_cnt"][g["up_base"][0]] == 16 and g["up_adj"][g["up_base"][0]].tolist() == list(range(1, 17))
    idx.close()
    # an uploaded row whose count claims more than the row holds is clamped to 32 entries
    a = og.export()
    a["l0_cnt"] = a["l0_cnt"].copy()
    a["l0_cnt"][0] = 37
    up = CudaHnswIndex.from_graph(a)
    assert up.export_graph(with_vectors=False)["l0_cnt"][0] == 32
    up.close()
