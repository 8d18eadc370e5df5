"""search_filtered on the device (search.rs:352-398, mod.rs:1176-1273) vs the oracle."""
import os

import numpy as np
import pytest

from oracle import binding as ob
from turdb_b200 import datasets as ds
from turdb_b200.hnsw import CudaHnswIndex, DistanceFunction, HnswSearchContext, visibility_bitmap

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hnsw_small.npz")


def check(idx, g, q, k, ef, mask, metric=ob.L2):
    words = visibility_bitmap(mask)
    gpu = idx.search_batch(q, k, ef, DistanceFunction(metric), visible=words)
    cpu = g.search(q, k, ef, metric, visible=words, n_threads=8)
    assert np.array_equal(gpu[3], cpu[3]), "counts differ"
    for i in range(len(q)):
        c = cpu[3][i]
        assert np.array_equal(gpu[1][i][:c], cpu[1][i][:c]), (i, gpu[1][i][:c], cpu[1][i][:c])
        assert np.array_equal(gpu[2][i][:c].view(np.uint32), cpu[2][i][:c].view(np.uint32))
        assert all(mask[n] for n in gpu[1][i][:c])
    for f in ("n_dist", "n_expanded"):
        assert np.array_equal(gpu[4][f], cpu[4][f]), f
    return gpu, cpu


@pytest.mark.parametrize("selectivity", [1.0, 0.9, 0.5, 0.1, 0.02, 0.0])
def test_filtered_matches_oracle(gpu_required, small_graph, selectivity):
    g, arrays = small_graph
    rng = np.random.default_rng(int(selectivity * 100) + 1)
    mask = rng.random(g.n) < selectivity
    q = ds.gaussian_latent(120, 128, seed=31)
    idx = CudaHnswIndex.from_graph(arrays)
    gpu, cpu = check(idx, g, q, 10, 64, mask)
    if selectivity == 0.0:
        assert (gpu[3] == 0).all()
    if selectivity == 1.0:  # everything visible: same answer as search()
        plain = idx.search_batch(q, 10, 64)
        assert np.array_equal(plain[1], gpu[1])
    idx.close()


def test_filtered_golden_and_reference_api(gpu_required):
    z = np.load(GOLD)
    arrays = {k: z[k] for k in ("vectors", "row_ids", "levels", "l0_adj", "l0_cnt", "up_base", "up_adj", "up_cnt")}
    arrays["entry"], arrays["max_level"] = int(z["entry"]), int(z["max_level"])
    idx = CudaHnswIndex.from_graph(arrays)
    words = visibility_bitmap(z["visible_mask"])
    rows, nodes, dist, cnt, _ = idx.search_batch(z["queries"], 10, 40, DistanceFunction.L2, visible=words)
    assert np.array_equal(nodes, z["filtered_nodes"]) and np.array_equal(cnt, z["filtered_counts"])
    assert np.array_equal(dist.view(np.uint32), z["filtered_dist"].view(np.uint32))
    # the reference-shaped entry point: is_visible(row_id) closure
    vis_rows = set(int(r) for r, v in zip(arrays["row_ids"], z["visible_mask"]) if v)
    hits = idx.search_filtered(z["queries"][0], 10, HnswSearchContext(40), lambda r: r in vis_rows,
                               row_ids=arrays["row_ids"])
    assert [h.node_id for h in hits] == z["filtered_nodes"][0][:len(hits)].tolist()
    idx.close()


def test_filtered_small_ef_and_cosine(gpu_required):
    x = ds.gaussian_latent(4000, 48, seed=41)
    q = ds.gaussian_latent(60, 48, seed=42)
    g = ob.OracleGraph.build(x, seed=6)
    idx = CudaHnswIndex.from_graph(g.export())
    mask = np.random.default_rng(7).random(4000) < 0.3
    check(idx, g, q, 5, 8, mask, ob.COSINE)
    check(idx, g, q, 20, 16, mask, ob.IP)
    idx.close()
