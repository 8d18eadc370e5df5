"""The oracle reproduces the committed golden fixture (tests/golden/make_golden.py) bit for bit."""
import os

import numpy as np

from oracle import binding as ob

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hnsw_small.npz")


def load():
    z = np.load(GOLD)
    arrays = {k: z[k] for k in ("vectors", "row_ids", "levels", "l0_adj", "l0_cnt", "up_base", "up_adj", "up_cnt")}
    arrays["entry"] = int(z["entry"])
    arrays["max_level"] = int(z["max_level"])
    return z, arrays


def test_oracle_search_matches_golden():
    z, arrays = load()
    g = ob.OracleGraph.from_arrays(arrays)
    for metric, name in ((ob.L2, "l2"), (ob.COSINE, "cosine"), (ob.IP, "ip")):
        rows, nodes, dist, cnt, st = g.search(z["queries"], 10, 40, metric, n_threads=4)
        assert np.array_equal(nodes, z[f"{name}_nodes"])
        assert np.array_equal(dist.view(np.uint32), z[f"{name}_dist"].view(np.uint32))
        assert np.array_equal(cnt, z[f"{name}_counts"]) and np.array_equal(st, z[f"{name}_stats"])
        assert np.array_equal(rows, z[f"{name}_rows"])


def test_oracle_rebuild_matches_golden_graph():
    """The insert path is deterministic: rebuilding from the vectors gives the same adjacency."""
    z, arrays = load()
    g = ob.OracleGraph.build(arrays["vectors"], m=16, ef_construction=100, mode=ob.BUILD_INTENT, seed=103,
                             row_ids=arrays["row_ids"])
    a = g.export()
    for k in ("levels", "l0_adj", "l0_cnt", "up_base", "up_adj", "up_cnt"):
        assert np.array_equal(a[k], arrays[k]), k
    assert a["entry"] == arrays["entry"] and a["max_level"] == arrays["max_level"]


def test_oracle_filtered_and_sql_match_golden():
    z, arrays = load()
    g = ob.OracleGraph.from_arrays(arrays)
    vis = z["visible_mask"]
    n = len(vis)
    words = np.packbits(np.pad(vis, (0, (-n) % 64)).reshape(-1, 64), axis=1, bitorder="little").view(np.uint64).ravel()
    rows, nodes, dist, cnt, _ = g.search(z["queries"], 10, 40, ob.L2, visible=words)
    assert np.array_equal(nodes, z["filtered_nodes"]) and np.array_equal(cnt, z["filtered_counts"])
    erows, edist, _ = ob.sql_topk(arrays["vectors"], z["queries"], 10, op=ob.L2, n_threads=4)
    assert np.array_equal(erows, z["sql_l2_rows"]) and np.array_equal(edist, z["sql_l2_dist"])
    # the HNSW result at ef=40 on this small corpus is the exact answer for almost every query
    rec = np.mean([len(set(z["l2_nodes"][i].tolist()) & set(erows[i].tolist())) / 10 for i in range(len(erows))])
    assert rec >= 0.95
