"""`.hnsw` reader (csrc/hnsw_file.inl, host code) against files produced by the oracle's restatement of the
reference's write path (storage.rs / mod.rs:776-904).  CPU tests parse only; the GPU test uploads and
searches.  The reference holds no test or fixture for this format ("parity unpinned" applies)."""
import struct

import numpy as np
import pytest

from oracle import binding as ob
from turdb_b200 import datasets as ds
from turdb_b200.hnsw_file import (FLAG_MAX_LEVEL_CLAMPED, FLAG_NODE_COUNT_MISMATCH, FLAG_SUSPECT_PAGES, FLAG_TOMBSTONES,
                                  HnswFile)

PAGE = 16384


@pytest.fixture(scope="module")
def graph2k():
    x = ds.gaussian_latent(2000, 32, seed=5)
    g = ob.OracleGraph.build(x, m=16, ef_construction=60, mode=ob.BUILD_INTENT, seed=9,
                             row_ids=np.arange(2000, dtype=np.uint64) * 5 + 11)
    return g, g.export(), x


def test_header_and_graph_round_trip(graph2k, tmp_path):
    g, arrays, x = graph2k
    data, pages, slots = ob.hnsw_file_write(g, index_id=77, table_id=5, ef_search=48, distance_fn=ob.COSINE, mode=1)
    assert len(data) % PAGE == 0 and data[:16] == b"TurDB HNSW\0\0\0\0\0\0"
    p = tmp_path / "t.hnsw"
    p.write_bytes(data)
    for f in (HnswFile.open(str(p)), HnswFile.from_bytes(data)):
        i = f.info
        assert (i["index_id"], i["table_id"], i["dimensions"], i["m"], i["m0"], i["ef_construction"], i["ef_search"]) == \
            (77, 5, 32, 16, 32, 60, 48)
        assert i["distance_fn"] == 1 and i["quantization"] == 0 and i["has_entry"] == 1
        assert i["n_nodes"] == 2000 == i["header_node_count"] and i["n_tombstones"] == 0
        assert i["flags"] == 0 and i["n_suspect_pages"] == 0 and i["n_pages"] == len(data) // PAGE
        assert i["max_level"] == g.max_level == i["header_max_level"] and i["entry"] == g.entry
        rows, pg, sl = f.nodes()
        assert np.array_equal(rows, arrays["row_ids"]) and np.array_equal(pg, pages) and np.array_equal(sl, slots)
        got = f.graph()
        for k in ("levels", "l0_cnt", "up_base", "up_cnt"):
            assert np.array_equal(got[k], arrays[k]), k
        # neighbour lists in stored order (padding beyond the count is INVALID on both sides)
        for k, c in (("l0_adj", "l0_cnt"), ("up_adj", "up_cnt")):
            w = got[k].shape[1]
            mask = np.arange(w)[None, :] < arrays[c][:, None]
            assert np.array_equal(got[k][mask], arrays[k][mask]), k
            assert (got[k][~mask] == 0xFFFFFFFF).all()
        f.close()


def test_node_ids_follow_allocate_node_order(graph2k):
    g, arrays, _ = graph2k
    data, pages, slots = ob.hnsw_file_write(g, mode=1)
    # dense id order == (page, slot) order (allocate_node appends: mod.rs:883-904)
    key = pages.astype(np.int64) * 65536 + slots
    assert (np.diff(key) > 0).all()
    # every record parses with the oracle's own HnswNode::read_from through the slot directory
    for i in (0, 1, 777, 1999):
        pg = data[int(pages[i]) * PAGE:(int(pages[i]) + 1) * PAGE]
        assert pg[0] == 0x10
        os_, size = struct.unpack_from("<HH", pg, 64 + 4 * int(slots[i]))
        off, status = os_ & 0x1FFF, (os_ >> 13) & 3
        assert status == 1
        nd = ob.node_read(pg[off:off + size])
        assert nd["row_id"] == arrays["row_ids"][i] and nd["max_level"] == arrays["levels"][i]
        assert nd["l0"] == [(int(pages[j]), int(slots[j])) for j in arrays["l0_adj"][i, :arrays["l0_cnt"][i]]]


def test_verbatim_page_fill_is_flagged(graph2k):
    """The reference's can_fit rule packs ~77 records per page although slot offsets keep 13 bits
    (storage.rs:338-344): records overlap.  The reader must flag it, not crash."""
    g, _, _ = graph2k
    small = ob.OracleGraph.build(ds.gaussian_latent(120, 16, seed=3), m=16, ef_construction=40, seed=4)
    try:
        data, _, _ = ob.hnsw_file_write(small, mode=0)
    except RuntimeError as e:  # -2: the writer itself hit the overwritten page header, as the reference would
        assert "-2" in str(e)
        return
    f = HnswFile.from_bytes(data)
    assert f.info["flags"] & FLAG_SUSPECT_PAGES and f.info["n_suspect_pages"] >= 1


def test_deleted_and_dangling_nodes_become_tombstones(graph2k):
    g, arrays, _ = graph2k
    data, pages, slots = ob.hnsw_file_write(g, mode=1)
    buf = bytearray(data)
    victim = 123
    so = int(pages[victim]) * PAGE + 64 + 4 * int(slots[victim])
    os_, = struct.unpack_from("<H", buf, so)
    struct.pack_into("<H", buf, so, (os_ & 0x1FFF) | (2 << 13))  # HnswPage::mark_deleted, storage.rs:671-688
    f = HnswFile.from_bytes(bytes(buf))
    i = f.info
    assert i["n_nodes"] == 1999 and i["n_deleted_slots"] == 1 and i["n_tombstones"] == 1
    assert i["flags"] & FLAG_TOMBSTONES and i["flags"] & FLAG_NODE_COUNT_MISMATCH
    rows, pg, sl = f.nodes()
    assert rows[-1] == 0 and pg[-1] == pages[victim] and sl[-1] == slots[victim]
    got = f.graph()
    tomb = 1999
    assert got["levels"][tomb] == 0 and got["l0_cnt"][tomb] == 0
    # every former neighbour of the victim now points at the tombstone; dense ids above it shifted down by one
    refs_before = int((arrays["l0_adj"] == victim).sum())
    assert refs_before > 0 and int((got["l0_adj"] == tomb).sum()) == refs_before
    v = np.zeros((1999, 32), np.float32)
    full = f.graph(v)["vectors"]
    assert full.shape == (2000, 32) and np.isinf(full[tomb]).all()


def test_empty_index_and_bad_magic():
    empty = ob.OracleGraph.new(8)
    data, _, _ = ob.hnsw_file_write(empty, mode=1)
    assert len(data) == PAGE
    f = HnswFile.from_bytes(data)
    assert f.info["n_nodes"] == 0 and f.info["has_entry"] == 0 and f.info["entry"] == 0xFFFFFFFF
    with pytest.raises(ValueError, match="magic bytes mismatch"):  # storage.rs:166-169
        HnswFile.from_bytes(b"NotTurDB" + bytes(PAGE - 8))
    with pytest.raises(ValueError, match="buffer too small"):      # storage.rs:159-164
        HnswFile.from_bytes(b"TurDB")


def test_header_max_level_above_entry_level_is_clamped(graph2k):
    g, _, _ = graph2k
    data, _, _ = ob.hnsw_file_write(g, mode=1)
    buf = bytearray(data)
    buf[50] = g.max_level + 3
    f = HnswFile.from_bytes(bytes(buf))
    assert f.info["header_max_level"] == g.max_level + 3 and f.info["max_level"] == g.max_level
    assert f.info["flags"] & FLAG_MAX_LEVEL_CLAMPED


@pytest.mark.gpu
def test_uploaded_file_searches_like_the_oracle(gpu_required, graph2k):
    from turdb_b200.hnsw import DistanceFunction
    g, arrays, x = graph2k
    data, _, _ = ob.hnsw_file_write(g, mode=1, distance_fn=ob.L2)
    f = HnswFile.from_bytes(data)
    q = ds.gaussian_latent(300, 32, seed=6)
    cpu = g.search(q, 10, 64, ob.L2, n_threads=4)
    table = {int(r): x[i] for i, r in enumerate(arrays["row_ids"])}
    for idx in (f.upload(vectors=x), f.upload(get_vector=lambda r: table.get(r))):
        gpu = idx.search_batch(q, 10, 64, DistanceFunction.L2)
        assert np.array_equal(gpu[1], cpu[1]) and np.array_equal(gpu[0], cpu[0])
        assert np.array_equal(gpu[2].view(np.uint32), cpu[2].view(np.uint32))
        idx.close()


@pytest.mark.gpu
def test_missing_vectors_and_tombstones_rank_last(gpu_required, graph2k):
    """get_vector -> None and unreadable nodes evaluate to +inf (mod.rs:1111-1121): never in a full top-k."""
    from turdb_b200.hnsw import DistanceFunction
    g, arrays, x = graph2k
    data, pages, slots = ob.hnsw_file_write(g, mode=1)
    buf = bytearray(data)
    for victim in (50, 900):
        so = int(pages[victim]) * PAGE + 64 + 4 * int(slots[victim])
        os_, = struct.unpack_from("<H", buf, so)
        struct.pack_into("<H", buf, so, (os_ & 0x1FFF) | (2 << 13))
    f = HnswFile.from_bytes(bytes(buf))
    rows, _, _ = f.nodes()
    gone = {int(arrays["row_ids"][7]), int(arrays["row_ids"][1500])}
    table = {int(r): x[i] for i, r in enumerate(arrays["row_ids"])}
    idx = f.upload(get_vector=lambda r: None if r in gone else table.get(r))
    q = ds.gaussian_latent(200, 32, seed=8)
    bad_rows = gone | {int(arrays["row_ids"][50]), int(arrays["row_ids"][900])}
    # same graph arrays + same +inf rows through the oracle: identical results
    ga = f.graph()
    vec = np.stack([np.full(32, np.inf, np.float32) if (int(r) in gone or i >= f.info["n_nodes"]) else table[int(r)]
                    for i, r in enumerate(rows)])
    ga["vectors"] = vec
    og = ob.OracleGraph.from_arrays(ga)
    # the reference's closure yields INFINITY for such nodes whatever the metric (mod.rs:1111-1121): cosine and inner
    # product must not turn the +inf row into NaN / -inf and rank it first
    for form in (0, 1, 2):
        idx.set_traversal_form(form)
        for metric in (DistanceFunction.L2, DistanceFunction.Cosine, DistanceFunction.InnerProduct):
            got = idx.search_batch(q, 10, 64, metric)
            assert np.isfinite(got[2]).all()
            assert not (set(got[0].ravel().tolist()) & bad_rows)
            cpu = og.search(q, 10, 64, int(metric), n_threads=4)
            assert np.array_equal(got[1], cpu[1]) and np.array_equal(got[2].view(np.uint32), cpu[2].view(np.uint32))
    idx.close()


def test_reader_survives_corrupted_files(graph2k):
    """Random byte damage and truncation: the reader either rejects the file or returns a graph whose every
    neighbour id, entry point and count is in range (index_create re-checks them on upload)."""
    g, _, _ = graph2k
    data, _, _ = ob.hnsw_file_write(g, mode=1)
    rng = np.random.default_rng(0)
    parsed = 0
    for it in range(120):
        b = bytearray(data)
        for _ in range(int(rng.integers(1, 40))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        if it % 7 == 0:
            b = b[: int(rng.integers(0, len(b)))]
        try:
            f = HnswFile.from_bytes(bytes(b))
        except (ValueError, RuntimeError):
            continue
        gr = f.graph()
        n = f.n_total
        assert gr["l0_adj"].shape == (n, 32) and (gr["l0_cnt"] <= 32).all() and (gr["up_cnt"] <= 16).all()
        assert (gr["l0_adj"][gr["l0_adj"] != 0xFFFFFFFF] < n).all()
        assert (gr["up_adj"][gr["up_adj"] != 0xFFFFFFFF] < n).all()
        assert f.info["entry"] == 0xFFFFFFFF or f.info["entry"] < n
        f.close()
        parsed += 1
    assert parsed > 50


def test_vector_literal_parser():
    from turdb_b200.sql_operator import parse_vector_literal
    assert parse_vector_literal(" [0.1, 2,-3.5e-1] ").tolist() == [np.float32(0.1), 2.0, np.float32(-0.35)]
    with pytest.raises(ValueError):
        parse_vector_literal("0.1, 0.2")


def test_vector_topk_planner_rule():
    from turdb_b200.sql_operator import VectorOp, plan_vector_topk
    # the statements of the reference's own k-NN tests (tests/hnsw_integration.rs:229, 249, 269)
    p = plan_vector_topk("SELECT id, name FROM embeddings ORDER BY vec <-> '[0.1, 0.1, 0.1, 0.1]' LIMIT 2")
    assert p["table"] == "embeddings" and p["column"] == "vec" and p["op"] == VectorOp.L2Distance
    assert p["limit"] == 2 and p["offset"] == 0 and p["projection"] == ["id", "name"] and p["literal"].tolist() == [np.float32(0.1)] * 4
    p = plan_vector_topk("select id from embeddings order by vec <=> '[0.5,0.5]' limit 3 offset 4;")
    assert p["op"] == VectorOp.CosineDistance and (p["limit"], p["offset"]) == (3, 4)
    # not rewritten: <#> is a NULL sort key in the reference, DESC asks for the farthest rows, WHERE changes the input
    assert plan_vector_topk("SELECT id FROM t ORDER BY vec <#> '[1,2]' LIMIT 5") is None
    assert plan_vector_topk("SELECT id FROM t ORDER BY vec <-> '[1,2]' DESC LIMIT 5") is None
    assert plan_vector_topk("SELECT id FROM t WHERE id > 3 ORDER BY vec <-> '[1,2]' LIMIT 5") is None
    assert plan_vector_topk("SELECT id FROM t ORDER BY vec <-> '[1,2]', id LIMIT 5") is None
    assert plan_vector_topk("SELECT id FROM t ORDER BY id LIMIT 5") is None
