"""SQ8 arena: encode parity with SQ8Vector::from_f32 (src/hnsw/quantization.rs:68-95, oracle restatement) and
traversal parity — the SQ8 kernels must equal the FP32 search over the DECODED vectors bit for bit (the contract
stated in include/turdb_cuda.h; the reference declares SQ8 but never wires it into its index)."""
import numpy as np
import pytest
import torch

from oracle import binding as ob
from turdb_b200 import datasets as ds
from turdb_b200.hnsw import CudaHnswIndex, DistanceFunction

pytestmark = pytest.mark.gpu


def run_sq8(idx, q, k, ef, metric):
    dev = torch.device("cuda:0")
    nq = q.shape[0]
    dq = torch.from_numpy(q).to(dev)
    rows = torch.empty((nq, k), dtype=torch.int64, device=dev)
    dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
    nodes = torch.empty((nq, k), dtype=torch.int32, device=dev)
    cnt = torch.empty(nq, dtype=torch.int32, device=dev)
    stats = torch.empty((nq, 4), dtype=torch.int32, device=dev)
    idx.search_batch_sq8_device(dq.data_ptr(), nq, k, ef, metric, rows.data_ptr(), dist.data_ptr(), cnt.data_ptr(),
                                nodes.data_ptr(), stats.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return (rows.cpu().numpy().astype(np.uint64), nodes.cpu().numpy().view(np.uint32), dist.cpu().numpy(),
            cnt.cpu().numpy().view(np.uint32), stats.cpu().numpy().view(np.uint32))


@pytest.mark.parametrize("dim", [128, 100, 7])
def test_encode_matches_from_f32(gpu_required, dim):
    x = ds.gaussian_latent(3000, dim, seed=11)
    x[5] = 0.25          # constant row: range 0 -> scale 1.0, codes 0
    x[6, :] = np.linspace(-1, 1, dim, dtype=np.float32)
    n = x.shape[0]
    g = dict(vectors=x, row_ids=np.arange(n, dtype=np.uint64), levels=np.zeros(n, np.uint8),
             l0_adj=np.full((n, 32), 0xFFFFFFFF, np.uint32), l0_cnt=np.zeros(n, np.uint8), up_base=np.full(n, 0xFFFFFFFF, np.uint32),
             up_adj=np.zeros((0, 16), np.uint32), up_cnt=np.zeros(0, np.uint8), entry=0, max_level=0)
    idx = CudaHnswIndex.from_graph(g)
    codes, mn, sc = idx.enable_sq8(return_rows=True)
    o_codes, o_mn, o_sc = ob.sq8_encode(x)
    assert np.array_equal(mn.view(np.uint32), o_mn.view(np.uint32))
    assert np.array_equal(sc.view(np.uint32), o_sc.view(np.uint32))
    assert np.array_equal(codes, o_codes)
    idx.close()


@pytest.mark.parametrize("metric", [ob.L2, ob.COSINE, ob.IP])
def test_sq8_search_equals_fp32_search_over_decoded_vectors(gpu_required, small_graph, metric):
    g, arrays = small_graph
    idx = CudaHnswIndex.from_graph(arrays)
    codes, mn, sc = idx.enable_sq8(return_rows=True)
    decoded = ob.sq8_decode(codes, mn, sc)
    a2 = dict(arrays)
    a2["vectors"] = decoded
    og = ob.OracleGraph.from_arrays(a2)
    q = ds.gaussian_latent(500, 128, seed=2)
    gpu = run_sq8(idx, q, 10, 64, metric)
    cpu = og.search(q, 10, 64, metric, n_threads=8)
    assert np.array_equal(gpu[3], cpu[3])
    assert np.array_equal(gpu[1], cpu[1]), "ids differ from the FP32 search over the decoded vectors"
    assert np.array_equal(gpu[2].view(np.uint32), cpu[2].view(np.uint32))
    for j, f in enumerate(("n_dist", "n_dist_upper", "n_expanded", "n_upper_hops")):
        assert np.array_equal(gpu[4][:, j], cpu[4][f])
    # and it stays a good approximation of the FP32 index
    fp = idx.search_batch(q, 10, 64, DistanceFunction(metric))
    agree = np.mean([len(set(gpu[1][i].tolist()) & set(fp[1][i].tolist())) / 10 for i in range(500)])
    assert agree >= 0.95
    idx.close()


def test_sq8_requires_enable(gpu_required, small_graph):
    g, arrays = small_graph
    idx = CudaHnswIndex.from_graph(arrays)
    with pytest.raises(ValueError, match="enable_sq8"):
        idx.search_batch_sq8_device(0, 1, 1, 8, 0, 0, 0, 0)
    idx.close()
