"""Short target for ncu: build the config-2 graph, run the traversal kernel a few times."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from turdb_b200 import datasets as ds
from turdb_b200.graph_build import build_graph
from turdb_b200.hnsw import CudaHnswIndex

n, dim, nq, k, ef = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 384, 10_000, 10, 128
x = ds.gaussian_latent(n, dim, seed=1, latent=16, normalise=True)
q = ds.gaussian_latent(nq, dim, seed=2, latent=16, normalise=True)
arrays = build_graph(x, seed=42)
idx = CudaHnswIndex.from_graph(arrays)
dev = torch.device("cuda:0")
dq = torch.from_numpy(q).to(dev)
rows = torch.empty((nq, k), dtype=torch.int64, device=dev)
dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
cnt = torch.empty(nq, dtype=torch.int32, device=dev)
stats = torch.empty((nq, 4), dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    idx.search_batch_device(dq.data_ptr(), nq, k, ef, 1, rows.data_ptr(), dist.data_ptr(), cnt.data_ptr(), 0,
                            stats.data_ptr(), 0, stream)
torch.cuda.synchronize()
st = stats.cpu().numpy().astype(np.int64)
print("algorithmic bytes per launch", int((st[:, 0] * dim * 4 + st[:, 2] * 129 + st[:, 3] * 65 + dim * 4 + k * 12).sum()))
