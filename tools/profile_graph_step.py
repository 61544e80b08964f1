#!/usr/bin/env python
"""Kernel list of ONE replay of the CUDA-graph training step (train.GraphedSageTrainer), Reddit-shaped, bf16 GEMMs:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/graph_step.csv python tools/profile_graph_step.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import dgll_b200.nn as dnn  # noqa: E402
from dgll_b200 import graphs as G, train as T  # noqa: E402

dev = torch.device("cuda", 0)
N, NNZ, F = 232965, 114615892, 602
rp, col = G.rmat_csr(N, NNZ, seed=0, device=dev)
table = G.feature_table(N, F, seed=0, device=dev)
gen = torch.Generator(device=dev).manual_seed(0)
labels = torch.randint(0, 41, (N,), device=dev, generator=gen)
seeds = torch.randperm(int(0.66 * N), device=dev, generator=gen)[:1024 * 4]
model = dnn.GraphSAGE(F, 256, 41, 2, torch.relu, 0.0).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=0.003, fused=True, capturable=True)
pre = T.make_batches(rp, col, seeds, (25, 10), 1024, rng_seed=3)
tr = T.GraphedSageTrainer(model, opt, table, labels, 1024, (25, 10), precision=sys.argv[1] if len(sys.argv) > 1 else "bf16")
tr.load(*pre[0])
tr.capture()
tr.epoch(pre)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.load(*pre[1])
tr.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
