#!/usr/bin/env python
"""Kernel list of ONE replay of the pipelined training graph (dgll_b200.pipelined): both branches of one mini-batch.
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/pipelined_step.csv python tools/profile_pipelined_step.py [reddit|papers] [bf16|fp32]
"reddit": N=232,965 F=602 hidden 256, 41 classes, resident table.  "papers": a 1/50 papers100M-shaped graph, F=128,
172 classes, one shard (the kernels and shapes of the partitioned run on one GPU)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import dgll_b200.nn as dnn  # noqa: E402
from dgll_b200 import graphs as G, parallel as P, pipelined as PL  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "reddit"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(0)
if which == "reddit":
    N, NNZ, F, C = 232965, 114615892, 602, 41
    rp, col = G.rmat_csr(N, NNZ, seed=0, device=dev)
    table = G.feature_table(N, F, seed=0, device=dev)
    kw = {"table": table}
else:
    N, NNZ, F, C = 2221199, 32313717, 128, 172
    rp, col = G.rmat_csr_large(N, NNZ, seed=0, device=dev)
    table = G.feature_table(N, F, seed=0, device=dev)
    kw = {"sharded": P.PeerShardedTable(N, table)}
labels = torch.randint(0, C, (N,), device=dev, generator=gen)
seeds = torch.randperm(int(0.66 * N), device=dev, generator=gen)[:1024 * 8]
model = dnn.GraphSAGE(F, 256, C, 2, torch.relu, 0.0).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=0.003, fused=True, capturable=True)
tr = PL.PipelinedSageTrainer(model, opt, labels, rp, col, F, batch_size=1024, fanouts=(25, 10), precision=prec, **kw)
tr.epoch(seeds)
tr.set_seeds(seeds)
tr._prologue.replay()
tr.graphs[0].replay()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.graphs[1].replay()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
