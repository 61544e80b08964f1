"""Host-side profile of the training step (cProfile) — where the per-batch milliseconds go."""
import cProfile, pstats, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgll_b200.nn as dnn
from dgll_b200 import graphs as G, train as T

dev = torch.device("cuda", 0)
N, NNZ, F = 232965, 114615892, 602
rp, col = G.rmat_csr(N, NNZ, seed=0, device=dev)
table = G.feature_table(N, F, seed=0, device=dev)
gen = torch.Generator(device=dev).manual_seed(0)
labels = torch.randint(0, 41, (N,), device=dev, generator=gen)
seeds = torch.randperm(int(0.66 * N), device=dev, generator=gen)[:1024 * 40]
model = dnn.GraphSAGE(F, 256, 41, 2, torch.relu, 0.0).to(dev)
from dgll_b200 import ops
ops.set_gemm_precision("bf16")
opt = torch.optim.Adam(model.parameters(), lr=0.003, fused=True)
pre = T.make_batches(rp, col, seeds, (25, 10), 1024, rng_seed=3)
T.sage_epoch(model, opt, table, labels, F, batches=pre[:8])
pr = cProfile.Profile()
pr.enable()
r = T.sage_epoch(model, opt, table, labels, F, batches=pre)
pr.disable()
print(r)
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
pstats.Stats(pr).sort_stats("cumulative").print_stats(40)
