#!/usr/bin/env python
"""One full-graph mean aggregation per graph / width inside a profiler range, for the DRAM bytes that go next to the
algorithmic-byte fractions above 1.0:
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
      --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_fullgraph_dram.csv \
      python tools/profile_fullgraph.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from dgll_b200 import graphs as G, kernels as K  # noqa: E402

dev = torch.device("cuda", 0)
N, E, F, _ = G.SHAPES["reddit"]
cases = []
rp, col = G.rmat_csr(N, E, seed=0, device=dev)
cases.append(("reddit-shaped R-MAT", rp, col, N))
urp, ucol = G.uniform_csr(N, E // N, seed=1, device=dev)
cases.append(("uniform control", urp, ucol, N))
Np, Ep, _, _ = G.SHAPES["products"]
prp, pcol = G.rmat_csr(Np, 2 * Ep, seed=2, device=dev, symmetric=True)
cases.append(("products-shaped R-MAT", prp, pcol, Np))
for name, r, c, n in cases:
    plan = K.CsrPlan(r, chunk_edges=4096)
    plan = plan if plan.n_heavy_rows > 0 else None
    for width in ((602, 256) if n == N else (256,)):
        x = G.feature_table(n, width, seed=3, device=dev)
        out = torch.empty((n, width), device=dev)
        for it in range(2):
            if it == 1:
                torch.cuda.synchronize()
                torch.cuda.cudart().cudaProfilerStart()
            K.spmm_csr(r, c, x[:, :width], reduce="mean", out=out, plan=plan)
            if it == 1:
                torch.cuda.synchronize()
                torch.cuda.cudart().cudaProfilerStop()
        print("CASE %s N=%d nnz=%d F=%d alg_bytes=%d" % (name, n, c.numel(), width,
                                                      c.numel() * (4 + width * 4) + n * (width * 4 + 8)), flush=True)
        del x, out
