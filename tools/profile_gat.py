#!/usr/bin/env python
"""One fused-GAT forward and backward (products-shaped, default kernels + long-row plans) inside a profiler range:
  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r01_gat2 \
      python tools/profile_gat.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from dgll_b200 import graphs as G, kernels as K  # noqa: E402

dev = torch.device("cuda", 0)
Np, E, _, _ = G.SHAPES["products"]
rp, col = G.rmat_csr(Np, 2 * E, seed=2, device=dev, symmetric=True)
heads, D = 4, 64
g = torch.Generator(device=dev).manual_seed(4)
wh = torch.randn((Np, heads * D), device=dev, generator=g)
el = torch.randn((Np, heads), device=dev, generator=g)
er = torch.randn((Np, heads), device=dev, generator=g)
gout = torch.randn_like(wh)
plan = K.CsrPlan(rp, chunk_edges=1024)
trp, tcol, _, perm = K.csr_transpose(rp, col, Np, want_perm=True)
bp, btp = K.CsrPlan(rp, chunk_edges=256), K.CsrPlan(trp, chunk_edges=256)
for it in range(2):
    if it == 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    out, rmax, rsum = K.gat_forward(rp, col, wh, el, er, heads, 0.2, save_stats=True, plan=plan)
    K.gat_backward(rp, col, trp, tcol, perm, wh, el, er, out, rmax, rsum, gout, heads, 0.2, plan=bp, t_plan=btp)
    if it == 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
print("done")
