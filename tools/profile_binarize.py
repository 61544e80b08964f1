#!/usr/bin/env python
"""binarize_pack on the Reddit-shaped table inside a profiler range (ncu -k regex:binarize)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from dgll_b200 import graphs as G, kernels as K  # noqa: E402

dev = torch.device("cuda", 0)
N, _, F, _ = G.SHAPES["reddit"]
x = G.feature_table(N, F, seed=1, device=dev)
for it in range(3):
    if it == 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    K.binarize_pack(x[:, :F])
    if it == 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
print("done")
