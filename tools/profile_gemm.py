#!/usr/bin/env python
"""A few launches of the TF32 dense transform on one shape inside a profiler range:
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tf32 -c 1 \
      -o gpurun_out/r02_gemm_x python tools/profile_gemm.py 2449029 256 100 [precision [gemm_kernel option]]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from dgll_b200 import kernels as K  # noqa: E402

M, N, Kd = (int(v) for v in sys.argv[1:4])
prec = sys.argv[4] if len(sys.argv) > 4 else "tf32"          # tf32 | tf32x3 | bf16 | fp32
if len(sys.argv) > 5:
    K.set_option("gemm_kernel", sys.argv[5])
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
a = torch.randn((M, Kd), device=dev, generator=g)
w = torch.randn((N, Kd), device=dev, generator=g)
out = torch.empty((M, N), device=dev)
for it in range(3):
    if it == 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    K.gemm(a, w, trans_b=True, out=out, precision=prec)
    if it == 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
print("done")
