import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgll_b200 import kernels as K
torch.manual_seed(0)
for (M, N, Kd) in ((128, 128, 32), (128, 128, 64), (256, 256, 96), (300, 64, 50)):
    a = torch.randn(M, Kd, device="cuda"); b = torch.randn(Kd, N, device="cuda")
    ref = a.double() @ b.double()
    at, bt = a.t().contiguous(), b.t().contiguous()
    for name, fn in (("A K-major, B K-major", lambda: K.gemm(a, bt, trans_b=True, precision="tf32")),
                     ("A K-major, B MN-major", lambda: K.gemm(a, b, precision="tf32")),
                     ("A MN-major, B K-major", lambda: K.gemm(at, bt, trans_a=True, trans_b=True, precision="tf32")),
                     ("A MN-major, B MN-major", lambda: K.gemm(at, b, trans_a=True, precision="tf32"))):
        o = fn(); torch.cuda.synchronize()
        err = ((o.double() - ref).abs().max() / ref.abs().max()).item()
        print(M, N, Kd, name, "rel_err %.3e" % err, "out[0,:4]", o[0, :4].tolist(), "ref", ref[0, :4].tolist(), flush=True)
