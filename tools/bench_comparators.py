#!/usr/bin/env python
"""Same-box GPU comparators and control runs (BASELINE.md §3 "GPU-vs-GPU comparators"):

  k1       the reference's own fused forward kernel (dgll/FusedKernel/gcn_fused_kernel.cu:5-74, compiled in place for
           sm_100a into oracle/_ref/) against dgllb_gcn_fused_forward on the PPI graph shapes it can launch
           (F <= 384: it asks for 128*F bytes of dynamic shared memory, gcn_fused_kernel.cu:203-210)
  cusparse torch.sparse.mm on a CUDA CSR matrix (cuSPARSE; the call behind gcnconv.py:31 / gcn_model.py:76) against
           dgllb_spmm_csr on the full Reddit-shaped graph (F = 602 and 256) and the products-shaped graph (F = 256)
  uniform  the same full-graph aggregation on the uniform-random control graph of the same (N, nnz)
  bin      BASELINE configs[3] at model level: 2-layer binarized-feature GCN (nn.BinGCN) against the 2-layer fp32 GCN
           with mean aggregation on the Reddit-shaped graph, forward and forward+backward+Adam
One JSON line per measurement.  python tools/bench_comparators.py [--which k1,cusparse,uniform,bin]"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from dgll_b200 import graphs as G, kernels as K, ops  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def out(**kw):
    print(json.dumps(kw), flush=True)


def run_k1(dev):
    import oracle
    path = oracle.ref_kernel_path()
    if path is None:
        out(kernel="K1", error="oracle/_ref/libgcn_fused_ref.so not built")
        return
    lib = ctypes.CDLL(path)
    lib.launch_gcn_fused_kernel.restype = None
    P = lambda z: ctypes.c_void_p(z.data_ptr())  # noqa: E731
    g = torch.Generator(device=dev).manual_seed(0)
    # PPI train graphs: smallest, a middle one, largest; F = 50 padded to 52, hidden 64 / 256
    for (N, nnz), Hd in (((591, 7708), 64), ((2263, 59644), 64), ((3480, 106754), 64), ((3480, 106754), 256)):
        F, Fp = 50, 52
        deg = torch.full((N,), nnz // N, dtype=torch.int64, device=dev)
        deg[: nnz - (nnz // N) * N] += 1
        rp = torch.zeros(N + 1, dtype=torch.int32, device=dev)
        rp[1:] = torch.cumsum(deg, 0).to(torch.int32)
        col = torch.randint(0, N, (nnz,), device=dev, generator=g, dtype=torch.int32)
        vals = torch.rand(nnz, device=dev, generator=g)
        X = torch.zeros((N, Fp), device=dev)
        X[:, :F] = torch.randn((N, F), device=dev, generator=g)
        W = torch.randn((Fp, Hd), device=dev, generator=g)
        nn_ = deg.to(torch.int32)
        H = torch.zeros((N, Hd), device=dev)

        def ref_call():
            H.zero_()
            lib.launch_gcn_fused_kernel(P(rp), P(col), P(vals), P(X), P(W), P(H), P(nn_), ctypes.c_int(N), ctypes.c_int(Fp),
                                        ctypes.c_int(F), ctypes.c_int(Hd), ctypes.c_int(nnz))

        ms_ref = timeit(ref_call, 20)
        ours = None

        def our_call():
            nonlocal ours
            ours = K.gcn_fused_forward_v2(rp, col, vals, X, W, nn_, F)

        ms_ours = timeit(our_call, 20)
        ref_call()
        torch.cuda.synchronize()
        err = ((ours - H).abs().max() / H.abs().max().clamp(min=1e-30)).item()
        out(kernel="fused GCN forward relu(A(XW))", shape="N=%d nnz=%d F=%d H=%d" % (N, nnz, F, Hd),
            reference_K1_ms=round(ms_ref, 4), ours_ms=round(ms_ours, 4), speedup=round(ms_ref / ms_ours, 2),
            max_rel_diff=err, note="K1 launch is blocking (cudaDeviceSynchronize inside, gcn_fused_kernel.cu:225) and "
                                   "recomputes X.W per edge; ours = one GEMM per node + SpMM with fused ReLU")


def spmm_pair(name, rp, col, N, Fw, dev, iters):
    nnz = col.numel()
    x = G.feature_table(N, Fw, seed=1, device=dev)
    view = x[:, :Fw]
    nbytes = nnz * (4 + Fw * 4) + N * (Fw * 4 + 8)
    o = torch.empty((N, x.size(1)), device=dev)[:, :Fw]
    g = ops.CsrGraph(rp, col)
    plan = g.plan()
    ms = timeit(lambda: K.spmm_csr(rp, col, view, reduce="mean", out=o, plan=plan), iters)
    res = dict(kernel="spmm mean, full graph", graph=name, shape="N=%d nnz=%d F=%d" % (N, nnz, Fw), ours_ms=round(ms, 3),
               ours_alg_GBps=round(nbytes / ms / 1e6, 1), ours_frac_of_measured_hbm=round(nbytes / ms / 1e6 / peak(), 3),
               plan_heavy_rows=(plan.n_heavy_rows if plan is not None else 0))
    try:
        deg = (rp[1:] - rp[:-1])
        vals = torch.repeat_interleave(1.0 / deg.clamp(min=1).to(torch.float32), deg)
        A = torch.sparse_csr_tensor(rp, col.long(), vals, size=(N, N))
        xc = view.contiguous()
        ms_lib = timeit(lambda: torch.sparse.mm(A, xc), max(3, iters // 2), warmup=1)
        ref = torch.sparse.mm(A, xc)
        res.update(cusparse_ms=round(ms_lib, 3), speedup_vs_cusparse=round(ms_lib / ms, 2),
                   max_rel_diff=((ref - o).abs().max() / ref.abs().max()).item())
        del A, xc, ref, vals
    except Exception as ex:
        res.update(cusparse_error="%s: %s" % (type(ex).__name__, str(ex)[:160]))
    out(**res)
    del x, o


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="k1,cusparse,uniform,bin")
    ap.add_argument("--iters", type=int, default=8)
    args = ap.parse_args()
    which = set(args.which.split(","))
    dev = torch.device("cuda", 0)
    if "k1" in which:
        run_k1(dev)
    N, NNZ, F, C = G.SHAPES["reddit"]
    if which & {"cusparse", "bin"}:
        rp, col = G.rmat_csr(N, NNZ, seed=0, device=dev)
    if "cusparse" in which:
        for Fw in (602, 256):
            spmm_pair("reddit-shaped R-MAT", rp, col, N, Fw, dev, args.iters)
    if "uniform" in which:
        urp, ucol = G.uniform_csr(N, NNZ // N, seed=0, device=dev)
        for Fw in (602, 256):
            spmm_pair("uniform control (same N, deg %d)" % (NNZ // N), urp, ucol, N, Fw, dev, args.iters)
        del urp, ucol
    if "cusparse" in which:
        Np, E, _, _ = G.SHAPES["products"]
        prp, pcol = G.rmat_csr(Np, 2 * E, seed=2, device=dev, symmetric=True)
        spmm_pair("products-shaped R-MAT (symmetric)", prp, pcol, Np, 256, dev, args.iters)
        del prp, pcol
    if "bin" in which:
        import dgll_b200.nn as dnn
        x = G.feature_table(N, F, seed=1, device=dev)[:, :F]
        y = torch.randint(0, C, (N,), device=dev)
        g = ops.CsrGraph(rp, col)
        g.plan(), g.bin_plan(), g.transpose().plan()
        ops.set_gemm_precision("tf32")

        class MeanGCN(torch.nn.Module):   # fp32 comparator: the same two-layer model with the fp32 mean aggregation
            def __init__(self):
                super().__init__()
                self.l1, self.l2 = torch.nn.Linear(F, 256), torch.nn.Linear(256, C)

            def forward(self, x, g):
                h = torch.relu(ops.spmm(g, ops.linear(x, self.l1.weight, bias=self.l1.bias, trans_w=True), reduce="mean"))
                return torch.log_softmax(ops.spmm(g, ops.linear(h, self.l2.weight, bias=self.l2.bias, trans_w=True),
                                                  reduce="mean"), dim=1)

        for name, model in (("fp32 mean-aggregation GCN", MeanGCN().to(dev)), ("binarized-feature GCN (nn.BinGCN)", dnn.BinGCN(F, 256, C, 0.0).to(dev))):
            opt = torch.optim.Adam(model.parameters(), lr=0.01, fused=True)
            with torch.no_grad():
                ms_f = timeit(lambda: model(x, g), 5)

            def step():
                opt.zero_grad(set_to_none=True)
                loss = torch.nn.functional.nll_loss(model(x, g), y)
                loss.backward()
                opt.step()
                return loss

            ms_s = timeit(step, 5)
            l0 = step().item()
            for _ in range(10):
                l1 = step().item()
            out(model=name, graph="reddit-shaped full batch N=%d nnz=%d F=%d hidden 256" % (N, NNZ, F),
                forward_ms=round(ms_f, 2), train_step_ms=round(ms_s, 2), loss_first=round(l0, 4), loss_after_10_more=round(l1, 4))
        ops.set_gemm_precision("fp32")


if __name__ == "__main__":
    main()
