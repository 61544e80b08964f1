#!/usr/bin/env python
"""Per-kernel throughput on the BASELINE.json shapes (not the driver's bench line — see bench.py).

Prints one JSON object per measurement: kernel, shape, ms, algorithmic GB/s (SURVEY.md §8 d formulas) and the
fraction of the measured HBM peak.  Inputs are larger than L2 or rotated between iterations.

  python tools/bench_kernels.py [--which full,gather,bin,gat,block] [--iters 10]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from dgll_b200 import graphs as G  # noqa: E402
from dgll_b200 import kernels as K  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timeit(fn, iters, warmup=3):
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


def report(name, shape, ms, nbytes, extra=None):
    d = {"kernel": name, "shape": shape, "ms": round(ms, 4), "alg_GB": round(nbytes / 1e9, 3),
         "GBps": round(nbytes / ms / 1e6, 1), "frac_of_measured_hbm": round(nbytes / ms / 1e6 / peak(), 3)}
    if extra:
        d.update(extra)
    print(json.dumps(d), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="full,gather,bin,gat,block,gemm")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    which = set(args.which.split(","))
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    N, NNZ, F, _ = G.SHAPES["reddit"]

    if which & {"full", "bin", "block", "gather"}:
        rp, col = G.rmat_csr(N, NNZ, seed=0, device=dev)
        deg = rp[1:] - rp[:-1]
        print(json.dumps({"graph": "reddit-shaped rmat", "N": N, "nnz": int(rp[-1]), "max_deg": int(deg.max()),
                          "rows_gt_4096": int((deg > 4096).sum()), "edges_in_rows_gt_4096": int(deg[deg > 4096].sum())}),
              flush=True)

    if "full" in which:
        for Fw, dt in ((602, torch.float32), (256, torch.float32), (602, torch.bfloat16), (256, torch.bfloat16)):
            x = G.feature_table(N, Fw, seed=1, device=dev, dtype=dt)
            b = x.element_size()
            out = torch.empty((N, (Fw + 3) // 4 * 4), device=dev)[:, :Fw]
            nbytes = NNZ * (4 + Fw * b) + N * (Fw * 4 + 8)
            plan = K.CsrPlan(rp, chunk_edges=4096)
            for fam in ("rowsplit", "stream", "wholerow"):
                K.set_option("spmm_kernel", fam)
                ms = timeit(lambda: K.spmm_csr(rp, col, x, reduce="mean", out=out, F=Fw), args.iters)
                report("spmm_full_graph[%s]" % fam, "F=%d %s" % (Fw, str(dt).split(".")[-1]), ms, nbytes)
            K.set_option("spmm_kernel", "rowsplit")
            ms = timeit(lambda: K.spmm_csr(rp, col, x, reduce="mean", out=out, F=Fw, plan=plan), args.iters)
            report("spmm_full_graph[rowsplit+plan]", "F=%d %s" % (Fw, str(dt).split(".")[-1]), ms, nbytes,
                   {"heavy_rows": plan.n_heavy_rows, "chunks": plan.n_chunks})
            K.set_option("spmm_kernel", None)
            del x, out

    if "gather" in which:
        table = G.feature_table(N, F, seed=1, device=dev)
        g = torch.Generator(device=dev).manual_seed(3)
        for M in (160000, 1000000):
            ids = torch.randint(0, N, (M,), device=dev, generator=g)
            out = torch.empty((M, table.size(1)), device=dev)
            ms = timeit(lambda: K.gather_rows(table, ids, out=out), args.iters)
            report("gather_rows[tma]", "M=%d row=%dB" % (M, table.size(1) * 4), ms, M * (8 + 2 * table.size(1) * 4))
            ms = timeit(lambda: torch.index_select(table, 0, ids, out=out), args.iters)
            report("torch.index_select (library comparator)", "M=%d" % M, ms, M * (8 + 2 * table.size(1) * 4))
        del table

    if "bin" in which:
        x = G.feature_table(N, F, seed=1, device=dev)
        packed = K.binarize_pack(x[:, :F])
        wpr = packed.size(1)
        nbytes = NNZ * (4 + 4 * wpr) + N * (4 * F + 8)
        ms_np = timeit(lambda: K.bin_spmm_csr(rp, col, packed, F, mode="mean"), args.iters)
        report("bin_spmm_csr[no plan]", "F=%d words/row=%d" % (F, wpr), ms_np, nbytes)
        bplan = K.CsrPlan(rp, chunk_edges=1024)
        ms = timeit(lambda: K.bin_spmm_csr(rp, col, packed, F, mode="mean", plan=bplan), args.iters)
        report("bin_spmm_csr[plan 1024]", "F=%d words/row=%d" % (F, wpr), ms, nbytes, {"chunks": bplan.n_chunks})
        ms_pack = timeit(lambda: K.binarize_pack(x[:, :F]), args.iters)
        report("binarize_pack", "N=%d F=%d" % (N, F), ms_pack, N * (F * 4 + wpr * 4))
        out = torch.empty((N, 604), device=dev)[:, :F]
        plan32 = K.CsrPlan(rp, chunk_edges=4096)
        ms32 = timeit(lambda: K.spmm_csr(rp, col, x, reduce="mean", out=out, F=F, plan=plan32), args.iters)
        report("spmm_full_graph fp32 (C4 comparator)", "F=%d" % F, ms32, NNZ * (4 + F * 4) + N * (F * 4 + 8),
               {"speedup_binarized_vs_fp32": round(ms32 / ms, 2)})
        del x, packed, out

    if "block" in which:
        table = G.feature_table(N, F, seed=1, device=dev)
        gen = torch.Generator(device=dev).manual_seed(5)
        for batch in (1024, 8192):
            seeds = torch.randperm(int(0.66 * N), device=dev, generator=gen)[:batch]
            b0, b1 = G.sample_blocks(rp, col, seeds, (25, 10), rng_seed=9)
            out = torch.empty((b0.num_dst, 604), device=dev)[:, :F]
            nbytes = b0.num_edges() * (4 + F * 4) + b0.num_dst * (F * 4 + 4)
            for fam in ("rowsplit", "stream", "wholerow"):
                K.set_option("spmm_kernel", fam)
                ms = timeit(lambda: K.spmm_csr(b0.row_ptr, b0.col_global, table, reduce="mean", out=out, F=F), 30)
                report("spmm_block0[%s]" % fam, "batch=%d n_dst=%d nnz=%d" % (batch, b0.num_dst, b0.num_edges()), ms, nbytes)
            K.set_option("spmm_kernel", None)
        del table

    if "gemm" in which:
        try:
            tf = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
        except Exception:
            tf = 1590.0
        g = torch.Generator(device=dev).manual_seed(6)
        hbm = peak()
        for (M, Kd, Nd) in ((160000, 602, 256), (11264, 602, 256), (11264, 128, 256), (11264, 256, 172),
                            (2449029, 100, 256), (8192, 8192, 8192)):
            ld = (Kd + 3) // 4 * 4                                     # 16-byte rows, as the feature tables are laid out
            a = torch.randn((M, ld), device=dev, generator=g)[:, :Kd]
            w = torch.randn((Nd, ld), device=dev, generator=g)[:, :Kd]  # nn.Linear layout [out, in]: K-major for x @ W^T
            go = torch.randn((M, Nd), device=dev, generator=g)
            out = torch.empty((M, Nd), device=dev)
            flop = 2.0 * M * Kd * Nd
            io_bytes = M * Kd * 4 + Kd * Nd * 4 + M * Nd * 4

            def rep(name, ms, extra=None):
                d = {"kernel": name, "shape": "%dx%dx%d" % (M, Kd, Nd), "ms": round(ms, 4), "TFLOPs": round(flop / ms / 1e9, 1),
                     "frac_of_measured_bf16_peak": round(flop / ms / 1e9 / tf, 3), "io_GBps": round(io_bytes / ms / 1e6, 1),
                     "io_frac_of_measured_hbm": round(io_bytes / ms / 1e6 / hbm, 3)}
                d.update(extra or {})
                print(json.dumps(d), flush=True)

            ref64 = None
            if M * Kd * Nd <= 2e10:
                ref64 = a.double() @ w.double().t()
            for prec in ("tf32", "tf32x3", "bf16", "fp32"):
                if prec == "fp32" and M * Kd * Nd > 4e11:
                    continue
                ms = timeit(lambda: K.gemm(a, w, out=out, trans_b=True, precision=prec), args.iters)
                extra = None
                if ref64 is not None:
                    extra = {"max_err_over_max_ref": float((out.double() - ref64).abs().max() / ref64.abs().max())}
                rep("gemm fwd x@W^T [%s]" % {"tf32": "tcgen05 tf32, TMA on fp32", "tf32x3": "tcgen05 3xTF32 (hi/lo split in smem)",
                                              "bf16": "tcgen05 bf16 + pack", "fp32": "simt fp32"}[prec], ms, extra)
            del ref64
            if M * Kd * Nd <= 4e11:
                for prec in ("tf32", "tf32x3", "bf16"):
                    dw = torch.empty((Nd, Kd), device=dev)
                    ms = timeit(lambda: K.gemm(go, a, trans_a=True, out=dw, precision=prec), args.iters)
                    rep("gemm dW = G^T X [%s]" % prec, ms)
                    dx = torch.empty((M, Kd), device=dev)
                    ms = timeit(lambda: K.gemm(go, w, out=dx, precision=prec), args.iters)
                    rep("gemm dX = G W [%s]" % prec, ms)
                    del dw, dx
            ab, wb = a.to(torch.bfloat16), w.to(torch.bfloat16)
            ms = timeit(lambda: torch.matmul(ab, wb.t()), args.iters)
            rep("torch.matmul bf16 (cuBLAS comparator, operands pre-converted, bf16 out)", ms)
            torch.backends.cuda.matmul.allow_tf32 = True
            ac, wc = a.contiguous(), w.contiguous()
            ms = timeit(lambda: torch.matmul(ac, wc.t(), out=out), args.iters)
            rep("torch.matmul fp32 operands, TF32 allowed (cuBLAS comparator, same I/O as ours)", ms)
            torch.backends.cuda.matmul.allow_tf32 = False
            ms = timeit(lambda: torch.matmul(ac, wc.t(), out=out), args.iters)
            rep("torch.matmul fp32 operands, exact fp32 (cuBLAS SGEMM comparator for fp32 / 3xTF32)", ms)
            del a, w, go, out, ab, wb, ac, wc

    if "gat" in which:
        Np, E, Fp, _ = G.SHAPES["products"]
        rp, col = G.rmat_csr(Np, 2 * E, seed=2, device=dev, symmetric=True)
        heads, D = 4, 64
        g = torch.Generator(device=dev).manual_seed(4)
        wh = torch.randn((Np, heads * D), device=dev, generator=g)
        el = torch.randn((Np, heads), device=dev, generator=g)
        er = torch.randn((Np, heads), device=dev, generator=g)
        out = torch.empty_like(wh)
        nnz = col.numel()
        nbytes = nnz * (4 + heads * D * 4 + heads * 4) + Np * (heads * D * 4 + heads * 4 + 8)
        gplan = K.CsrPlan(rp, chunk_edges=1024)
        for fam in ("group", "row"):
            K.set_option("gat_kernel", fam)
            ms = timeit(lambda: K.gat_forward(rp, col, wh, el, er, heads, 0.2, out=out), args.iters)
            report("gat_forward[%s] (fused SDDMM+softmax+SpMM)" % fam, "products-shaped N=%d nnz=%d heads=4 D=64" % (Np, nnz), ms, nbytes)
        K.set_option("gat_kernel", None)
        ms = timeit(lambda: K.gat_forward(rp, col, wh, el, er, heads, 0.2, out=out, plan=gplan), args.iters)
        report("gat_forward[row + plan 1024]", "products-shaped N=%d nnz=%d heads=4 D=64" % (Np, nnz), ms, nbytes,
               {"heavy_rows": gplan.n_heavy_rows, "chunks": gplan.n_chunks})
        gplan256 = K.CsrPlan(rp, chunk_edges=256)
        for warps in ("8", "4", "1"):
            K.set_option("gat_row_warps", warps)
            ms = timeit(lambda: K.gat_forward(rp, col, wh, el, er, heads, 0.2, out=out, plan=gplan256), args.iters)
            report("gat_forward[row + plan 256, %s warp(s) per block]" % warps, "products-shaped", ms, nbytes,
                   {"heavy_rows": gplan256.n_heavy_rows, "chunks": gplan256.n_chunks})
        K.set_option("gat_row_warps", None)
        ms = timeit(lambda: K.spmm_csr(rp, col, wh, reduce="sum", out=out), args.iters)
        report("spmm_full_graph (same graph, F=256)", "products-shaped", ms, nnz * (4 + 256 * 4) + Np * (256 * 4 + 8))
        # backward: pass 1 over the CSR (alpha, dz per edge), pass 2 over the transposed CSR (d_Wh, d_er)
        o2, rmax, rsum = K.gat_forward(rp, col, wh, el, er, heads, 0.2, save_stats=True, plan=gplan)
        trp, tcol, _, perm = K.csr_transpose(rp, col, Np, want_perm=True)
        gout = torch.randn_like(wh)
        ms_t = timeit(lambda: K.csr_transpose(rp, col, Np, want_perm=True), 3)
        # two-pass backward (round 1): pass 1 reads Wh_j + writes 8 B/(edge, head); pass 2 reads g_i + 8 B/(edge, head)
        bytes_two = nnz * (4 + heads * D * 4 + heads * 8) * 2 + Np * (heads * D * 4) * 3
        # fused single pass over CSR^T (round 2): per edge t_col + perm + one g_i row + the 64-byte destination record +
        # dz written once and read once; per node g, out (pre-pass), Wh_j, d_Wh_j
        bytes_fused = nnz * (8 + heads * D * 4 + 64 + heads * 8) + Np * (heads * D * 4) * 4 + Np * 64
        for fam, nb in (("twopass", bytes_two), ("fused", bytes_fused)):
            K.set_option("gat_bwd_kernel", fam)
            ms = timeit(lambda: K.gat_backward(rp, col, trp, tcol, perm, wh, el, er, o2, rmax, rsum, gout, heads, 0.2), 5)
            report("gat_backward[%s, no plans]" % fam, "products-shaped", ms, nb, {"csr_transpose_ms_once": round(ms_t, 3)})
            for ce in (1024, 256):
                pl, tpl = K.CsrPlan(rp, chunk_edges=ce), K.CsrPlan(trp, chunk_edges=ce)
                ms = timeit(lambda: K.gat_backward(rp, col, trp, tcol, perm, wh, el, er, o2, rmax, rsum, gout, heads, 0.2,
                                                   plan=pl, t_plan=tpl), 5)
                report("gat_backward[%s, plans %d]" % (fam, ce), "products-shaped", ms, nb,
                       {"heavy_rows": pl.n_heavy_rows, "chunks": pl.n_chunks, "t_heavy_rows": tpl.n_heavy_rows,
                        "t_chunks": tpl.n_chunks})
        K.set_option("gat_bwd_kernel", "fused")
        pl, tpl = K.CsrPlan(rp, chunk_edges=256), K.CsrPlan(trp, chunk_edges=256)
        for depth in ("4", "8"):
            K.set_option("gat_bwd_depth", depth)
            ms = timeit(lambda: K.gat_backward(rp, col, trp, tcol, perm, wh, el, er, o2, rmax, rsum, gout, heads, 0.2,
                                               plan=pl, t_plan=tpl), 5)
            report("gat_backward[fused, plans 256, %s rows in flight]" % depth, "products-shaped", ms, bytes_fused)
        K.set_option("gat_bwd_depth", None)
        K.set_option("gat_bwd_kernel", None)
        # the uniform control graph of the same (N, nnz): no skew, no plans
        del rp, col, trp, tcol, perm
        rp, col = G.uniform_csr(Np, nnz // Np, seed=3, device=dev)
        trp, tcol, _, perm = K.csr_transpose(rp, col, Np, want_perm=True)
        o2, rmax, rsum = K.gat_forward(rp, col, wh, el, er, heads, 0.2, save_stats=True)
        nz = col.numel()
        ms = timeit(lambda: K.gat_forward(rp, col, wh, el, er, heads, 0.2, out=out), args.iters)
        report("gat_forward[row]", "uniform control N=%d nnz=%d" % (Np, nz), ms,
               nz * (4 + heads * D * 4 + heads * 4) + Np * (heads * D * 4 + heads * 4 + 8))
        for fam, per_edge in (("twopass", (4 + heads * D * 4 + heads * 8) * 2), ("fused", 8 + heads * D * 4 + 64 + heads * 8)):
            K.set_option("gat_bwd_kernel", fam)
            ms = timeit(lambda: K.gat_backward(rp, col, trp, tcol, perm, wh, el, er, o2, rmax, rsum, gout, heads, 0.2), 5)
            report("gat_backward[%s]" % fam, "uniform control N=%d nnz=%d" % (Np, nz), ms,
                   nz * per_edge + Np * (heads * D * 4) * (3 if fam == "twopass" else 4))
        K.set_option("gat_bwd_kernel", None)


if __name__ == "__main__":
    main()
