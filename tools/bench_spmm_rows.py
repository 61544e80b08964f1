#!/usr/bin/env python
"""Sweep of the SpMM kernel families on sampled blocks (the headline shape of bench.py and its neighbours).

For each (shape, family, options) prints one JSON line: median kernel ms over distinct mini-batches (CUDA events, the
563 MB table exceeds L2 and 16 mini-batches are cycled), algorithmic GB/s (SURVEY.md §8 d) and the fraction of the
measured HBM peak.  Used to choose the defaults of csrc/spmm_rows.cu; summary in profiles/r02_spmm_rows_sweep.md.

  python tools/bench_spmm_rows.py [--quick]
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


def time_batches(fn, batches, rounds=4):
    """fn(batch) timed per launch; returns the mean of per-launch times over rounds x batches (after one warm round)."""
    for bt in batches:
        fn(bt)
    torch.cuda.synchronize()
    evs = []
    for _ in range(rounds):
        for bt in batches:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn(bt)
            b.record()
            evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return sum(ts) / len(ts), ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    N, NNZ, F, _ = G.SHAPES["reddit"]
    rp, col = G.rmat_csr(N, NNZ, seed=0, device=dev)
    gen = torch.Generator(device=dev).manual_seed(0)
    perm = torch.randperm(int(0.66 * N), device=dev, generator=gen)

    def blocks(batch, n_batches, fanouts=(25, 10)):
        out = []
        for b in range(n_batches):
            seeds = perm[b * batch:(b + 1) * batch]
            b0, b1 = G.sample_blocks(rp, col, seeds, fanouts, rng_seed=b + 1)
            out.append((b0, b1))
        return out

    def run(name, shape, fn, batches, nbytes, opts):
        K.set_option("spmm_kernel", opts.get("family"))
        for k in ("rows_tb", "rows_ns", "rows_d", "rows_stream", "spmm_tb"):
            K.set_option(k, opts.get(k))
        mean, med = time_batches(fn, batches)
        print(json.dumps({"case": name, "shape": shape, **{k: v for k, v in opts.items() if v is not None},
                          "ms_mean": round(mean, 4), "ms_median": round(med, 4), "alg_MB": round(nbytes / 1e6, 1),
                          "GBps": round(nbytes / mean / 1e6, 1), "frac": round(nbytes / mean / 1e6 / peak(), 3)}),
              flush=True)
        for k in ("spmm_kernel", "rows_tb", "rows_ns", "rows_d", "rows_stream", "spmm_tb"):
            K.set_option(k, None)

    variants = [dict(family="rowsplit"), dict(family="stream"), dict(family="wholerow", rows_stream=1)]
    for st in (0, 2, 3, 4, 5, 6, 8):
        variants.append(dict(family="wholerow", rows_stream=st))
    for tb in (32, 128):
        variants.append(dict(family="wholerow", rows_stream=0, rows_tb=tb))
    variants.append(dict(family="wholerow", rows_stream=0, rows_ns=3))
    variants.append(dict(family="wholerow", rows_stream=0, rows_d=2))
    variants.append(dict(family="wholerow", rows_stream=0, rows_d=8))

    # ---- headline: block0 of batch 1024, F=602 fp32 (ld 604), output ld 602 and 604 -----------------------------
    table = G.feature_table(N, F, seed=0, device=dev, pad_to=604)
    bl = blocks(1024, 16)
    nbytes = sum(b0.num_edges() * (4 + F * 4) + b0.num_dst * (F * 4 + 4) for b0, _ in bl) / len(bl)
    for ldo in (602, 604):
        outs = [torch.empty((b0.num_dst, ldo), device=dev)[:, :F] for b0, _ in bl]
        items = list(zip(bl, outs))
        for v in variants:
            if ldo == 602 and v.get("rows_stream") not in (None, 0, 1):
                continue
            run("block0 F=602 fp32", "batch=1024 ldo=%d n_dst~%d nnz~%d" % (ldo, bl[0][0].num_dst, bl[0][0].num_edges()),
                lambda it: K.spmm_csr(it[0][0].row_ptr, it[0][0].col_global, table, reduce="mean", out=it[1], F=F),
                items, nbytes, v)
    # ---- bf16 storage of the same table ---------------------------------------------------------------------------
    tb16 = G.feature_table(N, F, seed=0, device=dev, dtype=torch.bfloat16)
    nbytes16 = sum(b0.num_edges() * (4 + F * 2) + b0.num_dst * (F * 4 + 4) for b0, _ in bl) / len(bl)
    outs = [torch.empty((b0.num_dst, 604), device=dev)[:, :F] for b0, _ in bl]
    items = list(zip(bl, outs))
    for v in [dict(family="rowsplit"), dict(family="stream"), dict(family="wholerow", rows_stream=1),
              dict(family="wholerow", rows_stream=0), dict(family="wholerow", rows_stream=4), dict(family="wholerow", rows_stream=8),
              dict(family="wholerow", rows_stream=0, rows_d=8), dict(family="wholerow", rows_stream=0, rows_ns=2)]:
        run("block0 F=602 bf16", "batch=1024 ld=%d" % tb16.size(1),
            lambda it: K.spmm_csr(it[0][0].row_ptr, it[0][0].col_global, tb16, reduce="mean", out=it[1], F=F),
            items, nbytes16, v)
    del tb16
    # ---- block1 (hidden rows, F=256), small and L2 resident -------------------------------------------------------
    hs = [torch.randn((b0.num_dst, 256), device=dev) for b0, _ in bl]
    outs = [torch.empty((b1.num_dst, 256), device=dev) for _, b1 in bl]
    items = list(zip(bl, hs, outs))
    nb1 = sum(b1.num_edges() * (4 + 256 * 4) + b1.num_dst * (256 * 4 + 4) for _, b1 in bl) / len(bl)
    for v in [dict(family="rowsplit"), dict(family="wholerow", rows_stream=1), dict(family="wholerow", rows_stream=0),
              dict(family="wholerow", rows_stream=2)]:
        run("block1 F=256 fp32", "batch=1024 n_dst=1024",
            lambda it: K.spmm_csr(it[0][1].row_ptr, it[0][1].col, it[1], reduce="mean", out=it[2]), items, nb1, v)
    del hs, outs
    # ---- large block (batch 8192): stream vs whole-row ------------------------------------------------------------
    if not args.quick:
        bl8 = blocks(8192, 4)
        nb8 = sum(b0.num_edges() * (4 + F * 4) + b0.num_dst * (F * 4 + 4) for b0, _ in bl8) / len(bl8)
        outs = [torch.empty((b0.num_dst, 604), device=dev)[:, :F] for b0, _ in bl8]
        items = list(zip(bl8, outs))
        for v in [dict(family="rowsplit"), dict(family="stream"), dict(family="wholerow", rows_stream=1),
                  dict(family="wholerow", rows_stream=0), dict(family="wholerow", rows_stream=16)]:
            run("block0 F=602 fp32 batch 8192", "nnz~%d" % bl8[0][0].num_edges(),
                lambda it: K.spmm_csr(it[0][0].row_ptr, it[0][0].col_global, table, reduce="mean", out=it[1], F=F),
                items, nb8, v)
        del bl8, outs
    del table
    # ---- F=128 (papers100M-shaped rows of 512 bytes) on the same blocks ---------------------------------------------
    t128 = G.feature_table(N, 128, seed=0, device=dev)
    outs = [torch.empty((b0.num_dst, 128), device=dev) for b0, _ in bl]
    items = list(zip(bl, outs))
    nb128 = sum(b0.num_edges() * (4 + 128 * 4) + b0.num_dst * (128 * 4 + 4) for b0, _ in bl) / len(bl)
    for v in [dict(family="rowsplit"), dict(family="wholerow", rows_stream=1), dict(family="wholerow", rows_stream=0),
              dict(family="wholerow", rows_stream=2), dict(family="wholerow", rows_stream=4), dict(family="wholerow", rows_stream=0, rows_d=8)]:
        run("block0 F=128 fp32", "batch=1024 (table 119 MB: L2 resident)",
            lambda it: K.spmm_csr(it[0][0].row_ptr, it[0][0].col_global, t128, reduce="mean", out=it[1]), items, nb128, v)
        # the same through the sharded entry point with ONE shard (what the partitioned trainer launches on 1 GPU)
    ptrs = torch.tensor([t128.data_ptr()], dtype=torch.int64, device=dev)
    mean, med = time_batches(lambda it: K.spmm_csr_sharded(it[0][0].row_ptr, it[0][0].col_global, ptrs, N, 128 * 4, 128,
                                                           out=it[1]), items)
    print(json.dumps({"case": "block0 F=128 fp32 sharded entry, 1 shard", "ms_mean": round(mean, 4),
                      "GBps": round(nb128 / mean / 1e6, 1)}), flush=True)
    ref = K.spmm_csr(bl[0][0].row_ptr, bl[0][0].col_global, t128, reduce="mean")
    got = K.spmm_csr_sharded(bl[0][0].row_ptr, bl[0][0].col_global, ptrs, N, 128 * 4, 128)
    print(json.dumps({"sharded_equals_plain": bool(torch.equal(ref, got))}), flush=True)


if __name__ == "__main__":
    main()
