#!/usr/bin/env python
"""LADIES / FastGCN layer-wise sampling on a Reddit-shaped graph — the device pipeline
(csrc/layerwise.cu) next to the host scipy pipeline the reference's scripts run (oracle.layerwise.scipy_ladies_batch).

  python tools/bench_layerwise.py [--batches 50] [--cpu-batches 3] [--scale 1.0]

Batch 1,024 seeds, fanouts n_samp * 2**l = [512, 1024] (MQLadiesFlat.py:30-32).  Prints one JSON line per sampler:
ms per mini-batch (wall clock around sampler.sample, which ends in a read-back, so device time is included), the
per-stage device times (CUDA events), and the host baseline."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from dgll_b200 import graphs as G, kernels as K  # noqa: E402
import dgll_b200.data as D  # noqa: E402


def stage_times(s, batch, iters):
    """Per-stage device time of one LADIES layer-0 + layer-1 pass, events around each C-ABI call."""
    names = ["slice_rows", "col_sqsum", "weighted_choice", "importance_scale", "select_cols"]
    acc = dict.fromkeys(names, 0.0)
    for it in range(iters):
        prev = batch
        for l, fanout in enumerate(s.fanouts):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            ev[0].record()
            q_rp, q_col, q_val = K.csr_slice_rows(s.lap_rp, s.lap_col, s.lap_val, prev)
            ev[1].record()
            if s.kind == "ladies":
                cc, cp, stats = K.col_sqsum(q_col, q_val, s.num_nodes, s.flat)
            else:
                cc, cp = None, s.prob
            ev[2].record()
            sel, picks, count = K.weighted_choice(cc, cp, fanout, seed=it * 7 + l)
            ev[3].record()
            scale = K.importance_scale(cp, sel, count, s.num_nodes, "wrs")
            ev[4].record()
            K.scatter_pos(s._pos, picks, count)
            b_rp, b_col, b_val = K.csr_select_cols(q_rp, q_col, q_val, s._pos, scale)
            K.scatter_pos(s._pos, picks, count, reset=True)
            ev[5].record()
            torch.cuda.synchronize()
            for k, nm in enumerate(names):
                acc[nm] += ev[k].elapsed_time(ev[k + 1])
            prev = picks[:int(count.item())]
    return {k: round(v / iters, 4) for k, v in acc.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, default=50)
    ap.add_argument("--cpu-batches", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--batch", type=int, default=1024)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    N0, NNZ0, _, _ = G.SHAPES["reddit"]
    N, NNZ = int(N0 * args.scale), int(NNZ0 * args.scale)
    rp, col = G.rmat_csr(N, NNZ, seed=0, device=dev)
    fanouts = [512, 1024]
    gen = torch.Generator(device=dev).manual_seed(5)
    seeds = torch.randperm(N, device=dev, generator=gen)
    for cls, kw in (("LadiesFlat", {"flat": True}), ("Ladies", {}), ("FastGCNSamplerFlatWrs", {"flat": True, "wrs": True})):
        t0 = time.perf_counter()
        s = getattr(D, cls)(fanouts, (rp, col), rng_seed=1, **kw)
        torch.cuda.synchronize()
        setup_s = time.perf_counter() - t0
        for w in range(3):
            s.sample(None, seeds[w * args.batch:(w + 1) * args.batch])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        edges = srcs = 0
        for b in range(args.batches):
            inp, _, blocks = s.sample(None, seeds[(3 + b) * args.batch:(4 + b) * args.batch])
            edges += sum(bl.num_edges() for bl in blocks)
            srcs += inp.numel()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / args.batches
        st = stage_times(s, seeds[:args.batch], 5)
        line = {"sampler": cls, "graph": "reddit-shaped rmat N=%d nnz=%d" % (N, int(rp[-1])), "batch": args.batch,
                "fanouts": fanouts, "ms_per_batch_device_pipeline": round(ms, 3), "stage_ms_both_layers": st,
                "block_edges_per_batch": edges / args.batches, "input_nodes_per_batch": srcs / args.batches,
                "laplacian_setup_s": round(setup_s, 2), "lap_nnz": int(s.lap_col.numel())}
        if cls.startswith("Ladies") and args.cpu_batches > 0:
            import scipy.sparse as sp
            from oracle import layerwise as LW
            t0 = time.perf_counter()
            lap = sp.csr_matrix((s.lap_val.cpu().numpy(), s.lap_col.cpu().numpy(), s.lap_rp.cpu().numpy()), shape=(N, N))
            np.random.seed(0)
            LW.scipy_ladies_batch(lap, seeds[:args.batch].cpu().numpy(), fanouts, flat=s.flat)   # warm
            t0 = time.perf_counter()
            for b in range(args.cpu_batches):
                LW.scipy_ladies_batch(lap, seeds[(3 + b) * args.batch:(4 + b) * args.batch].cpu().numpy(), fanouts,
                                      flat=s.flat)
            cpu_ms = (time.perf_counter() - t0) * 1e3 / args.cpu_batches
            line["ms_per_batch_host_scipy"] = round(cpu_ms, 2)
            line["host_cores"] = os.cpu_count()
            line["speedup_vs_host"] = round(cpu_ms / ms, 1)
            del lap
        print(json.dumps(line), flush=True)
        del s


if __name__ == "__main__":
    main()
