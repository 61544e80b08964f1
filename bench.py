#!/usr/bin/env python
"""bench.py — the hot path of BASELINE.json config[1] on B200.

Workload (config.workload): GraphSAGE-mean 2-layer, fanout 25/10, batch 1,024 seeds, on a synthetic
Reddit-shaped graph (N=232,965, nnz=114,615,892, F=602 fp32, row stride 604).  One STEP = the neighbourhood
aggregation path of one mini-batch with pre-sampled blocks resident in HBM:
    layer 0: mean-aggregate block0 straight from the feature table (the sampled-block feature gather is fused
             into the aggregation: col ids are global), F=602      -> [n_dst0, 602]
             + TMA row gather of the dst (self) rows                -> [n_dst0, 604]
    layer 1: mean-aggregate block1 over the hidden rows, F=256      -> [1024, 256]
``value`` = algorithmic bytes of those kernels (SURVEY.md §8 d formulas) / step time, whole job, GB/s.
``e2e``   = the same metric through the public API with the blocks in pinned HOST memory: H2D of the block
            arrays + the kernels + D2H of the layer-1 aggregate, all inside the timed region.
``roofline`` = the dominant kernel (layer-0 SpMM) timed with CUDA events on its own stream, inside the timed region.
``cpu_baseline`` / ``--impl reference`` = the reference's CPU aggregation (torch.sparse.mm on a COO block, as
            dgll/nn/Convolution/gcnconv.py:31 / Evaluation/PPI/gcn_model.py:76 do) and a best-effort CSR port
            (oracle/oracle.c, OpenMP) on the same blocks, all host threads.
Multi-GPU (--gpus N under torchrun): the batch axis shards — every rank holds the graph + features and runs its own
mini-batches (the reference's data-parallel scheme, GPU Accelerator/MQGCN.py:94-157); no data-path collective;
scaling = weak.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_NODES, NNZ, FEAT, HIDDEN = 232965, 114615892, 602, 256
BATCH, FANOUTS = 1024, (25, 10)
N_BATCHES = 16  # distinct pre-sampled mini-batches cycled through the timed steps


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed ncu summary
    (profiles/r01_spmm_headline.txt, one `--set full` capture of the same command).  None when absent."""
    path = os.path.join(ROOT, "profiles", "r01_spmm_headline.txt")
    try:
        rd = wr = None
        for line in open(path):
            f = line.split()
            if len(f) >= 3 and f[0] == "dram__bytes_read.sum" and rd is None:
                rd = float(f[1]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[f[2]]
            if len(f) >= 3 and f[0] == "dram__bytes_write.sum" and wr is None:
                wr = float(f[1]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[f[2]]
            if rd is not None and wr is not None:
                return rd + wr
    except Exception:
        pass
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].startswith("Active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def alg_bytes_spmm(nnz, n_dst, F, b=4, idx=4, vals=0, rp=8):
    """nnz*(i + v + F*b) + n_dst*(F*4 + r)."""
    return nnz * (idx + vals + F * b) + n_dst * (F * 4 + rp)


def alg_bytes_gather(m, row_bytes, id_bytes=8):
    return m * (id_bytes + 2 * row_bytes)


# --------------------------------------------------------------------- CPU arm --
def cpu_blocks(n_batches, seed=0):
    """The same workload built on the host with numpy (uniform-degree control graph slice is NOT used: the blocks
    are sampled from a host copy of a seeded random Reddit-shaped neighbourhood model): per mini-batch, block0 has
    ~11K dst rows x 25 sampled neighbours over 232,965 feature rows of 602 floats; block1 1,024 x 10."""
    import numpy as np
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_batches):
        n_dst1 = BATCH
        nb1 = rng.integers(0, N_NODES, size=n_dst1 * FANOUTS[1])
        dst0 = np.unique(np.concatenate([rng.choice(N_NODES, n_dst1, replace=False), nb1]))
        n_dst0 = dst0.size
        col0 = rng.integers(0, N_NODES, size=n_dst0 * FANOUTS[0]).astype(np.int32)
        rp0 = np.arange(0, n_dst0 * FANOUTS[0] + 1, FANOUTS[0], dtype=np.int64)
        col1 = rng.integers(0, n_dst0, size=n_dst1 * FANOUTS[1]).astype(np.int32)
        rp1 = np.arange(0, n_dst1 * FANOUTS[1] + 1, FANOUTS[1], dtype=np.int64)
        out.append((rp0, col0, dst0, rp1, col1))
    return out


def cpu_step_bytes(blk):
    rp0, col0, dst0, rp1, col1 = blk
    return (alg_bytes_spmm(col0.size, rp0.size - 1, FEAT) + alg_bytes_gather(dst0.size, 604 * 4)
            + alg_bytes_spmm(col1.size, rp1.size - 1, HIDDEN))


def run_cpu_arm(steps, warmup, budget_s=25.0, which=("coo", "csr")):
    """Times the reference CPU aggregation on host cores.  Returns dict(value GB/s, per-variant numbers)."""
    import numpy as np
    import torch
    import oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((N_NODES, FEAT), dtype=np.float32)
    xt = torch.from_numpy(x)
    h1 = rng.standard_normal((12000, HIDDEN), dtype=np.float32)
    blocks = cpu_blocks(max(2, min(steps + warmup, 4)))
    res = {}

    def step_coo(blk):
        rp0, col0, dst0, rp1, col1 = blk
        rows0 = torch.from_numpy(np.repeat(np.arange(rp0.size - 1), np.diff(rp0)))
        a0 = torch.sparse_coo_tensor(torch.stack([rows0, torch.from_numpy(col0).long()]),
                                     torch.full((col0.size,), 1.0 / FANOUTS[0]), (rp0.size - 1, N_NODES))
        agg0 = torch.sparse.mm(a0, xt)                       # gcnconv.py:31 / gcn_model.py:76
        self0 = xt[torch.from_numpy(dst0)]                   # dgraph.py:105 features[nodes]
        rows1 = torch.from_numpy(np.repeat(np.arange(rp1.size - 1), np.diff(rp1)))
        h = torch.from_numpy(h1[:rp0.size - 1])
        a1 = torch.sparse_coo_tensor(torch.stack([rows1, torch.from_numpy(col1).long()]),
                                     torch.full((col1.size,), 1.0 / FANOUTS[1]), (rp1.size - 1, rp0.size - 1))
        return agg0, self0, torch.sparse.mm(a1, h)

    def step_csr(blk):
        rp0, col0, dst0, rp1, col1 = blk
        agg0 = oracle.spmm_csr(rp0, col0, x, reduce="mean")
        self0, _ = oracle.gather_rows(x, dst0)
        return agg0, self0, oracle.spmm_csr(rp1, col1, h1[:rp0.size - 1], reduce="mean")

    for name, fn in (("coo", step_coo), ("csr", step_csr)):
        if name not in which:
            continue
        for w in range(min(warmup, 2)):
            fn(blocks[w % len(blocks)])
        t0, n, nbytes = time.perf_counter(), 0, 0
        while n < steps and (time.perf_counter() - t0) < budget_s / len(which):
            blk = blocks[n % len(blocks)]
            fn(blk)
            nbytes += cpu_step_bytes(blk)
            n += 1
        dt = time.perf_counter() - t0
        res[name] = {"gbs": nbytes / dt / 1e9, "ms_per_step": dt / max(n, 1) * 1e3, "steps": n}
    best = max(res, key=lambda k: res[k]["gbs"])
    return {"value": res[best]["gbs"], "best": best, "variants": res, "cores": cores}


# --------------------------------------------------------------------- GPU arm --
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-epoch", action="store_true", help="skip the sampled-GraphSAGE training-epoch extra")
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = ("GraphSAGE-mean 2-layer fanout 25/10 batch 1024, aggregation path of one mini-batch, synthetic "
                "Reddit-shaped graph N=232965 nnz=114615892 F=602 fp32")
    metric = "SpMM aggregation GB/s (algorithmic bytes, sampled GraphSAGE-mean blocks)"

    if args.impl == "reference":
        if rank != 0:
            return
        r = run_cpu_arm(max(args.steps, 1), args.warmup, budget_s=60.0)
        v = r["variants"][r["best"]]
        line = {"impl": "reference", "metric": metric, "value": r["value"], "unit": "GB/s", "n_gpus": args.gpus,
                "steps": v["steps"], "warmup": min(args.warmup, 2), "ms_per_step": v["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": workload, "flush": "feature table 563 MB > L2"},
                "cpu_baseline": {"value": r["value"], "unit": "GB/s", "cores": r["cores"], "kind": "port",
                                 "sample": "%d mini-batches of the same block shapes; best of torch.sparse.mm COO "
                                           "(reference-faithful, gcnconv.py:31) and OpenMP CSR port (oracle.c): %s"
                                           % (v["steps"], json.dumps({k: round(x["gbs"], 2) for k, x in r["variants"].items()}))},
                "e2e": {"value": r["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    from dgll_b200 import _lib, graphs as G, kernels as K

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic inputs, resident in HBM -------------------------------------------------------------
    t_setup = time.perf_counter()
    row_ptr, col_idx = G.rmat_csr(N_NODES, NNZ, seed=args.seed, device=dev)
    LD = int(os.environ.get("BENCH_LD", "604"))
    table = G.feature_table(N_NODES, FEAT, seed=args.seed, device=dev, pad_to=LD)  # [N, 604] fp32, 563 MB > L2
    gen = torch.Generator(device=dev).manual_seed(args.seed + 1000 * rank)
    n_train = int(0.66 * N_NODES)
    perm = torch.randperm(n_train, device=dev, generator=gen)
    batches = []
    for b in range(N_BATCHES):
        seeds = perm[b * BATCH:(b + 1) * BATCH]
        blocks = G.sample_blocks(row_ptr, col_idx, seeds, FANOUTS, rng_seed=args.seed * 7919 + b + 131 * rank)
        b0, b1 = blocks
        batches.append({
            "rp0": b0.row_ptr, "col0": b0.col_global, "dst0": b0.dst_ids.contiguous(), "n_dst0": b0.num_dst,
            "rp1": b1.row_ptr, "col1": b1.col, "n_dst1": b1.num_dst,
            "agg0": torch.empty((b0.num_dst, FEAT), device=dev), "self0": torch.empty((b0.num_dst, LD), device=dev),
            "agg1": torch.empty((b1.num_dst, HIDDEN), device=dev),
            "h1": torch.randn((b0.num_dst, HIDDEN), device=dev, generator=gen),
        })
    for bt in batches:
        bt["bytes0"] = alg_bytes_spmm(bt["col0"].numel(), bt["n_dst0"], FEAT, rp=4)
        bt["bytes"] = (bt["bytes0"] + alg_bytes_gather(bt["n_dst0"], 604 * 4)
                       + alg_bytes_spmm(bt["col1"].numel(), bt["n_dst1"], HIDDEN, rp=4))
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup

    view = table[:, :FEAT]

    def step(bt, ev=None):
        if ev is not None:
            ev[0].record()
        K.spmm_csr(bt["rp0"], bt["col0"], view, reduce="mean", out=bt["agg0"])
        if ev is not None:
            ev[1].record()
        K.gather_rows(table, bt["dst0"], out=bt["self0"])
        K.spmm_csr(bt["rp1"], bt["col1"], bt["h1"], reduce="mean", out=bt["agg1"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------------------
    for w in range(args.warmup):
        step(batches[w % N_BATCHES])
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = _lib.launch_count()
    # the dominant kernel is bracketed by its own event pair on every 4th step (events cost ~1-2 us of stream time each)
    KEV = 4
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range((args.steps + KEV - 1) // KEV)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    prof = os.environ.get("BENCH_PROFILE") == "1"   # ncu --profile-from-start off: capture the timed steps only
    if prof:
        torch.cuda.profiler.start()
    e0.record()
    total_bytes = 0
    for s in range(args.steps):
        bt = batches[s % N_BATCHES]
        step(bt, kev[s // KEV] if s % KEV == 0 else None)
        total_bytes += bt["bytes"]
    e1.record()
    barrier()
    if prof:
        torch.cuda.profiler.stop()
    launches = _lib.launch_count() - l0
    ms = e0.elapsed_time(e1)
    k_ms = sum(a.elapsed_time(b) for a, b in kev) / len(kev)
    k_bytes = sum(batches[s % N_BATCHES]["bytes0"] for s in range(0, args.steps, KEV)) / len(kev)

    # ---- end to end: blocks in pinned host memory -> H2D -> kernels -> D2H of the layer-1 aggregate -----
    # Each mini-batch's block arrays (row_ptr0, col0, dst ids, row_ptr1, col1; all int32) live in ONE pinned host
    # buffer, so a step costs one H2D copy (copy stream, double-buffered: the copy of batch k+1 overlaps the kernels of
    # batch k, the way the reference's MQ-GNN queues overlap loading and compute, GPU Accelerator/buffer_queues.py:22-119),
    # the three kernels, and one D2H copy of the [1024, 256] result that the host waits for one step later.
    def pack(bt):
        parts = [bt["rp0"].to(torch.int32), bt["col0"].to(torch.int32), bt["dst0"].to(torch.int32),
                 bt["rp1"].to(torch.int32), bt["col1"].to(torch.int32)]
        offs, n = [], 0
        for t in parts:
            offs.append((n, t.numel()))
            n += (t.numel() + 3) // 4 * 4          # keep every segment 16-byte aligned
        buf = torch.zeros(n, dtype=torch.int32).pin_memory()
        for (o, m), t in zip(offs, parts):
            buf[o:o + m].copy_(t)
        return buf, offs

    host = [pack(bt) for bt in batches]
    max_len = max(h[0].numel() for h in host)
    dev_buf = [torch.empty(max_len, dtype=torch.int32, device=dev) for _ in range(2)]
    out_host = [torch.empty((BATCH, HIDDEN), dtype=torch.float32).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream()
    ev_h2d = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    for e in ev_done:
        e.record(main_stream)

    def e2e_submit(s):
        i, b = s % N_BATCHES, s % 2
        hbuf, offs = host[i]
        bt = batches[i]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_done[b])                  # device buffer b is free again
            dev_buf[b][:hbuf.numel()].copy_(hbuf, non_blocking=True)
            ev_h2d[b].record(copy_stream)
        main_stream.wait_event(ev_h2d[b])
        v = [dev_buf[b][o:o + m] for (o, m) in offs]
        K.spmm_csr(v[0], v[1], view, reduce="mean", out=bt["agg0"])
        K.gather_rows(table, v[2], out=bt["self0"])
        K.spmm_csr(v[3], v[4], bt["h1"], reduce="mean", out=bt["agg1"])
        out_host[b][:bt["n_dst1"]].copy_(bt["agg1"], non_blocking=True)
        ev_done[b].record(main_stream)
        return hbuf.numel() * 4, bt["agg1"].numel() * 4

    def e2e_run(n_steps):
        h2d = d2h = nbytes = 0
        for s in range(n_steps):
            a, b = e2e_submit(s)
            h2d += a
            d2h += b
            nbytes += batches[s % N_BATCHES]["bytes"]
            if s >= 1:
                ev_done[(s - 1) % 2].synchronize()               # the host consumes the previous step's result
        ev_done[(n_steps - 1) % 2].synchronize()
        return h2d, d2h, nbytes

    # CUDA-graph form of the same step (the C-ABI calls are stream-ordered and capture-safe): graph i = the three
    # kernels of batch i, with the D2H of the PREVIOUS batch's result and the H2D of the NEXT batch's blocks on a parallel
    # branch (PCIe is full duplex, both copies hide under the kernels).  One graph launch per step instead of ~12 host
    # calls; every step still moves one batch host->device and one result device->host.
    e2e_mode = "eager, double-buffered copy stream"
    graphs = None
    try:
        side = torch.cuda.Stream(device=dev)
        graphs = []
        torch.cuda.synchronize()
        for i in range(N_BATCHES):
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph):
                cur = torch.cuda.current_stream()
                side.wait_stream(cur)
                nb = (i + 1) % N_BATCHES
                pb = (i - 1) % N_BATCHES
                with torch.cuda.stream(side):
                    dev_buf[nb % 2][:host[nb][0].numel()].copy_(host[nb][0], non_blocking=True)
                    out_host[pb % 2][:batches[pb]["n_dst1"]].copy_(batches[pb]["agg1"], non_blocking=True)
                bt, offs = batches[i], host[i][1]
                v = [dev_buf[i % 2][o:o + m] for (o, m) in offs]
                K.spmm_csr(v[0], v[1], view, reduce="mean", out=bt["agg0"])
                K.gather_rows(table, v[2], out=bt["self0"])
                K.spmm_csr(v[3], v[4], bt["h1"], reduce="mean", out=bt["agg1"])
                cur.wait_stream(side)
            graphs.append(gph)
        e2e_mode = "CUDA graph per mini-batch (3 kernels || D2H of the previous result + H2D of the next batch)"
    except Exception as ex:  # capture unsupported: keep the eager pipeline
        graphs = None
        e2e_mode += " (graph capture failed: %s)" % type(ex).__name__
        torch.cuda.synchronize()

    def e2e_run_graphs(n_steps):
        h2d = d2h = nbytes = 0
        dev_buf[0][:host[0][0].numel()].copy_(host[0][0], non_blocking=True)   # batch 0 primes the pipeline
        for s in range(n_steps):
            i = s % N_BATCHES
            graphs[i].replay()
            ev_done[s % 2].record(main_stream)
            h2d += host[(i + 1) % N_BATCHES][0].numel() * 4
            d2h += batches[i]["agg1"].numel() * 4
            nbytes += batches[i]["bytes"]
            if s >= 1:
                ev_done[(s - 1) % 2].synchronize()           # results up to step s-2 are on the host now
        last = (n_steps - 1) % N_BATCHES                     # the final step's result has no following graph
        out_host[last % 2][:batches[last]["n_dst1"]].copy_(batches[last]["agg1"], non_blocking=True)
        main_stream.synchronize()
        return h2d, d2h, nbytes

    run = e2e_run_graphs if graphs is not None else e2e_run
    run(min(args.warmup, 5) + 2)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    h2d, d2h, e2e_bytes = run(args.steps)
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    # the graph path must produce the same result as the eager kernels (checked on the last batch run)
    if graphs is not None:
        last = (args.steps - 1) % N_BATCHES
        ref_out = K.spmm_csr(batches[last]["rp1"], batches[last]["col1"], batches[last]["h1"], reduce="mean")
        torch.cuda.synchronize()
        assert torch.equal(out_host[(args.steps - 1) % 2][:batches[last]["n_dst1"]].to(dev), ref_out), "e2e graph result"
    clk = clocks.stop() if rank == 0 else None

    # ---- extra: sampled GraphSAGE TRAINING epoch (second half of BASELINE.json's metric) ------------------------
    # 2-layer SAGE 602 -> 256 -> 41, fanout 25/10, batch 1024/GPU, Adam; every rank trains on its shard of the
    # 153,756 train seeds (first 66 % of the nodes), gradients all-reduced as one flat buffer per step.
    epoch = None
    if not args.no_epoch:
        import dgll_b200.nn as dnn
        from dgll_b200 import train as T
        torch.manual_seed(args.seed)
        labels = torch.randint(0, 41, (N_NODES,), device=dev, generator=gen)
        model = dnn.GraphSAGE(FEAT, HIDDEN, 41, 2, torch.relu, 0.0).to(dev)
        opt = torch.optim.Adam(model.parameters(), lr=0.003, fused=True)
        perm_e = torch.randperm(n_train, device=dev, generator=torch.Generator(device=dev).manual_seed(args.seed))
        shard = perm_e[rank::world].contiguous()              # use_ddp-style split of the shuffled train seeds
        res = {}
        for prec in ("fp32", "bf16"):
            T.sage_epoch(model, opt, table, labels, FEAT, row_ptr, col_idx, shard[:8 * BATCH], FANOUTS, BATCH,
                         rng_seed=1, precision=prec)                                  # warm-up: 8 batches
            barrier()
            r_e2e = T.sage_epoch(model, opt, table, labels, FEAT, row_ptr, col_idx, shard, FANOUTS, BATCH,
                                 rng_seed=2, precision=prec)                           # sampler in the loop
            pre = T.make_batches(row_ptr, col_idx, shard, FANOUTS, BATCH, rng_seed=3)
            barrier()
            r_pre = T.sage_epoch(model, opt, table, labels, FEAT, batches=pre, precision=prec)
            # the same step captured once as a CUDA graph on fixed-capacity block buffers (train.GraphedSageTrainer):
            # (a) pre-sampled blocks, (b) device sampler + block builder run eagerly, training step replayed
            g_pre = g_loop = float("nan")
            g_err = None
            try:
                opt_g = torch.optim.Adam(model.parameters(), lr=0.003, fused=True, capturable=True)
                tr = T.GraphedSageTrainer(model, opt_g, table, labels, BATCH, FANOUTS, precision=prec)
                tr.load(*pre[0])
                tr.capture()
                barrier()
                g_pre = tr.epoch(pre)["time_s"]
                barrier()
                g_loop = tr.epoch(T.iter_batches(row_ptr, col_idx, shard, FANOUTS, BATCH, rng_seed=4))["time_s"]
                del tr, opt_g
            except Exception as ex:                                  # report, keep the eager numbers
                g_err = "%s: %s" % (type(ex).__name__, str(ex)[:200])
            del pre
            t = torch.tensor([r_e2e["time_s"], r_pre["time_s"], g_pre, g_loop], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[prec] = {"epoch_s_sampler_in_loop": round(t[0].item(), 4), "epoch_s_presampled": round(t[1].item(), 4),
                         "epoch_s_presampled_cuda_graph": round(t[2].item(), 4),
                         "epoch_s_sampler_in_loop_cuda_graph": round(t[3].item(), 4),
                         "batches_per_gpu": r_e2e["n_batches"], "loss": round(r_e2e["loss"], 4)}
            if g_err:
                res[prec]["cuda_graph_error"] = g_err
        epoch = {"model": "GraphSAGE-mean 2-layer 602-256-41, fanout 25/10, batch 1024/GPU, Adam, fwd+bwd+step",
                 "train_seeds": int(perm.numel()), "gemm": res}

    # ---- max over ranks ----------------------------------------------------------------------------------
    stats = torch.tensor([ms, e2e_ms, float(total_bytes), float(e2e_bytes), float(launches)], device=dev,
                         dtype=torch.float64)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_ms = mx[0].item(), mx[1].item()
        total_bytes, e2e_bytes, launches = sm[2].item(), sm[3].item(), int(sm[4].item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    value = total_bytes / (ms * 1e-3) / 1e9
    achieved = k_bytes / (k_ms * 1e-3) / 1e9
    line = {
        "metric": metric, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "batch": BATCH, "fanouts": list(FANOUTS), "hidden": HIDDEN,
                   "parallelism": "dp%d (batch axis; graph+features replicated)" % world,
                   "flush": "inputs larger than L2: 563 MB feature table, %d distinct mini-batches cycled" % N_BATCHES,
                   "block0": {"n_dst": batches[0]["n_dst0"], "nnz": batches[0]["col0"].numel()},
                   "setup_s": round(setup_s, 1)},
        "roofline": {"bound": "hbm", "kernel": "spmm_rowslab_kernel<float,4,32> (layer-0 mean aggregation, F=602)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(), "traffic_source": "profiles/r01_spmm_headline.txt (ncu --set full)",
                     "peak_source": peak_src, "kernel_ms": k_ms,
                     "algorithmic_bytes_per_launch": k_bytes},
        "e2e": {"value": e2e_bytes / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d / args.steps,
                "d2h_bytes_per_step": d2h / args.steps, "ms_per_step": e2e_ms / args.steps, "mode": e2e_mode},
        "gpu_launches": launches, "clocks": clk, "epoch": epoch,
    }
    if not args.no_cpu_baseline and world == 1:
        r = run_cpu_arm(6, 1, budget_s=24.0)
        line["cpu_baseline"] = {
            "value": r["value"], "unit": "GB/s", "cores": r["cores"], "kind": "port",
            "sample": "<=6 mini-batches of the same block shapes per variant, ~12 s each; best=%s; %s" % (
                r["best"], json.dumps({k: round(x["gbs"], 2) for k, x in r["variants"].items()}))}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
