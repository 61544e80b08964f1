#!/usr/bin/env python
"""bench.py — the hot path of BASELINE.json config[1] on B200.

Workload (config.workload): GraphSAGE-mean 2-layer, fanout 25/10, batch 1,024 seeds, on a synthetic
Reddit-shaped graph (N=232,965, nnz=114,615,892, F=602 fp32, row stride 604).  One STEP = the neighbourhood
aggregation path of one mini-batch with pre-sampled blocks resident in HBM:
    layer 0: mean-aggregate block0 straight from the feature table (the sampled-block feature gather is fused
             into the aggregation: col ids are global), F=602      -> [n_dst0, 602] (row stride 604)
             + TMA row gather of the dst (self) rows                -> [n_dst0, 604]
    layer 1: mean-aggregate block1 over the hidden rows, F=256      -> [1024, 256]
``value`` = algorithmic bytes of those kernels (SURVEY.md §8 d formulas) / step time, whole job, GB/s.
``e2e``   = the same metric through the public API with the blocks in pinned HOST memory: H2D of the block
            arrays + the kernels + D2H of the layer-1 aggregate, all inside the timed region.
``roofline`` = the dominant kernel (layer-0 SpMM) timed with CUDA events on its own stream, inside the timed region;
            ``frac`` on algorithmic bytes, ``dram_frac`` on the DRAM bytes ncu measured for the same launch.
``cpu_baseline`` / ``--impl reference`` = the reference's CPU aggregation (torch.sparse.mm on a COO block, as
            dgll/nn/Convolution/gcnconv.py:31 / Evaluation/PPI/gcn_model.py:76 do) and a best-effort CSR port
            (oracle/oracle.c, OpenMP) on THE SAME seeded blocks and feature table, all host threads.
``gpu_baseline`` = the library call behind the reference's layers on the same box and inputs: torch.sparse.mm on a
            CUDA CSR matrix (cuSPARSE).
Extras (do not change ``value``):
  ``bf16_table``   the same step with the feature table stored in bf16 (row stride 608): half the bytes per edge.
  ``epoch``        sampled-GraphSAGE TRAINING epoch on the Reddit-shaped graph (second half of BASELINE's metric).
  ``partitioned``  BASELINE configs[4]: papers100M-shaped R-MAT graph, features node-range partitioned over the N GPUs,
                   one CUDA graph per mini-batch (dgll_b200.pipelined), halo rows read over NVLink by the aggregation
                   kernel itself ("peer") and, for comparison, exchanged with NCCL all_to_all ("nccl").
Multi-GPU (--gpus N under torchrun): the headline step shards over the batch axis — every rank holds the graph +
features and runs its own mini-batches (the reference's data-parallel scheme, GPU Accelerator/MQGCN.py:94-157); no
data-path collective; scaling = weak.  The partitioned extra is the path with a real exchange step.
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
LD32, LD16 = 604, 608                       # row strides (elements) of the fp32 / bf16 feature table: 16-byte rows
BATCH, FANOUTS = 1024, (25, 10)
N_BATCHES = 16  # distinct pre-sampled mini-batches cycled through the timed steps
NCU_SUMMARY = os.path.join("profiles", "r02_spmm_headline.txt")
WORKLOAD = ("GraphSAGE-mean 2-layer fanout 25/10 batch 1024, aggregation path of one mini-batch, synthetic "
            "Reddit-shaped graph N=232965 nnz=114615892 F=602 fp32")
METRIC = "SpMM aggregation GB/s (algorithmic bytes, sampled GraphSAGE-mean blocks)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed ncu summary
    (one `--set full` capture of the same command).  None when absent."""
    for rel in (NCU_SUMMARY, os.path.join("profiles", "r01_spmm_headline.txt")):
        path = os.path.join(ROOT, rel)
        try:
            rd = wr = None
            for line in open(path):
                f = line.split()
                if len(f) >= 3 and f[0] == "dram__bytes_read.sum" and rd is None:
                    rd = float(f[1]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[f[2]]
                if len(f) >= 3 and f[0] == "dram__bytes_write.sum" and wr is None:
                    wr = float(f[1]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[f[2]]
                if rd is not None and wr is not None:
                    return rd + wr, rel
        except Exception:
            pass
    return None, None


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


def bench_config(world, n_dst0, nnz0):
    """The ``config`` object — identical in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "batch": BATCH, "fanouts": list(FANOUTS), "hidden": HIDDEN,
            "parallelism": "dp%d (batch axis; graph+features replicated)" % world,
            "flush": "inputs larger than L2: 563 MB feature table, %d distinct mini-batches cycled" % N_BATCHES,
            "block0": {"n_dst": n_dst0, "nnz": nnz0},
            "inputs": "R-MAT graph + N(0,1) features + device-sampled blocks, all from seed 0 (same arrays in both arms)"}


# ------------------------------------------------------------------ seeded inputs --
def make_inputs(seed, rank, dev):
    """Graph, feature table and the N_BATCHES sampled mini-batches of the workload, on ``dev``.  Both arms call this
    with the same seed, so the CPU arm times the reference path on exactly the arrays the GPU arm aggregates."""
    import torch
    from dgll_b200 import graphs as G
    row_ptr, col_idx = G.rmat_csr(N_NODES, NNZ, seed=seed, device=dev)
    table = G.feature_table(N_NODES, FEAT, seed=seed, device=dev, pad_to=LD32)     # [N, 604] fp32, 563 MB > L2
    gen = torch.Generator(device=dev).manual_seed(seed + 1000 * rank)
    n_train = int(0.66 * N_NODES)
    perm = torch.randperm(n_train, device=dev, generator=gen)
    batches = []
    for b in range(N_BATCHES):
        seeds = perm[b * BATCH:(b + 1) * BATCH]
        b0, b1 = G.sample_blocks(row_ptr, col_idx, seeds, FANOUTS, rng_seed=seed * 7919 + b + 131 * rank)
        batches.append({"rp0": b0.row_ptr, "col0": b0.col_global, "dst0": b0.dst_ids.contiguous(), "n_dst0": b0.num_dst,
                        "rp1": b1.row_ptr, "col1": b1.col, "n_dst1": b1.num_dst})
    return row_ptr, col_idx, table, gen, batches


# --------------------------------------------------------------------- CPU arm --
def numpy_blocks(n_batches, seed=0):
    """Fallback inputs when no CUDA device is visible (CPU-only containers): uniform-random blocks of the workload's
    shapes.  On the GPU box the CPU arm uses ``make_inputs`` instead — the same arrays as the GPU arm."""
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


def host_inputs(seed, n_blocks):
    """(x float32[N, 602] on the host, blocks, origin string) for the CPU arm."""
    import numpy as np
    try:
        import torch
        if torch.cuda.is_available():
            torch.cuda.set_device(0)
            dev = torch.device("cuda", 0)
            _, _, table, _, batches = make_inputs(seed, 0, dev)
            x = table[:, :FEAT].contiguous().cpu().numpy()
            blocks = [(bt["rp0"].cpu().numpy().astype(np.int64), bt["col0"].cpu().numpy().astype(np.int32),
                       bt["dst0"].cpu().numpy().astype(np.int64), bt["rp1"].cpu().numpy().astype(np.int64),
                       bt["col1"].cpu().numpy().astype(np.int32)) for bt in batches[:n_blocks]]
            del table, batches
            torch.cuda.empty_cache()
            return x, blocks, "same seeded graph, feature table and sampled blocks as the GPU arm (generated on the device, copied to the host before timing)"
    except Exception:
        pass
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((N_NODES, FEAT), dtype=np.float32), numpy_blocks(n_blocks, seed),
            "no CUDA device visible: numpy uniform-random blocks of the workload's shapes")


def cpu_step_bytes(blk):
    rp0, col0, dst0, rp1, col1 = blk
    return (alg_bytes_spmm(col0.size, rp0.size - 1, FEAT, rp=4) + alg_bytes_gather(dst0.size, LD32 * 4)
            + alg_bytes_spmm(col1.size, rp1.size - 1, HIDDEN, rp=4))


def run_cpu_arm(steps, warmup, budget_s=25.0, which=("coo", "csr"), seed=0, n_blocks=N_BATCHES):
    """Times the reference CPU aggregation on host cores.  Returns dict(value GB/s, per-variant numbers)."""
    import numpy as np
    import torch
    import oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x, blocks, origin = host_inputs(seed, n_blocks)
    xt = torch.from_numpy(x)
    h1 = np.random.default_rng(seed).standard_normal((BATCH * (1 + FANOUTS[1]), HIDDEN), dtype=np.float32)
    res = {}

    def step_coo(blk):
        # the reference rebuilds the COO adjacency inside every forward (Evaluation/PPI/gcn_model.py:68-76): timed
        rp0, col0, dst0, rp1, col1 = blk
        rows0 = torch.from_numpy(np.repeat(np.arange(rp0.size - 1), np.diff(rp0)))
        deg0 = np.maximum(np.diff(rp0), 1).astype(np.float32)
        a0 = torch.sparse_coo_tensor(torch.stack([rows0, torch.from_numpy(col0).long()]),
                                     torch.from_numpy(np.repeat(1.0 / deg0, np.diff(rp0))), (rp0.size - 1, N_NODES))
        agg0 = torch.sparse.mm(a0, xt)                       # gcnconv.py:31 / gcn_model.py:76
        self0 = xt[torch.from_numpy(dst0)]                   # dgraph.py:105 features[nodes]
        rows1 = torch.from_numpy(np.repeat(np.arange(rp1.size - 1), np.diff(rp1)))
        deg1 = np.maximum(np.diff(rp1), 1).astype(np.float32)
        h = torch.from_numpy(h1[:rp0.size - 1])
        a1 = torch.sparse_coo_tensor(torch.stack([rows1, torch.from_numpy(col1).long()]),
                                     torch.from_numpy(np.repeat(1.0 / deg1, np.diff(rp1))), (rp1.size - 1, rp0.size - 1))
        return agg0, self0, torch.sparse.mm(a1, h)

    def step_csr(blk):
        rp0, col0, dst0, rp1, col1 = blk
        agg0 = oracle.spmm_csr(rp0, col0, x, reduce="mean")
        self0, _ = oracle.gather_rows(x, dst0)
        return agg0, self0, oracle.spmm_csr(rp1, col1, h1[:rp0.size - 1], reduce="mean")

    for name, fn in (("coo", step_coo), ("csr", step_csr)):
        if name not in which:
            continue
        for w in range(warmup):
            fn(blocks[w % len(blocks)])
        t0, n, nbytes = time.perf_counter(), 0, 0
        while n < steps and (n == 0 or (time.perf_counter() - t0) < budget_s / len(which)):
            blk = blocks[n % len(blocks)]
            fn(blk)
            nbytes += cpu_step_bytes(blk)
            n += 1
        dt = time.perf_counter() - t0
        res[name] = {"gbs": nbytes / dt / 1e9, "ms_per_step": dt / max(n, 1) * 1e3, "steps": n}
    best = max(res, key=lambda k: res[k]["gbs"])
    return {"value": res[best]["gbs"], "best": best, "variants": res, "cores": cores, "origin": origin,
            "block0": {"n_dst": int(blocks[0][0].size - 1), "nnz": int(blocks[0][1].size)}}


def reference_arm(args):
    r = run_cpu_arm(max(args.steps, 1), args.warmup, budget_s=120.0, seed=args.seed)
    v = r["variants"][r["best"]]
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "GB/s", "n_gpus": args.gpus,
            "steps": v["steps"], "warmup": args.warmup, "ms_per_step": v["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": bench_config(args.gpus, r["block0"]["n_dst"], r["block0"]["nnz"]),
            "cpu_baseline": {"value": r["value"], "unit": "GB/s", "cores": r["cores"], "kind": "port",
                             "sample": "%d steps over the %d mini-batches of the workload; inputs: %s; best of "
                                       "torch.sparse.mm on a COO block rebuilt inside every step as the reference's forward "
                                       "does (gcnconv.py:31, gcn_model.py:68-76) and the OpenMP CSR port (oracle.c): %s"
                                       % (v["steps"], N_BATCHES, r["origin"],
                                          json.dumps({k: round(x["gbs"], 2) for k, x in r["variants"].items()}))},
            "e2e": {"value": r["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------- extras --
def extra_epoch(args, rank, world, dev, row_ptr, col_idx, table, gen, barrier):
    """Sampled GraphSAGE TRAINING epoch, Reddit-shaped: 2-layer SAGE 602 -> 256 -> 41, fanout 25/10, batch 1024/GPU,
    Adam; every rank trains on its shard of the 153,756 train seeds (first 66 % of the nodes, use_ddp-style split),
    gradients all-reduced as one flat buffer per step."""
    import torch
    import torch.distributed as dist
    import dgll_b200.nn as dnn
    from dgll_b200 import pipelined as PL, train as T
    n_train = int(0.66 * N_NODES)
    torch.manual_seed(args.seed)
    labels = torch.randint(0, 41, (N_NODES,), device=dev, generator=gen)
    model = dnn.GraphSAGE(FEAT, HIDDEN, 41, 2, torch.relu, 0.0).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=0.003, fused=True)
    perm_e = torch.randperm(n_train, device=dev, generator=torch.Generator(device=dev).manual_seed(args.seed))
    shard = perm_e[rank::world].contiguous()
    res = {}
    for prec in ("fp32", "tf32", "bf16"):
        out = {}
        T.sage_epoch(model, opt, table, labels, FEAT, row_ptr, col_idx, shard[:8 * BATCH], FANOUTS, BATCH,
                     rng_seed=1, precision=prec)                                  # warm-up: 8 batches
        barrier()
        r_e2e = T.sage_epoch(model, opt, table, labels, FEAT, row_ptr, col_idx, shard, FANOUTS, BATCH,
                             rng_seed=2, precision=prec)                           # Python-dispatched, sampler in the loop
        out["epoch_s_sampler_in_loop"] = r_e2e["time_s"]
        out["loss"] = round(r_e2e["loss"], 4)
        g_err = None
        g_loop = g_loop2 = float("nan")
        try:
            # ONE graph launch per mini-batch: sampler, block builder, fused gather+aggregation, training step, flat
            # all-reduce and Adam inside two ping-pong CUDA graphs (dgll_b200.pipelined)
            opt_g = torch.optim.Adam(model.parameters(), lr=0.003, fused=True, capturable=True)
            tr = PL.PipelinedSageTrainer(model, opt_g, labels, row_ptr, col_idx, FEAT, table=table, batch_size=BATCH,
                                         fanouts=FANOUTS, precision=prec, rng_seed=4, max_seeds=shard.numel())
            tr.set_seeds(shard)
            tr.capture()
            tr.epoch(shard[:4 * BATCH])
            barrier()
            g_loop = tr.epoch(shard)["time_s"]
            barrier()
            g_loop2 = tr.epoch(shard)["time_s"]
            for p in model.parameters():
                p.grad = None
            del tr, opt_g
        except Exception as ex:
            g_err = "%s: %s" % (type(ex).__name__, str(ex)[:200])
        t = torch.tensor([out["epoch_s_sampler_in_loop"], g_loop, g_loop2], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["epoch_s_sampler_in_loop"] = round(t[0].item(), 4)
        out["epoch_s_sampler_in_loop_cuda_graph"] = round(min(t[1].item(), t[2].item()), 5)
        out["epoch_s_sampler_in_loop_cuda_graph_runs"] = [round(t[1].item(), 5), round(t[2].item(), 5)]
        out["batches_per_gpu"] = r_e2e["n_batches"]
        if g_err:
            out["cuda_graph_error"] = g_err
        res[prec] = out
    return {"model": "GraphSAGE-mean 2-layer 602-256-41, fanout 25/10, batch 1024/GPU, Adam, fwd+bwd+step",
            "train_seeds": n_train, "scaling": "strong (the 153,756 train seeds are split over the ranks)",
            "cuda_graph": "one launch per mini-batch: device sampler + block builder + fused gather/aggregation || "
                          "training step + flat all-reduce + Adam (dgll_b200.pipelined)",
            "gemm": res}


def extra_partitioned(args, rank, world, dev, barrier):
    """BASELINE configs[4]: GraphSAGE on a papers100M-shaped graph, features node-range partitioned over the ranks."""
    import torch
    import torch.distributed as dist
    import dgll_b200.nn as dnn
    from dgll_b200 import graphs as G, parallel as P, pipelined as PL, train as T
    N0, NNZ0, F, C = G.SHAPES["papers100m"]
    scale = float(os.environ.get("BENCH_C5_SCALE", "1.0"))
    N, NNZ = int(N0 * scale), int(NNZ0 * scale)
    t0 = time.perf_counter()
    row_ptr, col = G.rmat_csr_large(N, NNZ, seed=args.seed, device=dev)            # topology replicated on every rank
    deg = row_ptr[1:] - row_ptr[:-1]
    max_deg = int(deg.max().item())
    del deg
    lo, hi = P.local_range(rank, N, world)
    table = G.feature_table(hi - lo, F, seed=100 + rank, device=dev)               # this rank's feature rows, fp32
    labels = torch.randint(0, C, (hi - lo,), device=dev, generator=torch.Generator(device=dev).manual_seed(1 + rank))
    n_train = int(1207179 * scale)
    per_rank = max(n_train // world, 64 * BATCH)
    seeds_any = lo + torch.randperm(hi - lo, device=dev, generator=torch.Generator(device=dev).manual_seed(7 + rank))[:per_rank]
    # Training seeds: papers100M's labelled nodes are arXiv papers with real citation lists, while a uniformly drawn node
    # of an R-MAT graph of this size mostly has 0-2 in-neighbours (a 1,024-seed batch then carries ~10 K input-layer edges
    # instead of ~200 K).  The seeds are therefore drawn uniformly among this rank's nodes with in-degree >= fanout[1],
    # so that the output-layer block is full-size; the any-node variant is kept beside it.
    cand = ((row_ptr[lo + 1:hi + 1] - row_ptr[lo:hi]) >= FANOUTS[1]).nonzero().flatten()
    n_cand = int(cand.numel())
    pick = torch.randperm(n_cand, device=dev, generator=torch.Generator(device=dev).manual_seed(7 + rank))[:per_rank]
    seeds = lo + cand[pick]
    if seeds.numel() < per_rank:                              # not enough candidates (tiny scaled-down runs): top up
        seeds = torch.cat([seeds, seeds_any[:per_rank - seeds.numel()]])
    del cand, pick
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    out = {"graph": "R-MAT (0.57,0.19,0.19,0.05), N=%d nnz=%d max in-degree %d, topology replicated" % (N, NNZ, max_deg),
           "features": "F=%d fp32, node-range partitioned: %d rows (%.1f GB) per GPU" % (F, hi - lo, (hi - lo) * F * 4 / 1e9),
           "model": "GraphSAGE-mean 2-layer %d-256-%d, fanout 25/10, batch 1024 per GPU, Adam, tcgen05 TF32 transforms (fp32 operands read in place by TMA)" % (F, C),
           "train_seeds": n_train, "scaling": "weak per step (1,024 seeds per GPU); the epoch is the fixed 1,207,179 seeds",
           "seed_choice": "uniform among the nodes with in-degree >= %d (%d candidates on rank 0's range); "
                          "'peer_fp32_any_node_seeds' draws from all nodes instead" % (FANOUTS[1], n_cand),
           "setup_s": round(setup_s, 1), "mechanisms": {}}

    def run_peer(tbl, tag, row_ptr=row_ptr, col=col, seeds=seeds):
        sharded = P.PeerShardedTable(N, tbl)
        torch.manual_seed(args.seed)
        model = dnn.GraphSAGE(F, HIDDEN, C, 2, torch.relu, 0.0).to(dev)
        opt = torch.optim.Adam(model.parameters(), lr=0.003, fused=True, capturable=True)
        tr = PL.PipelinedSageTrainer(model, opt, labels, row_ptr, col, F, sharded=sharded, batch_size=BATCH,
                                     fanouts=FANOUTS, precision="tf32", rng_seed=11, label_offset=lo, max_seeds=per_rank)
        tr.set_seeds(seeds)
        tr.capture()
        tr.epoch(seeds[:16 * BATCH])
        barrier()
        r = tr.epoch(seeds)                                  # the whole epoch, measured (not extrapolated)
        barrier()
        tr.set_seeds(seeds[:BATCH])
        tr._prologue.replay()                                # a FULL mini-batch in slot 0 for the traffic statistics
        halo = tr.halo_stats()
        stage = tr.stage_times(seeds[:24 * BATCH], steps=24)
        t = torch.tensor([r["time_s"], stage["produce_ms"], stage["train_ms"], halo["remote_edge_fraction"],
                          float(halo["remote_bytes_per_step"]), float(halo["block0_edges"]), float(halo["block0_dst_rows"])],
                         device=dev, dtype=torch.float64)
        mx = t.clone()
        if world > 1:
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t /= world
        n_b = r["n_batches"]
        res = {"halo": "peer: remote rows read over NVLink by the input-layer aggregation kernel itself "
                       "(dgllb_spmm_csr_sharded), no collective, no staging",
               "step": "one CUDA graph per mini-batch: {sample + build block + sharded aggregation of batch i+1} || "
                       "{train batch i + flat all-reduce + Adam}; overlap on at every N",
               "epoch_s": round(mx[0].item(), 4), "batches_per_gpu": n_b,
               "ms_per_step": round(mx[0].item() * 1e3 / n_b, 4),
               "seeds_per_s": round(world * per_rank / mx[0].item(), 1),
               "stage_ms_alone": {"sample+halo+aggregate (branch B)": round(mx[1].item(), 4),
                                         "fwd+bwd+all-reduce+Adam (branch A)": round(mx[2].item(), 4)},
               "remote_edge_fraction_measured": round(t[3].item(), 4),
               "nvlink_bytes_in_per_gpu_per_step": int(t[4].item()),
               "block0_edges_per_step": int(t[5].item()), "block0_dst_rows_per_step": int(t[6].item()),
               "loss": round(r["loss"], 4)}
        tr.close()
        for p in model.parameters():
            p.grad = None
        barrier()                                            # no peer may still be reading this rank's shard
        sharded.close()
        out["mechanisms"][tag] = res

    def run_nccl():
        hx = P.HaloExchange(N, table)
        torch.manual_seed(args.seed)
        model = dnn.GraphSAGE(F, HIDDEN, C, 2, torch.relu, 0.0).to(dev)
        opt = torch.optim.Adam(model.parameters(), lr=0.003, fused=True, capturable=True)
        tr = T.GraphedSageTrainer(model, opt, None, labels, BATCH, FANOUTS, n_feat=F, precision="tf32",
                                  capture_collectives=True, label_offset=lo)
        steps = 40

        def produce(first, count):
            for k in range(count):
                s = seeds[(first + k) * BATCH:(first + k + 1) * BATCH]
                blocks = G.sample_blocks(row_ptr, col, s, FANOUTS, rng_seed=rank * 100003 + first + k)
                yield s, blocks, hx.fetch(blocks[0].src_ids)

        first = next(produce(0, 1))
        tr.load(*first)
        tr.capture()
        tr.epoch(produce(0, 5), overlap=False)
        barrier()
        hx.stats = {"rows": 0, "remote_rows": 0, "calls": 0}
        r = tr.epoch(produce(5, steps), overlap=False)       # exchange and step on ONE stream: same communicator, ordered
        t = torch.tensor([r["time_s"]], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t[0].item() * 1e3 / steps
        out["mechanisms"]["nccl_all_to_all"] = {
            "halo": "NCCL all_to_all_single: counts, ids, rows (parallel.HaloExchange, one host read-back per step), then "
                    "the step graph with the flat all-reduce and Adam captured; exchange and step on one stream",
            "steps": steps, "ms_per_step": round(ms, 4), "seeds_per_s": round(world * BATCH / (ms * 1e-3), 1),
            "epoch_s_extrapolated": round(per_rank / BATCH * ms * 1e-3, 4),
            "remote_row_fraction_measured": round(hx.stats["remote_rows"] / max(hx.stats["rows"], 1), 4),
            "unique_src_rows_per_step": round(hx.stats["rows"] / max(hx.stats["calls"], 1), 1), "loss": round(r["loss"], 4)}
        for p in model.parameters():
            p.grad = None

    def run_uniform_control():
        """Same N, nnz, features and trainer on the uniform-random control graph (SURVEY.md §8 d): every node has 14-15
        in-neighbours, so blocks are full-size and nearly every sampled source row is distinct — the heaviest halo load
        this model can put on NVLink (the R-MAT graph's random seeds mostly have tiny in-degree)."""
        g = torch.Generator(device=dev).manual_seed(args.seed)
        deg = torch.full((N,), NNZ // N, dtype=torch.int64, device=dev)
        deg[: NNZ - (NNZ // N) * N] += 1
        u_rp = torch.zeros(N + 1, dtype=torch.int64, device=dev)
        torch.cumsum(deg, 0, out=u_rp[1:])
        del deg
        u_col = torch.empty(NNZ, dtype=torch.int32, device=dev)
        for o in range(0, NNZ, 1 << 28):
            m = min(1 << 28, NNZ - o)
            u_col[o:o + m] = torch.randint(0, N, (m,), device=dev, generator=g, dtype=torch.int32)
        run_peer(table, "peer_fp32_uniform_control_graph", u_rp, u_col)

    def run_cpu_baseline():
        """the reference CPU aggregation on K = 20 of this run's own mini-batches (device-sampled,
        copied to the host), source rows taken from a host table slice; epoch figure extrapolated."""
        import numpy as np
        import oracle
        torch.set_num_threads(os.cpu_count() or 1)
        Kb, rows_host = 20, 2_000_000
        torch.manual_seed(args.seed)
        m = dnn.GraphSAGE(F, HIDDEN, C, 2, torch.relu, 0.0).to(dev)
        o = torch.optim.Adam(m.parameters(), lr=0.003, fused=True, capturable=True)
        sh = P.PeerShardedTable(N, table) if world == 1 else None
        if sh is None:
            return                                           # CPU arm only at N = 1 (the contract's rank-0 rule)
        tr = PL.PipelinedSageTrainer(m, o, labels, row_ptr, col, F, sharded=sh, batch_size=BATCH, fanouts=FANOUTS,
                                     precision="tf32", rng_seed=11, label_offset=lo, max_seeds=Kb * BATCH)
        tr.set_seeds(seeds[:Kb * BATCH])
        blocks = []
        for _ in range(Kb):
            tr._produce(tr.slots[0])
            sl = tr.slots[0]
            n0, e0, e1 = int(sl.cnt1[0]), int(sl.rp0[-1]), int(sl.rp1[-1])
            blocks.append((sl.rp0[:n0 + 1].cpu().numpy().astype(np.int64), (sl.nbr0[:e0].cpu().numpy() % rows_host).astype(np.int32),
                           sl.rp1.cpu().numpy().astype(np.int64), sl.col1[:e1].cpu().numpy().astype(np.int32)))
        for p_ in m.parameters():
            p_.grad = None
        x = np.random.default_rng(args.seed).standard_normal((rows_host, F), dtype=np.float32)
        xt = torch.from_numpy(x)
        h = np.random.default_rng(1).standard_normal((BATCH * (1 + FANOUTS[1]), HIDDEN), dtype=np.float32)

        def step_csr(b):
            rp0, c0, rp1, c1 = b
            oracle.spmm_csr(rp0, c0, x, reduce="mean")
            oracle.spmm_csr(rp1, c1, h, reduce="mean")

        def step_coo(b):
            rp0, c0, rp1, c1 = b
            for rp_, c_, src in ((rp0, c0, xt), (rp1, c1, torch.from_numpy(h))):
                deg = np.diff(rp_)
                rows_ = torch.from_numpy(np.repeat(np.arange(rp_.size - 1), deg))
                a = torch.sparse_coo_tensor(torch.stack([rows_, torch.from_numpy(c_).long()]),
                                            torch.from_numpy(np.repeat(1.0 / np.maximum(deg, 1).astype(np.float32), deg)),
                                            (rp_.size - 1, src.size(0)))
                torch.sparse.mm(a, src)

        res = {}
        for name, fn in (("coo", step_coo), ("csr", step_csr)):
            fn(blocks[0])
            t0 = time.perf_counter()
            for b in blocks:
                fn(b)
            res[name] = (time.perf_counter() - t0) / Kb
        best = min(res, key=res.get)
        out["cpu_baseline_extrapolated"] = {
            "kind": "port", "cores": os.cpu_count(), "sample": "K=%d of this run's mini-batches (device-sampled, copied to the "
            "host), aggregation of both blocks only (no transforms, no backward: a LOWER bound on the CPU step), source rows "
            "from a %d-row host table slice; torch.sparse.mm COO as gcnconv.py:31 and the OpenMP CSR port" % (Kb, rows_host),
            "ms_per_step": {k: round(v * 1e3, 3) for k, v in res.items()}, "best": best,
            "seeds_per_s": round(BATCH / res[best], 1), "epoch_s_extrapolated": round(n_train / BATCH * res[best], 2)}

    for tag, fn in (("peer_fp32", lambda: run_peer(table, "peer_fp32")), ("nccl_all_to_all", run_nccl),
                    ("peer_bf16_table", lambda: run_peer(table.to(torch.bfloat16), "peer_bf16_table")),
                    ("peer_fp32_any_node_seeds", lambda: run_peer(table, "peer_fp32_any_node_seeds", seeds=seeds_any)),
                    ("peer_fp32_uniform_control_graph", run_uniform_control)):
        try:
            fn()
        except Exception as ex:
            out["mechanisms"][tag] = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:300])}
        barrier()
    if world == 1 and not args.no_cpu_baseline:
        try:
            run_cpu_baseline()
        except Exception as ex:
            out["cpu_baseline_extrapolated"] = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:300])}
    return out


# --------------------------------------------------------------------- GPU arm --
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-epoch", action="store_true", help="skip the sampled-GraphSAGE training-epoch extra")
    ap.add_argument("--no-partitioned", action="store_true", help="skip the papers100M-shaped partitioned extra")
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from dgll_b200 import _lib, kernels as K

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic inputs, resident in HBM -------------------------------------------------------------
    t_setup = time.perf_counter()
    row_ptr, col_idx, table, gen, batches = make_inputs(args.seed, rank, dev)
    for bt in batches:
        bt["agg0"] = torch.empty((bt["n_dst0"], LD32), device=dev)[:, :FEAT]      # 16-byte rows: 128-bit stores
        bt["self0"] = torch.empty((bt["n_dst0"], LD32), device=dev)
        bt["agg1"] = torch.empty((bt["n_dst1"], HIDDEN), device=dev)
        bt["h1"] = torch.randn((bt["n_dst0"], HIDDEN), device=dev, generator=gen)
        bt["bytes0"] = alg_bytes_spmm(bt["col0"].numel(), bt["n_dst0"], FEAT, rp=4)
        bt["bytes_rest"] = (alg_bytes_gather(bt["n_dst0"], LD32 * 4)
                            + alg_bytes_spmm(bt["col1"].numel(), bt["n_dst1"], HIDDEN, rp=4))
        bt["bytes"] = bt["bytes0"] + bt["bytes_rest"]
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup

    def make_step(tbl):
        view = tbl[:, :FEAT]

        def step(bt, ev=None):
            if ev is not None:
                ev[0].record()
            K.spmm_csr(bt["rp0"], bt["col0"], view, reduce="mean", out=bt["agg0"])
            if ev is not None:
                ev[1].record()
            K.gather_rows(table, bt["dst0"], out=bt["self0"])
            K.spmm_csr(bt["rp1"], bt["col1"], bt["h1"], reduce="mean", out=bt["agg1"])
        return step

    step = make_step(table)
    view = table[:, :FEAT]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(step_fn, steps, warmup, profile=False):
        """W warm-up steps, then exactly ``steps`` steps between two events; the dominant kernel is bracketed by its own
        event pair on every 4th step (events cost ~1-2 us of stream time each)."""
        for w in range(warmup):
            step_fn(batches[w % N_BATCHES])
        barrier()
        KEV = 4
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range((steps + KEV - 1) // KEV)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if profile:
            torch.cuda.profiler.start()
        e0.record()
        for s in range(steps):
            step_fn(batches[s % N_BATCHES], kev[s // KEV] if s % KEV == 0 else None)
        e1.record()
        barrier()
        if profile:
            torch.cuda.profiler.stop()
        return e0.elapsed_time(e1), sum(a.elapsed_time(b) for a, b in kev) / len(kev)

    # ---- device-resident timing ------------------------------------------------------------------------
    clocks = ClockSampler(local_rank)
    for w in range(args.warmup):
        step(batches[w % N_BATCHES])
    barrier()
    if rank == 0:
        clocks.start()
    l0 = _lib.launch_count()
    prof = os.environ.get("BENCH_PROFILE") == "1"   # ncu --profile-from-start off: capture the timed steps only
    ms, k_ms = timed_loop(step, args.steps, 0, profile=prof)
    launches = _lib.launch_count() - l0
    total_bytes = sum(batches[s % N_BATCHES]["bytes"] for s in range(args.steps))
    # supporting figure: the dominant kernel alone, the distinct mini-batches back to back between ONE event pair — an
    # event pair around a single ~85 us launch also counts the launch gap on both sides (ncu on the same launch: 80 us)
    bb0, bb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bb_reps = 4
    bb0.record()
    for _ in range(bb_reps):
        for bt in batches:
            K.spmm_csr(bt["rp0"], bt["col0"], view, reduce="mean", out=bt["agg0"])
    bb1.record()
    torch.cuda.synchronize()
    k_ms_bb = bb0.elapsed_time(bb1) / (bb_reps * N_BATCHES)
    k_bytes_bb = sum(bt["bytes0"] for bt in batches) / N_BATCHES
    k_bytes = sum(batches[s % N_BATCHES]["bytes0"] for s in range(0, args.steps, 4)) / len(range(0, args.steps, 4))

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
        # (a high-priority branch for the two short kernels made the step slower: 0.107 -> 0.118 ms)
        side, side2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        graphs = []
        torch.cuda.synchronize()
        for i in range(N_BATCHES):
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph):
                cur = torch.cuda.current_stream()
                side.wait_stream(cur)
                side2.wait_stream(cur)
                nb = (i + 1) % N_BATCHES
                pb = (i - 1) % N_BATCHES
                with torch.cuda.stream(side):
                    dev_buf[nb % 2][:host[nb][0].numel()].copy_(host[nb][0], non_blocking=True)
                    out_host[pb % 2][:batches[pb]["n_dst1"]].copy_(batches[pb]["agg1"], non_blocking=True)
                bt, offs = batches[i], host[i][1]
                v = [dev_buf[i % 2][o:o + m] for (o, m) in offs]
                # the three kernels of a mini-batch do not depend on each other (layer 1 aggregates the given hidden
                # table): the dst-row gather and the layer-1 aggregation run on a branch beside the layer-0 aggregation
                with torch.cuda.stream(side2):
                    K.gather_rows(table, v[2], out=bt["self0"])
                    K.spmm_csr(v[3], v[4], bt["h1"], reduce="mean", out=bt["agg1"])
                K.spmm_csr(v[0], v[1], view, reduce="mean", out=bt["agg0"])
                cur.wait_stream(side2)
                cur.wait_stream(side)
            graphs.append(gph)
        e2e_mode = ("CUDA graph per mini-batch ({layer-0 aggregation} || {dst-row gather, layer-1 aggregation} || "
                    "{D2H of the previous result, H2D of the next batch})")
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
    # the same step as a BLOCKING call (what a caller without a pipeline sees): H2D of this batch's blocks, the three kernels,
    # D2H of the result, host synchronisation — every step, nothing overlapped
    blk_steps = min(args.steps, 200)
    torch.cuda.synchronize()
    t_blk = time.perf_counter()
    for s_ in range(blk_steps):
        i_ = s_ % N_BATCHES
        hbuf, offs = host[i_]
        bt = batches[i_]
        dev_buf[0][:hbuf.numel()].copy_(hbuf, non_blocking=True)
        v = [dev_buf[0][o:o + m] for (o, m) in offs]
        K.spmm_csr(v[0], v[1], view, reduce="mean", out=bt["agg0"])
        K.gather_rows(table, v[2], out=bt["self0"])
        K.spmm_csr(v[3], v[4], bt["h1"], reduce="mean", out=bt["agg1"])
        out_host[0][:bt["n_dst1"]].copy_(bt["agg1"], non_blocking=True)
        torch.cuda.synchronize()
    blk_ms = (time.perf_counter() - t_blk) * 1e3 / blk_steps
    blk_bytes = sum(batches[s_ % N_BATCHES]["bytes"] for s_ in range(blk_steps)) / blk_steps
    clk = clocks.stop() if rank == 0 else None
    del graphs

    # ---- same-box GPU comparator: the library SpMM behind the reference's layers -------------------------------
    gpu_baseline = None
    if rank == 0:
        try:
            mats = []
            for bt in batches:
                deg = (bt["rp0"][1:] - bt["rp0"][:-1]).to(torch.float32).clamp(min=1)
                vals = torch.repeat_interleave(1.0 / deg, (bt["rp0"][1:] - bt["rp0"][:-1]).long())
                mats.append(torch.sparse_csr_tensor(bt["rp0"].long(), bt["col0"].long(), vals, size=(bt["n_dst0"], N_NODES)))
            xc = view.contiguous()                                   # torch.sparse.mm needs a dense contiguous operand
            for m in mats[:3]:
                torch.sparse.mm(m, xc)
            torch.cuda.synchronize()
            evs = []
            for s in range(min(args.steps, 64)):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                torch.sparse.mm(mats[s % N_BATCHES], xc)
                b.record()
                evs.append((a, b))
            torch.cuda.synchronize()
            t_lib = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
            chk = (torch.sparse.mm(mats[0], xc) - batches[0]["agg0"]).abs().max().item()
            gpu_baseline = {"kernel": "torch.sparse.mm(CSR on CUDA = cuSPARSE, the call behind gcnconv.py:31 / gcn_model.py:76), "
                                      "layer-0 aggregation of the same mini-batches", "kernel_ms": t_lib,
                            "value": k_bytes / (t_lib * 1e-3) / 1e9, "unit": "GB/s", "ours_kernel_ms": k_ms,
                            "speedup": t_lib / k_ms, "max_abs_diff_vs_ours": chk}
            del mats, xc
        except Exception as ex:
            gpu_baseline = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:200])}

    # ---- extra: the same step with the feature table stored in bf16 ----------------------------------------------
    bf16 = None
    try:
        t16 = torch.zeros((N_NODES, LD16), dtype=torch.bfloat16, device=dev)
        t16[:, :FEAT] = view.to(torch.bfloat16)
        step16 = make_step(t16)
        ms16, k16 = timed_loop(step16, args.steps, max(args.warmup, 3))
        b16 = sum(alg_bytes_spmm(batches[s % N_BATCHES]["col0"].numel(), batches[s % N_BATCHES]["n_dst0"], FEAT, b=2, rp=4)
                  + batches[s % N_BATCHES]["bytes_rest"] for s in range(args.steps))
        kb16 = sum(alg_bytes_spmm(batches[s % N_BATCHES]["col0"].numel(), batches[s % N_BATCHES]["n_dst0"], FEAT, b=2, rp=4)
                   for s in range(0, args.steps, 4)) / len(range(0, args.steps, 4))
        ref16 = K.spmm_csr(batches[0]["rp0"], batches[0]["col0"], view, reduce="mean")
        got16 = K.spmm_csr(batches[0]["rp0"], batches[0]["col0"], t16[:, :FEAT], reduce="mean")
        bf16 = {"table": "bf16, row stride 608 (283 MB)", "ms_per_step": ms16 / args.steps,
                "value_alg_GBps_at_2_bytes": b16 / (ms16 * 1e-3) / 1e9, "kernel_ms": k16,
                "kernel_alg_GBps": kb16 / (k16 * 1e-3) / 1e9, "speedup_vs_fp32_step": (ms / args.steps) / (ms16 / args.steps),
                "max_rel_err_vs_fp32_table": ((got16 - ref16).abs().max() / ref16.abs().max()).item(),
                "scope": "per GPU (rank 0)"}
        del t16, ref16, got16
    except Exception as ex:
        bf16 = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:200])}

    # ---- extras: training epoch (Reddit-shaped) and the partitioned papers100M-shaped run -------------------------
    epoch = None
    if not args.no_epoch:
        try:
            epoch = extra_epoch(args, rank, world, dev, row_ptr, col_idx, table, gen, barrier)
        except Exception as ex:
            epoch = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:300])}
    n_dst0, nnz0 = batches[0]["n_dst0"], batches[0]["col0"].numel()
    del row_ptr, col_idx, table, view, batches, host, dev_buf, step
    torch.cuda.empty_cache()
    partitioned = None
    if not args.no_partitioned:
        try:
            partitioned = extra_partitioned(args, rank, world, dev, barrier)
        except Exception as ex:
            partitioned = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:300])}
        torch.cuda.empty_cache()

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
    traffic, traffic_src = ncu_traffic()
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(world, n_dst0, nnz0),
        "roofline": {"bound": "hbm", "kernel": "spmm_rows_stream_kernel<float,5,4> (layer-0 mean aggregation, F=602: whole rows per warp, window rolling across rows)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": "%s (ncu --set full)" % traffic_src if traffic_src else None,
                     "dram_frac": (traffic / (k_ms * 1e-3) / 1e9 / peak) if traffic else None,
                     "peak_source": peak_src, "kernel_ms": k_ms, "algorithmic_bytes_per_launch": k_bytes,
                     "back_to_back": {"kernel_ms": k_ms_bb, "achieved": k_bytes_bb / (k_ms_bb * 1e-3) / 1e9,
                                      "dram_frac": (traffic / (k_ms_bb * 1e-3) / 1e9 / peak) if traffic else None,
                                      "how": "%d launches over the %d distinct mini-batches between one event pair, outside "
                                             "the timed steps" % (bb_reps * N_BATCHES, N_BATCHES)},
                     "note": "frac counts every edge's full source row (SURVEY §8 d); dram_frac counts the bytes DRAM "
                             "actually moved — repeated source rows of a mini-batch are served by L2"},
        "e2e": {"value": e2e_bytes / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d / args.steps,
                "d2h_bytes_per_step": d2h / args.steps, "ms_per_step": e2e_ms / args.steps, "mode": e2e_mode,
                "blocking_call": {"value": blk_bytes / (blk_ms * 1e-3) / 1e9, "ms_per_step": blk_ms, "steps": blk_steps,
                                  "how": "per step: H2D of the batch's blocks, three kernels, D2H of the result, host "
                                         "synchronisation; wall clock, nothing overlapped (rank 0)"}},
        "gpu_launches": launches, "clocks": clk, "gpu_baseline": gpu_baseline, "bf16_table": bf16,
        "epoch": epoch, "partitioned": partitioned, "setup_s": round(setup_s, 1),
    }
    if not args.no_cpu_baseline and world == 1:
        r = run_cpu_arm(8, 1, budget_s=24.0, seed=args.seed, n_blocks=4)
        line["cpu_baseline"] = {
            "value": r["value"], "unit": "GB/s", "cores": r["cores"], "kind": "port",
            "sample": "<=8 steps per variant (~12 s each) on 4 of the GPU arm's own mini-batches (%s); best=%s; %s" % (
                r["origin"], r["best"], json.dumps({k: round(x["gbs"], 2) for k, x in r["variants"].items()}))}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
