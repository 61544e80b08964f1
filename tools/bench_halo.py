#!/usr/bin/env python
"""BASELINE.json configs[4]: sampled GraphSAGE on a papers100M-shaped graph, features node-range partitioned across the
GPUs of one box, halo feature rows exchanged with NCCL all-to-all (dgll_b200.parallel.HaloExchange).

  python tools/bench_halo.py                                   # 1 GPU (everything local)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         tools/bench_halo.py [--scale 0.125] [--steps 60]

Every rank: replicated topology (CSR by destination), its shard of the fp32 feature table [ceil(N/P), 128], its own
seeds (local by destination).  One step = sample 25/10 blocks on the device -> bucket block0's source ids by owner ->
all_to_all ids -> owners gather rows (TMA gather kernel) -> all_to_all rows -> 2-layer SAGE forward + backward on the
aggregation kernels -> flat gradient all-reduce -> Adam.  Prints ONE JSON line on rank 0: whole-job seeds/s, ms per
step (max over ranks, CUDA events), the stage breakdown and the remote-row fraction.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import dgll_b200.nn as dnn  # noqa: E402
from dgll_b200 import graphs as G, ops, parallel as P, train as T  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="shrink N and nnz by this factor (1.0 = papers100M-shaped)")
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--graph", default="uniform", choices=["uniform", "rmat"])
    ap.add_argument("--mode", default="graph", choices=["graph", "eager"],
                    help="graph = the training step replayed as one CUDA graph on fixed-capacity buffers "
                         "(train.GraphedSageTrainer); eager = Python-dispatched step")
    ap.add_argument("--capture-collectives", action="store_true",
                    help="capture the flat gradient all-reduce and Adam inside the step graph (N > 1)")
    ap.add_argument("--no-overlap", dest="overlap", action="store_false",
                    help="graph mode: produce step i+1 (sample, build blocks, fetch rows) on the main stream instead of a "
                         "side stream that runs under the replay of step i")
    ap.add_argument("--halo", default="peer", choices=["padded", "exact", "peer"],
                    help="peer = shards mapped over NVLink (CUDA IPC), ONE gather kernel pulls remote rows; exact = NCCL "
                         "all_to_all with counts exchanged first; padded = NCCL with fixed-capacity buckets, no host sync")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N0, NNZ0, F, C = G.SHAPES["papers100m"]
    N = int(N0 * args.scale)
    NNZ = int(NNZ0 * args.scale)
    fanouts, hidden = (25, 10), 256

    t0 = time.perf_counter()
    if args.graph == "rmat":
        row_ptr, col = G.rmat_csr(N, NNZ, seed=0, device=dev)
    else:
        # degree-14/15 control graph (papers100M's average in-degree is 14.5); same on every rank
        gen = torch.Generator(device=dev).manual_seed(0)
        deg = torch.full((N,), NNZ // N, dtype=torch.int64, device=dev)
        deg[: NNZ - (NNZ // N) * N] += 1
        row_ptr = torch.zeros(N + 1, dtype=torch.int64, device=dev)
        torch.cumsum(deg, 0, out=row_ptr[1:])
        del deg
        col = torch.empty(NNZ, dtype=torch.int32, device=dev)
        step = 1 << 28
        for o in range(0, NNZ, step):
            m = min(step, NNZ - o)
            col[o:o + m] = torch.randint(0, N, (m,), device=dev, generator=gen, dtype=torch.int32)
    lo, hi = P.local_range(rank, N, world)
    table = G.feature_table(hi - lo, F, seed=100 + rank, device=dev)          # this rank's feature rows
    labels = torch.randint(0, C, (N,), device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    n_train = int(N * 1207179 / N0) if args.scale == 1.0 else int(N * 0.0109)
    gs = torch.Generator(device=dev).manual_seed(7 + rank)
    seeds_all = lo + torch.randperm(hi - lo, device=dev, generator=gs)[: max(n_train // world, args.batch * (args.steps + args.warmup))]
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0

    hx = P.PeerShardedTable(N, table) if args.halo == "peer" else P.HaloExchange(N, table)
    torch.manual_seed(0)
    model = dnn.GraphSAGE(F, hidden, C, 2, torch.relu, 0.0).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=0.003, fused=True,
                           capturable=((world == 1 or args.capture_collectives) and args.mode == "graph"))
    params = list(model.parameters())
    ops.set_gemm_precision(args.precision)
    fetch = (lambda ids: hx.fetch_padded(ids)) if args.halo == "padded" else (lambda ids: hx.fetch(ids))
    trainer, mode, mode_err = None, args.mode, None
    if args.mode == "graph":
        try:
            trainer = T.GraphedSageTrainer(model, opt, None, labels, args.batch, fanouts, n_feat=F,
                                           capture_collectives=args.capture_collectives)
            b0 = G.sample_blocks(row_ptr, col, seeds_all[:args.batch], fanouts, rng_seed=999 + rank)
            trainer.load(seeds_all[:args.batch], b0, fetch(b0[0].src_ids))
            trainer.capture()
        except Exception as ex:   # keep the run: fall back to the eager step and say so in the JSON line
            trainer, mode, mode_err = None, "eager", "%s: %s" % (type(ex).__name__, str(ex)[:160])
            opt = torch.optim.Adam(model.parameters(), lr=0.003, fused=True)

    def one_step(i, ev=None):
        s = seeds_all[i * args.batch:(i + 1) * args.batch]
        if ev: ev[0].record()
        blocks = G.sample_blocks(row_ptr, col, s, fanouts, rng_seed=rank * 100003 + i)
        if ev: ev[1].record()
        x = fetch(blocks[0].src_ids)                                           # peer: ONE kernel
        if ev: ev[2].record()
        if trainer is not None:
            trainer.load(s, blocks, x)
            trainer.step()                      # graph replay (+ flat gradient all-reduce and Adam when world > 1)
        else:
            logits = model(blocks, x)
            loss = torch.nn.functional.cross_entropy(logits, labels[s])
            opt.zero_grad(set_to_none=True)
            loss.backward()
            P.allreduce_gradients(params)
            opt.step()
        if ev: ev[3].record()
        return blocks[0].num_src, blocks[0].num_edges()

    def produce(first, count, evs=None):
        """(seeds, blocks, x) per step — sampling, block building and the halo fetch; runs on the trainer's side stream
        when the production of step i+1 overlaps the replay of step i (events are recorded on whatever stream is current)."""
        for k in range(count):
            i = first + k
            s = seeds_all[i * args.batch:(i + 1) * args.batch]
            if evs: evs[k][0].record()
            blocks = G.sample_blocks(row_ptr, col, s, fanouts, rng_seed=rank * 100003 + i)
            if evs: evs[k][1].record()
            x = fetch(blocks[0].src_ids)
            if evs: evs[k][2].record()
            counts[0] += blocks[0].num_src
            counts[1] += blocks[0].num_edges()
            yield s, blocks, x

    counts = [0, 0]
    if trainer is not None:
        trainer.epoch(produce(0, args.warmup), overlap=args.overlap)
    else:
        for w in range(args.warmup):
            one_step(w)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    hx.stats = {"rows": 0, "remote_rows": 0, "calls": 0}
    verified = None
    if args.halo == "peer" and world > 1:
        # bit-exact cross-check of the NVLink peer gather against the NCCL all-to-all exchange on one real batch
        blocks = G.sample_blocks(row_ptr, col, seeds_all[:args.batch], fanouts, rng_seed=12345 + rank)
        a = hx.fetch(blocks[0].src_ids)
        b = P.HaloExchange(N, table).fetch(blocks[0].src_ids)
        ok = torch.tensor([int(torch.equal(a, b))], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        verified = bool(ok.item())
        assert verified, "peer gather and NCCL halo exchange disagree"
        hx.stats = {"rows": 0, "remote_rows": 0, "calls": 0}
    hx_overflow = (lambda: hx.check_overflow()) if hasattr(hx, "check_overflow") else (lambda: False)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    counts = [0, 0]
    if trainer is not None:
        def cb(stage, i):
            evs[i][3 if stage == "before" else 4].record()
        r = trainer.epoch(produce(args.warmup, args.steps, evs), overlap=args.overlap, callback=cb)
        ms = r["time_s"] * 1e3
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        stage = [sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps, sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps,
                 sum(e[3].elapsed_time(e[4]) for e in evs) / args.steps]
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            a, b = one_step(args.warmup + i, evs[i])
            counts[0] += a
            counts[1] += b
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        stage = [sum(e[k].elapsed_time(e[k + 1]) for e in evs) / args.steps for k in range(3)]
    n_src, n_edge = counts
    t = torch.tensor([ms] + stage, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = t[0].item()
        seeds_per_s = world * args.batch * args.steps / (ms * 1e-3)
        rows_per_step = hx.stats["rows"] / args.steps
        remote = (hx.stats["remote_rows"] / max(hx.stats["rows"], 1)) if args.halo != "peer" else (world - 1) / world
        print(json.dumps({
            "metric": "sampled GraphSAGE training throughput, papers100M-shaped, node-range partitioned features",
            "value": seeds_per_s, "unit": "seeds/s", "n_gpus": world, "steps": args.steps, "ms_per_step": ms / args.steps,
            "epoch_s_extrapolated": (1207179 * args.scale) / seeds_per_s,
            "stage_ms_max_over_ranks": {"sample": t[1].item(), "halo_exchange": t[2].item(), "fwd_bwd_step": t[3].item()},
            "block0_src_rows_per_step": rows_per_step, "block0_edges_per_step": n_edge / args.steps,
            "remote_row_fraction": remote,
            "halo_bytes_in_per_gpu_per_step": rows_per_step * remote * F * 4,
            "config": {"N": N, "nnz": NNZ, "F": F, "fanouts": list(fanouts), "batch_per_gpu": args.batch, "hidden": hidden,
                       "graph": args.graph, "gemm": args.precision, "step": mode, "step_fallback_reason": mode_err,
                       "overlap_production_with_replay": bool(args.overlap and trainer is not None), "halo": args.halo, "halo_overflow": hx_overflow(), "peer_vs_nccl_bit_exact": verified, "scale": args.scale, "setup_s": round(setup_s, 1)},
            "scaling": "weak"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
