#!/usr/bin/env python
"""A/B of the pipelined trainer's overlap knobs on the papers100M-shaped uniform control graph (the heaviest halo
load): ms per step with the training branch captured on a high-priority stream vs default priorities, and with the
sharded aggregation's residency capped (option rows_sharded_bps).
  python tools/ab_train_priority.py            (1 GPU)
  torchrun --nproc-per-node N tools/ab_train_priority.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import dgll_b200.nn as dnn  # noqa: E402
from dgll_b200 import graphs as G, parallel as P, pipelined as PL  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N, NNZ, F, C = G.SHAPES["papers100m"]
BATCH, FANOUTS, HIDDEN = 1024, (25, 10), 256
g = torch.Generator(device=dev).manual_seed(0)
deg = torch.full((N,), NNZ // N, dtype=torch.int64, device=dev)
deg[: NNZ - (NNZ // N) * N] += 1
rp = torch.zeros(N + 1, dtype=torch.int64, device=dev)
torch.cumsum(deg, 0, out=rp[1:])
del deg
col = torch.empty(NNZ, dtype=torch.int32, device=dev)
for o in range(0, NNZ, 1 << 28):
    m = min(1 << 28, NNZ - o)
    col[o:o + m] = torch.randint(0, N, (m,), device=dev, generator=g, dtype=torch.int32)
lo, hi = P.local_range(rank, N, world)
table = G.feature_table(hi - lo, F, seed=100 + rank, device=dev)
labels = torch.randint(0, C, (hi - lo,), device=dev, generator=torch.Generator(device=dev).manual_seed(1 + rank))
per_rank = 128 * BATCH
seeds = lo + torch.randperm(hi - lo, device=dev, generator=torch.Generator(device=dev).manual_seed(7 + rank))[:per_rank]
out = {"world": world}
def log(msg):
    if rank == 0:
        print("[ab] " + msg, file=sys.stderr, flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "ab_priority_progress.log"), "a") as f:
            f.write(msg + "\n")


log("setup done")
VARIANTS = [("prio_base", True, None), ("prio_bps8", True, "8"), ("noprio_base", False, None), ("noprio_bps8", False, "8"),
            ("noprio_bps5", False, "5"), ("noprio_bps12", False, "12"), ("prio_bps5", True, "5")]
for tag, prio, bps in VARIANTS:
    log("variant " + tag)
    sharded = P.PeerShardedTable(N, table)
    torch.manual_seed(0)
    model = dnn.GraphSAGE(F, HIDDEN, C, 2, torch.relu, 0.0).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=0.003, fused=True, capturable=True)
    tr = PL.PipelinedSageTrainer(model, opt, labels, rp, col, F, sharded=sharded, batch_size=BATCH, fanouts=FANOUTS,
                                 precision="tf32", rng_seed=11, label_offset=lo, max_seeds=per_rank, train_priority=prio,
                                 sharded_blocks_per_sm=int(bps) if bps else 0)
    tr.set_seeds(seeds)
    tr.capture()
    log("captured")
    tr.epoch(seeds[:16 * BATCH])
    log("warm epoch done")
    if world > 1:
        dist.barrier()
    r = tr.epoch(seeds)
    t = torch.tensor([r["time_s"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    log("epoch done")
    out[tag] = {"ms_per_step": round(t.item() * 1e3 / r["n_batches"], 4), "loss": round(r["loss"], 4)}
    log(json.dumps(out))
    tr.close()
    for p in model.parameters():
        p.grad = None
    if world > 1:
        dist.barrier()
    sharded.close()
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
