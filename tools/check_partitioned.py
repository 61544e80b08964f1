#!/usr/bin/env python
"""Multi-GPU correctness of the partitioned one-graph-per-step trainer (run under torchrun, >= 2 GPUs):
  (1) training on the node-range-partitioned table read over NVLink (PeerShardedTable + dgllb_spmm_csr_sharded) takes
      bit-identical steps to training on a replicated copy of the same table;
  (2) every rank ends with the same weights (captured flat all-reduce);
  (3) the peer gather equals the NCCL all_to_all halo exchange bit for bit.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_partitioned.py
Prints one JSON line on rank 0 and exits non-zero on any mismatch."""
import copy
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import dgll_b200.nn as dnn  # noqa: E402
from dgll_b200 import graphs as G, parallel as P, pipelined as PL  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N, F, C, B = 300000, 128, 19, 512
rp, col = G.rmat_csr_large(N, N * 15, seed=0, device=dev)
full = G.feature_table(N, F, seed=5, device=dev)                     # same table on every rank
lo, hi = P.local_range(rank, N, world)
shard = full[lo:hi].clone()
labels = torch.randint(0, C, (N,), device=dev, generator=torch.Generator(device=dev).manual_seed(1))
seeds = lo + torch.randperm(hi - lo, device=dev, generator=torch.Generator(device=dev).manual_seed(7 + rank))[:B * 6 + 77]
torch.manual_seed(0)
m0 = dnn.GraphSAGE(F, 64, C, 2, torch.relu, 0.0).to(dev)
res, models = [], []
sharded = P.PeerShardedTable(N, shard)
for kw in ({"table": full}, {"sharded": sharded}):
    m = copy.deepcopy(m0)
    o = torch.optim.Adam(m.parameters(), lr=0.01, fused=True, capturable=True)
    lab = labels if "table" in kw else labels[lo:hi].contiguous()
    tr = PL.PipelinedSageTrainer(m, o, lab, rp, col, F, batch_size=B, fanouts=(10, 5), precision="bf16", rng_seed=3,
                                 label_offset=0 if "table" in kw else lo, **kw)
    res.append(tr.epoch(seeds))
    if "sharded" in kw:
        halo = tr.halo_stats()
    models.append(m)
ok = res[0]["loss"] == res[1]["loss"]
for p, q in zip(models[0].parameters(), models[1].parameters()):
    ok = ok and torch.equal(p, q)
flat = torch.cat([p.detach().reshape(-1) for p in models[1].parameters()])
ref = flat.clone()
dist.broadcast(ref, 0)
same_across_ranks = torch.equal(flat, ref)
ids = torch.randint(0, N, (50000,), device=dev, generator=torch.Generator(device=dev).manual_seed(rank))
peer_eq_nccl = torch.equal(sharded.fetch(ids), P.HaloExchange(N, shard).fetch(ids)) and torch.equal(sharded.fetch(ids), full[ids])
flags = torch.tensor([int(ok), int(same_across_ranks), int(peer_eq_nccl)], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
dist.barrier()
del tr, o, models, m          # captured graphs hold NCCL kernels: drop them before the process group goes away
import gc  # noqa: E402
gc.collect()
torch.cuda.synchronize()
sharded.close()
if rank == 0:
    print(json.dumps({"world": world, "partitioned_equals_replicated_bitwise": bool(flags[0]),
                      "weights_identical_on_all_ranks": bool(flags[1]), "peer_gather_equals_nccl_exchange": bool(flags[2]),
                      "loss": res[1]["loss"], "remote_edge_fraction_rank0": halo["remote_edge_fraction"]}), flush=True)
dist.destroy_process_group()
sys.exit(0 if bool(flags.min()) else 1)
