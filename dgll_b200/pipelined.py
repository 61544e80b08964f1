"""Sampled-GraphSAGE training with ONE CUDA-graph launch per mini-batch, on a resident or a node-range-partitioned
feature table (BASELINE.json configs[1] and configs[4]; SURVEY.md §8 e).

What the reference does per mini-batch (GPU Accelerator/MQGCN.py:117-157, CommGNN_train.py:102-145, buffer_queues.py):
host-side DGL sampling -> feature fetch -> forward/backward on DGL kernels -> one NCCL all_reduce PER PARAMETER
(MQGCN.py:55-67) -> optimizer, with Python threads and bounded queues overlapping the stages.  Here every stage is a
device kernel and the whole step is one graph with two branches:

    branch A (train, slot s)                          branch B (produce, slot 1-s)
      layer 0: self0 W_s + agg0 W_n + b, relu           next seeds <- seed table[counter]           (no host copy)
      layer 1: SAGEConv on block1 (aggregation kernels) sample output layer, fanout f1              dgllb_sample_neighbors_cap
      loss, backward (A^T G on the transposed block)    dst-first compaction                        dgllb_build_block_cap
      ONE flat gradient all-reduce (NCCL, captured)     sample input layer, fanout f0
      Adam (captured)                                   agg0 = mean over the sampled in-neighbours, read STRAIGHT from the
                                                        feature table — over NVLink from the owning GPU when the table
                                                        is partitioned (dgllb_spmm_csr_sharded: the halo exchange IS the
                                                        aggregation's own load) — and self0 = the dst rows

The input-layer aggregation depends on no weight, which is what lets mini-batch i+1 be produced while step i trains
without staleness; it is the MQ-GNN pipeline (README.md:26-37) with CUDA-graph branches instead of threads and
queues.  Nothing is read back to the host inside an epoch; the host's work per step is one graph launch.

Fixed capacities (CUDA graphs need static shapes): B seeds, B*(1+f1) input-layer destination rows, B*f1 and
B*(1+f1)*f0 edges; padding slots carry negative ids (degree-0 rows, zero gradient), a short last batch is masked out
of the loss.
"""
import time

import torch
import torch.distributed as dist

from . import graphs as G
from . import kernels as K
from . import ops


class _Slot:
    """Device buffers of one mini-batch in flight."""

    def __init__(self, B, f0, f1, n_feat, ld_self, self_dtype, dev):
        i32, i64 = torch.int32, torch.int64
        self.cap_d0 = B * (1 + f1)
        self.cap_e1, self.cap_e0 = B * f1, self.cap_d0 * f0
        self.seeds = torch.full((B,), -1, dtype=i64, device=dev)
        self.rp1 = torch.zeros(B + 1, dtype=i32, device=dev)
        self.nbr1 = torch.zeros(self.cap_e1, dtype=i32, device=dev)
        self.src1 = torch.full((self.cap_d0,), -1, dtype=i64, device=dev)
        self.col1 = torch.full((self.cap_e1,), self.cap_d0, dtype=i32, device=dev)
        self.cnt1 = torch.zeros(3, dtype=i32, device=dev)
        self.rp0 = torch.zeros(self.cap_d0 + 1, dtype=i32, device=dev)
        self.nbr0 = torch.zeros(self.cap_e0, dtype=i32, device=dev)
        # rows are 16-byte multiples so that TMA (the TF32 transform) and 128-bit stores can address them in place
        self.agg0 = torch.zeros((self.cap_d0, (n_feat + 3) // 4 * 4), dtype=torch.float32, device=dev)[:, :n_feat]
        self.self0 = torch.zeros((self.cap_d0, ld_self), dtype=self_dtype, device=dev)
        self._shape = torch.empty(self.cap_d0, dtype=torch.int8, device=dev)   # only its length is used (num_src of block1)
        # transpose of block1 (for grad_x = A^T g): structure only, so it is built on the produce branch as well
        self.t_rp1 = torch.zeros(self.cap_d0 + 1, dtype=i32, device=dev)
        self.t_col1 = torch.zeros(self.cap_e1, dtype=i32, device=dev)
        self.t_perm1 = torch.zeros(self.cap_e1, dtype=i32, device=dev)
        self.target = torch.full((B,), -100, dtype=i64, device=dev)            # labels of the seeds, -100 = padding slot
        self.loss_scale = torch.ones((), dtype=torch.float32, device=dev)      # 1 / (valid seeds * world size)


class PipelinedSageTrainer:
    """2-layer ``dgll_b200.nn.GraphSAGE`` trained with one graph launch per mini-batch (see the module docstring).

    Feature source (exactly one):
      ``table``    resident ``[N, ld]`` fp32/bf16 table on this GPU (every rank holds it: data parallel over seeds);
      ``sharded``  a ``parallel.PeerShardedTable`` — the table node-range partitioned over the ranks, peers mapped
                   over NVLink; ``labels`` may then be this rank's slice with ``label_offset`` = its first node id.
    ``row_ptr`` / ``col_idx``: the (replicated) topology, CSR by destination.  ``group``: process group of the
    gradient all-reduce (default group when initialised; world 1 = no collective)."""

    def __init__(self, model, opt, labels, row_ptr, col_idx, n_feat, table=None, sharded=None, batch_size=1024,
                 fanouts=(25, 10), group=None, precision=None, rng_seed=0, label_offset=0, max_seeds=None,
                 train_priority=True, sharded_blocks_per_sm=5):
        if len(fanouts) != 2 or len(model.layers) != 2:
            raise ValueError("PipelinedSageTrainer: 2-layer models / two fanouts")
        if (table is None) == (sharded is None):
            raise ValueError("PipelinedSageTrainer: pass exactly one of table= / sharded=")
        if not opt.defaults.get("capturable", False):
            raise ValueError("PipelinedSageTrainer: the optimizer is captured; build it with capturable=True")
        self.model, self.opt, self.labels, self.group = model, opt, labels, group
        self.row_ptr, self.col_idx = row_ptr, col_idx
        self.table, self.sharded = table, sharded
        self.F = int(n_feat)
        self.B, self.f0, self.f1 = int(batch_size), int(fanouts[0]), int(fanouts[1])
        self.label_offset = int(label_offset)
        self._precision = precision
        self._train_priority = bool(train_priority)
        self._sharded_bps = int(sharded_blocks_per_sm) if sharded_blocks_per_sm else 0
        dev = labels.device
        src = table if table is not None else sharded.table
        self._tdtype = src.dtype
        ld_self = src.size(1)
        self.slots = [_Slot(self.B, self.f0, self.f1, self.F, ld_self, src.dtype, dev) for _ in range(2)]
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in model.parameters() if p.requires_grad]
        # gradients live in ONE flat buffer (p.grad are views): zeroed, accumulated into by autograd, all-reduced in place
        self._flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self._flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        # the epoch's seeds and the step counter live on the device: the graph itself picks its next batch
        self.max_seeds = int(max_seeds) if max_seeds is not None else 0
        self._seed_table = None
        self._n_seeds = torch.zeros((), dtype=torch.int64, device=dev)
        self._ctr = torch.zeros((), dtype=torch.int64, device=dev)
        self._arange = torch.arange(self.B, dtype=torch.int64, device=dev)
        self._rng_base = int(rng_seed) * 7919 * 1000003
        self._rng_off = torch.zeros(1, dtype=torch.int64, device=dev)
        self.loss_sum = torch.zeros((), device=dev)
        self.graphs = None
        self._g_train = None
        self._side = None

    # ------------------------------------------------------------------------------------------ branch B --
    def _produce(self, s):
        """Mini-batch number ``counter`` into slot ``s``: seeds, both sampled layers, block1, agg0, self0."""
        idx = self._ctr * self.B + self._arange
        ok = idx < self._n_seeds
        picked = self._seed_table[torch.minimum(idx, self._n_seeds - 1).clamp(min=0)]
        s.seeds.copy_(torch.where(ok, picked, torch.full_like(picked, -1)))
        K.sample_neighbors_cap(self.row_ptr, self.col_idx, s.seeds, self.f1, rng_seed=self._rng_base,
                               rng_offset=self._rng_off, out_row_ptr=s.rp1, out_col=s.nbr1)
        K.build_block_cap(s.seeds, s.rp1, s.nbr1, col_pad=s.cap_d0, src_ids=s.src1, col_local=s.col1, counts=s.cnt1)
        K.csr_transpose(s.rp1, s.col1, s.cap_d0, want_perm=True, out=(s.t_rp1, s.t_col1, s.t_perm1))
        K.sample_neighbors_cap(self.row_ptr, self.col_idx, s.src1, self.f0, rng_seed=self._rng_base + 1,
                               rng_offset=self._rng_off, out_row_ptr=s.rp0, out_col=s.nbr0)
        if self.sharded is not None:
            t = self.sharded
            K.spmm_csr_sharded(s.rp0, s.nbr0, t.shard_ptrs, t.part, t.stride_bytes, self.F, dtype=self._tdtype,
                               reduce="mean", out=s.agg0)
            t.fetch(s.src1, out=s.self0)                      # negative (padding) ids read nothing, rows come back zero
        else:
            K.spmm_csr(s.rp0, s.nbr0, self.table, reduce="mean", out=s.agg0, F=self.F)
            K.gather_rows(self.table, s.src1.clamp(min=0), out=s.self0)
        # the loss inputs do not depend on the weights either: the targets and the 1/(valid seeds x world) scale are
        # produced here, off the training branch
        valid = s.seeds >= 0
        lab = self.labels[(s.seeds - self.label_offset).clamp(min=0)]
        s.target.copy_(torch.where(valid, lab, torch.full_like(lab, -100)))
        s.loss_scale.copy_(1.0 / (valid.sum().clamp(min=1).to(torch.float32) * self.world))
        self._ctr += 1
        self._rng_off += 1000003

    # ------------------------------------------------------------------------------------------ branch A --
    def _train(self, s):
        block1 = G.Block(s.rp1, s.col1, s.col1, s._shape, self.B)
        g1 = ops.CsrGraph(s.rp1, s.col1, n_src=s.cap_d0)               # with the transpose the produce branch built
        g1._t, g1._perm = ops.CsrGraph(s.t_rp1, s.t_col1, n_src=self.B), s.t_perm1
        block1._csr = g1
        self0 = s.self0 if s.self0.dtype == torch.float32 else s.self0.float()
        logits = self.model([None, block1], None, pre=(s.agg0, self0))
        # mean over the valid seeds (a slot that holds only padding gives 0, not 0/0), already divided by the world size
        # so that the SUM all-reduce of the gradients is their average
        loss = torch.nn.functional.cross_entropy(logits, s.target, ignore_index=-100, reduction="sum") * s.loss_scale
        self._flat.zero_()
        with ops.direct_weight_grads():                                # dW GEMMs accumulate straight into the flat buffer
            loss.backward()
        if self.world > 1:
            dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=self.group)
        self.opt.step()
        self.loss_sum.add_(loss.detach(), alpha=float(self.world))     # this rank's mean loss

    # ------------------------------------------------------------------------------------------- capture --
    def set_seeds(self, seeds, first_batch=0):
        """Load an epoch's seed list (device tensor) and rewind the in-graph batch counter / RNG offset; batch b then
        draws what ``train.sage_epoch(..., rng_seed=)`` draws for its batch b."""
        n = seeds.numel()
        if self._seed_table is None:
            cap = max(self.max_seeds, n, 1)
            self._seed_table = torch.zeros(cap, dtype=torch.int64, device=self.labels.device)
            if self.graphs is not None:
                raise RuntimeError("seed table created after capture")
        if n > self._seed_table.numel():
            raise ValueError("PipelinedSageTrainer: %d seeds exceed the seed-table capacity %d (max_seeds=)" %
                             (n, self._seed_table.numel()))
        self._seed_table[:n].copy_(seeds)
        self._n_seeds.fill_(n)
        self._ctr.fill_(first_batch)
        self._rng_off.fill_(first_batch * 1000003)

    def capture(self):
        """Warm up (allocator, lazy init, NCCL communicator, optimizer state) and capture the two ping-pong graphs.
        Model weights and optimizer state are put back afterwards: capturing does not train."""
        if self._seed_table is None:
            raise RuntimeError("call set_seeds() before capture()")
        prev = None
        if self._precision is not None:
            prev = ops.get_gemm_precision()
            ops.set_gemm_precision(self._precision)
        self.model.train()
        keep = (self._ctr.clone(), self._rng_off.clone())
        state = [p.detach().clone() for p in self.params]
        saved = {p: {k: v.detach().clone() for k, v in self.opt.state[p].items() if torch.is_tensor(v)}
                 for p in self.params if p in self.opt.state}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for s in self.slots:
                self._produce(s)
                self._train(s)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad():
            for p, st in zip(self.params, state):
                p.copy_(st)
            for p in self.params:
                for k, v in self.opt.state.get(p, {}).items():
                    if torch.is_tensor(v):
                        if p in saved and k in saved[p]:
                            v.copy_(saved[p][k])
                        else:
                            v.zero_()
        self._ctr.copy_(keep[0])
        self._rng_off.copy_(keep[1])
        self.loss_sum.zero_()
        self._side = torch.cuda.Stream()
        # The sharded aggregation of the produce branch sits on NVLink round trips while holding registers: at full
        # residency it takes ~94 % of an SM's register file and the training branch's CTAs cannot start until it drains.
        # Inside the step graph it is therefore captured with its residency capped at 5 blocks (10 warps) per SM — still
        # ~6 MB in flight per GPU, three times what NVLink needs (8 GPUs, heaviest halo load: 0.488 -> 0.462 ms per step
        # together with the stream priority below; either alone: 0.473 / 0.490; tools/ab_train_priority.py).
        bps_prev = K.get_option("rows_sharded_bps")
        if self.sharded is not None and self._sharded_bps:
            K.set_option("rows_sharded_bps", self._sharded_bps)
        # the training branch is the critical path of a step (its last kernels are the all-reduce and Adam): it is captured
        # on a HIGH-priority stream, so that when both branches have blocks waiting the SMs go to the training kernels and
        # the produce branch (whose sharded aggregation holds SMs while its loads cross NVLink) fills what is left
        self._main = torch.cuda.Stream(priority=-1) if self._train_priority else torch.cuda.Stream()
        self.graphs = []
        for k in range(2):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._main):
                cur = torch.cuda.current_stream()
                self._side.wait_stream(cur)
                with torch.cuda.stream(self._side):
                    self._produce(self.slots[1 - k])          # mini-batch i+1 ...
                self._train(self.slots[k])                    # ... while step i trains
                cur.wait_stream(self._side)
            self.graphs.append(g)
        K.set_option("rows_sharded_bps", bps_prev if bps_prev else None)
        self._prologue = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._prologue):
            self._produce(self.slots[0])
        if prev is not None:
            ops.set_gemm_precision(prev)

    # --------------------------------------------------------------------------------------------- epoch --
    def epoch(self, seeds, first_batch=0, stage_events=False):
        """One epoch over ``seeds`` (device int64).  Returns dict(time_s, wall_s, n_batches, loss)."""
        self.set_seeds(seeds, first_batch)
        if self.graphs is None:
            self.capture()
        n = (seeds.numel() + self.B - 1) // self.B
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t_wall = time.perf_counter()
        self.loss_sum.zero_()
        e0.record()
        self._prologue.replay()                               # mini-batch 0 into slot 0
        for i in range(n):
            self.graphs[i & 1].replay()                       # train batch i ‖ produce batch i+1 (all padding after the last)
        e1.record()
        torch.cuda.synchronize()
        return {"time_s": e0.elapsed_time(e1) * 1e-3, "wall_s": time.perf_counter() - t_wall, "n_batches": n,
                "loss": float(self.loss_sum.item()) / max(n, 1)}

    def close(self):
        """Drop the captured graphs (they hold the NCCL all-reduce of this trainer's process group).  Call it — or let the
        trainer be garbage-collected — BEFORE ``dist.destroy_process_group()``: tearing the communicator down while a
        graph that captured its kernels is still alive blocks (seen with tools/ab_train_priority.py at 2 GPUs)."""
        self.graphs = None
        self._prologue = None
        self._g_train = None
        torch.cuda.synchronize()

    def stage_times(self, seeds, steps=20):
        """Device time of each branch ALONE, replayed as its own CUDA graph (no overlap, no Python dispatch) — the
        numbers behind the ``stage_ms`` of bench.py's partitioned extra.  Trains ``steps`` steps on one mini-batch as a
        side effect; call it after the measured epochs."""
        self.set_seeds(seeds)
        if self.graphs is None:
            self.capture()
        if getattr(self, "_g_train", None) is None:
            prev = None
            if self._precision is not None:
                prev = ops.get_gemm_precision()
                ops.set_gemm_precision(self._precision)
            self._g_train = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._g_train):
                self._train(self.slots[0])
            if prev is not None:
                ops.set_gemm_precision(prev)
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        self._prologue.replay()
        torch.cuda.synchronize()
        for i in range(steps):
            ev[i][0].record()
            self._prologue.replay()                           # branch B: the next mini-batch into slot 0
            ev[i][1].record()
            self._g_train.replay()                            # branch A on slot 0
            ev[i][2].record()
        torch.cuda.synchronize()
        pm = sorted(e[0].elapsed_time(e[1]) for e in ev)[steps // 2]
        tm = sorted(e[1].elapsed_time(e[2]) for e in ev)[steps // 2]
        return {"produce_ms": pm, "train_ms": tm}

    def halo_stats(self):
        """Measured share of the input layer's source rows that live on another GPU, from the last produced slot
        (device read-back; call outside the timed region)."""
        s = max(self.slots, key=lambda sl: int(sl.rp0[-1].item()))   # the slot that holds a real mini-batch
        nnz = int(s.rp0[-1].item())
        nbr = s.nbr0[:nnz].long()
        n_dst = int(s.cnt1[0].item())
        if self.sharded is None:
            return {"block0_edges": nnz, "block0_dst_rows": n_dst, "remote_edge_fraction": 0.0,
                    "remote_bytes_per_step": 0}
        t = self.sharded
        owner = nbr // t.part
        remote = int((owner != t.rank).sum().item())
        dst_owner = s.src1[:n_dst] // t.part
        remote_dst = int((dst_owner != t.rank).sum().item())
        esz = t.table.element_size()
        return {"block0_edges": nnz, "block0_dst_rows": n_dst, "remote_edge_fraction": remote / max(nnz, 1),
                "remote_dst_fraction": remote_dst / max(n_dst, 1),
                "remote_bytes_per_step": (remote + remote_dst) * self.F * esz}
