"""Sampled GraphSAGE training epoch on the B200 path — the loop of the reference's mini-batch trainers
(GPU Accelerator/CommGNN_train.py:102-145, MQGCN.py:117-157) with every stage on the device:
sampling (``dgllb_sample_neighbors`` + dst-first compaction), layer-0 aggregation straight from the HBM feature table
(gather fused), dense transforms (exact fp32 or tcgen05 bf16), loss, backward through the same kernels, fused flat
gradient all-reduce (``parallel.allreduce_gradients``) and the optimizer step.
"""
import time

import torch

from . import _nvtx
from . import graphs as G
from . import ops, parallel


def make_batches(row_ptr, col_idx, seeds, fanouts, batch_size, rng_seed=0):
    """Pre-sample every mini-batch of an epoch (mode A, "aggregation epoch": blocks resident on the device)."""
    out = []
    for b, i in enumerate(range(0, seeds.numel(), batch_size)):
        s = seeds[i:i + batch_size]
        out.append((s, G.sample_blocks(row_ptr, col_idx, s, fanouts, rng_seed=rng_seed * 7919 + b)))
    return out


def iter_batches(row_ptr, col_idx, seeds, fanouts, batch_size, rng_seed=0):
    """The sampler in the loop: yields (seeds, blocks) one mini-batch at a time (device sampler + block builder)."""
    for b, i in enumerate(range(0, seeds.numel(), batch_size)):
        s = seeds[i:i + batch_size]
        yield s, G.sample_blocks(row_ptr, col_idx, s, fanouts, rng_seed=rng_seed * 7919 + b)


def sage_epoch(model, opt, table, labels, n_feat, row_ptr=None, col_idx=None, seeds=None, fanouts=(25, 10),
               batch_size=1024, batches=None, rng_seed=0, group=None, precision=None):
    """One epoch.  ``batches`` (from ``make_batches``) = pre-sampled blocks; otherwise the sampler runs in the loop.
    Returns dict(time_s, n_batches, loss, sample_s) — time from CUDA events on the current stream."""
    if precision is not None:
        prev = ops.get_gemm_precision()
        ops.set_gemm_precision(precision)
    model.train()
    params = [p for p in model.parameters() if p.requires_grad]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_wall = time.perf_counter()
    e0.record()
    loss_sum = torch.zeros((), device=table.device)
    n = 0
    it = batches if batches is not None else range(0, seeds.numel(), batch_size)
    for b, item in enumerate(it):
        with _nvtx.range("gpu-load"):                    # FeatureCache/gcn.py:83-87 (sampling + feature fetch)
            if batches is not None:
                s, blocks = item
            else:
                s = seeds[item:item + batch_size]
                blocks = G.sample_blocks(row_ptr, col_idx, s, fanouts, rng_seed=rng_seed * 7919 + b)
        with _nvtx.range("gpu-compute"):                 # FeatureCache/gcn.py:88-93
            logits = model(blocks, None, feat_table=table)   # padded table: SAGEConv slices the logical width itself
            loss = torch.nn.functional.cross_entropy(logits, labels[s])
            opt.zero_grad(set_to_none=True)
            loss.backward()
            parallel.allreduce_gradients(params, group=group)
            opt.step()
        loss_sum += loss.detach()
        n += 1
    e1.record()
    torch.cuda.synchronize()
    if precision is not None:
        ops.set_gemm_precision(prev)
    return {"time_s": e0.elapsed_time(e1) * 1e-3, "wall_s": time.perf_counter() - t_wall, "n_batches": n,
            "loss": float(loss_sum.item()) / max(n, 1)}


def sage_epoch_pipelined(model, opt, table, labels, row_ptr, col_idx, seeds, fanouts=(25, 10), batch_size=1024,
                         rng_seed=0, group=None, precision=None, buffer_size=4):
    """The same epoch as ``sage_epoch`` (sampler in the loop) run through the MQ-GNN style producer/consumer pipeline
    (``dgll_b200.pipeline``): sampling + block construction on a producer thread / stream, training on the consumer."""
    from . import pipeline

    if precision is not None:
        prev = ops.get_gemm_precision()
        ops.set_gemm_precision(precision)

    def batches():
        for b, i in enumerate(range(0, seeds.numel(), batch_size)):
            s = seeds[i:i + batch_size]
            blocks = G.sample_blocks(row_ptr, col_idx, s, fanouts, rng_seed=rng_seed * 7919 + b)
            yield blocks[0].src_ids, s, blocks

    res = pipeline.run_epoch(batches(), model, opt, fetch=None, labels=labels, group=group, BUFFER_SIZE=buffer_size,
                             forward=lambda m, mfgs, feat: m(mfgs, None, feat_table=table))
    if precision is not None:
        ops.set_gemm_precision(prev)
    return res


class GraphedSageTrainer:
    """The training step of ``sage_epoch`` captured ONCE as a CUDA graph and replayed for every mini-batch.

    Block shapes differ per mini-batch, so the step is captured on fixed-capacity buffers: ``batch_size`` seeds,
    ``cap_d0 = batch_size * (1 + fanouts[-1])`` destination rows for the input layer and ``cap_d0 * fanouts[0]`` /
    ``batch_size * fanouts[-1]`` edges.  Loading a mini-batch copies its arrays into those buffers on the device (no
    read-back): rows past the real count get degree 0 (they aggregate to zero and receive zero gradient), unused edge
    slots of the output-layer block hold the padding column id that ``dgllb_csr_transpose`` drops, padded seeds of a
    short last batch are masked out of the loss.  The result is the same update as the eager step (checked in
    tests/test_gpu_layers.py) with the host doing ~10 small copies and one graph launch per mini-batch instead of
    ~100 Python-dispatched calls — the eager epoch is host-bound.

    2-layer ``dgll_b200.nn.GraphSAGE``.  Input features either come straight from a resident ``table`` (gather fused
    into the input layer's aggregation) or, with ``table=None, n_feat=F``, as the rows of the block's source nodes
    handed to ``load(..., x=)`` — the partitioned case, where they were just fetched from the owning GPUs
    (``parallel.PeerShardedTable``).  The optimizer is captured too when it is capturable
    (``torch.optim.Adam(..., fused=True, capturable=True)``) and there is no gradient all-reduce, otherwise it (and
    the all-reduce) run eagerly after the replay."""

    def __init__(self, model, opt, table, labels, batch_size=1024, fanouts=(25, 10), group=None, precision=None,
                 n_feat=None, capture_collectives=False, label_offset=0):
        if len(fanouts) != 2 or len(model.layers) != 2:
            raise ValueError("GraphedSageTrainer: 2-layer models / two fanouts")
        if table is None and n_feat is None:
            raise ValueError("GraphedSageTrainer: pass the resident feature table, or n_feat for per-batch feature rows")
        dev = labels.device if table is None else table.device
        self.model, self.opt, self.table, self.labels, self.group = model, opt, table, labels, group
        self.label_offset = int(label_offset)   # labels may be this rank's slice of a node-range partition
        self.B = int(batch_size)
        self.cap_d0 = self.B * (1 + int(fanouts[-1]))
        self.cap_e0 = self.cap_d0 * int(fanouts[0])
        self.cap_e1 = self.B * int(fanouts[-1])
        self.params = [p for p in model.parameters() if p.requires_grad]
        i64, i32 = torch.int64, torch.int32
        self.seeds = torch.zeros(self.B, dtype=i64, device=dev)
        self.valid = torch.ones(self.B, dtype=torch.bool, device=dev)
        self.rp0 = torch.zeros(self.cap_d0 + 1, dtype=i32, device=dev)
        self.col0 = torch.zeros(self.cap_e0, dtype=i32, device=dev)
        self.ids0 = torch.zeros(self.cap_d0, dtype=i64, device=dev)
        self.x = None
        if table is None:
            self.cap_src = self.cap_d0 * (1 + int(fanouts[0]))
            self.x = torch.zeros((self.cap_src, int(n_feat)), dtype=torch.float32, device=dev)
            self._src_shape = torch.empty(self.cap_src, dtype=torch.int8, device=dev)   # only its length is used
        self.rp1 = torch.zeros(self.B + 1, dtype=i32, device=dev)
        self.col1 = torch.full((self.cap_e1,), self.cap_d0, dtype=i32, device=dev)
        self.loss = torch.zeros((), device=dev)
        self.loss_sum = torch.zeros((), device=dev)
        self.graph = None
        self._precision = precision
        world = parallel.dist.get_world_size(group) if parallel.dist.is_initialized() else 1
        self._opt_in_graph = world == 1 and bool(opt.defaults.get("capturable", False))
        # With more than one rank and capture_collectives, the gradient all-reduce (NCCL is graph-capturable) and the
        # optimizer are captured too.  Gradients live in ONE flat buffer (p.grad are views): zeroed, accumulated into by autograd,
        # all-reduced in place; the 1/world average is folded into the loss scale.  One launch per step on every rank.
        self._world = world
        self._flat = None
        if capture_collectives and world > 1:
            if not opt.defaults.get("capturable", False):
                raise ValueError("capture_collectives needs a capturable optimizer")
            self._flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=dev)
            off = 0
            for p in self.params:
                p.grad = self._flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            self._opt_in_graph = True

    # ---- sampler inside the graph (fixed-capacity sampler + block builder, no size read-backs) ------------------
    def enable_device_sampler(self, row_ptr, col_idx, rng_seed=0):
        """Put the neighbour sampler and the block builder INSIDE the captured step (resident-table mode): one graph
        launch per mini-batch, nothing read back.  Uses the fixed-capacity kernels (``dgllb_sample_neighbors_cap`` /
        ``dgllb_build_block_cap``: negative ids are padding, unused edge slots hold the padding column) writing straight
        into the capture buffers; the random seed lives in device memory and is bumped by the graph itself, following
        the same schedule as ``sage_epoch(..., rng_seed=)`` / ``graphs.sample_blocks`` so the two draw identical blocks.
        Call before ``capture()``; then drive with ``step_sampled(seeds)`` / ``epoch_sampled(seeds)``."""
        if self.x is not None:
            raise ValueError("device sampler in the graph: resident-table mode only")
        dev = self.seeds.device
        self._smp = (row_ptr, col_idx)
        self._rng_base = int(rng_seed) * 7919 * 1000003           # (rng_seed*7919 + b)*1000003 + layer, b via the offset
        self._rng_off = torch.zeros(1, dtype=torch.int64, device=dev)
        self._nbr1 = torch.zeros(self.cap_e1, dtype=torch.int32, device=dev)     # global ids of block1's neighbours
        self._src1 = torch.full((self.cap_d0,), -1, dtype=torch.int64, device=dev)
        self._cnt1 = torch.zeros(3, dtype=torch.int32, device=dev)
        self.graph = None

    def _sample_into_buffers(self):
        from . import kernels as K
        rp, col = self._smp
        f0, f1 = self.cap_e0 // self.cap_d0, self.cap_e1 // self.B
        K.sample_neighbors_cap(rp, col, self.seeds, f1, rng_seed=self._rng_base, rng_offset=self._rng_off,
                               out_row_ptr=self.rp1, out_col=self._nbr1)
        K.build_block_cap(self.seeds, self.rp1, self._nbr1, col_pad=self.cap_d0, src_ids=self._src1, col_local=self.col1,
                          counts=self._cnt1)
        K.sample_neighbors_cap(rp, col, self._src1, f0, rng_seed=self._rng_base + 1, rng_offset=self._rng_off,
                               out_row_ptr=self.rp0, out_col=self.col0)
        torch.clamp(self._src1, min=0, out=self.ids0)            # padding rows gather row 0 (their output is unused)
        self.valid.copy_(self.seeds >= 0)
        self._rng_off += 1000003

    def step_sampled(self, seeds):
        """One training step on ``seeds`` (<= batch_size ids) with the sampler inside the replayed graph."""
        ns = seeds.numel()
        if ns > self.B:
            raise ValueError("GraphedSageTrainer: more seeds than the captured batch size")
        self.seeds[:ns].copy_(seeds)
        if ns < self.B:
            self.seeds[ns:].fill_(-1)
        self.step()

    def epoch_sampled(self, seeds, first_batch=0):
        """One epoch over ``seeds`` in batches of ``batch_size``; the in-graph RNG offset is set so that batch b draws
        what ``sage_epoch`` draws for it."""
        if getattr(self, "_smp", None) is None:
            raise ValueError("call enable_device_sampler(row_ptr, col_idx) first")
        if self.graph is None:
            self.seeds.copy_(torch.nn.functional.pad(seeds[:self.B], (0, max(0, self.B - seeds[:self.B].numel())), value=-1))
            self.capture()
        self._rng_off.fill_(first_batch * 1000003)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t_wall = time.perf_counter()
        self.loss_sum.zero_()
        e0.record()
        n = 0
        for i in range(0, seeds.numel(), self.B):
            self.step_sampled(seeds[i:i + self.B])
            n += 1
        e1.record()
        torch.cuda.synchronize()
        return {"time_s": e0.elapsed_time(e1) * 1e-3, "wall_s": time.perf_counter() - t_wall, "n_batches": n,
                "loss": float(self.loss_sum.item()) / max(n, 1)}

    def _blocks(self):
        b0 = G.Block(self.rp0, self.col0, self.col0, self.ids0 if self.x is None else self._src_shape, self.cap_d0)
        b1 = G.Block(self.rp1, self.col1, self.col1, self.ids0, self.B)
        return [b0, b1]

    def load(self, seeds, blocks, x=None):
        """Copy one mini-batch into the capture buffers (device-side, stream-ordered, no synchronisation).
        ``x``: feature rows of ``blocks[0]``'s source nodes (per-batch feature mode only)."""
        b0, b1 = blocks
        n0, n1, ns = b0.num_dst, b1.num_dst, seeds.numel()
        col0 = b0.col_global if self.x is None else b0.col
        e0, e1 = col0.numel(), b1.col.numel()
        if n0 > self.cap_d0 or e0 > self.cap_e0 or e1 > self.cap_e1 or ns > self.B or n1 != ns:
            raise ValueError("GraphedSageTrainer: mini-batch exceeds the captured capacities")
        if self.x is not None:
            if x is None or x.size(0) > self.cap_src or x.size(1) != self.x.size(1):
                raise ValueError("GraphedSageTrainer: per-batch feature rows missing or larger than the capacity")
            self.x[:x.size(0)].copy_(x)
        self.seeds[:ns].copy_(seeds)
        if ns < self.B:
            self.seeds[ns:].zero_()
        self.valid[:ns] = True
        self.valid[ns:] = False
        self.rp0[:n0 + 1].copy_(b0.row_ptr)
        self.rp0[n0 + 1:].copy_(b0.row_ptr[-1:].expand(self.cap_d0 - n0))
        self.col0[:e0].copy_(col0)
        if self.x is None:
            self.ids0[:n0].copy_(b0.src_ids[:n0])
            self.ids0[n0:].zero_()
        self.rp1[:n1 + 1].copy_(b1.row_ptr)
        if n1 < self.B:
            self.rp1[n1 + 1:].copy_(b1.row_ptr[-1:].expand(self.B - n1))
        self.col1[:e1].copy_(b1.col)
        self.col1[e1:].fill_(self.cap_d0)

    def _step_body(self):
        if getattr(self, "_smp", None) is not None:
            self._sample_into_buffers()
        if self.x is None:
            logits = self.model(self._blocks(), None, feat_table=self.table)
        else:
            logits = self.model(self._blocks(), self.x)
        target = torch.where(self.valid, self.labels[(self.seeds - self.label_offset).clamp(min=0)], torch.full_like(self.seeds, -100))
        loss = torch.nn.functional.cross_entropy(logits, target, ignore_index=-100)
        if self._flat is not None:
            self._flat.zero_()
            (loss / self._world).backward()                     # accumulates into the views of the flat buffer
            parallel.dist.all_reduce(self._flat, op=parallel.dist.ReduceOp.SUM, group=self.group)
        else:
            loss.backward()
        self.loss.copy_(loss.detach())
        self.loss_sum += loss.detach()
        if self._opt_in_graph:
            self.opt.step()

    def capture(self):
        prev = None
        if self._precision is not None:
            prev = ops.get_gemm_precision()
            ops.set_gemm_precision(self._precision)
        self.model.train()
        state = [p.detach().clone() for p in self.params]
        # optimizer state must exist BEFORE the capture (state created inside it would be re-zeroed by every replay):
        # the warm-up steps create it, then it is put back to its pre-warm-up values in place
        saved = {p: {k: v.detach().clone() for k, v in self.opt.state[p].items() if torch.is_tensor(v)}
                 for p in self.params if p in self.opt.state}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # warm-up off the capture stream (allocator, lazy init)
            for _ in range(2):
                if self._flat is None:
                    self.opt.zero_grad(set_to_none=True)
                self._step_body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad():                               # the warm-up must not count as training
            for p, s in zip(self.params, state):
                p.copy_(s)
        if self._opt_in_graph:
            for p in self.params:
                for k, v in self.opt.state.get(p, {}).items():
                    if torch.is_tensor(v):
                        if p in saved and k in saved[p]:
                            v.copy_(saved[p][k])
                        else:
                            v.zero_()
        self.loss_sum.zero_()
        if self._flat is None:
            self.opt.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._step_body()
        if prev is not None:
            ops.set_gemm_precision(prev)

    def step(self):
        """Replay the captured step on the loaded mini-batch."""
        self.graph.replay()
        if not self._opt_in_graph:
            parallel.allreduce_gradients(self.params, group=self.group)
            self.opt.step()

    def epoch(self, batches, overlap=None, callback=None):
        """One epoch over pre-sampled ``batches`` (``make_batches``) or any iterable of (seeds, blocks[, x]).

        ``overlap`` (default: on for iterators, off for lists): mini-batch i+1 is PRODUCED — whatever the iterator does:
        device sampler, block builder, feature fetch from the peers — on a side stream while the GPU replays step i.
        The producer's read-backs then synchronise only the side stream, so the host no longer waits for the training
        graph before it can sample the next batch; the copy into the capture buffers stays on the main stream, ordered
        after the previous replay.  ``callback(stage, i)`` with stage "before" (the load of step i is about to be
        enqueued) / "after" (step i is enqueued) runs on the main stream — a hook for per-stage CUDA events."""
        if self.graph is None:
            first = batches[0] if isinstance(batches, (list, tuple)) else None
            if first is not None:
                self.load(*first)
            self.capture()
        if overlap is None:
            overlap = not isinstance(batches, (list, tuple))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t_wall = time.perf_counter()
        self.loss_sum.zero_()
        e0.record()
        n = 0
        if not overlap:
            for item in batches:
                if callback:
                    callback("before", n)
                self.load(*item)
                self.step()
                if callback:
                    callback("after", n)
                n += 1
        else:
            main = torch.cuda.current_stream()
            if getattr(self, "_side", None) is None:
                self._side = torch.cuda.Stream()
            side = self._side
            side.wait_stream(main)
            it = iter(batches)

            def produce():
                with torch.cuda.stream(side):
                    try:
                        item = next(it)
                    except StopIteration:
                        return None
                    ev = torch.cuda.Event()
                    ev.record(side)
                return item, ev

            nxt = produce()
            while nxt is not None:
                item, ev = nxt
                main.wait_event(ev)
                if callback:
                    callback("before", n)
                self.load(*item)
                for t in _tensors_of(item):
                    t.record_stream(main)      # allocated on the side stream, read by the copies on the main stream
                self.step()
                if callback:
                    callback("after", n)
                n += 1
                nxt = produce()                 # the host samples batch i+1 while the GPU replays step i
        e1.record()
        torch.cuda.synchronize()
        return {"time_s": e0.elapsed_time(e1) * 1e-3, "wall_s": time.perf_counter() - t_wall, "n_batches": n,
                "loss": float(self.loss_sum.item()) / max(n, 1)}


def _tensors_of(item):
    """Every tensor of a (seeds, blocks[, x]) mini-batch."""
    seeds, blocks = item[0], item[1]
    out = [seeds]
    for b in blocks:
        for name in ("row_ptr", "col", "col_global", "src_ids"):
            t = getattr(b, name, None)
            if torch.is_tensor(t) and t.is_cuda:
                out.append(t)
    if len(item) > 2 and torch.is_tensor(item[2]):
        out.append(item[2])
    return out
