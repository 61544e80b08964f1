"""One-box multi-GPU plumbing for the aggregation path: node-range partition, halo feature exchange, fused gradient
all-reduce.  One process per GPU, ``torch.distributed`` (NCCL over NVLink on the box, gloo in the CPU tests).

The reference has no partitioned training (SURVEY.md §8 e): its multi-GPU scheme is data parallelism with a
per-parameter ``all_reduce`` (GPU Accelerator/MQGCN.py:55-67).  ``allreduce_gradients`` keeps that semantics in ONE
collective; ``HaloExchange`` is the new piece for graphs whose feature table is sharded by node range
(papers100M-shaped, BASELINE.json configs[4]).

Host logic only lives here.  Row gathers go through ``gather_fn`` — by default the TMA gather kernel
(``kernels.gather_rows``, CUDA only, no fallback); the gloo tests inject a CPU gather to exercise the exchange logic.
"""
import torch
import torch.distributed as dist


def part_size(n_nodes, world):
    return (n_nodes + world - 1) // world


def owner_of(ids, n_nodes, world):
    """Node-range partition: node u belongs to rank ``u // ceil(N/P)`` — no lookup table."""
    return torch.div(ids, part_size(n_nodes, world), rounding_mode="floor")


def local_range(rank, n_nodes, world):
    p = part_size(n_nodes, world)
    return min(rank * p, n_nodes), min((rank + 1) * p, n_nodes)


def _default_gather(table, ids):
    from . import kernels as K  # CUDA only; raises on CPU tensors
    return K.gather_rows(table, ids)


class HaloExchange:
    """Fetch feature rows of arbitrary GLOBAL node ids from a table sharded by node range.

    fetch(ids):  1. bucket ids by owner (stable)                 2. all_to_all_single(counts)
                 3. all_to_all_single(local row offsets)         4. owners gather rows (TMA gather kernel) into the
                 send buffer                                     5. all_to_all_single(rows)   6. undo the bucketing
    Rows owned by the caller never leave the device (they are gathered locally in step 4 like any other bucket, the
    self-bucket of all_to_all_single is a device copy).  ``stats`` counts local/remote rows for the scaling report.
    """

    def __init__(self, n_nodes, local_table, group=None, gather_fn=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_nodes = n_nodes
        self.table = local_table
        self.part = part_size(n_nodes, self.world)
        lo, hi = local_range(self.rank, n_nodes, self.world)
        if local_table.size(0) != hi - lo:
            raise ValueError("rank %d must hold rows [%d, %d) of the table, got %d rows" %
                             (self.rank, lo, hi, local_table.size(0)))
        self.gather_fn = gather_fn or _default_gather
        self.stats = {"rows": 0, "remote_rows": 0, "calls": 0}

    def plan(self, ids):
        """Bucketing + the two small exchanges.  Returns a dict reused by ``exchange`` (lets the caller overlap the id
        exchange of batch k+1 with the aggregation of batch k)."""
        ids = ids.to(torch.int64)
        if self.world == 1:
            self.stats["rows"] += ids.numel()
            self.stats["calls"] += 1
            return {"order": None, "send_split": None, "recv_split": None, "req": ids, "n": ids.numel()}
        owner = owner_of(ids, self.n_nodes, self.world)
        order = torch.sort(owner, stable=True).indices
        send_counts = torch.bincount(owner, minlength=self.world)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=self.group)
        both = torch.stack([send_counts, recv_counts]).tolist()        # ONE device read-back per fetch
        send_split, recv_split = both
        local_off = (ids - owner * self.part)[order].contiguous()
        req = torch.empty(sum(recv_split), dtype=torch.int64, device=ids.device)
        dist.all_to_all_single(req, local_off, recv_split, send_split, group=self.group)
        self.stats["rows"] += ids.numel()
        self.stats["remote_rows"] += ids.numel() - send_split[self.rank]
        self.stats["calls"] += 1
        return {"order": order, "send_split": send_split, "recv_split": recv_split, "req": req, "n": ids.numel()}

    def exchange(self, plan):
        rows_out = self.gather_fn(self.table, plan["req"])            # rows other ranks (and we) asked for
        if plan["order"] is None:
            return rows_out                                           # single rank: everything is local
        width = rows_out.shape[1:]
        recv = torch.empty((plan["n"],) + tuple(width), dtype=rows_out.dtype, device=rows_out.device)
        if self.world > 1:
            dist.all_to_all_single(recv, rows_out.contiguous(), plan["send_split"], plan["recv_split"],
                                   group=self.group)
        else:
            recv.copy_(rows_out)
        out = torch.empty_like(recv)
        out[plan["order"]] = recv
        return out

    def fetch(self, ids):
        return self.exchange(self.plan(ids))

    # ---- fixed-capacity variant: no host read-back, two collectives, static shapes -------------------------
    def fetch_padded(self, ids, slack=1.3, cap=None):
        """Same result as ``fetch`` without any device->host synchronisation: every rank sends each peer a
        fixed-capacity id bucket (``cap`` = ceil(slack * max_rank len(ids) / world) rounded up to 128, agreed once; padded with -1) and
        receives fixed-capacity row buckets back, so both all_to_all_single calls use equal splits known to the host.
        Costs ``slack`` x the bandwidth; under the uniform node-range partition bucket sizes concentrate tightly
        around len(ids)/world.  A bucket that would overflow sets ``self.overflow`` (device flag, checked by the
        caller when convenient, e.g. once per epoch via ``check_overflow``); the overflowing rows come back as zeros."""
        ids = ids.to(torch.int64)
        n = ids.numel()
        if self.world == 1:
            self.stats["rows"] += n
            self.stats["calls"] += 1
            return self.gather_fn(self.table, ids)
        W = self.world
        if cap is None:
            # the capacity must be IDENTICAL on every rank (equal-split collectives): agree on it once, from the
            # largest request of the first call (one all-reduce + read-back), then keep it
            if getattr(self, "_cap", None) is None:
                nmax = torch.tensor([n], dtype=torch.int64, device=ids.device)
                dist.all_reduce(nmax, op=dist.ReduceOp.MAX, group=self.group)
                self._cap = (int(int(nmax.item()) * slack / W) + 128) // 128 * 128
            cap = self._cap
        dev = ids.device
        owner = owner_of(ids, self.n_nodes, W)
        order = torch.sort(owner, stable=True).indices
        owner_s = owner[order]
        counts = torch.bincount(owner, minlength=W)
        start = torch.cumsum(counts, 0) - counts
        pos = torch.arange(n, device=dev) - start[owner_s]                 # position inside the bucket
        ok = pos < cap
        if getattr(self, "overflow", None) is None:
            self.overflow = torch.zeros((), dtype=torch.bool, device=dev)
        self.overflow |= ~ok.all()
        slot = owner_s * cap + pos.clamp(max=cap - 1)                        # flat index into [W, cap]
        # (no boolean-mask indexing anywhere: it would read a size back to the host)
        send_buf = torch.full((W * cap + 1,), -1, dtype=torch.int64, device=dev)
        send_buf[torch.where(ok, slot, torch.full_like(slot, W * cap))] = (ids - owner * self.part)[order]
        send_ids = send_buf[:W * cap]
        recv_ids = torch.empty_like(send_ids)
        dist.all_to_all_single(recv_ids, send_ids, group=self.group)
        rows_out = self.gather_fn(self.table, recv_ids.clamp(min=0))        # padding gathers row 0 (discarded)
        rows_in = torch.empty_like(rows_out)
        dist.all_to_all_single(rows_in, rows_out, group=self.group)
        got = self.gather_fn(rows_in, slot)                                  # bucketed order
        got = got * ok.view((n,) + (1,) * (got.dim() - 1)).to(got.dtype)     # overflowed rows -> zeros
        out = torch.empty_like(got)
        out[order] = got
        self.stats["rows"] += n
        self.stats["remote_rows"] += n - n // W                              # expectation; exact counts stay on device
        self.stats["calls"] += 1
        return out

    def set_bucket_capacity(self, max_ids_per_fetch, slack=1.3):
        """Fix the per-peer bucket capacity of ``fetch_padded`` from a bound every rank agrees on (no collective)."""
        self._cap = (int(max_ids_per_fetch * slack / self.world) + 128) // 128 * 128
        return self._cap

    def check_overflow(self):
        """One read-back: True if any ``fetch_padded`` bucket overflowed since the last check."""
        flag = getattr(self, "overflow", None)
        if flag is None:
            return False
        v = bool(flag.item())
        flag.zero_()
        return v


class PeerShardedTable:
    """The node-range-partitioned feature table of one NVSwitch box, readable in place from every GPU.

    Each rank passes its own shard ``[ceil(N/P), ld]``; the constructor exchanges CUDA-IPC handles once
    (``all_gather_object``), maps the peers' shards and keeps a device array of the P shard pointers.  ``fetch(ids)``
    is then ONE kernel launch (``dgllb_gather_rows_sharded``): remote rows are pulled over NVLink by the gather kernel
    itself — no bucketing, no all_to_all, no host read-back.  All shards must have the same row stride; the last
    shard may be shorter.  (Same result as ``HaloExchange.fetch``; this is the B200-native mechanism.)"""

    def __init__(self, n_nodes, local_table, group=None):
        from . import kernels as K
        self._K = K
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_nodes, self.part = n_nodes, part_size(n_nodes, self.world)
        lo, hi = local_range(self.rank, n_nodes, self.world)
        if local_table.size(0) != hi - lo or not local_table.is_cuda or local_table.stride(1) != 1:
            raise ValueError("rank %d must hold rows [%d, %d) of the table as a row-major CUDA tensor" % (self.rank, lo, hi))
        self.table = local_table
        self.width, self.dtype = local_table.size(1), local_table.dtype
        self.stride_bytes = local_table.stride(0) * local_table.element_size()
        self._mapped = []
        ptrs = [0] * self.world
        ptrs[self.rank] = local_table.data_ptr()
        if self.world > 1:
            handle, off = K.ipc_export(local_table)
            infos = [None] * self.world
            dist.all_gather_object(infos, (handle, off, self.stride_bytes), group=group)
            for r, (h, o, sb) in enumerate(infos):
                if sb != self.stride_bytes:
                    raise ValueError("all shards must share one row stride")
                if r != self.rank:
                    ptrs[r] = K.ipc_import(h, o)
                    self._mapped.append((ptrs[r], o))
            dist.barrier(group=group)
        self.shard_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=local_table.device)
        self.stats = {"rows": 0, "calls": 0}

    def fetch(self, ids, out=None):
        self.stats["rows"] += ids.numel()
        self.stats["calls"] += 1
        return self._K.gather_rows_sharded(self.shard_ptrs, self.part, self.stride_bytes, ids, self.width, self.dtype,
                                           out=out)

    def close(self):
        for ptr, off in self._mapped:
            self._K.ipc_release(ptr, off)
        self._mapped = []


def allreduce_gradients(params, group=None, average=True):
    """Sum (and average) every ``.grad`` in ONE flat-buffer all-reduce (replaces the per-parameter loop of
    GPU Accelerator/MQGCN.py:55-67 — 4-6 latency-bound NCCL calls per step — and DDP buckets, :141-144)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    views = [v.view_as(g) for v, g in zip(flat.split([g.numel() for g in grads]), grads)]
    torch._foreach_copy_(grads, views)          # one multi-tensor launch instead of a copy per parameter


def shard_seeds(seeds, n_nodes, rank, world):
    """Seeds owned by ``rank`` under the node-range partition (batches are local by destination)."""
    lo, hi = local_range(rank, n_nodes, world)
    return seeds[(seeds >= lo) & (seeds < hi)]
