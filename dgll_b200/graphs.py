"""Synthetic graphs of the BASELINE.json shapes and device-side block (MFG) construction.

Graph generator (SURVEY.md §8 d): R-MAT (a,b,c,d) = (0.57,0.19,0.19,0.05), ids hash-permuted then ``mod N``,
self loops removed, duplicates removed, then padded with uniform edges / truncated to the exact nnz; CSR by
destination (in-edges) with ``col_idx`` sorted within a row.  Everything is generated on the device with a
seeded ``torch.Generator`` so every rank/box reproduces the same graph.

Block construction follows the DGL convention the reference's GPU-Accelerator scripts rely on
(GPU Accelerator/MQGCN.py:45,48): the src id space of a block starts with its dst nodes in the same order
(``to_block`` first-occurrence order), so ``h_dst = h_src[:num_dst]``.
"""
import math

import torch

from . import kernels as K

SHAPES = {
    # name: (N, nnz, F, classes)
    "reddit": (232965, 114615892, 602, 41),
    "products": (2449029, 61859140, 100, 47),       # undirected E; symmetrised nnz = 2E (+N self loops)
    "papers100m": (111059956, 1615685872, 128, 172),
}


def _rmat_edges(n_nodes, n_edges, gen, device, abcd=(0.57, 0.19, 0.19, 0.05), chunk=1 << 26):
    scale = max(1, math.ceil(math.log2(max(n_nodes, 2))))
    a, b, c, _ = abcd
    src_parts, dst_parts = [], []
    done = 0
    while done < n_edges:
        m = min(chunk, n_edges - done)
        src = torch.zeros(m, dtype=torch.int64, device=device)
        dst = torch.zeros(m, dtype=torch.int64, device=device)
        for _ in range(scale):
            r = torch.rand(m, device=device, generator=gen)
            src_bit = (r >= a + b).to(torch.int64)                     # quadrants c, d
            dst_bit = ((r >= a) & (r < a + b) | (r >= a + b + c)).to(torch.int64)  # quadrants b, d
            src = (src << 1) | src_bit
            dst = (dst << 1) | dst_bit
        src_parts.append(src)
        dst_parts.append(dst)
        done += m
    return torch.cat(src_parts), torch.cat(dst_parts), scale


def _hash_perm(ids, scale, mult=0x9E3779B1, add=0x7F4A7C15):
    """Bijective scramble on [0, 2^scale): odd multiplier + xor-shift, breaks R-MAT's id/degree correlation."""
    mask = (1 << scale) - 1
    x = (ids * mult + add) & mask
    x = x ^ (x >> max(scale // 2, 1))
    x = (x * (0x85EBCA6B | 1)) & mask
    return x


def rmat_csr(n_nodes, nnz, seed=0, device="cuda", symmetric=False, index64=None):
    """CSR by destination of a synthetic R-MAT graph with exactly ``nnz`` distinct directed edges, no self loops.
    Returns (row_ptr int64[N+1], col_idx int32[nnz])."""
    gen = torch.Generator(device=device).manual_seed(seed)
    want = nnz // 2 if symmetric else nnz
    keys = torch.empty(0, dtype=torch.int64, device=device)
    need = want
    rounds = 0
    while need > 0:
        extra = int(need * (1.25 if rounds == 0 else 1.5)) + 1024
        if rounds < 2:
            src, dst, scale = _rmat_edges(n_nodes, extra, gen, device)
            src = _hash_perm(src, scale) % n_nodes
            dst = _hash_perm(dst, scale, mult=0xC2B2AE35, add=0x27D4EB2F) % n_nodes
        else:  # top up with uniform edges (R-MAT saturates its hot corner)
            src = torch.randint(0, n_nodes, (extra,), device=device, generator=gen)
            dst = torch.randint(0, n_nodes, (extra,), device=device, generator=gen)
        ok = src != dst
        src, dst = src[ok], dst[ok]
        if symmetric:
            lo, hi = torch.minimum(src, dst), torch.maximum(src, dst)
            k = hi * n_nodes + lo
        else:
            k = dst * n_nodes + src
        del src, dst, ok
        keys = torch.unique(torch.cat([keys, k]))
        del k
        if keys.numel() > want:
            # drop a seeded random subset of the surplus, keep sorted order
            perm = torch.randperm(keys.numel(), device=device, generator=gen)[:want]
            keys = keys[torch.sort(perm).values]
        need = want - keys.numel()
        rounds += 1
    if symmetric:
        hi, lo = keys // n_nodes, keys % n_nodes
        keys = torch.sort(torch.cat([hi * n_nodes + lo, lo * n_nodes + hi])).values
    dst = keys // n_nodes
    col = (keys % n_nodes).to(torch.int32)
    del keys
    counts = torch.bincount(dst, minlength=n_nodes)
    row_ptr = torch.zeros(n_nodes + 1, dtype=torch.int64, device=device)
    torch.cumsum(counts, 0, out=row_ptr[1:])
    return row_ptr, col


def _rmat_keys_into(buf, n_nodes, gen, device, chunk=1 << 26):
    """Fill ``buf`` (int64) with dst*N+src keys of hash-permuted R-MAT edges, a chunk at a time; self loops get key -1."""
    done = 0
    while done < buf.numel():
        m = min(chunk, buf.numel() - done)
        src, dst, scale = _rmat_edges(n_nodes, m, gen, device, chunk=chunk)
        src = _hash_perm(src, scale) % n_nodes
        dst = _hash_perm(dst, scale, mult=0xC2B2AE35, add=0x27D4EB2F) % n_nodes
        k = dst * n_nodes + src
        k[src == dst] = -1
        buf[done:done + m] = k
        del src, dst, k
        done += m


def rmat_csr_large(n_nodes, nnz, seed=0, device="cuda"):
    """``rmat_csr`` for graphs of billions of edges (papers100M-shaped: 1.6 B): the same generator, hash permutation,
    self-loop and duplicate removal and uniform top-up, but keys are produced chunk-wise into ONE buffer and the surplus
    is dropped by uniform thinning of the sorted keys, so the peak is ~4 key arrays (~50 GB at 1.6 B edges) instead of
    ~10.  Returns (row_ptr int64[N+1], col_idx int32[nnz])."""
    gen = torch.Generator(device=device).manual_seed(seed)
    keys = torch.empty(0, dtype=torch.int64, device=device)
    rounds = 0
    while keys.numel() < nnz:
        need = nnz - keys.numel()
        extra = int(need * (1.2 if rounds == 0 else 1.5)) + 1024
        buf = torch.empty(extra, dtype=torch.int64, device=device)
        if rounds < 2:
            _rmat_keys_into(buf, n_nodes, gen, device)
        else:  # top up with uniform edges (R-MAT saturates its hot corner)
            for o in range(0, extra, 1 << 27):
                m = min(1 << 27, extra - o)
                src = torch.randint(0, n_nodes, (m,), device=device, generator=gen)
                dst = torch.randint(0, n_nodes, (m,), device=device, generator=gen)
                k = dst * n_nodes + src
                k[src == dst] = -1
                buf[o:o + m] = k
                del src, dst, k
        if keys.numel():
            buf = torch.cat([keys, buf])
        del keys
        keys = torch.unique(buf)
        del buf
        if keys.numel() and int(keys[0].item()) < 0:
            keys = keys[1:]
        rounds += 1
    if keys.numel() > nnz:
        # uniform thinning of the sorted key list: keep nnz entries at evenly spaced ranks
        idx = torch.div(torch.arange(nnz, device=device, dtype=torch.int64) * keys.numel(), nnz, rounding_mode="floor")
        keys = keys[idx]
        del idx
    col = (keys % n_nodes).to(torch.int32)
    keys = torch.div(keys, n_nodes, rounding_mode="floor")
    row_ptr = torch.zeros(n_nodes + 1, dtype=torch.int64, device=device)
    torch.cumsum(torch.bincount(keys, minlength=n_nodes), 0, out=row_ptr[1:])
    return row_ptr, col


def uniform_csr(n_nodes, deg, seed=0, device="cuda"):
    """Control graph: every row has exactly ``deg`` uniform random in-neighbours (duplicates allowed)."""
    gen = torch.Generator(device=device).manual_seed(seed)
    col = torch.randint(0, n_nodes, (n_nodes * deg,), device=device, generator=gen, dtype=torch.int32)
    row_ptr = torch.arange(0, n_nodes * deg + 1, deg, device=device, dtype=torch.int64)
    return row_ptr, col


def feature_table(n_rows, F, seed=0, device="cuda", dtype=torch.float32, pad_to=None):
    """N(0,1) features in a table whose row stride is a 16-byte multiple (602 -> 604 fp32 / 608 bf16); pad
    columns are zero.  Returns the padded table; use ``table[:, :F]`` as the logical view."""
    if pad_to is None:
        q = 16 // torch.empty(0, dtype=dtype).element_size()
        pad_to = (F + q - 1) // q * q
    gen = torch.Generator(device=device).manual_seed(seed)
    t = torch.zeros((n_rows, pad_to), dtype=dtype, device=device)
    step = max(1, (1 << 28) // max(pad_to, 1))
    for r0 in range(0, n_rows, step):
        r1 = min(n_rows, r0 + step)
        t[r0:r1, :F] = torch.randn((r1 - r0, F), device=device, generator=gen).to(dtype)
    return t


class Block:
    """One message-flow-graph block: CSR by destination over a compact src id space whose first ``num_dst``
    entries are the dst nodes (DGL ``to_block`` convention).  ``src_ids`` maps compact src -> global node id;
    ``col_global`` keeps the global ids so layer 0 can aggregate straight from the feature table (gather fused
    into the aggregation)."""
    is_block = True

    def __init__(self, row_ptr, col_local, col_global, src_ids, num_dst):
        self.row_ptr = row_ptr
        self.col = col_local
        self.col_global = col_global
        self.src_ids = src_ids
        self.num_dst = int(num_dst)
        self.num_src = int(src_ids.numel())
        self.srcdata, self.dstdata = {}, {}

    def num_dst_nodes(self):
        return self.num_dst

    def num_src_nodes(self):
        return self.num_src

    def num_edges(self):
        return int(self.col.numel())

    @property
    def dst_ids(self):
        return self.src_ids[:self.num_dst]


def compact_dst_first(dst_ids, nbr_global):
    """Relabel: unique ids of cat(dst_ids, nbr_global) in first-occurrence order (dst first, in order).
    Returns (src_ids int64[num_src], col_local int32[len(nbr_global)]).  ``dst_ids`` must be unique."""
    allv = torch.cat([dst_ids.to(torch.int64), nbr_global.to(torch.int64)])
    uniq, inv = torch.unique(allv, return_inverse=True)
    first = torch.full((uniq.numel(),), allv.numel(), dtype=torch.int64, device=allv.device)
    first.scatter_reduce_(0, inv, torch.arange(allv.numel(), device=allv.device), reduce="amin")
    order = torch.argsort(first)
    rank = torch.empty_like(order)
    rank[order] = torch.arange(order.numel(), device=order.device)
    local = rank[inv]
    return uniq[order], local[dst_ids.numel():].to(torch.int32)


def sample_blocks(row_ptr, col_idx, seeds, fanouts, rng_seed=0, builder="device"):
    """Device-side DGL-style neighbour sampling: ``fanouts[0]`` applies to the INPUT layer, the last entry to the
    seed/output layer (the order DGLLNeighborSampler consumes them, dgll/sampling/dgllsampler.py:14).
    Returns blocks[0..L-1] (input layer first), each dst-first compacted.
    ``builder='device'``: ``dgllb_build_block`` (hash-based, one read-back per layer); ``'torch'``: the sort-based
    formulation with torch.unique (kept as the cross-check; identical output)."""
    blocks = []
    cur = seeds.to(torch.int64)
    for li, fanout in enumerate(reversed(list(fanouts))):
        b_rp, b_col = K.sample_neighbors(row_ptr, col_idx, cur, fanout, rng_seed=rng_seed * 1000003 + li)
        if builder == "device":
            src_cap, col_cap, counts = K.build_block(cur, b_rp, b_col)
            num_src, nnz = counts.tolist()                      # the layer's only device->host read-back
            src_ids, col_local, b_col = src_cap[:num_src], col_cap[:nnz], b_col[:nnz]
        else:
            nnz = int(b_rp[-1].item())
            b_col = b_col[:nnz]
            src_ids, col_local = compact_dst_first(cur, b_col)
        blocks.insert(0, Block(b_rp, col_local, b_col, src_ids, cur.numel()))
        cur = src_ids
    return blocks
