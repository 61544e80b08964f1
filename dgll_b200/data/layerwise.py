"""Layer-wise importance samplers with the reference's class names — FastGCN and LADIES, "flat" and "WRS" switches
(SURVEY.md §8 f-4).  Constructor and ``sample`` signatures follow the scripts:

    Ladies / LadiesFlat / LadiesWrs / LadiesFlatWrs (fanouts, g, flat=False, HW_row_norm=False)
        GPU Accelerator/MQLadies.py:62-89, MQLadiesFlat.py:62, MQLadiesWrs.py:62, MQLadiesFlatWrs.py:63-90
    FastGCNSampler (fanouts, g)                                              MQFastGCN.py:60-88
    FastGCNSamplerFlat / FastGCNSamplerWrs / FastGCNSamplerFlatWrs (fanouts, g, HW_row_norm=False, flat=False, wrs=False)
        MQFastGCNFlat.py:62, MQFastGCNWrs.py:63, MQFastGCNFlatWrs.py:62-101
    sampler.sample(g, batch_nodes) -> (input_nodes, batch_nodes, blocks)      blocks[0] = input layer

What the reference does per mini-batch with scipy on the host (row slice of the normalised Laplacian, column
probabilities, ``np.random.choice(..., replace=False, p=)``, importance weights, column slice + rescale, CSR emission)
runs here as device kernels (csrc/layerwise.cu) on a Laplacian resident in HBM, stream-ordered with the layer kernels;
two small read-backs per layer.  The Laplacian itself is built once at construction.

``g`` is whatever carries the matrix the reference obtains from ``g.adj_external(scipy_fmt="csr")``: an object with
``.csr()`` -> (row_ptr, col) (``dgll_b200.data.DGraph``), a ``CsrGraph``, or a ``(row_ptr, col)`` pair; entries are
taken as 1.0 (a binary adjacency).  For a symmetric graph the orientation does not matter; for a directed one pass
the same orientation the reference would see.

Differences from the scripts, all deliberate:
  * the draw uses a counter-based generator on the device (same distribution as ``np.random.choice`` without
    replacement, not numpy's stream); pass ``chooser=numpy_chooser`` to draw on the host from ``np.random`` exactly as
    the reference does, or ``replay=[picks_layer0, ...]`` to replay recorded draws (what the parity tests do);
  * the importance weights the reference computes and then drops at ``create_block(('csc', (indptr, indices, [])))``
    are kept on the block as ``edge_weight`` (fp32) / ``edge_weight64``; ``GraphConv(block, h)`` ignores them unless
    asked, exactly like the reference;
  * ``carry``: the FastGCN scripts feed ``block.srcnodes()`` (= arange(num_src), LOCAL ids) back as the next layer's
    rows (MQFastGCN.py:85); ``carry="local"`` reproduces that, the default ``"global"`` carries the picked node ids the
    way the LADIES scripts do (MQLadies.py:86).
"""
import numpy as np
import torch

from .. import graphs as G
from .. import kernels as K

__all__ = ["Ladies", "LadiesFlat", "LadiesWrs", "LadiesFlatWrs", "FastGCNSampler", "FastGCNSamplerFlat",
           "FastGCNSamplerWrs", "FastGCNSamplerFlatWrs", "LayerwiseSampler", "build_laplacian", "numpy_chooser"]


def _adjacency_of(g):
    if hasattr(g, "csr"):
        rp, col = g.csr()
    elif hasattr(g, "row_ptr") and hasattr(g, "col"):
        rp, col = g.row_ptr, g.col
    else:
        rp, col = g
    if not rp.is_cuda:
        rp, col = rp.cuda(), col.cuda()
    return rp.to(torch.int64), col.to(torch.int64)


def build_laplacian(g, kind):
    """(row_ptr int64, col int32, values float64) of the normalised (A + I), columns sorted per row.
    ``kind='row'``: diag(1/rowsum)·(A+I) (utils.py:11-19).  ``kind='sym'``: D^-1/2 (A+I)^T D^-1/2 with
    D = rowsum(A+I) + 1e-20 (utils.py:215-222).  One-time setup on the device; the two degree vectors are raised to
    their powers with numpy on the host so the values are bit-identical to the reference's."""
    rp, col = _adjacency_of(g)
    dev = rp.device
    n = rp.numel() - 1
    rows = torch.repeat_interleave(torch.arange(n, device=dev), rp[1:] - rp[:-1])
    diag = torch.arange(n, device=dev)
    key = torch.cat([rows * n + col, diag * n + diag])
    ukey, cnt = torch.unique(key, return_counts=True)          # sorted by (row, col); A[i,i] = 1 becomes 2
    r, c, v = ukey // n, ukey % n, cnt.to(torch.float64)
    rowsum = torch.zeros(n, dtype=torch.float64, device=dev).index_add_(0, r, v)   # small integers: exact
    if kind == "row":
        with np.errstate(divide="ignore"):
            r_inv = np.power(rowsum.cpu().numpy(), -1.0)
        r_inv[np.isinf(r_inv)] = 0.0
        vals = torch.from_numpy(r_inv).to(dev)[r] * v
        out_rp = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        torch.cumsum(torch.bincount(r, minlength=n), 0, out=out_rp[1:])
        return out_rp, c.to(torch.int32), vals
    if kind != "sym":
        raise ValueError("build_laplacian: kind must be 'row' or 'sym'")
    d = np.power(rowsum.cpu().numpy() + 1e-20, -0.5)
    d[np.isinf(d)] = 0.0
    d = torch.from_numpy(d).to(dev)
    vals = (v * d[c]) * d[r]                                    # (M·D)ᵀ·D: first the column's factor, then the row's
    order = torch.argsort(c * n + r)                            # transpose: new row = c, new col = r
    out_rp = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.bincount(c, minlength=n), 0, out=out_rp[1:])
    return out_rp, r[order].to(torch.int32), vals[order]


def numpy_chooser(n_total, cand_cols, cand_prob, n_cand, fanout):
    """Draw on the host exactly as the reference: ``np.random.choice(n, s_num, p=prob, replace=False)`` on the dense
    probability vector, consuming numpy's global stream.  Returns picks (int64 ndarray, drawing order)."""
    prob = np.zeros(n_total)
    p = cand_prob[:n_cand].cpu().numpy()
    if cand_cols is None:
        prob[:n_cand] = p
    else:
        prob[cand_cols[:n_cand].cpu().numpy()] = p
    s_num = int(min(np.sum(prob > 0), fanout))
    return np.random.choice(n_total, s_num, p=prob, replace=False)


class LayerwiseSampler:
    kind = "ladies"             # 'ladies': probabilities from the sliced rows per layer; 'fastgcn': global, static
    include_batch = False       # MQFastGCN.py:81: union the batch into the picks (sorted), scale 1/p/s_num

    def __init__(self, fanouts, g, flat=False, HW_row_norm=False, wrs=False, carry="global", rng_seed=0,
                 chooser=None, keep_weights=True):
        self.fanouts = list(int(f) for f in fanouts)
        self.layers = len(self.fanouts)
        self.flat, self.wrs = bool(flat), bool(wrs)
        if carry not in ("global", "local"):
            raise ValueError("carry must be 'global' or 'local'")
        self.carry, self.rng_seed, self.chooser, self.keep_weights = carry, int(rng_seed), chooser, keep_weights
        self.lap_rp, self.lap_col, self.lap_val = build_laplacian(g, "row" if self.kind == "ladies" else "sym")
        self.num_nodes = self.lap_rp.numel() - 1
        dev = self.lap_rp.device
        self._pos = torch.full((self.num_nodes,), -1, dtype=torch.int32, device=dev)
        self._calls = 0
        self.prob = None
        if self.kind == "fastgcn":
            # MQFastGCN.py:73-74: one probability vector from the whole Laplacian (recomputed per call there, constant)
            cc, cp, stats = K.col_sqsum(self.lap_col, self.lap_val, self.num_nodes, self.flat)
            n_cand = int(stats[0].item())
            self.prob = torch.zeros(self.num_nodes, dtype=torch.float64, device=dev)
            self.prob[cc[:n_cand].long()] = cp[:n_cand]

    # -- one mini-batch --------------------------------------------------------------------------------------------
    def sample(self, g, batch_nodes, replay=None):
        dev = self.lap_rp.device
        batch = torch.as_tensor(batch_nodes).to(dev).to(torch.int64)
        prev = batch
        blocks = []
        for l, fanout in enumerate(self.fanouts):
            q_rp, q_col, q_val = K.csr_slice_rows(self.lap_rp, self.lap_col, self.lap_val, prev)
            if self.kind == "ladies":
                cand_cols, cand_prob, stats = K.col_sqsum(q_col, q_val, self.num_nodes, self.flat)
            else:
                cand_cols, cand_prob, stats = None, self.prob, None
            sel, picks, count = self._draw(cand_cols, cand_prob, stats, fanout, l, None if replay is None else replay[l])
            use_wrs = self.wrs or self.kind == "ladies"     # every LADIES script goes through estWRS_weights
            drawn = None
            if self.include_batch and not use_wrs:
                cnt = int(count.item())
                drawn = picks[:cnt]
                nxt = torch.unique(torch.cat([drawn, batch]))                        # sorted, as np.unique
                scale = (1.0 / cand_prob[nxt]) / cnt
                weights, count_nxt = None, None
            else:
                scale = K.importance_scale(cand_prob, sel, count, self.num_nodes, "wrs" if use_wrs else "inverse")
                nxt, weights, count_nxt = picks, scale, count
            K.scatter_pos(self._pos, nxt, count_nxt)
            b_rp, b_col, b_val = K.csr_select_cols(q_rp, q_col, q_val, self._pos, scale, with_values=self.keep_weights)
            K.scatter_pos(self._pos, nxt, count_nxt, reset=True)
            if count_nxt is None:
                s_num, nnz = nxt.numel(), int(b_rp[-1].item())
            else:
                s_num, nnz = torch.stack([count[0], b_rp[-1]]).tolist()   # the layer's second read-back
            nxt = nxt[:s_num]
            b_col = b_col[:nnz]
            blk = G.Block(b_rp, b_col, None, nxt, prev.numel())
            blk.col_global = nxt[b_col.long()].to(torch.int32) if nnz else b_col
            blk.dst_global = prev
            blk.picks = picks[:s_num] if drawn is None else drawn
            blk.weights = None if weights is None else weights[:s_num]
            blk.prob = (cand_cols, cand_prob)
            if b_val is not None:
                blk.edge_weight64 = b_val[:nnz]
                blk.edge_weight = blk.edge_weight64.to(torch.float32)
            blocks.append(blk)
            if self.carry == "local":
                num_src = int(b_col.max().item()) + 1 if nnz else 0           # what DGL infers for the block
                prev = torch.arange(num_src, device=dev, dtype=torch.int64)
            else:
                prev = nxt
        blocks.reverse()
        ndata = getattr(g, "ndata", None)
        if ndata is not None:
            if "feat" in ndata:
                blocks[0].srcdata["feat"] = ndata["feat"][prev.to(ndata["feat"].device)]
            if "label" in ndata:
                blocks[-1].dstdata["label"] = ndata["label"][batch.to(ndata["label"].device)]
        self._calls += 1
        return prev, batch, blocks

    def _draw(self, cand_cols, cand_prob, stats, fanout, layer, replay):
        dev = cand_prob.device
        if replay is None and self.chooser is None:
            seed = (self.rng_seed * 1000003 + self._calls) * 131 + layer
            return K.weighted_choice(cand_cols, cand_prob, fanout, seed)
        n_cand = self.num_nodes if stats is None else int(stats[0].item())
        if replay is None:
            replay = self.chooser(self.num_nodes, cand_cols, cand_prob, n_cand, fanout)
        picks_h = torch.as_tensor(np.asarray(replay)).to(torch.int64).to(dev)
        m = picks_h.numel()
        cap = max(fanout, m)
        picks = torch.full((cap,), -1, dtype=torch.int64, device=dev)
        picks[:m] = picks_h
        sel = torch.full((cap,), -1, dtype=torch.int32, device=dev)
        if cand_cols is None:
            sel[:m] = picks_h.to(torch.int32)
        else:
            sel[:m] = torch.searchsorted(cand_cols[:n_cand].contiguous(), picks_h.to(torch.int32)).to(torch.int32)
        return sel, picks, torch.tensor([m], dtype=torch.int64, device=dev)


class Ladies(LayerwiseSampler):
    kind = "ladies"


class LadiesFlat(Ladies):
    pass


class LadiesWrs(Ladies):
    pass


class LadiesFlatWrs(Ladies):
    pass


class FastGCNSampler(LayerwiseSampler):
    """MQFastGCN.py:60-88: plain probabilities, the batch is unioned into every layer's picks."""
    kind = "fastgcn"
    include_batch = True

    def __init__(self, fanouts, g, **kw):
        super().__init__(fanouts, g, flat=False, wrs=False, **kw)


class FastGCNSamplerFlat(LayerwiseSampler):
    kind = "fastgcn"

    def __init__(self, fanouts, g, HW_row_norm=False, flat=False, wrs=False, **kw):
        super().__init__(fanouts, g, flat=flat, HW_row_norm=HW_row_norm, wrs=wrs, **kw)


class FastGCNSamplerWrs(FastGCNSamplerFlat):
    pass


class FastGCNSamplerFlatWrs(FastGCNSamplerFlat):
    pass
