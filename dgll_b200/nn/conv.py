"""``dgll.nn.Convolution`` on the B200 kernels — same class names, constructor and forward signatures, parameter
names/shapes and initialisers as the reference (dgll/nn/Convolution/*.py), so state dicts and call sites carry over.

The aggregation of every layer runs on the hand-written kernels through ``dgll_b200.ops``:
  gcnConv / GraphConvolution   x@W on the device GEMM, Â·(xW)+b on the CSR SpMM (bias fused)      gcnconv.py:9-40, gcn.py:17-48
  NeighborAggregator/sageConv  reduce over K as a fixed-fanout segment SpMM (mean/sum/max)        sageconv.py:10-83
  gatConv                      softmax_j(leakyrelu(a1.Wh_i + a2.Wh_j)) via the fused GAT kernel   gatconv.py:10-57
  sparseGatConv                exp(-leakyrelu(.)) / rowsum via the same kernel (EXP_NEG mode)     gatconv.py:89-151
  SpecialSpmm(Function)        sparse x dense with the SDDMM backward, no dense N x N             gatconv.py:60-86
  GinConv                      Linear(X + A·X) with the dense batched adjacency sparsified        ginconv.py:10-30
Differences from the reference, all documented in SURVEY.md §8: ``NeighborAggregator`` really reduces over K (the
reference discards the reduction, sageconv.py:33-38) and ``sageConv.weight`` is initialised (never is upstream);
attention dropout runs INSIDE the fused kernel with a counter-based mask (same distribution as ``F.dropout`` on the
attention coefficients, not the same draws).
"""
import math

import torch
import torch.nn.functional as Fn

from .. import backend as F
from .. import ops


# ------------------------------------------------------------------- GCN ---
class gcnConv(F.nn.Module):
    """gcnconv.py:9-40 — ``spmm(adj, x @ W) + b``; W, b ~ U(-1/sqrt(out), 1/sqrt(out))."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.weight = F.Parameter(torch.empty(in_features, out_features))
        if bias:
            self.bias = F.Parameter(torch.empty(out_features))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1.0 / math.sqrt(self.weight.size(1))
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    def forward(self, x, adj):
        support = ops.linear(x, self.weight)
        return ops.spmm(adj, support, bias=self.bias)   # bias fused into the aggregation epilogue

    def __repr__(self):
        return "%s (%d -> %d)" % (self.__class__.__name__, self.in_features, self.out_features)


class GraphConvolution(gcnConv):
    """dgll/nn/Convolution/gcn.py:17-48 — duplicate of gcnConv upstream."""


class GCN(F.nn.Module):
    """gcnconv.py:43-58: relu -> dropout -> gcnConv -> log_softmax."""

    def __init__(self, in_features, nhid, nclass, dropout):
        super().__init__()
        self.in_features, self.nhid, self.nclass, self.dropout = in_features, nhid, nclass, dropout
        self.gcn1 = gcnConv(in_features, nhid)
        self.gcn2 = gcnConv(nhid, nclass)

    def forward(self, x, adj):
        h1 = Fn.relu(self.gcn1(x, adj))
        h1_d = Fn.dropout(h1, self.dropout, training=self.training)
        logits = self.gcn2(h1_d, adj)
        return Fn.log_softmax(logits, dim=1)


# ------------------------------------------------------------- GraphSAGE ---
def _fixed_fanout_graph(batch, k, device):
    """CSR with exactly ``k`` entries per row over a [batch*k, F] neighbour table (identity columns)."""
    rp = torch.arange(0, batch * k + 1, k, device=device, dtype=torch.int64)
    col = torch.arange(batch * k, device=device, dtype=torch.int32)
    return ops.CsrGraph(rp, col, n_src=batch * k)


class BinGCNConv(F.nn.Module):
    """Binarized-feature GCN layer (the "Quantization/Binarization" feature the reference names in README.md:11 and its
    architecture diagram but ships no code for; BASELINE.json configs[3]; semantics SURVEY.md §8 a18):

        h_i = ( mean_{j in N(i)} sign(x_j) ) W + b,        sign(x) = +1 for x >= 0, -1 otherwise

    The aggregation runs on the bit-packed popcount SpMM (``dgllb_binarize_pack`` + ``dgllb_bin_spmm_csr``): 4 bytes
    per 32 features per edge instead of 128, integer-exact counts.  Aggregation comes first (the binarized rows are
    the cheap thing to move), the dense transform second.  Backward: straight-through estimator — the sign is treated as
    the identity where ``|x| <= ste_clip`` (``None`` = everywhere) — so dL/dx is ONE transposed aggregation on the fp32
    SpMM kernels; W and b get their ordinary gradients.  ``adj``: anything ``ops.as_csr`` takes (edge values are
    ignored: the binarized mean is over the unweighted neighbourhood)."""

    def __init__(self, in_features, out_features, bias=True, ste_clip=1.0):
        super().__init__()
        self.in_features, self.out_features, self.ste_clip = in_features, out_features, ste_clip
        self.weight = F.Parameter(torch.empty(in_features, out_features))
        if bias:
            self.bias = F.Parameter(torch.empty(out_features))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1.0 / math.sqrt(self.weight.size(1))
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    def forward(self, x, adj):
        agg = ops.binarized_aggregate_ste(adj, x, clip=self.ste_clip)
        return ops.linear(agg, self.weight, bias=self.bias)

    def __repr__(self):
        return "%s (%d -> %d)" % (self.__class__.__name__, self.in_features, self.out_features)


class BinGCN(F.nn.Module):
    """Two ``BinGCNConv`` layers stacked the way ``GCN`` stacks ``gcnConv`` (gcnconv.py:43-58):
    relu -> dropout -> layer -> log_softmax."""

    def __init__(self, in_features, nhid, nclass, dropout, ste_clip=1.0):
        super().__init__()
        self.in_features, self.nhid, self.nclass, self.dropout = in_features, nhid, nclass, dropout
        self.gcn1 = BinGCNConv(in_features, nhid, ste_clip=ste_clip)
        self.gcn2 = BinGCNConv(nhid, nclass, ste_clip=ste_clip)

    def forward(self, x, adj):
        h1 = Fn.relu(self.gcn1(x, adj))
        h1_d = Fn.dropout(h1, self.dropout, training=self.training)
        # centre the hidden activations before taking their sign: after a ReLU every feature would binarize to +1
        logits = self.gcn2(h1_d - h1_d.mean(dim=0, keepdim=True), adj)
        return Fn.log_softmax(logits, dim=1)


class NeighborAggregator(F.nn.Module):
    """sageconv.py:10-45 — reduce over the K axis (mean / sum / max), then ``@ W`` (+ b)."""

    def __init__(self, input_dim, output_dim, use_bias=False, aggr_method="mean"):
        super().__init__()
        self.input_dim, self.output_dim = input_dim, output_dim
        self.use_bias, self.aggr_method = use_bias, aggr_method
        self.weight = F.Parameter(torch.empty(input_dim, output_dim))
        if use_bias:
            self.bias = F.Parameter(torch.empty(output_dim))
        self.reset_parameters()

    def reset_parameters(self):
        F.init.kaiming_uniform_(self.weight)
        if self.use_bias:
            F.init.zeros_(self.bias)

    def aggregate(self, neighbor_feature):
        if self.aggr_method not in ("mean", "sum", "max"):
            raise ValueError("Unsupported aggr_method, expected mean, sum, max, but got {}".format(self.aggr_method))
        b, k, f = neighbor_feature.shape
        flat = neighbor_feature.reshape(b * k, f)
        return ops.spmm(_fixed_fanout_graph(b, k, flat.device), flat, reduce=self.aggr_method)

    def forward(self, neighbor_feature):
        agg = self.aggregate(neighbor_feature)
        return ops.linear(agg, self.weight, bias=self.bias if self.use_bias else None)


class sageConv(F.nn.Module):
    """sageconv.py:48-83 — ``act(src @ W  (+ | cat)  reduce_K(neigh) @ W_n)``."""

    def __init__(self, input_dim, hidden_dim, activation=Fn.relu, aggr_neighbor_method="mean",
                 aggr_hid_method="sum"):
        super().__init__()
        self.input_dim, self.hidden_dim = input_dim, hidden_dim
        self.activation = activation
        self.aggr_hid_method, self.aggr_neighbor_method = aggr_hid_method, aggr_neighbor_method
        self.weight = F.Parameter(torch.empty(input_dim, hidden_dim))
        self.neighborAgg = NeighborAggregator(input_dim, hidden_dim, aggr_method=aggr_neighbor_method)
        self.reset_parameters()

    def reset_parameters(self):
        F.init.kaiming_uniform_(self.weight)

    def forward(self, src_node_features, neighbor_node_features):
        neighbor_hidden = self.neighborAgg(neighbor_node_features)
        self_hidden = ops.linear(src_node_features, self.weight)
        if self.aggr_hid_method == "sum":
            hidden = self_hidden + neighbor_hidden
        elif self.aggr_hid_method == "concat":
            hidden = torch.cat([self_hidden, neighbor_hidden], dim=1)
        else:
            raise ValueError("Expected sum or concat, got {}".format(self.aggr_hid_method))
        return self.activation(hidden) if self.activation else hidden


class GraphSage(F.nn.Module):
    """sageconv.py:86-114 — layer l runs on hops 0..L-l-1 with ``hidden[hop+1].view(B, K_l, -1)``."""

    def __init__(self, input_dim, hidden_dim=[64, 64], num_neighbors_list=[10, 10]):
        super().__init__()
        self.input_dim, self.hidden_dim, self.num_neighbors_list = input_dim, hidden_dim, num_neighbors_list
        self.gcn1 = sageConv(input_dim, hidden_dim[0])
        self.gcn2 = sageConv(hidden_dim[0], hidden_dim[1])
        self.gcn = [self.gcn1, self.gcn2]
        self.num_layers = len(num_neighbors_list)

    def forward(self, node_feature_list):
        hidden = node_feature_list
        for l in range(self.num_layers):
            next_hidden = []
            gcn = self.gcn[l]
            for hop in range(self.num_layers - l):
                src_nodes = hidden[hop]
                src_nums = len(src_nodes)
                h = gcn(src_nodes, hidden[hop + 1].view(src_nums, self.num_neighbors_list[l], -1))
                next_hidden.append(h)
            hidden = next_hidden
        return hidden[0]


# ------------------------------------------------------------------- GAT ---
def _attn_dropout(p, training):
    """Attention dropout probability for this call (gatconv.py:37 ``F.dropout(attention, p, training=)``, :132)."""
    return float(p) if (training and p > 0) else 0.0


def _fused_attention(x, adj, Ws, a_ls, a_rs, slope, mode, elu, dropout):
    """All heads of one attention layer: ``[Wh | el | er] = x @ [W_1..W_H | W_h a_l,h | W_h a_r,h]`` in ONE dense
    transform (``el_h = (x W_h) a_l,h = x (W_h a_l,h)``: the score projections are folded into the weight matrix, a
    parameter-sized product), then the fused aggregation on that buffer (``ops.gat_aggregate_ext``).  Head widths that
    are not a multiple of 4 are zero-padded to one so the kernels take their 128-bit path; the pad columns are
    dropped from the result.  Falls back to separate tensors on non-square graphs (blocks)."""
    H, D = len(Ws), Ws[0].size(1)
    graph = ops.as_csr(adj, binary=True)
    if graph.n_dst != graph.n_src:
        Wh = ops.linear(x, torch.cat(list(Ws), dim=1) if H > 1 else Ws[0])
        Whv = Wh.view(-1, H, D)
        # a block's destination nodes are its first n_dst source nodes (dst-first compaction, MQGCN.py:45,48): their
        # "left" scores are the first n_dst rows; autograd zero-pads the gradient of the slice for the other rows
        el = (Whv * torch.stack(list(a_ls))).sum(-1)[:graph.n_dst]
        er = (Whv * torch.stack(list(a_rs))).sum(-1)
        return ops.gat_aggregate(graph, Wh, el.contiguous(), er.contiguous(), heads=H, slope=slope, mode=mode, elu=elu,
                                 dropout=dropout)
    Dp = (D + 3) // 4 * 4
    cols = [W if Dp == D else Fn.pad(W, (0, Dp - D)) for W in Ws]
    cols += [(W @ a)[:, None] for W, a in zip(Ws, a_ls)]
    cols += [(W @ a)[:, None] for W, a in zip(Ws, a_rs)]
    width = H * Dp + 2 * H
    if width % 4:
        cols.append(x.new_zeros((Ws[0].size(0), 4 - width % 4)))
    ext = ops.linear(x, torch.cat(cols, dim=1))
    out = ops.gat_aggregate_ext(graph, ext, H, Dp, slope=slope, mode=mode, elu=elu, dropout=dropout)
    if Dp != D:
        out = out.view(-1, H, Dp)[:, :, :D].reshape(-1, H * D)
    return out


class gatConv(F.nn.Module):
    """gatconv.py:10-57 — dense-adjacency GAT layer; attention = softmax over ``adj > 0`` of
    ``leakyrelu(Wh a[:D] + (Wh a[D:])^T)``.  The N x N score matrix is never formed: the kernel walks the edges."""

    def __init__(self, in_features, out_features, dropout, alpha, concat=True):
        super().__init__()
        self.dropout, self.in_features, self.out_features = dropout, in_features, out_features
        self.alpha, self.concat = alpha, concat
        self.W = F.Parameter(torch.empty(in_features, out_features))
        F.init.xavier_uniform_(self.W.data, gain=1.414)
        self.a = F.Parameter(torch.empty(2 * out_features, 1))
        F.init.xavier_uniform_(self.a.data, gain=1.414)
        self.leakyrelu = F.LeakyReLU(self.alpha)

    def forward(self, h, adj):
        D = self.out_features
        return _fused_attention(h, adj, [self.W], [self.a[:D, 0]], [self.a[D:, 0]], self.alpha, "softmax", self.concat,
                                _attn_dropout(self.dropout, self.training))

    def __repr__(self):
        return "%s (%d -> %d)" % (self.__class__.__name__, self.in_features, self.out_features)


class SpecialSpmmFunction(torch.autograd.Function):
    """gatconv.py:60-81 — ``sparse_coo(indices, values, shape) @ b`` whose backward is an SDDMM
    (``grad_values[e] = <grad_out[i_e], b[j_e]>``) and ``a^T @ grad_out``; no dense N x N product."""

    @staticmethod
    def forward(ctx, indices, values, shape, b):
        assert indices.requires_grad is False
        n = int(shape[0])
        rows = indices[0].to(torch.int64)
        order = torch.sort(rows, stable=True).indices
        counts = torch.bincount(rows, minlength=n)
        rp = torch.zeros(n + 1, dtype=torch.int64, device=rows.device)
        torch.cumsum(counts, 0, out=rp[1:])
        g = ops.CsrGraph(rp, indices[1][order].to(torch.int32), values.detach()[order], n_src=int(shape[1]))
        ctx.graph, ctx.order = g, order
        ctx.save_for_backward(b)
        from .. import kernels as K
        return K.spmm_csr(g.row_ptr, g.col, b, values=g.values, reduce="sum", n_dst=n)

    @staticmethod
    def backward(ctx, grad_output):
        (b,) = ctx.saved_tensors
        from .. import kernels as K
        g, order = ctx.graph, ctx.order
        grad_values = grad_b = None
        go = grad_output.contiguous()
        if ctx.needs_input_grad[1]:
            gv_sorted = K.sddmm_csr(g.row_ptr, g.col, go, b)
            grad_values = torch.empty_like(gv_sorted)
            grad_values[order] = gv_sorted
        if ctx.needs_input_grad[3]:
            gt = g.transpose()
            grad_b = K.spmm_csr(gt.row_ptr, gt.col, go, values=gt.values, reduce="sum", n_dst=gt.n_dst)
        return None, grad_values, None, grad_b


class SpecialSpmm(F.nn.Module):
    def forward(self, indices, values, shape, b):
        return SpecialSpmmFunction.apply(indices, values, shape, b)


class sparseGatConv(F.nn.Module):
    """gatconv.py:89-151 — ``e_ij = exp(-leakyrelu(a.[Wh_i || Wh_j]))``, ``h'_i = sum_j e_ij Wh_j / sum_j e_ij``.
    Computed with a running-max online softmax (same ratio, no overflow); edges = nonzeros of ``adj``."""

    def __init__(self, in_features, out_features, dropout, alpha, concat=True):
        super().__init__()
        self.in_features, self.out_features, self.alpha, self.concat = in_features, out_features, alpha, concat
        self.W = F.Parameter(torch.zeros(in_features, out_features))
        F.init.xavier_normal_(self.W.data, gain=1.414)
        self.a = F.Parameter(torch.zeros(1, 2 * out_features))
        F.init.xavier_normal_(self.a.data, gain=1.414)
        self.dropout = F.Dropout(dropout)
        self.leakyrelu = F.LeakyReLU(self.alpha)
        self.special_spmm = SpecialSpmm()

    def forward(self, input, adj):
        D = self.out_features
        return _fused_attention(input, adj, [self.W], [self.a[0, :D]], [self.a[0, D:]], self.alpha, "exp_neg",
                                self.concat, _attn_dropout(self.dropout.p, self.training))

    def __repr__(self):
        return "%s (%d -> %d)" % (self.__class__.__name__, self.in_features, self.out_features)


class _MultiHeadMixin:
    """Runs all ``nheads`` attention modules of a reference GAT/SpGAT as ONE multi-head kernel launch: the per-head
    W (and a) are concatenated on the fly, so parameters stay per-module (state-dict compatible, gatconv.py:159-161)."""

    def _heads_forward(self, x, adj, mode):
        atts = self.attentions
        D = atts[0].out_features
        if mode == "softmax":
            a_l, a_r = [m.a[:D, 0] for m in atts], [m.a[D:, 0] for m in atts]
        else:
            a_l, a_r = [m.a[0, :D] for m in atts], [m.a[0, D:] for m in atts]
        pdrop = atts[0].dropout.p if isinstance(atts[0].dropout, torch.nn.Dropout) else atts[0].dropout
        return _fused_attention(x, adj, [m.W for m in atts], a_l, a_r, atts[0].alpha, mode, True,
                                _attn_dropout(pdrop, self.training))


class GAT(F.nn.Module, _MultiHeadMixin):
    """gatconv.py:154-172 — dense-adjacency GAT model."""

    def __init__(self, nfeat, nhid, nclass, dropout, alpha, nheads):
        super().__init__()
        self.dropout = dropout
        self.attentions = [gatConv(nfeat, nhid, dropout=dropout, alpha=alpha, concat=True) for _ in range(nheads)]
        for i, attention in enumerate(self.attentions):
            self.add_module("attention_{}".format(i), attention)
        self.out_att = gatConv(nhid * nheads, nclass, dropout=dropout, alpha=alpha, concat=False)

    def forward(self, x, adj):
        x = Fn.dropout(x, self.dropout, training=self.training)
        x = self._heads_forward(x, adj, "softmax")
        x = Fn.dropout(x, self.dropout, training=self.training)
        x = Fn.elu(self.out_att(x, adj))
        return Fn.log_softmax(x, dim=1)


class SpGAT(F.nn.Module, _MultiHeadMixin):
    """gatconv.py:175-199 — sparse GAT model."""

    def __init__(self, nfeat, nhid, nclass, dropout, alpha, nheads):
        super().__init__()
        self.dropout = dropout
        self.attentions = [sparseGatConv(nfeat, nhid, dropout=dropout, alpha=alpha, concat=True)
                           for _ in range(nheads)]
        for i, attention in enumerate(self.attentions):
            self.add_module("attention_{}".format(i), attention)
        self.out_att = sparseGatConv(nhid * nheads, nclass, dropout=dropout, alpha=alpha, concat=False)

    def forward(self, x, adj):
        x = Fn.dropout(x, self.dropout, training=self.training)
        x = self._heads_forward(x, adj, "exp_neg")
        x = Fn.dropout(x, self.dropout, training=self.training)
        x = Fn.elu(self.out_att(x, adj))
        return Fn.log_softmax(x, dim=1)


# ------------------------------------------------------------------- GIN ---
def _batched_block_diag(Adj):
    """[B, N, N] dense batch -> one CSR over B*N nodes (block diagonal), values kept."""
    B, N, _ = Adj.shape
    nz = (Adj != 0).nonzero()
    rows = nz[:, 0] * N + nz[:, 1]
    cols = nz[:, 0] * N + nz[:, 2]
    vals = Adj[nz[:, 0], nz[:, 1], nz[:, 2]]
    return ops.CsrGraph.from_coo(rows, cols, B * N, B * N, vals)


class GinConv(F.nn.Module):
    """ginconv.py:10-30 — ``relu(Linear(Feat + Adj @ Feat))``; the dense batched product runs as one sparse
    aggregation over the block-diagonal graph (self term fused as the SpMM addend)."""

    def __init__(self, hidden_dim):
        super().__init__()
        self.linear = F.nn.Linear(hidden_dim, hidden_dim)

    def forward(self, Adj, Feat):
        B, N, Fd = Feat.shape
        g = Adj if isinstance(Adj, ops.CsrGraph) else _batched_block_diag(Adj)
        flat = Feat.reshape(B * N, Fd)
        agg = ops.spmm(g, flat) + flat
        X = ops.linear(agg, self.linear.weight, trans_w=True, bias=self.linear.bias, relu=True)
        return X.reshape(B, N, Fd)


class GIN(F.nn.Module):
    """ginconv.py:34-66."""

    def __init__(self, input_dim, hidden_dim, output_dim, n_layers):
        super().__init__()
        self.in_proj = F.nn.Linear(input_dim, hidden_dim)
        self.convs = F.nn.ModuleList()
        for _ in range(n_layers):
            self.convs.append(GinConv(hidden_dim))
        self.out_proj = F.nn.Linear(hidden_dim * (1 + n_layers), output_dim)

    def forward(self, A, X):
        g = _batched_block_diag(A)
        X = ops.linear(X, self.in_proj.weight, trans_w=True, bias=self.in_proj.bias)
        hidden_states = [X]
        for layer in self.convs:
            X = layer(g, X)
            hidden_states.append(X)
        X = torch.cat(hidden_states, dim=2).sum(dim=1)
        return ops.linear(X, self.out_proj.weight, trans_w=True, bias=self.out_proj.bias)
