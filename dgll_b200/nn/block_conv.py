"""Block (MFG) layers with the DGL names the reference's GPU-Accelerator scripts call under the alias ``dgll``:
``dgll.nn.GraphConv(in, out, activation=)`` and ``dgll.nn.SAGEConv(in, out, 'mean')`` used as ``layer(block, h)`` or
``layer(block, (h_src, h_dst))`` (GPU Accelerator/CommGNNModel.py:23-28,36-39,72-77,96-100; MQGCN.py:41-50).

The arithmetic upstream lives in DGL 2.4 (third-party, not in the reference tree); semantics restated from its
public documentation — "parity unpinned" by reference tests, pinned by oracle/layers.py:
  GraphConv(norm='both'): h = D_in^-1/2 * (A (D_out^-1/2 * x_src)) W + b, degrees from the block, clamped >= 1,
                          W applied before the aggregation iff in > out
  SAGEConv('mean'):       h = W_self x_dst + W_neigh mean_{j in N(i)} x_j + b, empty mean = 0,
                          W_neigh applied before the aggregation iff in > out
A block is any object with ``row_ptr`` / ``col`` (CSR by destination over a dst-first compact src space),
``num_dst_nodes()`` and ``num_src_nodes()`` — ``dgll_b200.graphs.Block`` or ``dgll_b200.data.create_block``.
Layer 0 may aggregate straight from the global feature table (gather fused): pass ``feat_table=`` and the block's
``col_global``.
"""
import torch

from .. import backend as F
from .. import ops


def _graph_of(block, use_global=False, n_table_rows=None):
    """CsrGraph view of a block, cached on the block (no device read-back: n_src comes from the caller)."""
    key = "_csr_global" if use_global else "_csr"
    g = getattr(block, key, None)
    if g is None:
        if use_global:
            g = ops.CsrGraph(block.row_ptr, block.col_global, n_src=n_table_rows)
        else:
            g = ops.CsrGraph(block.row_ptr, block.col, n_src=block.num_src_nodes())
        setattr(block, key, g)
    return g


class GraphConv(F.nn.Module):
    def __init__(self, in_feats, out_feats, norm="both", weight=True, bias=True, activation=None,
                 allow_zero_in_degree=False):
        super().__init__()
        if norm not in ("none", "both", "right", "left"):
            raise ValueError('Invalid norm value. Must be either "none", "both", "right" or "left".')
        self._in_feats, self._out_feats, self._norm = in_feats, out_feats, norm
        self._activation = activation
        if weight:
            self.weight = F.Parameter(torch.empty(in_feats, out_feats))
        else:
            self.register_parameter("weight", None)
        if bias:
            self.bias = F.Parameter(torch.empty(out_feats))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        if self.weight is not None:
            F.init.xavier_uniform_(self.weight)
        if self.bias is not None:
            F.init.zeros_(self.bias)

    def forward(self, block, feat, edge_weight=None):
        """``edge_weight`` (DGL's keyword): per-edge scalars multiplied into the messages; the degree normalisation
        stays that of the unweighted block.  The layer-wise samplers leave their importance weights on
        ``block.edge_weight``; like the reference's scripts the layer ignores them unless they are passed here."""
        feat_src = feat[0] if isinstance(feat, tuple) else feat
        g = _graph_of(block)
        n_dst = block.num_dst_nodes()
        if self._norm in ("left", "both"):
            out_deg = torch.bincount(g.col.long(), minlength=feat_src.size(0)).to(feat_src.dtype).clamp(min=1)
            norm = out_deg.pow(-0.5) if self._norm == "both" else 1.0 / out_deg
            feat_src = feat_src * norm[:, None]
        w_first = self.weight is not None and self._in_feats > self._out_feats
        if w_first:
            feat_src = ops.linear(feat_src, self.weight)
        rst = ops.spmm(g if edge_weight is None else g.with_values(edge_weight.to(torch.float32)), feat_src,
                       reduce="sum")
        if self._norm in ("right", "both"):
            in_deg = g.degrees().clamp(min=1)
            rst = rst * (in_deg.pow(-0.5) if self._norm == "both" else 1.0 / in_deg)[:, None]
        if self.weight is not None and not w_first:
            rst = ops.linear(rst, self.weight)
        if self.bias is not None:
            rst = rst + self.bias
        if self._activation is not None:
            rst = self._activation(rst)
        return rst[:n_dst]


class SAGEConv(F.nn.Module):
    def __init__(self, in_feats, out_feats, aggregator_type="mean", feat_drop=0.0, bias=True, norm=None,
                 activation=None):
        super().__init__()
        if aggregator_type not in ("mean", "gcn", "pool"):
            raise KeyError("Invalid aggregator_type. Must be one of mean/gcn/pool (lstm is not supported on this path)")
        self._in_src_feats = self._in_dst_feats = in_feats
        self._out_feats, self._aggre_type = out_feats, aggregator_type
        self.norm, self.activation = norm, activation
        self.feat_drop = F.nn.Dropout(feat_drop)
        if aggregator_type == "pool":
            self.fc_pool = F.nn.Linear(in_feats, in_feats)
        self.fc_neigh = F.nn.Linear(in_feats, out_feats, bias=False)
        if aggregator_type != "gcn":
            self.fc_self = F.nn.Linear(in_feats, out_feats, bias=False)
        if bias:
            self.bias = F.Parameter(torch.zeros(out_feats))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        gain = F.init.calculate_gain("relu")
        if self._aggre_type == "pool":
            F.init.xavier_uniform_(self.fc_pool.weight, gain=gain)
        if self._aggre_type != "gcn":
            F.init.xavier_uniform_(self.fc_self.weight, gain=gain)
        F.init.xavier_uniform_(self.fc_neigh.weight, gain=gain)

    def forward_preaggregated(self, h_neigh_mean, feat_dst, relu=False):
        """The layer applied to an input-feature neighbour MEAN computed upstream (``h_neigh_mean`` [n_dst, in]) and
        the destination rows ``feat_dst`` [n_dst, >= in]: ``W_self x_dst + W_neigh mean_j x_j + b``.  The mean of raw
        input features does not depend on the weights, so a pipelined trainer aggregates mini-batch i+1 — straight from
        the (possibly partitioned) feature table — while step i trains (``dgll_b200.pipelined``).  Same result as
        ``forward`` up to the order of the linear map and the mean (DGL's ``lin_before_mp`` choice)."""
        if self._aggre_type != "mean":
            raise ValueError("forward_preaggregated: 'mean' aggregator only")
        # two GEMMs into one output; bias (and, when the caller's activation is a ReLU, the ReLU) in the second's epilogue
        rst = ops.linear2(feat_dst[:, :self._in_dst_feats], self.fc_self.weight, h_neigh_mean, self.fc_neigh.weight,
                          bias=self.bias, relu=relu, trans_w=True)
        if self.activation is not None:
            rst = self.activation(rst)
        if self.norm is not None:
            rst = self.norm(rst)
        return rst

    def forward(self, block, feat, feat_table=None):
        """``feat`` = h_src or (h_src, h_dst).  With ``feat_table`` (layer 0) the neighbour mean is taken straight
        from the global table through ``block.col_global`` and only the dst rows are gathered."""
        n_dst = block.num_dst_nodes()
        if feat_table is not None:
            g = _graph_of(block, use_global=True, n_table_rows=feat_table.size(0))
            feat_src = feat_table
            feat_dst = ops.gather_rows(feat_table, block.src_ids[:n_dst])
        else:
            g = _graph_of(block)
            if isinstance(feat, tuple):
                feat_src, feat_dst = self.feat_drop(feat[0]), self.feat_drop(feat[1])
            else:
                feat_src = self.feat_drop(feat)
                feat_dst = feat_src[:n_dst]
        lin_before_mp = self._in_src_feats > self._out_feats and feat_table is None
        wn = self.fc_neigh.weight            # [out, in]: used with trans_w=True, no transpose copy
        if self._aggre_type == "mean":
            x_dst = feat_dst[:, :self._in_dst_feats]
            if lin_before_mp:    # self term and bias ride in the aggregation's epilogue: no separate add kernels
                rst = ops.spmm(g, ops.linear(feat_src, wn, trans_w=True), reduce="mean",
                               addend=ops.linear(x_dst, self.fc_self.weight, trans_w=True), bias=self.bias)
            else:                # two GEMMs into one output
                rst = ops.linear2(x_dst, self.fc_self.weight, ops.spmm(g, feat_src, reduce="mean", F=self._in_src_feats),
                                  wn, bias=self.bias, trans_w=True)
            if self.activation is not None:
                rst = self.activation(rst)
            if self.norm is not None:
                rst = self.norm(rst)
            return rst
        elif self._aggre_type == "gcn":
            s = ops.spmm(g, feat_src, reduce="sum", F=self._in_src_feats)
            h_neigh = ops.linear((s + feat_dst[:, :self._in_src_feats]) / (g.degrees()[:, None] + 1), wn, trans_w=True)
        else:  # pool
            pooled = ops.linear(feat_src, self.fc_pool.weight, trans_w=True, bias=self.fc_pool.bias, relu=True)
            h_neigh = ops.linear(ops.spmm(g, pooled, reduce="max"), wn, trans_w=True)
        if self._aggre_type == "gcn":
            rst = h_neigh
        else:
            rst = ops.linear(feat_dst[:, :self._in_dst_feats], self.fc_self.weight, trans_w=True,
                             bias=self.bias) + h_neigh
        if self.bias is not None and self._aggre_type == "gcn":
            rst = rst + self.bias
        if self.activation is not None:
            rst = self.activation(rst)
        if self.norm is not None:
            rst = self.norm(rst)
        return rst
