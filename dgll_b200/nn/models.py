"""Block (mini-batch) models with the reference's class names and constructor signatures
(GPU Accelerator/CommGNNModel.py:10-56 ``GCN``, :61-114 ``GraphSAGE``): the "normal" sampled-block forward paths
(``blocks[0].is_block``).  The ``mos`` variants of the reference take DGL graph objects and are out of scope.

``forward(blocks, x, feat_table=None)``: with ``feat_table`` (the HBM-resident feature table) layer 0 aggregates
straight from the table through the block's global column ids — the sampled-block gather is fused into the aggregation
and ``x`` may be ``None``.
"""
import torch

from .. import backend as F
from .block_conv import GraphConv, SAGEConv


class BlockGCN(F.nn.Module):
    """CommGNNModel.py:10-41 (class ``GCN`` there; renamed to avoid clashing with dgll.nn.Convolution.GCN)."""

    def __init__(self, in_feats, n_hidden, n_classes, n_layers, activation, dropout):
        super().__init__()
        assert n_layers > 1
        self.n_layers, self.n_hidden, self.n_classes = n_layers, n_hidden, n_classes
        self.layers = F.nn.ModuleList()
        self.layers.append(GraphConv(in_feats, n_hidden, activation=activation))
        for _ in range(1, n_layers - 1):
            self.layers.append(GraphConv(n_hidden, n_hidden, activation=activation))
        self.layers.append(GraphConv(n_hidden, n_classes))
        self.dropout = F.nn.Dropout(dropout)

    def forward(self, blocks, x):
        h = x
        for l, (layer, block) in enumerate(zip(self.layers, blocks)):
            h = layer(block, h)
            if l != len(self.layers) - 1:
                h = self.dropout(h)
        return h


class GraphSAGE(F.nn.Module):
    """CommGNNModel.py:61-100."""

    def __init__(self, in_feats, n_hidden, n_classes, n_layers, activation, dropout):
        super().__init__()
        self.n_layers, self.n_hidden, self.n_classes = n_layers, n_hidden, n_classes
        self.layers = F.nn.ModuleList()
        if n_layers > 1:
            self.layers.append(SAGEConv(in_feats, n_hidden, "mean"))
            for _ in range(1, n_layers - 1):
                self.layers.append(SAGEConv(n_hidden, n_hidden, "mean"))
            self.layers.append(SAGEConv(n_hidden, n_classes, "mean"))
        else:
            self.layers.append(SAGEConv(in_feats, n_classes, "mean"))
        self.dropout = F.nn.Dropout(dropout)
        self.activation = activation

    def forward(self, blocks, x, feat_table=None, pre=None):
        """``pre=(neigh_mean, feat_dst)``: layer 0's input-feature neighbour mean and destination rows were produced
        upstream (``SAGEConv.forward_preaggregated``); ``blocks[0]`` is then unused and may be ``None``."""
        assert isinstance(blocks, list) and (pre is not None or blocks[0].is_block)
        h = x
        last = len(self.layers) - 1
        for l, (layer, block) in enumerate(zip(self.layers, blocks)):
            if l == 0 and pre is not None:
                fuse = l != last and self.activation is torch.relu        # ReLU in the transform's epilogue
                h = layer.forward_preaggregated(pre[0], pre[1], relu=fuse)
                if fuse:
                    h = self.dropout(h)
                    continue
            elif l == 0 and feat_table is not None:
                h = layer(block, None, feat_table=feat_table)
            else:
                h = layer(block, h)
            if l != len(self.layers) - 1:
                h = self.activation(h)
                h = self.dropout(h)
        return h
