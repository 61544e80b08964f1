"""``Evaluation/PPI`` model classes (BASELINE.json configs[0]) on the B200 kernels: same names, constructor and forward
signatures, parameter names (``layers.N.weight``, ``out_layer.weight/bias``) and initialisers as
Evaluation/PPI/gcn_model.py:63-94.

Differences in mechanism only: the binary adjacency is converted from ``edge_index`` to CSR ONCE per edge_index
tensor (the reference rebuilds an uncoalesced COO in every layer of every step, gcn_model.py:73), ``X·W`` runs on the
device GEMM and ``relu(A·(XW))`` is one aggregation launch with the ReLU fused.  Duplicate edges keep summing exactly
as ``torch.sparse.mm`` on the uncoalesced COO does.
"""
import torch

from .. import backend as F
from .. import ops


def create_sparse_adj(edge_index, num_nodes):
    """gcn_model.py:44-57 — returns the CSR graph the kernels consume (row index = edge_index[0], values = 1)."""
    holder = getattr(edge_index, "_dgllb_csr", None)
    if holder is not None and holder[0] == (edge_index._version, num_nodes):
        return holder[1]
    g = ops.CsrGraph.from_edge_index(edge_index, num_nodes)
    try:
        edge_index._dgllb_csr = ((edge_index._version, num_nodes), g)
    except Exception:
        pass
    return g


class GCNLayer(F.nn.Module):
    """gcn_model.py:63-77 — ``relu(A @ (X @ W))``, W ~ N(0, 1)."""

    def __init__(self, in_features, out_features):
        super().__init__()
        self.weight = F.Parameter(torch.randn(in_features, out_features))

    def forward(self, edge_index, features, num_nodes):
        support = ops.linear(features, self.weight)
        return ops.spmm(create_sparse_adj(edge_index, num_nodes), support, relu=True)


class PPIGCN(F.nn.Module):
    """gcn_model.py:80-94 (class ``GCN`` there)."""

    def __init__(self, in_features, hidden_features, out_features, num_layers):
        super().__init__()
        self.layers = F.nn.ModuleList()
        self.layers.append(GCNLayer(in_features, hidden_features))
        for _ in range(num_layers - 1):
            self.layers.append(GCNLayer(hidden_features, hidden_features))
        self.out_layer = F.nn.Linear(hidden_features, out_features)

    def forward(self, edge_index, features):
        num_nodes = features.size(0)
        x = features
        for layer in self.layers:
            x = layer(edge_index, x, num_nodes)
        return ops.linear(x, self.out_layer.weight, trans_w=True, bias=self.out_layer.bias)
