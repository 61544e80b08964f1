"""Edge-value construction helpers with the reference's names.

Host (scipy) functions keep the reference signatures and arithmetic:
  normalize / matrix_row_normalize    D^-1 M                     dgll/nn/utils/utils.py:240-247, GPU Accelerator/utils.py:12-20
  normalize_lap                       D^-1/2 M D^-1/2            GPU Accelerator/utils.py:215-222
  sparse_mx_to_torch_sparse_tensor    scipy -> torch sparse COO  dgll/nn/utils/utils.py:250-257
  accuracy                                                        dgll/nn/utils/utils.py:260-264
Device functions build the same edge values straight on a ``CsrGraph`` (no host round trip), for graphs that
already live in HBM:
  row_normalize_csr(g)   values / row sum        (= normalize)
  sym_normalize_csr(g)   d_i^-1/2 v d_j^-1/2     (= FusedKernel/train_gcn.py:64-71 and normalize_lap; 1/sqrt(0) := 0)
"""
import numpy as np
import scipy.sparse as sp
import torch

from .. import ops


def normalize(mx):
    """Row-normalize sparse matrix (utils.py:240-247)."""
    rowsum = np.array(mx.sum(1))
    with np.errstate(divide="ignore"):
        r_inv = np.power(rowsum, -1.0).flatten()
    r_inv[np.isinf(r_inv)] = 0.0
    return sp.diags(r_inv).dot(mx)


matrix_row_normalize = normalize


def normalize_lap(adj):
    """D^-1/2 adj D^-1/2 (GPU Accelerator/utils.py:215-222; row sums + 1e-20)."""
    rowsum = np.array(adj.sum(1)) + 1e-20
    d_inv_sqrt = np.power(rowsum, -0.5).flatten()
    d_inv_sqrt[np.isinf(d_inv_sqrt)] = 0.0
    d = sp.diags(d_inv_sqrt, 0)
    return adj.dot(d).transpose().dot(d)


def sparse_mx_to_torch_sparse_tensor(sparse_mx, device=None):
    """scipy sparse -> torch sparse COO float32 (utils.py:250-257); ``device`` places it on the GPU directly."""
    sparse_mx = sparse_mx.tocoo().astype(np.float32)
    indices = torch.from_numpy(np.vstack((sparse_mx.row, sparse_mx.col)).astype(np.int64))
    values = torch.from_numpy(sparse_mx.data)
    t = torch.sparse_coo_tensor(indices, values, torch.Size(sparse_mx.shape))
    return t.to(device) if device is not None else t


def accuracy(output, labels):
    """utils.py:260-264."""
    preds = output.max(1)[1].type_as(labels)
    correct = preds.eq(labels).double()
    return correct.sum() / len(labels)


# ---- device-side equivalents on a CsrGraph -------------------------------------------------------------------
def _values_or_ones(g):
    return g.values if g.values is not None else torch.ones(g.col.numel(), dtype=torch.float32, device=g.device)


def _row_of_edge(g):
    deg = (g.row_ptr[1:] - g.row_ptr[:-1])
    return torch.repeat_interleave(torch.arange(g.n_dst, device=g.device), deg)


def row_normalize_csr(g):
    """``normalize`` on the device: values[e] / sum of row(e); empty / zero-sum rows stay zero."""
    g = ops.as_csr(g)
    v = _values_or_ones(g)
    rows = _row_of_edge(g)
    rowsum = torch.zeros(g.n_dst, dtype=torch.float32, device=g.device).index_add_(0, rows, v)
    inv = torch.where(rowsum != 0, 1.0 / rowsum, torch.zeros_like(rowsum))
    return g.with_values(v * inv[rows])


def sym_normalize_csr(g):
    """``D^-1/2 A D^-1/2`` on the device with D = row sums and 1/sqrt(0) := 0 (square graphs)."""
    g = ops.as_csr(g)
    v = _values_or_ones(g)
    rows = _row_of_edge(g)
    d = torch.zeros(g.n_dst, dtype=torch.float32, device=g.device).index_add_(0, rows, v)
    dis = torch.where(d > 0, d.rsqrt(), torch.zeros_like(d))
    return g.with_values(dis[rows] * v * dis[g.col.long()])
