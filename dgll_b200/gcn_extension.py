"""Drop-in for the reference's ``gcn_extension`` operator module (dgll/FusedKernel/gcn_extension.cpp:103-110).

Same function names, argument order and meaning, same error behaviour (a non-CUDA tensor raises ``RuntimeError``
as the reference's ``TORCH_CHECK(x.is_cuda())`` does, gcn_extension.cpp:31-36,70-76), fresh torch-owned outputs.
Differences, all deliberate: the launch honours the CURRENT torch stream and device and does
not synchronise; the backward returns the TRUE gradients of ``relu(A_hat (X W))`` (the reference kernel's are wrong);
dtype/contiguity are validated instead of reinterpreted.

``GCNFusedFunction`` / ``GCNLayer`` / ``GCN`` mirror dgll/FusedKernel/train_gcn.py:8-57.
"""
import math

import torch

from . import kernels as K


def _check(name, t, dtype):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if t.dtype != dtype:
        raise RuntimeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t.contiguous()


def gcn_fused_forward(row_ptr, col_idx, values, X, W, num_neighbors, actual_F):
    """H[N, H_dim] = relu(A_hat (X[:, :actual_F] W[:actual_F])) — gcn_extension.cpp:22-58."""
    row_ptr = _check("row_ptr", row_ptr, torch.int32)
    col_idx = _check("col_idx", col_idx, torch.int32)
    values = _check("values", values, torch.float32)
    X = _check("X", X, torch.float32)
    W = _check("W", W, torch.float32)
    num_neighbors = _check("num_neighbors", num_neighbors, torch.int32)
    if X.dim() != 2 or W.dim() != 2 or W.size(0) != X.size(1):
        raise RuntimeError("X must be [N, F_padded] and W [F_padded, H]; got %s and %s" % (tuple(X.shape), tuple(W.shape)))
    if row_ptr.numel() != X.size(0) + 1 or num_neighbors.numel() != X.size(0):
        raise RuntimeError("row_ptr must have N+1 and num_neighbors N entries")
    with torch.cuda.device(X.device):
        return K.gcn_fused_forward_v2(row_ptr, col_idx, values, X, W, num_neighbors, int(actual_F))


def gcn_fused_backward(grad_output, row_ptr, col_idx, values, X, W, num_neighbors, actual_F, H=None):
    """[grad_X, grad_W] — gcn_extension.cpp:60-101.  ``H`` (the forward output) is optional: the reference
    signature does not carry it, so it is recomputed for the ReLU mask when absent."""
    grad_output = _check("grad_output", grad_output, torch.float32)
    row_ptr = _check("row_ptr", row_ptr, torch.int32)
    col_idx = _check("col_idx", col_idx, torch.int32)
    values = _check("values", values, torch.float32)
    X = _check("X", X, torch.float32)
    W = _check("W", W, torch.float32)
    num_neighbors = _check("num_neighbors", num_neighbors, torch.int32)
    with torch.cuda.device(X.device):
        if H is None:
            H = K.gcn_fused_forward_v2(row_ptr, col_idx, values, X, W, num_neighbors, int(actual_F))
        gX, gW = K.gcn_fused_backward_v2(grad_output, row_ptr, col_idx, values, X, W, H, num_neighbors, int(actual_F))
    return [gX, gW]


class GCNFusedFunction(torch.autograd.Function):
    """dgll/FusedKernel/train_gcn.py:8-22 (saves the output too, so the backward has the ReLU mask)."""

    @staticmethod
    def forward(ctx, row_ptr, col_idx, values, X, W, num_neighbors, actual_F):
        out = gcn_fused_forward(row_ptr, col_idx, values, X, W, num_neighbors, actual_F)
        ctx.save_for_backward(row_ptr, col_idx, values, X, W, num_neighbors, out)
        ctx.actual_F = actual_F
        return out

    @staticmethod
    def backward(ctx, grad_output):
        row_ptr, col_idx, values, X, W, num_neighbors, out = ctx.saved_tensors
        gX, gW = gcn_fused_backward(grad_output.contiguous(), row_ptr, col_idx, values, X, W, num_neighbors,
                                    ctx.actual_F, H=out)
        return None, None, None, gX, gW, None, None


class GCNLayer(torch.nn.Module):
    """train_gcn.py:24-41, same positional arguments and attribute names: ``GCNLayer(in_features_padded,
    actual_in_features, out_features)``; ``W`` ~ N(0,1)/sqrt(actual_in_features), shape [in_features_padded,
    out_features], created on the device; ``actual_F`` = the number of real input columns."""

    def __init__(self, in_features_padded, actual_in_features, out_features):
        super().__init__()
        scale = 1.0 / math.sqrt(actual_in_features)
        self.W = torch.nn.Parameter(torch.randn(in_features_padded, out_features, dtype=torch.float32, device="cuda") * scale)
        self.actual_F = actual_in_features

    def forward(self, row_ptr, col_idx, values, X, num_neighbors):
        return GCNFusedFunction.apply(row_ptr, col_idx, values, X, self.W, num_neighbors, self.actual_F)


class GCN(torch.nn.Module):
    """train_gcn.py:43-57, same constructor (``GCN(input_dim, hidden_dim, output_dim)``) and parameter shapes:
    ``layer1.W`` [pad4(input_dim), hidden_dim], ``layer2.W`` [hidden_dim, output_dim] — the hidden width is NOT padded
    (the kernels fall back to scalar loads when a row is not a 16-byte multiple), so reference state dicts load."""

    def __init__(self, input_dim, hidden_dim, output_dim):
        super().__init__()
        self.input_dim_padded = ((input_dim + 3) // 4) * 4
        self.layer1 = GCNLayer(self.input_dim_padded, input_dim, hidden_dim)
        self.layer2 = GCNLayer(hidden_dim, hidden_dim, output_dim)

    def forward(self, row_ptr, col_idx, values, X, num_neighbors):
        if X.size(1) != self.input_dim_padded:
            X_padded = torch.zeros(X.size(0), self.input_dim_padded, device=X.device, dtype=torch.float32)
            X_padded[:, :X.size(1)] = X
            X = X_padded
        h1 = self.layer1(row_ptr, col_idx, values, X, num_neighbors)
        return self.layer2(row_ptr, col_idx, values, h1, num_neighbors)
