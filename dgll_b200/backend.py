"""``dgll.backend`` for the B200 path (reference: dgll/__init__.py:1 ``import torch as backend``; rule in example.py:3-10).

The reference's layers are written against ``from dgll import backend as F`` and use the UNION of ``torch``,
``torch.nn`` and ``torch.nn.functional`` names (``F.nn``, ``F.Parameter``, ``F.FloatTensor``, ``F.init``, ``F.mm``,
``F.spmm``, ``F.dropout(x, p, training=)``, ``F.LeakyReLU`` ...; plain ``torch`` lacks several — SURVEY.md §8 b).
This module provides that union by falling through ``torch.nn.functional -> torch.nn -> torch`` and ROUTES THE HOT
PATH to the sm_100a kernels: ``F.mm`` / ``F.matmul`` (2-D fp32 CUDA) -> device GEMM, ``F.spmm`` / ``F.sparse.mm`` ->
CSR aggregation kernel.  CPU tensors passed to the routed ops raise (no CPU fallback).
"""
import sys
import types

import torch
import torch.nn
import torch.nn.functional

from . import ops

nn = torch.nn
init = torch.nn.init
autograd = torch.autograd
optim = torch.optim
Parameter = torch.nn.Parameter


def mm(a, b):
    """F.mm (gcnconv.py:30, gatconv.py:32,117): dense transform on the device GEMM."""
    return ops.linear(a, b)


def matmul(a, b):
    """F.matmul (sageconv.py:40,72; gatconv.py:38,49-50): 2-D x 2-D and batched [B,K,F] x [F,H] go to the device GEMM;
    a sparse left operand goes to the aggregation kernel."""
    if isinstance(a, torch.Tensor) and a.layout != torch.strided:
        return ops.spmm(a, b)
    if b.dim() == 2 and a.dim() >= 2 and a.is_cuda:
        return ops.linear(a, b)
    return torch.matmul(a, b)


def spmm(adj, dense):
    """F.spmm (gcnconv.py:31): sparse(COO/CSR) x dense through the CSR aggregation kernel."""
    return ops.spmm(adj, dense)


class _Sparse(types.ModuleType):
    """F.sparse.mm (Evaluation/PPI/gcn_model.py:76) routed; everything else falls through to torch.sparse."""

    def __getattr__(self, name):
        return getattr(torch.sparse, name)

    @staticmethod
    def mm(adj, dense):
        return ops.spmm(adj, dense)


sparse = _Sparse("dgll_b200.backend.sparse")


def __getattr__(name):
    for mod in (torch.nn.functional, torch.nn, torch):
        if hasattr(mod, name):
            return getattr(mod, name)
    raise AttributeError("dgll backend has no attribute %r" % name)


def install_as_dgll():
    """Make ``from dgll import backend as F`` resolve to this module (for running reference-style scripts)."""
    pkg = sys.modules.get("dgll")
    if pkg is None:
        pkg = types.ModuleType("dgll")
        pkg.__path__ = []
        sys.modules["dgll"] = pkg
    pkg.backend = sys.modules[__name__]
    sys.modules["dgll.backend"] = sys.modules[__name__]
    return pkg
