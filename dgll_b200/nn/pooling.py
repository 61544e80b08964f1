"""``dgll.nn.GlobalPooling`` (dgll/nn/GlobalPooling/Pooling.py:18-119) on the segment-reduce kernel.

``torch_scatter.scatter(x, batch, dim=0, dim_size=size, reduce=...)`` over a batch vector is a segment reduction:
it runs as the CSR aggregation kernel with identity columns (sum / mean / max; empty graphs give 0 like
torch_scatter).  ``batch=None`` pools everything into one row exactly as the reference's early return does.
"""
from typing import List, Optional, Union

import torch

from .. import backend as F
from .. import ops


def sumPooling(x, batch: Optional[torch.Tensor], size: Optional[int] = None):
    """Pooling.py:18-37."""
    return ops.segment_reduce(x, batch, size, "sum")


def meanPooling(x, batch: Optional[torch.Tensor], size: Optional[int] = None):
    """Pooling.py:40-59."""
    return ops.segment_reduce(x, batch, size, "mean")


def maxPooling(x, batch: Optional[torch.Tensor], size: Optional[int] = None):
    """Pooling.py:62-81."""
    return ops.segment_reduce(x, batch, size, "max")


class Pooling(F.nn.Module):
    """Pooling.py:83-119 — one or several of 'sum'/'add'/'mean'/'max', concatenated on the last dimension."""

    def __init__(self, aggr: Union[str, List[str]]):
        super().__init__()
        self.aggrs = [aggr] if isinstance(aggr, str) else aggr
        assert len(self.aggrs) > 0
        assert len(set(self.aggrs) | {"sum", "add", "mean", "max"}) == 4

    def forward(self, x, batch: Optional[torch.Tensor], size: Optional[int] = None):
        xs = []
        for aggr in self.aggrs:
            if aggr in ("sum", "add"):
                xs.append(sumPooling(x, batch, size))
            elif aggr == "mean":
                xs.append(meanPooling(x, batch, size))
            elif aggr == "max":
                xs.append(maxPooling(x, batch, size))
        return xs[0] if len(xs) == 1 else torch.cat(xs, dim=-1)

    def __repr__(self):
        aggr = self.aggrs[0] if len(self.aggrs) == 1 else self.aggrs
        return f"{self.__class__.__name__}(aggr={aggr})"
