"""``DGraph`` — the reference's in-tree graph store (dgll/data/dgraph.py:18-132) with the adjacency also held as a
device CSR so samplers and gathers run on the GPU.  Constructor and the 7 getters keep the reference signatures.

``edges`` is the reference's format: a list where ``edges[v]`` is the python list of v's neighbours (stored order is
significant: the samplers keep it).  A ``(row_ptr, col)`` tensor pair is accepted too (large synthetic graphs).
"""
import torch

from .. import kernels as K
from .. import ops


class DGraph(object):
    def __init__(self, nodes=None, edges=None, labels=None, features=None, train_mask=None, test_mask=None,
                 validation_mask=None, device=None):
        self.nodes = nodes
        self.edges = edges
        self.labels = labels
        self.features = features
        self.train_mask = train_mask
        self.test_mask = test_mask
        self.validation_mask = validation_mask
        self.device = torch.device(device) if device is not None else (
            features.device if isinstance(features, torch.Tensor) and features.is_cuda else None)
        self._csr = None

    # ---- device CSR of the adjacency lists (built lazily, order preserved) ----
    def csr(self):
        if self._csr is None:
            if isinstance(self.edges, (tuple, list)) and len(self.edges) == 2 and isinstance(self.edges[0], torch.Tensor):
                rp, col = self.edges
            else:
                deg = torch.tensor([len(e) for e in self.edges], dtype=torch.int64)
                rp = torch.zeros(len(self.edges) + 1, dtype=torch.int64)
                torch.cumsum(deg, 0, out=rp[1:])
                flat = [w for e in self.edges for w in e]
                col = torch.tensor(flat, dtype=torch.int32) if flat else torch.zeros(0, dtype=torch.int32)
            dev = self.device if self.device is not None else torch.device("cuda")
            self._csr = (rp.to(dev), col.to(dev).to(torch.int32))
        return self._csr

    def get_neighbors(self, nodes):
        """dgraph.py:49-62 — list of neighbour lists, one per node (host lists, as the reference returns)."""
        if isinstance(self.edges, (tuple, list)) and len(self.edges) == 2 and isinstance(self.edges[0], torch.Tensor):
            rp, col = (t.cpu() for t in self.edges)
            return [col[rp[int(v)]:rp[int(v) + 1]].tolist() for v in nodes]
        return [self.edges[int(v)] for v in nodes]

    def get_induced_subgraph(self, nodes):
        """dgraph.py:64-81 — dense int32 [n, n]; ``adj[pos(u), pos(w)] = 1`` for ``w in edges[u]`` with w in nodes."""
        rp, col = self.csr()
        dev = rp.device
        nodes_d = nodes.to(dev).to(torch.int64)
        n = nodes_d.numel()
        n_total = rp.numel() - 1
        pos = torch.full((n_total,), -1, dtype=torch.int64, device=dev)
        # later occurrences win, as in the reference's dict comprehension (:77)
        pos.scatter_reduce_(0, nodes_d, torch.arange(n, device=dev), reduce="amax", include_self=True)
        result = torch.zeros((n, n), dtype=torch.int32, device=dev)
        if n == 0:
            return result
        deg = (rp[1:] - rp[:-1])[nodes_d]
        row_of_edge = torch.repeat_interleave(pos[nodes_d], deg)
        starts = rp[nodes_d]
        offs = torch.arange(int(deg.sum().item()), device=dev) - torch.repeat_interleave(
            torch.cumsum(deg, 0) - deg, deg)
        nbr = col[(torch.repeat_interleave(starts, deg) + offs)].to(torch.int64)
        cpos = pos[nbr]
        keep = cpos >= 0
        result[row_of_edge[keep], cpos[keep]] = 1
        return result.to(nodes.device) if not nodes.is_cuda else result

    def get_labels(self, nodes):
        """dgraph.py:83-93."""
        return self.labels[nodes.to(self.labels.device)] if isinstance(self.labels, torch.Tensor) else self.labels[nodes]

    def get_features(self, nodes):
        """dgraph.py:95-105 — ``features[nodes]`` through the TMA row-gather kernel when the table is on the GPU."""
        if isinstance(self.features, torch.Tensor) and self.features.is_cuda:
            return ops.gather_rows(self.features, nodes.to(self.features.device))
        raise RuntimeError("dgll_b200: DGraph.features must be a CUDA tensor (there is no CPU fallback); "
                           "use GraphCacheServer for host-resident tables")

    def get_train_nodes(self):
        return self.nodes[self.train_mask]

    def get_validation_nodes(self):
        return self.nodes[self.validation_mask]

    def get_test_nodes(self):
        return self.nodes[self.test_mask]
