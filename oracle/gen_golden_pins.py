"""Generate tests/golden/cache_server.npz and tests/golden/nn_sageconv_fixed.npz by EXECUTING the reference's own
``GraphCacheServer`` (dgll/FeatureCache/storage.py:12-221) and ``sageConv`` / ``GraphSage``
(dgll/nn/Convolution/sageconv.py:10-114) in place.

Run in the build container only (needs /root/reference mounted, read-only):

    python oracle/gen_golden_pins.py [--ref /root/reference]

Nothing is copied: each file is parsed with ``ast`` and its class definitions are executed where they lie.

* storage.py imports numba / dgl (0.4 contrib API) and calls ``.cuda(gpuid)``; here those names resolve to the smallest
  possible stand-ins — ``Frame`` / ``FrameRef`` keep the dict they are given, ``.cuda(...)`` returns the tensor itself,
  ``torch.cuda.LongTensor/FloatTensor`` are the CPU constructors, the memory queries ``auto_cache`` uses return numbers
  chosen so that the capacity formula (:71-78) yields the wanted capability.  Every line of cache bookkeeping, masking,
  gather and miss accounting that runs is the reference's.
* sageconv.py is executed with the two one-line repairs SURVEY.md §8 a8 documents, applied to the AST, not to a copy:
  (1) the reductions of NeighborAggregator.forward (:33-38) are assigned (``neighbor_feature = neighbor_feature.mean(dim=1)``;
  ``.max(dim=1)`` additionally takes ``.values``); (2) ``sageConv.__init__`` ends with ``self.reset_parameters()`` so
  ``weight`` is initialised (:63-68).  Nothing else changes.

TEST INFRASTRUCTURE ONLY.
"""
import argparse
import ast
import contextlib
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import make_backend  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


# ------------------------------------------------------------------ GraphCacheServer --
class _Frame:
    def __init__(self, d):
        self.d = d


def _frameref(frame):
    return frame.d


class _Col:
    def __init__(self, t):
        self.data = t


class _Store:
    """graph._node_frame._frame[name].data -> host tensor (storage.py:120-122)."""

    def __init__(self, tables):
        self._node_frame = types.SimpleNamespace(_frame={k: _Col(v) for k, v in tables.items()})


class _Mapping:
    def __init__(self, t):
        self.t = t

    def tousertensor(self):
        return self.t


class _NodeFlow:
    """The DGL 0.4 NodeFlow surface fetch_data / fetch_from_cache touch (:163-166,:204)."""

    def __init__(self, layers):
        self.num_layers = len(layers)
        self._layers = layers
        self._node_mapping = _Mapping(torch.cat(layers))
        offs = [0]
        for l in layers:
            offs.append(offs[-1] + len(l))
        self._layer_offsets = offs
        self._node_frames = [None] * self.num_layers

    def layer_parent_nid(self, i):
        return self._layers[i]


class _FakeCuda(types.ModuleType):
    """torch.cuda as storage.py uses it, on the CPU.  ``mem`` makes auto_cache's formula return the wanted capability."""

    def __init__(self):
        super().__init__("torch.cuda")
        self.avail = 0

    @staticmethod
    def LongTensor(*a):
        return torch.LongTensor(*a)

    @staticmethod
    def FloatTensor(*a):
        return torch.FloatTensor(*a)

    @staticmethod
    def device(_):
        return contextlib.nullcontext()

    def max_memory_allocated(self, device=None):
        return 0

    def max_memory_cached(self, device=None):
        return 0

    def get_device_properties(self, _):
        return types.SimpleNamespace(total_memory=self.avail + 1024 * 1024 * 1024)


class _TorchProxy(types.ModuleType):
    def __init__(self, cuda):
        super().__init__("torch")
        self.cuda = cuda

    def __getattr__(self, name):
        return getattr(torch, name)


def run_cache_server(ref):
    path = os.path.join(ref, "dgll", "FeatureCache", "storage.py")
    tree = ast.parse(open(path).read(), filename=path)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "GraphCacheServer"]
    assert len(cls) == 1
    fake_cuda = _FakeCuda()
    ns = {"torch": _TorchProxy(fake_cuda), "Frame": _Frame, "FrameRef": _frameref, "np": np, "print": lambda *a, **k: None}
    exec(compile(ast.Module(body=cls, type_ignores=[]), path, "exec"), ns)
    GraphCacheServer = ns["GraphCacheServer"]
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self           # the build container has no GPU
    try:
        rng = np.random.default_rng(7)
        n_full, n_local = 900, 600
        feats = torch.from_numpy(rng.standard_normal((n_full, 13)).astype(np.float32))
        norm = torch.from_numpy(rng.random((n_full, 1)).astype(np.float32))
        nid_map = torch.from_numpy(rng.permutation(n_full)[:n_local].astype(np.int64))
        out_deg = torch.from_numpy(rng.integers(0, 50, size=n_local).astype(np.int64))
        names = ["features", "norm"]
        layers = [torch.from_numpy(rng.integers(0, n_local, size=m).astype(np.int64)) for m in (257, 64, 9)]
        out = {"feats": feats.numpy(), "norm": norm.numpy(), "nid_map": nid_map.numpy(), "out_deg": out_deg.numpy(),
               "layer0": layers[0].numpy(), "layer1": layers[1].numpy(), "layer2": layers[2].numpy()}
        for tag, cap in (("part", 150), ("full", n_local)):
            srv = GraphCacheServer(_Store({"features": feats, "norm": norm}), n_local, nid_map, 0)
            srv.init_field(names)
            assert srv.total_dim == 14
            fake_cuda.avail = cap * srv.total_dim * 4            # storage.py:71-78 -> capability == cap
            srv.auto_cache(types.SimpleNamespace(out_degrees=lambda: out_deg), names)
            assert srv.capability == cap and srv.full_cached == (tag == "full")
            srv.log = True
            nf = _NodeFlow(layers)
            srv.fetch_data(nf)
            out[tag + "_gpu_flag"] = srv.gpu_flag.numpy().copy()
            out[tag + "_localid2cacheid"] = srv.localid2cacheid.numpy().copy()
            out[tag + "_cached_num"] = np.int64(srv.cached_num)
            for name in names:
                out["%s_cache_%s" % (tag, name)] = srv.gpu_fix_cache[name].numpy().copy()
                for i in range(nf.num_layers):
                    out["%s_frame%d_%s" % (tag, i, name)] = nf._node_frames[i][name].numpy().copy()
            if tag == "part":
                out["part_try_num"] = np.int64(srv.try_num)
                out["part_miss_num"] = np.int64(srv.miss_num)
                out["part_miss_rate"] = np.float64(srv.get_miss_rate())
        np.savez_compressed(os.path.join(GOLD, "cache_server.npz"), **out)
        return out
    finally:
        torch.Tensor.cuda = orig_cuda


# ------------------------------------------------------------------------ sageConv --
class _Repair(ast.NodeTransformer):
    """The two documented one-line repairs of sageconv.py (module docstring)."""

    def __init__(self):
        self.assigned, self.reset = 0, 0
        self._cls = None

    def visit_ClassDef(self, node):
        self._cls = node.name
        self.generic_visit(node)
        self._cls = None
        return node

    def visit_FunctionDef(self, node):
        self.generic_visit(node)
        if self._cls == "sageConv" and node.name == "__init__":
            node.body.append(ast.Expr(ast.Call(ast.Attribute(ast.Name("self", ast.Load()), "reset_parameters", ast.Load()), [], [])))
            self.reset += 1
        return node

    def visit_Expr(self, node):
        v = node.value
        if (self._cls == "NeighborAggregator" and isinstance(v, ast.Call) and isinstance(v.func, ast.Attribute)
                and isinstance(v.func.value, ast.Name) and v.func.value.id == "neighbor_feature"
                and v.func.attr in ("mean", "sum", "max")):
            val = ast.Attribute(v, "values", ast.Load()) if v.func.attr == "max" else v
            self.assigned += 1
            return ast.Assign([ast.Name("neighbor_feature", ast.Store())], val)
        return node


def run_sageconv(ref):
    path = os.path.join(ref, "dgll", "nn", "Convolution", "sageconv.py")
    tree = ast.parse(open(path).read(), filename=path)
    fix = _Repair()
    tree = fix.visit(tree)
    assert fix.assigned == 3 and fix.reset == 1, (fix.assigned, fix.reset)
    ast.fix_missing_locations(tree)
    body = [n for n in tree.body if isinstance(n, ast.ClassDef)]
    ns = {"F": make_backend()}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    sageConv, GraphSage = ns["sageConv"], ns["GraphSage"]
    out = {}
    rng = np.random.default_rng(11)
    B, Kn, Fi, H = 37, 6, 19, 8
    src = torch.from_numpy(rng.standard_normal((B, Fi)).astype(np.float32))
    neigh = torch.from_numpy(rng.standard_normal((B, Kn, Fi)).astype(np.float32))
    out["src"], out["neigh"] = src.numpy(), neigh.numpy()
    for aggr in ("mean", "sum", "max"):
        for comb in ("sum", "concat"):
            torch.manual_seed(5)
            layer = sageConv(Fi, H, aggr_neighbor_method=aggr, aggr_hid_method=comb)
            s = src.clone().requires_grad_(True)
            nb = neigh.clone().requires_grad_(True)
            y = layer(s, nb)
            g = torch.from_numpy(rng.standard_normal(tuple(y.shape)).astype(np.float32))
            y.backward(g)
            k = "%s_%s_" % (aggr, comb)
            out[k + "w_self"] = layer.weight.detach().numpy().copy()
            out[k + "w_neigh"] = layer.neighborAgg.weight.detach().numpy().copy()
            out[k + "out"] = y.detach().numpy().copy()
            out[k + "g"] = g.numpy()
            out[k + "d_src"] = s.grad.numpy().copy()
            out[k + "d_neigh"] = nb.grad.numpy().copy()
            out[k + "d_w_self"] = layer.weight.grad.numpy().copy()
            out[k + "d_w_neigh"] = layer.neighborAgg.weight.grad.numpy().copy()
    # the two-layer model on three hops of fixed-fanout features (sageconv.py:103-114)
    torch.manual_seed(9)
    fan = [4, 4]   # layer l views EVERY hop with num_neighbors_list[l] (:111), so the hops of one layer share a fanout
    model = GraphSage(Fi, hidden_dim=[12, 5], num_neighbors_list=fan)
    hops = [torch.from_numpy(rng.standard_normal((n, Fi)).astype(np.float32)) for n in (10, 10 * 4, 10 * 4 * 4)]
    y = model(hops)
    out["model_fan"] = np.asarray(fan)
    for i, h in enumerate(hops):
        out["model_hop%d" % i] = h.numpy()
    for i, g in enumerate(model.gcn):
        out["model_l%d_w_self" % i] = g.weight.detach().numpy().copy()
        out["model_l%d_w_neigh" % i] = g.neighborAgg.weight.detach().numpy().copy()
    out["model_out"] = y.detach().numpy().copy()
    np.savez_compressed(os.path.join(GOLD, "nn_sageconv_fixed.npz"), **out)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    a = run_cache_server(args.ref)
    b = run_sageconv(args.ref)
    print("cache_server.npz: %d arrays, miss rate %.4f; nn_sageconv_fixed.npz: %d arrays" %
          (len(a), float(a["part_miss_rate"]), len(b)))


if __name__ == "__main__":
    main()
