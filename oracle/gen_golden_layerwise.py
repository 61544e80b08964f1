"""Generate tests/golden/layerwise_*.npz by executing the REFERENCE's own layer-wise sampler code.

Run in the build container only (needs /root/reference mounted, read-only):

    python oracle/gen_golden_layerwise.py [--ref /root/reference]

The sampler scripts under ``dgll/GPU Accelerator`` are not importable (module-level dataset downloads, DGL under the
alias ``dgll``, ogb).  This script parses each file with ``ast`` and executes ONLY the definitions under test, in place,
with no source copied into this repository:

  utils.py            matrix_row_normalize :11-19, estWRS_weights :199-213, normalize_lap :215-222
  MQLadies.py         class Ladies :62-89                 (flat=False / flat=True)
  MQLadiesFlatWrs.py  class LadiesFlatWrs :63-90
  MQFastGCN.py        class FastGCNSampler :60-88         (the np.unique(concat(batch)) variant)
  MQFastGCNFlatWrs.py class FastGCNSamplerFlatWrs :63-101 (flat / wrs switches)

DGL is replaced by the smallest possible stand-in: ``dgll.dataloading.Sampler`` = object,
``dgll.create_block(('csc', (indptr, indices, [])))`` records the arrays and answers ``srcnodes()`` with
``arange(num_src)`` the way a DGL block does, and the graph object answers ``adj_external(scipy_fmt='csr')``,
``num_nodes()`` and ``ndata``.  Every line of sampling arithmetic executed is the reference's.  The edge values of the
importance-weighted adjacency (``adj.data``) are computed by the reference and then dropped at ``create_block``; the
fixtures hold the picks and the WRS weights, from which tests rebuild them.

TEST INFRASTRUCTURE ONLY.
"""
import argparse
import ast
import os
import types

import numpy as np
import scipy.sparse as sp
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def defs_from(path, names, namespace):
    """Execute the named top-level FunctionDef / ClassDef nodes of ``path`` inside ``namespace``."""
    with open(path) as f:
        tree = ast.parse(f.read(), filename=path)
    picked = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    assert sorted(n.name for n in picked) == sorted(names), (path, [n.name for n in picked])
    mod = ast.Module(body=picked, type_ignores=[])
    exec(compile(mod, path, "exec"), namespace)


class _Block:
    def __init__(self, data):
        fmt, (indptr, indices, _e) = data
        assert fmt == "csc"
        self.indptr = np.asarray(indptr).copy()
        self.indices = np.asarray(indices).copy()
        self.num_dst = len(self.indptr) - 1
        self.num_src = int(self.indices.max()) + 1 if len(self.indices) else 0
        self.srcdata, self.dstdata = {}, {}

    def srcnodes(self):
        # a DGL block answers with a torch arange; this scipy rejects torch index arrays, so hand back the same
        # values as an ndarray that also answers .clone().detach() (MQFastGCN.py:88)
        return _Ids(np.arange(self.num_src))


class _Ids(np.ndarray):
    def __new__(cls, a):
        return np.asarray(a).view(cls)

    def clone(self):
        return self

    def detach(self):
        return np.asarray(self)


class _Graph:
    def __init__(self, adj, feat, label):
        self._adj = adj
        self.ndata = {"feat": feat, "label": label}

    def adj_external(self, scipy_fmt="csr"):
        return self._adj.copy()

    def num_nodes(self):
        return self._adj.shape[0]


def make_namespace(ref):
    created = []
    dgll = types.SimpleNamespace(
        dataloading=types.SimpleNamespace(Sampler=object),
        create_block=lambda data: created.append(_Block(data)) or created[-1])
    ns = {"np": np, "sp": sp, "torch": torch, "dgll": dgll}
    gpu_acc = os.path.join(ref, "dgll", "GPU Accelerator")
    defs_from(os.path.join(gpu_acc, "utils.py"), ["matrix_row_normalize", "estWRS_weights", "normalize_lap"], ns)
    return ns, gpu_acc, created


def random_digraph(n, avg_deg, seed, symmetric):
    rng = np.random.RandomState(seed)
    m = n * avg_deg
    src = rng.randint(0, n, size=m)
    dst = (src + 1 + rng.zipf(1.5, size=m) % (n - 1)) % n   # skewed, never a self loop
    a = sp.coo_matrix((np.ones(m), (src, dst)), shape=(n, n)).tocsr()
    a.data[:] = 1.0                                           # duplicates collapse to 1
    if symmetric:
        a = ((a + a.T) > 0).astype(np.float64).tocsr()
    a.sort_indices()
    return a


def run_case(ns, created, cls_name, kwargs, adj, batch, fanouts, seed):
    """Run the reference sampler class; capture per layer the recorded block arrays, the picks and the weights."""
    n = adj.shape[0]
    g = _Graph(adj, torch.arange(n * 3, dtype=torch.float32).view(n, 3), torch.arange(n) % 7)
    sampler = ns[cls_name](list(fanouts), g, **kwargs)
    # capture picks / weights: wrap np.random.choice and the estimator without changing what they compute
    picks, wts = [], []
    real_choice = np.random.choice
    real_est = ns["estWRS_weights"]

    def choice(*a, **k):
        r = real_choice(*a, **k)
        picks.append(np.asarray(r).copy())
        return r

    def est(p, m):
        idx, w = real_est(p, m)
        wts.append(np.asarray(w).copy())
        return idx, w

    np.random.seed(seed)
    del created[:]
    rnd = types.SimpleNamespace(choice=choice)
    # the definitions were exec'd in `ns`, so that dict is their module globals: observe np.random.choice and the estimator
    ns["np"], ns["estWRS_weights"] = _NumpyProxy(rnd), est
    try:
        input_nodes, out_nodes, subgs = sampler.sample(g, batch)
    finally:
        ns["np"], ns["estWRS_weights"] = np, real_est
    lap = sampler.lap_matrix.tocsr().copy()
    lap.sort_indices()      # scipy's diag·csr product leaves rows unsorted; canonical order, same values
    blocks = list(created)  # creation order = output layer first (the reference reverses subgs afterwards)
    out = {"lap_indptr": lap.indptr.astype(np.int64), "lap_indices": lap.indices.astype(np.int64),
           "lap_data": lap.data.astype(np.float64),
           "adj_indptr": adj.indptr.astype(np.int64), "adj_indices": adj.indices.astype(np.int64),
           "batch": np.asarray(batch).astype(np.int64), "fanouts": np.asarray(fanouts, dtype=np.int64),
           "np_seed": np.int64(seed), "input_nodes": np.asarray(input_nodes).astype(np.int64),
           "n_layers": np.int64(len(blocks)),
           "feat_rows": subgs[0].srcdata["feat"].numpy(), "labels": subgs[-1].dstdata["label"].numpy()}
    for li, b in enumerate(blocks):
        out["l%d_indptr" % li] = b.indptr.astype(np.int64)
        out["l%d_indices" % li] = b.indices.astype(np.int64)
        out["l%d_picks" % li] = picks[li].astype(np.int64)
        if wts:
            out["l%d_weights" % li] = wts[li].astype(np.float64)
    return out


class _NumpyProxy:
    """numpy with ``random.choice`` observed (results recorded); everything else is numpy itself."""

    def __init__(self, rnd):
        self.random = rnd

    def __getattr__(self, name):
        return getattr(np, name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    ns, gpu_acc, created = make_namespace(args.ref)
    defs_from(os.path.join(gpu_acc, "MQLadies.py"), ["Ladies"], ns)
    defs_from(os.path.join(gpu_acc, "MQLadiesFlatWrs.py"), ["LadiesFlatWrs"], ns)
    defs_from(os.path.join(gpu_acc, "MQFastGCN.py"), ["FastGCNSampler"], ns)
    defs_from(os.path.join(gpu_acc, "MQFastGCNFlatWrs.py"), ["FastGCNSamplerFlatWrs"], ns)
    os.makedirs(GOLD, exist_ok=True)
    sym = random_digraph(400, 6, 3, symmetric=True)
    dig = random_digraph(350, 5, 4, symmetric=False)
    batch_s = np.random.RandomState(9).choice(400, 48, replace=False)
    batch_d = np.random.RandomState(10).choice(350, 40, replace=False)
    cases = [
        ("ladies_sym", "Ladies", {}, sym, batch_s, [64, 96]),
        ("ladies_flat_dir", "Ladies", {"flat": True}, dig, batch_d, [50, 500]),   # 500 > candidates: s_num clamps
        ("ladiesflatwrs_sym", "LadiesFlatWrs", {"flat": True}, sym, batch_s, [64, 96]),
        ("fastgcn_sym", "FastGCNSampler", {}, sym, batch_s, [64, 96]),
        ("fastgcnflatwrs_plain_dir", "FastGCNSamplerFlatWrs", {}, dig, batch_d, [50, 80]),
        ("fastgcnflatwrs_flat_sym", "FastGCNSamplerFlatWrs", {"flat": True}, sym, batch_s, [64, 96]),
        ("fastgcnflatwrs_wrs_sym", "FastGCNSamplerFlatWrs", {"flat": True, "wrs": True}, sym, batch_s, [64, 96]),
    ]
    for name, cls, kw, adj, batch, fan in cases:
        out = run_case(ns, created, cls, kw, adj, batch, fan, seed=1234)
        path = os.path.join(GOLD, "layerwise_%s.npz" % name)
        np.savez_compressed(path, **out)
        print("wrote %s (%.1f KB): layers %s" % (path, os.path.getsize(path) / 1024,
                                                [(len(out["l%d_indptr" % i]) - 1, len(out["l%d_indices" % i]))
                                                 for i in range(int(out["n_layers"]))]))
    # the weight estimator on its own
    np.random.seed(77)
    p = np.random.rand(200)
    p[::7] = 0
    p /= p.sum()
    np.random.seed(78)
    idx, w = ns["estWRS_weights"](p, 60)
    np.savez_compressed(os.path.join(GOLD, "layerwise_estwrs.npz"), p=p, idx=idx.astype(np.int64), w=w, m=np.int64(60),
                        np_seed=np.int64(78))
    print("wrote layerwise_estwrs.npz")


if __name__ == "__main__":
    main()
