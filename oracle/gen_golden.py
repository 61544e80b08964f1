"""Generate tests/golden/*.npz by running the REFERENCE's own Python modules.

Run in the build container only (needs /root/reference mounted, read-only):

    python oracle/gen_golden.py [--ref /root/reference] [--ppi-scratch /tmp/ppi]

Nothing is copied from the reference: its modules are imported in place through
a ``sys.modules`` shim (the package is not importable as shipped — SURVEY.md §8c:
``nn/__inti__.py`` typo, absolute ``from gatconv import *``, ``dgll.backend`` =
bare torch which has no ``Parameter`` / ``dropout(training=)``).  The shim only
supplies a working ``dgll.backend``; every line of layer arithmetic executed is
the reference's.  Outputs are small seeded input/output vectors that pin
``oracle/layers.py`` / ``oracle/samplers.py`` / ``oracle/oracle.c``.

TEST INFRASTRUCTURE ONLY.
"""
import argparse
import importlib.util
import json
import os
import random
import sys
import tarfile
import types

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def make_backend():
    """dgll.backend as the layers need it: names resolved through torch.nn.functional -> torch.nn -> torch."""
    import torch.nn as nn
    import torch.nn.functional as Fn

    class _Backend(types.ModuleType):
        def __getattr__(self, name):
            for mod in (Fn, nn, torch):
                if hasattr(mod, name):
                    return getattr(mod, name)
            raise AttributeError(name)

    b = _Backend("dgll.backend")
    b.nn = nn
    b.init = nn.init
    b.Tensor = torch.Tensor
    b.FloatTensor = torch.FloatTensor
    b.LongTensor = torch.LongTensor
    b.autograd = torch.autograd
    b.sparse = torch.sparse
    b.optim = torch.optim
    return b


def install_shim(ref):
    pkg = types.ModuleType("dgll")
    pkg.__path__ = [os.path.join(ref, "dgll")]
    pkg.backend = make_backend()
    sys.modules["dgll"] = pkg
    sys.modules["dgll.backend"] = pkg.backend
    return pkg


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def t2n(t):
    return t.detach().cpu().numpy()


def save(name, **arrays):
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024))


def random_graph(n, avg_deg, seed, symmetric=True, self_loops=False):
    rng = np.random.RandomState(seed)
    m = n * avg_deg // (2 if symmetric else 1)
    src = rng.randint(0, n, size=m)
    dst = (src + 1 + rng.zipf(1.6, size=m) % (n - 1)) % n  # skewed offsets, never a self loop
    a = np.zeros((n, n), dtype=np.float32)
    a[dst, src] = 1.0
    if symmetric:
        a[src, dst] = 1.0
    if self_loops:
        a[np.arange(n), np.arange(n)] = 1.0
    a[n - 1, :] = 0.0  # one isolated destination row (edge case) unless self loops
    if symmetric:
        a[:, n - 1] = 0.0
    if self_loops:
        a[n - 1, n - 1] = 1.0
    return a


# ------------------------------------------------------------------ PPI (config C1) --
def gen_ppi(ref, scratch):
    ppi_dir = os.path.join(scratch, "PPI")
    if not os.path.exists(os.path.join(ppi_dir, "train_graph.json")):
        os.makedirs(scratch, exist_ok=True)
        with tarfile.open(os.path.join(ref, "Evaluation", "PPI.tar.xz")) as tf:
            tf.extractall(scratch)
    # networkx >= 3.4 renamed the node-link key; the reference calls node_link_graph(data) (ppi_dataloader.py:24)
    from networkx.readwrite import json_graph
    orig = json_graph.node_link_graph
    json_graph.node_link_graph = lambda data, *a, **k: orig(data, *a, edges="links", **k)
    sys.path.insert(0, os.path.join(ref, "Evaluation", "PPI"))
    loader = load_by_path("ref_ppi_dataloader", os.path.join(ref, "Evaluation", "PPI", "ppi_dataloader.py"))
    model_mod = load_by_path("ref_gcn_model", os.path.join(ref, "Evaluation", "PPI", "gcn_model.py"))
    train = loader.load_ppi_dataset(ppi_dir, "train")
    sizes = [(g[1].shape[0], g[0].shape[1]) for g in train]
    print("PPI train graphs (N, nnz):", sizes)
    for gi in (8, 5):  # the two smallest graphs: (591, 7708) and (1021, 18216)
        edge_index, feats, labels = train[gi]
        torch.manual_seed(0)
        model = model_mod.GCN(feats.shape[1], 64, labels.shape[1], 2)  # "2-layer" = num_layers=2
        crit = torch.nn.CrossEntropyLoss()  # as train_gcn.py:26 — float multi-hot targets
        out = model(edge_index, feats)
        loss = crit(out, labels)
        loss.backward()
        sd = {k: t2n(v) for k, v in model.state_dict().items()}
        grads = {k: t2n(p.grad) for k, p in model.named_parameters()}
        # one layer in isolation: relu(A (X W)) — the aggregation under test
        h1 = model.layers[0](edge_index, feats, feats.size(0))
        save("ppi_gcn_g%d" % gi,
             edge_index=t2n(edge_index).astype(np.int32), feats=t2n(feats), labels=t2n(labels).astype(np.uint8),
             w0=sd["layers.0.weight"], w1=sd["layers.1.weight"], w_out=sd["out_layer.weight"],
             b_out=sd["out_layer.bias"], logits=t2n(out), loss=np.float64(loss.item()), h1=t2n(h1),
             g_w0=grads["layers.0.weight"], g_w1=grads["layers.1.weight"], g_w_out=grads["out_layer.weight"],
             g_b_out=grads["out_layer.bias"], all_sizes=np.array(sizes, dtype=np.int64))


# ------------------------------------------------------- dgll.nn layers (GCN/GAT/SpGAT/GIN) --
def gen_nn(ref):
    conv = os.path.join(ref, "dgll", "nn", "Convolution")
    gcnconv = load_by_path("ref_gcnconv", os.path.join(conv, "gcnconv.py"))
    gatconv = load_by_path("ref_gatconv", os.path.join(conv, "gatconv.py"))
    ginconv = load_by_path("ref_ginconv", os.path.join(conv, "ginconv.py"))
    utils = load_by_path("ref_nn_utils", os.path.join(ref, "dgll", "nn", "utils", "utils.py"))
    import scipy.sparse as sp

    # --- GCN: adjacency prepared by the reference's own helpers (utils.py:168-171,240-257)
    n, f, nhid, ncls = 200, 32, 16, 7
    a = random_graph(n, 8, seed=1, symmetric=False)
    adj = sp.coo_matrix(a)
    adj = adj + adj.T.multiply(adj.T > adj) - adj.multiply(adj.T > adj)   # utils.py:168
    adj_n = utils.normalize(adj + sp.eye(adj.shape[0]))                    # utils.py:171
    adj_t = utils.sparse_mx_to_torch_sparse_tensor(adj_n)                  # utils.py:179
    torch.manual_seed(0)
    x = torch.randn(n, f)
    model = gcnconv.GCN(f, nhid, ncls, dropout=0.5)
    model.eval()
    x.requires_grad_(True)
    out = model(x, adj_t)
    labels = torch.randint(0, ncls, (n,), generator=torch.Generator().manual_seed(3))
    loss = torch.nn.functional.nll_loss(out, labels)
    loss.backward()
    coo = adj_t.coalesce()
    layer_out = model.gcn1(x, adj_t)
    save("nn_gcn", x=t2n(x), adj_indices=t2n(coo.indices()).astype(np.int32), adj_values=t2n(coo.values()),
         adj_raw=a, w1=t2n(model.gcn1.weight), b1=t2n(model.gcn1.bias), w2=t2n(model.gcn2.weight),
         b2=t2n(model.gcn2.bias), out=t2n(out), layer_out=t2n(layer_out), labels=t2n(labels),
         loss=np.float64(loss.item()), g_x=t2n(x.grad), g_w1=t2n(model.gcn1.weight.grad),
         g_b1=t2n(model.gcn1.bias.grad), g_w2=t2n(model.gcn2.weight.grad), g_b2=t2n(model.gcn2.bias.grad))

    # --- GAT dense + sparse: same graph (with self loops so no row is empty), dropout 0, eval
    n, f, nhid, ncls, heads, alpha = 150, 24, 8, 5, 4, 0.2
    a = random_graph(n, 6, seed=2, symmetric=True, self_loops=True)
    adj_d = torch.from_numpy(a)
    labels = torch.randint(0, ncls, (n,), generator=torch.Generator().manual_seed(4))
    for cls_name, tag in (("GAT", "gat_dense"), ("SpGAT", "gat_sparse")):
        torch.manual_seed(0)
        x = torch.randn(n, f).requires_grad_(True)
        model = getattr(gatconv, cls_name)(f, nhid, ncls, dropout=0.0, alpha=alpha, nheads=heads)
        model.eval()
        out = model(x, adj_d)
        loss = torch.nn.functional.nll_loss(out, labels)
        loss.backward()
        first = torch.cat([att(x, adj_d) for att in model.attentions], dim=1)
        arrays = dict(x=t2n(x), adj=a, out=t2n(out), first_layer=t2n(first), labels=t2n(labels),
                      loss=np.float64(loss.item()), g_x=t2n(x.grad), alpha=np.float32(alpha))
        for i, att in enumerate(model.attentions):
            arrays["W%d" % i], arrays["a%d" % i] = t2n(att.W), t2n(att.a)
            arrays["g_W%d" % i], arrays["g_a%d" % i] = t2n(att.W.grad), t2n(att.a.grad)
        arrays["W_out"], arrays["a_out"] = t2n(model.out_att.W), t2n(model.out_att.a)
        arrays["g_W_out"], arrays["g_a_out"] = t2n(model.out_att.W.grad), t2n(model.out_att.a.grad)
        save("nn_" + tag, **arrays)

    # --- SpecialSpmmFunction forward/backward (gatconv.py:60-81)
    torch.manual_seed(1)
    n, d = 60, 12
    a = random_graph(n, 5, seed=5, symmetric=False)
    idx = torch.from_numpy(a).nonzero().t()
    vals = torch.rand(idx.shape[1]).requires_grad_(True)
    b = torch.randn(n, d).requires_grad_(True)
    y = gatconv.SpecialSpmmFunction.apply(idx, vals, torch.Size([n, n]), b)
    g = torch.randn(n, d)
    y.backward(g)
    save("nn_special_spmm", indices=t2n(idx).astype(np.int32), values=t2n(vals), b=t2n(b), y=t2n(y), g=t2n(g),
         g_values=t2n(vals.grad), g_b=t2n(b.grad))

    # --- GIN (ginconv.py:10-66)
    torch.manual_seed(2)
    B, n, f, hid, od = 3, 20, 6, 10, 4
    A = torch.from_numpy(np.stack([random_graph(n, 4, seed=10 + i) for i in range(B)]))
    X = torch.randn(B, n, f)
    gin = ginconv.GIN(f, hid, od, 2)
    out = gin(A, X)
    sd = {k.replace(".", "__"): t2n(v) for k, v in gin.state_dict().items()}
    save("nn_gin", A=t2n(A), X=t2n(X), out=t2n(out), **sd)

    # --- normalisation helpers (utils.py:240-257) on a raw matrix
    m = sp.coo_matrix(random_graph(40, 4, seed=6, symmetric=False))
    mn = utils.normalize(m)
    t = utils.sparse_mx_to_torch_sparse_tensor(mn).coalesce()
    save("nn_normalize", raw=m.toarray().astype(np.float32), indices=t2n(t.indices()).astype(np.int32),
         values=t2n(t.values()), dense=t2n(t.to_dense()))

    # --- fixed-fanout sampler with replacement (utils.py:52-68), legacy numpy RNG
    rng = np.random.RandomState(11)
    n = 120
    nbr_tab = {v: list(rng.choice(n, size=rng.randint(1, 9), replace=False)) for v in range(n)}
    np.random.seed(7)
    hops = utils.multihop_sampling(np.arange(0, 16), [5, 3], nbr_tab)
    flat = np.concatenate([np.array(nbr_tab[v], dtype=np.int64) for v in range(n)])
    ptr = np.cumsum([0] + [len(nbr_tab[v]) for v in range(n)])
    save("sampler_multihop", nbr_flat=flat, nbr_ptr=np.array(ptr, dtype=np.int64), seeds=np.arange(0, 16),
         hop1=np.asarray(hops[1], dtype=np.int64), hop2=np.asarray(hops[2], dtype=np.int64),
         fanouts=np.array([5, 3]), np_seed=np.int64(7))


# ------------------------------------------------------- in-tree graph store + sampler --
def gen_sampler(ref):
    dgraph = load_by_path("dgll.data.dgraph", os.path.join(ref, "dgll", "data", "dgraph.py"))
    sys.modules.setdefault("dgll.sampling", types.ModuleType("dgll.sampling"))
    sys.modules["dgll.sampling"].__path__ = [os.path.join(ref, "dgll", "sampling")]
    base = load_by_path("dgll.sampling.base_sampler", os.path.join(ref, "dgll", "sampling", "base_sampler.py"))
    smp = load_by_path("dgll.sampling.dgllsampler", os.path.join(ref, "dgll", "sampling", "dgllsampler.py"))
    rng = np.random.RandomState(21)
    n, f = 300, 10
    edges = []
    for v in range(n):
        deg = int(rng.choice([0, 1, 2, 3, 5, 8, 13, 21, 40], p=[.05, .1, .15, .2, .2, .15, .08, .05, .02]))
        edges.append(sorted(rng.choice(n, size=min(deg, n), replace=False).tolist()))
    feats = torch.from_numpy(rng.randn(n, f).astype(np.float32))
    labels = torch.from_numpy(rng.randint(0, 7, size=n))
    g = dgraph.DGraph(nodes=torch.arange(n), edges=edges, labels=labels, features=feats,
                      train_mask=torch.arange(n) < 200, test_mask=torch.arange(n) >= 250,
                      validation_mask=(torch.arange(n) >= 200) & (torch.arange(n) < 250))
    seeds = torch.tensor([0, 2, 5, 6, 9, 23, 77, 150, 299, 5])  # includes a duplicate seed
    random.seed(1234)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):  # the reference prints the graph object (dgllsampler.py:13)
        input_nodes, output_nodes, subgs = smp.DGLLNeighborSampler([5, 3]).sample(g, seeds)
        adj = smp.DGLLNeighborSampler([5, 3]).get_adj(g, subgs)
        gathered = subgs[0].get_features(g, subgs)
    induced = g.get_induced_subgraph(torch.tensor([0, 2, 5, 6, 9, 23]))
    flat = np.concatenate([np.array(e, dtype=np.int64) for e in edges])
    ptr = np.cumsum([0] + [len(e) for e in edges])
    save("sampler_neighbor", nbr_flat=flat, nbr_ptr=np.array(ptr, dtype=np.int64), feats=t2n(feats),
         labels=t2n(labels), seeds=t2n(seeds), py_seed=np.int64(1234), fanouts=np.array([5, 3]),
         input_nodes=t2n(input_nodes), output_nodes=t2n(output_nodes),
         b0_src=t2n(subgs[0].src_nodes()), b0_dst=t2n(subgs[0].dst_nodes()), b0_nodes=t2n(subgs[0].nodes()),
         b1_src=t2n(subgs[1].src_nodes()), b1_dst=t2n(subgs[1].dst_nodes()), b1_nodes=t2n(subgs[1].nodes()),
         adj=t2n(adj), gathered=t2n(gathered), induced=t2n(induced),
         neighbors_json=np.frombuffer(json.dumps(g.get_neighbors(torch.tensor([0, 2, 5]))).encode(), dtype=np.uint8),
         train_nodes=t2n(g.get_train_nodes()), labels_sel=t2n(g.get_labels(seeds)),
         feats_sel=t2n(g.get_features(seeds)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--ppi-scratch", default="/tmp/ppi")
    args = ap.parse_args()
    install_shim(args.ref)
    torch.set_num_threads(1)  # deterministic CPU reductions
    gen_ppi(args.ref, args.ppi_scratch)
    gen_nn(args.ref)
    gen_sampler(args.ref)


if __name__ == "__main__":
    main()
