"""GPU parity of the reference-facing layer / data-path API (dgll_b200.nn, dgll_b200.data, dgll_b200.gcn_extension,
dgll_b200.backend) against golden vectors generated from the reference's own modules (oracle/gen_golden.py) and the
oracle restatements.  fp32: <= 1e-5 relative on outputs/losses; gradients <= 1e-4 (the golden gradients are fp32
autograd of the reference, themselves ~1e-6 noisy)."""
import json
import random

import numpy as np
import pytest
import torch

import oracle
from oracle import layers as L
from oracle import samplers as S
from conftest import golden, rel_err

pytestmark = pytest.mark.gpu
OUT_TOL, GRAD_TOL = 1e-5, 1e-4


def cu(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a)).to(dtype).cuda()


@pytest.fixture(scope="module")
def nn():
    assert torch.cuda.is_available()
    import dgll_b200.nn as m
    return m


# ---------------------------------------------------------------- GCN ------
def test_gcn_model_matches_reference_golden(nn):
    g = golden("nn_gcn")
    n = g["x"].shape[0]
    adj = torch.sparse_coo_tensor(cu(g["adj_indices"], torch.int64), cu(g["adj_values"]), (n, n))
    model = nn.GCN(32, 16, 7, dropout=0.5).cuda().eval()
    with torch.no_grad():
        model.gcn1.weight.copy_(cu(g["w1"])); model.gcn1.bias.copy_(cu(g["b1"]))
        model.gcn2.weight.copy_(cu(g["w2"])); model.gcn2.bias.copy_(cu(g["b2"]))
    x = cu(g["x"]).requires_grad_(True)
    out = model(x, adj)
    loss = torch.nn.functional.nll_loss(out, cu(g["labels"], torch.int64))
    loss.backward()
    assert rel_err(out.detach().cpu(), g["out"]) <= OUT_TOL
    assert abs(loss.item() - float(g["loss"])) <= OUT_TOL * abs(float(g["loss"]))
    assert rel_err(x.grad.cpu(), g["g_x"]) <= GRAD_TOL
    assert rel_err(model.gcn1.weight.grad.cpu(), g["g_w1"]) <= GRAD_TOL
    assert rel_err(model.gcn1.bias.grad.cpu(), g["g_b1"]) <= GRAD_TOL
    assert rel_err(model.gcn2.weight.grad.cpu(), g["g_w2"]) <= GRAD_TOL
    assert rel_err(model.gcn2.bias.grad.cpu(), g["g_b2"]) <= GRAD_TOL
    lo = model.gcn1(cu(g["x"]), adj)
    assert rel_err(lo.detach().cpu(), g["layer_out"]) <= OUT_TOL
    assert sorted(model.state_dict()) == ["gcn1.bias", "gcn1.weight", "gcn2.bias", "gcn2.weight"]


def test_backend_routes_reference_style_code(nn):
    """Code written against ``from dgll import backend as F`` (gcnconv.py:29-35) runs on the kernels."""
    from dgll_b200 import _lib, backend as F
    g = golden("nn_gcn")
    n = g["x"].shape[0]
    adj = F.sparse_coo_tensor(cu(g["adj_indices"], torch.int64), cu(g["adj_values"]), F.Size([n, n]))
    w = F.Parameter(cu(g["w1"]))
    before = _lib.launch_count()
    support = F.mm(cu(g["x"]), w)
    output = F.spmm(adj, support) + cu(g["b1"])
    assert _lib.launch_count() >= before + 2
    assert rel_err(output.detach().cpu(), g["layer_out"]) <= OUT_TOL
    out2 = F.sparse.mm(adj, support)
    assert torch.equal(out2, F.spmm(adj, support))
    with pytest.raises(RuntimeError):
        F.spmm(adj.cpu(), support.cpu())
    assert F.dropout(support, 0.5, training=False) is not None and hasattr(F, "FloatTensor") and hasattr(F, "LeakyReLU")


# ----------------------------------------------------------------- PPI -----
@pytest.mark.parametrize("gi", [8, 5])
def test_ppi_gcn_config1_forward_loss_grads(nn, gi):
    """C1: Evaluation/PPI GCN (2 GCNLayers + Linear) reproduced with ops.linear + ops.spmm(relu) — golden from the
    reference's gcn_model.py on the bundled PPI graphs.  Activations reach ~1e6: relative to max|ref|."""
    from dgll_b200 import ops
    g = golden("ppi_gcn_g%d" % gi)
    n = g["feats"].shape[0]
    graph = ops.CsrGraph.from_edge_index(cu(g["edge_index"], torch.int64), n)
    x = cu(g["feats"])
    w0 = cu(g["w0"]).requires_grad_(True)
    w1 = cu(g["w1"]).requires_grad_(True)
    w_out = cu(g["w_out"]).requires_grad_(True)
    b_out = cu(g["b_out"]).requires_grad_(True)
    h1 = ops.spmm(graph, ops.linear(x, w0), relu=True)
    h2 = ops.spmm(graph, ops.linear(h1, w1), relu=True)
    logits = ops.linear(h2, w_out.t(), bias=b_out)
    loss = torch.nn.functional.cross_entropy(logits, cu(g["labels"]))
    loss.backward()
    assert rel_err(h1.detach().cpu(), g["h1"]) <= OUT_TOL
    assert rel_err(logits.detach().cpu(), g["logits"]) <= OUT_TOL
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert rel_err(w0.grad.cpu(), g["g_w0"]) <= GRAD_TOL
    assert rel_err(w1.grad.cpu(), g["g_w1"]) <= GRAD_TOL
    assert rel_err(w_out.grad.cpu(), g["g_w_out"]) <= GRAD_TOL
    assert rel_err(b_out.grad.cpu(), g["g_b_out"]) <= GRAD_TOL


def test_ppi_model_class_drop_in(nn):
    """The Evaluation/PPI GCN class itself (same state-dict keys) on the golden graph: logits, loss, one Adam step."""
    g = golden("ppi_gcn_g8")
    model = nn.PPIGCN(50, 64, 121, 2).cuda()
    assert sorted(model.state_dict()) == ["layers.0.weight", "layers.1.weight", "out_layer.bias", "out_layer.weight"]
    model.load_state_dict({"layers.0.weight": cu(g["w0"]), "layers.1.weight": cu(g["w1"]),
                           "out_layer.weight": cu(g["w_out"]), "out_layer.bias": cu(g["b_out"])})
    ei, x, y = cu(g["edge_index"], torch.int64), cu(g["feats"]), cu(g["labels"])
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    out = model(ei, x)
    loss = torch.nn.CrossEntropyLoss()(out, y)
    assert rel_err(out.detach().cpu(), g["logits"]) <= OUT_TOL
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    opt.zero_grad(); loss.backward(); opt.step()
    assert rel_err(model.layers[0].weight.grad.cpu(), g["g_w0"]) <= GRAD_TOL
    out2 = model(ei, x)                        # second call reuses the cached CSR of this edge_index tensor
    assert torch.isfinite(out2).all() and hasattr(ei, "_dgllb_csr")


# ----------------------------------------------------------------- GAT -----
def _load_gat(model, g, heads):
    with torch.no_grad():
        for i in range(heads):
            getattr(model, "attention_%d" % i).W.copy_(cu(g["W%d" % i]))
            getattr(model, "attention_%d" % i).a.copy_(cu(g["a%d" % i]))
        model.out_att.W.copy_(cu(g["W_out"]))
        model.out_att.a.copy_(cu(g["a_out"]))


@pytest.mark.parametrize("tag,cls", [("nn_gat_dense", "GAT"), ("nn_gat_sparse", "SpGAT")])
def test_gat_models_match_reference_golden(nn, tag, cls):
    g = golden(tag)
    heads = 4
    model = getattr(nn, cls)(24, 8, 5, dropout=0.0, alpha=float(g["alpha"]), nheads=heads).cuda()
    _load_gat(model, g, heads)
    adj = cu(g["adj"])
    x = cu(g["x"]).requires_grad_(True)
    out = model(x, adj)
    loss = torch.nn.functional.nll_loss(out, cu(g["labels"], torch.int64))
    loss.backward()
    assert rel_err(out.detach().cpu(), g["out"]) <= OUT_TOL
    assert abs(loss.item() - float(g["loss"])) <= OUT_TOL * abs(float(g["loss"]))
    assert rel_err(x.grad.cpu(), g["g_x"]) <= GRAD_TOL
    for i in range(heads):
        att = getattr(model, "attention_%d" % i)
        assert rel_err(att.W.grad.cpu(), g["g_W%d" % i]) <= GRAD_TOL
        assert rel_err(att.a.grad.cpu(), g["g_a%d" % i]) <= GRAD_TOL
    assert rel_err(model.out_att.W.grad.cpu(), g["g_W_out"]) <= GRAD_TOL
    assert rel_err(model.out_att.a.grad.cpu(), g["g_a_out"]) <= GRAD_TOL
    # single-layer call path (heads=1 kernel) equals the multi-head fused launch, head by head
    model.eval()
    with torch.no_grad():
        first = torch.cat([getattr(model, "attention_%d" % i)(cu(g["x"]), adj) for i in range(heads)], dim=1)
    assert rel_err(first.cpu(), g["first_layer"]) <= OUT_TOL


def test_attention_dropout_in_training_mode(nn):
    """gatconv.py:37 / :132: dropout on the attention coefficients in training mode.  The fused kernel draws its own
    counter-based mask: eval mode is deterministic and equals p=0; training mode differs call to call, follows
    torch.manual_seed, keeps the expectation, and backpropagates."""
    torch.manual_seed(0)
    n = 400
    adj = (torch.rand(n, n, device="cuda") < 0.05).float()
    adj.fill_diagonal_(1.0)
    x = torch.randn(n, 16, device="cuda")
    for cls in (nn.gatConv, nn.sparseGatConv):
        m = cls(16, 8, dropout=0.6, alpha=0.2).cuda()
        m.eval()
        e1, e2 = m(x, adj), m(x, adj)
        assert torch.equal(e1, e2)
        m.train()
        torch.manual_seed(1)
        t1 = m(x, adj)
        t2 = m(x, adj)
        torch.manual_seed(1)
        t3 = m(x, adj)
        assert not torch.equal(t1, t2) and torch.equal(t1, t3) and not torch.equal(t1, e1)
        t1.sum().backward()
        assert torch.isfinite(m.W.grad).all() and torch.isfinite(m.a.grad).all()
    model = nn.SpGAT(16, 8, 5, dropout=0.6, alpha=0.2, nheads=4).cuda().train()
    out = model(x, adj)
    torch.nn.functional.nll_loss(out, torch.randint(0, 5, (n,), device="cuda")).backward()
    assert all(torch.isfinite(p.grad).all() for p in model.parameters())


def test_special_spmm_matches_reference_golden(nn):
    g = golden("nn_special_spmm")
    n = g["b"].shape[0]
    idx = cu(g["indices"], torch.int64)
    vals = cu(g["values"]).requires_grad_(True)
    b = cu(g["b"]).requires_grad_(True)
    y = nn.SpecialSpmm()(idx, vals, torch.Size([n, n]), b)
    y.backward(cu(g["g"]))
    assert rel_err(y.detach().cpu(), g["y"]) <= OUT_TOL
    assert rel_err(vals.grad.cpu(), g["g_values"]) <= OUT_TOL
    assert rel_err(b.grad.cpu(), g["g_b"]) <= OUT_TOL


# ----------------------------------------------------------------- GIN / pooling
def test_gin_model_matches_reference_golden(nn):
    g = golden("nn_gin")
    model = nn.GIN(6, 10, 4, 2).cuda()
    sd = {k.replace("__", "."): cu(v) for k, v in g.items() if "__" in k}
    model.load_state_dict(sd)
    out = model(cu(g["A"]), cu(g["X"]))
    assert rel_err(out.detach().cpu(), g["out"]) <= OUT_TOL


@pytest.mark.parametrize("sorted_batch", [True, False])
def test_pooling_matches_scatter_semantics(nn, sorted_batch):
    rng = np.random.default_rng(0)
    sizes = rng.integers(0, 30, size=25)
    batch = np.repeat(np.arange(25), sizes)
    if not sorted_batch:
        batch = rng.permutation(batch)
    x = rng.standard_normal((batch.size, 19)).astype(np.float32)
    for red, fn in (("sum", nn.sumPooling), ("mean", nn.meanPooling), ("max", nn.maxPooling)):
        ref = L.pooling(torch.from_numpy(x).double(), torch.from_numpy(batch), size=25, reduce=red).numpy()
        out = fn(cu(x), cu(batch, torch.int64), 25)
        assert rel_err(out.cpu(), ref) <= OUT_TOL
    both = nn.Pooling(["sum", "max"])(cu(x), cu(batch, torch.int64), 25)
    assert both.shape == (25, 38)
    one = nn.sumPooling(cu(x), None)
    assert rel_err(one.cpu(), x.astype(np.float64).sum(0, keepdims=True)) <= OUT_TOL
    # gradient of mean pooling = 1/|segment| scattered back
    xs = cu(x).requires_grad_(True)
    nn.meanPooling(xs, cu(batch, torch.int64), 25).sum().backward()
    cnt = np.bincount(batch, minlength=25)
    assert rel_err(xs.grad.cpu(), np.repeat((1.0 / cnt[batch])[:, None], 19, axis=1)) <= OUT_TOL


def test_as_csr_cache_follows_in_place_edits_of_sparse_values():
    from dgll_b200 import ops
    idx = torch.tensor([[0, 1, 2, 2], [1, 0, 0, 1]], device="cuda")
    adj = torch.sparse_coo_tensor(idx, torch.ones(4, device="cuda"), (3, 3)).coalesce()
    x = torch.randn(3, 8, device="cuda")
    a = ops.spmm(adj, x)
    assert ops.as_csr(adj) is ops.as_csr(adj)                   # converted once per tensor object
    adj.values().mul_(2.0)                                      # in-place edit: the cached CSR must not be reused
    b = ops.spmm(adj, x)
    assert torch.allclose(b, 2 * a)


@pytest.mark.timeout(120)
def test_pipeline_surfaces_thread_failures_instead_of_hanging(nn):
    """An exception in the producer (dataloader / fetch) or in the consumer (forward) ends run_epoch with that exception:
    the sentinel is always queued, a failing consumer releases a producer blocked on the full queue."""
    from dgll_b200 import graphs as G, pipeline
    N, F = 3000, 16
    rp, col = G.rmat_csr(N, N * 10, seed=1, device="cuda")
    table = G.feature_table(N, F, seed=2)
    labels = torch.randint(0, 3, (N,), device="cuda")
    model = nn.GraphSAGE(F, 8, 3, 2, torch.relu, 0.0).cuda()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)

    def loader(fail_at=None, n=12):
        for b in range(n):
            if fail_at is not None and b == fail_at:
                raise ValueError("loader failed")
            s = torch.arange(b * 64, (b + 1) * 64, device="cuda")
            blocks = G.sample_blocks(rp, col, s, (5, 5), rng_seed=b)
            yield blocks[0].src_ids, s, blocks

    fwd = lambda m, mfgs, feat: m(mfgs, None, feat_table=table)  # noqa: E731
    ok = pipeline.run_epoch(loader(), model, opt, labels=labels, forward=fwd, BUFFER_SIZE=2)
    assert ok["n_batches"] == 12
    with pytest.raises(ValueError, match="loader failed"):
        pipeline.run_epoch(loader(fail_at=5), model, opt, labels=labels, forward=fwd, BUFFER_SIZE=2)

    def bad_forward(m, mfgs, feat):
        raise RuntimeError("forward failed")

    with pytest.raises(RuntimeError, match="forward failed"):
        pipeline.run_epoch(loader(), model, opt, labels=labels, forward=bad_forward, BUFFER_SIZE=2)


def test_gat_layer_trains_on_a_sampled_block(nn):
    """A GAT layer on a NON-square graph (a sampled block: n_dst < n_src, dst nodes first): forward and all gradients
    against an fp64 restatement of gatconv.py:30-54 restricted to the block's edges."""
    from dgll_b200 import ops
    rng = np.random.default_rng(12)
    n_src, n_dst, Fi, D = 400, 90, 20, 16
    deg = rng.integers(1, 12, size=n_dst)
    rp = np.zeros(n_dst + 1, dtype=np.int64)
    rp[1:] = np.cumsum(deg)
    col = rng.integers(0, n_src, size=int(rp[-1])).astype(np.int32)
    g = ops.CsrGraph(cu(rp, torch.int64), cu(col, torch.int32), n_src=n_src)
    layer = nn.gatConv(Fi, D, dropout=0.0, alpha=0.2, concat=True).cuda()
    x = rng.standard_normal((n_src, Fi)).astype(np.float32)
    tx = cu(x).requires_grad_(True)
    out = layer(tx, g)
    assert tuple(out.shape) == (n_dst, D)
    W = layer.W.detach().double().cpu().requires_grad_(True)
    a = layer.a.detach().double().cpu().requires_grad_(True)
    rx = torch.from_numpy(x).double().requires_grad_(True)
    Wh = rx @ W
    rows = torch.from_numpy(np.repeat(np.arange(n_dst), deg))
    cols = torch.from_numpy(col.astype(np.int64))
    e = torch.nn.functional.leaky_relu((Wh[:n_dst] @ a[:D])[rows, 0] + (Wh @ a[D:])[cols, 0], 0.2)
    emax = torch.full((n_dst,), -1e30, dtype=torch.float64).scatter_reduce(0, rows, e.detach(), reduce="amax")
    w = torch.exp(e - emax[rows])
    alpha = w / torch.zeros(n_dst, dtype=torch.float64).index_add(0, rows, w)[rows]
    ref = torch.nn.functional.elu(torch.zeros((n_dst, D), dtype=torch.float64).index_add(0, rows, alpha[:, None] * Wh[cols]))
    assert rel_err(out.detach().cpu(), ref.detach()) <= OUT_TOL
    gout = torch.from_numpy(rng.standard_normal((n_dst, D)).astype(np.float32))
    out.backward(gout.cuda())
    ref.backward(gout.double())
    assert rel_err(tx.grad.cpu(), rx.grad) <= GRAD_TOL
    assert rel_err(layer.W.grad.cpu(), W.grad) <= GRAD_TOL
    assert rel_err(layer.a.grad.cpu(), a.grad) <= GRAD_TOL


# ------------------------------------------------------ binarized GCN ------
def test_bin_gcn_conv_forward_exact_and_ste_backward(nn):
    """BinGCNConv (README.md:11): forward = (mean of the neighbours' sign bits) W + b with integer-exact
    counts from the popcount kernel; backward = straight-through estimator, checked against torch autograd of the same
    definition (sign replaced by a clipped identity in the backward pass)."""
    rng = np.random.default_rng(5)
    n, Fi, H = 700, 70, 24
    deg = rng.integers(0, 30, size=n)
    rp = np.zeros(n + 1, dtype=np.int64)
    rp[1:] = np.cumsum(deg)
    col = rng.integers(0, n, size=int(rp[-1])).astype(np.int32)
    x = (rng.standard_normal((n, Fi)) * 1.5).astype(np.float32)
    from dgll_b200 import ops
    g = ops.CsrGraph(cu(rp, torch.int64), cu(col, torch.int32))
    layer = nn.BinGCNConv(Fi, H).cuda()
    tx = cu(x).requires_grad_(True)
    out = layer(tx, g)
    # reference: dense restatement in fp64 with a straight-through sign
    A = np.zeros((n, n))
    np.add.at(A, (np.repeat(np.arange(n), deg), col), 1.0)
    Am = torch.from_numpy(A / np.maximum(deg, 1)[:, None])
    rx = torch.from_numpy(x).double().requires_grad_(True)

    class SignSTE(torch.autograd.Function):
        @staticmethod
        def forward(ctx, v):
            ctx.save_for_backward(v)
            return torch.where(v >= 0, torch.ones_like(v), -torch.ones_like(v))

        @staticmethod
        def backward(ctx, gr):
            (v,) = ctx.saved_tensors
            return gr * (v.abs() <= 1.0)

    w, b = layer.weight.detach().double().cpu().requires_grad_(True), layer.bias.detach().double().cpu().requires_grad_(True)
    ref = (Am @ SignSTE.apply(rx)) @ w + b
    assert rel_err(out.detach().cpu(), ref.detach()) <= OUT_TOL
    gout = torch.from_numpy(rng.standard_normal((n, H)).astype(np.float32))
    out.backward(gout.cuda())
    ref.backward(gout.double())
    assert rel_err(tx.grad.cpu(), rx.grad) <= GRAD_TOL
    assert rel_err(layer.weight.grad.cpu(), w.grad) <= GRAD_TOL
    assert rel_err(layer.bias.grad.cpu(), b.grad) <= GRAD_TOL
    # the aggregation itself is integer-exact: equal to the fp32 kernel on sign(x)
    agg = ops.binarized_aggregate(g, x=cu(x), mode="sum")
    sx = torch.where(cu(x) >= 0, 1.0, -1.0)
    assert torch.equal(agg, ops.spmm(g, sx, reduce="sum"))
    # the two-layer model trains (loss goes down on a fixed batch)
    torch.manual_seed(0)
    model = nn.BinGCN(Fi, 32, 5, 0.0).cuda()
    opt = torch.optim.Adam(model.parameters(), lr=0.02)
    y = cu(rng.integers(0, 5, size=n), torch.int64)
    losses = []
    for _ in range(30):
        opt.zero_grad()
        loss = torch.nn.functional.nll_loss(model(cu(x), g), y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < 0.9 * losses[0]


# ----------------------------------------------------------------- SAGE ----
@pytest.mark.parametrize("aggr,combine", [("mean", "sum"), ("sum", "concat"), ("max", "sum")])
def test_sage_conv_matches_restated_oracle(nn, aggr, combine):
    rng = np.random.default_rng(1)
    B, Kf, Fi, H = 64, 10, 33, 16
    layer = nn.sageConv(Fi, H, aggr_neighbor_method=aggr, aggr_hid_method=combine).cuda()
    src = rng.standard_normal((B, Fi)).astype(np.float32)
    neigh = rng.standard_normal((B, Kf, Fi)).astype(np.float32)
    ts, tn = cu(src).requires_grad_(True), cu(neigh).requires_grad_(True)
    out = layer(ts, tn)
    out.sum().backward()
    ws = layer.weight.detach().double().cpu().requires_grad_(True)
    wn = layer.neighborAgg.weight.detach().double().cpu().requires_grad_(True)
    rs = torch.from_numpy(src).double().requires_grad_(True)
    rn = torch.from_numpy(neigh).double().requires_grad_(True)
    ref = L.sage_conv(rs, rn, ws, wn, aggr=aggr, combine=combine)
    ref.sum().backward()
    assert rel_err(out.detach().cpu(), ref.detach()) <= OUT_TOL
    assert rel_err(ts.grad.cpu(), rs.grad) <= GRAD_TOL
    assert rel_err(tn.grad.cpu(), rn.grad) <= GRAD_TOL
    assert rel_err(layer.weight.grad.cpu(), ws.grad) <= GRAD_TOL
    assert rel_err(layer.neighborAgg.weight.grad.cpu(), wn.grad) <= GRAD_TOL


def test_graphsage_model_on_multihop_sampler_output(nn):
    """GraphSage.forward (sageconv.py:103-114) fed by the reference's fixed-fanout sampler (restated bit-exactly in
    dgll_b200.data.multihop_sampling).  Equal fanouts: the reference indexes ``num_neighbors_list[l]`` by LAYER for
    every hop (:111), which only type-checks when all fanouts are equal — kept as is."""
    from dgll_b200 import ops
    from dgll_b200.data import multihop_sampling
    g = golden("sampler_multihop")
    ptr, flat = g["nbr_ptr"], g["nbr_flat"]
    table_nb = {i: flat[ptr[i]:ptr[i + 1]] for i in range(len(ptr) - 1)}
    fan = [4, 4]
    np.random.seed(11)
    hops = multihop_sampling(g["seeds"], fan, table_nb)
    rng = np.random.default_rng(2)
    feats = rng.standard_normal((len(ptr) - 1, 20)).astype(np.float32)
    model = nn.GraphSage(20, hidden_dim=[12, 6], num_neighbors_list=fan).cuda()
    table = cu(feats)
    flist = [ops.gather_rows(table, cu(h, torch.int64)) for h in hops]
    out = model(flist)
    layers = [(m.weight.detach().double().cpu(), m.neighborAgg.weight.detach().double().cpu()) for m in model.gcn]
    ref = L.graphsage_model([torch.from_numpy(feats[h]).double() for h in hops], layers, fan)
    assert out.shape == (len(g["seeds"]), 6)
    assert rel_err(out.detach().cpu(), ref) <= OUT_TOL


def _rand_block(rng, n_dst, n_src, fanout):
    deg = rng.integers(0, fanout + 1, size=n_dst)
    rp = np.zeros(n_dst + 1, dtype=np.int64)
    rp[1:] = np.cumsum(deg)
    col = rng.integers(0, n_src, size=int(rp[-1])).astype(np.int32)
    return rp, col


@pytest.mark.parametrize("fin,fout", [(40, 16), (16, 40)])
def test_block_sage_and_graphconv_match_dgl_semantics(nn, fin, fout):
    from dgll_b200.data import create_block
    rng = np.random.default_rng(fin)
    n_dst, n_src = 120, 400
    rp, col = _rand_block(rng, n_dst, n_src, 10)
    blk = create_block(("csc", (rp, col, [])), num_src_nodes=n_src, num_dst_nodes=n_dst)
    x = rng.standard_normal((n_src, fin)).astype(np.float32)
    sage = nn.SAGEConv(fin, fout, "mean").cuda()
    with torch.no_grad():
        sage.bias.uniform_(-1, 1)
    xs = cu(x).requires_grad_(True)
    out = sage(blk, xs)
    out.sum().backward()
    rx = torch.from_numpy(x).double().requires_grad_(True)
    ref = L.dgl_sage_conv_mean(rp, col, rx, n_dst, sage.fc_self.weight.detach().double().cpu().t(),
                               sage.fc_neigh.weight.detach().double().cpu().t(), sage.bias.detach().double().cpu())
    ref.sum().backward()
    assert rel_err(out.detach().cpu(), ref.detach()) <= OUT_TOL
    assert rel_err(xs.grad.cpu(), rx.grad) <= GRAD_TOL
    conv = nn.GraphConv(fin, fout, activation=torch.relu).cuda()
    with torch.no_grad():
        conv.bias.uniform_(-1, 1)
    out = conv(blk, cu(x))
    ref = L.dgl_graph_conv_both(rp, col, torch.from_numpy(x).double(), n_dst, conv.weight.detach().double().cpu(),
                                conv.bias.detach().double().cpu(), relu=True)
    assert rel_err(out.detach().cpu(), ref) <= OUT_TOL


def test_sage_layer0_gather_fused_equals_materialised(nn):
    """Layer 0 aggregating straight from the padded global table (gather fused) == gather-then-aggregate."""
    from dgll_b200 import graphs as G
    g = torch.Generator(device="cuda").manual_seed(0)
    N, F = 5000, 602
    rp, col = G.uniform_csr(N, 40, seed=1)
    table = G.feature_table(N, F, seed=2)
    seeds = torch.randperm(N, device="cuda", generator=g)[:256]
    b0, b1 = G.sample_blocks(rp, col, seeds, (25, 10), rng_seed=3)
    sage = nn.SAGEConv(F, 64, "mean").cuda()
    fused = sage(b0, None, feat_table=table)
    from dgll_b200 import ops
    h_src = ops.gather_rows(table, b0.src_ids)[:, :F].contiguous()
    plain = sage(b0, h_src)
    assert rel_err(fused.detach().cpu(), plain.detach().cpu()) <= OUT_TOL
    assert b0.src_ids[:b0.num_dst].equal(b1.src_ids)  # dst-first convention: block0's dst = block1's src


# ------------------------------------------------------- gcn_extension -----
def test_gcn_extension_drop_in(nn):
    from dgll_b200 import gcn_extension as ext
    rng = np.random.default_rng(7)
    N, F, Hd = 591, 50, 64
    a = (rng.random((N, N)) < 0.02)
    a_hat = L.sym_norm_adjacency(a.astype(np.float64))
    rows, cols = np.nonzero(a_hat)
    rp, col, val = oracle.coo_to_csr(rows, cols, N, a_hat[rows, cols])
    X = np.zeros((N, 52), dtype=np.float32)
    X[:, :F] = rng.standard_normal((N, F))
    model = ext.GCN(F, Hd, 121)
    args = (cu(rp, torch.int32), cu(col, torch.int32), cu(val), cu(X), cu(np.diff(rp), torch.int32))
    h = model(*args)
    ref1 = L.fused_gcn_layer(torch.from_numpy(X[:, :F]).double(), torch.from_numpy(a_hat),
                             model.layer1.W.detach().double().cpu()[:F])
    ref2 = L.fused_gcn_layer(ref1, torch.from_numpy(a_hat), model.layer2.W.detach().double().cpu())
    assert rel_err(h.detach().cpu(), ref2) <= OUT_TOL
    loss = torch.nn.functional.binary_cross_entropy_with_logits(h, torch.ones_like(h))
    loss.backward()
    assert model.layer1.W.grad is not None and torch.isfinite(model.layer1.W.grad).all()
    # reference error behaviour: non-CUDA tensor -> RuntimeError (TORCH_CHECK is_cuda, gcn_extension.cpp:31-36)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        ext.gcn_fused_forward(args[0].cpu(), *args[1:4], model.layer1.W, args[4], F)
    # backward without H (the reference signature) == backward with H
    H = ext.gcn_fused_forward(args[0], args[1], args[2], args[3], model.layer1.W.detach(), args[4], F)
    gout = torch.randn_like(H)
    a_, b_ = ext.gcn_fused_backward(gout, args[0], args[1], args[2], args[3], model.layer1.W.detach(), args[4], F)
    c_, d_ = ext.gcn_fused_backward(gout, args[0], args[1], args[2], args[3], model.layer1.W.detach(), args[4], F, H=H)
    assert torch.equal(a_, c_) and torch.equal(b_, d_)


def test_gcn_extension_reference_constructors_and_state_dict():
    """The classes take the reference's positional arguments (train_gcn.py:25,44-47) and carry its parameter shapes, also
    when the hidden width is not a multiple of 4: a reference-shaped state dict loads and the model runs."""
    from dgll_b200 import gcn_extension as ext
    layer = ext.GCNLayer(52, 50, 30)                       # (in_features_padded, actual_in_features, out_features)
    assert tuple(layer.W.shape) == (52, 30) and layer.actual_F == 50
    model = ext.GCN(50, 30, 121)                           # hidden 30 is not a multiple of 4
    assert model.input_dim_padded == 52
    assert tuple(model.layer1.W.shape) == (52, 30) and tuple(model.layer2.W.shape) == (30, 121)
    assert model.layer2.actual_F == 30
    sd = {"layer1.W": torch.randn(52, 30), "layer2.W": torch.randn(30, 121)}
    model.load_state_dict(sd)
    rng = np.random.default_rng(3)
    N = 300
    a = (rng.random((N, N)) < 0.03)
    a_hat = L.sym_norm_adjacency(a.astype(np.float64))
    rows, cols = np.nonzero(a_hat)
    rp, col, val = oracle.coo_to_csr(rows, cols, N, a_hat[rows, cols])
    X = rng.standard_normal((N, 50)).astype(np.float32)
    h = model(cu(rp, torch.int32), cu(col, torch.int32), cu(val), cu(X), cu(np.diff(rp), torch.int32))
    ref1 = L.fused_gcn_layer(torch.from_numpy(X).double(), torch.from_numpy(a_hat), sd["layer1.W"].double()[:50])
    ref2 = L.fused_gcn_layer(ref1, torch.from_numpy(a_hat), sd["layer2.W"].double())
    assert rel_err(h.detach().cpu(), ref2) <= OUT_TOL
    h.sum().backward()
    assert torch.isfinite(model.layer2.W.grad).all() and tuple(model.layer2.W.grad.shape) == (30, 121)


# ------------------------------------------------------------ data path ----
def _golden_dgraph(g):
    from dgll_b200.data import DGraph
    ptr, flat = g["nbr_ptr"], g["nbr_flat"]
    edges = [[int(v) for v in flat[ptr[i]:ptr[i + 1]]] for i in range(len(ptr) - 1)]
    n = len(edges)
    return DGraph(nodes=torch.arange(n), edges=edges, labels=cu(g["labels"], torch.int64), features=cu(g["feats"]),
                  train_mask=torch.zeros(n, dtype=torch.bool), test_mask=None, validation_mask=None, device="cuda")


def test_neighbor_sampler_and_graph_store_bit_exact_with_reference():
    """The reference's DGLLNeighborSampler / DGraph / sugbraph run under random.seed -> golden; ours must reproduce
    the index lists, induced adjacency and gathered features bit for bit."""
    from dgll_b200.data import DGLLNeighborSampler
    g = golden("sampler_neighbor")
    dg = _golden_dgraph(g)
    random.seed(int(g["py_seed"]))
    sampler = DGLLNeighborSampler([int(v) for v in g["fanouts"]])
    inp, outp, subgs = sampler.sample(dg, torch.as_tensor(g["seeds"]))
    assert np.array_equal(inp.numpy(), g["input_nodes"]) and np.array_equal(outp.numpy(), g["output_nodes"])
    for i, sg in enumerate(subgs):
        assert np.array_equal(sg.src_nodes().numpy(), g["b%d_src" % i])
        assert np.array_equal(sg.dst_nodes().numpy(), g["b%d_dst" % i])
        assert np.array_equal(sg.nodes().numpy(), g["b%d_nodes" % i])
    adj = sampler.get_adj(dg, subgs)
    assert adj.dtype == torch.int32 and np.array_equal(adj.cpu().numpy(), g["adj"])
    feats = subgs[0].get_features(dg, subgs)
    assert np.array_equal(feats.cpu().numpy(), g["gathered"])
    assert np.array_equal(dg.get_induced_subgraph(torch.tensor([0, 2, 5, 6, 9, 23])).cpu().numpy(), g["induced"])
    assert dg.get_neighbors(torch.tensor([0, 2, 5])) == json.loads(bytes(g["neighbors_json"]).decode())
    assert np.array_equal(dg.get_features(torch.as_tensor(g["seeds"])).cpu().numpy(), g["feats_sel"])
    assert np.array_equal(dg.get_labels(torch.as_tensor(g["seeds"])).cpu().numpy(), g["labels_sel"])
    # the host-sampled block aggregates on the device exactly like the oracle on the same lists
    from dgll_b200 import kernels as K
    seeds1 = subgs[1].src_nodes()          # seeds of the inner layer = raw src list with duplicates
    rp0, col0 = subgs[0].to_csr(seeds1)
    ref = oracle.spmm_csr(rp0.numpy(), col0.numpy(), g["feats"], reduce="mean")
    out = K.spmm_csr(rp0.cuda(), col0.cuda(), dg.features, reduce="mean")
    assert rel_err(out.cpu(), ref) <= OUT_TOL


def test_multihop_sampling_bit_exact_with_reference():
    from dgll_b200.data import multihop_sampling
    g = golden("sampler_multihop")
    ptr, flat = g["nbr_ptr"], g["nbr_flat"]
    table = {i: flat[ptr[i]:ptr[i + 1]] for i in range(len(ptr) - 1)}
    np.random.seed(int(g["np_seed"]))
    hops = multihop_sampling(g["seeds"], [int(v) for v in g["fanouts"]], table)
    assert np.array_equal(hops[1], g["hop1"]) and np.array_equal(hops[2], g["hop2"])


def test_device_sampler_structure_and_dataloader():
    from dgll_b200.data import BlockDataLoader, DataLoader, DGLLNeighborSampler, NeighborSampler
    g = golden("sampler_neighbor")
    dg = _golden_dgraph(g)
    sampler = DGLLNeighborSampler([4, 3], device_sampling=True, rng_seed=1)
    train = torch.as_tensor(g["train_nodes"])
    n_batches = 0
    for inp, outp, subgs in DataLoader(dg, train, sampler, batch_size=64):
        n_batches += 1
        assert len(subgs) == 2 and outp.numel() <= 64
        ptr, flat = g["nbr_ptr"], g["nbr_flat"]
        src, dst = subgs[1].src_nodes().cpu().numpy(), subgs[1].dst_nodes().cpu().numpy()
        for s, d in zip(src, dst):
            assert s in flat[ptr[d]:ptr[d + 1]]
        cnt = np.bincount(dst, minlength=len(ptr) - 1)
        deg = np.diff(ptr)
        assert np.all(cnt[outp.numpy()] == np.minimum(deg[outp.numpy()], 3) * np.bincount(outp.numpy(), minlength=len(ptr) - 1)[outp.numpy()])
    assert n_batches == (len(train) + 63) // 64
    loader = BlockDataLoader(dg, train, NeighborSampler([5, 5]), batch_size=50, shuffle=True, drop_last=True)
    seen = 0
    for input_nodes, output_nodes, mfgs in loader:
        seen += 1
        assert mfgs[0].is_block and mfgs[1].num_dst_nodes() == 50
        assert mfgs[0].num_dst_nodes() == mfgs[1].num_src_nodes()
        assert input_nodes.equal(mfgs[0].src_ids) and input_nodes[:50].equal(output_nodes)
    assert seen == len(loader) == len(train) // 50


def test_graph_cache_server_matches_reference_semantics():
    """GraphCacheServer (FeatureCache/storage.py) vs the oracle's CacheServer restatement: same fill policy, same
    frames (bit-exact), same miss rate."""
    from dgll_b200.data import GraphCacheServer, NodeFlow
    rng = np.random.RandomState(0)
    n = 3000
    host = {"features": rng.randn(n, 604).astype(np.float32), "norm": rng.rand(n, 1).astype(np.float32)}
    nid_map = rng.permutation(n)
    deg = rng.randint(0, 50, size=n)
    ref = S.CacheServer(host, n, nid_map)
    ref.auto_cache(deg, 700, ["features", "norm"])
    srv = GraphCacheServer({k: torch.from_numpy(v) for k, v in host.items()}, n, torch.from_numpy(nid_map), 0)
    srv.init_field(["features", "norm"])
    assert srv.total_dim == 605
    srv.auto_cache(torch.from_numpy(deg), ["features", "norm"], capability=700)
    assert np.array_equal(srv.gpu_flag.cpu().numpy(), ref.gpu_flag)
    assert np.array_equal(srv.localid2cacheid.cpu().numpy()[ref.gpu_flag], ref.localid2cacheid[ref.gpu_flag])
    srv.log = True
    layers = [rng.randint(0, n, size=m) for m in (2000, 400, 50)]
    nf = NodeFlow([torch.from_numpy(l) for l in layers])
    srv.fetch_data(nf)
    for i, ids in enumerate(layers):
        fr = ref.fetch(ids)
        assert np.array_equal(nf._node_frames[i]["features"].cpu().numpy(), fr["features"])
        assert np.array_equal(nf._node_frames[i]["norm"].cpu().numpy(), fr["norm"])
    assert abs(srv.get_miss_rate() - ref.get_miss_rate()) < 1e-12
    # full cache path (capability >= node_num): fetch_from_cache
    srv2 = GraphCacheServer({k: torch.from_numpy(v) for k, v in host.items()}, n, torch.from_numpy(nid_map), 0)
    srv2.auto_cache(torch.from_numpy(deg), ["features", "norm"], capability=n)
    assert srv2.full_cached
    nf2 = NodeFlow([torch.from_numpy(layers[0])])
    srv2.fetch_data(nf2)
    assert np.array_equal(nf2._node_frames[0]["features"].cpu().numpy(), host["features"][nid_map[layers[0]]])


def test_graph_cache_server_pinned_by_the_reference_run():
    """tests/golden/cache_server.npz = what the reference's own GraphCacheServer produced (storage.py executed in place
    by oracle/gen_golden_pins.py).  With the reference's cache set, bookkeeping arrays, per-layer frames (bit-exact) and
    the miss statistics must be identical; auto_cache must pick a set that obeys the same out-degree policy."""
    from dgll_b200.data import GraphCacheServer, NodeFlow
    g = golden("cache_server")
    host = {"features": torch.from_numpy(g["feats"]), "norm": torch.from_numpy(g["norm"])}
    n = g["nid_map"].size
    nid_map = torch.from_numpy(g["nid_map"])
    layers = [torch.from_numpy(g["layer%d" % i]) for i in range(3)]
    flag = g["part_gpu_flag"].astype(bool)
    ids = np.nonzero(flag)[0]
    nids = torch.from_numpy(ids[np.argsort(g["part_localid2cacheid"][ids])])
    srv = GraphCacheServer(host, n, nid_map, 0)
    srv.init_field(["features", "norm"])
    assert srv.total_dim == 14
    srv.cache_fix_data(nids.cuda(), srv.get_feat_from_server(nids.cuda(), ["features", "norm"]))
    assert np.array_equal(srv.gpu_flag.cpu().numpy(), flag)
    assert np.array_equal(srv.localid2cacheid.cpu().numpy()[flag], g["part_localid2cacheid"][flag])
    for name in host:
        assert np.array_equal(srv.gpu_fix_cache[name].cpu().numpy(), g["part_cache_" + name])
    srv.log = True
    nf = NodeFlow(layers)
    srv.fetch_data(nf)
    for i in range(3):
        for name in host:
            assert np.array_equal(nf._node_frames[i][name].cpu().numpy(), g["part_frame%d_%s" % (i, name)])
    assert srv.get_miss_rate() == float(g["part_miss_rate"])
    # to_gpu=True (storage.py:120-122): rows come back on the device
    fr = srv.get_feat_from_server(layers[2].cuda(), ["features"], to_gpu=True)
    assert fr["features"].is_cuda and np.array_equal(fr["features"].cpu().numpy(), g["feats"][g["nid_map"][g["layer2"]]])
    # the fill policy on its own: `capability` nodes, none with a smaller out-degree than an uncached one
    srv2 = GraphCacheServer(host, n, nid_map, 0)
    srv2.auto_cache(torch.from_numpy(g["out_deg"]), ["features", "norm"], capability=150)
    f2 = srv2.gpu_flag.cpu().numpy()
    assert f2.sum() == 150 and g["out_deg"][f2].min() >= g["out_deg"][~f2].max()
    strict = g["out_deg"] > g["out_deg"][flag].min()
    assert np.array_equal(f2[strict], flag[strict])
    # fully cached
    srv3 = GraphCacheServer(host, n, nid_map, 0)
    srv3.auto_cache(torch.from_numpy(g["out_deg"]), ["features", "norm"], capability=n)
    assert srv3.full_cached and np.array_equal(srv3.localid2cacheid.cpu().numpy(), g["full_localid2cacheid"])
    nf3 = NodeFlow(layers)
    srv3.fetch_data(nf3)
    for i in range(3):
        for name in host:
            assert np.array_equal(nf3._node_frames[i][name].cpu().numpy(), g["full_frame%d_%s" % (i, name)])


@pytest.mark.parametrize("aggr", ["mean", "sum", "max"])
@pytest.mark.parametrize("combine", ["sum", "concat"])
def test_sage_conv_pinned_by_the_repaired_reference(nn, aggr, combine):
    """tests/golden/nn_sageconv_fixed.npz: the reference's sageconv.py run with its two documented one-line repairs
    applied to the AST (oracle/gen_golden_pins.py) — outputs and all four gradients."""
    g = golden("nn_sageconv_fixed")
    k = "%s_%s_" % (aggr, combine)
    layer = nn.sageConv(g["src"].shape[1], g[k + "w_self"].shape[1], aggr_neighbor_method=aggr,
                        aggr_hid_method=combine).cuda()
    with torch.no_grad():
        layer.weight.copy_(cu(g[k + "w_self"]))
        layer.neighborAgg.weight.copy_(cu(g[k + "w_neigh"]))
    ts, tn = cu(g["src"]).requires_grad_(True), cu(g["neigh"]).requires_grad_(True)
    out = layer(ts, tn)
    assert rel_err(out.detach().cpu(), g[k + "out"]) <= OUT_TOL
    out.backward(cu(g[k + "g"]))
    assert rel_err(ts.grad.cpu(), g[k + "d_src"]) <= GRAD_TOL
    assert rel_err(tn.grad.cpu(), g[k + "d_neigh"]) <= GRAD_TOL
    assert rel_err(layer.weight.grad.cpu(), g[k + "d_w_self"]) <= GRAD_TOL
    assert rel_err(layer.neighborAgg.weight.grad.cpu(), g[k + "d_w_neigh"]) <= GRAD_TOL


def test_graphsage_model_pinned_by_the_repaired_reference(nn):
    g = golden("nn_sageconv_fixed")
    fan = [int(v) for v in g["model_fan"]]
    model = nn.GraphSage(g["model_hop0"].shape[1], hidden_dim=[g["model_l0_w_self"].shape[1], g["model_l1_w_self"].shape[1]],
                         num_neighbors_list=fan).cuda()
    with torch.no_grad():
        for i, layer in enumerate(model.gcn):
            layer.weight.copy_(cu(g["model_l%d_w_self" % i]))
            layer.neighborAgg.weight.copy_(cu(g["model_l%d_w_neigh" % i]))
    out = model([cu(g["model_hop%d" % i]) for i in range(3)])
    assert rel_err(out.detach().cpu(), g["model_out"]) <= OUT_TOL


# ------------------------------------------------------- normalisation -----
def test_normalisation_helpers_match_reference():
    """a17: host helpers reproduce the reference's golden ``normalize`` output; device versions equal them."""
    import scipy.sparse as sp
    from dgll_b200 import ops
    from dgll_b200.nn import utils as U
    g = golden("nn_normalize")
    raw = sp.csr_matrix(g["raw"])
    out = U.normalize(raw)
    assert rel_err(np.asarray(out.todense()), g["dense"]) <= 1e-6
    t = U.sparse_mx_to_torch_sparse_tensor(out, device="cuda").coalesce()
    assert np.array_equal(t.indices().cpu().numpy(), g["indices"]) and rel_err(t.values().cpu().numpy(), g["values"]) <= 1e-6
    gcsr = ops.CsrGraph.from_dense(cu(g["raw"]))
    nz = (cu(g["raw"]) > 0).nonzero()
    gcsr = gcsr.with_values(cu(g["raw"])[nz[:, 0], nz[:, 1]])
    dv = U.row_normalize_csr(gcsr)
    dense = torch.zeros(g["raw"].shape, device="cuda")
    rows = torch.repeat_interleave(torch.arange(dv.n_dst, device="cuda"), dv.row_ptr[1:] - dv.row_ptr[:-1])
    dense[rows, dv.col.long()] = dv.values
    assert rel_err(dense.cpu().numpy(), g["dense"]) <= 1e-6
    a = (np.random.default_rng(0).random((50, 50)) < 0.1).astype(np.float64)
    a = ((a + a.T) > 0).astype(np.float64)
    ref = L.sym_norm_adjacency(a)
    sv = U.sym_normalize_csr(ops.CsrGraph.from_dense(cu(a)))
    dense = torch.zeros((50, 50), device="cuda")
    rows = torch.repeat_interleave(torch.arange(50, device="cuda"), sv.row_ptr[1:] - sv.row_ptr[:-1])
    dense[rows, sv.col.long()] = sv.values
    assert rel_err(dense.cpu().numpy(), ref) <= 1e-6
    assert rel_err(np.asarray(U.normalize_lap(sp.csr_matrix(a)).todense()), ref) <= 1e-6


# ------------------------------------------------------------ pipeline -----
def test_pipelined_epoch_equals_sequential_epoch(nn):
    """MQ-GNN style producer/consumer epoch (dgll_b200.pipeline) == the sequential loop on the same seeds: every kernel
    on the path is deterministic, so the losses and the final weights agree bit for bit."""
    import copy
    from dgll_b200 import graphs as G, train as T
    N, F = 20000, 100
    rp, col = G.rmat_csr(N, N * 20, seed=1, device="cuda")
    table = G.feature_table(N, F, seed=2)
    gen = torch.Generator(device="cuda").manual_seed(3)
    labels = torch.randint(0, 7, (N,), device="cuda", generator=gen)
    seeds = torch.randperm(N, device="cuda", generator=gen)[:1024 * 6]
    torch.manual_seed(0)
    m1 = nn.GraphSAGE(F, 64, 7, 2, torch.relu, 0.0).cuda()
    m2 = copy.deepcopy(m1)
    o1 = torch.optim.SGD(m1.parameters(), lr=0.05)
    o2 = torch.optim.SGD(m2.parameters(), lr=0.05)
    a = T.sage_epoch(m1, o1, table, labels, F, rp, col, seeds, (10, 5), 1024, rng_seed=4)
    b = T.sage_epoch_pipelined(m2, o2, table, labels, rp, col, seeds, (10, 5), 1024, rng_seed=4)
    assert a["n_batches"] == b["n_batches"] == 6
    assert abs(a["loss"] - b["loss"]) <= 1e-6 * abs(a["loss"])
    for p, q in zip(m1.parameters(), m2.parameters()):
        assert torch.equal(p, q)


# ------------------------------------------------- CUDA-graph training step --
@pytest.mark.parametrize("optimizer", ["sgd_eager_step", "adam_captured"])
def test_graphed_training_step_equals_eager_epoch(nn, optimizer):
    """train.GraphedSageTrainer replays ONE captured step on fixed-capacity block buffers (rows past the real count
    have degree 0, unused edge slots hold the padding column the transpose drops, a short last batch is masked out of
    the loss).  Same losses and weights as the eager loop on the same pre-sampled mini-batches (fp32 GEMMs)."""
    import copy
    from dgll_b200 import graphs as G, train as T
    N, F = 20000, 100
    rp, col = G.rmat_csr(N, N * 20, seed=1, device="cuda")
    table = G.feature_table(N, F, seed=2)
    gen = torch.Generator(device="cuda").manual_seed(3)
    labels = torch.randint(0, 7, (N,), device="cuda", generator=gen)
    seeds = torch.randperm(N, device="cuda", generator=gen)[:512 * 5 + 77]        # short last batch
    torch.manual_seed(0)
    m1 = nn.GraphSAGE(F, 64, 7, 2, torch.relu, 0.0).cuda()
    m2 = copy.deepcopy(m1)
    if optimizer == "adam_captured":
        o1 = torch.optim.Adam(m1.parameters(), lr=0.01, fused=True)
        o2 = torch.optim.Adam(m2.parameters(), lr=0.01, fused=True, capturable=True)
    else:
        o1 = torch.optim.SGD(m1.parameters(), lr=0.05)
        o2 = torch.optim.SGD(m2.parameters(), lr=0.05)
    pre = T.make_batches(rp, col, seeds, (10, 5), 512, rng_seed=4)
    assert len(pre) == 6 and pre[-1][0].numel() == 77
    a = T.sage_epoch(m1, o1, table, labels, F, batches=pre, precision="fp32")
    tr = T.GraphedSageTrainer(m2, o2, table, labels, 512, (10, 5), precision="fp32")
    w0 = [p.detach().clone() for p in m2.parameters()]
    tr.load(*pre[0])
    tr.capture()
    for p, q in zip(m2.parameters(), w0):
        assert torch.equal(p, q)                     # capture (with its warm-up steps) leaves the weights untouched
    b = tr.epoch(pre)
    assert a["n_batches"] == b["n_batches"] == 6
    assert abs(a["loss"] - b["loss"]) <= 1e-5 * abs(a["loss"])
    for p, q in zip(m1.parameters(), m2.parameters()):
        assert rel_err(q.detach().cpu().numpy(), p.detach().cpu().numpy()) <= 2e-5
    # production of batch i+1 on a side stream under the replay of step i (the default for iterators): same update,
    # bit for bit, as the serial replay
    m3 = copy.deepcopy(m1)
    for p, q in zip(m3.parameters(), w0):
        p.data.copy_(q)
    o3 = (torch.optim.Adam(m3.parameters(), lr=0.01, fused=True, capturable=True) if optimizer == "adam_captured"
          else torch.optim.SGD(m3.parameters(), lr=0.05))
    tr3 = T.GraphedSageTrainer(m3, o3, table, labels, 512, (10, 5), precision="fp32")
    tr3.load(*pre[0])
    tr3.capture()
    d = tr3.epoch(iter(pre))
    assert d["n_batches"] == 6 and d["loss"] == b["loss"]
    for p, q in zip(m2.parameters(), m3.parameters()):
        assert torch.equal(p, q)
    # a second epoch through the sampler-in-the-loop iterator replays the same graph
    c = tr.epoch(T.iter_batches(rp, col, seeds, (10, 5), 512, rng_seed=9))
    assert c["n_batches"] == 6 and np.isfinite(c["loss"])
    with pytest.raises(ValueError):
        tr.load(seeds[:600], pre[0][1])              # more seeds than the captured capacity


def test_graphed_training_step_with_per_batch_feature_rows(nn):
    """The partitioned case: the input layer's source rows arrive per mini-batch (fetched from the owning GPUs) instead
    of being read from a resident table.  Same update as the eager ``model(blocks, x)`` step."""
    import copy
    from dgll_b200 import graphs as G, ops, train as T
    N, F = 20000, 64
    rp, col = G.rmat_csr(N, N * 20, seed=5, device="cuda")
    table = torch.randn((N, F), device="cuda", generator=torch.Generator(device="cuda").manual_seed(6))
    gen = torch.Generator(device="cuda").manual_seed(7)
    labels = torch.randint(0, 5, (N,), device="cuda", generator=gen)
    seeds = torch.randperm(N, device="cuda", generator=gen)[:256 * 4]
    torch.manual_seed(1)
    m1 = nn.GraphSAGE(F, 32, 5, 2, torch.relu, 0.0).cuda()
    m2 = copy.deepcopy(m1)
    o1 = torch.optim.SGD(m1.parameters(), lr=0.05)
    o2 = torch.optim.SGD(m2.parameters(), lr=0.05)
    pre = T.make_batches(rp, col, seeds, (8, 4), 256, rng_seed=2)
    items = [(s, blocks, table[blocks[0].src_ids]) for s, blocks in pre]
    prev = ops.get_gemm_precision()
    ops.set_gemm_precision("fp32")
    try:
        m1.train()
        losses = []
        for s, blocks, x in items:
            loss = torch.nn.functional.cross_entropy(m1(blocks, x), labels[s])
            o1.zero_grad(set_to_none=True)
            loss.backward()
            o1.step()
            losses.append(loss.item())
        tr = T.GraphedSageTrainer(m2, o2, None, labels, 256, (8, 4), n_feat=F, precision="fp32")
        tr.load(*items[0])
        tr.capture()
        r = tr.epoch(iter(items))                      # iterator => production overlapped with the replay
    finally:
        ops.set_gemm_precision(prev)
    assert r["n_batches"] == 4 and abs(r["loss"] - sum(losses) / 4) <= 1e-5 * abs(sum(losses) / 4)
    for p, q in zip(m1.parameters(), m2.parameters()):
        assert rel_err(q.detach().cpu().numpy(), p.detach().cpu().numpy()) <= 2e-5
    with pytest.raises(ValueError):
        tr.load(items[0][0], items[0][1])              # feature rows are mandatory in this mode
