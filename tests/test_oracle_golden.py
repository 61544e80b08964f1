"""Pin the CPU oracle against golden vectors produced by the reference's own modules
(oracle/gen_golden.py).  CPU-only: runs under -m "not gpu"."""
import json
import random

import numpy as np
import pytest
import torch

import oracle
from oracle import layers as L
from oracle import samplers as S
from conftest import golden, rel_err

T = lambda a: torch.as_tensor(np.asarray(a)).double()


def test_ppi_gcn_forward_loss_and_grads():
    for gi in (8, 5):
        g = golden("ppi_gcn_g%d" % gi)
        x = T(g["feats"])
        ws = [T(g["w0"]).requires_grad_(True), T(g["w1"]).requires_grad_(True)]
        w_out, b_out = T(g["w_out"]).requires_grad_(True), T(g["b_out"]).requires_grad_(True)
        logits = L.ppi_gcn(g["edge_index"], x, ws, w_out, b_out)
        loss = L.ppi_loss(logits, T(g["labels"]))
        loss.backward()
        assert rel_err(logits.detach(), g["logits"]) < 1e-5
        assert abs(loss.item() - float(g["loss"])) / abs(float(g["loss"])) < 1e-5
        assert rel_err(ws[0].grad, g["g_w0"]) < 1e-4
        assert rel_err(ws[1].grad, g["g_w1"]) < 1e-4
        assert rel_err(w_out.grad, g["g_w_out"]) < 1e-4
        h1 = L.ppi_gcn_layer(g["edge_index"], x, ws[0].detach())
        assert rel_err(h1, g["h1"]) < 1e-5


def test_c_spmm_matches_reference_layer_on_ppi():
    g = golden("ppi_gcn_g8")
    ei = g["edge_index"]
    n = g["feats"].shape[0]
    rp, col, _ = oracle.coo_to_csr(ei[0], ei[1], n)
    support = oracle.gemm(g["feats"], g["w0"])
    h1 = oracle.spmm_csr(rp, col, support, reduce="sum", relu=True)
    assert rel_err(h1, g["h1"]) < 1e-5


def test_gcn_model_and_adjacency():
    g = golden("nn_gcn")
    n = g["x"].shape[0]
    adj_dense = L.gcn_adjacency(g["adj_raw"])
    ref_dense = np.zeros((n, n))
    ref_dense[g["adj_indices"][0], g["adj_indices"][1]] = g["adj_values"]
    assert rel_err(adj_dense, ref_dense) < 1e-6
    adj = L.coo_adj(g["adj_indices"], g["adj_values"], n)
    x = T(g["x"]).requires_grad_(True)
    p = [T(g[k]).requires_grad_(True) for k in ("w1", "b1", "w2", "b2")]
    out = L.gcn_model(x, adj, *p)
    loss = torch.nn.functional.nll_loss(out, torch.as_tensor(g["labels"]).long())
    loss.backward()
    assert rel_err(out.detach(), g["out"]) < 1e-5
    assert abs(loss.item() - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    for t, k in zip([x] + p, ("g_x", "g_w1", "g_b1", "g_w2", "g_b2")):
        assert rel_err(t.grad, g[k]) < 1e-4, k
    lo = L.gcn_conv(T(g["x"]), adj, T(g["w1"]), T(g["b1"]))
    assert rel_err(lo, g["layer_out"]) < 1e-5
    # C oracle on the same layer
    rp, col, val = oracle.coo_to_csr(g["adj_indices"][0], g["adj_indices"][1], n, g["adj_values"])
    c = oracle.spmm_csr(rp, col, oracle.gemm(g["x"], g["w1"]), values=val, bias=g["b1"])
    assert rel_err(c, g["layer_out"]) < 1e-5


def _gat_check(tag, sparse):
    g = golden(tag)
    x = T(g["x"]).requires_grad_(True)
    adj = T(g["adj"])
    heads = 4
    Ws = [T(g["W%d" % i]).requires_grad_(True) for i in range(heads)]
    As = [T(g["a%d" % i]).requires_grad_(True) for i in range(heads)]
    Wo, ao = T(g["W_out"]).requires_grad_(True), T(g["a_out"]).requires_grad_(True)
    out = L.gat_model(x, adj, Ws, As, Wo, ao, float(g["alpha"]), sparse)
    loss = torch.nn.functional.nll_loss(out, torch.as_tensor(g["labels"]).long())
    loss.backward()
    assert rel_err(out.detach(), g["out"]) < 1e-5
    assert rel_err(x.grad, g["g_x"]) < 1e-4
    for i in range(heads):
        assert rel_err(Ws[i].grad, g["g_W%d" % i]) < 1e-4
        assert rel_err(As[i].grad, g["g_a%d" % i]) < 1e-4
    assert rel_err(Wo.grad, g["g_W_out"]) < 1e-4
    # C oracle: first layer (4 heads fused), el/er from the per-head attention vectors
    adjn = g["adj"]
    rows, cols = np.nonzero(adjn)
    rp, col, _ = oracle.coo_to_csr(rows, cols, adjn.shape[0])
    D = g["W0"].shape[1]
    wh = np.concatenate([g["x"] @ g["W%d" % i] for i in range(heads)], axis=1)
    if sparse:
        al = [g["a%d" % i][0, :D] for i in range(heads)]
        ar = [g["a%d" % i][0, D:] for i in range(heads)]
    else:
        al = [g["a%d" % i][:D, 0] for i in range(heads)]
        ar = [g["a%d" % i][D:, 0] for i in range(heads)]
    el = np.stack([wh[:, i * D:(i + 1) * D] @ al[i] for i in range(heads)], axis=1)
    er = np.stack([wh[:, i * D:(i + 1) * D] @ ar[i] for i in range(heads)], axis=1)
    c = oracle.gat_forward(rp, col, wh, el, er, heads, float(g["alpha"]), "exp_neg" if sparse else "softmax", elu=True)
    assert rel_err(c, g["first_layer"]) < 1e-5


def test_gat_dense_model():
    _gat_check("nn_gat_dense", sparse=False)


def test_gat_sparse_model():
    _gat_check("nn_gat_sparse", sparse=True)


def test_special_spmm_and_sddmm():
    g = golden("nn_special_spmm")
    idx = torch.as_tensor(g["indices"])
    gv, gb = L.special_spmm_backward(idx, T(g["values"]), T(g["b"]), T(g["g"]))
    assert rel_err(gv, g["g_values"]) < 1e-5
    assert rel_err(gb, g["g_b"]) < 1e-5
    n = g["b"].shape[0]
    rp, col, val = oracle.coo_to_csr(g["indices"][0], g["indices"][1], n, g["values"])
    y = oracle.spmm_csr(rp, col, g["b"], values=val)
    assert rel_err(y, g["y"]) < 1e-5
    # SDDMM in CSR edge order == grad_values in COO order (indices are already row-major sorted by nonzero())
    sd = oracle.sddmm_csr(rp, col, g["g"], g["b"])
    assert rel_err(sd, g["g_values"]) < 1e-5


def test_gin_model():
    g = golden("nn_gin")
    out = L.gin_model(T(g["A"]), T(g["X"]), g)
    assert rel_err(out, g["out"]) < 1e-5


def test_normalize_helper():
    g = golden("nn_normalize")
    assert rel_err(L.normalize_rows(g["raw"]), g["dense"]) < 1e-6


def test_neighbor_sampler_bit_exact():
    g = golden("sampler_neighbor")
    ptr, flat = g["nbr_ptr"], g["nbr_flat"]
    random.seed(int(g["py_seed"]))
    inp, outp, blocks = S.neighbor_sampler(ptr, flat, g["seeds"], list(g["fanouts"]))
    assert np.array_equal(inp, g["input_nodes"])
    assert np.array_equal(outp, g["output_nodes"])
    for i, (s, d) in enumerate(blocks):
        assert np.array_equal(s, g["b%d_src" % i])
        assert np.array_equal(d, g["b%d_dst" % i])
        assert np.array_equal(S.block_nodes(s, d), g["b%d_nodes" % i])
    adj, nodes = S.get_adj(ptr, flat, blocks)
    assert np.array_equal(adj, g["adj"])
    gathered, _ = oracle.gather_rows(g["feats"], nodes)
    assert np.array_equal(gathered, g["gathered"])
    assert np.array_equal(S.induced_subgraph(ptr, flat, [0, 2, 5, 6, 9, 23]), g["induced"])
    nb = json.loads(bytes(g["neighbors_json"]).decode())
    assert nb == [S.neighbors_of(ptr, flat, v) for v in (0, 2, 5)]
    sel, _ = oracle.gather_rows(g["feats"], g["seeds"])
    assert np.array_equal(sel, g["feats_sel"])


def test_multihop_sampler_bit_exact():
    g = golden("sampler_multihop")
    np.random.seed(int(g["np_seed"]))
    hops = S.multihop_sampling(g["nbr_ptr"], g["nbr_flat"], g["seeds"], list(g["fanouts"]))
    assert np.array_equal(hops[1], g["hop1"])
    assert np.array_equal(hops[2], g["hop2"])


def test_cache_server_split_gather():
    rng = np.random.RandomState(0)
    n = 500
    host = {"features": rng.randn(n, 12).astype(np.float32), "norm": rng.rand(n, 1).astype(np.float32)}
    nid_map = rng.permutation(n)
    cs = S.CacheServer(host, n, nid_map)
    deg = rng.randint(0, 50, size=n)
    cs.auto_cache(deg, 120, ["features", "norm"])
    ids = rng.randint(0, n, size=300)
    fr = cs.fetch(ids)
    assert np.array_equal(fr["features"], host["features"][nid_map[ids]])
    assert np.array_equal(fr["norm"], host["norm"][nid_map[ids]])
    out, miss = oracle.gather_rows(cs.cache["features"], ids, host_table=host["features"], gpu_flag=cs.gpu_flag,
                                   local2cache=cs.localid2cacheid, nid_map=nid_map)
    assert np.array_equal(out, fr["features"])
    assert miss == cs.miss_num
    assert 0.0 < cs.get_miss_rate() < 1.0


def _ref_cache_order(g, tag="part"):
    """The node ids the reference cached, in its cache order (storage.py:139: localid2cacheid[nids] = arange)."""
    flag = g[tag + "_gpu_flag"].astype(bool)
    ids = np.nonzero(flag)[0]
    return ids[np.argsort(g[tag + "_localid2cacheid"][ids])]


def test_cache_server_restatement_pinned_by_the_reference_run():
    """tests/golden/cache_server.npz holds what the reference's own GraphCacheServer did (oracle/gen_golden_pins.py
    executes dgll/FeatureCache/storage.py in place): fill policy, bookkeeping arrays, per-layer frames, miss accounting."""
    g = golden("cache_server")
    host = {"features": g["feats"], "norm": g["norm"]}
    n = g["nid_map"].size
    layers = [g["layer0"], g["layer1"], g["layer2"]]
    # fill policy (:84-98): exactly `capability` nodes, none with a smaller out-degree than an uncached one
    flag = g["part_gpu_flag"].astype(bool)
    assert flag.sum() == g["part_cached_num"] == 150
    assert g["out_deg"][flag].min() >= g["out_deg"][~flag].max()
    cs = S.CacheServer(host, n, g["nid_map"])
    cs.auto_cache(g["out_deg"], 150, ["features", "norm"])
    strict = g["out_deg"] > g["out_deg"][flag].min()             # ties at the boundary are the sort's free choice
    assert np.array_equal(cs.gpu_flag[strict], flag[strict]) and cs.gpu_flag.sum() == 150
    # same cache set and order as the reference -> identical bookkeeping, frames and miss statistics
    cs = S.CacheServer(host, n, g["nid_map"])
    nids = _ref_cache_order(g)
    cs.cache_fix_data(nids, {k: host[k][g["nid_map"][nids]] for k in host})
    assert np.array_equal(cs.gpu_flag, flag)
    assert np.array_equal(cs.localid2cacheid[flag], g["part_localid2cacheid"][flag])
    for name in host:
        assert np.array_equal(cs.cache[name], g["part_cache_" + name])
    for i, ids in enumerate(layers):
        fr = cs.fetch(ids)
        for name in host:
            assert np.array_equal(fr[name], g["part_frame%d_%s" % (i, name)])
    assert cs.try_num == g["part_try_num"] and cs.miss_num == g["part_miss_num"]
    assert cs.get_miss_rate() == float(g["part_miss_rate"])
    # fully cached (:80-83, fetch_from_cache :201-210)
    cf = S.CacheServer(host, n, g["nid_map"])
    cf.auto_cache(g["out_deg"], n, ["features", "norm"])
    assert cf.full_cached and np.array_equal(cf.localid2cacheid, g["full_localid2cacheid"])
    for i, ids in enumerate(layers):
        fr = cf.fetch(ids)
        for name in host:
            assert np.array_equal(fr[name], g["full_frame%d_%s" % (i, name)])


@pytest.mark.parametrize("aggr", ["mean", "sum", "max"])
@pytest.mark.parametrize("combine", ["sum", "concat"])
def test_sage_conv_restatement_pinned_by_the_repaired_reference(aggr, combine):
    """tests/golden/nn_sageconv_fixed.npz: the reference's sageconv.py executed with its two documented one-line
    repairs applied to the AST (oracle/gen_golden_pins.py)."""
    g = golden("nn_sageconv_fixed")
    k = "%s_%s_" % (aggr, combine)
    src = T(g["src"]).requires_grad_(True)
    neigh = T(g["neigh"]).requires_grad_(True)
    ws, wn = T(g[k + "w_self"]).requires_grad_(True), T(g[k + "w_neigh"]).requires_grad_(True)
    out = L.sage_conv(src, neigh, ws, wn, aggr=aggr, combine=combine)
    assert rel_err(out.detach(), g[k + "out"]) < 1e-6
    out.backward(T(g[k + "g"]))
    for t, name in ((src, "d_src"), (neigh, "d_neigh"), (ws, "d_w_self"), (wn, "d_w_neigh")):
        assert rel_err(t.grad, g[k + name]) < 1e-6, name


def test_graphsage_model_restatement_pinned_by_the_repaired_reference():
    g = golden("nn_sageconv_fixed")
    layers = [(T(g["model_l%d_w_self" % i]), T(g["model_l%d_w_neigh" % i])) for i in range(2)]
    out = L.graphsage_model([T(g["model_hop%d" % i]) for i in range(3)], layers, list(g["model_fan"]))
    assert rel_err(out, g["model_out"]) < 1e-6


def test_fused_kernel_restatement_matches_dense():
    rng = np.random.RandomState(3)
    n, f, fp, h = 40, 10, 12, 7
    a = (rng.rand(n, n) < 0.15).astype(np.float64)
    a_hat = L.sym_norm_adjacency(a)
    rows, cols = np.nonzero(a_hat)
    rp, col, val = oracle.coo_to_csr(rows, cols, n, a_hat[rows, cols])
    X = np.zeros((n, fp), dtype=np.float32)
    X[:, :f] = rng.randn(n, f)
    W = rng.randn(fp, h).astype(np.float32)
    H = oracle.gcn_fused_forward(rp, col, val, X, W, np.diff(rp), f)
    ref = L.fused_gcn_layer(T(X[:, :f]), T(a_hat), T(W[:f]))
    assert rel_err(H, ref) < 1e-5


def test_binarized_counts_definition():
    rng = np.random.RandomState(4)
    n, f = 64, 70
    x = rng.randn(n, f).astype(np.float32)
    a = rng.rand(n, n) < 0.2
    rows, cols = np.nonzero(a)
    rp, col, _ = oracle.coo_to_csr(rows, cols, n)
    packed = oracle.binarize_pack(x)
    assert packed.shape[1] == oracle.packed_words(f) == 4
    cnt = oracle.bin_spmm_counts(rp, col, packed, f)
    assert np.array_equal(cnt, a.astype(np.int64) @ (x >= 0).astype(np.int64))
