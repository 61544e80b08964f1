"""GPU parity tests for the layer-wise importance samplers: the device pipeline (csrc/layerwise.cu
through the C ABI) against oracle/layerwise.py and the fixtures produced by the reference's own class bodies
(tests/golden/layerwise_*.npz).  Index work bit-exact; fp64 probabilities / weights / edge values <= 1e-12 relative."""
import itertools

import numpy as np
import pytest
import torch

from oracle import layerwise as LW
from conftest import golden, rel_err

pytestmark = pytest.mark.gpu

F64_TOL = 1e-12

CASES = {
    # golden file                 class                    kwargs                        oracle (kind, flat, wrs, include_batch)
    "ladies_sym":               ("Ladies",                {},                            ("ladies", False, True, False), "global"),
    "ladies_flat_dir":          ("LadiesFlat",            {"flat": True},                ("ladies", True, True, False), "global"),
    "ladiesflatwrs_sym":        ("LadiesFlatWrs",         {"flat": True},                ("ladies", True, True, False), "global"),
    "fastgcn_sym":              ("FastGCNSampler",        {},                            ("fastgcn", False, False, True), "local"),
    "fastgcnflatwrs_plain_dir": ("FastGCNSamplerFlatWrs", {},                            ("fastgcn", False, False, False), "local"),
    "fastgcnflatwrs_flat_sym":  ("FastGCNSamplerFlat",    {"flat": True},                ("fastgcn", True, False, False), "local"),
    "fastgcnflatwrs_wrs_sym":   ("FastGCNSamplerFlatWrs", {"flat": True, "wrs": True},   ("fastgcn", True, True, False), "local"),
}


@pytest.fixture(scope="module")
def K():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from dgll_b200 import kernels
    return kernels


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def make_sampler(name, **extra):
    import dgll_b200.data as D
    g = golden("layerwise_" + name)
    cls, kw, _, carry = CASES[name]
    adj = (dev(g["adj_indptr"]), dev(g["adj_indices"].astype(np.int32)))
    return g, getattr(D, cls)(g["fanouts"].tolist(), adj, carry=carry, **kw, **extra)


@pytest.mark.parametrize("name", sorted(CASES))
def test_laplacian_bit_exact(K, name):
    g, s = make_sampler(name)
    assert np.array_equal(s.lap_rp.cpu().numpy(), g["lap_indptr"])
    assert np.array_equal(s.lap_col.cpu().numpy(), g["lap_indices"])
    assert np.array_equal(s.lap_val.cpu().numpy(), g["lap_data"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_replayed_draws_give_the_reference_blocks(K, name):
    """With the reference's recorded draws replayed, every layer's block (indptr, indices), the nodes handed to the
    next layer and the returned input nodes are bit-identical to what the reference class produced; weights and the
    importance-scaled edge values match the oracle to fp64 round-off."""
    g, s = make_sampler(name)
    n_layers = int(g["n_layers"])
    replay = [g["l%d_picks" % li] for li in range(n_layers)]
    input_nodes, batch, blocks = s.sample(None, g["batch"], replay=replay)
    kind, flat, wrs, include_batch = CASES[name][2]
    lap = (g["lap_indptr"], g["lap_indices"], g["lap_data"])
    ref = LW.layerwise_sample(lap, g["batch"], g["fanouts"].tolist(), kind, flat=flat, wrs=wrs,
                              include_batch=include_batch, carry=CASES[name][3], picks=replay, rng=None)
    assert np.array_equal(batch.cpu().numpy(), g["batch"])
    assert np.array_equal(input_nodes.cpu().numpy(), g["input_nodes"])
    for li in range(n_layers):
        blk = blocks[n_layers - 1 - li]                  # golden / oracle order: output layer first
        nnz = len(g["l%d_indices" % li])
        assert np.array_equal(blk.row_ptr.cpu().numpy(), g["l%d_indptr" % li])
        assert blk.col.numel() == nnz
        assert np.array_equal(blk.col.cpu().numpy().astype(np.int64), g["l%d_indices" % li])
        assert np.array_equal(blk.picks.cpu().numpy(), g["l%d_picks" % li])
        assert np.array_equal(blk.src_ids.cpu().numpy(), ref[li]["next_nodes"])
        assert np.array_equal(blk.col_global.cpu().numpy().astype(np.int64), ref[li]["next_nodes"][ref[li]["indices"]])
        if "l%d_weights" % li in g:
            assert rel_err(blk.weights.cpu().numpy(), g["l%d_weights" % li]) <= F64_TOL
        assert rel_err(blk.edge_weight64.cpu().numpy(), ref[li]["data"]) <= F64_TOL
        # probabilities of the layer
        cc, cp = blk.prob
        if cc is None:
            assert rel_err(cp.cpu().numpy(), ref[li]["prob"]) <= F64_TOL
        else:
            dense = np.zeros(len(ref[li]["prob"]))
            ccn = cc.cpu().numpy()
            keep = ccn >= 0
            dense[ccn[keep]] = cp.cpu().numpy()[keep]
            assert rel_err(dense, ref[li]["prob"]) <= F64_TOL
            assert np.array_equal(dense > 0, ref[li]["prob"] > 0)


def test_slice_rows_and_column_sums_bit_exact(K):
    g = golden("layerwise_ladies_sym")
    lap = (g["lap_indptr"], g["lap_indices"], g["lap_data"])
    rows = np.concatenate([g["batch"], g["batch"][:3], [0, 399]])          # duplicates allowed
    q_ptr, q_idx, q_val = LW.slice_rows(*lap, rows)
    rp, col, val = K.csr_slice_rows(dev(lap[0]), dev(lap[1].astype(np.int32)), dev(lap[2]), dev(rows))
    assert np.array_equal(rp.cpu().numpy(), q_ptr)
    assert np.array_equal(col.cpu().numpy().astype(np.int64), q_idx)
    assert np.array_equal(val.cpu().numpy(), q_val)
    for flat in (False, True):
        cc, cp, stats = K.col_sqsum(col, val, 400, flat)
        n_cand, total, n_pos = stats.tolist()
        prob_i = LW.column_sq_sums(q_idx, q_val, 400)
        if flat:
            prob_i = np.sqrt(prob_i)
        cols = np.nonzero(prob_i)[0]
        assert int(n_cand) == len(cols) == int(n_pos)
        assert np.array_equal(cc[:len(cols)].cpu().numpy().astype(np.int64), cols)
        assert (cc[len(cols):] == -1).all() and (cp[len(cols):] == 0).all()
        # un-normalise with the device's own total: the per-column sums themselves are bit-identical to scipy's order
        assert abs(total - prob_i.sum()) <= 1e-13 * prob_i.sum()
        assert rel_err(cp[:len(cols)].cpu().numpy(), prob_i[cols] / prob_i.sum()) <= 1e-14
    # empty slice
    rp0, col0, val0 = K.csr_slice_rows(dev(lap[0]), dev(lap[1].astype(np.int32)), dev(lap[2]),
                                       torch.zeros(0, dtype=torch.int64, device="cuda"))
    assert rp0.tolist() == [0] and col0.numel() == 0
    _, _, st0 = K.col_sqsum(col0, val0, 400)
    assert st0.tolist() == [0.0, 0.0, 0.0]


def test_column_sums_are_bitwise_those_of_the_sequential_order(K):
    """One long column: 3,000 entries summed in stored order, products and sums rounded separately (no FMA)."""
    rng = np.random.default_rng(0)
    col = np.concatenate([np.full(3000, 7), rng.integers(0, 50, 2000)]).astype(np.int32)
    perm = rng.permutation(col.size)
    col = col[perm]
    val = rng.standard_normal(col.size)
    want = LW.column_sq_sums(col.astype(np.int64), val, 50)
    cc, cp, stats = K.col_sqsum(dev(col), dev(val), 50)
    n_cand = int(stats[0].item())
    got = (cp[:n_cand] * stats[1]).cpu().numpy()       # p * total: not exact in general, so compare the ratio instead
    assert rel_err(got, want[cc[:n_cand].cpu().numpy()]) <= 1e-15
    a = K.col_sqsum(dev(col), dev(val), 50)[1]
    assert torch.equal(a, cp)                            # run-to-run bit-identical


def test_select_columns_matches_oracle(K):
    rng = np.random.default_rng(5)
    g = golden("layerwise_ladies_flat_dir")
    lap = (g["lap_indptr"], g["lap_indices"], g["lap_data"])
    rows = rng.choice(350, 60, replace=False)
    q_ptr, q_idx, q_val = LW.slice_rows(*lap, rows)
    picks = rng.permutation(np.unique(q_idx))[:40]                        # unsorted picks
    picks = np.concatenate([picks, [int(np.setdiff1d(np.arange(350), q_idx)[0])]])   # one column absent from Q
    scale = rng.random(len(picks)) + 0.5
    want = LW.select_columns(q_ptr, q_idx, q_val, picks, scale)
    pos = torch.full((350,), -1, dtype=torch.int32, device="cuda")
    K.scatter_pos(pos, dev(picks))
    cap = len(q_idx) + 17                                                  # capacity larger than nnz(Q)
    qc = torch.zeros(cap, dtype=torch.int32, device="cuda")
    qv = torch.zeros(cap, dtype=torch.float64, device="cuda")
    qc[:len(q_idx)] = dev(q_idx.astype(np.int32))
    qv[:len(q_idx)] = dev(q_val)
    rp, col, val = K.csr_select_cols(dev(q_ptr), qc, qv, pos, dev(scale))
    nnz = int(rp[-1].item())
    assert np.array_equal(rp.cpu().numpy(), want[0])
    assert np.array_equal(col[:nnz].cpu().numpy().astype(np.int64), want[1])
    assert np.array_equal(val[:nnz].cpu().numpy(), want[2])               # one rounded multiply: bit-exact
    rp2, col2, val2 = K.csr_select_cols(dev(q_ptr), qc, None, pos, None, with_values=False)   # topology only
    assert val2 is None and np.array_equal(col2[:nnz].cpu().numpy().astype(np.int64), want[1])
    K.scatter_pos(pos, dev(picks), reset=True)
    assert (pos == -1).all()


def test_wrs_weights_bit_exact_given_the_same_probabilities(K):
    g = golden("layerwise_estwrs")
    p, idx = g["p"], g["idx"]
    scale = K.importance_scale(dev(p), dev(idx.astype(np.int32)), torch.tensor([len(idx)], device="cuda"), len(p), "wrs")
    assert np.array_equal(scale.cpu().numpy(), g["w"])                    # same operations, same order, no FMA
    cap = np.concatenate([idx, [-1, -1, -1]]).astype(np.int32)            # capacity beyond the count: zeros
    scale = K.importance_scale(dev(p), dev(cap), torch.tensor([len(idx)], device="cuda"), len(p), "wrs")
    assert np.array_equal(scale[:len(idx)].cpu().numpy(), g["w"]) and (scale[len(idx):] == 0).all()
    inv = K.importance_scale(dev(p), dev(idx.astype(np.int32)), torch.tensor([len(idx)], device="cuda"), len(p), "inverse")
    assert np.array_equal(inv.cpu().numpy(), 1 / p[idx] / len(idx))


def test_device_draw_is_a_valid_weighted_sample_without_replacement(K):
    rng = np.random.default_rng(2)
    p = rng.random(500)
    p[rng.random(500) < 0.4] = 0
    p /= p.sum()
    n_pos = int((p > 0).sum())
    for fanout in (1, 64, n_pos, n_pos + 50):
        sel, picks, count = K.weighted_choice(None, dev(p), fanout, seed=11)
        m = int(count.item())
        assert m == min(fanout, n_pos)
        got = picks[:m].cpu().numpy()
        assert len(np.unique(got)) == m and (p[got] > 0).all()
        assert (picks[m:] == -1).all() and (sel[m:] == -1).all()
        again = K.weighted_choice(None, dev(p), fanout, seed=11)[1]
        assert torch.equal(again, picks)
    other = K.weighted_choice(None, dev(p), 64, seed=12)[1]
    assert not torch.equal(other, K.weighted_choice(None, dev(p), 64, seed=11)[1])
    # candidate form: same nodes drawn whatever the candidate order (generator keyed by node id)
    cols = np.nonzero(p)[0].astype(np.int32)
    perm = rng.permutation(len(cols))
    a = K.weighted_choice(dev(cols), dev(p[cols]), 64, seed=11)[1]
    b = K.weighted_choice(dev(cols[perm]), dev(p[cols][perm]), 64, seed=11)[1]
    c = K.weighted_choice(None, dev(p), 64, seed=11)[1]
    assert torch.equal(a, b) and torch.equal(a, c)


def test_device_draw_follows_the_sequential_weighted_distribution(K):
    """choice(n, m, replace=False, p) draws sequentially, renormalising after each draw.  For n=5, m=2 the exact law
    of the ordered pair is p_a * p_b / (1 - p_a); 20,000 seeds reproduce it within 5 sigma for every pair."""
    p = np.array([0.4, 0.25, 0.2, 0.1, 0.05])
    trials = 20000
    counts = np.zeros((5, 5))
    pd = dev(p)
    for s in range(trials):
        a, b = K.weighted_choice(None, pd, 2, seed=s)[1].tolist()
        counts[a, b] += 1
    for a, b in itertools.permutations(range(5), 2):
        q = p[a] * p[b] / (1 - p[a])
        sigma = np.sqrt(trials * q * (1 - q))
        assert abs(counts[a, b] - trials * q) <= 5 * sigma, (a, b, counts[a, b], trials * q)


@pytest.mark.parametrize("cls,kw", [("Ladies", {}), ("LadiesFlatWrs", {"flat": True}),
                                    ("FastGCNSamplerFlat", {"flat": True}), ("FastGCNSamplerWrs", {"wrs": True}),
                                    ("FastGCNSampler", {})])
def test_device_sampler_end_to_end_properties(K, cls, kw):
    """Device draw in the loop: block structure is consistent with the Laplacian, deterministic per seed, and the
    blocks drive the GraphConv layers (forward + backward)."""
    import dgll_b200.data as D
    import dgll_b200.nn as dnn
    g = golden("layerwise_ladies_sym")
    n = len(g["adj_indptr"]) - 1
    adj = (dev(g["adj_indptr"]), dev(g["adj_indices"].astype(np.int32)))
    s = getattr(D, cls)([32, 64], adj, rng_seed=3, **kw)
    batch = dev(g["batch"])
    inp, out, blocks = s.sample(None, batch)
    s2 = getattr(D, cls)([32, 64], adj, rng_seed=3, **kw)
    inp2, _, blocks2 = s2.sample(None, batch)
    assert torch.equal(inp, inp2) and all(torch.equal(a.col, b.col) for a, b in zip(blocks, blocks2))
    lap = (s.lap_rp.cpu().numpy(), s.lap_col.cpu().numpy().astype(np.int64), s.lap_val.cpu().numpy())
    rows = g["batch"]
    for blk in reversed(blocks):
        nxt = blk.src_ids.cpu().numpy()
        assert len(np.unique(nxt)) == len(nxt) and blk.num_dst_nodes() == len(rows)
        if cls != "FastGCNSampler":
            assert len(nxt) <= blk.picks.numel() and np.array_equal(nxt, blk.picks.cpu().numpy())
        else:
            assert np.isin(g["batch"], nxt).all()         # the batch is unioned in
        q = LW.slice_rows(*lap, rows)
        scale = np.ones(len(nxt))
        want = LW.select_columns(*q, nxt, scale)
        assert np.array_equal(blk.row_ptr.cpu().numpy(), want[0])
        assert np.array_equal(blk.col.cpu().numpy().astype(np.int64), want[1])
        rows = nxt
    assert torch.equal(inp, blocks[0].src_ids)
    # the blocks feed the block layers exactly like neighbour-sampled ones (MQLadies.py:49-60)
    torch.manual_seed(0)
    c1, c2 = dnn.GraphConv(12, 16).cuda(), dnn.GraphConv(16, 5).cuda()
    x = torch.randn(n, 12, device="cuda")
    h = torch.relu(c1(blocks[0], x[inp]))
    h = c2(blocks[1], h)
    assert h.shape == (len(g["batch"]), 5) and torch.isfinite(h).all()
    h.sum().backward()
    assert torch.isfinite(c1.weight.grad).all() and c1.weight.grad.abs().sum() > 0
    # importance-weighted aggregation (the weights the reference computes and drops): matches a dense product
    hw = c2(blocks[1], torch.relu(c1(blocks[0], x[inp], edge_weight=blocks[0].edge_weight)),
            edge_weight=blocks[1].edge_weight)
    assert torch.isfinite(hw).all()


def test_numpy_chooser_replays_the_reference_stream(K):
    """chooser=numpy_chooser draws with np.random.choice on the host exactly as the scripts do: seeding numpy like the
    fixture generator reproduces the reference's picks (probabilities agree to the last bits that matter)."""
    import dgll_b200.data as D
    g, s = make_sampler("ladies_sym", chooser=D.numpy_chooser)
    np.random.seed(int(g["np_seed"]))
    _, _, blocks = s.sample(None, g["batch"])
    for li in range(int(g["n_layers"])):
        blk = blocks[int(g["n_layers"]) - 1 - li]
        assert np.array_equal(blk.picks.cpu().numpy(), g["l%d_picks" % li])
        assert np.array_equal(blk.col.cpu().numpy().astype(np.int64), g["l%d_indices" % li])
