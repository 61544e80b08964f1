"""Pin oracle/layerwise.py against golden vectors produced by running the reference's own layer-wise sampler class
bodies (oracle/gen_golden_layerwise.py).  CPU-only: runs under -m "not gpu"."""
import numpy as np
import pytest

from oracle import layerwise as LW
from conftest import golden

CASES = {
    # golden file                 kind       flat   wrs    include_batch carry
    "ladies_sym":               ("ladies",  False, True,  False, "global"),
    "ladies_flat_dir":          ("ladies",  True,  True,  False, "global"),
    "ladiesflatwrs_sym":        ("ladies",  True,  True,  False, "global"),
    "fastgcn_sym":              ("fastgcn", False, False, True,  "local"),
    "fastgcnflatwrs_plain_dir": ("fastgcn", False, False, False, "local"),
    "fastgcnflatwrs_flat_sym":  ("fastgcn", True,  False, False, "local"),
    "fastgcnflatwrs_wrs_sym":   ("fastgcn", True,  True,  False, "local"),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_laplacian_bit_exact(name):
    g = golden("layerwise_" + name)
    kind = CASES[name][0]
    ptr, idx, val = LW.laplacian(g["adj_indptr"], g["adj_indices"], "row" if kind == "ladies" else "sym")
    assert np.array_equal(ptr, g["lap_indptr"])
    assert np.array_equal(idx, g["lap_indices"])
    assert np.array_equal(val, g["lap_data"])          # same multiplication order as scipy: bit-identical


@pytest.mark.parametrize("name", sorted(CASES))
def test_sampler_replays_the_reference_rng_stream(name):
    """Seeding numpy as the generator did makes rng.choice draw the same nodes: blocks, picks, WRS weights and the
    returned input nodes are bit-identical to what the reference's class produced."""
    g = golden("layerwise_" + name)
    kind, flat, wrs, include_batch, carry = CASES[name]
    lap = (g["lap_indptr"], g["lap_indices"], g["lap_data"])
    np.random.seed(int(g["np_seed"]))
    layers = LW.layerwise_sample(lap, g["batch"], g["fanouts"].tolist(), kind, flat=flat, wrs=wrs,
                                 include_batch=include_batch, carry=carry, rng=np.random)
    assert len(layers) == int(g["n_layers"])
    for li, lay in enumerate(layers):
        assert np.array_equal(lay["picks"], g["l%d_picks" % li])
        assert np.array_equal(lay["indptr"], g["l%d_indptr" % li])
        assert np.array_equal(lay["indices"], g["l%d_indices" % li])
        if "l%d_weights" % li in g:
            assert np.array_equal(lay["weights"], g["l%d_weights" % li])
    last = layers[-1]
    want = np.arange(int(last["indices"].max()) + 1) if carry == "local" else last["next_nodes"]
    assert np.array_equal(want, g["input_nodes"])


def test_replayed_picks_give_the_same_blocks_without_an_rng():
    g = golden("layerwise_ladies_sym")
    lap = (g["lap_indptr"], g["lap_indices"], g["lap_data"])
    picks = [g["l%d_picks" % li] for li in range(int(g["n_layers"]))]
    layers = LW.layerwise_sample(lap, g["batch"], g["fanouts"].tolist(), "ladies", picks=picks, rng=None)
    for li, lay in enumerate(layers):
        assert np.array_equal(lay["indices"], g["l%d_indices" % li])
        # importance-weighted values: row-normalised Laplacian entry times the estimator weight of its column
        assert lay["data"].shape == lay["indices"].shape and np.all(lay["data"] > 0)


def test_fanout_larger_than_candidates_clamps():
    g = golden("layerwise_ladies_flat_dir")
    assert int(g["fanouts"][1]) == 500 and len(g["l1_picks"]) < 500   # s_num = min(#(prob > 0), fanout)


def test_wrs_estimator_known_answer():
    g = golden("layerwise_estwrs")
    w = LW.est_wrs_weights(g["p"], g["idx"])
    assert np.array_equal(w, g["w"])
    np.random.seed(int(g["np_seed"]))
    assert np.array_equal(np.random.choice(len(g["p"]), int(g["m"]), False, g["p"]), g["idx"])


def test_restatement_equals_the_scipy_pipeline_on_a_larger_graph():
    """Independent of the fixtures: the plain-CSR restatement against the scipy calls the scripts make
    (``lap[rows, :]``, ``Q.multiply(Q).sum(0)``, ``Q[:, picks].multiply(w).tocsr()``) on a 3,000-node graph, same
    numpy stream — picks, block structure and values identical."""
    sp = pytest.importorskip("scipy.sparse")
    rng = np.random.default_rng(12)
    n = 3000
    src = rng.integers(0, n, 40000)
    dst = (src + 1 + rng.zipf(1.4, 40000) % (n - 1)) % n
    a = sp.coo_matrix((np.ones(40000), (src, dst)), shape=(n, n)).tocsr()
    a.data[:] = 1.0
    a = ((a + a.T) > 0).astype(np.float64).tocsr()
    a.sort_indices()
    lap = LW.laplacian(a.indptr.astype(np.int64), a.indices.astype(np.int64), "row")
    lap_sp = sp.csr_matrix((lap[2], lap[1], lap[0]), shape=(n, n))
    batch = rng.choice(n, 200, replace=False)
    for flat in (False, True):
        np.random.seed(5)
        want = LW.scipy_ladies_batch(lap_sp, batch, [128, 256], flat=flat)
        np.random.seed(5)
        got = LW.layerwise_sample(lap, batch, [128, 256], "ladies", flat=flat)
        for (w_ptr, w_idx, w_val, w_picks), g in zip(want, got):
            assert np.array_equal(g["picks"], w_picks)
            assert np.array_equal(g["indptr"], w_ptr) and np.array_equal(g["indices"], w_idx)
            assert np.array_equal(g["data"], w_val)
