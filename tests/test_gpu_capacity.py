"""Fixed-capacity sampler / block builder (negative ids = padding, nothing read back) and the training step with the
sampler captured inside the CUDA graph (verified on the B200 at the start of round 2)."""
import copy

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from dgll_b200 import kernels
    return kernels


def test_capacity_sampler_and_builder_equal_the_exact_ones_on_the_valid_prefix(K):
    from dgll_b200 import graphs as G
    N = 5000
    rp, col = G.rmat_csr(N, N * 30, seed=3, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(1)
    seeds = torch.randperm(N, device="cuda", generator=gen)[:300]
    fan1, fan0 = 5, 8
    # exact path
    b_rp, b_col = K.sample_neighbors(rp, col, seeds, fan1, rng_seed=77)
    src_cap, col_cap, counts = K.build_block(seeds, b_rp, b_col)
    num_src, nnz = counts.tolist()
    # capacity path: 300 real seeds in 384 slots, seed offset in device memory
    cap = 384
    seeds_cap = torch.full((cap,), -1, dtype=torch.int64, device="cuda")
    seeds_cap[:300] = seeds
    off = torch.tensor([7], dtype=torch.int64, device="cuda")
    c_rp, c_col = K.sample_neighbors_cap(rp, col, seeds_cap, fan1, rng_seed=70, rng_offset=off)
    assert torch.equal(c_rp[:301], b_rp) and bool((c_rp[301:] == c_rp[300]).all())
    assert torch.equal(c_col[:nnz], b_col[:nnz])
    pad = cap * (1 + fan1)
    s_ids, c_loc, cnt = K.build_block_cap(seeds_cap, c_rp, c_col, col_pad=pad)
    assert cnt.tolist() == [num_src, nnz, 300]
    assert torch.equal(s_ids[:num_src], src_cap[:num_src]) and bool((s_ids[num_src:] == -1).all())
    assert torch.equal(c_loc[:nnz], col_cap[:nnz]) and bool((c_loc[nnz:] == pad).all())
    # the padded src array is the next layer's seed array as it is
    n_rp, n_col = K.sample_neighbors_cap(rp, col, s_ids, fan0, rng_seed=71, rng_offset=off)
    e_rp, e_col = K.sample_neighbors(rp, col, s_ids[:num_src].contiguous(), fan0, rng_seed=78)
    assert torch.equal(n_rp[:num_src + 1], e_rp) and bool((n_rp[num_src + 1:] == n_rp[num_src]).all())
    assert torch.equal(n_col[:int(e_rp[-1])], e_col[:int(e_rp[-1])])


@pytest.mark.parametrize("optimizer", ["sgd_eager_step", "adam_captured"])
def test_step_with_the_sampler_inside_the_graph_equals_the_eager_epoch(optimizer):
    import dgll_b200.nn as nn
    from dgll_b200 import graphs as G, train as T
    N, F = 20000, 100
    rp, col = G.rmat_csr(N, N * 20, seed=1, device="cuda")
    table = G.feature_table(N, F, seed=2)
    gen = torch.Generator(device="cuda").manual_seed(3)
    labels = torch.randint(0, 7, (N,), device="cuda", generator=gen)
    seeds = torch.randperm(N, device="cuda", generator=gen)[:512 * 4 + 100]
    torch.manual_seed(0)
    m1 = nn.GraphSAGE(F, 64, 7, 2, torch.relu, 0.0).cuda()
    m2 = copy.deepcopy(m1)
    if optimizer == "adam_captured":
        o1 = torch.optim.Adam(m1.parameters(), lr=0.01, fused=True)
        o2 = torch.optim.Adam(m2.parameters(), lr=0.01, fused=True, capturable=True)
    else:
        o1 = torch.optim.SGD(m1.parameters(), lr=0.05)
        o2 = torch.optim.SGD(m2.parameters(), lr=0.05)
    a = T.sage_epoch(m1, o1, table, labels, F, rp, col, seeds, (10, 5), 512, rng_seed=6, precision="fp32")
    tr = T.GraphedSageTrainer(m2, o2, table, labels, 512, (10, 5), precision="fp32")
    tr.enable_device_sampler(rp, col, rng_seed=6)
    b = tr.epoch_sampled(seeds)
    assert a["n_batches"] == b["n_batches"] == 5
    assert abs(a["loss"] - b["loss"]) <= 1e-5 * abs(a["loss"])
    for p, q in zip(m1.parameters(), m2.parameters()):
        assert rel_err(q.detach().cpu().numpy(), p.detach().cpu().numpy()) <= 2e-5
