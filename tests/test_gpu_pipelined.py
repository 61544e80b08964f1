"""The one-graph-per-mini-batch trainer (dgll_b200.pipelined): sampler, block builder, input-layer aggregation straight
from the (resident or node-range-partitioned) feature table, training step, flat gradient all-reduce and Adam inside two
ping-pong CUDA graphs.  It must take the same steps as the Python-dispatched epoch (train.sage_epoch) that samples
the same blocks."""
import copy

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _setup(N=20000, F=100, n_seeds=512 * 4 + 100, classes=7):
    import dgll_b200.nn as nn
    from dgll_b200 import graphs as G
    rp, col = G.rmat_csr(N, N * 20, seed=1, device="cuda")
    table = G.feature_table(N, F, seed=2)
    gen = torch.Generator(device="cuda").manual_seed(3)
    labels = torch.randint(0, classes, (N,), device="cuda", generator=gen)
    seeds = torch.randperm(N, device="cuda", generator=gen)[:n_seeds]
    torch.manual_seed(0)
    model = nn.GraphSAGE(F, 64, classes, 2, torch.relu, 0.0).cuda()
    return rp, col, table, labels, seeds, model


@pytest.mark.parametrize("source", ["table", "sharded"])
def test_pipelined_epoch_equals_eager_epoch(source):
    from dgll_b200 import parallel as P, pipelined as PL, train as T
    rp, col, table, labels, seeds, m1 = _setup()
    F = 100
    m2 = copy.deepcopy(m1)
    o1 = torch.optim.Adam(m1.parameters(), lr=0.01, fused=True)
    o2 = torch.optim.Adam(m2.parameters(), lr=0.01, fused=True, capturable=True)
    a = T.sage_epoch(m1, o1, table, labels, F, rp, col, seeds, (10, 5), 512, rng_seed=6, precision="fp32")
    kw = {"table": table} if source == "table" else {"sharded": P.PeerShardedTable(table.size(0), table)}
    tr = PL.PipelinedSageTrainer(m2, o2, labels, rp, col, F, batch_size=512, fanouts=(10, 5), precision="fp32",
                                 rng_seed=6, **kw)
    b = tr.epoch(seeds)
    assert a["n_batches"] == b["n_batches"] == 5
    assert abs(a["loss"] - b["loss"]) <= 1e-5 * abs(a["loss"])
    for p, q in zip(m1.parameters(), m2.parameters()):
        assert rel_err(q.detach().cpu().numpy(), p.detach().cpu().numpy()) <= 2e-5
    # a second epoch replays the same graphs on a new seed list (shorter: the seed table keeps its capacity)
    seeds2 = seeds.flip(0)[:512 * 3]
    a2 = T.sage_epoch(m1, o1, table, labels, F, rp, col, seeds2, (10, 5), 512, rng_seed=6, precision="fp32")
    b2 = tr.epoch(seeds2)
    assert a2["n_batches"] == b2["n_batches"] == 3
    assert abs(a2["loss"] - b2["loss"]) <= 2e-5 * abs(a2["loss"])
    for p, q in zip(m1.parameters(), m2.parameters()):
        assert rel_err(q.detach().cpu().numpy(), p.detach().cpu().numpy()) <= 5e-5


def test_pipelined_table_and_one_shard_are_bit_identical():
    """The sharded aggregation with a single shard does the arithmetic of the resident-table kernel."""
    from dgll_b200 import parallel as P, pipelined as PL
    rp, col, table, labels, seeds, m1 = _setup(n_seeds=512 * 3)
    m2 = copy.deepcopy(m1)
    res = []
    for m, kw in ((m1, {"table": table}), (m2, {"sharded": P.PeerShardedTable(table.size(0), table)})):
        o = torch.optim.Adam(m.parameters(), lr=0.01, fused=True, capturable=True)
        tr = PL.PipelinedSageTrainer(m, o, labels, rp, col, 100, batch_size=512, fanouts=(10, 5), precision="bf16",
                                     rng_seed=1, **kw)
        res.append(tr.epoch(seeds))
        st = tr.halo_stats()
        assert st["block0_edges"] > 0 and st["remote_edge_fraction"] == 0.0
    assert res[0]["loss"] == res[1]["loss"]
    for p, q in zip(m1.parameters(), m2.parameters()):
        assert torch.equal(p, q)


def test_pipelined_bf16_table_storage():
    """bf16 storage of the feature table (half the bytes of the input-layer aggregation): same training run within the
    bf16 tolerance of the stored features."""
    from dgll_b200 import pipelined as PL
    rp, col, table, labels, seeds, m1 = _setup(n_seeds=512 * 3)
    m2 = copy.deepcopy(m1)
    t16 = torch.zeros((table.size(0), 104), dtype=torch.bfloat16, device="cuda")
    t16[:, :100] = table[:, :100].to(torch.bfloat16)
    out = []
    for m, t in ((m1, table), (m2, t16)):
        o = torch.optim.Adam(m.parameters(), lr=0.01, fused=True, capturable=True)
        tr = PL.PipelinedSageTrainer(m, o, labels, rp, col, 100, table=t, batch_size=512, fanouts=(10, 5),
                                     precision="fp32", rng_seed=1)
        out.append(tr.epoch(seeds)["loss"])
    assert abs(out[0] - out[1]) <= 1e-2 * abs(out[0])
