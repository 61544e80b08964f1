#!/usr/bin/env python
"""BASELINE.json configs[2]: GAT 3-layer, 4 heads, full-batch on a synthetic ogbn-products-shaped graph
(N=2,449,029, symmetrised nnz=123,718,280 + self loops, F=100 -> 4x64 -> 4x64 -> 47), fused SDDMM + edge-softmax +
aggregation kernels, tcgen05 bf16 dense transforms.

  python tools/bench_gat_model.py [--steps 5] [--definition softmax|exp_neg] [--precision bf16|fp32]

The layers are the reference's ``sparseGatConv`` / ``gatConv`` modules (dgll/nn/Convolution/gatconv.py:10-57, :89-151)
stacked the way ``SpGAT`` / ``GAT`` stack them (:154-199: nheads modules concatenated, ELU), one more hidden layer as
configs[2] asks, every multi-head layer run as ONE fused kernel launch (nn.conv._MultiHeadMixin).  The reference
itself cannot run this size: its dense layer needs an N x N matrix and its sparse layer materialises [E, 2D]
(SURVEY.md §8 a9/a10).  Prints one JSON line: ms per full-batch training step (forward + loss + backward + Adam) and
per forward, edges/s, and the share of the step spent in the fused aggregation kernels."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import dgll_b200.nn as dnn  # noqa: E402
from dgll_b200 import graphs as G, ops  # noqa: E402
from dgll_b200.nn.conv import _MultiHeadMixin, gatConv, sparseGatConv  # noqa: E402


class HeadsLayer(torch.nn.Module, _MultiHeadMixin):
    def __init__(self, fin, fout, heads, definition):
        super().__init__()
        cls = gatConv if definition == "softmax" else sparseGatConv
        self.attentions = torch.nn.ModuleList([cls(fin, fout, dropout=0.0, alpha=0.2, concat=True) for _ in range(heads)])
        self.mode = definition

    def forward(self, x, adj):
        return self._heads_forward(x, adj, self.mode)


class GAT3(torch.nn.Module):
    def __init__(self, fin, hid, n_cls, heads, definition):
        super().__init__()
        cls = gatConv if definition == "softmax" else sparseGatConv
        self.l0 = HeadsLayer(fin, hid, heads, definition)
        self.l1 = HeadsLayer(hid * heads, hid, heads, definition)
        self.out = cls(hid * heads, n_cls, dropout=0.0, alpha=0.2, concat=False)

    def forward(self, x, adj):
        return self.out(self.l1(self.l0(x, adj), adj), adj)


def cpu_baseline(adj, x, Np, nnz, slices):
    """The reference's sparse GAT layer on the host (gatconv.py:111-148: h W, [E, 2D] edge features, exp(-leakyrelu),
    row sums, weighted sum, ELU; forward + backward through torch autograd, all host threads) for ONE 4-head hidden layer,
    on the destination rows of a 1/`slices` node range with all their in-edges; the full-graph, three-layer figure is
    extrapolated (x slices x 3 layers, the 47-class layer counted like a hidden one).  The reference's own backward
    (SpecialSpmmFunction, :76) is a dense N x N product and cannot run at this size at all."""
    import os
    import time
    torch.set_num_threads(os.cpu_count() or 1)
    n_dst = Np // slices
    rp = adj.row_ptr[:n_dst + 1].cpu()
    e = int(rp[-1])
    cols = adj.col[:e].long().cpu()
    rows = torch.repeat_interleave(torch.arange(n_dst), rp[1:] - rp[:-1])
    src_ids, inv = torch.unique(cols, return_inverse=True)           # the slice's source nodes, compact
    h = x[src_ids.to(x.device)].cpu()
    h_dst = x[:n_dst].cpu()
    heads, D = 4, 64
    Ws = [torch.randn(h.size(1), D, requires_grad=True) for _ in range(heads)]
    As = [torch.randn(1, 2 * D, requires_grad=True) for _ in range(heads)]

    def layer():
        outs = []
        for W, a in zip(Ws, As):
            hw_s, hw_d = h @ W, h_dst @ W
            edge_h = torch.cat((hw_d[rows], hw_s[inv]), dim=1).t()                      # [2D, E] as gatconv.py:121
            edge_e = torch.exp(-torch.nn.functional.leaky_relu(a.mm(edge_h).squeeze(), 0.2))
            rowsum = torch.zeros(n_dst, 1).index_add_(0, rows, edge_e[:, None])
            hp = torch.zeros(n_dst, D).index_add_(0, rows, edge_e[:, None] * hw_s[inv])
            outs.append(torch.nn.functional.elu(hp / rowsum.clamp(min=1e-30)))
        return torch.cat(outs, dim=1)

    layer().sum().backward()                                                             # warm-up
    t0 = time.perf_counter()
    out = layer()
    t1 = time.perf_counter()
    out.sum().backward()
    t2 = time.perf_counter()
    fwd, bwd = t1 - t0, t2 - t1
    return {"kind": "port", "cores": os.cpu_count(), "slice": "1/%d of the destination rows: %d rows, %d edges, %d source rows" % (slices, n_dst, e, src_ids.numel()),
            "layer_forward_s_on_slice": round(fwd, 3), "layer_backward_s_on_slice": round(bwd, 3),
            "training_step_s_extrapolated_full_graph_3_layers": round((fwd + bwd) * slices * 3, 1),
            "forward_s_extrapolated_full_graph_3_layers": round(fwd * slices * 3, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--definition", default="exp_neg", choices=["softmax", "exp_neg"])
    ap.add_argument("--precision", default="tf32", choices=["tf32", "bf16", "fp32"])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--cpu-slice", type=int, default=64,
                    help="time the reference CPU layer (gatconv.py:111-148 restated on an edge list) on a 1/N node-range "
                         "slice of the destination rows and extrapolate (SURVEY.md §8 d-2); 0 = skip")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    Np, E, Fin, C = G.SHAPES["products"]
    Np, E = int(Np * args.scale), int(E * args.scale)
    rp, col = G.rmat_csr(Np, 2 * E, seed=2, device=dev, symmetric=True)
    # + self loops (SURVEY §8 d), merged into the CSR once
    rows = torch.repeat_interleave(torch.arange(Np, device=dev), rp[1:] - rp[:-1])
    key = torch.unique(torch.cat([rows * Np + col.long(), torch.arange(Np, device=dev) * (Np + 1)]))
    del rows
    rp2 = torch.zeros(Np + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.bincount(key // Np, minlength=Np), 0, out=rp2[1:])
    adj = ops.CsrGraph(rp2, (key % Np).to(torch.int32), n_src=Np)
    nnz = int(adj.col.numel())
    del key, rp, col
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn((Np, Fin), device=dev, generator=g)
    y = torch.randint(0, C, (Np,), device=dev, generator=g)
    torch.manual_seed(0)
    model = GAT3(Fin, 64, C, 4, args.definition).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=0.005, fused=True)
    ops.set_gemm_precision(args.precision)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(model(x, adj), y)
        loss.backward()
        opt.step()
        return loss

    for _ in range(2):
        step()                                        # warm-up: plans, transposes, allocator
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    for i in range(args.steps):
        loss = step()
        evs[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps))
    ms_step = ts[len(ts) // 2]
    model.eval()
    with torch.no_grad():
        model(x, adj)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            model(x, adj)
        f1.record()
    torch.cuda.synchronize()
    ms_fwd = f0.elapsed_time(f1) / args.steps
    # algorithmic bytes of the three fused aggregations, forward (SURVEY §8 d formula)
    def gat_bytes(H, D):
        return nnz * (4 + H * D * 4 + H * 4) + Np * (H * D * 4 + H * 4 + 8)
    fwd_bytes = 2 * gat_bytes(4, 64) + gat_bytes(1, C)
    cpu = None
    if args.cpu_slice > 0:
        cpu = cpu_baseline(adj, x, Np, nnz, args.cpu_slice)
    print(json.dumps({
        "cpu_baseline_extrapolated": cpu,
        "workload": "GAT 3-layer 4 heads x 64 -> 47, full batch, products-shaped (BASELINE configs[2])",
        "N": Np, "nnz_with_self_loops": nnz, "definition": args.definition, "gemm": args.precision,
        "ms_per_training_step": round(ms_step, 2), "ms_per_forward": round(ms_fwd, 2),
        "edges_per_s_training": round(3 * nnz / (ms_step * 1e-3)), "loss": round(float(loss.item()), 4),
        "forward_aggregation_alg_GB": round(fwd_bytes / 1e9, 1),
        "forward_alg_GBps_if_all_aggregation": round(fwd_bytes / ms_fwd / 1e6, 1),
        "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1)}), flush=True)


if __name__ == "__main__":
    main()
