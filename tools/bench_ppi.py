#!/usr/bin/env python
"""BASELINE.json configs[0]: 2-layer GCN, full batch per graph, PPI (Evaluation/PPI) — the reference's own
CPU-runnable case, timed beside the B200 path on the same inputs.

  python tools/bench_ppi.py [--epochs 5]

Inputs: the two real PPI train graphs committed as fixtures (tests/golden/ppi_gcn_g5/g8.npz, graphs 5 and 8 of the
bundle) plus 18 synthetic graphs with the (N, directed nnz) of the other 18 train graphs (the fixture carries the size
list; symmetric, no duplicates, 50 features, 121 multi-hot labels like the real ones) — so one "epoch" is the reference's
loop over 20 graphs (Evaluation/PPI/train_gcn.py:29-57: forward, CrossEntropyLoss on float multi-hot targets, backward,
Adam step per graph), total N = 44,906, nnz = 1,226,368.
  reference arm  oracle.layers.ppi_gcn: the restatement of gcn_model.py:44-94 (COO rebuilt in every layer of every step,
                 torch.sparse.mm) in fp32 on the host cores, torch autograd + Adam — what Evaluation/PPI runs
  B200 arm       dgll_b200.nn.ppi.PPIGCN (same class layout / parameter names) on the aggregation kernels
Prints one JSON line: epoch seconds of both arms (median over --epochs), and the final-loss agreement after training both
from the same initial weights on the same graphs."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from dgll_b200.nn.ppi import PPIGCN  # noqa: E402
from oracle import layers as L  # noqa: E402   (bench reference arm: the one place outside tests that may use oracle/)


def load_graphs():
    gold = os.path.join(ROOT, "tests", "golden")
    real = {5: dict(np.load(os.path.join(gold, "ppi_gcn_g5.npz"))), 8: dict(np.load(os.path.join(gold, "ppi_gcn_g8.npz")))}
    sizes = real[5]["all_sizes"]
    rng = np.random.default_rng(0)
    graphs = []
    for gi, (n, nnz) in enumerate(sizes):
        if gi in real:
            g = real[gi]
            graphs.append((g["edge_index"].astype(np.int64), g["feats"].astype(np.float32), g["labels"].astype(np.float32)))
            continue
        n, half = int(n), int(nnz) // 2
        seen = set()
        while len(seen) < half:
            a, b = rng.integers(0, n, size=2 * (half - len(seen)) + 16).reshape(2, -1)
            for u, v in zip(a.tolist(), b.tolist()):
                if u != v and (min(u, v), max(u, v)) not in seen and len(seen) < half:
                    seen.add((min(u, v), max(u, v)))
        e = np.array(sorted(seen), dtype=np.int64).T
        ei = np.concatenate([e, e[::-1]], axis=1)
        feats = rng.standard_normal((n, 50)).astype(np.float32)
        labels = (rng.random((n, 121)) < 0.3).astype(np.float32)
        graphs.append((ei, feats, labels))
    return graphs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    args = ap.parse_args()
    graphs = load_graphs()
    torch.manual_seed(0)
    ref_model = PPIGCN(50, 64, 121, 2)                          # same initial weights for both arms
    init = {k: v.clone() for k, v in ref_model.state_dict().items()}

    # ---- reference arm (CPU) ----
    torch.set_num_threads(args.threads)
    ws = [init["layers.0.weight"].clone().requires_grad_(True), init["layers.1.weight"].clone().requires_grad_(True)]
    w_out = init["out_layer.weight"].clone().requires_grad_(True)
    b_out = init["out_layer.bias"].clone().requires_grad_(True)
    opt_c = torch.optim.Adam(ws + [w_out, b_out], lr=0.01)
    cpu_graphs = [(torch.from_numpy(ei), torch.from_numpy(x), torch.from_numpy(y)) for ei, x, y in graphs]
    cpu_t, cpu_loss = [], None
    for ep in range(args.epochs):
        t0 = time.perf_counter()
        tot = 0.0
        for ei, x, y in cpu_graphs:
            opt_c.zero_grad()
            loss = L.ppi_loss(L.ppi_gcn(ei, x, ws, w_out, b_out), y)
            loss.backward()
            opt_c.step()
            tot += loss.item()
        cpu_t.append(time.perf_counter() - t0)
        cpu_loss = tot / len(cpu_graphs)

    # ---- B200 arm ----
    dev = torch.device("cuda", 0)
    model = PPIGCN(50, 64, 121, 2)
    model.load_state_dict(init)
    model = model.to(dev)
    opt_g = torch.optim.Adam(model.parameters(), lr=0.01, fused=True)
    gpu_graphs = [(torch.from_numpy(ei).to(dev), torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev))
                  for ei, x, y in graphs]
    gpu_t, gpu_loss = [], None
    for ep in range(args.epochs):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tot = torch.zeros((), device=dev)
        for ei, x, y in gpu_graphs:
            opt_g.zero_grad(set_to_none=True)
            loss = torch.nn.functional.cross_entropy(model(ei, x), y)
            loss.backward()
            opt_g.step()
            tot += loss.detach()
        torch.cuda.synchronize()
        gpu_t.append(time.perf_counter() - t0)
        gpu_loss = float(tot.item()) / len(gpu_graphs)
    med = lambda v: sorted(v)[len(v) // 2]
    print(json.dumps({
        "workload": "2-layer GCN 50-64-64-121, full batch per graph, 20 PPI-shaped train graphs (2 real, 18 synthetic of "
                    "the real sizes), forward + loss + backward + Adam per graph (BASELINE configs[0])",
        "total_nodes": int(sum(x.shape[0] for _, x, _ in graphs)), "total_nnz": int(sum(ei.shape[1] for ei, _, _ in graphs)),
        "epoch_s_reference_cpu": round(med(cpu_t), 4), "cpu_threads": args.threads,
        "epoch_s_b200": round(med(gpu_t[1:] or gpu_t), 4), "first_epoch_s_b200_incl_csr_build": round(gpu_t[0], 4),
        "speedup": round(med(cpu_t) / med(gpu_t[1:] or gpu_t), 1),
        "mean_loss_last_epoch": {"reference_cpu": round(cpu_loss, 4), "b200": round(gpu_loss, 4)},
        "loss_rel_diff": abs(cpu_loss - gpu_loss) / abs(cpu_loss)}), flush=True)


if __name__ == "__main__":
    main()
