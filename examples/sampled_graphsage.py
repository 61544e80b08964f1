"""Sampled GraphSAGE mini-batch training with the DGL-named API the reference's GPU-Accelerator scripts use
(GPU Accelerator/MQGCN.py:114-157, CommGNNModel.py:61-100): NeighborSampler -> DataLoader -> GraphSAGE(blocks, x).
Everything (sampling, block construction, feature gather, aggregation, transforms) runs on the device."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout

import torch

from dgll_b200 import graphs as G
from dgll_b200 import ops
from dgll_b200.data import BlockDataLoader, DGraph, NeighborSampler
from dgll_b200.nn import GraphSAGE


def main(n=50000, deg=30, feats=100, classes=10, batches=30, seed=0):
    dev = torch.device("cuda")
    row_ptr, col = G.rmat_csr(n, n * deg, seed=seed, device=dev)
    table = G.feature_table(n, feats, seed=seed, device=dev)          # [n, 100] fp32, rows 16-byte aligned
    labels = torch.randint(0, classes, (n,), device=dev)
    graph = DGraph(nodes=torch.arange(n, device=dev), edges=(row_ptr, col), labels=labels, features=table, device=dev)
    loader = BlockDataLoader(graph, torch.arange(n // 2, device=dev), NeighborSampler([25, 10]), batch_size=1024,
                             shuffle=True, drop_last=True)
    model = GraphSAGE(feats, 128, classes, 2, torch.relu, 0.0).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=0.003, fused=True)
    ops.set_gemm_precision("bf16")                                     # tcgen05 tensor cores for the dense transforms
    losses = []
    for step, (input_nodes, output_nodes, mfgs) in enumerate(loader):
        x = graph.get_features(input_nodes)                            # TMA row gather of the block's source rows
        logits = model(mfgs, x)
        loss = torch.nn.functional.cross_entropy(logits, graph.get_labels(output_nodes))
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
        if step + 1 == batches:
            break
    ops.set_gemm_precision("fp32")
    print("sampled_graphsage: %d mini-batches, loss %.4f -> %.4f" % (len(losses), losses[0], losses[-1]))
    return losses


if __name__ == "__main__":
    main()
