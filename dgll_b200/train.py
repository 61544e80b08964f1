"""Sampled GraphSAGE training epoch on the B200 path — the loop of the reference's mini-batch trainers
(GPU Accelerator/CommGNN_train.py:102-145, MQGCN.py:117-157) with every stage on the device:
sampling (``dgllb_sample_neighbors`` + dst-first compaction), layer-0 aggregation straight from the HBM feature table
(gather fused), dense transforms (exact fp32 or tcgen05 bf16), loss, backward through the same kernels, fused flat
gradient all-reduce (``parallel.allreduce_gradients``) and the optimizer step.
"""
import time

import torch

from . import graphs as G
from . import ops, parallel


def make_batches(row_ptr, col_idx, seeds, fanouts, batch_size, rng_seed=0):
    """Pre-sample every mini-batch of an epoch (mode A, "aggregation epoch": blocks resident on the device)."""
    out = []
    for b, i in enumerate(range(0, seeds.numel(), batch_size)):
        s = seeds[i:i + batch_size]
        out.append((s, G.sample_blocks(row_ptr, col_idx, s, fanouts, rng_seed=rng_seed * 7919 + b)))
    return out


def sage_epoch(model, opt, table, labels, n_feat, row_ptr=None, col_idx=None, seeds=None, fanouts=(25, 10),
               batch_size=1024, batches=None, rng_seed=0, group=None, precision=None):
    """One epoch.  ``batches`` (from ``make_batches``) = pre-sampled blocks; otherwise the sampler runs in the loop.
    Returns dict(time_s, n_batches, loss, sample_s) — time from CUDA events on the current stream."""
    if precision is not None:
        prev = ops.get_gemm_precision()
        ops.set_gemm_precision(precision)
    model.train()
    params = [p for p in model.parameters() if p.requires_grad]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_wall = time.perf_counter()
    e0.record()
    loss_sum = torch.zeros((), device=table.device)
    n = 0
    it = batches if batches is not None else range(0, seeds.numel(), batch_size)
    for b, item in enumerate(it):
        if batches is not None:
            s, blocks = item
        else:
            s = seeds[item:item + batch_size]
            blocks = G.sample_blocks(row_ptr, col_idx, s, fanouts, rng_seed=rng_seed * 7919 + b)
        logits = model(blocks, None, feat_table=table)   # padded table: SAGEConv slices the logical width itself
        loss = torch.nn.functional.cross_entropy(logits, labels[s])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        parallel.allreduce_gradients(params, group=group)
        opt.step()
        loss_sum += loss.detach()
        n += 1
    e1.record()
    torch.cuda.synchronize()
    if precision is not None:
        ops.set_gemm_precision(prev)
    return {"time_s": e0.elapsed_time(e1) * 1e-3, "wall_s": time.perf_counter() - t_wall, "n_batches": n,
            "loss": float(loss_sum.item()) / max(n, 1)}


def sage_epoch_pipelined(model, opt, table, labels, row_ptr, col_idx, seeds, fanouts=(25, 10), batch_size=1024,
                         rng_seed=0, group=None, precision=None, buffer_size=4):
    """The same epoch as ``sage_epoch`` (sampler in the loop) run through the MQ-GNN style producer/consumer pipeline
    (``dgll_b200.pipeline``): sampling + block construction on a producer thread / stream, training on the consumer."""
    from . import pipeline

    if precision is not None:
        prev = ops.get_gemm_precision()
        ops.set_gemm_precision(precision)

    def batches():
        for b, i in enumerate(range(0, seeds.numel(), batch_size)):
            s = seeds[i:i + batch_size]
            blocks = G.sample_blocks(row_ptr, col_idx, s, fanouts, rng_seed=rng_seed * 7919 + b)
            yield blocks[0].src_ids, s, blocks

    res = pipeline.run_epoch(batches(), model, opt, fetch=None, labels=labels, group=group, BUFFER_SIZE=buffer_size,
                             forward=lambda m, mfgs, feat: m(mfgs, None, feat_table=table))
    if precision is not None:
        ops.set_gemm_precision(prev)
    return res
