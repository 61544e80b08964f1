"""MQ-GNN style producer/consumer pipeline around the device hot path.

Mirrors the reference's host-side overlap (GPU Accelerator/buffer_queues.py:22-119, MQGCN.py:55-79,150-157):
  sample_generator   producer thread, stream ``d_stream``: draw mini-batches from the dataloader (device sampling +
                     block construction + feature gather / halo fetch) and put ``[mfgs, feat, label, step]`` into a
                     bounded queue (BUFFER_SIZE = 4), ``None`` sentinel at the end        (buffer_queues.py:22-70)
  sample_consumer    consumer thread, stream ``c_stream``: forward, loss, backward; gradients are shared on ``g_stream``
                     (gradient_generator = all-reduce, gradient_consumer = set .grad + opt.step)   (:74-119, MQGCN.py:55-79)
Differences, all deliberate: cross-stream ordering is explicit (every queue item carries the CUDA event recorded after
its tensors were produced; the reference shares tensors between streams without any event), the gradient all-reduce is
ONE flat-buffer collective (parallel.allreduce_gradients) instead of one per parameter, and blocks come from the
device sampler so the producer thread spends its time in launches that release the GIL.
"""
import threading
import time
from queue import Queue

import torch

from . import parallel


def _put(gpu_queue, condition, item, abort):
    """Blocking put that gives up when the other side has failed (``abort`` set); returns False in that case."""
    with condition:
        while gpu_queue.full():
            if abort is not None and abort.is_set():
                return False
            condition.wait(timeout=0.2)
        gpu_queue.put(item)
        condition.notify_all()
    return True


def sample_generator(gpu_queue, condition, train_dataloader, fetch=None, labels=None, d_stream=None, abort=None):
    """Producer (buffer_queues.py:22-70).  ``fetch(input_nodes, mfgs)`` returns the layer-0 input features (or None
    when the model aggregates straight from the HBM table); ``labels[output_nodes]`` are the targets.  The ``None``
    sentinel is queued even when the dataloader or ``fetch`` raises, so the consumer never waits for ever; when the
    consumer has failed (``abort`` set) the producer stops instead of blocking on the full queue."""
    d_stream = d_stream or torch.cuda.Stream()
    step = -1
    try:
        with torch.cuda.stream(d_stream):
            for step, (input_nodes, output_nodes, mfgs) in enumerate(train_dataloader):
                feat = fetch(input_nodes, mfgs) if fetch is not None else None
                lab = labels[output_nodes] if labels is not None else None
                ev = torch.cuda.Event()
                ev.record(d_stream)
                if not _put(gpu_queue, condition, [mfgs, feat, lab, step, ev], abort):
                    break
    finally:
        if abort is None or not abort.is_set():
            _put(gpu_queue, condition, None, abort)
    return step + 1


def _record_stream(obj, stream):
    """Tell the caching allocator that ``stream`` uses tensors that were allocated on the producer's stream, so their
    memory is not recycled by the producer while the consumer's kernels still read them."""
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda:
            obj.record_stream(stream)
    elif isinstance(obj, (list, tuple)):
        for o in obj:
            _record_stream(o, stream)
    elif hasattr(obj, "__dict__"):
        for v in vars(obj).values():
            if isinstance(v, (torch.Tensor, list, tuple)):
                _record_stream(v, stream)


def gradient_generator(model, group=None):
    """MQGCN.py:55-67 — sum the gradients over the ranks and average them (one flat collective)."""
    parallel.allreduce_gradients([p for p in model.parameters() if p.requires_grad], group=group, average=True)


def gradient_consumer(opt):
    """MQGCN.py:70-79 — the averaged gradients are already in ``.grad``: take the optimizer step."""
    opt.step()


def sample_consumer(gpu_queue, condition, opt, model, forward=None, loss_fn=None, group=None, c_stream=None,
                    g_stream=None, stats=None, abort=None):
    """Consumer (buffer_queues.py:74-119).  ``forward(model, mfgs, feat)`` defaults to ``model(mfgs, feat)``.  On a
    failure it sets ``abort`` and drains the queue so a producer blocked on the full queue wakes up."""
    try:
        return _consume(gpu_queue, condition, opt, model, forward, loss_fn, group, c_stream, g_stream, stats)
    except BaseException:
        if abort is not None:
            abort.set()
        with condition:
            while not gpu_queue.empty():
                gpu_queue.get()
            condition.notify_all()
        raise


def _consume(gpu_queue, condition, opt, model, forward, loss_fn, group, c_stream, g_stream, stats):
    c_stream = c_stream or torch.cuda.Stream()
    g_stream = g_stream or torch.cuda.Stream()
    loss_fn = loss_fn or torch.nn.functional.cross_entropy
    forward = forward or (lambda m, mfgs, feat: m(mfgs, feat))
    multi = torch.distributed.is_available() and torch.distributed.is_initialized() and \
        torch.distributed.get_world_size(group) > 1
    n, loss_sum = 0, None
    model.train()
    with torch.cuda.stream(c_stream):
        while True:
            with condition:
                while gpu_queue.empty():
                    condition.wait()
                item = gpu_queue.get()
                condition.notify_all()
            if item is None:
                break
            mfgs, feat, lab, step, ev = item
            c_stream.wait_event(ev)                       # the producer's tensors are complete
            _record_stream([mfgs, feat, lab], c_stream)
            opt.zero_grad(set_to_none=True)
            predictions = forward(model, mfgs, feat)
            loss = loss_fn(predictions, lab)
            loss.backward()
            if multi:
                done = torch.cuda.Event()
                done.record(c_stream)
                with torch.cuda.stream(g_stream):         # gradient sharing on its own stream (g_stream, :108)
                    g_stream.wait_event(done)
                    gradient_generator(model, group)
                    shared = torch.cuda.Event()
                    shared.record(g_stream)
                c_stream.wait_event(shared)
            gradient_consumer(opt)
            loss_sum = loss.detach() if loss_sum is None else loss_sum + loss.detach()
            n += 1
        end = torch.cuda.Event()
        end.record(c_stream)
    end.synchronize()
    if stats is not None:
        stats["n_batches"] = n
        stats["loss"] = float(loss_sum.item()) / max(n, 1) if n else float("nan")
    return n


def run_epoch(train_dataloader, model, opt, fetch=None, labels=None, forward=None, loss_fn=None, group=None,
              BUFFER_SIZE=4):
    """One pipelined epoch (MQGCN.py:150-157: a 2-worker ThreadPoolExecutor around a bounded queue).
    Returns dict(time_s (wall), n_batches, loss)."""
    import concurrent.futures
    condition = threading.Condition()
    gpu_queue = Queue(maxsize=BUFFER_SIZE)
    stats = {}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev = torch.cuda.current_device()

    abort = threading.Event()

    def _prod():
        torch.cuda.set_device(dev)
        return sample_generator(gpu_queue, condition, train_dataloader, fetch=fetch, labels=labels, abort=abort)

    def _cons():
        torch.cuda.set_device(dev)
        return sample_consumer(gpu_queue, condition, opt, model, forward=forward, loss_fn=loss_fn, group=group,
                               stats=stats, abort=abort)

    with concurrent.futures.ThreadPoolExecutor(max_workers=2) as ex:
        fp, fc = ex.submit(_prod), ex.submit(_cons)
        concurrent.futures.wait([fp, fc])
        for f in (fc, fp):               # surface the failure (the consumer's first: it is the one that aborts the other)
            if f.exception() is not None:
                raise f.exception()
    torch.cuda.synchronize()
    stats["time_s"] = time.perf_counter() - t0
    return stats
