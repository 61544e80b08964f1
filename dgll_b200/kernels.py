"""Host-side launchers: torch tensors in, C-ABI calls out.

PyTorch is used for device memory and streams only; every function here hands
raw device pointers, sizes and the CURRENT torch CUDA stream to
``libdgll_b200.so`` (``include/dgll_b200.h``).  Calls are stream-ordered and never
synchronise.  Nothing here has a CPU path: non-CUDA tensors raise.
"""
import ctypes

import torch

from . import _lib
from ._lib import (BF16, EPI_ELU, EPI_RELU, F32, GAT_EXP_NEG, GAT_SOFTMAX, MAX, MEAN, SUM, check, lib)

_REDUCE = {"sum": SUM, "add": SUM, "mean": MEAN, "max": MAX}

set_option = _lib.set_option
get_option = _lib.get_option


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream():
    """The current torch stream of the current device as a raw ``cudaStream_t``.  Goes through torch's C entry points:
    ``torch.cuda.current_stream()`` builds a Python Stream object per call (~18 us, 12 % of a training step's host
    time in tools/profile_epoch.py)."""
    if _raw_stream is not None and _raw_device is not None:
        return ctypes.c_void_p(_raw_stream(_raw_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("dgll_b200: expected a CUDA tensor (there is no CPU fallback), got %s" % t.device)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _rowmajor(t, name):
    """(tensor, ld) for a 2-D tensor whose rows are contiguous (column slices are fine)."""
    if t.dim() != 2:
        raise ValueError("%s must be 2-D" % name)
    if t.size(1) > 1 and t.stride(1) != 1:
        t = t.contiguous()
    if t.size(0) > 1 and t.stride(0) < t.size(1):
        t = t.contiguous()
    ld = t.stride(0) if t.size(0) > 1 else max(t.size(1), 1)
    return t, ld


def _index32(t, name):
    if t is None:
        return None
    if t.dtype == torch.int32:
        return t.contiguous()
    if t.dtype == torch.int64:
        return t.to(torch.int32)
    raise TypeError("%s must be int32/int64" % name)


def _rowptr(t):
    if t.dtype not in (torch.int32, torch.int64):
        raise TypeError("row_ptr must be int32/int64")
    return t.contiguous(), int(t.dtype == torch.int64)


class CsrPlan:
    """nnz-split schedule for a static CSR (rows longer than ``chunk_edges`` are split)."""

    def __init__(self, row_ptr, chunk_edges=0):
        _need_cuda(row_ptr)
        rp, is64 = _rowptr(row_ptr)
        self._h = ctypes.c_void_p()
        self.n_rows = rp.numel() - 1
        check(lib().dgllb_csr_plan_create(_p(rp), is64, self.n_rows, int(chunk_edges), _stream(),
                                          ctypes.byref(self._h)), "csr_plan_create")
        a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
        check(lib().dgllb_csr_plan_info(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        self.n_heavy_rows, self.n_chunks, self.chunk_edges = a.value, b.value, c.value

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().dgllb_csr_plan_destroy(h)
            except Exception:
                pass


def spmm_csr(row_ptr, col_idx, x, values=None, reduce="sum", n_dst=None, out=None, row_scale=None,
             addend=None, bias=None, relu=False, elu=False, return_argmax=False, plan=None, F=None):
    """Neighbourhood aggregation ``out[i] = epi(row_scale[i]*reduce_e(values[e]*x[col[e]]) + addend[i] + bias)``.

    ``col_idx=None`` is a segment reduce over consecutive rows of ``x`` (pooling).
    ``x`` may be fp32 or bf16; the result is fp32.
    """
    _need_cuda(row_ptr, col_idx, x, values, out, row_scale, addend, bias)
    rp, is64 = _rowptr(row_ptr)
    n_dst = rp.numel() - 1 if n_dst is None else n_dst
    col = _index32(col_idx, "col_idx")
    x, ldx = _rowmajor(x, "x")
    F = x.size(1) if F is None else F
    if x.dtype == torch.float32:
        xd = F32
    elif x.dtype == torch.bfloat16:
        xd = BF16
    else:
        raise TypeError("x must be float32 or bfloat16")
    if out is None:
        out = torch.empty((n_dst, F), dtype=torch.float32, device=x.device)
    out_c, ldo = _rowmajor(out, "out")
    if out_c is not out:
        raise ValueError("out must have contiguous rows")
    if values is not None:
        values = values.to(torch.float32).contiguous()
    ld_add = 0
    if addend is not None:
        addend, ld_add = _rowmajor(addend.to(torch.float32), "addend")
    if bias is not None:
        bias = bias.to(torch.float32).contiguous()
    if row_scale is not None:
        row_scale = row_scale.to(torch.float32).contiguous()
    red = _REDUCE[reduce]
    argmax = None
    if return_argmax:
        argmax = torch.empty((n_dst, F), dtype=torch.int32, device=x.device)
    epi = (EPI_RELU if relu else 0) | (EPI_ELU if elu else 0)
    check(lib().dgllb_spmm_csr(_p(rp), is64, _p(col), _p(values), _p(x), xd, ldx, _p(out), ldo, n_dst,
                               x.size(0), (col.numel() if col is not None else -1), F, red, _p(row_scale), _p(addend), ld_add, _p(bias), epi,
                               _p(argmax), plan._h if plan is not None else None, _stream()), "spmm_csr")
    return (out, argmax) if return_argmax else out


def spmm_csr_sharded(row_ptr, col_idx, shard_ptrs, rows_per_shard, stride_bytes, F, dtype=torch.float32, values=None,
                     reduce="mean", n_dst=None, out=None):
    """Sum/mean aggregation straight from a node-range-partitioned table: ``shard_ptrs`` is a device int64 tensor of
    shard base addresses (``parallel.PeerShardedTable.shard_ptrs``), column ids are GLOBAL row ids.  fp32 result."""
    _need_cuda(row_ptr, col_idx, shard_ptrs, values, out)
    rp, is64 = _rowptr(row_ptr)
    n_dst = rp.numel() - 1 if n_dst is None else n_dst
    col = _index32(col_idx, "col_idx")
    if dtype == torch.float32:
        xd = F32
    elif dtype == torch.bfloat16:
        xd = BF16
    else:
        raise TypeError("table dtype must be float32 or bfloat16")
    if reduce not in ("sum", "mean"):
        raise ValueError("spmm_csr_sharded: reduce must be 'sum' or 'mean'")
    if out is None:
        out = torch.empty((n_dst, F), dtype=torch.float32, device=rp.device)
    o, ldo = _rowmajor(out, "out")
    if o is not out:
        raise ValueError("out must have contiguous rows")
    if values is not None:
        values = values.to(torch.float32).contiguous()
    check(lib().dgllb_spmm_csr_sharded(_p(rp), is64, _p(col), _p(values), _p(shard_ptrs), shard_ptrs.numel(),
                                       int(rows_per_shard), int(stride_bytes), xd, _p(out), ldo, n_dst, int(F),
                                       _REDUCE[reduce], _stream()), "spmm_csr_sharded")
    return out


def sddmm_csr(row_ptr, col_idx, a, b):
    """Per-edge dot products ``out[e] = <a[row(e)], b[col[e]]>`` (SpecialSpmm backward, gatconv.py:76-78)."""
    _need_cuda(row_ptr, col_idx, a, b)
    rp, is64 = _rowptr(row_ptr)
    col = _index32(col_idx, "col_idx")
    a, lda = _rowmajor(a.to(torch.float32), "a")
    b, ldb = _rowmajor(b.to(torch.float32), "b")
    out = torch.empty(col.numel(), dtype=torch.float32, device=a.device)
    check(lib().dgllb_sddmm_csr(_p(rp), is64, _p(col), _p(a), lda, _p(b), ldb, _p(out), rp.numel() - 1,
                                a.size(1), _stream()), "sddmm_csr")
    return out


def spmm_max_backward(col_idx, argmax, grad_out, n_src):
    _need_cuda(col_idx, argmax, grad_out)
    col = _index32(col_idx, "col_idx")
    g, ldg = _rowmajor(grad_out.to(torch.float32), "grad_out")
    gx = torch.zeros((n_src, g.size(1)), dtype=torch.float32, device=g.device)
    check(lib().dgllb_spmm_max_backward(_p(col), _p(argmax.contiguous()), _p(g), ldg, _p(gx), gx.stride(0),
                                        g.size(0), g.size(1), _stream()), "spmm_max_backward")
    return gx


def csr_transpose(row_ptr, col_idx, n_cols, values=None, want_perm=False, out=None):
    """CSR of A^T: returns (t_row_ptr, t_col_idx, t_values or None, perm or None).  ``out=(t_row_ptr, t_col_idx, perm)``
    writes into caller-owned buffers (fixed-capacity pipelines)."""
    _need_cuda(row_ptr, col_idx, values)
    rp, is64 = _rowptr(row_ptr)
    col = _index32(col_idx, "col_idx")
    nnz = col.numel()
    dev = rp.device
    if out is not None:
        t_rp, t_col, perm = out
        if t_rp.dtype != rp.dtype or t_rp.numel() != n_cols + 1 or t_col.numel() < nnz or (perm is not None and perm.numel() < nnz):
            raise ValueError("csr_transpose: output buffers do not fit")
    else:
        t_rp = torch.empty(n_cols + 1, dtype=rp.dtype, device=dev)
        t_col = torch.empty(nnz, dtype=torch.int32, device=dev)
        perm = torch.empty(nnz, dtype=torch.int32, device=dev) if want_perm else None
    t_val = torch.empty(nnz, dtype=torch.float32, device=dev) if values is not None else None
    if values is not None:
        values = values.to(torch.float32).contiguous()
    check(lib().dgllb_csr_transpose(_p(rp), is64, _p(col), _p(values), rp.numel() - 1, n_cols, nnz, _p(t_rp),
                                    _p(t_col), _p(t_val), _p(perm), _stream()), "csr_transpose")
    return t_rp, t_col, t_val, perm


def gather_rows(table, ids, out=None):
    """``out[i] = table[ids[i]]`` — exact byte copy through TMA bulk copies (dgraph.py:105)."""
    _need_cuda(table, ids, out)
    if ids.dtype not in (torch.int32, torch.int64):
        raise TypeError("ids must be int32/int64")
    ids = ids.contiguous()
    squeeze = table.dim() == 1
    t2 = table.unsqueeze(1) if squeeze else table
    if t2.dim() != 2:
        t2 = t2.reshape(t2.size(0), -1)
    t2, ld = _rowmajor(t2, "table")
    esz = t2.element_size()
    if out is None:
        out = torch.empty((ids.numel(), t2.size(1)), dtype=t2.dtype, device=t2.device)
    o2, ldo = _rowmajor(out if out.dim() == 2 else out.reshape(out.size(0), -1), "out")
    check(lib().dgllb_gather_rows(_p(t2), ld * esz, _p(ids), int(ids.dtype == torch.int64), _p(o2), ldo * esz,
                                  ids.numel(), t2.size(1) * esz, _stream()), "gather_rows")
    if squeeze:
        return out.reshape(-1)
    if table.dim() > 2:
        return out.reshape((ids.numel(),) + tuple(table.shape[1:]))
    return out


def gather_rows_cached(cache_table, host_table, ids, gpu_flag, localid2cacheid, nid_map=None, out=None,
                       miss_counter=None):
    """GraphCacheServer.fetch_data split gather (FeatureCache/storage.py:151-198) in one launch."""
    _need_cuda(cache_table, ids, gpu_flag, localid2cacheid, nid_map, out, miss_counter)
    ids = ids.to(torch.int64).contiguous()
    flag = gpu_flag.to(torch.uint8) if gpu_flag.dtype != torch.uint8 and gpu_flag.dtype != torch.bool else gpu_flag
    ref = cache_table if cache_table is not None else host_table
    width, esz = ref.size(1), ref.element_size()
    if out is None:
        out = torch.empty((ids.numel(), width), dtype=ref.dtype, device=ids.device)
    hp, hs = None, 0
    if host_table is not None:
        # pinned host memory is addressed directly by the kernel (UVA); device tensors work too
        if not host_table.is_cuda and not host_table.is_pinned():
            raise RuntimeError("host_table must be pinned (page-locked) or a CUDA tensor")
        hp, hs = ctypes.c_void_p(host_table.data_ptr()), host_table.stride(0) * esz
    cp, cs = (None, 0) if cache_table is None else (_p(cache_table), cache_table.stride(0) * esz)
    check(lib().dgllb_gather_rows_cached(cp, cs, hp, hs, _p(ids), _p(flag), _p(localid2cacheid.contiguous()),
                                         _p(nid_map), _p(out), out.stride(0) * esz, ids.numel(), width * esz,
                                         _p(miss_counter), _stream()), "gather_rows_cached")
    return out


def gather_rows_sharded(shard_ptrs, rows_per_shard, stride_bytes, ids, width, dtype, out=None):
    """``out[i] = shard[id // rows_per_shard][id % rows_per_shard]`` — ``shard_ptrs`` is a device int64 tensor of
    device pointers (local shard + peer shards mapped over NVLink, see ``parallel.PeerShardedTable``)."""
    _need_cuda(shard_ptrs, ids, out)
    if ids.dtype not in (torch.int32, torch.int64):
        raise TypeError("ids must be int32/int64")
    ids = ids.contiguous()
    esz = torch.empty(0, dtype=dtype).element_size()
    if out is None:
        out = torch.empty((ids.numel(), width), dtype=dtype, device=ids.device)
    check(lib().dgllb_gather_rows_sharded(_p(shard_ptrs), shard_ptrs.numel(), int(rows_per_shard), int(stride_bytes),
                                          _p(ids), int(ids.dtype == torch.int64), _p(out), out.stride(0) * esz,
                                          ids.numel(), width * esz, _stream()), "gather_rows_sharded")
    return out


def ipc_export(t):
    """(64-byte handle, offset) of the device allocation holding tensor ``t``."""
    _need_cuda(t)
    buf = (ctypes.c_ubyte * 64)()
    off = ctypes.c_int64()
    check(lib().dgllb_ipc_export(_p(t), buf, ctypes.byref(off)), "ipc_export")
    return bytes(buf), off.value


def ipc_import(handle, offset):
    """Device pointer (int) of a peer allocation mapped into this process."""
    buf = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
    out = ctypes.c_void_p()
    check(lib().dgllb_ipc_import(buf, int(offset), ctypes.byref(out)), "ipc_import")
    return out.value


def ipc_release(ptr, offset):
    check(lib().dgllb_ipc_release(ctypes.c_void_p(ptr), int(offset)), "ipc_release")


def gemm(a, b, bias=None, relu=False, elu=False, trans_a=False, trans_b=False, out=None, accumulate=False,
         precision="fp32"):
    """Dense transform ``op(a) @ op(b) (+bias)``: ``fp32`` = exact SIMT path, ``bf16`` = tcgen05 on packed bf16 operands,
    ``tf32`` = tcgen05 on the fp32 operands as they lie in memory (TMA, no packing), ``tf32x3`` = the same kernel with
    every operand word split into tf32 hi + lo in shared memory and three MMAs per k-step (fp32-grade results)."""
    _need_cuda(a, b, bias, out)
    a, lda = _rowmajor(a.to(torch.float32), "a")
    b, ldb = _rowmajor(b.to(torch.float32), "b")
    M = a.size(1) if trans_a else a.size(0)
    K = a.size(0) if trans_a else a.size(1)
    N = b.size(0) if trans_b else b.size(1)
    Kb = b.size(1) if trans_b else b.size(0)
    if K != Kb:
        raise ValueError("gemm: inner dimensions differ (%d vs %d)" % (K, Kb))
    if out is None:
        if accumulate:
            raise ValueError("accumulate=True needs out")
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    o, ldc = _rowmajor(out, "out")
    if o is not out:
        raise ValueError("out must have contiguous rows")
    if bias is not None:
        bias = bias.to(torch.float32).contiguous()
    epi = (EPI_RELU if relu else 0) | (EPI_ELU if elu else 0)
    prec = {"fp32": 0, "bf16": 1, "tf32": 2, "tf32x3": 3}[precision]
    check(lib().dgllb_gemm_f32(_p(a), lda, int(trans_a), _p(b), ldb, int(trans_b), _p(out), ldc, M, N, K,
                               _p(bias), epi, int(accumulate), prec, _stream()), "gemm")
    return out


def gat_forward(row_ptr, col_idx, wh, el, er, heads, slope, mode="softmax", elu=False, save_stats=False,
                n_dst=None, out=None, plan=None, dropout=0.0, seed=0):
    """Fused multi-head GAT aggregation; ``wh`` is [n_src, heads*D], ``el``/``er`` are [n, heads]."""
    _need_cuda(row_ptr, col_idx, wh, el, er, out)
    rp, is64 = _rowptr(row_ptr)
    n_dst = rp.numel() - 1 if n_dst is None else n_dst
    col = _index32(col_idx, "col_idx")
    wh, ldw = _rowmajor(wh, "wh")
    el, lde = _rowmajor(el, "el")
    er, lde2 = _rowmajor(er, "er")
    if lde != lde2:
        el, er = el.contiguous(), er.contiguous()
        lde = heads
    FD = wh.size(1)
    if FD % heads:
        raise ValueError("wh width %d not divisible by heads %d" % (FD, heads))
    if out is None:
        out = torch.empty((n_dst, FD), dtype=torch.float32, device=wh.device)
    o, ldo = _rowmajor(out, "out")
    rmax = rsum = None
    if save_stats:
        rmax = torch.empty((n_dst, heads), dtype=torch.float32, device=wh.device)
        rsum = torch.empty((n_dst, heads), dtype=torch.float32, device=wh.device)
    md = {"softmax": GAT_SOFTMAX, "exp_neg": GAT_EXP_NEG}[mode]
    check(lib().dgllb_gat_forward(_p(rp), is64, _p(col), _p(wh), ldw, _p(el), _p(er), lde, _p(out), ldo,
                                  _p(rmax), _p(rsum), n_dst, wh.size(0), heads, FD // heads, float(slope), md,
                                  EPI_ELU if elu else 0, float(dropout), ctypes.c_uint64(seed & 0xFFFFFFFFFFFFFFFF),
                                  plan._h if plan is not None else None, _stream()),
          "gat_forward")
    return (out, rmax, rsum) if save_stats else out


def gat_backward(row_ptr, col_idx, t_row_ptr, t_col_idx, perm, wh, el, er, out, rmax, rsum, grad_out, heads,
                 slope, mode="softmax", d_ext=None, dropout=0.0, seed=0, plan=None, t_plan=None):
    """Backward of :func:`gat_forward`.  Returns (d_wh, d_el, d_er) — views into one [n, heads*D+2*heads]
    buffer (``d_ext``) when n_src == n_dst so the dense-transform backward runs as one GEMM."""
    _need_cuda(row_ptr, col_idx, t_row_ptr, t_col_idx, perm, wh, el, er, out, rmax, rsum, grad_out)
    rp, is64 = _rowptr(row_ptr)
    trp, _ = _rowptr(t_row_ptr)
    col = _index32(col_idx, "col_idx")
    tcol = _index32(t_col_idx, "t_col_idx")
    wh, ldw = _rowmajor(wh, "wh")
    el, lde = _rowmajor(el, "el")
    er, lde2 = _rowmajor(er, "er")
    if lde != lde2:
        el, er = el.contiguous(), er.contiguous()
        lde = heads
    out, ldo = _rowmajor(out, "out")
    g, ldg = _rowmajor(grad_out.to(torch.float32), "grad_out")
    n_dst, n_src, FD = rp.numel() - 1, wh.size(0), wh.size(1)
    dev = wh.device
    if d_ext is not None:
        d_wh, d_el, d_er = d_ext[:, :FD], d_ext[:, FD:FD + heads], d_ext[:, FD + heads:FD + 2 * heads]
        ldd, ldde = d_ext.stride(0), d_ext.stride(0)
        if n_src != n_dst:
            raise ValueError("d_ext needs n_src == n_dst")
    else:
        d_wh = torch.empty((n_src, FD), dtype=torch.float32, device=dev)
        d_el = torch.empty((n_dst, heads), dtype=torch.float32, device=dev)
        d_er = torch.empty((n_src, heads), dtype=torch.float32, device=dev)
        ldd, ldde = FD, heads
    ws = torch.empty(2 * col.numel() * heads, dtype=torch.float32, device=dev)
    md = {"softmax": GAT_SOFTMAX, "exp_neg": GAT_EXP_NEG}[mode]
    check(lib().dgllb_gat_backward(_p(rp), is64, _p(col), _p(trp), _p(tcol), _p(perm), _p(wh), ldw, _p(el), _p(er),
                                   lde, _p(out), ldo, _p(rmax), _p(rsum), _p(g), ldg, _p(d_wh), ldd, _p(d_el),
                                   _p(d_er), ldde, _p(ws), n_dst, n_src, heads, FD // heads, float(slope), md,
                                   float(dropout), ctypes.c_uint64(seed & 0xFFFFFFFFFFFFFFFF),
                                   plan._h if plan is not None else None, t_plan._h if t_plan is not None else None,
                                   _stream()), "gat_backward")
    return d_wh, d_el, d_er


def gat_dropout_mask(seed, nnz, heads, p, device="cuda"):
    """The [nnz, heads] multiplier (0 or 1/(1-p)) the GAT kernels apply for this seed — for tests / replay."""
    out = torch.empty((nnz, heads), dtype=torch.float32, device=device)
    check(lib().dgllb_gat_dropout_mask(ctypes.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), nnz, heads, float(p), _p(out),
                                       _stream()), "gat_dropout_mask")
    return out


def packed_words(F):
    """Words per packed row: ceil(F/32) rounded up to a multiple of 4 (16-byte rows)."""
    return ((F + 31) // 32 + 3) // 4 * 4


def binarize_pack(x):
    """Bit-pack ``x >= 0`` along features: uint32 words as an int32 tensor."""
    _need_cuda(x)
    x, ldx = _rowmajor(x.to(torch.float32), "x")
    wpr = packed_words(x.size(1))
    packed = torch.empty((x.size(0), wpr), dtype=torch.int32, device=x.device)
    check(lib().dgllb_binarize_pack(_p(x), ldx, _p(packed), wpr, x.size(0), x.size(1), _stream()), "binarize_pack")
    return packed


def bin_spmm_csr(row_ptr, col_idx, packed, F, mode="count", n_dst=None, plan=None):
    """Binarized aggregation: ``count`` (int32), ``sum`` (+-1 sum, fp32) or ``mean`` (+-1 mean, fp32)."""
    _need_cuda(row_ptr, col_idx, packed)
    rp, is64 = _rowptr(row_ptr)
    n_dst = rp.numel() - 1 if n_dst is None else n_dst
    col = _index32(col_idx, "col_idx")
    packed = packed.contiguous()
    md = {"count": 0, "sum": 1, "mean": 2}[mode]
    out = torch.empty((n_dst, F), dtype=torch.int32 if md == 0 else torch.float32, device=packed.device)
    check(lib().dgllb_bin_spmm_csr(_p(rp), is64, _p(col), _p(packed), packed.size(1), _p(out), F, n_dst, F, md,
                                   plan._h if plan is not None else None, _stream()), "bin_spmm_csr")
    return out


def sample_neighbors(row_ptr, col_idx, seeds, fanout, rng_seed=0):
    """Uniform sampling without replacement, min(deg, fanout) per seed → (block_row_ptr int32, block_col int32)."""
    _need_cuda(row_ptr, col_idx, seeds)
    rp, is64 = _rowptr(row_ptr)
    col = _index32(col_idx, "col_idx")
    if seeds.dtype not in (torch.int32, torch.int64):
        raise TypeError("seeds must be int32/int64")
    seeds = seeds.contiguous()
    n = seeds.numel()
    dev = rp.device
    out_rp = torch.empty(n + 1, dtype=torch.int32, device=dev)
    if fanout is None or fanout < 0:
        fo = -1
        deg = (rp[1:] - rp[:-1])[seeds.long()]
        cap = int(deg.sum().item())
    else:
        fo = int(fanout)
        cap = n * fo
    out_col = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    check(lib().dgllb_sample_neighbors(_p(rp), is64, _p(col), _p(seeds), int(seeds.dtype == torch.int64), n, fo,
                                       ctypes.c_uint64(rng_seed & 0xFFFFFFFFFFFFFFFF), _p(out_rp), _p(out_col),
                                       _stream()), "sample_neighbors")
    return out_rp, out_col


def build_block(dst_ids, row_ptr, nbr_global):
    """dst-first compaction on the device.  Returns (src_ids int64[cap], col_local int32[cap_nnz], counts int32[2] on
    the device = {num_src, nnz}); slice with the counts after reading them back once."""
    _need_cuda(dst_ids, row_ptr, nbr_global)
    dst_ids = dst_ids.to(torch.int64).contiguous()
    row_ptr = row_ptr.contiguous()
    if row_ptr.dtype != torch.int32:
        raise TypeError("build_block: row_ptr must be int32 (output of sample_neighbors)")
    nbr = _index32(nbr_global, "nbr_global")
    n_dst, cap = dst_ids.numel(), nbr.numel()
    dev = dst_ids.device
    src_ids = torch.empty(n_dst + cap, dtype=torch.int64, device=dev)
    col_local = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    counts = torch.empty(2, dtype=torch.int32, device=dev)
    check(lib().dgllb_build_block(_p(dst_ids), n_dst, _p(row_ptr), _p(nbr), cap, _p(src_ids), _p(col_local), _p(counts),
                                  _stream()), "build_block")
    return src_ids, col_local, counts


# ------------------------------------------- layer-wise importance sampling ---
def csr_slice_rows(row_ptr, col_idx, values, rows):
    """Q = M[rows, :] on the device.  Returns (q_row_ptr int64[n+1], q_col int32[nnz], q_values f64[nnz] | None);
    one read-back (nnz) to size the outputs."""
    _need_cuda(row_ptr, col_idx, values, rows)
    rp, is64 = _rowptr(row_ptr)
    col = _index32(col_idx, "col_idx")
    rows = rows.to(torch.int64).contiguous()
    n = rows.numel()
    dev = col.device
    q_rp = torch.empty(n + 1, dtype=torch.int64, device=dev)
    check(lib().dgllb_csr_slice_rows_ptr(_p(rp), is64, _p(rows), n, _p(q_rp), _stream()), "csr_slice_rows")
    nnz = int(q_rp[-1].item())
    q_col = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)[:nnz]
    q_val = None
    if values is not None:
        if values.dtype != torch.float64:
            raise TypeError("csr_slice_rows: values must be float64 (the Laplacian's dtype)")
        values = values.contiguous()
        q_val = torch.empty(max(nnz, 1), dtype=torch.float64, device=dev)[:nnz]
    check(lib().dgllb_csr_slice_rows_fill(_p(rp), is64, _p(col), _p(values), _p(rows), n, _p(q_rp), _p(q_col),
                                          _p(q_val), _stream()), "csr_slice_rows")
    return q_rp, q_col, q_val


def col_sqsum(col_idx, values, n_cols, flat=False):
    """Distinct columns (ascending) and normalised probabilities sum(v^2) (sqrt if flat) / total.
    Returns (cand_cols int32[nnz], cand_prob f64[nnz], stats f64[3] on the device = {n_cand, total, n_positive})."""
    _need_cuda(col_idx, values)
    col = _index32(col_idx, "col_idx")
    if values.dtype != torch.float64:
        raise TypeError("col_sqsum: values must be float64")
    values = values.contiguous()
    nnz, dev = col.numel(), col.device
    cand_cols = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)[:nnz]
    cand_prob = torch.empty(max(nnz, 1), dtype=torch.float64, device=dev)[:nnz]
    stats = torch.empty(3, dtype=torch.float64, device=dev)
    check(lib().dgllb_col_sqsum(_p(col), _p(values), nnz, int(n_cols), int(bool(flat)), _p(cand_cols), _p(cand_prob),
                                _p(stats), _stream()), "col_sqsum")
    return cand_cols, cand_prob, stats


def weighted_choice(cand_cols, cand_prob, fanout, seed):
    """Weighted draw without replacement, in drawing order.  Returns (sel int32[fanout], picks int64[fanout],
    count int64[1] on the device)."""
    _need_cuda(cand_cols, cand_prob)
    if cand_prob.dtype != torch.float64:
        raise TypeError("weighted_choice: probabilities must be float64")
    cand_prob = cand_prob.contiguous()
    cc = None if cand_cols is None else _index32(cand_cols, "cand_cols")
    dev = cand_prob.device
    fanout = int(fanout)
    sel = torch.empty(max(fanout, 1), dtype=torch.int32, device=dev)[:fanout]
    picks = torch.empty(max(fanout, 1), dtype=torch.int64, device=dev)[:fanout]
    count = torch.empty(1, dtype=torch.int64, device=dev)
    check(lib().dgllb_weighted_choice(_p(cc), _p(cand_prob), cand_prob.numel(), fanout,
                                      ctypes.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), _p(sel), _p(picks), _p(count),
                                      _stream()), "weighted_choice")
    return sel, picks, count


def importance_scale(cand_prob, sel, count, n_total, mode):
    """Column scales of the drawn candidates: ``mode='inverse'`` 1/p/count, ``mode='wrs'`` the WRS estimator."""
    _need_cuda(cand_prob, sel, count)
    sel = sel.contiguous()
    if sel.dtype != torch.int32 or count.dtype != torch.int64 or cand_prob.dtype != torch.float64:
        raise TypeError("importance_scale: sel int32, count int64, cand_prob float64 expected")
    cap = sel.numel()
    scale = torch.empty(max(cap, 1), dtype=torch.float64, device=sel.device)[:cap]
    check(lib().dgllb_importance_scale(_p(cand_prob.contiguous()), _p(sel), _p(count), cap, int(n_total),
                                       {"inverse": 0, "wrs": 1}[mode], _p(scale), _stream()), "importance_scale")
    return scale


def scatter_pos(pos, picks, count=None, reset=False):
    """pos[picks[k]] = k (or -1 when ``reset``) for k < count; in place."""
    _need_cuda(pos, picks, count)
    if pos.dtype != torch.int32 or not pos.is_contiguous():
        raise TypeError("scatter_pos: pos must be a contiguous int32 tensor")
    picks = picks.to(torch.int64).contiguous()
    check(lib().dgllb_scatter_pos(_p(pos), _p(picks), _p(count), picks.numel(), int(bool(reset)), _stream()),
          "scatter_pos")
    return pos


def csr_select_cols(q_row_ptr, q_col, q_values, pos, scale=None, with_values=True):
    """adj = Q[:, picks] * scale with columns relabelled through ``pos`` and rows sorted by the new label.
    Returns (row_ptr int64[n+1], col int32[cap], values f64[cap] | None); the true nnz is row_ptr[-1] (device)."""
    _need_cuda(q_row_ptr, q_col, q_values, pos, scale)
    if q_row_ptr.dtype != torch.int64:
        raise TypeError("csr_select_cols: q_row_ptr must be int64 (output of csr_slice_rows)")
    q_rp = q_row_ptr.contiguous()
    q_col = _index32(q_col, "q_col")
    n, cap, dev = q_rp.numel() - 1, q_col.numel(), q_col.device
    out_rp = torch.empty(n + 1, dtype=torch.int64, device=dev)
    out_col = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)[:cap]
    out_val = torch.empty(max(cap, 1), dtype=torch.float64, device=dev)[:cap] if with_values else None
    check(lib().dgllb_csr_select_cols(_p(q_rp), _p(q_col), _p(q_values), n, cap, _p(pos), _p(scale), _p(out_rp),
                                      _p(out_col), _p(out_val), _stream()), "csr_select_cols")
    return out_rp, out_col, out_val


def sample_neighbors_cap(row_ptr, col_idx, seeds, fanout, rng_seed=0, rng_offset=None, out_row_ptr=None, out_col=None):
    """Fixed-capacity ``sample_neighbors``: negative entries of ``seeds`` are padding slots (degree-0 rows), the random
    seed is ``rng_seed + rng_offset[0]`` with ``rng_offset`` a uint64/int64 device scalar.  No size is read back:
    returns (row_ptr int32[n+1], col int32[n*fanout]); the true nnz is row_ptr[-1] on the device."""
    _need_cuda(row_ptr, col_idx, seeds, rng_offset, out_row_ptr, out_col)
    rp, is64 = _rowptr(row_ptr)
    col = _index32(col_idx, "col_idx")
    if seeds.dtype not in (torch.int32, torch.int64) or not seeds.is_contiguous():
        raise TypeError("seeds must be a contiguous int32/int64 tensor")
    if fanout is None or fanout < 0:
        raise ValueError("sample_neighbors_cap needs a fixed fanout (the output capacity is n * fanout)")
    n, dev = seeds.numel(), rp.device
    if out_row_ptr is None:
        out_row_ptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    if out_col is None:
        out_col = torch.empty(max(n * int(fanout), 1), dtype=torch.int32, device=dev)
    if out_row_ptr.numel() != n + 1 or out_col.numel() < n * int(fanout):
        raise ValueError("sample_neighbors_cap: output buffers too small")
    check(lib().dgllb_sample_neighbors_cap(_p(rp), is64, _p(col), _p(seeds), int(seeds.dtype == torch.int64), n,
                                           int(fanout), ctypes.c_uint64(rng_seed & 0xFFFFFFFFFFFFFFFF), _p(rng_offset),
                                           _p(out_row_ptr), _p(out_col), _stream()), "sample_neighbors_cap")
    return out_row_ptr, out_col


def build_block_cap(dst_ids, row_ptr, nbr_global, col_pad, src_ids=None, col_local=None, counts=None):
    """Fixed-capacity ``build_block``: negative ``dst_ids`` are padding; ``src_ids`` comes back -1 padded (usable as the
    next layer's seed array as it is), unused ``col_local`` slots = ``col_pad``; counts int32[3] on the device =
    {num_src, nnz, n_dst_valid}.  Nothing is read back."""
    _need_cuda(dst_ids, row_ptr, nbr_global, src_ids, col_local, counts)
    if dst_ids.dtype != torch.int64 or not dst_ids.is_contiguous():
        raise TypeError("build_block_cap: dst_ids must be a contiguous int64 tensor")
    if row_ptr.dtype != torch.int32 or not row_ptr.is_contiguous():
        raise TypeError("build_block_cap: row_ptr must be int32 (output of sample_neighbors_cap)")
    nbr = _index32(nbr_global, "nbr_global")
    n_dst, cap, dev = dst_ids.numel(), nbr.numel(), dst_ids.device
    if src_ids is None:
        src_ids = torch.empty(n_dst + cap, dtype=torch.int64, device=dev)
    if col_local is None:
        col_local = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    if counts is None:
        counts = torch.empty(3, dtype=torch.int32, device=dev)
    if src_ids.numel() < n_dst + cap or col_local.numel() < cap or counts.numel() < 3:
        raise ValueError("build_block_cap: output buffers too small")
    check(lib().dgllb_build_block_cap(_p(dst_ids), n_dst, _p(row_ptr), _p(nbr), cap, int(col_pad), _p(src_ids),
                                      _p(col_local), _p(counts), _stream()), "build_block_cap")
    return src_ids, col_local, counts


def gcn_fused_forward_v2(row_ptr, col_idx, values, X, W, num_neighbors, actual_F):
    _need_cuda(row_ptr, col_idx, values, X, W, num_neighbors)
    N, Fp, Hd = X.size(0), X.size(1), W.size(1)
    H = torch.empty((N, Hd), dtype=torch.float32, device=X.device)
    check(lib().dgllb_gcn_fused_forward(_p(row_ptr), _p(col_idx), _p(values), _p(X), _p(W), _p(H),
                                        _p(num_neighbors), N, Fp, int(actual_F), Hd, col_idx.numel(), _stream()),
          "gcn_fused_forward")
    return H


def gcn_fused_backward_v2(grad_output, row_ptr, col_idx, values, X, W, H, num_neighbors, actual_F):
    _need_cuda(grad_output, row_ptr, col_idx, values, X, W, H, num_neighbors)
    N, Fp, Hd = X.size(0), X.size(1), W.size(1)
    gW, gX = torch.zeros_like(W), torch.zeros_like(X)
    check(lib().dgllb_gcn_fused_backward(_p(row_ptr), _p(col_idx), _p(values), _p(X), _p(W), _p(H),
                                         _p(grad_output.contiguous()), _p(gW), _p(gX), _p(num_neighbors), N, Fp,
                                         int(actual_F), Hd, col_idx.numel(), _stream()), "gcn_fused_backward")
    return gX, gW
