"""CPU oracle for the neighbourhood-aggregation path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; nothing under ``dgll_b200/``
does (the product path has no CPU fallback).

``oracle.c``     plain-C restatement of the kernels' arithmetic (numpy in/out here)
``layers.py``    torch-CPU restatement of the reference's layer classes, line by line
``samplers.py``  restatement of the reference's host samplers / graph store / cache
``gen_golden.py`` runs the REAL reference modules (build container only) and writes
                 ``tests/golden/*.npz``; the restatements are pinned against those.

Pinning status: pinned by the committed golden vectors (generated from the
reference's own code) for GCN / PPI-GCN / dense GAT / sparse GAT / samplers /
normalisation; the reference has no numeric tests of its own.
DGL ``GraphConv``/``SAGEConv`` (third-party, not in the tree), the fixed
``sageConv`` (broken as shipped) and the binarized SpMM (no reference code) are
"parity unpinned": restated from published semantics / SURVEY.md §8 a18.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_REF_LIB = os.path.join(_HERE, "_ref", "libgcn_fused_ref.so")
_lib = None

SUM, MEAN, MAX = 0, 1, 2
EPI_RELU, EPI_ELU = 1, 2


def build(ref=True):
    """Compile oracle.c (and oracle/_ref from the mounted reference, when present)."""
    targets = ["oracle"] + (["ref"] if ref else [])
    r = subprocess.run(["make", "-C", _HERE] + targets, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout.decode(errors="replace"))
    return _LIB


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "oracle.c")):
            build(ref=False)
        _lib = ctypes.CDLL(_LIB)
    return _lib


def ref_kernel_path():
    """Path of the reference's own forward kernel built for sm_100a, or None."""
    return _REF_LIB if os.path.exists(_REF_LIB) else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def spmm_csr(row_ptr, col, x, values=None, reduce="sum", row_scale=None, addend=None, bias=None, relu=False,
             elu=False, return_argmax=False):
    L = _load()
    rp = np.ascontiguousarray(row_ptr, dtype=np.int64)
    n_dst = rp.size - 1
    c = None if col is None else np.ascontiguousarray(col, dtype=np.int32)
    x = _f32(x)
    F = x.shape[1]
    v = None if values is None else _f32(values)
    rs = None if row_scale is None else _f32(row_scale)
    ad = None if addend is None else _f32(addend)
    b = None if bias is None else _f32(bias)
    out = np.empty((n_dst, F), dtype=np.float32)
    am = np.empty((n_dst, F), dtype=np.int32) if return_argmax else None
    red = {"sum": SUM, "add": SUM, "mean": MEAN, "max": MAX}[reduce]
    epi = (EPI_RELU if relu else 0) | (EPI_ELU if elu else 0)
    L.orc_spmm_csr(_ptr(rp), _ptr(c), _ptr(v), _ptr(x), ctypes.c_int64(F), _ptr(out), ctypes.c_int64(F),
                   ctypes.c_int64(n_dst), ctypes.c_int(F), ctypes.c_int(red), _ptr(rs), _ptr(ad),
                   ctypes.c_int64(F if ad is not None else 0), _ptr(b), ctypes.c_int(epi), _ptr(am))
    return (out, am) if return_argmax else out


def gcn_fused_forward(row_ptr, col_idx, values, X, W, num_neighbors, actual_F):
    L = _load()
    rp = np.ascontiguousarray(row_ptr, dtype=np.int32)
    ci = np.ascontiguousarray(col_idx, dtype=np.int32)
    nn = np.ascontiguousarray(num_neighbors, dtype=np.int32)
    v, X, W = _f32(values), _f32(X), _f32(W)
    N, Fp, Hd = X.shape[0], X.shape[1], W.shape[1]
    H = np.empty((N, Hd), dtype=np.float32)
    L.orc_gcn_fused_forward(_ptr(rp), _ptr(ci), _ptr(v), _ptr(X), _ptr(W), _ptr(H), _ptr(nn), ctypes.c_int(N),
                            ctypes.c_int(Fp), ctypes.c_int(int(actual_F)), ctypes.c_int(Hd), ctypes.c_int(ci.size))
    return H


def gat_forward(row_ptr, col, wh, el, er, heads, slope, mode="softmax", elu=False):
    L = _load()
    rp = np.ascontiguousarray(row_ptr, dtype=np.int64)
    c = np.ascontiguousarray(col, dtype=np.int32)
    wh, el, er = _f32(wh), _f32(el), _f32(er)
    n_dst, FD = rp.size - 1, wh.shape[1]
    out = np.empty((n_dst, FD), dtype=np.float32)
    L.orc_gat_forward(_ptr(rp), _ptr(c), _ptr(wh), ctypes.c_int64(FD), _ptr(el), _ptr(er), ctypes.c_int64(heads),
                      _ptr(out), ctypes.c_int64(FD), ctypes.c_int64(n_dst), ctypes.c_int(heads),
                      ctypes.c_int(FD // heads), ctypes.c_float(slope),
                      ctypes.c_int({"softmax": 0, "exp_neg": 1}[mode]), ctypes.c_int(EPI_ELU if elu else 0))
    return out


def sddmm_csr(row_ptr, col, a, b):
    L = _load()
    rp = np.ascontiguousarray(row_ptr, dtype=np.int64)
    c = np.ascontiguousarray(col, dtype=np.int32)
    a, b = _f32(a), _f32(b)
    out = np.empty(c.size, dtype=np.float32)
    L.orc_sddmm_csr(_ptr(rp), _ptr(c), _ptr(a), ctypes.c_int64(a.shape[1]), _ptr(b), ctypes.c_int64(b.shape[1]),
                    _ptr(out), ctypes.c_int64(rp.size - 1), ctypes.c_int(a.shape[1]))
    return out


def gather_rows(table, ids, host_table=None, gpu_flag=None, local2cache=None, nid_map=None):
    """Byte-exact row gather; returns (out, miss_count)."""
    L = _load()
    L.orc_gather_rows.restype = ctypes.c_int64
    table = None if table is None else np.ascontiguousarray(table)
    host_table = None if host_table is None else np.ascontiguousarray(host_table)
    ref = table if table is not None else host_table
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    out = np.empty((ids.size,) + ref.shape[1:], dtype=ref.dtype)
    rb = int(np.prod(ref.shape[1:], dtype=np.int64)) * ref.itemsize
    gf = None if gpu_flag is None else np.ascontiguousarray(gpu_flag, dtype=np.uint8)
    l2c = None if local2cache is None else np.ascontiguousarray(local2cache, dtype=np.int64)
    nm = None if nid_map is None else np.ascontiguousarray(nid_map, dtype=np.int64)
    miss = L.orc_gather_rows(_ptr(table), ctypes.c_int64(rb), _ptr(host_table), ctypes.c_int64(rb), _ptr(ids),
                             _ptr(gf), _ptr(l2c), _ptr(nm), _ptr(out), ctypes.c_int64(rb),
                             ctypes.c_int64(ids.size), ctypes.c_int64(rb))
    return out, int(miss)


def packed_words(F):
    return ((F + 31) // 32 + 3) // 4 * 4


def binarize_pack(x, wpr=None):
    L = _load()
    x = _f32(x)
    wpr = packed_words(x.shape[1]) if wpr is None else wpr
    packed = np.empty((x.shape[0], wpr), dtype=np.uint32)
    L.orc_binarize_pack(_ptr(x), ctypes.c_int64(x.shape[1]), _ptr(packed), ctypes.c_int64(wpr),
                        ctypes.c_int64(x.shape[0]), ctypes.c_int(x.shape[1]))
    return packed


def bin_spmm_counts(row_ptr, col, packed, F):
    L = _load()
    rp = np.ascontiguousarray(row_ptr, dtype=np.int64)
    c = np.ascontiguousarray(col, dtype=np.int32)
    packed = np.ascontiguousarray(packed, dtype=np.uint32)
    cnt = np.empty((rp.size - 1, F), dtype=np.int32)
    L.orc_bin_spmm_counts(_ptr(rp), _ptr(c), _ptr(packed), ctypes.c_int64(packed.shape[1]), _ptr(cnt),
                          ctypes.c_int64(rp.size - 1), ctypes.c_int(F))
    return cnt


def gemm(a, b, bias=None, relu=False, elu=False):
    L = _load()
    a, b = _f32(a), _f32(b)
    bs = None if bias is None else _f32(bias)
    M, K, N = a.shape[0], a.shape[1], b.shape[1]
    c = np.empty((M, N), dtype=np.float32)
    epi = (EPI_RELU if relu else 0) | (EPI_ELU if elu else 0)
    L.orc_gemm(_ptr(a), ctypes.c_int64(K), _ptr(b), ctypes.c_int64(N), _ptr(c), ctypes.c_int64(N), ctypes.c_int64(M),
               ctypes.c_int64(N), ctypes.c_int64(K), _ptr(bs), ctypes.c_int(epi))
    return c


def coo_to_csr(rows, cols, n_rows, values=None):
    """Stable COO -> CSR by row (row = destination); keeps duplicate edges like torch.sparse.mm's sum."""
    rows = np.asarray(rows, dtype=np.int64)
    order = np.argsort(rows, kind="stable")
    rp = np.zeros(n_rows + 1, dtype=np.int64)
    np.add.at(rp, rows + 1, 1)
    rp = np.cumsum(rp)
    col = np.asarray(cols, dtype=np.int64)[order].astype(np.int32)
    val = None if values is None else np.asarray(values, dtype=np.float32)[order]
    return rp, col, val
