"""Autograd-aware operators over the C-ABI kernels: the layer classes in ``dgll_b200.nn`` are built from these.

Every forward AND backward runs on the hand-written kernels (``kernels.py`` -> ``libdgll_b200.so``); torch is used
for memory, streams, elementwise glue on [n, F] tensors and autograd bookkeeping.  There is no CPU path: CPU tensors
raise.

  CsrGraph        static CSR by destination (+ lazily built transpose / nnz-split plan), converted once from the
                  reference's adjacency formats (torch sparse COO as gcnconv.py:31 takes, dense 0/1 as gatconv.py:34
                  / :115 take, edge_index as Evaluation/PPI/gcn_model.py:44-57 builds)
  spmm            out = epi(reduce_e(values[e] * x[col[e]]) + bias)         (gcnconv.py:31, gcn_model.py:76, sageconv.py:32-38)
  linear / mm     dense transform x @ W (+ b)                               (gcnconv.py:30, gatconv.py:117, sageconv.py:40,72)
  gat_aggregate   fused SDDMM + edge softmax + aggregation                  (gatconv.py:30-54, 111-148)
  gather_rows     feature row gather                                        (dgraph.py:105, storage.py:185-209)
"""
import torch

from . import kernels as K

_PRECISION = {"gemm": "fp32"}


def set_gemm_precision(p):
    """'fp32' = exact SIMT fp32 FMA (parity path, <=1e-5); 'tf32' = tcgen05 tensor cores reading the fp32 operands in
    place (TF32, fp32 accumulation in TMEM, <=2e-3); 'bf16' = tcgen05 on operands packed to bf16 (<=1e-2)."""
    if p not in ("fp32", "bf16", "tf32", "tf32x3"):
        raise ValueError("precision must be 'fp32', 'tf32x3', 'tf32' or 'bf16'")
    _PRECISION["gemm"] = p


def get_gemm_precision():
    return _PRECISION["gemm"]


# ------------------------------------------------------------------ graph ---
class CsrGraph:
    """CSR by destination row: ``row_ptr[n_dst+1]``, ``col[nnz]`` (source ids), optional ``values[nnz]``."""

    HEAVY_ROW = 4096  # rows longer than this get an nnz-split plan
    PLAN_MIN_EDGES = 1 << 22

    def __init__(self, row_ptr, col, values=None, n_src=None):
        if not row_ptr.is_cuda:
            raise RuntimeError("dgll_b200: CsrGraph needs CUDA tensors (there is no CPU fallback)")
        self.row_ptr = row_ptr.contiguous()
        self.col = None if col is None else (col if col.dtype == torch.int32 else col.to(torch.int32)).contiguous()
        self.values = None if values is None else values.to(torch.float32).contiguous()
        self.n_dst = self.row_ptr.numel() - 1
        self.n_src = int(n_src) if n_src is not None else self.n_dst
        self._t = None
        self._perm = None
        self._plan = False
        self._deg = None
        self._inv_deg = None

    @property
    def nnz(self):
        return int(self.col.numel()) if self.col is not None else int(self.row_ptr[-1].item())

    @property
    def device(self):
        return self.row_ptr.device

    def degrees(self):
        if self._deg is None:
            self._deg = (self.row_ptr[1:] - self.row_ptr[:-1]).to(torch.float32)
            self._inv_deg = torch.where(self._deg > 0, 1.0 / self._deg.clamp(min=1), torch.zeros_like(self._deg))
        return self._deg

    def inv_degrees(self):
        self.degrees()
        return self._inv_deg

    @staticmethod
    def _may_probe():
        """Plan probing reads one scalar back (a device synchronisation): never while a CUDA graph is being captured —
        a capture-time graph simply runs without a split plan (fixed-capacity blocks have bounded rows anyway)."""
        return not torch.cuda.is_current_stream_capturing()

    def plan(self):
        """nnz-split schedule when the graph has very long rows (Reddit-shaped skew); None otherwise."""
        if self._plan is False:
            if not self._may_probe():
                return None
            self._plan = None
            # only graphs big enough to have such rows are inspected (the check reads one scalar back = one sync
            # per STATIC graph); sampled blocks never pay it
            if self.n_dst > 0 and self.col is not None and self.col.numel() >= self.PLAN_MIN_EDGES and \
                    float(self.degrees().max().item()) > self.HEAVY_ROW:
                self._plan = K.CsrPlan(self.row_ptr, chunk_edges=self.HEAVY_ROW)
        return self._plan

    def bin_plan(self):
        """nnz-split into 1,024-edge items: the binarized kernel (integer atomics, exact) and the fused GAT forward
        (partial softmax states merged exactly)."""
        if getattr(self, "_bin_plan", False) is False:
            if not self._may_probe():
                return None
            self._bin_plan = None
            if self.n_dst > 0 and self.col is not None and self.col.numel() >= self.PLAN_MIN_EDGES and \
                    float(self.degrees().max().item()) > 1024:
                self._bin_plan = K.CsrPlan(self.row_ptr, chunk_edges=1024)
        return self._bin_plan

    def gat_plan(self):
        """nnz-split for the two GAT backward passes (256-edge items: 55.5 ms vs 57.5 ms at 1,024 and 84.1 ms unsplit on
        the products-shaped graph; the forward is best with the 1,024-edge ``bin_plan``: 19.7 ms vs 20.9 ms —
        profiles/r01_kernels.jsonl)."""
        if getattr(self, "_gat_plan", False) is False:
            if not self._may_probe():
                return None
            self._gat_plan = None
            if self.n_dst > 0 and self.col is not None and self.col.numel() >= self.PLAN_MIN_EDGES and \
                    float(self.degrees().max().item()) > 256:
                self._gat_plan = K.CsrPlan(self.row_ptr, chunk_edges=256)
        return self._gat_plan

    def transpose(self):
        """CsrGraph of A^T (values carried along); ``perm[e_T] = e`` kept for per-edge gradients."""
        if self._t is None:
            t_rp, t_col, t_val, perm = K.csr_transpose(self.row_ptr, self.col, self.n_src, values=self.values,
                                                       want_perm=True)
            self._t = CsrGraph(t_rp, t_col, t_val, n_src=self.n_dst)
            self._perm = perm
        return self._t

    def with_values(self, values):
        g = CsrGraph(self.row_ptr, self.col, values, n_src=self.n_src)
        g._deg, g._inv_deg = self._deg, self._inv_deg
        return g

    # -- constructors from the reference's adjacency formats --
    @staticmethod
    def from_coo(rows, cols, n_dst, n_src=None, values=None):
        """Stable COO -> CSR by row.  Duplicate (row, col) pairs are kept: they sum, as torch.sparse.mm does on an
        uncoalesced COO (Evaluation/PPI/gcn_model.py:44-57 builds exactly that)."""
        rows = rows.to(torch.int64)
        order = torch.sort(rows, stable=True).indices
        counts = torch.bincount(rows, minlength=n_dst)
        row_ptr = torch.zeros(n_dst + 1, dtype=torch.int64, device=rows.device)
        torch.cumsum(counts, 0, out=row_ptr[1:])
        col = cols[order].to(torch.int32)
        vals = None if values is None else values[order]
        return CsrGraph(row_ptr, col, vals, n_src=n_src if n_src is not None else n_dst)

    @staticmethod
    def from_edge_index(edge_index, n, values=None):
        """edge_index int64[2,E]; row index = edge_index[0] (gcn_model.py:56)."""
        return CsrGraph.from_coo(edge_index[0], edge_index[1], n, n, values)

    @staticmethod
    def from_torch_sparse(adj):
        if adj.layout == torch.sparse_csr:
            return CsrGraph(adj.crow_indices(), adj.col_indices().to(torch.int32), adj.values(), n_src=adj.size(1))
        idx = adj._indices() if not adj.is_coalesced() else adj.indices()
        val = adj._values() if not adj.is_coalesced() else adj.values()
        return CsrGraph.from_coo(idx[0], idx[1], adj.size(0), adj.size(1), val)

    @staticmethod
    def from_dense(adj):
        """Edges = positions with adj > 0 in row-major order (gatconv.py:34 ``adj > 0``; :115 ``adj.nonzero()``)."""
        nz = (adj > 0).nonzero()
        return CsrGraph.from_coo(nz[:, 0], nz[:, 1], adj.size(0), adj.size(1), None)


class _Key:
    """holder attached to the adjacency tensor so a converted graph lives exactly as long as that tensor object."""
    __slots__ = ("graph", "version")


def as_csr(adj, binary=False):
    """Accepts CsrGraph | torch sparse COO/CSR | dense [n, m] | (row_ptr, col[, values]).  Conversions of tensor
    inputs are cached per tensor object (the reference passes the same ``adj`` every step)."""
    if isinstance(adj, CsrGraph):
        return adj
    if isinstance(adj, (tuple, list)):
        return CsrGraph(*adj)
    if not isinstance(adj, torch.Tensor):
        raise TypeError("unsupported adjacency type %r" % type(adj))
    if not adj.is_cuda:
        raise RuntimeError("dgll_b200: adjacency must live on a CUDA device (there is no CPU fallback)")
    holder = getattr(adj, "_dgllb_csr", None)
    if adj.layout == torch.strided:
        ver = adj._version
    elif adj.layout == torch.sparse_csr:
        ver = adj.values()._version
    else:                                   # in-place edits of a sparse tensor's values bump the values' version
        ver = adj._values()._version
    if holder is not None and holder.version == (ver, binary):
        return holder.graph
    if adj.layout == torch.strided:
        g = CsrGraph.from_dense(adj)
        if not binary:
            nz = (adj > 0).nonzero()
            g = g.with_values(adj[nz[:, 0], nz[:, 1]])
    else:
        g = CsrGraph.from_torch_sparse(adj)
        if binary:
            g = g.with_values(None)
    holder = _Key()
    holder.graph, holder.version = g, (ver, binary)
    try:
        adj._dgllb_csr = holder
    except Exception:
        pass
    return g


# ------------------------------------------------------------------- SpMM ---
class _SpmmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, values, bias, graph, reduce, relu, F, addend=None):
        g = graph if values is None else graph.with_values(values)
        want_argmax = reduce == "max" and (x.requires_grad or (values is not None and values.requires_grad))
        plan = graph.plan() if reduce != "max" else None
        res = K.spmm_csr(g.row_ptr, g.col, x, values=g.values, reduce=reduce, n_dst=g.n_dst, bias=bias, relu=relu,
                         return_argmax=want_argmax, plan=plan, F=F, addend=addend)
        out, argmax = res if want_argmax else (res, None)
        ctx.graph, ctx.reduce, ctx.relu = graph, reduce, relu
        ctx.has_values, ctx.has_bias, ctx.has_addend = values is not None, bias is not None, addend is not None
        ctx.bias_ref = bias
        ctx.save_for_backward(x, values, out if relu else None, argmax)
        return out

    @staticmethod
    def backward(ctx, grad):
        x, values, out, argmax = ctx.saved_tensors
        graph = ctx.graph
        g = grad.contiguous()
        if ctx.relu:
            g = g * (out > 0)
        gx = gv = gb = None
        ga = g if (ctx.has_addend and ctx.needs_input_grad[7]) else None   # d out / d addend = identity (after the ReLU mask)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = _bias_grad(g, ctx)
        if ctx.reduce == "max":
            if ctx.needs_input_grad[0]:
                if ctx.has_values:
                    raise NotImplementedError("max aggregation with learnable edge values")
                gx = K.spmm_max_backward(graph.col, argmax, g, x.size(0))
                if gx.size(1) != x.size(1):
                    gx = torch.nn.functional.pad(gx, (0, x.size(1) - gx.size(1)))
            return gx, None, gb, None, None, None, None, ga
        if ctx.reduce == "mean":
            g = g * graph.inv_degrees()[:, None]
        if ctx.needs_input_grad[0]:
            gt = graph.transpose()
            tvals = None
            if ctx.has_values:
                tvals = values.detach()[graph._perm.long()]
            elif gt.values is not None:
                tvals = gt.values
            gx = K.spmm_csr(gt.row_ptr, gt.col, g, values=tvals, reduce="sum", n_dst=gt.n_dst,
                            plan=gt.plan())
            if gx.size(1) != x.size(1):
                gx = torch.nn.functional.pad(gx, (0, x.size(1) - gx.size(1)))
            if gx.dtype != x.dtype:
                gx = gx.to(x.dtype)
        if ctx.has_values and ctx.needs_input_grad[1]:
            # d values[e] = <g[row(e)], x[col[e]]>  — the SDDMM SpecialSpmmFunction.backward computes densely (gatconv.py:76-78)
            gv = K.sddmm_csr(graph.row_ptr, graph.col, g, x[:, :g.size(1)].float())
        return gx, gv, gb, None, None, None, None, ga


def spmm(adj, x, values=None, reduce="sum", bias=None, relu=False, F=None, addend=None):
    """Neighbourhood aggregation with autograd: ``epi(reduce_e(values[e] * x[col[e]]) + addend + bias)``.  ``values``
    overrides the graph's edge values (and may require grad); ``addend`` [n_dst, F] is added in the kernel's epilogue
    (the self term of a SAGE layer) and receives the output gradient."""
    graph = as_csr(adj)
    if values is None and graph.values is not None:
        values_arg = None  # static values ride inside the graph object
    else:
        values_arg = values
    return _SpmmFn.apply(x, values_arg, bias, graph, reduce, relu, F, addend)


# ----------------------------------------------------------------- linear ---
_DIRECT_GRADS = {"on": False}


class direct_weight_grads:
    """Context manager for trainers that own pre-allocated, pre-zeroed ``.grad`` buffers (``pipelined``): inside it the
    weight-gradient GEMMs of ``linear`` / ``linear2`` ACCUMULATE straight into ``w.grad`` (split-K reduction included) and
    return no gradient, and bias gradients are added in place — autograd's per-parameter accumulate kernels disappear."""

    def __enter__(self):
        self._prev = _DIRECT_GRADS["on"]
        _DIRECT_GRADS["on"] = True

    def __exit__(self, *a):
        _DIRECT_GRADS["on"] = self._prev


def _weight_grad(a, b, w, precision, trans_a=True):
    """dW = a^T b as a GEMM; written into ``w.grad`` when direct gradient writes are on (returns None then)."""
    if _DIRECT_GRADS["on"] and w.grad is not None and w.grad.is_contiguous():
        K.gemm(a, b, trans_a=trans_a, out=w.grad, accumulate=True, precision=precision)
        return None
    return K.gemm(a, b, trans_a=trans_a, precision=precision)


def _bias_grad(g, ctx_or_bias):
    bias = getattr(ctx_or_bias, "bias_ref", None) if not torch.is_tensor(ctx_or_bias) else ctx_or_bias
    gb = g.sum(0)
    if _DIRECT_GRADS["on"] and bias is not None and bias.grad is not None:
        bias.grad.add_(gb)
        return None
    return gb


class _Linear2Fn(torch.autograd.Function):
    """``act(x1 @ op(w1) + x2 @ op(w2) + bias)`` — the two transforms of a SAGE layer (self + neighbour) as two GEMMs
    into ONE output (the second accumulates, bias and ReLU ride in its epilogue); no add / activation kernels."""

    @staticmethod
    def forward(ctx, x1, w1, x2, w2, bias, relu, precision, trans_w):
        out = K.gemm(x1, w1, trans_b=trans_w, precision=precision)
        K.gemm(x2, w2, bias=bias, relu=relu, trans_b=trans_w, out=out, accumulate=True, precision=precision)
        ctx.relu, ctx.precision, ctx.trans_w = relu, precision, trans_w
        ctx.bias_ref = bias
        ctx.save_for_backward(x1, w1, x2, w2, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, grad):
        x1, w1, x2, w2, out = ctx.saved_tensors
        g = grad.contiguous()
        if ctx.relu:
            g = g * (out > 0)
        p, tw = ctx.precision, ctx.trans_w
        gx1 = gx2 = gw1 = gw2 = gb = None
        if ctx.needs_input_grad[0]:
            gx1 = K.gemm(g, w1, trans_b=not tw, precision=p)
        if ctx.needs_input_grad[2]:
            gx2 = K.gemm(g, w2, trans_b=not tw, precision=p)
        if ctx.needs_input_grad[1]:
            gw1 = _weight_grad(g, x1, w1, p) if tw else _weight_grad(x1, g, w1, p)
        if ctx.needs_input_grad[3]:
            gw2 = _weight_grad(g, x2, w2, p) if tw else _weight_grad(x2, g, w2, p)
        if ctx.bias_ref is not None and ctx.needs_input_grad[4]:
            gb = _bias_grad(g, ctx.bias_ref)
        return gx1, gw1, gx2, gw2, gb, None, None, None


def linear2(x1, w1, x2, w2, bias=None, relu=False, precision=None, trans_w=False):
    """``act(x1 @ w1 + x2 @ w2 + bias)`` on two GEMMs into one output (see ``_Linear2Fn``)."""
    return _Linear2Fn.apply(x1, w1, x2, w2, bias, relu, precision or _PRECISION["gemm"], trans_w)


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, relu, precision, trans_w):
        out = K.gemm(x, w, bias=bias, relu=relu, trans_b=trans_w, precision=precision)
        ctx.relu, ctx.precision, ctx.trans_w = relu, precision, trans_w
        ctx.bias_ref = bias
        ctx.save_for_backward(x, w, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, grad):
        x, w, out = ctx.saved_tensors
        g = grad.contiguous()
        if ctx.relu:
            g = g * (out > 0)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = K.gemm(g, w, trans_b=not ctx.trans_w, precision=ctx.precision)      # dX = G W^T
        if ctx.needs_input_grad[1]:
            if ctx.trans_w:
                gw = _weight_grad(g, x, w, ctx.precision)                             # dW[N,K] = G^T X
            else:
                gw = _weight_grad(x, g, w, ctx.precision)                             # dW[K,N] = X^T G
        if ctx.bias_ref is not None and ctx.needs_input_grad[2]:
            gb = _bias_grad(g, ctx.bias_ref)
        return gx, gw, gb, None, None, None


def linear(x, w, bias=None, relu=False, precision=None, trans_w=False):
    """x[M,K] @ w[K,N] (+ bias)(relu) on the device GEMM (fp32 exact or tcgen05 bf16).  ``trans_w=True`` takes ``w``
    stored [N,K] (an ``nn.Linear`` weight) without materialising its transpose."""
    lead = None
    if x.dim() > 2:
        lead = x.shape[:-1]
        x = x.reshape(-1, x.size(-1))
    out = _LinearFn.apply(x, w, bias, relu, precision or _PRECISION["gemm"], trans_w)
    return out if lead is None else out.reshape(*lead, out.size(-1))


def mm(a, b):
    return linear(a, b)


# -------------------------------------------------------------------- GAT ---
class _GatFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, wh, el, er, graph, heads, slope, mode, dropout, seed):
        out, rmax, rsum = K.gat_forward(graph.row_ptr, graph.col, wh, el, er, heads, slope, mode=mode,
                                        save_stats=True, n_dst=graph.n_dst, plan=graph.bin_plan(), dropout=dropout,
                                        seed=seed)
        ctx.graph, ctx.heads, ctx.slope, ctx.mode = graph, heads, slope, mode
        ctx.dropout, ctx.seed = dropout, seed
        ctx.save_for_backward(wh, el, er, out, rmax, rsum)
        return out

    @staticmethod
    def backward(ctx, grad):
        wh, el, er, out, rmax, rsum = ctx.saved_tensors
        graph = ctx.graph
        gt = graph.transpose()
        d_wh, d_el, d_er = K.gat_backward(graph.row_ptr, graph.col, gt.row_ptr, gt.col, graph._perm, wh, el, er, out,
                                          rmax, rsum, grad.contiguous(), ctx.heads, ctx.slope, mode=ctx.mode,
                                          dropout=ctx.dropout, seed=ctx.seed, plan=graph.gat_plan(),
                                          t_plan=gt.gat_plan())
        return d_wh, d_el, d_er, None, None, None, None, None, None


class _GatExtFn(torch.autograd.Function):
    """GAT aggregation on ONE activation buffer ``ext = [Wh | el | er | pad]`` ([n, heads*D + 2*heads (+pad)], produced by
    a single dense transform): the kernels read the three parts in place through leading dimensions, and the backward
    writes d_Wh, d_el, d_er into one buffer of the same layout, so the transform's backward is one pair of GEMMs and no
    slice / cat / broadcast kernels run.  Square graphs only (el and er index the same node set)."""

    @staticmethod
    def forward(ctx, ext, graph, heads, D, slope, mode, dropout, seed):
        FD = heads * D
        wh, el, er = ext[:, :FD], ext[:, FD:FD + heads], ext[:, FD + heads:FD + 2 * heads]
        out, rmax, rsum = K.gat_forward(graph.row_ptr, graph.col, wh, el, er, heads, slope, mode=mode,
                                        save_stats=True, n_dst=graph.n_dst, plan=graph.bin_plan(), dropout=dropout,
                                        seed=seed)
        ctx.graph, ctx.heads, ctx.D, ctx.slope, ctx.mode = graph, heads, D, slope, mode
        ctx.dropout, ctx.seed = dropout, seed
        ctx.save_for_backward(ext, out, rmax, rsum)
        return out

    @staticmethod
    def backward(ctx, grad):
        ext, out, rmax, rsum = ctx.saved_tensors
        graph, heads, FD = ctx.graph, ctx.heads, ctx.heads * ctx.D
        wh, el, er = ext[:, :FD], ext[:, FD:FD + heads], ext[:, FD + heads:FD + 2 * heads]
        gt = graph.transpose()
        d_ext = torch.empty_like(ext)
        if ext.size(1) > FD + 2 * heads:
            d_ext[:, FD + 2 * heads:].zero_()
        K.gat_backward(graph.row_ptr, graph.col, gt.row_ptr, gt.col, graph._perm, wh, el, er, out, rmax, rsum,
                       grad.contiguous(), heads, ctx.slope, mode=ctx.mode, d_ext=d_ext, dropout=ctx.dropout,
                       seed=ctx.seed, plan=graph.gat_plan(), t_plan=gt.gat_plan())
        return d_ext, None, None, None, None, None, None, None


def gat_aggregate_ext(adj, ext, heads, D, slope=0.2, mode="softmax", elu=False, dropout=0.0, seed=None):
    """``gat_aggregate`` for ``ext = [Wh | el | er | pad]`` rows (see ``_GatExtFn``); returns [n, heads*D]."""
    graph = as_csr(adj, binary=True)
    if graph.n_dst != graph.n_src or ext.size(0) != graph.n_src:
        raise ValueError("gat_aggregate_ext: square graphs only")
    if dropout and seed is None:
        seed = _fresh_seed()
    seed = seed or 0
    FD = heads * D
    if torch.is_grad_enabled() and ext.requires_grad:
        out = _GatExtFn.apply(ext, graph, heads, D, slope, mode, float(dropout), seed)
        return torch.nn.functional.elu(out) if elu else out
    return K.gat_forward(graph.row_ptr, graph.col, ext[:, :FD], ext[:, FD:FD + heads], ext[:, FD + heads:FD + 2 * heads],
                         heads, slope, mode=mode, elu=elu, n_dst=graph.n_dst, plan=graph.bin_plan(),
                         dropout=float(dropout), seed=seed)


def _fresh_seed():
    """64-bit seed from torch's CPU generator (follows torch.manual_seed; no device synchronisation)."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def gat_aggregate(adj, wh, el, er, heads=1, slope=0.2, mode="softmax", elu=False, dropout=0.0, seed=None):
    """Fused SDDMM + edge-softmax + aggregation.  ``wh`` [n_src, heads*D]; ``el`` [n_dst, heads]; ``er`` [n_src, heads].
    ``dropout`` > 0 drops attention coefficients inside the kernel (mask = f(seed, edge, head), replayed by the
    backward) exactly where the reference applies ``F.dropout`` to them (gatconv.py:37, :132)."""
    graph = as_csr(adj, binary=True)
    if dropout and seed is None:
        seed = _fresh_seed()
    seed = seed or 0
    if torch.is_grad_enabled() and (wh.requires_grad or el.requires_grad or er.requires_grad):
        out = _GatFn.apply(wh, el, er, graph, heads, slope, mode, float(dropout), seed)
        return torch.nn.functional.elu(out) if elu else out
    return K.gat_forward(graph.row_ptr, graph.col, wh, el, er, heads, slope, mode=mode, elu=elu, n_dst=graph.n_dst,
                         plan=graph.bin_plan(), dropout=float(dropout), seed=seed)


# -------------------------------------------------------------- binarized ---
def binarized_aggregate(adj, x=None, packed=None, F=None, mode="mean"):
    """Binarized neighbourhood aggregation (README.md:11; semantics SURVEY.md §8 a18): features are reduced to their
    sign bit (``x >= 0``), bit-packed 32 per word, and aggregated with the bit-sliced popcount kernel.
    ``mode``: 'count' (int32 #neighbours with the bit set), 'sum' / 'mean' of the +-1 values (fp32).
    Pass ``packed`` (from ``kernels.binarize_pack``) to reuse a packed table across layers/steps.  Forward only."""
    graph = as_csr(adj, binary=True)
    if packed is None:
        packed = K.binarize_pack(x)
        F = x.size(1)
    plan = graph.bin_plan()
    return K.bin_spmm_csr(graph.row_ptr, graph.col, packed, F, mode=mode, n_dst=graph.n_dst, plan=plan)


class _BinAggFn(torch.autograd.Function):
    """mean_{j in N(i)} sign(x_j) on the bit-packed popcount kernel, with the straight-through estimator in backward:
    d sign(x)/dx := 1 where |x| <= clip (everywhere when clip is None), so dL/dx = A_mean^T g masked by |x| <= clip —
    one transposed aggregation on the fp32 SpMM kernels."""

    @staticmethod
    def forward(ctx, x, graph, clip):
        packed = K.binarize_pack(x)
        out = K.bin_spmm_csr(graph.row_ptr, graph.col, packed, x.size(1), mode="mean", n_dst=graph.n_dst,
                             plan=graph.bin_plan())
        ctx.graph, ctx.clip = graph, clip
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, grad):
        (x,) = ctx.saved_tensors
        graph = ctx.graph
        g = grad.contiguous() * graph.inv_degrees()[:, None]
        gt = graph.transpose()
        gx = K.spmm_csr(gt.row_ptr, gt.col, g, reduce="sum", n_dst=gt.n_dst, plan=gt.plan())
        if ctx.clip is not None:
            gx = gx * (x.abs() <= ctx.clip)
        return gx, None, None


def binarized_aggregate_ste(adj, x, clip=1.0):
    """Differentiable binarized mean aggregation (README.md:11): forward = ``binarized_aggregate(...,
    mode='mean')`` of the sign bits of ``x``; backward = straight-through estimator (see ``_BinAggFn``)."""
    graph = as_csr(adj, binary=True)
    if torch.is_grad_enabled() and x.requires_grad:
        return _BinAggFn.apply(x, graph, clip)
    return binarized_aggregate(graph, x=x, mode="mean")


# ----------------------------------------------------------------- gather ---
class _GatherFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, ids):
        ctx.n = table.size(0)
        ctx.save_for_backward(ids)
        return K.gather_rows(table, ids)

    @staticmethod
    def backward(ctx, grad):
        (ids,) = ctx.saved_tensors
        # scatter-add of gradient rows = aggregation over the transposed "gather graph": rows sorted by id
        order = torch.sort(ids.to(torch.int64), stable=True).indices
        counts = torch.bincount(ids.to(torch.int64), minlength=ctx.n)
        rp = torch.zeros(ctx.n + 1, dtype=torch.int64, device=ids.device)
        torch.cumsum(counts, 0, out=rp[1:])
        g2 = grad.reshape(grad.size(0), -1).contiguous()
        out = K.spmm_csr(rp, order.to(torch.int32), g2, reduce="sum", n_dst=ctx.n)
        return out.reshape((ctx.n,) + tuple(grad.shape[1:])), None


def gather_rows(table, ids):
    """``table[ids]`` through the TMA row-gather kernel (bit-exact copy); differentiable w.r.t. ``table``."""
    if table.requires_grad and torch.is_grad_enabled():
        return _GatherFn.apply(table, ids)
    return K.gather_rows(table, ids)


# ---------------------------------------------------------------- pooling ---
def segment_reduce(x, batch, size=None, reduce="sum"):
    """scatter(x, batch, dim=0, reduce=...) for a SORTED batch vector = segment reduce (Pooling.py:37,59,81).
    An unsorted ``batch`` is sorted first (stable), which keeps the result identical to scatter()."""
    if batch is None:
        rp = torch.tensor([0, x.size(0)], dtype=torch.int64, device=x.device)
        col = torch.arange(x.size(0), device=x.device, dtype=torch.int32)
        return spmm(CsrGraph(rp, col, n_src=x.size(0)), x, reduce=reduce)
    batch = batch.to(torch.int64)
    size = int(batch.max().item() + 1) if size is None else int(size)
    if batch.numel() > 1 and bool((batch[1:] < batch[:-1]).any().item()):
        order = torch.sort(batch, stable=True).indices
        x = gather_rows(x, order)
        batch = batch[order]
    counts = torch.bincount(batch, minlength=size)
    rp = torch.zeros(size + 1, dtype=torch.int64, device=x.device)
    torch.cumsum(counts, 0, out=rp[1:])
    g = CsrGraph(rp, torch.arange(x.size(0), device=x.device, dtype=torch.int32), n_src=x.size(0))
    return spmm(g, x, reduce=reduce)
