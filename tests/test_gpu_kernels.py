"""GPU parity tests proper: every kernel is called through the C ABI (ctypes) and compared with the CPU oracle
on the same seeded inputs.  Tolerances (BASELINE.json north_star): bit-exact for integer / byte / index work,
<= 1e-5 relative (fp32) and <= 1e-2 (bf16) for aggregated features."""
import numpy as np
import pytest
import torch

import oracle
from oracle import samplers as S
from conftest import rel_err

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
BF16_TOL = 1e-2


@pytest.fixture(scope="module")
def K():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from dgll_b200 import kernels
    return kernels


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rand_csr(rng, n_dst, n_src, max_deg, heavy=(), empty_frac=0.1, sort_cols=False):
    deg = rng.integers(1, max_deg + 1, size=n_dst)
    deg[rng.random(n_dst) < empty_frac] = 0
    for r, d in heavy:
        deg[r] = d
    rp = np.zeros(n_dst + 1, dtype=np.int64)
    rp[1:] = np.cumsum(deg)
    col = rng.integers(0, n_src, size=int(rp[-1])).astype(np.int32)
    if sort_cols:
        for i in range(n_dst):
            col[rp[i]:rp[i + 1]].sort()
    return rp, col


def padded(x, ld):
    t = torch.zeros((x.shape[0], ld), dtype=torch.float32, device="cuda")
    t[:, :x.shape[1]] = dev(x)
    return t


# ---------------------------------------------------------------- SpMM ------
@pytest.mark.parametrize("F,ld", [(602, 604), (602, 602), (256, 256), (64, 64), (50, 52), (100, 100), (7, 7),
                                  (128, 128), (1, 1), (33, 36)])
@pytest.mark.parametrize("reduce", ["sum", "mean", "max"])
def test_spmm_reduce_parity(K, F, ld, reduce):
    rng = np.random.default_rng(F * 7 + len(reduce))
    n_dst, n_src = 700, 1500
    rp, col = rand_csr(rng, n_dst, n_src, 60, heavy=[(3, 900)])
    x = rng.standard_normal((n_src, F)).astype(np.float32)
    vals = rng.random(col.size).astype(np.float32) + 0.1
    tx = padded(x, ld)
    for v in (None, vals):
        ref = oracle.spmm_csr(rp, col, x, values=v, reduce=reduce)
        out = K.spmm_csr(dev(rp), dev(col), tx, values=None if v is None else dev(v), reduce=reduce, F=F)
        assert out.shape == (n_dst, F)
        assert rel_err(out.cpu().numpy(), ref) <= FP32_TOL


def test_spmm_int32_rowptr_and_epilogue(K):
    rng = np.random.default_rng(11)
    n_dst, n_src, F = 513, 900, 96
    rp, col = rand_csr(rng, n_dst, n_src, 30)
    x = rng.standard_normal((n_src, F)).astype(np.float32)
    vals = rng.standard_normal(col.size).astype(np.float32)
    rs = rng.random(n_dst).astype(np.float32)
    add = rng.standard_normal((n_dst, F)).astype(np.float32)
    bias = rng.standard_normal(F).astype(np.float32)
    for relu, elu in ((True, False), (False, True), (False, False)):
        ref = oracle.spmm_csr(rp, col, x, values=vals, reduce="sum", row_scale=rs, addend=add, bias=bias,
                              relu=relu, elu=elu)
        out = K.spmm_csr(dev(rp.astype(np.int32)), dev(col), dev(x), values=dev(vals), reduce="sum",
                         row_scale=dev(rs), addend=dev(add), bias=dev(bias), relu=relu, elu=elu)
        assert rel_err(out.cpu().numpy(), ref) <= FP32_TOL


def test_spmm_max_argmax_and_backward(K):
    rng = np.random.default_rng(5)
    n_dst, n_src, F = 300, 500, 40
    rp, col = rand_csr(rng, n_dst, n_src, 25)
    x = rng.standard_normal((n_src, F)).astype(np.float32)
    ref, ref_am = oracle.spmm_csr(rp, col, x, reduce="max", return_argmax=True)
    out, am = K.spmm_csr(dev(rp), dev(col), dev(x), reduce="max", return_argmax=True)
    assert np.array_equal(out.cpu().numpy(), ref)  # max of copies is exact
    am = am.cpu().numpy()
    assert np.array_equal(am, ref_am)
    g = rng.standard_normal((n_dst, F)).astype(np.float32)
    gx = K.spmm_max_backward(dev(col), dev(am), dev(g), n_src).cpu().numpy()
    ref_gx = np.zeros((n_src, F), dtype=np.float64)
    for i in range(n_dst):
        for f in range(F):
            if ref_am[i, f] >= 0:
                ref_gx[col[ref_am[i, f]], f] += g[i, f]
    assert rel_err(gx, ref_gx) <= FP32_TOL


@pytest.mark.parametrize("chunk", [64, 256])
def test_spmm_plan_heavy_rows(K, chunk):
    rng = np.random.default_rng(chunk)
    n_dst, n_src, F = 400, 3000, 200
    rp, col = rand_csr(rng, n_dst, n_src, 40, heavy=[(0, 5000), (17, 777), (399, 2049)])
    x = rng.standard_normal((n_src, F)).astype(np.float32)
    vals = rng.random(col.size).astype(np.float32)
    bias = rng.standard_normal(F).astype(np.float32)
    plan = K.CsrPlan(dev(rp), chunk_edges=chunk)
    deg = rp[1:] - rp[:-1]
    assert plan.n_heavy_rows == int((deg > chunk).sum())
    assert plan.n_chunks == int(sum((d + chunk - 1) // chunk for d in deg if d > chunk))
    for reduce in ("sum", "mean"):
        ref = oracle.spmm_csr(rp, col, x, values=vals, reduce=reduce, bias=bias, relu=True)
        out = K.spmm_csr(dev(rp), dev(col), dev(x), values=dev(vals), reduce=reduce, bias=dev(bias), relu=True,
                         plan=plan)
        assert rel_err(out.cpu().numpy(), ref) <= FP32_TOL
        # split rows are merged in chunk order from a workspace (no atomics): run-to-run bit-identical, also for an
        # output whose rows are not 16-byte aligned and for bf16 features
        for _ in range(3):
            assert torch.equal(K.spmm_csr(dev(rp), dev(col), dev(x), values=dev(vals), reduce=reduce, bias=dev(bias),
                                          relu=True, plan=plan), out)
        odd = torch.zeros((n_dst, F + 1), device="cuda")[:, 1:]
        K.spmm_csr(dev(rp), dev(col), dev(x), values=dev(vals), reduce=reduce, bias=dev(bias), relu=True, plan=plan,
                   out=odd)
        assert torch.equal(odd, out)
        xb = dev(x).to(torch.bfloat16)
        ob = K.spmm_csr(dev(rp), dev(col), xb, values=dev(vals), reduce=reduce, plan=plan)
        assert torch.equal(K.spmm_csr(dev(rp), dev(col), xb, values=dev(vals), reduce=reduce, plan=plan), ob)
        assert rel_err(ob.cpu().numpy(), oracle.spmm_csr(rp, col, xb.float().cpu().numpy(), values=vals,
                                                         reduce=reduce)) <= FP32_TOL


def test_spmm_bf16_features(K):
    rng = np.random.default_rng(2)
    n_dst, n_src, F = 600, 1000, 608
    rp, col = rand_csr(rng, n_dst, n_src, 50)
    x = rng.standard_normal((n_src, F)).astype(np.float32)
    xb = dev(x).to(torch.bfloat16)
    ref_exact = oracle.spmm_csr(rp, col, xb.float().cpu().numpy(), reduce="mean")
    out = K.spmm_csr(dev(rp), dev(col), xb, reduce="mean")
    # same bf16-rounded inputs, fp32 accumulate: matches to fp32 tolerance ...
    assert rel_err(out.cpu().numpy(), ref_exact) <= FP32_TOL
    # ... and the fp32 oracle to the bf16 tolerance of north_star
    assert rel_err(out.cpu().numpy(), oracle.spmm_csr(rp, col, x, reduce="mean")) <= BF16_TOL


def test_spmm_segment_reduce_pooling(K):
    """col_idx=None: scatter()-style pooling over a sorted batch vector (GlobalPooling/Pooling.py:18-81)."""
    from oracle import layers as L
    rng = np.random.default_rng(9)
    sizes = rng.integers(0, 50, size=40)
    batch = np.repeat(np.arange(40), sizes)
    x = rng.standard_normal((batch.size, 24)).astype(np.float32)
    rp = np.zeros(41, dtype=np.int64)
    rp[1:] = np.cumsum(sizes)
    for red in ("sum", "mean", "max"):
        ref = L.pooling(torch.from_numpy(x).double(), torch.from_numpy(batch), size=40, reduce=red).numpy()
        out = K.spmm_csr(dev(rp), None, dev(x), reduce=red).cpu().numpy()
        assert rel_err(out, ref) <= FP32_TOL


def test_spmm_empty_and_ragged(K):
    x = torch.randn(10, 8, device="cuda")
    rp = torch.zeros(1, dtype=torch.int64, device="cuda")
    out = K.spmm_csr(rp, torch.zeros(0, dtype=torch.int32, device="cuda"), x)
    assert out.shape == (0, 8)
    rp = torch.zeros(6, dtype=torch.int64, device="cuda")  # 5 rows, no edges
    out = K.spmm_csr(rp, torch.zeros(0, dtype=torch.int32, device="cuda"), x, reduce="mean")
    assert torch.count_nonzero(out).item() == 0
    out = K.spmm_csr(rp, torch.zeros(0, dtype=torch.int32, device="cuda"), x, reduce="max")
    assert torch.count_nonzero(out).item() == 0


def test_spmm_linearity_large(K):
    """Size-independent property at a larger size: A(x+y) = Ax + Ay and mean of ones = 1 on non-empty rows."""
    g = torch.Generator(device="cuda").manual_seed(0)
    n, F, deg = 50000, 256, 32
    col = torch.randint(0, n, (n * deg,), device="cuda", generator=g, dtype=torch.int32)
    rp = torch.arange(0, n * deg + 1, deg, device="cuda", dtype=torch.int64)
    x = torch.randn(n, F, device="cuda", generator=g)
    y = torch.randn(n, F, device="cuda", generator=g)
    a = K.spmm_csr(rp, col, x + y)
    b = K.spmm_csr(rp, col, x) + K.spmm_csr(rp, col, y)
    assert (a - b).abs().max().item() <= 1e-4 * a.abs().max().item()
    ones = K.spmm_csr(rp, col, torch.ones(n, F, device="cuda"), reduce="mean")
    assert torch.equal(ones, torch.ones_like(ones))


# ---------------------------------------------------------------- SDDMM -----
@pytest.mark.parametrize("F", [64, 100, 7, 256])
def test_sddmm_parity(K, F):
    rng = np.random.default_rng(F)
    n = 400
    rp, col = rand_csr(rng, n, n, 20)
    a = rng.standard_normal((n, F)).astype(np.float32)
    b = rng.standard_normal((n, F)).astype(np.float32)
    ref = oracle.sddmm_csr(rp, col, a, b)
    out = K.sddmm_csr(dev(rp), dev(col), dev(a), dev(b)).cpu().numpy()
    assert rel_err(out, ref) <= FP32_TOL


# --------------------------------------------------------------- gather -----
@pytest.mark.parametrize("F,ld,dtype", [(602, 604, np.float32), (602, 602, np.float32), (128, 128, np.float32),
                                        (1, 1, np.float32), (2000, 2000, np.float32), (19, 20, np.int32),
                                        (3, 3, np.uint8), (304, 304, np.float16)])
@pytest.mark.parametrize("ids64", [True, False])
def test_gather_rows_bit_exact(K, F, ld, dtype, ids64):
    rng = np.random.default_rng(F)
    n_src, m = 3000, 5001
    if np.issubdtype(dtype, np.floating):
        x = rng.standard_normal((n_src, ld)).astype(dtype)
    else:
        x = rng.integers(0, 200, size=(n_src, ld)).astype(dtype)
    ids = rng.integers(0, n_src, size=m).astype(np.int64 if ids64 else np.int32)
    table = dev(x)
    view = table[:, :F]
    ref, _ = oracle.gather_rows(np.ascontiguousarray(x[:, :F]), ids)
    out = K.gather_rows(view, dev(ids))
    assert out.dtype == table.dtype and out.shape == (m, F)
    assert np.array_equal(out.cpu().numpy(), ref)


def test_gather_rows_empty_and_1d(K):
    table = torch.arange(100, device="cuda", dtype=torch.int64)
    ids = torch.tensor([5, 0, 99, 5], device="cuda")
    assert K.gather_rows(table, ids).tolist() == [5, 0, 99, 5]
    out = K.gather_rows(torch.randn(10, 16, device="cuda"), torch.zeros(0, dtype=torch.int64, device="cuda"))
    assert out.shape == (0, 16)


def test_gather_rows_cached_split(K):
    """GraphCacheServer.fetch_data semantics (FeatureCache/storage.py:151-198): hits from the HBM cache, misses
    from the pinned host table, miss counter — against the oracle's restatement."""
    rng = np.random.default_rng(3)
    n, F, m = 2000, 604, 3333
    host = rng.standard_normal((n, F)).astype(np.float32)
    deg = rng.integers(0, 1000, size=n)
    cached = np.argsort(-deg, kind="stable")[:700]
    flag = np.zeros(n, dtype=np.uint8)
    flag[cached] = 1
    l2c = np.zeros(n, dtype=np.int64)
    l2c[cached] = np.arange(cached.size)
    cache = host[cached].copy()
    ids = rng.integers(0, n, size=m).astype(np.int64)
    ref, ref_miss = oracle.gather_rows(cache, ids, host_table=host, gpu_flag=flag, local2cache=l2c)
    assert np.array_equal(ref, host[ids])
    host_t = torch.from_numpy(host).pin_memory()
    counter = torch.zeros(1, dtype=torch.int64, device="cuda")
    out = K.gather_rows_cached(dev(cache), host_t, dev(ids), dev(flag), dev(l2c), miss_counter=counter)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), ref)
    assert int(counter.item()) == ref_miss == int((flag[ids] == 0).sum())


# ------------------------------------------------------------ binarized -----
@pytest.mark.parametrize("F", [602, 32, 33, 128, 1, 1000])
def test_binarize_and_bin_spmm_bit_exact(K, F):
    rng = np.random.default_rng(F)
    n_dst, n_src = 500, 1200
    rp, col = rand_csr(rng, n_dst, n_src, 70, heavy=[(1, 1500)])
    x = rng.standard_normal((n_src, F)).astype(np.float32)
    x[rng.random(x.shape) < 0.05] = 0.0   # exact zeros count as +1 (x >= 0)
    x[0, 0] = -0.0                        # -0.0 >= 0 is true
    packed = K.binarize_pack(dev(x))
    ref_packed = oracle.binarize_pack(x)
    assert np.array_equal(packed.cpu().numpy().view(np.uint32), ref_packed)
    # 16-byte rows (the layout of the feature tables): the vectorised streaming pack, incl. a ragged row tail
    packed_v = K.binarize_pack(padded(x, (F + 3) // 4 * 4 + 4)[:, :F])
    assert np.array_equal(packed_v.cpu().numpy().view(np.uint32), ref_packed)
    ref_cnt = oracle.bin_spmm_counts(rp, col, ref_packed, F)
    cnt = K.bin_spmm_csr(dev(rp), dev(col), packed, F, mode="count").cpu().numpy()
    assert cnt.dtype == np.int32 and np.array_equal(cnt, ref_cnt)
    deg = (rp[1:] - rp[:-1]).astype(np.float32)[:, None]
    s = K.bin_spmm_csr(dev(rp), dev(col), packed, F, mode="sum").cpu().numpy()
    assert np.array_equal(s, 2.0 * ref_cnt.astype(np.float32) - deg)
    m = K.bin_spmm_csr(dev(rp), dev(col), packed, F, mode="mean").cpu().numpy()
    with np.errstate(divide="ignore", invalid="ignore"):
        ref_m = np.where(deg > 0, (2.0 * ref_cnt - deg) / deg, 0.0)
    assert rel_err(m, ref_m) <= FP32_TOL
    # nnz-split plan for long rows: integer atomics, still bit-exact
    plan = K.CsrPlan(dev(rp), chunk_edges=200)
    assert plan.n_heavy_rows >= 1
    assert np.array_equal(K.bin_spmm_csr(dev(rp), dev(col), packed, F, mode="count", plan=plan).cpu().numpy(), ref_cnt)
    assert np.array_equal(K.bin_spmm_csr(dev(rp), dev(col), packed, F, mode="sum", plan=plan).cpu().numpy(), s)
    assert rel_err(K.bin_spmm_csr(dev(rp), dev(col), packed, F, mode="mean", plan=plan).cpu().numpy(), ref_m) <= FP32_TOL
    # cross-check with the fp32 kernel on sign(x): same aggregation, different formulation
    sx = np.where(x >= 0, 1.0, -1.0).astype(np.float32)
    agg = K.spmm_csr(dev(rp), dev(col), dev(sx), reduce="sum").cpu().numpy()
    assert np.array_equal(agg, s)


# ------------------------------------------------------------------ GAT -----
@pytest.mark.parametrize("heads,D", [(4, 64), (1, 47), (4, 16), (2, 8), (1, 256), (3, 5), (4, 128), (2, 48), (3, 100)])
@pytest.mark.parametrize("mode", ["softmax", "exp_neg"])
def test_gat_forward_parity(K, heads, D, mode):
    rng = np.random.default_rng(heads * 100 + D)
    n = 600
    rp, col = rand_csr(rng, n, n, 40, heavy=[(2, 700)])
    wh = rng.standard_normal((n, heads * D)).astype(np.float32)
    el = rng.standard_normal((n, heads)).astype(np.float32)
    er = rng.standard_normal((n, heads)).astype(np.float32)
    for elu in (False, True):
        ref = oracle.gat_forward(rp, col, wh, el, er, heads, 0.2, mode=mode, elu=elu)
        out = K.gat_forward(dev(rp), dev(col), dev(wh), dev(el), dev(er), heads, 0.2, mode=mode, elu=elu)
        assert rel_err(out.cpu().numpy(), ref) <= FP32_TOL


@pytest.mark.parametrize("heads,D", [(4, 64), (1, 256), (4, 128), (2, 48)])
def test_gat_whole_row_kernel_matches_oracle(K, heads, D, gat_kernel):
    """Option gat_kernel=row pins the one-warp-per-row (all heads) forward kernel."""
    gat_kernel("row")
    rng = np.random.default_rng(heads * 7 + D)
    n = 500
    rp, col = rand_csr(rng, n, n, 45, heavy=[(7, 600)])
    wh = rng.standard_normal((n, heads * D)).astype(np.float32)
    el = rng.standard_normal((n, heads)).astype(np.float32)
    er = rng.standard_normal((n, heads)).astype(np.float32)
    for mode in ("softmax", "exp_neg"):
        ref = oracle.gat_forward(rp, col, wh, el, er, heads, 0.2, mode=mode, elu=True)
        out, rmax, rsum = K.gat_forward(dev(rp), dev(col), dev(wh), dev(el), dev(er), heads, 0.2, mode=mode, elu=True,
                                        save_stats=True)
        assert rel_err(out.cpu().numpy(), ref) <= FP32_TOL
        gat_kernel("group")
        out2, rmax2, rsum2 = K.gat_forward(dev(rp), dev(col), dev(wh), dev(el), dev(er), heads, 0.2, mode=mode,
                                           elu=True, save_stats=True)
        gat_kernel("row")
        assert torch.equal(rmax, rmax2) and rel_err(rsum.cpu().numpy(), rsum2.cpu().numpy()) <= FP32_TOL
        # long rows split into chunks, partial softmax states merged
        plan = K.CsrPlan(dev(rp), chunk_edges=100)
        assert plan.n_heavy_rows >= 1
        out3, rmax3, rsum3 = K.gat_forward(dev(rp), dev(col), dev(wh), dev(el), dev(er), heads, 0.2, mode=mode,
                                           elu=True, save_stats=True, plan=plan)
        assert rel_err(out3.cpu().numpy(), ref) <= FP32_TOL
        assert torch.equal(rmax3, rmax2) and rel_err(rsum3.cpu().numpy(), rsum2.cpu().numpy()) <= FP32_TOL


@pytest.mark.parametrize("heads,D", [(4, 64), (1, 47), (2, 8)])
@pytest.mark.parametrize("mode", ["softmax", "exp_neg"])
def test_gat_backward_matches_autograd(K, heads, D, mode):
    rng = np.random.default_rng(heads + D)
    n = 300
    rp, col = rand_csr(rng, n, n, 20)
    wh = rng.standard_normal((n, heads * D)).astype(np.float32)
    el = rng.standard_normal((n, heads)).astype(np.float32)
    er = rng.standard_normal((n, heads)).astype(np.float32)
    g = rng.standard_normal((n, heads * D)).astype(np.float32)
    # fp64 autograd reference of the same definition (oracle/layers.py restates gatconv.py; here per-edge form)
    rows = torch.from_numpy(np.repeat(np.arange(n), rp[1:] - rp[:-1])).long()
    cols = torch.from_numpy(col).long()
    twh = torch.from_numpy(wh).double().requires_grad_(True)
    tel = torch.from_numpy(el).double().requires_grad_(True)
    ter = torch.from_numpy(er).double().requires_grad_(True)
    z = torch.nn.functional.leaky_relu(tel[rows] + ter[cols], 0.2)
    s = z if mode == "softmax" else -z
    smax = torch.full((n, heads), -float("inf"), dtype=torch.float64).scatter_reduce(
        0, rows[:, None].expand(-1, heads), s.detach(), reduce="amax")
    ex = torch.exp(s - smax[rows])
    den = torch.zeros(n, heads, dtype=torch.float64).index_add_(0, rows, ex)
    alpha = ex / den[rows]
    msg = alpha[:, :, None] * twh[cols].view(-1, heads, D)
    out_ref = torch.zeros(n, heads, D, dtype=torch.float64).index_add_(0, rows, msg).view(n, heads * D)
    out_ref.backward(torch.from_numpy(g).double())

    drp, dcol = dev(rp), dev(col)
    out, rmax, rsum = K.gat_forward(drp, dcol, dev(wh), dev(el), dev(er), heads, 0.2, mode=mode, save_stats=True)
    assert rel_err(out.cpu().numpy(), out_ref.detach().numpy()) <= FP32_TOL
    trp, tcol, _, perm = K.csr_transpose(drp, dcol, n, want_perm=True)
    d_wh, d_el, d_er = K.gat_backward(drp, dcol, trp, tcol, perm, dev(wh), dev(el), dev(er), out, rmax, rsum,
                                      dev(g), heads, 0.2, mode=mode)
    assert rel_err(d_wh.cpu().numpy(), twh.grad.numpy()) <= 2e-5
    assert rel_err(d_el.cpu().numpy(), tel.grad.numpy()) <= 2e-5
    assert rel_err(d_er.cpu().numpy(), ter.grad.numpy()) <= 2e-5


@pytest.mark.parametrize("heads,D", [(4, 64), (2, 20), (1, 8)])
@pytest.mark.parametrize("mode", ["softmax", "exp_neg"])
def test_gat_backward_with_row_splitting_plans(K, heads, D, mode):
    """Long rows of the forward CSR (pass 1) and of the transposed CSR (pass 2) are split into (row, chunk) items whose
    partials are merged in a fixed order: same gradients as the unsplit kernels (<= 2e-5), run-to-run bit-identical."""
    rng = np.random.default_rng(11 * heads + D)
    n = 400
    deg = rng.integers(0, 12, n)
    deg[[3, 77, 399]] = [300, 97, 64]                       # long destination rows
    rp = np.zeros(n + 1, dtype=np.int64)
    rp[1:] = np.cumsum(deg)
    col = rng.integers(0, n, rp[-1]).astype(np.int32)
    col[rng.random(col.size) < 0.3] = 5                     # one hub source: a long row of the transposed CSR
    col[rng.random(col.size) < 0.05] = 390
    wh = rng.standard_normal((n, heads * D)).astype(np.float32)
    el = rng.standard_normal((n, heads)).astype(np.float32)
    er = rng.standard_normal((n, heads)).astype(np.float32)
    g = rng.standard_normal((n, heads * D)).astype(np.float32)
    drp, dcol = dev(rp), dev(col)
    out, rmax, rsum = K.gat_forward(drp, dcol, dev(wh), dev(el), dev(er), heads, 0.2, mode=mode, save_stats=True)
    trp, tcol, _, perm = K.csr_transpose(drp, dcol, n, want_perm=True)
    base = K.gat_backward(drp, dcol, trp, tcol, perm, dev(wh), dev(el), dev(er), out, rmax, rsum, dev(g), heads, 0.2,
                          mode=mode)
    plan, t_plan = K.CsrPlan(drp, chunk_edges=32), K.CsrPlan(trp, chunk_edges=32)
    assert plan.n_heavy_rows == 3 and t_plan.n_heavy_rows >= 2
    for pl, tpl in ((plan, t_plan), (plan, None), (None, t_plan)):
        got = K.gat_backward(drp, dcol, trp, tcol, perm, dev(wh), dev(el), dev(er), out, rmax, rsum, dev(g), heads,
                             0.2, mode=mode, plan=pl, t_plan=tpl)
        again = K.gat_backward(drp, dcol, trp, tcol, perm, dev(wh), dev(el), dev(er), out, rmax, rsum, dev(g), heads,
                               0.2, mode=mode, plan=pl, t_plan=tpl)
        for a, b, c in zip(got, base, again):
            assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 2e-5
            assert torch.equal(a, c)
    # with attention dropout as well (the mask is a function of the edge position, not of the schedule)
    o2, m2, s2 = K.gat_forward(drp, dcol, dev(wh), dev(el), dev(er), heads, 0.2, mode=mode, save_stats=True,
                               dropout=0.4, seed=5)
    b2 = K.gat_backward(drp, dcol, trp, tcol, perm, dev(wh), dev(el), dev(er), o2, m2, s2, dev(g), heads, 0.2,
                        mode=mode, dropout=0.4, seed=5)
    g2 = K.gat_backward(drp, dcol, trp, tcol, perm, dev(wh), dev(el), dev(er), o2, m2, s2, dev(g), heads, 0.2,
                        mode=mode, dropout=0.4, seed=5, plan=plan, t_plan=t_plan)
    for a, b in zip(g2, b2):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 2e-5


@pytest.mark.parametrize("heads,D", [(4, 64), (3, 16), (1, 32), (2, 128), (4, 128), (2, 20)])
@pytest.mark.parametrize("mode", ["softmax", "exp_neg"])
def test_gat_backward_fused_single_pass(K, heads, D, mode):
    """The single-pass backward over the transposed CSR (whole-row shapes; option gat_bwd_kernel) on a NON-square graph
    with long rows on both sides, with and without plans and attention dropout: fp64 autograd of the same definition
    (<= 2e-5), the two-pass kernels (<= 2e-5), run-to-run bit-identical."""
    rng = np.random.default_rng(7 * heads + D)
    n_dst, n_src = 250, 400
    rp, col = rand_csr(rng, n_dst, n_src, 24, heavy=[(3, 300), (77, 97)])
    col[rng.random(col.size) < 0.25] = 5                    # a hub source: long row of the transposed CSR
    wh = rng.standard_normal((n_src, heads * D)).astype(np.float32)
    el = rng.standard_normal((n_dst, heads)).astype(np.float32)
    er = rng.standard_normal((n_src, heads)).astype(np.float32)
    g = rng.standard_normal((n_dst, heads * D)).astype(np.float32)
    drp, dcol = dev(rp), dev(col)
    trp, tcol, _, perm = K.csr_transpose(drp, dcol, n_src, want_perm=True)
    plan, t_plan = K.CsrPlan(drp, chunk_edges=32), K.CsrPlan(trp, chunk_edges=32)
    assert plan.n_heavy_rows == 2 and t_plan.n_heavy_rows >= 1
    rows = torch.from_numpy(np.repeat(np.arange(n_dst), rp[1:] - rp[:-1])).long()
    cols = torch.from_numpy(col).long()
    for drop in (0.0, 0.4):
        out, rmax, rsum = K.gat_forward(drp, dcol, dev(wh), dev(el), dev(er), heads, 0.2, mode=mode, save_stats=True,
                                        dropout=drop, seed=9)
        mask = K.gat_dropout_mask(9, col.size, heads, drop).cpu().double() if drop else torch.ones(col.size, heads).double()
        twh = torch.from_numpy(wh).double().requires_grad_(True)
        tel = torch.from_numpy(el).double().requires_grad_(True)
        ter = torch.from_numpy(er).double().requires_grad_(True)
        z = torch.nn.functional.leaky_relu(tel[rows] + ter[cols], 0.2)
        sc = z if mode == "softmax" else -z
        smax = torch.full((n_dst, heads), -float("inf"), dtype=torch.float64).scatter_reduce(
            0, rows[:, None].expand(-1, heads), sc.detach(), reduce="amax")
        ex = torch.exp(sc - smax[rows])
        den = torch.zeros(n_dst, heads, dtype=torch.float64).index_add_(0, rows, ex)
        alpha = ex / den[rows] * mask
        msg = alpha[:, :, None] * twh[cols].view(-1, heads, D)
        out_ref = torch.zeros(n_dst, heads, D, dtype=torch.float64).index_add_(0, rows, msg).view(n_dst, heads * D)
        assert rel_err(out.cpu().numpy(), out_ref.detach().numpy()) <= FP32_TOL
        out_ref.backward(torch.from_numpy(g).double())
        ref = (twh.grad.numpy(), tel.grad.numpy(), ter.grad.numpy())
        args = (drp, dcol, trp, tcol, perm, dev(wh), dev(el), dev(er), out, rmax, rsum, dev(g), heads, 0.2)
        try:
            K.set_option("gat_bwd_kernel", "twopass")
            two = K.gat_backward(*args, mode=mode, dropout=drop, seed=9)
            K.set_option("gat_bwd_kernel", "fused")
            for pl, tpl in ((None, None), (plan, t_plan), (plan, None), (None, t_plan)):
                got = K.gat_backward(*args, mode=mode, dropout=drop, seed=9, plan=pl, t_plan=tpl)
                again = K.gat_backward(*args, mode=mode, dropout=drop, seed=9, plan=pl, t_plan=tpl)
                for a, b, c, r in zip(got, two, again, ref):
                    assert rel_err(a.cpu().numpy(), r) <= 2e-5
                    assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 2e-5
                    assert torch.equal(a, c)
        finally:
            K.set_option("gat_bwd_kernel", None)


# ------------------------------------------------------------ transpose -----
def test_csr_transpose_bit_exact(K):
    rng = np.random.default_rng(4)
    n_rows, n_cols = 700, 500
    rp, col = rand_csr(rng, n_rows, n_cols, 30)
    vals = rng.standard_normal(col.size).astype(np.float32)
    for rpt in (np.int64, np.int32):
        t_rp, t_col, t_val, perm = K.csr_transpose(dev(rp.astype(rpt)), dev(col), n_cols, values=dev(vals),
                                                   want_perm=True)
        rows = np.repeat(np.arange(n_rows), rp[1:] - rp[:-1])
        order = np.argsort(col, kind="stable")
        ref_rp = np.zeros(n_cols + 1, dtype=np.int64)
        np.add.at(ref_rp, col.astype(np.int64) + 1, 1)
        ref_rp = np.cumsum(ref_rp)
        assert np.array_equal(t_rp.cpu().numpy(), ref_rp)
        assert np.array_equal(t_col.cpu().numpy(), rows[order])
        assert np.array_equal(perm.cpu().numpy(), order)
        assert np.array_equal(t_val.cpu().numpy(), vals[order])


# ----------------------------------------------------------------- GEMM -----
@pytest.mark.parametrize("M,N,K_", [(300, 64, 50), (1000, 256, 602), (17, 5, 3), (128, 121, 64)])
def test_gemm_fp32_parity(K, M, N, K_):
    rng = np.random.default_rng(M)
    a = rng.standard_normal((M, K_)).astype(np.float32)
    b = rng.standard_normal((K_, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    ref = oracle.gemm(a, b, bias=bias, relu=True)
    out = K.gemm(dev(a), dev(b), bias=dev(bias), relu=True).cpu().numpy()
    assert rel_err(out, ref) <= FP32_TOL
    # transposed operands + accumulate (dW = H^T G)
    g = rng.standard_normal((M, N)).astype(np.float32)
    acc = rng.standard_normal((K_, N)).astype(np.float32)
    ref2 = acc.astype(np.float64) + a.astype(np.float64).T @ g.astype(np.float64)
    o = dev(acc)
    K.gemm(dev(a), dev(g), trans_a=True, out=o, accumulate=True)
    assert rel_err(o.cpu().numpy(), ref2) <= FP32_TOL
    ref3 = g.astype(np.float64) @ b.astype(np.float64).T
    assert rel_err(K.gemm(dev(g), dev(b), trans_b=True).cpu().numpy(), ref3) <= FP32_TOL


# -------------------------------------------------------------- sampler -----
@pytest.mark.parametrize("fanout", [10, 25, 1, 32, -1])
def test_sample_neighbors_structure(K, fanout):
    rng = np.random.default_rng(8)
    n = 2000
    rp, col = rand_csr(rng, n, n, 80, empty_frac=0.05)
    # distinct neighbours per row so "no duplicates" is checkable
    for i in range(n):
        d = rp[i + 1] - rp[i]
        col[rp[i]:rp[i + 1]] = rng.choice(n, size=d, replace=False)
    seeds = rng.integers(0, n, size=1024).astype(np.int64)
    o_rp, o_col = K.sample_neighbors(dev(rp), dev(col), dev(seeds), fanout, rng_seed=123)
    o_rp, o_col = o_rp.cpu().numpy(), o_col.cpu().numpy()
    deg = (rp[1:] - rp[:-1])[seeds]
    want = deg if fanout < 0 else np.minimum(deg, fanout)
    assert np.array_equal(np.diff(o_rp), want) and o_rp[0] == 0
    for i, v in enumerate(seeds):
        got = o_col[o_rp[i]:o_rp[i + 1]]
        nb = col[rp[v]:rp[v + 1]]
        if fanout < 0 or deg[i] <= fanout:
            assert np.array_equal(got, nb)  # all neighbours, original order (base_sampler.py:53-54)
        else:
            assert len(set(got.tolist())) == len(got)          # without replacement
            assert set(got.tolist()) <= set(nb.tolist())       # subset of the neighbourhood
            pos = [int(np.where(nb == g_)[0][0]) for g_ in got]
            assert pos == sorted(pos)                          # neighbour order kept
    # determinism for a given rng_seed, different draw for another
    o2 = K.sample_neighbors(dev(rp), dev(col), dev(seeds), fanout, rng_seed=123)[1].cpu().numpy()
    assert np.array_equal(o2, o_col)
    if fanout in (10, 25):
        o3 = K.sample_neighbors(dev(rp), dev(col), dev(seeds), fanout, rng_seed=124)[1].cpu().numpy()
        assert not np.array_equal(o3, o_col)


def test_sample_neighbors_uniformity(K):
    """Every neighbour of a degree-50 node is chosen with probability fanout/deg (chi-square-ish bound)."""
    deg, fanout, trials = 50, 10, 20000
    rp = np.array([0, deg], dtype=np.int64)
    col = np.arange(deg, dtype=np.int32)
    seeds = np.zeros(trials, dtype=np.int64)
    _, o_col = K.sample_neighbors(dev(rp), dev(col), dev(seeds), fanout, rng_seed=7)
    hist = np.bincount(o_col.cpu().numpy(), minlength=deg)
    expect = trials * fanout / deg
    assert np.all(np.abs(hist - expect) < 6 * np.sqrt(expect))


def test_host_sampled_blocks_aggregate_exactly(K):
    """Parity protocol of SURVEY.md App. B: the reference's host sampler (restated) produces the index lists; the
    device aggregates them; result equals the oracle on the same lists."""
    import random
    rng = np.random.default_rng(21)
    n, F = 500, 602
    rp, col = rand_csr(rng, n, n, 30)
    x = rng.standard_normal((n, F)).astype(np.float32)
    seeds = list(range(0, 64))
    random.seed(0)
    src, dst = S.sample_neighbours(rp, col, seeds, 10, rng=random)
    b_rp = S.block_to_csr(dst, seeds)
    b_rp = np.asarray(b_rp, dtype=np.int64)
    b_col = np.asarray(src, dtype=np.int32)
    ref = oracle.spmm_csr(b_rp, b_col, x, reduce="mean")
    out = K.spmm_csr(dev(b_rp), dev(b_col), padded(x, 604), reduce="mean", F=F).cpu().numpy()
    assert rel_err(out, ref) <= FP32_TOL


# ---------------------------------------------------------- legacy fused ----
def _fused_inputs(rng, N, F, Hd):
    rp, col = rand_csr(rng, N, N, 25, empty_frac=0.05)
    vals = (rng.random(col.size).astype(np.float32) + 0.05) / 10
    Fp = (F + 3) // 4 * 4
    X = np.zeros((N, Fp), dtype=np.float32)
    X[:, :F] = rng.standard_normal((N, F)).astype(np.float32)
    W = (rng.standard_normal((Fp, Hd)) / np.sqrt(F)).astype(np.float32)
    nn = (rp[1:] - rp[:-1]).astype(np.int32)
    return rp.astype(np.int32), col, vals, X, W, nn, Fp


@pytest.mark.parametrize("N,F,Hd", [(591, 50, 64), (1021, 64, 121), (300, 602, 41)])
def test_gcn_fused_forward_backward(K, N, F, Hd):
    rng = np.random.default_rng(N)
    rp, col, vals, X, W, nn, Fp = _fused_inputs(rng, N, F, Hd)
    ref = oracle.gcn_fused_forward(rp, col, vals, X, W, nn, F)
    H = K.gcn_fused_forward_v2(dev(rp), dev(col), dev(vals), dev(X), dev(W), dev(nn), F)
    assert rel_err(H.cpu().numpy(), ref) <= FP32_TOL
    # true gradients: torch autograd over relu(A (X W)) in fp64
    rows = torch.from_numpy(np.repeat(np.arange(N), np.diff(rp))).long()
    A = torch.sparse_coo_tensor(torch.stack([rows, torch.from_numpy(col).long()]),
                                torch.from_numpy(vals).double(), (N, N))
    tX = torch.from_numpy(X).double().requires_grad_(True)
    tW = torch.from_numpy(W).double().requires_grad_(True)
    Wm = tW.clone()
    out = torch.relu(torch.sparse.mm(A, tX[:, :F] @ Wm[:F]))
    g = rng.standard_normal((N, Hd)).astype(np.float32)
    out.backward(torch.from_numpy(g).double())
    gX, gW = K.gcn_fused_backward_v2(dev(g), dev(rp), dev(col), dev(vals), dev(X), dev(W), H, dev(nn), F)
    assert rel_err(gX.cpu().numpy()[:, :F], tX.grad.numpy()[:, :F]) <= 2e-5
    assert rel_err(gW.cpu().numpy()[:F], tW.grad.numpy()[:F]) <= 2e-5


def test_legacy_symbols_blocking_abi(K):
    """launch_gcn_fused_kernel keeps the reference's C ABI (gcn_fused_kernel.cu:190-195): void, blocking."""
    import ctypes
    from dgll_b200 import _lib
    rng = np.random.default_rng(1)
    N, F, Hd = 400, 50, 64
    rp, col, vals, X, W, nn, Fp = _fused_inputs(rng, N, F, Hd)
    t = [dev(a) for a in (rp, col, vals, X, W, nn)]
    H = torch.zeros((N, Hd), device="cuda")
    torch.cuda.synchronize()
    P = lambda z: ctypes.c_void_p(z.data_ptr())
    _lib.lib().launch_gcn_fused_kernel(P(t[0]), P(t[1]), P(t[2]), P(t[3]), P(t[4]), P(H), P(t[5]), N, Fp, F, Hd,
                                       col.size)
    ref = oracle.gcn_fused_forward(rp, col, vals, X, W, nn, F)
    assert rel_err(H.cpu().numpy(), ref) <= FP32_TOL


def test_reference_kernel_agrees_with_oracle_when_built(K):
    """oracle/_ref holds the reference's own forward kernel compiled for sm_100a (when it was built in the
    container that has /root/reference): the restated oracle must match it on PPI-like shapes."""
    import ctypes
    path = oracle.ref_kernel_path()
    if path is None:
        pytest.skip("oracle/_ref/libgcn_fused_ref.so not built")
    ref_lib = ctypes.CDLL(path)
    rng = np.random.default_rng(6)
    N, F, Hd = 591, 50, 64
    rp, col, vals, X, W, nn, Fp = _fused_inputs(rng, N, F, Hd)
    t = [dev(a) for a in (rp, col, vals, X, W, nn)]
    H = torch.zeros((N, Hd), device="cuda")
    torch.cuda.synchronize()
    P = lambda z: ctypes.c_void_p(z.data_ptr())
    ref_lib.launch_gcn_fused_kernel.restype = None
    ref_lib.launch_gcn_fused_kernel(P(t[0]), P(t[1]), P(t[2]), P(t[3]), P(t[4]), P(H), P(t[5]), ctypes.c_int(N),
                                    ctypes.c_int(Fp), ctypes.c_int(F), ctypes.c_int(Hd), ctypes.c_int(col.size))
    torch.cuda.synchronize()
    orc = oracle.gcn_fused_forward(rp, col, vals, X, W, nn, F)
    assert rel_err(H.cpu().numpy(), orc) <= FP32_TOL
    ours = K.gcn_fused_forward_v2(*t[:5], t[5], F)
    assert rel_err(ours.cpu().numpy(), H.cpu().numpy()) <= FP32_TOL


def test_ppi_golden_layer_on_device(K):
    """C1: the reference's PPI GCNLayer output (golden, generated from Evaluation/PPI/gcn_model.py) reproduced by
    GEMM + SpMM(sum, relu) on the device."""
    from conftest import golden
    g = golden("ppi_gcn_g8")
    ei = g["edge_index"]
    n = g["feats"].shape[0]
    rp, col, _ = oracle.coo_to_csr(ei[0], ei[1], n)
    support = K.gemm(dev(g["feats"].astype(np.float32)), dev(g["w0"].astype(np.float32)))
    h1 = K.spmm_csr(dev(rp), dev(col), support, reduce="sum", relu=True)
    assert rel_err(h1.cpu().numpy(), g["h1"]) <= FP32_TOL


# ------------------------------------------------ SpMM kernel families -----
@pytest.fixture
def gat_kernel(K):
    yield lambda name: K.set_option("gat_kernel", name)
    K.set_option("gat_kernel", None)


@pytest.fixture
def spmm_family(K):
    """Pin one SpMM kernel family through the library's option table for the duration of a test."""
    def pin(name, **opts):
        K.set_option("spmm_kernel", name)
        for k, v in opts.items():
            K.set_option(k, v)
    yield pin
    for k in ("spmm_kernel", "rows_tb", "rows_ns", "rows_d", "rows_stream", "spmm_tb"):
        K.set_option(k, None)


@pytest.mark.parametrize("family", ["rowsplit", "stream", "wholerow", "wholerow:ns2", "wholerow:ns3:d8", "wholerow:st1",
                                    "wholerow:st2", "wholerow:st7", "wholerow:st32", "wholerow:ns1:st3"])
@pytest.mark.parametrize("F,ld,dt", [(602, 604, "f32"), (256, 256, "f32"), (128, 128, "f32"), (100, 100, "f32"),
                                     (602, 608, "bf16"), (1000, 1000, "f32"), (36, 36, "f32"), (1500, 1504, "bf16")])
def test_spmm_kernel_families_agree_with_oracle(K, family, F, ld, dt, spmm_family):
    """Option spmm_kernel pins the (row, slab)-per-warp, the rolling-LDG streaming or the whole-row rolling-window
    kernel (with its slabs-per-warp / window-depth variants): all must match the oracle on ragged blocks (empty rows,
    short rows, one long row, rows straddling chunk borders)."""
    parts = family.split(":")
    names = {"ns": "rows_ns", "d": "rows_d", "st": "rows_stream"}      # st1 = per-row kernel, st<n> = n rows per warp
    spmm_family(parts[0], **{names[o.rstrip("0123456789")]: int(o.lstrip("nsdt")) for o in parts[1:]})
    rng = np.random.default_rng(F + len(family))
    n_dst, n_src = 2500, 4000
    rp, col = rand_csr(rng, n_dst, n_src, 37, heavy=[(5, 1500), (2499, 300)], empty_frac=0.15)
    rp[-3:] = rp[-3]  # trailing empty rows
    col = col[:rp[-1]]
    x = rng.standard_normal((n_src, F)).astype(np.float32)
    vals = rng.random(col.size).astype(np.float32) + 0.1
    bias = rng.standard_normal(F).astype(np.float32)
    add = rng.standard_normal((n_dst, F)).astype(np.float32)
    rs = rng.random(n_dst).astype(np.float32)
    tx = padded(x, ld)
    tol = FP32_TOL
    if dt == "bf16":
        tx = tx.to(torch.bfloat16)
        x = tx[:, :F].float().cpu().numpy()
    for rpt in (np.int64, np.int32):
        for red in ("sum", "mean"):
            ref = oracle.spmm_csr(rp, col, x, reduce=red)
            out = K.spmm_csr(dev(rp.astype(rpt)), dev(col), tx, reduce=red, F=F)
            assert rel_err(out.cpu().numpy(), ref) <= tol, (family, red)
    ref = oracle.spmm_csr(rp, col, x, values=vals, reduce="mean", row_scale=rs, addend=add, bias=bias, relu=True)
    out = K.spmm_csr(dev(rp), dev(col), tx, values=dev(vals), reduce="mean", row_scale=dev(rs), addend=dev(add),
                     bias=dev(bias), relu=True, F=F)
    assert rel_err(out.cpu().numpy(), ref) <= tol
    # all rows empty / no edges at all
    rp0 = np.zeros(n_dst + 1, dtype=np.int64)
    out = K.spmm_csr(dev(rp0), torch.zeros(0, dtype=torch.int32, device="cuda"), tx, reduce="mean", bias=dev(bias), F=F)
    assert rel_err(out.cpu().numpy(), np.broadcast_to(bias, (n_dst, F))) <= tol


@pytest.mark.parametrize("family", ["rowsplit", "stream", "wholerow"])
def test_spmm_families_deterministic(K, family, spmm_family):
    spmm_family(family)
    g = torch.Generator(device="cuda").manual_seed(1)
    n, F = 20000, 602
    deg = torch.randint(0, 40, (n,), device="cuda", generator=g)
    rp = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    rp[1:] = torch.cumsum(deg, 0)
    col = torch.randint(0, n, (int(rp[-1].item()),), device="cuda", generator=g, dtype=torch.int32)
    x = torch.zeros(n, 604, device="cuda")
    x[:, :F] = torch.randn(n, F, device="cuda", generator=g)
    a = K.spmm_csr(rp, col, x, reduce="mean", F=F)
    b = K.spmm_csr(rp, col, x, reduce="mean", F=F)
    assert torch.equal(a, b)


# ------------------------------------------------------- tcgen05 TF32 GEMM --
TF32_TOL = 2e-3


@pytest.mark.parametrize("M,N,K_", [(128, 128, 32), (300, 64, 50), (1000, 256, 602), (17, 5, 3), (4096, 41, 256),
                                    (129, 130, 65), (2000, 256, 1204), (11264, 172, 256), (257, 300, 4096)])
def test_gemm_tcgen05_tf32_parity_all_orientations(K, M, N, K_):
    """precision='tf32': TMA loads the fp32 operands as they lie in memory (K-major or MN-major), tcgen05.mma kind::tf32
    accumulates in TMEM.  All four storage orientations, epilogue, accumulate, ragged M/N/K, operands whose rows are not
    16-byte multiples (aligned-copy fallback) and the deterministic split-K shape.  Bar: 2e-3 of max|ref| against the
    fp64 product of the fp32 operands; the result does not depend on how the operands are stored."""
    rng = np.random.default_rng(M + 3 * N)
    a = rng.standard_normal((M, K_)).astype(np.float32)
    b = rng.standard_normal((K_, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    da, db = dev(a), dev(b)
    at, bt = da.t().contiguous(), db.t().contiguous()          # stored [K, M] / [N, K]
    outs = [K.gemm(da, db, precision="tf32"), K.gemm(at, db, trans_a=True, precision="tf32"),
            K.gemm(da, bt, trans_b=True, precision="tf32"), K.gemm(at, bt, trans_a=True, trans_b=True, precision="tf32")]
    for o in outs:
        assert rel_err(o.cpu().numpy(), ref) <= TF32_TOL
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    # row strides that are 16-byte multiples (the layout of the feature tables): no copy at all
    ld_a, ld_b = (K_ + 3) // 4 * 4 + 4, (N + 3) // 4 * 4
    pa, pb = padded(a, ld_a)[:, :K_], padded(b, ld_b)[:, :N]
    assert torch.equal(K.gemm(pa, pb, precision="tf32"), outs[0])
    out = K.gemm(da, db, bias=dev(bias), relu=True, precision="tf32").cpu().numpy()
    assert rel_err(out, np.maximum(ref + bias, 0)) <= TF32_TOL
    acc = rng.standard_normal((M, N)).astype(np.float32)
    o = dev(acc)
    K.gemm(da, db, out=o, accumulate=True, precision="tf32")
    assert rel_err(o.cpu().numpy(), acc + ref) <= TF32_TOL
    # deterministic
    assert torch.equal(K.gemm(da, db, precision="tf32"), outs[0])


@pytest.mark.parametrize("M,N,K_", [(128, 128, 32), (300, 64, 50), (1000, 256, 602), (17, 5, 3), (129, 130, 65),
                                    (4096, 300, 256), (40000, 256, 100), (20000, 172, 256), (30000, 100, 256)])
def test_gemm_tf32_persistent_kernel(K, M, N, K_):
    """The persistent TF32 kernel (one CTA per SM walking tiles, BN = 256 when N > 128, two TMEM accumulators; option
    gemm_kernel=4 pins it, 3 pins the one-tile-per-CTA kernel): all four operand orientations, epilogue, accumulate,
    ragged edges, several tiles per CTA.  Same bar as the one-tile kernel, and bit-identical to it (same k order)."""
    rng = np.random.default_rng(M + 5 * N)
    a = rng.standard_normal((M, K_)).astype(np.float32)
    b = rng.standard_normal((K_, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    da, db = dev(a), dev(b)
    at, bt = da.t().contiguous(), db.t().contiguous()
    try:
        K.set_option("gemm_kernel", "3")
        one = K.gemm(da, db, precision="tf32")
        K.set_option("gemm_kernel", "4")
        outs = [K.gemm(da, db, precision="tf32"), K.gemm(at, db, trans_a=True, precision="tf32"),
                K.gemm(da, bt, trans_b=True, precision="tf32"),
                K.gemm(at, bt, trans_a=True, trans_b=True, precision="tf32")]
        for o in outs:
            assert rel_err(o.cpu().numpy(), ref) <= TF32_TOL
            assert torch.equal(o, outs[0])
        if K_ < 512:                                            # above that the one-tile kernel may split K
            assert torch.equal(outs[0], one)
        out = K.gemm(da, db, bias=dev(bias), relu=True, precision="tf32").cpu().numpy()
        assert rel_err(out, np.maximum(ref + bias, 0)) <= TF32_TOL
        acc = rng.standard_normal((M, N)).astype(np.float32)
        o = dev(acc)
        K.gemm(da, db, out=o, accumulate=True, precision="tf32")
        assert rel_err(o.cpu().numpy(), acc + ref) <= TF32_TOL
        ld_c = (N + 3) // 4 * 4 + 4                             # output rows with a 16-byte-multiple stride
        oc = torch.zeros((M, ld_c), device="cuda")
        K.gemm(da, db, out=oc[:, :N], precision="tf32")
        assert torch.equal(oc[:, :N], outs[0]) and float(oc[:, N:].abs().max()) == 0.0
    finally:
        K.set_option("gemm_kernel", None)
    # default dispatch (persistent when there is a tile per SM) gives the same result
    assert torch.equal(K.gemm(da, db, precision="tf32"), outs[0]) or K_ >= 512


@pytest.mark.parametrize("M,N,K_", [(128, 128, 32), (300, 64, 50), (1000, 256, 602), (17, 5, 3), (129, 130, 65),
                                    (4096, 300, 256), (40000, 256, 100), (602, 256, 11264), (257, 300, 4096),
                                    (300, 128, 200000)])
def test_gemm_tf32x3_fp32_grade(K, M, N, K_):
    """precision='tf32x3': operands split into tf32 hi + lo inside the kernel, three MMAs per k-step, fp32 accumulation in
    TMEM.  Bar: the fp32 bar of north_star (1e-5 of max|ref| against the fp64 product) — measured ~1e-6 — in all four
    operand orientations, with epilogue / accumulate / ragged edges, and through the kernel's own split-K on the
    long-reduction shapes; deterministic."""
    rng = np.random.default_rng(M + 7 * N)
    a = rng.standard_normal((M, K_)).astype(np.float32)
    b = rng.standard_normal((K_, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    da, db = dev(a), dev(b)
    at, bt = da.t().contiguous(), db.t().contiguous()
    outs = [K.gemm(da, db, precision="tf32x3"), K.gemm(at, db, trans_a=True, precision="tf32x3"),
            K.gemm(da, bt, trans_b=True, precision="tf32x3"),
            K.gemm(at, bt, trans_a=True, trans_b=True, precision="tf32x3")]
    for o in outs:
        assert rel_err(o.cpu().numpy(), ref) <= 0.5 * FP32_TOL
        assert torch.equal(o, outs[0])
    out = K.gemm(da, db, bias=dev(bias), relu=True, precision="tf32x3").cpu().numpy()
    assert rel_err(out, np.maximum(ref + bias, 0)) <= FP32_TOL
    acc = rng.standard_normal((M, N)).astype(np.float32)
    o = dev(acc)
    K.gemm(da, db, out=o, accumulate=True, precision="tf32x3")
    assert rel_err(o.cpu().numpy(), acc + ref) <= FP32_TOL
    assert torch.equal(K.gemm(da, db, precision="tf32x3"), outs[0])


def test_layers_run_on_the_tf32_path():
    import dgll_b200.nn as nn
    from dgll_b200 import ops
    torch.manual_seed(0)
    lin = torch.nn.Linear(300, 64).cuda()
    x = torch.randn(500, 300, device="cuda", requires_grad=True)
    ref = torch.nn.functional.linear(x.double(), lin.weight.double(), lin.bias.double())
    out = ops.linear(x, lin.weight, bias=lin.bias, trans_w=True, precision="tf32")
    assert rel_err(out.detach().cpu().numpy(), ref.detach().cpu().numpy()) <= TF32_TOL
    g = torch.randn_like(out)
    out.backward(g)
    gx, gw = torch.autograd.grad(ref, [x, lin.weight], g.double(), allow_unused=True)
    assert rel_err(x.grad.cpu().numpy(), gx.cpu().numpy()) <= TF32_TOL
    assert rel_err(lin.weight.grad.cpu().numpy(), gw.cpu().numpy()) <= TF32_TOL


# ------------------------------------------------------- tcgen05 GEMM ------
@pytest.mark.parametrize("M,N,K_", [(128, 128, 64), (300, 64, 50), (1000, 256, 602), (17, 5, 3), (4096, 41, 256),
                                    (129, 130, 65), (2000, 256, 1204)])
def test_gemm_tcgen05_bf16_parity(K, M, N, K_):
    """precision='bf16': tcgen05.mma with bf16 operands, fp32 accumulation in TMEM.  Bar: 1e-2 of max|ref| against the
    fp64 product of the ORIGINAL fp32 operands, and 1e-5 against the product of the bf16-rounded operands."""
    rng = np.random.default_rng(M + N)
    a = rng.standard_normal((M, K_)).astype(np.float32)
    b = rng.standard_normal((K_, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    out = K.gemm(dev(a), dev(b), precision="bf16").cpu().numpy()
    ref = a.astype(np.float64) @ b.astype(np.float64)
    assert rel_err(out, ref) <= BF16_TOL
    ar = dev(a).to(torch.bfloat16).float().cpu().numpy().astype(np.float64)
    br = dev(b).to(torch.bfloat16).float().cpu().numpy().astype(np.float64)
    assert rel_err(out, ar @ br) <= 2e-5
    # epilogue: bias + relu; transposed operands; accumulate
    out = K.gemm(dev(a), dev(b), bias=dev(bias), relu=True, precision="bf16").cpu().numpy()
    assert rel_err(out, np.maximum(ar @ br + bias, 0)) <= 2e-5
    g = rng.standard_normal((M, N)).astype(np.float32)
    gr = dev(g).to(torch.bfloat16).float().cpu().numpy().astype(np.float64)
    acc = rng.standard_normal((K_, N)).astype(np.float32)
    o = dev(acc)
    K.gemm(dev(a), dev(g), trans_a=True, out=o, accumulate=True, precision="bf16")      # dW += A^T G
    assert rel_err(o.cpu().numpy(), acc + ar.T @ gr) <= 2e-5
    o2 = K.gemm(dev(g), dev(b), trans_b=True, precision="bf16").cpu().numpy()            # dX = G W^T
    assert rel_err(o2, gr @ br.T) <= 2e-5


@pytest.mark.parametrize("M,N,K_", [(256, 602, 11264), (41, 256, 11264), (130, 70, 5000), (128, 128, 512)])
def test_gemm_tcgen05_split_k(K, M, N, K_):
    """Few output tiles + a long reduction (dW = X^T G of a training step): the k-blocks are split over CTAs, partials
    summed in a fixed order by a second kernel.  Same bar as the unsplit kernel, epilogue / accumulate honoured,
    run-to-run bit-identical (no atomics)."""
    rng = np.random.default_rng(M + K_)
    x = rng.standard_normal((K_, M)).astype(np.float32)          # stored [K, M]: trans_a
    g = rng.standard_normal((K_, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    xr = dev(x).to(torch.bfloat16).float().cpu().numpy().astype(np.float64)
    gr = dev(g).to(torch.bfloat16).float().cpu().numpy().astype(np.float64)
    ref = xr.T @ gr
    out = K.gemm(dev(x), dev(g), trans_a=True, precision="bf16")
    assert rel_err(out.cpu().numpy(), ref) <= 2e-5
    assert rel_err(out.cpu().numpy(), x.astype(np.float64).T @ g.astype(np.float64)) <= BF16_TOL
    again = K.gemm(dev(x), dev(g), trans_a=True, precision="bf16")
    assert torch.equal(out, again)
    acc = rng.standard_normal((M, N)).astype(np.float32)
    o = dev(acc)
    K.gemm(dev(x), dev(g), trans_a=True, out=o, accumulate=True, bias=dev(bias), relu=True, precision="bf16")
    assert rel_err(o.cpu().numpy(), np.maximum(acc + ref + bias, 0)) <= 2e-5
    # output with a leading dimension (column slice of a wider buffer)
    wide = torch.zeros((M, N + 9), device="cuda")
    K.gemm(dev(x), dev(g), trans_a=True, out=wide[:, :N], precision="bf16")
    assert rel_err(wide[:, :N].cpu().numpy(), ref) <= 2e-5 and (wide[:, N:] == 0).all()


def test_layers_run_on_tensor_core_path(K):
    """ops.set_gemm_precision('bf16') routes every dense transform of a layer through tcgen05; GCN output stays within
    the bf16 bar of the fp32 path."""
    import dgll_b200.nn as nn
    from dgll_b200 import ops
    from conftest import golden
    g = golden("nn_gcn")
    n = g["x"].shape[0]
    adj = torch.sparse_coo_tensor(dev(g["adj_indices"]).long(), dev(g["adj_values"]).float(), (n, n))
    model = nn.GCN(32, 16, 7, dropout=0.0).cuda().eval()
    x = dev(g["x"]).float()
    ref = model(x, adj)
    ops.set_gemm_precision("bf16")
    try:
        out = model(x, adj)
    finally:
        ops.set_gemm_precision("fp32")
    assert rel_err(out.detach().cpu().numpy(), ref.detach().cpu().numpy()) <= BF16_TOL


# ---------------------------------------------------- sharded / peer gather --
@pytest.mark.parametrize("F,dtype", [(128, torch.float32), (602, torch.float32), (64, torch.bfloat16), (3, torch.float32)])
def test_gather_rows_sharded_bit_exact(K, F, dtype):
    """Node-range sharded table read through a device array of shard pointers (here all shards live on one GPU; on
    the 8-GPU box the same kernel dereferences peer pointers mapped with dgllb_ipc_import)."""
    g = torch.Generator(device="cuda").manual_seed(F)
    n, P = 10007, 4
    part = (n + P - 1) // P
    ld = (F + 3) // 4 * 4 if dtype == torch.float32 else (F + 7) // 8 * 8
    full = torch.randn((n, ld), device="cuda", generator=g).to(dtype)
    shards = [full[r * part:min(n, (r + 1) * part)].clone() for r in range(P)]
    ptrs = torch.tensor([s.data_ptr() for s in shards], dtype=torch.int64, device="cuda")
    for idt in (torch.int64, torch.int32):
        ids = torch.randint(0, n, (5000,), device="cuda", generator=g).to(idt)
        out = K.gather_rows_sharded(ptrs, part, ld * full.element_size(), ids, ld, dtype)
        assert torch.equal(out, full[ids.long()])
    from dgll_b200.parallel import PeerShardedTable
    t = PeerShardedTable(n, full)            # world size 1: the single shard is the local table
    ids = torch.randint(0, n, (777,), device="cuda", generator=g)
    assert torch.equal(t.fetch(ids), full[ids])


@pytest.mark.parametrize("F,dtype", [(128, torch.float32), (602, torch.float32), (100, torch.float32),
                                     (602, torch.bfloat16), (1000, torch.float32), (40, torch.float32)])
def test_spmm_csr_sharded_equals_oracle_and_single_table(K, F, dtype):
    """The aggregation that reads a node-range-partitioned table in place (the halo fetch fused into the SpMM): same
    numbers as the oracle and as the single-table kernel, for sum/mean, with and without edge values, ragged rows."""
    rng = np.random.default_rng(F)
    n_src, n_dst, P = 9001, 1200, 4
    part = (n_src + P - 1) // P
    ld = (F + 3) // 4 * 4 if dtype == torch.float32 else (F + 7) // 8 * 8
    rp, col = rand_csr(rng, n_dst, n_src, 30, heavy=[(3, 700)], empty_frac=0.1)
    x = rng.standard_normal((n_src, F)).astype(np.float32)
    full = padded(x, ld).to(dtype)
    xr = full[:, :F].float().cpu().numpy()
    vals = rng.random(col.size).astype(np.float32) + 0.1
    shards = [full[r * part:min(n_src, (r + 1) * part)].clone() for r in range(P)]
    ptrs = torch.tensor([s.data_ptr() for s in shards], dtype=torch.int64, device="cuda")
    sb = ld * full.element_size()
    for rpt in (np.int64, np.int32):
        for red in ("sum", "mean"):
            for v in (None, vals):
                ref = oracle.spmm_csr(rp, col, xr, values=v, reduce=red)
                out = K.spmm_csr_sharded(dev(rp.astype(rpt)), dev(col), ptrs, part, sb, F, dtype=dtype,
                                         values=None if v is None else dev(v), reduce=red)
                assert rel_err(out.cpu().numpy(), ref) <= FP32_TOL, (red, v is not None)
    one = torch.tensor([full.data_ptr()], dtype=torch.int64, device="cuda")
    a = K.spmm_csr_sharded(dev(rp), dev(col), one, n_src, sb, F, dtype=dtype)
    b = K.spmm_csr_sharded(dev(rp), dev(col), ptrs, part, sb, F, dtype=dtype)
    assert torch.equal(a, b)                                   # the shard lookup does not change the arithmetic
    with pytest.raises(Exception):
        K.spmm_csr_sharded(dev(rp), dev(col), ptrs, part, sb + 4, F, dtype=dtype)   # rows must be 16-byte multiples


def test_library_options_roundtrip(K):
    K.set_option("spmm_kernel", "stream")
    assert K.get_option("spmm_kernel") == 2
    K.set_option("spmm_kernel", None)
    assert K.get_option("spmm_kernel") == 0
    K.set_option("rows_d", 5)
    assert K.get_option("rows_d") == 5
    K.set_option("rows_d", None)
    with pytest.raises(Exception):
        K.set_option("no_such_option", 1)
    with pytest.raises(Exception):
        K.set_option("spmm_kernel", "no_such_family")


def test_ipc_export_import_roundtrip_same_process(K):
    """The IPC handle of a tensor's allocation + offset names the same bytes (cudaIpcOpenMemHandle cannot be opened in
    the exporting process, so only the export side and the offset arithmetic are checked here; the cross-process
    path runs in tools/bench_halo.py --halo peer on >= 2 GPUs)."""
    big = torch.empty(1 << 20, device="cuda")
    view = big[12345:]
    h, off = K.ipc_export(view)
    assert len(h) == 64 and off >= 12345 * 4 and off % 4 == 0
    h2, off2 = K.ipc_export(big)
    assert h2 == h and off - off2 == 12345 * 4


# ---------------------------------------------------------- block builder --
@pytest.mark.parametrize("fanouts", [(25, 10), (5,), (3, 3, 3)])
def test_device_block_builder_matches_sort_based_compaction(K, fanouts):
    """dgllb_build_block (hash table + first-occurrence flags + scan) == the torch.unique formulation, bit for bit:
    src ids in first-occurrence order (dst first), relabelled columns, counts; and the block invariants hold."""
    from dgll_b200 import graphs as G
    rng = np.random.default_rng(len(fanouts))
    n = 30000
    rp, col = rand_csr(rng, n, n, 40, heavy=[(11, 3000)], empty_frac=0.1)
    g = torch.Generator(device="cuda").manual_seed(3)
    seeds = torch.randperm(n, device="cuda", generator=g)[:512]
    a = G.sample_blocks(dev(rp), dev(col), seeds, fanouts, rng_seed=5, builder="device")
    b = G.sample_blocks(dev(rp), dev(col), seeds, fanouts, rng_seed=5, builder="torch")
    assert len(a) == len(b) == len(fanouts)
    for x, y in zip(a, b):
        assert torch.equal(x.src_ids, y.src_ids) and torch.equal(x.col, y.col)
        assert torch.equal(x.col_global, y.col_global) and torch.equal(x.row_ptr, y.row_ptr)
        assert x.num_dst == y.num_dst and x.num_src == y.num_src
        assert torch.equal(x.src_ids[x.col.long()], x.col_global.long())      # relabelling is consistent
        assert x.src_ids.unique().numel() == x.num_src                          # no duplicates
    assert a[-1].src_ids[:512].equal(seeds) and a[0].num_dst == a[1].num_src if len(a) > 1 else True
    # empty neighbourhoods only
    rp0 = torch.zeros(4, dtype=torch.int32, device="cuda")
    s, c, cnt = K.build_block(torch.tensor([7, 3, 9], device="cuda"), rp0, torch.zeros(0, dtype=torch.int32, device="cuda"))
    assert cnt.tolist() == [3, 0] and s[:3].tolist() == [7, 3, 9]


# ------------------------------------------------------ attention dropout --
@pytest.mark.parametrize("kernel", ["group", "row"])
@pytest.mark.parametrize("mode", ["softmax", "exp_neg"])
def test_gat_attention_dropout_forward_backward_with_replayed_mask(K, kernel, mode, gat_kernel):
    """The kernels' dropout mask is exported (dgllb_gat_dropout_mask) and replayed in an fp64 autograd restatement of
    gatconv.py:30-54 / :111-148 WITH dropout on the attention coefficients: outputs and all three gradients match."""
    gat_kernel(kernel)
    rng = np.random.default_rng(3)
    n, heads, D, pdrop, seed = 300, 4, 32, 0.4, 123456789
    rp, col = rand_csr(rng, n, n, 25, heavy=[(9, 400)])
    wh = rng.standard_normal((n, heads * D)).astype(np.float32)
    el = rng.standard_normal((n, heads)).astype(np.float32)
    er = rng.standard_normal((n, heads)).astype(np.float32)
    g = rng.standard_normal((n, heads * D)).astype(np.float32)
    mask = K.gat_dropout_mask(seed, col.size, heads, pdrop)
    kept = (mask > 0).float().mean().item()
    assert abs(kept - (1 - pdrop)) < 0.02 and torch.all((mask == 0) | ((mask - 1 / (1 - pdrop)).abs() < 1e-6))
    rows = torch.from_numpy(np.repeat(np.arange(n), rp[1:] - rp[:-1])).long()
    cols = torch.from_numpy(col).long()
    twh = torch.from_numpy(wh).double().requires_grad_(True)
    tel = torch.from_numpy(el).double().requires_grad_(True)
    ter = torch.from_numpy(er).double().requires_grad_(True)
    z = torch.nn.functional.leaky_relu(tel[rows] + ter[cols], 0.2)
    sgn = z if mode == "softmax" else -z
    smax = torch.full((n, heads), -float("inf"), dtype=torch.float64).scatter_reduce(
        0, rows[:, None].expand(-1, heads), sgn.detach(), reduce="amax")
    ex = torch.exp(sgn - smax[rows])
    den = torch.zeros(n, heads, dtype=torch.float64).index_add_(0, rows, ex)
    alpha = ex / den[rows] * mask.double().cpu()                 # dropout AFTER normalisation, as the reference
    msg = alpha[:, :, None] * twh[cols].view(-1, heads, D)
    ref = torch.zeros(n, heads, D, dtype=torch.float64).index_add_(0, rows, msg).view(n, heads * D)
    ref.backward(torch.from_numpy(g).double())
    drp, dcol = dev(rp), dev(col)
    plan = K.CsrPlan(drp, chunk_edges=128) if kernel == "row" else None
    out, rmax, rsum = K.gat_forward(drp, dcol, dev(wh), dev(el), dev(er), heads, 0.2, mode=mode, save_stats=True,
                                    dropout=pdrop, seed=seed, plan=plan)
    assert rel_err(out.cpu().numpy(), ref.detach().numpy()) <= FP32_TOL
    trp, tcol, _, perm = K.csr_transpose(drp, dcol, n, want_perm=True)
    d_wh, d_el, d_er = K.gat_backward(drp, dcol, trp, tcol, perm, dev(wh), dev(el), dev(er), out, rmax, rsum, dev(g),
                                      heads, 0.2, mode=mode, dropout=pdrop, seed=seed)
    assert rel_err(d_wh.cpu().numpy(), twh.grad.numpy()) <= 2e-5
    assert rel_err(d_el.cpu().numpy(), tel.grad.numpy()) <= 2e-5
    assert rel_err(d_er.cpu().numpy(), ter.grad.numpy()) <= 2e-5
    # p = 0 is bit-identical to the call without dropout arguments
    a = K.gat_forward(drp, dcol, dev(wh), dev(el), dev(er), heads, 0.2, mode=mode)
    b = K.gat_forward(drp, dcol, dev(wh), dev(el), dev(er), heads, 0.2, mode=mode, dropout=0.0, seed=99)
    assert torch.equal(a, b)
