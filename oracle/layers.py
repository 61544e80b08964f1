"""torch-CPU restatement of the reference's layer arithmetic — TEST INFRASTRUCTURE ONLY.

Functional (weights are arguments) so tests can load the golden weights that
``oracle/gen_golden.py`` captured from the reference's own classes.  Each
function cites the reference lines it restates.  All math in float64 unless
``dtype`` says otherwise, so the oracle is a tighter truth than either fp32 path.
Never imported by anything under ``dgll_b200/``.
"""
import numpy as np
import torch
import torch.nn.functional as Fn


def _t(a, dtype=torch.float64):
    if isinstance(a, torch.Tensor):
        return a.detach().to(dtype)
    return torch.as_tensor(np.asarray(a)).to(dtype)


def coo_adj(indices, values, n, dtype=torch.float64):
    """torch sparse COO [n,n] from int indices [2,E] (row = destination)."""
    idx = torch.as_tensor(np.asarray(indices)).long()
    return torch.sparse_coo_tensor(idx, _t(values, dtype), (n, n)).coalesce()


# ------------------------------------------------------------------- GCN ---
def gcn_conv(x, adj, weight, bias=None):
    """gcnConv.forward, dgll/nn/Convolution/gcnconv.py:29-35: spmm(adj, x @ W) + b."""
    support = torch.mm(x, weight)
    out = torch.sparse.mm(adj, support)
    return out + bias if bias is not None else out


def gcn_model(x, adj, w1, b1, w2, b2):
    """GCN.forward in eval mode, gcnconv.py:53-58: relu -> (dropout off) -> layer -> log_softmax."""
    h1 = torch.relu(gcn_conv(x, adj, w1, b1))
    return torch.log_softmax(gcn_conv(h1, adj, w2, b2), dim=1)


def normalize_rows(mx):
    """normalize(), dgll/nn/utils/utils.py:240-247: D^-1 M with 1/0 := 0 (dense numpy in/out)."""
    mx = np.asarray(mx, dtype=np.float64)
    rowsum = mx.sum(1)
    with np.errstate(divide="ignore"):
        r_inv = np.power(rowsum, -1.0)
    r_inv[np.isinf(r_inv)] = 0.0
    return r_inv[:, None] * mx


def gcn_adjacency(a_raw):
    """load_data's adjacency prep, utils.py:168-171: symmetrise by max, add I, row-normalise."""
    a = np.asarray(a_raw, dtype=np.float64)
    at = a.T
    a = a + at * (at > a) - a * (at > a)
    return normalize_rows(a + np.eye(a.shape[0]))


def sym_norm_adjacency(a_bin):
    """FusedKernel/train_gcn.py:64-71: D^-1/2 A D^-1/2 on the undirected binary A, no self loops, 1/sqrt(0) := 0."""
    a = np.asarray(a_bin, dtype=np.float64)
    a = ((a + a.T) > 0).astype(np.float64)
    d = a.sum(1)
    with np.errstate(divide="ignore"):
        dis = 1.0 / np.sqrt(d)
    dis[d == 0] = 0.0
    return dis[:, None] * a * dis[None, :]


def fused_gcn_layer(x, a_hat, w):
    """relu(A_hat (X W)), dgll/FusedKernel/gcn_fused_kernel.cu:26-70 (+ ReLU always, :67)."""
    return torch.relu(a_hat @ (x @ w))


# ---------------------------------------------------------- PPI GCN (C1) ---
def ppi_gcn_layer(edge_index, x, w):
    """GCNLayer.forward, Evaluation/PPI/gcn_model.py:68-77: relu(sparse.mm(coo(edge_index, ones), x @ W)).
    Row index of the COO is edge_index[0] (gcn_model.py:56)."""
    n = x.size(0)
    ei = torch.as_tensor(np.asarray(edge_index)).long()
    adj = torch.sparse_coo_tensor(ei, torch.ones(ei.shape[1], dtype=x.dtype), (n, n))
    return torch.relu(torch.sparse.mm(adj, torch.mm(x, w)))


def ppi_gcn(edge_index, x, ws, w_out, b_out):
    """GCN.forward, gcn_model.py:89-94: num_layers GCNLayers then nn.Linear."""
    h = x
    for w in ws:
        h = ppi_gcn_layer(edge_index, h, w)
    return h @ w_out.t() + b_out


def ppi_loss(logits, labels):
    """CrossEntropyLoss with float multi-hot targets, Evaluation/PPI/train_gcn.py:26,45."""
    return Fn.cross_entropy(logits, labels)


# ------------------------------------------------------------------- GAT ---
def gat_dense_layer(h, adj, W, a, alpha, concat=True):
    """gatConv.forward, gatconv.py:30-54 (attention dropout off)."""
    Wh = h @ W
    D = W.shape[1]
    e = Fn.leaky_relu(Wh @ a[:D, :] + (Wh @ a[D:, :]).T, alpha)
    att = torch.where(adj > 0, e, torch.full_like(e, -9e15))
    att = torch.softmax(att, dim=1)
    hp = att @ Wh
    return Fn.elu(hp) if concat else hp


def gat_sparse_layer(h, adj, W, a, alpha, concat=True):
    """sparseGatConv.forward, gatconv.py:111-148: exp(-leakyrelu(a.[h_i||h_j])), row-sum normalised."""
    n = h.size(0)
    edge = adj.nonzero().t()
    hw = h @ W
    edge_h = torch.cat((hw[edge[0]], hw[edge[1]]), dim=1).t()
    edge_e = torch.exp(-Fn.leaky_relu(a.mm(edge_h).squeeze(), alpha))
    rowsum = torch.zeros(n, 1, dtype=h.dtype).index_add_(0, edge[0], edge_e[:, None])
    hp = torch.zeros(n, hw.size(1), dtype=h.dtype).index_add_(0, edge[0], edge_e[:, None] * hw[edge[1]])
    hp = hp / rowsum
    return Fn.elu(hp) if concat else hp


def gat_model(x, adj, Ws, As, W_out, a_out, alpha, sparse):
    """GAT / SpGAT.forward in eval mode, gatconv.py:180-199: heads concat -> out layer -> elu -> log_softmax."""
    layer = gat_sparse_layer if sparse else gat_dense_layer
    h = torch.cat([layer(x, adj, W, a, alpha, True) for W, a in zip(Ws, As)], dim=1)
    return torch.log_softmax(Fn.elu(layer(h, adj, W_out, a_out, alpha, False)), dim=1)


def special_spmm_backward(indices, values, b, grad_out):
    """SpecialSpmmFunction.backward, gatconv.py:71-81 (the dense N x N product replaced by its definition)."""
    i, j = indices[0].long(), indices[1].long()
    g_values = (grad_out[i] * b[j]).sum(1)
    n = b.size(0)
    g_b = torch.zeros_like(b).index_add_(0, j, values[:, None] * grad_out[i])
    return g_values, g_b


# ------------------------------------------------------------- GraphSAGE ---
def sage_conv(src, neigh, w_self, w_neigh, b_neigh=None, aggr="mean", combine="sum", activation=True):
    """sageConv.forward as INTENDED, sageconv.py:32-45,70-83.  Two documented fixes to the shipped code
    (SURVEY.md §8 a8): the reduction over K is assigned (the reference discards it, :33-38) and max takes the
    values.  act(src @ W_s  (+|cat)  reduce_K(neigh) @ W_n [+ b_n])."""
    if aggr == "mean":
        agg = neigh.mean(dim=1)
    elif aggr == "sum":
        agg = neigh.sum(dim=1)
    elif aggr == "max":
        agg = neigh.max(dim=1).values
    else:
        raise ValueError("Unsupported aggr_method, expected mean, sum, max, but got {}".format(aggr))
    nh = agg @ w_neigh
    if b_neigh is not None:
        nh = nh + b_neigh
    sh = src @ w_self
    if combine == "sum":
        hid = sh + nh
    elif combine == "concat":
        hid = torch.cat([sh, nh], dim=1)
    else:
        raise ValueError("Expected sum or concat, got {}".format(combine))
    return torch.relu(hid) if activation else hid


def graphsage_model(feature_list, layers, fanouts):
    """GraphSage.forward, sageconv.py:103-114: layer l runs on hops 0..L-l-1 with hidden[hop+1].view(B, K_l, -1)."""
    hidden = list(feature_list)
    L = len(fanouts)
    for l in range(L):
        nxt = []
        for hop in range(L - l):
            src = hidden[hop]
            nxt.append(sage_conv(src, hidden[hop + 1].view(len(src), fanouts[l], -1), *layers[l]))
        hidden = nxt
    return hidden[0]


def dgl_sage_conv_mean(row_ptr, col, x_src, n_dst, w_self, w_neigh, bias=None):
    """DGL 2.4 SAGEConv('mean') on a block (third-party; restated from public docs):
    h_i = W_self x_i + W_neigh mean_{j in N(i)} x_j + b; mean over an empty set = 0; dst nodes are the first
    n_dst src nodes (GPU Accelerator/MQGCN.py:45,48).  Weights are [in, out]."""
    rp = np.asarray(row_ptr, dtype=np.int64)
    deg = torch.as_tensor(rp[1:] - rp[:-1]).to(x_src.dtype)
    rows = torch.as_tensor(np.repeat(np.arange(n_dst), rp[1:] - rp[:-1])).long()
    cols = torch.as_tensor(np.asarray(col)).long()
    agg = torch.zeros(n_dst, x_src.size(1), dtype=x_src.dtype).index_add_(0, rows, x_src[cols])
    agg = agg / deg.clamp(min=1)[:, None]
    out = x_src[:n_dst] @ w_self + agg @ w_neigh
    return out + bias if bias is not None else out


def dgl_graph_conv_both(row_ptr, col, x_src, n_dst, weight, bias=None, relu=False):
    """DGL 2.4 GraphConv(norm='both') on a block (restated):
    h = D_in^-1/2 * (A (D_out^-1/2 * x_src)) W + b, degrees from the block, clamped >= 1."""
    rp = np.asarray(row_ptr, dtype=np.int64)
    cols = torch.as_tensor(np.asarray(col)).long()
    n_src = x_src.size(0)
    out_deg = torch.bincount(cols, minlength=n_src).to(x_src.dtype).clamp(min=1)
    in_deg = torch.as_tensor(rp[1:] - rp[:-1]).to(x_src.dtype).clamp(min=1)
    rows = torch.as_tensor(np.repeat(np.arange(n_dst), rp[1:] - rp[:-1])).long()
    xs = x_src * out_deg.pow(-0.5)[:, None]
    agg = torch.zeros(n_dst, x_src.size(1), dtype=x_src.dtype).index_add_(0, rows, xs[cols])
    out = (agg * in_deg.pow(-0.5)[:, None]) @ weight
    if bias is not None:
        out = out + bias
    return torch.relu(out) if relu else out


# ------------------------------------------------------------ GIN / pool ---
def gin_conv(adj, feat, w, b):
    """GinConv.forward, ginconv.py:27-28: relu(Linear(Feat + Adj @ Feat)); w is nn.Linear weight [out,in]."""
    return torch.relu((feat + adj @ feat) @ w.t() + b)


def gin_model(A, X, sd):
    """GIN.forward, ginconv.py:53-66; sd = golden state-dict (keys with '.' replaced by '__')."""
    g = lambda k: _t(sd[k])
    X = X @ g("in_proj__weight").t() + g("in_proj__bias")
    hs = [X]
    i = 0
    while "convs__%d__linear__weight" % i in sd:
        X = gin_conv(A, X, g("convs__%d__linear__weight" % i), g("convs__%d__linear__bias" % i))
        hs.append(X)
        i += 1
    X = torch.cat(hs, dim=2).sum(dim=1)
    return X @ g("out_proj__weight").t() + g("out_proj__bias")


def pooling(x, batch, size=None, reduce="sum"):
    """sumPooling/meanPooling/maxPooling, GlobalPooling/Pooling.py:18-81 = torch_scatter.scatter(dim=0)."""
    if batch is None:
        if reduce in ("sum", "add"):
            return x.sum(0, keepdim=True)
        return x.mean(0, keepdim=True) if reduce == "mean" else x.max(0, keepdim=True)[0]
    size = int(batch.max().item() + 1) if size is None else size
    out = torch.zeros(size, x.size(1), dtype=x.dtype)
    if reduce in ("sum", "add", "mean"):
        out.index_add_(0, batch, x)
        if reduce == "mean":
            cnt = torch.bincount(batch, minlength=size).clamp(min=1).to(x.dtype)
            out = out / cnt[:, None]
        return out
    out = torch.full((size, x.size(1)), float("-inf"), dtype=x.dtype)
    out = out.scatter_reduce(0, batch[:, None].expand_as(x), x, reduce="amax")
    out[torch.isinf(out)] = 0  # torch_scatter fills empty segments with 0
    return out
