"""Full-batch GCN the way the reference's users write it (dgll/nn/Convolution/gcnconv.py:43-58 + nn/utils/utils.py
load_data): scipy adjacency -> D^-1(A+I) -> torch sparse COO -> GCN(x, adj) -> nll_loss, a few Adam steps.
Only the imports change: dgll_b200.nn / dgll_b200.nn.utils instead of dgll.nn.  Runs on a small synthetic graph."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout

import numpy as np
import scipy.sparse as sp
import torch

from dgll_b200.nn import GCN
from dgll_b200.nn.utils import accuracy, normalize, sparse_mx_to_torch_sparse_tensor


def main(n=3000, feats=64, classes=7, epochs=20, seed=0):
    rng = np.random.default_rng(seed)
    a = sp.random(n, n, density=8.0 / n, random_state=seed, format="coo")
    a.data[:] = 1.0
    a = a + a.T.multiply(a.T > a) - a.multiply(a.T > a)            # symmetrise (utils.py:168-169)
    adj = sparse_mx_to_torch_sparse_tensor(normalize(a + sp.eye(n)), device="cuda")
    x = torch.from_numpy(rng.standard_normal((n, feats)).astype(np.float32)).cuda()
    labels = torch.from_numpy(rng.integers(0, classes, size=n)).cuda()
    model = GCN(feats, 32, classes, dropout=0.5).cuda()
    opt = torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=5e-4)
    first = last = None
    for epoch in range(epochs):
        model.train()
        opt.zero_grad()
        out = model(x, adj)
        loss = torch.nn.functional.nll_loss(out, labels)
        loss.backward()
        opt.step()
        first = loss.item() if first is None else first
        last = loss.item()
    model.eval()
    acc = accuracy(model(x, adj), labels).item()
    print("gcn_full_batch: loss %.4f -> %.4f, train accuracy %.3f" % (first, last, acc))
    return first, last


if __name__ == "__main__":
    main()
