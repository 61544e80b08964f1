"""Samplers and the mini-batch iterator with the reference's names and return structure
(dgll/sampling/base_sampler.py:4-109, dgllsampler.py:6-21, dgll/dataloader/dataloader.py:4-24,
dgll/nn/utils/utils.py:52-68), plus the DGL-named block sampler the GPU-Accelerator scripts use.

Two modes:
  host (default)   exactly the reference's arithmetic on the host — Python ``random.sample`` without replacement /
                   ``np.random.choice`` with replacement — so a seeded run reproduces the reference's index lists
                   BIT-EXACTLY (the parity protocol of SURVEY.md App. B); the lists then feed the device kernels.
  device           ``dgllb_sample_neighbors`` (Floyd sampling, counter-based RNG) — same structure, same
                   distribution, not the same draws; validated structurally.
"""
import random

import numpy as np
import torch

from .. import graphs as G
from .. import kernels as K


class sugbraph():
    """base_sampler.py:65-109 (name kept as spelled upstream): an edge list block in GLOBAL ids."""

    def __init__(self, src_data, dst_data, counts=None):
        self.src_data = src_data
        self.dst_data = dst_data
        self.graph_nodes = torch.unique(torch.cat((self.src_data, dst_data)))
        self._counts = counts  # edges per seed position, when the sampler that built the block knows them

    def src_nodes(self):
        return self.src_data

    def dst_nodes(self):
        return self.dst_data

    def nodes(self):
        return self.graph_nodes

    def num_src_nodes(self):
        return self.src_data.shape[0]

    def num_dst_nodes(self):
        return self.dst_data.shape[0]

    def get_features(self, g, subgs):
        all_nodes = torch.cat([subg.nodes() for subg in subgs])
        return g.get_features(torch.unique(all_nodes))

    # -- device view: CSR by destination ROW POSITION (row r = r-th seed), global source ids --
    def to_csr(self, seeds):
        """Edges are emitted seed by seed (base_sampler.py:34-40) so each seed owns one contiguous run."""
        dst = self.dst_data
        seeds = seeds.to(dst.device)
        n = seeds.numel()
        if self._counts is not None and len(self._counts) == n:
            rp = torch.zeros(n + 1, dtype=torch.int64, device=dst.device)
            rp[1:] = torch.cumsum(torch.as_tensor(self._counts, dtype=torch.int64, device=dst.device), 0)
            return rp, self.src_data.to(torch.int32)
        if dst.numel() == 0:
            return torch.zeros(n + 1, dtype=torch.int64, device=dst.device), self.src_data.to(torch.int32)
        # run boundaries: a new run starts where dst changes OR where the seed list repeats a value consecutively;
        # the reference's per-seed loop makes run k belong to the k-th seed that has >= 1 sampled neighbour
        change = torch.ones(dst.numel(), dtype=torch.bool, device=dst.device)
        change[1:] = dst[1:] != dst[:-1]
        run_id = torch.cumsum(change, 0) - 1
        run_dst = dst[change]
        counts_run = torch.bincount(run_id, minlength=run_dst.numel())
        # map runs onto seed positions in order (seeds without neighbours have no run)
        rp = torch.zeros(n + 1, dtype=torch.int64, device=dst.device)
        seeds_l, run_l, cnt_l = seeds.tolist(), run_dst.tolist(), counts_run.tolist()
        k = 0
        counts = [0] * n
        for r, v in enumerate(seeds_l):
            if k < len(run_l) and run_l[k] == v:
                counts[r] = cnt_l[k]
                k += 1
        assert k == len(run_l), "dst is not a seed-ordered edge list"
        rp[1:] = torch.cumsum(torch.tensor(counts, dtype=torch.int64, device=dst.device), 0)
        return rp, self.src_data.to(torch.int32)


class Base_sampler(object):
    """base_sampler.py:4-63."""

    def __init__(self, device_sampling=False, rng_seed=0):
        self.device_sampling = device_sampling
        self.rng_seed = rng_seed
        self._calls = 0

    def sample(self, g, nodes):
        raise NotImplementedError

    def _subgraph(self, nodes, neighbors_list):
        src_list, dst_list = [], []
        for i, neighbors in enumerate(neighbors_list):
            dst_node = int(nodes[i])
            for src_node in neighbors:
                src_list.append(src_node)
                dst_list.append(dst_node)
        dev = nodes.device if isinstance(nodes, torch.Tensor) else None
        return sugbraph(torch.tensor(src_list, dtype=torch.int64, device=dev),
                        torch.tensor(dst_list, dtype=torch.int64, device=dev),
                        counts=[len(nb) for nb in neighbors_list])

    def sample_neighbours(self, g, nodes, fanout=None):
        """base_sampler.py:45-58."""
        if self.device_sampling:
            rp, col = g.csr()
            seeds = nodes.to(rp.device).to(torch.int64)
            b_rp, b_col = K.sample_neighbors(rp, col, seeds, -1 if fanout is None else fanout,
                                             rng_seed=self.rng_seed * 1000003 + self._calls)
            self._calls += 1
            cnt = (b_rp[1:] - b_rp[:-1]).to(torch.int64)
            nnz = int(b_rp[-1].item())
            return sugbraph(b_col[:nnz].to(torch.int64), torch.repeat_interleave(seeds, cnt), counts=cnt)
        neighbors_list = g.get_neighbors(nodes)
        random_neighbors = []
        for neighbors in neighbors_list:
            if len(neighbors) == 0:
                random_neighbors.append([])
            elif fanout is None:
                random_neighbors.append(neighbors)
            elif len(neighbors) <= fanout:
                random_neighbors.append(neighbors)
            else:
                random_neighbors.append(random.sample(neighbors, fanout))
        return self._subgraph(nodes, random_neighbors)

    def get_adj(self, g, subgs):
        all_nodes = torch.cat([subg.nodes() for subg in subgs])
        return g.get_induced_subgraph(torch.unique(all_nodes))


class DGLLNeighborSampler(Base_sampler):
    """dgllsampler.py:6-21 — fanouts consumed in REVERSE; next seeds = raw src list WITH duplicates (:17)."""

    def __init__(self, fanouts, device_sampling=False, rng_seed=0):
        super().__init__(device_sampling, rng_seed)
        self.fanouts = fanouts

    def sample(self, g, seed_nodes):
        output_nodes = seed_nodes
        subgs = []
        input_nodes = seed_nodes
        for fanout in reversed(self.fanouts):
            subg = self.sample_neighbours(g, seed_nodes, fanout)
            seed_nodes = subg.src_nodes()
            subgs.insert(0, subg)
            input_nodes = seed_nodes
        return input_nodes, output_nodes, subgs


class DataLoader:
    """dataloader.py:4-24 — mini-batches of ``train_nodes`` through ``sampler.sample`` (the reference's loop refers
    to undefined ``self.data`` / ``batch_size``; this is the evident intent)."""

    def __init__(self, Dgraph, train_nodes, sampler, batch_size=1, device=None, shuffle=False, drop_last=False):
        self.Dgraph, self.sampler, self.train_nodes = Dgraph, sampler, train_nodes
        self.batch_size, self.device = batch_size, device
        self.shuffle, self.drop_last = shuffle, drop_last

    def sample(self):
        nodes = self.train_nodes
        if self.shuffle:
            nodes = nodes[torch.randperm(len(nodes))]
        for i in range(0, len(nodes), self.batch_size):
            seed_nodes = nodes[i:i + self.batch_size]
            if self.drop_last and len(seed_nodes) < self.batch_size:
                break
            yield self.sampler.sample(self.Dgraph, seed_nodes)

    def __iter__(self):
        return self.sample()

    def __len__(self):
        n = len(self.train_nodes)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size


# ---- fixed-fanout sampler with replacement (dgll/nn/utils/utils.py:52-68) ----
def sampling(src_nodes, sample_num, neighbor_table):
    results = []
    for sid in src_nodes:
        res = np.random.choice(neighbor_table[sid], size=(sample_num,))
        results.append(res)
    return np.asarray(results).flatten()


def multihop_sampling(src_nodes, sample_nums, neighbor_table):
    sampling_result = [src_nodes]
    for k, hopk_num in enumerate(sample_nums):
        hopk_result = sampling(sampling_result[k], hopk_num, neighbor_table)
        sampling_result.append(hopk_result)
    return sampling_result


# ---- DGL-named block API (GPU Accelerator/MQGCN.py:114-137, MQFastGCN.py:82-84) ----
def create_block(data, num_src_nodes=None, num_dst_nodes=None, device=None):
    """``dgll.create_block(('csc', (indptr, indices, [])))``: a CSC of shape [num_src, num_dst] = CSR by destination."""
    fmt, (indptr, indices, _eids) = data
    if fmt != "csc":
        raise ValueError("create_block: only the 'csc' format the reference uses is supported")
    indptr = torch.as_tensor(indptr)
    indices = torch.as_tensor(indices)
    dev = torch.device(device) if device is not None else (indptr.device if indptr.is_cuda else torch.device("cuda"))
    n_dst = indptr.numel() - 1 if num_dst_nodes is None else num_dst_nodes
    n_src = int(indices.max().item()) + 1 if (num_src_nodes is None and indices.numel()) else (num_src_nodes or 0)
    n_src = max(n_src, n_dst)
    col = indices.to(dev).to(torch.int32)
    blk = G.Block(indptr.to(dev).to(torch.int64), col, col, torch.arange(n_src, device=dev), n_dst)
    blk.num_src = n_src
    return blk


class NeighborSampler:
    """``dgll.dataloading.NeighborSampler(fanouts)``: device-side sampling + dst-first compaction per layer."""

    def __init__(self, fanouts, rng_seed=0):
        self.fanouts = list(fanouts)
        self.rng_seed = rng_seed
        self._calls = 0

    def sample(self, g, seed_nodes):
        rp, col = g.csr() if hasattr(g, "csr") else g
        seeds = seed_nodes.to(rp.device).to(torch.int64)
        blocks = G.sample_blocks(rp, col, seeds, self.fanouts, rng_seed=self.rng_seed + 7919 * self._calls)
        self._calls += 1
        return blocks[0].src_ids, seeds, blocks


class BlockDataLoader:
    """``dgll.dataloading.DataLoader(graph, nids, sampler, device=, batch_size=, shuffle=, drop_last=, use_ddp=,
    num_workers=)`` yielding ``(input_nodes, output_nodes, mfgs)`` (MQGCN.py:117-137).  ``use_ddp`` shards ``nids``
    by rank exactly like a DistributedSampler; ``num_workers`` is accepted and ignored (sampling runs on the GPU)."""

    def __init__(self, graph, nids, sampler, device=None, batch_size=1, shuffle=False, drop_last=False,
                 use_ddp=False, num_workers=0, seed=0):
        self.graph, self.sampler = graph, sampler
        self.batch_size, self.shuffle, self.drop_last = batch_size, shuffle, drop_last
        self.device = device
        self.epoch, self.seed = 0, seed
        nids = torch.as_tensor(nids)
        if use_ddp and torch.distributed.is_available() and torch.distributed.is_initialized():
            r, w = torch.distributed.get_rank(), torch.distributed.get_world_size()
            per = (nids.numel() + w - 1) // w
            nids = nids[r * per:(r + 1) * per]
        self.nids = nids

    def set_epoch(self, epoch):
        self.epoch = epoch

    def __len__(self):
        n = self.nids.numel()
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        nids = self.nids
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            nids = nids[torch.randperm(nids.numel(), generator=g).to(nids.device)]
        for i in range(0, nids.numel(), self.batch_size):
            seeds = nids[i:i + self.batch_size]
            if self.drop_last and seeds.numel() < self.batch_size:
                break
            yield self.sampler.sample(self.graph, seeds)
