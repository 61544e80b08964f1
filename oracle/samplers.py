"""Restatement of the reference's host data path — TEST INFRASTRUCTURE ONLY.

Graph store, neighbour samplers, induced adjacency and the feature-cache split,
each following the cited reference lines, on plain numpy / Python so results are
bit-exact against the golden vectors (``tests/golden/sampler_*.npz``).
Adjacency is passed as CSR-like (nbr_ptr, nbr_flat) = the python adjacency lists
of ``DGraph.edges`` (dgll/data/dgraph.py:49-62), order preserved.
"""
import random

import numpy as np


def neighbors_of(nbr_ptr, nbr_flat, v):
    return [int(x) for x in nbr_flat[nbr_ptr[v]:nbr_ptr[v + 1]]]


def sample_neighbours(nbr_ptr, nbr_flat, nodes, fanout, rng=random):
    """Base_sampler.sample_neighbours + _subgraph, dgll/sampling/base_sampler.py:30-58.

    Per seed in order: [] if no neighbours; all (original order) if fanout is None or deg <= fanout;
    else ``random.sample(neighbors, fanout)`` (Python's global Mersenne Twister, without replacement).
    Block = concatenation of (src = neighbour, dst = seed).  Returns (src, dst) int64 arrays."""
    src, dst = [], []
    for v in nodes:
        v = int(v)
        nb = neighbors_of(nbr_ptr, nbr_flat, v)
        if len(nb) == 0:
            picked = []
        elif fanout is None or len(nb) <= fanout:
            picked = nb
        else:
            picked = rng.sample(nb, fanout)
        src.extend(picked)
        dst.extend([v] * len(picked))
    return np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64)


def neighbor_sampler(nbr_ptr, nbr_flat, seeds, fanouts, rng=random):
    """DGLLNeighborSampler.sample, dgll/sampling/dgllsampler.py:10-21: fanouts consumed in REVERSE (the last
    entry applies to the output layer); next seeds = raw src list WITH duplicates (:17); blocks are inserted
    at the front so blocks[0] is the input layer.  Returns (input_nodes, output_nodes, [(src, dst), ...])."""
    output_nodes = np.asarray(seeds, dtype=np.int64)
    cur = output_nodes
    blocks = []
    for fanout in reversed(list(fanouts)):
        src, dst = sample_neighbours(nbr_ptr, nbr_flat, cur, fanout, rng)
        cur = src
        blocks.insert(0, (src, dst))
    return cur, output_nodes, blocks


def block_nodes(src, dst):
    """sugbraph.graph_nodes, base_sampler.py:81: sorted unique of cat(src, dst)."""
    return np.unique(np.concatenate([src, dst]))


def induced_subgraph(nbr_ptr, nbr_flat, nodes):
    """DGraph.get_induced_subgraph, dgll/data/dgraph.py:64-81: dense int32 [n,n], adj[pos(u), pos(w)] = 1 for
    w in edges[u] if w in nodes; later duplicates in ``nodes`` win the position map (dict comprehension)."""
    nodes = [int(x) for x in nodes]
    n = len(nodes)
    out = np.zeros((n, n), dtype=np.int32)
    mapping = {j: i for i, j in enumerate(nodes)}
    for u in nodes:
        for w in neighbors_of(nbr_ptr, nbr_flat, u):
            if w in mapping:
                out[mapping[u], mapping[w]] = 1
    return out


def get_adj(nbr_ptr, nbr_flat, blocks):
    """Base_sampler.get_adj, base_sampler.py:60-63: induced adjacency over the sorted unique nodes of all blocks."""
    allnodes = np.unique(np.concatenate([block_nodes(s, d) for s, d in blocks]))
    return induced_subgraph(nbr_ptr, nbr_flat, allnodes), allnodes


def multihop_sampling(nbr_ptr, nbr_flat, src_nodes, sample_nums, rng=np.random):
    """sampling + multihop_sampling, dgll/nn/utils/utils.py:52-68: ``np.random.choice(neighbor_tab[v], size=(K,))``
    WITH replacement from the legacy global numpy RNG, flattened per hop."""
    result = [np.asarray(src_nodes)]
    for k, num in enumerate(sample_nums):
        hop = []
        for v in result[k]:
            hop.append(rng.choice(neighbors_of(nbr_ptr, nbr_flat, int(v)), size=(num,)))
        result.append(np.asarray(hop).flatten())
    return result


def block_to_csr(dst, seeds):
    """Row pointer of one block as a CSR by destination ROW POSITION (row r = seeds[r]).  The reference emits
    a block's edges seed by seed (base_sampler.py:34-40), so each seed's edges are one contiguous run of
    ``dst``; a seed listed twice owns two runs (consecutive duplicates are split by re-sampling counts, which
    the caller can pass instead).  Returns int64 row_ptr[len(seeds)+1]."""
    seeds = np.asarray(seeds, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    row_ptr = np.zeros(len(seeds) + 1, dtype=np.int64)
    pos = 0
    for r, v in enumerate(seeds):
        cnt = 0
        while pos + cnt < len(dst) and dst[pos + cnt] == v:
            cnt += 1
        row_ptr[r + 1] = row_ptr[r] + cnt
        pos += cnt
    assert pos == len(dst), "dst is not a seed-ordered edge list"
    return row_ptr


class CacheServer:
    """GraphCacheServer semantics, dgll/FeatureCache/storage.py:12-221, on numpy (DGL frames replaced by a dict of
    host tables).  cache_fix_data :129-148, fetch_data :151-198, fetch_from_cache :201-210, miss-rate :213-221,
    auto_cache policy :84-98 (full if capacity >= N else top-`capacity` nodes by out-degree, argsort descending)."""

    def __init__(self, host_tables, node_num, nid_map):
        self.host = host_tables
        self.node_num = node_num
        self.nid_map = np.asarray(nid_map, dtype=np.int64)
        self.gpu_flag = np.zeros(node_num, dtype=bool)
        self.localid2cacheid = np.zeros(node_num, dtype=np.int64)
        self.cache = {}
        self.full_cached = False
        self.try_num = 0
        self.miss_num = 0

    def cache_fix_data(self, nids, data, is_full=False):
        nids = np.asarray(nids, dtype=np.int64)
        self.localid2cacheid[nids] = np.arange(len(nids))
        for name in data:
            assert len(nids) == data[name].shape[0]
            self.cache[name] = np.array(data[name])
        self.gpu_flag[nids] = True
        self.full_cached = is_full

    def auto_cache(self, out_degrees, capability, names):
        if capability >= self.node_num:
            nids = np.arange(self.node_num)
            self.cache_fix_data(nids, {n: self.host[n][self.nid_map[nids]] for n in names}, is_full=True)
        else:
            order = np.argsort(-np.asarray(out_degrees), kind="stable")
            nids = order[:capability]
            self.cache_fix_data(nids, {n: self.host[n][self.nid_map[nids]] for n in names}, is_full=False)

    def fetch(self, tnid):
        tnid = np.asarray(tnid, dtype=np.int64)
        if self.full_cached:
            return {n: self.cache[n][tnid] for n in self.cache}
        mask = self.gpu_flag[tnid]
        frame = {}
        for name in self.cache:
            out = np.empty((len(tnid),) + self.cache[name].shape[1:], dtype=self.cache[name].dtype)
            out[mask] = self.cache[name][self.localid2cacheid[tnid[mask]]]
            out[~mask] = self.host[name][self.nid_map[tnid[~mask]]]
            frame[name] = out
        self.try_num += len(tnid)
        self.miss_num += int((~mask).sum())
        return frame

    def get_miss_rate(self):
        r = float(self.miss_num) / self.try_num
        self.miss_num = 0
        self.try_num = 0
        return r
