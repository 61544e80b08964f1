"""``GraphCacheServer`` — the reference's static GPU feature cache (dgll/FeatureCache/storage.py:12-221) on the
split-gather kernel.

Semantics kept: ``cache_fix_data`` fills ``localid2cacheid`` / ``gpu_flag`` / ``gpu_fix_cache`` (:129-148);
``auto_cache`` caches everything when the capacity allows, else the top-``capability`` nodes by out-degree,
``argsort(descending)`` (:64-98); ``fetch_data`` assembles per-layer frames from HBM hits and host misses (:151-198);
``fetch_from_cache`` (:201-210); ``get_miss_rate`` (:213-221).
Mechanism changed: the 5 index kernels + masked scatters + a synchronous CPU gather + H2D per field per layer of the
reference are ONE kernel launch per field (``dgllb_gather_rows_cached``): hits are read from the HBM cache, misses
straight from the pinned host table over PCIe/UVA, the miss counter is a device atomic.

``graph`` is the host feature store: a mapping ``{name: pinned or pageable host tensor}`` (pageable tensors are
pinned once), or an object exposing the reference's ``_node_frame._frame[name].data``.
"""
import torch

from .. import _nvtx
from .. import kernels as K


class NodeFlow:
    """Minimal stand-in for the DGL 0.4 NodeFlow surface ``fetch_data`` touches: per-layer parent node ids in,
    per-layer frames out (``_node_frames[i]`` = dict name -> tensor)."""

    def __init__(self, layer_nids):
        self.layer_nids = [torch.as_tensor(t) for t in layer_nids]
        self.num_layers = len(self.layer_nids)
        self._node_frames = [None] * self.num_layers

    def layer_parent_nid(self, i):
        return self.layer_nids[i]


class GraphCacheServer:
    def __init__(self, graph, node_num, nid_map, gpuid):
        self.graph = graph
        self.gpuid = gpuid
        self.device = torch.device("cuda", gpuid)
        self.node_num = node_num
        self.nid_map = nid_map.clone().detach().to(self.device).to(torch.int64)
        self.gpu_flag = torch.zeros(self.node_num, dtype=torch.bool, device=self.device)
        self.cached_num = 0
        self.capability = node_num
        self.full_cached = False
        self.dims = {}
        self.total_dim = 0
        self.gpu_fix_cache = dict()
        self.localid2cacheid = torch.zeros(node_num, dtype=torch.int64, device=self.device)
        self.log = False
        self.try_num = 0
        self.miss_num = 0
        self._miss_counter = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._host = {}

    # -- host store access --
    def _host_table(self, name):
        t = self._host.get(name)
        if t is None:
            if isinstance(self.graph, dict):
                t = self.graph[name]
            else:
                t = self.graph._node_frame._frame[name].data
            if not t.is_cuda and not t.is_pinned():
                t = t.contiguous().pin_memory()
            self._host[name] = t
        return t

    def init_field(self, embed_names):
        self.total_dim = 0
        for name in embed_names:
            t = self._host_table(name)
            self.dims[name] = t.size(1) if t.dim() > 1 else 1
            self.total_dim += self.dims[name]

    def get_feat_from_server(self, nids, embed_names, to_gpu=False):
        """storage.py:101-126 — rows ``nid_map[nids]`` of the host tables (host tensors unless ``to_gpu``)."""
        nids_in_full = self.nid_map[nids.to(self.device)]
        idx = nids_in_full.cpu()
        if to_gpu:     # storage.py:120-122: host-side index, then an asynchronous copy to the device
            return {name: self._host_table(name)[idx].to(self.device, non_blocking=True) for name in embed_names}
        return {name: self._host_table(name)[idx] for name in embed_names}

    def auto_cache(self, dgl_g, embed_names, capability=None):
        """storage.py:64-98.  ``dgl_g``: anything with ``out_degrees()`` or a degree tensor."""
        if not self.dims:
            self.init_field(embed_names)
        if capability is None:
            peak_allocated = torch.cuda.max_memory_allocated(device=self.gpuid)
            peak_cached = torch.cuda.max_memory_reserved(device=self.gpuid)
            total = torch.cuda.get_device_properties(self.gpuid).total_memory
            available = total - peak_allocated - peak_cached - 1024 * 1024 * 1024
            capability = int(available / (max(self.total_dim, 1) * 4))
        self.capability = capability
        if self.capability >= self.node_num:
            full_nids = torch.arange(self.node_num, device=self.device)
            self.cache_fix_data(full_nids, self._fetch_rows_for_cache(full_nids, embed_names), is_full=True)
        else:
            out_degrees = dgl_g.out_degrees() if hasattr(dgl_g, "out_degrees") else torch.as_tensor(dgl_g)
            sort_nid = torch.argsort(out_degrees.to(self.device), descending=True, stable=True)
            cache_nid = sort_nid[:self.capability]
            self.cache_fix_data(cache_nid, self._fetch_rows_for_cache(cache_nid, embed_names), is_full=False)

    def _fetch_rows_for_cache(self, nids, embed_names):
        idx = self.nid_map[nids].cpu()
        return {name: self._host_table(name)[idx] for name in embed_names}

    def cache_fix_data(self, nids, data, is_full=False):
        """storage.py:129-148."""
        nids = nids.to(self.device)
        rows = nids.size(0)
        self.localid2cacheid[nids] = torch.arange(rows, device=self.device)
        self.cached_num = rows
        for name in data:
            assert rows == data[name].size(0)
            self.dims[name] = data[name].size(1) if data[name].dim() > 1 else 1
            self.gpu_fix_cache[name] = data[name].to(self.device).contiguous()
        self.gpu_flag[nids] = True
        self.full_cached = is_full

    # -- the hot call --
    def fetch(self, tnid):
        """Frame ``{name: [len(tnid), dim]}`` for node ids ``tnid`` (hits from HBM, misses from the host table)."""
        tnid = tnid.to(self.device).to(torch.int64)
        frame = {}
        if self.full_cached:
            for name in self.gpu_fix_cache:
                frame[name] = K.gather_rows(self.gpu_fix_cache[name], tnid)
            return frame
        first = True
        for name in self.dims:
            cache = self.gpu_fix_cache.get(name)
            host = self._host_table(name)
            h2 = host if host.dim() > 1 else host.unsqueeze(1)
            c2 = None if cache is None else (cache if cache.dim() > 1 else cache.unsqueeze(1))
            # misses are counted once per row (on the first field), as log_miss_rate does per layer (storage.py:197-198)
            counter = self._miss_counter if (self.log and first) else None
            frame[name] = K.gather_rows_cached(c2, h2, tnid, self.gpu_flag, self.localid2cacheid, self.nid_map,
                                               miss_counter=counter)
            first = False
        if self.log:
            self.try_num += tnid.numel()
        return frame

    def fetch_data(self, nodeflow):
        """storage.py:151-198."""
        if self.full_cached:
            self.fetch_from_cache(nodeflow)
            return
        for i in range(nodeflow.num_layers):
            with _nvtx.range("cache-idxload"):
                tnid = nodeflow.layer_parent_nid(i).to(self.device)
            # the reference's cache-index / cache-allocate / cache-gpu / cache-cpu stages (:170-192) are ONE kernel here
            with _nvtx.range("cache-gpu+cache-cpu"):
                frame = self.fetch(tnid)
            with _nvtx.range("cache-asign"):
                nodeflow._node_frames[i] = frame

    def fetch_from_cache(self, nodeflow):
        """storage.py:201-210."""
        for i in range(nodeflow.num_layers):
            with _nvtx.range("cache-idxload"):
                tnid = nodeflow.layer_parent_nid(i).to(self.device)
            with _nvtx.range("cache-gpu"):
                nodeflow._node_frames[i] = {name: K.gather_rows(self.gpu_fix_cache[name], tnid)
                                            for name in self.gpu_fix_cache}

    def log_miss_rate(self, miss_num, total_num):
        self.try_num += total_num
        self.miss_num += miss_num

    def get_miss_rate(self):
        """storage.py:216-221 (the miss count lives in a device counter; reading it synchronises once)."""
        if self.log:
            self.miss_num += int(self._miss_counter.item())
            self._miss_counter.zero_()
        miss_rate = float(self.miss_num) / max(self.try_num, 1)
        self.miss_num = 0
        self.try_num = 0
        return miss_rate
