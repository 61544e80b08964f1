"""Data path with the reference's names: graph store, samplers, mini-batch iterators, feature cache, blocks."""
from .dgraph import DGraph
from .sampling import (Base_sampler, BlockDataLoader, DataLoader, DGLLNeighborSampler, NeighborSampler, create_block,
                       multihop_sampling, sampling, sugbraph)
from .cache import GraphCacheServer, NodeFlow

__all__ = ["DGraph", "Base_sampler", "DGLLNeighborSampler", "DataLoader", "sugbraph", "sampling", "multihop_sampling",
           "create_block", "NeighborSampler", "BlockDataLoader", "GraphCacheServer", "NodeFlow"]
