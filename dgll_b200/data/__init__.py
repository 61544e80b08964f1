"""Data path with the reference's names: graph store, samplers, mini-batch iterators, feature cache, blocks."""
from .dgraph import DGraph
from .sampling import (Base_sampler, BlockDataLoader, DataLoader, DGLLNeighborSampler, NeighborSampler, create_block,
                       multihop_sampling, sampling, sugbraph)
from .cache import GraphCacheServer, NodeFlow
from .layerwise import (FastGCNSampler, FastGCNSamplerFlat, FastGCNSamplerFlatWrs, FastGCNSamplerWrs, Ladies, LadiesFlat,
                        LadiesFlatWrs, LadiesWrs, LayerwiseSampler, build_laplacian, numpy_chooser)

__all__ = ["DGraph", "Base_sampler", "DGLLNeighborSampler", "DataLoader", "sugbraph", "sampling", "multihop_sampling",
           "create_block", "NeighborSampler", "BlockDataLoader", "GraphCacheServer", "NodeFlow",
           "Ladies", "LadiesFlat", "LadiesWrs", "LadiesFlatWrs", "FastGCNSampler", "FastGCNSamplerFlat",
           "FastGCNSamplerWrs", "FastGCNSamplerFlatWrs", "LayerwiseSampler", "build_laplacian", "numpy_chooser"]
