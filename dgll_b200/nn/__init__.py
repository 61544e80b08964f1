"""``dgll.nn`` for the B200 path: same exported names as dgll/nn/Convolution/__init__.py:7 and
dgll/nn/GlobalPooling/__init__.py:9, plus the DGL-named block layers the GPU-Accelerator scripts use."""
from .conv import (GAT, GCN, GIN, BinGCN, BinGCNConv, GinConv, GraphConvolution, GraphSage, NeighborAggregator, SpecialSpmm,
                   SpecialSpmmFunction, SpGAT, gatConv, gcnConv, sageConv, sparseGatConv)
from .pooling import Pooling, maxPooling, meanPooling, sumPooling
from .block_conv import GraphConv, SAGEConv
from .models import BlockGCN, GraphSAGE
from .ppi import GCNLayer, PPIGCN, create_sparse_adj

__all__ = ["gcnConv", "GraphConvolution", "GCN", "sageConv", "NeighborAggregator", "GraphSage", "gatConv",
           "sparseGatConv", "SpecialSpmm", "SpecialSpmmFunction", "GAT", "SpGAT", "GinConv", "GIN", "sumPooling",
           "meanPooling", "maxPooling", "Pooling", "GraphConv", "SAGEConv", "BlockGCN", "GraphSAGE", "GCNLayer", "PPIGCN", "create_sparse_adj", "BinGCNConv", "BinGCN"]
