"""gkgnet_b200 -- B200-native graph hot path of GKGNet (dynamic group-kNN graph +
max-relative aggregation) behind the reference's module API.  CUDA only: importing works
anywhere, running a kernel requires libgkg_b200.so and an sm_100a device."""
from . import _lib, ops  # noqa: F401
from .backbone import FFN, GKGNet, Downsample, Stem  # noqa: F401
from .graph import DenseDilated, DenseDilatedKnnGraph, edge_index_from_neighbors  # noqa: F401
from .head import LabelQueryHead  # noqa: F401
from .layers import BasicConv, act_layer, batched_index_select, norm_layer, set_norm_type  # noqa: F401
from .registry import BACKBONES, HEADS, build_backbone, build_head  # noqa: F401
from .vertex import (DyGraphConv2d, DyGraphConv2dMultiGroup, DyGraphLabel,  # noqa: F401
                     DyGraphLabelMultiGroup, EdgeConv2d, FFNLabel, GINConv2d, GraphAtten, GraphConv2d, Grapher,
                     GrapherLabel, GraphSAGE, MRConv2d)
