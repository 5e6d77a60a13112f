"""B200-native graph message-passing layers for multi-robot perception
(hot path of xjh19971/multi-robot-perception-gnn-1, dgl/model/models.py).

Import name: the directory is called ``multi-robot-perception-gnn-1_b200`` (not a Python identifier);
``import robot_gnn_b200`` (repo-root alias module) loads it.
"""
from . import _lib
from .graph import RobotGraph, graph, batch, robot_batch, function, relative_pose_edges, complete_k_of
from . import functional
from . import parallel
from .engine import HostPipeline, GraphedStep, GraphedForward, GraphedTrainSlice, gnn_slice_auto
from .functional import film_mean, film_mean_cat, copy_u_reduce, edge_mlp, compress, linear, axpby, to_gnn_layout, from_gnn_layout
from .layers import edge_encoder, GCN, GNNBlock, GraphConv, GCN2Conv, gnn_slice

__all__ = ["RobotGraph", "graph", "batch", "robot_batch", "function", "relative_pose_edges", "complete_k_of", "functional", "film_mean", "film_mean_cat",
           "copy_u_reduce", "edge_mlp", "compress", "linear", "axpby", "to_gnn_layout", "from_gnn_layout", "edge_encoder", "GCN", "GNNBlock",
           "GraphConv", "GCN2Conv", "HostPipeline", "GraphedStep", "GraphedForward", "GraphedTrainSlice", "gnn_slice_auto", "gnn_slice"]
