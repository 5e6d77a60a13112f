"""dgl.nn.pytorch.{GraphConv, GCN2Conv} on librobotgnn (model/models.py:7 imports GraphConv)."""
import robot_gnn_b200 as _rg

GraphConv = _rg.GraphConv
GCN2Conv = _rg.GCN2Conv
