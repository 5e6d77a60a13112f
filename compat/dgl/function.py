"""dgl.function builtins (model/models.py:8): message/reduce descriptors understood by RobotGraph.update_all."""
import robot_gnn_b200 as _rg

copy_u = _rg.function.copy_u
copy_src = _rg.function.copy_src
sum = _rg.function.sum      # noqa: A001
mean = _rg.function.mean
