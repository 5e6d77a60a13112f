"""dgl.ops: imported by the reference (model/models.py:9) and never called."""
