"""dgl.ops placeholder: imported (dgl/model/models.py:9) but never called by the reference."""
