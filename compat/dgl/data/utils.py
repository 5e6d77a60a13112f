"""dgl.data.utils.{save_graphs, load_graphs} aliases."""


def save_graphs(filename, g_list, labels=None):
    from .. import save_graphs as _s
    return _s(filename, g_list, labels)


def load_graphs(filename, idx_list=None):
    from .. import load_graphs as _l
    return _l(filename, idx_list)
