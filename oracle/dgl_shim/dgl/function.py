"""dgl.function builtins restated as UDFs (ORACLE / TEST INFRASTRUCTURE).
Used by the reference at dgl/model/dgl_models.py:126 (fn.copy_u)."""


def copy_u(u, out):
    return lambda edges: {out: edges.src[u]}


copy_src = copy_u


def sum(msg, out):  # noqa: A001
    return lambda nodes: {out: nodes.mailbox[msg].sum(1)}


def mean(msg, out):
    return lambda nodes: {out: nodes.mailbox[msg].mean(1)}
