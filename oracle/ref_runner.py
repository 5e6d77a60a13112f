"""Run the UNMODIFIED reference classes (byte-compiled into oracle/_ref by oracle/build_ref.py) on the CPU - the timed
`--impl reference` arm and `cpu_baseline` of bench.py (kind "reference"), and a second pin for the oracle (TEST INFRASTRUCTURE).

What is the reference's and what is restated here:
  * reference, unmodified (from oracle/_ref/model/models.bin): `multi_view_dgl_model.__init__` (it creates gcn1 / conv1 / gcn2 /
    conv2 exactly as the scripts do, models.py:158-171), `GCN`, `edge_encoder.forward` (models.py:152-155), `edge_udf`, `node_udf`
    (models.py:207-211);
  * restated: the DGL runtime underneath (oracle/dgl_shim - DGL is not installable here), and the two statements of
    `GCN.forward` around `update_all` (models.py:219-226) so that the layer returns ndata['images'] - the aggregate `node_udf`
    writes - instead of its input: with the literal return the backward would skip the aggregation entirely and the
    timed work would not be the layer's;
  * the torchvision encoder and the decoder are out of scope: their constructors are stubbed while the model object is
    built, and the timed region is the GNN slice of `multi_view_dgl_model.forward` (models.py:181-189) only.
"""
import importlib
import marshal
import os
import sys
import time
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(HERE, "dgl_shim")
_REF = os.path.join(HERE, "_ref")


def available():
    return all(os.path.exists(os.path.join(_REF, "model", f)) for f in ("models.bin", "blocks.bin"))


def _load_compiled(name):
    path = os.path.join(_REF, "model", name + ".bin")
    with open(path, "rb") as fh:
        ver = fh.readline().decode().strip()
        if ver != sys.version.split()[0]:
            raise RuntimeError(f"{path} was compiled by Python {ver}, this is {sys.version.split()[0]}: rebuild oracle/_ref")
        code = marshal.loads(fh.read())
    mod = types.ModuleType("model." + name)
    mod.__package__, mod.__file__ = "model", path
    sys.modules["model." + name] = mod
    exec(code, mod.__dict__)
    return mod


def load():
    """-> (dgl shim module, reference `model.models` module executed from its compiled code object)."""
    if not available():
        raise RuntimeError("oracle/_ref is not built (run __graft_entry__.build() where /root/reference is mounted)")
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    pkg = sys.modules.get("model")
    if pkg is not None and getattr(pkg, "__robot_gnn_ref__", False):
        return importlib.import_module("dgl"), sys.modules["model.models"]
    if pkg is not None:
        raise RuntimeError("another package named `model` is already imported (compat/model?); run the reference in its own process")
    dgl = importlib.import_module("dgl")
    pkg = types.ModuleType("model")
    pkg.__path__, pkg.__package__, pkg.__robot_gnn_ref__ = [os.path.join(_REF, "model")], "model", True
    sys.modules["model"] = pkg
    try:
        pkg.blocks = _load_compiled("blocks")          # models.py:5 does `from .blocks import TransBlock`
        pkg.models = _load_compiled("models")
    except Exception:
        for k in ("model", "model.blocks", "model.models"):
            sys.modules.pop(k, None)
        raise
    return dgl, pkg.models


def scripts_available():
    return all(os.path.exists(os.path.join(_REF, "scripts", f + ".bin")) for f in ("training", "dataloader", "dataloader_utils", "utils"))


def load_script_module(name, argv=None):
    """Execute one of the reference's top-level script modules (dgl/training.py and the modules it imports) from its compiled code
    object under its own name, so `import <name>` inside the others resolves to it.  The caller decides which `dgl` / `model`
    packages are importable (the compat ones, to run the scripts over the B200 path).  argv: sys.argv while the module body runs
    (training.py parses its options at import time, training.py:21-52)."""
    path = os.path.join(_REF, "scripts", name + ".bin")
    with open(path, "rb") as fh:
        ver = fh.readline().decode().strip()
        if ver != sys.version.split()[0]:
            raise RuntimeError(f"{path} was compiled by Python {ver}, this is {sys.version.split()[0]}: rebuild oracle/_ref")
        code = marshal.loads(fh.read())
    mod = types.ModuleType(name)
    mod.__file__ = path
    sys.modules[name] = mod
    old_argv = sys.argv
    if argv is not None:
        sys.argv = list(argv)
    try:
        exec(code, mod.__dict__)
    except Exception:
        sys.modules.pop(name, None)
        raise
    finally:
        sys.argv = old_argv
    return mod


def make_model(ref, f, k, compress_gcn, multi_gcn):
    opt = types.SimpleNamespace(feature_dim=f, compress_gcn=compress_gcn, multi_gcn=multi_gcn, camera_num=k, image_size=256,
                                task='depth', skip_level=False, backbone='resnet18', pretrained=False)
    enc, dec = ref.encoder, ref.decoder
    ref.encoder = lambda o: torch.nn.Identity()
    ref.decoder = lambda o, d: torch.nn.Identity()
    try:
        return ref.multi_view_dgl_model(opt)
    finally:
        ref.encoder, ref.decoder = enc, dec


def gnn_layer(ref, gcn, g, h):
    """GCN.forward (models.py:219-226) with the intended return key."""
    with g.local_scope():
        g.ndata['image'] = h
        with g.local_scope():
            g.edata['pose_gamma'], g.edata['pose_beta'] = gcn.edge_encoder(g.edata['pose'])
            g.update_all(ref.edge_udf, ref.node_udf)
            return g.ndata['images']


def gnn_slice(ref, model, g, h):
    """models.py:181-189."""
    g_h = gnn_layer(ref, model.gcn1, g, h)
    h = torch.cat((h, g_h), dim=1)
    if model.opt.compress_gcn:
        h = model.conv1(h)
    if model.opt.multi_gcn:
        g_h = gnn_layer(ref, model.gcn2, g, h)
        h = torch.cat((h, g_h), dim=1)
        h = model.conv2(h)
    return h


def time_block(f, k, graphs, compress_gcn, multi_gcn, reps=3, threads=None, seed=1):
    """fwd+bwd seconds (best of `reps` after one warm-up) of the reference's GNN slice on `graphs` K-robot graphs, fp32."""
    from . import gnn_oracle as O
    dgl, ref = load()
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(seed)
    sb = O.synthetic_batch(graphs, k, f, p=64, seed=seed)
    n = graphs * k
    model = make_model(ref, f, k, compress_gcn, multi_gcn)
    g = dgl.graph((torch.as_tensor(sb["src"]), torch.as_tensor(sb["dst"])), num_nodes=n)
    g.edata['pose'] = sb["pose"]
    h = sb["h"].reshape(n, f, 8, 8).clone().requires_grad_()
    dy = torch.randn((n, f if compress_gcn else 2 * f, 8, 8), generator=torch.Generator().manual_seed(2))
    best = float("inf")
    for it in range(reps + 1):
        model.zero_grad(set_to_none=True)
        h.grad = None
        t0 = time.perf_counter()
        y = gnn_slice(ref, model, g, h)
        y.backward(dy)
        dt = time.perf_counter() - t0
        if it > 0:
            best = min(best, dt)
    return best
