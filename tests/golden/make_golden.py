"""Generate the frozen golden vectors under tests/golden/ (run in the BUILD container only).

    python tests/golden/make_golden.py

Runs the UNMODIFIED reference classes from /root/reference/dgl/model/models.py
(``edge_encoder``, ``GCN``, ``edge_udf``, ``node_udf``, ``multi_view_dgl_model``)
over oracle/dgl_shim (DGL itself is not installable here) and stores inputs +
outputs as small .npz fixtures.  /root/reference does not exist on the GPU box;
tests only read the .npz files.

Import hazard: the reference directory is itself called ``dgl``;
/root/reference/dgl (NOT /root/reference) goes on sys.path so ``import dgl``
resolves to the shim and ``import model.models`` to the reference.
"""
import argparse
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "dgl_shim"))
sys.path.insert(0, "/root/reference/dgl")
sys.path.insert(0, ROOT)

import dgl  # the shim  # noqa: E402
from model import models as ref  # the unmodified reference  # noqa: E402
from oracle import gnn_oracle as O  # noqa: E402

torch.set_num_threads(1)  # deterministic accumulation order
torch.use_deterministic_algorithms(True)


def ref_graph(batch):
    g = dgl.graph((batch["src"], batch["dst"]), num_nodes=int(batch["node_offsets"][-1]))
    return g


def load_layer(mod_gcn, conv, p):
    sd = {"edge_encoder.layers.0.weight": p["w1"], "edge_encoder.layers.0.bias": p["b1"],
          "edge_encoder.layers.2.weight": p["w2"], "edge_encoder.layers.2.bias": p["b2"]}
    mod_gcn.load_state_dict(sd, strict=True)
    if conv is not None:
        conv.load_state_dict({"weight": p["wc"].reshape(p["wc"].shape[0], -1, 1, 1), "bias": p["bc"]}, strict=True)


def run_ref_layer(gcn, g, h, pose):
    """Body of the reference GCN.forward (models.py:219-223) reading the key node_udf
    actually writes ('images', models.py:208) - the intended output (dgl_models.py:189)."""
    with g.local_scope():
        g.ndata['image'] = h
        g.edata['pose'] = pose
        with g.local_scope():
            g.edata['pose_gamma'], g.edata['pose_beta'] = gcn.edge_encoder(g.edata['pose'])
            g.update_all(ref.edge_udf, ref.node_udf)
            return g.ndata['images']


def npz(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **out)
    print(f"wrote {name}: {os.path.getsize(path)/1024:.1f} KiB")


# ----------------------------------------------------------------------------------------------
def golden_csr():
    cases = []
    for k in (2, 3, 5, 8):
        for g in (1, 3, 8):
            s, d = O.fully_connected_edges(k)
            cases.append(([s] * g, [d] * g, [k] * g))
    ks = [2, 5, 3, 8, 4, 1]   # heterogeneous incl. a single isolated robot (no edges)
    cases.append(([O.fully_connected_edges(k)[0] for k in ks], [O.fully_connected_edges(k)[1] for k in ks], ks))
    rng = np.random.RandomState(0)
    for _ in range(4):  # random multigraphs with isolated nodes and duplicate edges
        nn_ = rng.randint(1, 12, size=rng.randint(1, 6)).tolist()
        ss, dd = [], []
        for n in nn_:
            e = rng.randint(0, 3 * n + 1)
            ss.append(rng.randint(0, n, size=e).astype(np.int64))
            dd.append(rng.randint(0, n, size=e).astype(np.int64))
        cases.append((ss, dd, nn_))
    out = {"num_cases": len(cases)}
    for i, (ss, dd, nn_) in enumerate(cases):
        b = O.batch_graphs(ss, dd, nn_)
        c = O.build_csr(b["src"], b["dst"], b["node_offsets"][-1])
        # cross-check against the DGL-semantics shim: batch() ids and mailbox (ascending eid) order
        gs = [dgl.graph((torch.from_numpy(s), torch.from_numpy(d)), num_nodes=n) for s, d, n in zip(ss, dd, nn_)]
        bg = dgl.batch(gs)
        assert np.array_equal(bg.edges()[0].numpy(), b["src"]) and np.array_equal(bg.edges()[1].numpy(), b["dst"])
        assert np.array_equal(bg.in_degrees().numpy(), c["in_deg"]) and np.array_equal(bg.out_degrees().numpy(), c["out_deg"])
        assert np.array_equal(bg.batch_num_nodes().numpy(), b["batch_num_nodes"])
        if bg.num_edges():
            bg.edata['id'] = torch.arange(bg.num_edges(), dtype=torch.float64).unsqueeze(1)
            seen = {}

            def red(nodes, seen=seen):
                seen.setdefault("mb", []).append(nodes.mailbox['m'])
                return {'x': nodes.mailbox['m'].sum(1)}
            bg.ndata['z'] = torch.zeros(bg.num_nodes(), 1)
            bg.update_all(lambda e: {'m': e.data['id']}, red)
            for mb in seen["mb"]:
                assert (mb[:, 1:, 0] > mb[:, :-1, 0]).all()  # ascending edge id per mailbox row
                for row in mb[:, :, 0].long().numpy():
                    v = b["dst"][row[0]]
                    assert np.array_equal(row, c["in_eid"][c["in_ptr"][v]:c["in_ptr"][v + 1]])
        out[f"c{i}_local_src"] = np.concatenate(ss) if len(ss) else np.zeros(0, np.int64)
        out[f"c{i}_local_dst"] = np.concatenate(dd) if len(dd) else np.zeros(0, np.int64)
        for k2, v in {**b, **c}.items():
            out[f"c{i}_{k2}"] = v
    npz("csr_cases.npz", **out)


# ----------------------------------------------------------------------------------------------
def golden_layers():
    """GCN layer forward + every gradient, from the reference classes (fp32, as the reference runs)."""
    cfgs = [  # (seed, F, (ph, pw), K list)
        (1, 16, (1, 1), [5, 5, 5]),
        (2, 32, (8, 8), [5, 5]),
        (3, 64, (8, 8), [3, 5, 2, 4]),
        (1, 24, (2, 3), [4, 1, 3]),      # ragged P, isolated robot
    ]
    out = {"num_cases": len(cfgs)}
    for i, (seed, f, (ph, pw), ks) in enumerate(cfgs):
        gen = torch.Generator().manual_seed(seed)
        ed = [O.fully_connected_edges(k) for k in ks]
        b = O.batch_graphs([e[0] for e in ed], [e[1] for e in ed], ks)
        n, e = int(b["node_offsets"][-1]), int(b["edge_offsets"][-1])
        h = torch.relu(torch.randn((n, f, ph, pw), generator=gen)).requires_grad_()
        pose = torch.randn((e, 9), generator=gen) * torch.tensor([3., 3, 3, 1, 1, 1, 1, 1, 1])
        dagg = torch.randn((n, f, ph, pw), generator=gen)
        p = O.init_layer_params(f, compress_gcn=False, seed=seed)
        opt = types.SimpleNamespace(feature_dim=f)
        gcn = ref.GCN(opt)
        load_layer(gcn, None, p)
        g = ref_graph(b)
        agg = run_ref_layer(gcn, g, h, pose)
        gam, bet = gcn.edge_encoder(pose)
        z = torch.stack((gam.reshape(e, f), bet.reshape(e, f)), 2).reshape(e, 2 * f)
        agg.backward(dagg)
        # literal reference behaviour (F5): forward returns its input
        with g.local_scope():
            g.ndata['image'] = h.detach()
            g.edata['pose'] = pose
            lit = gcn(g)
        assert torch.equal(lit, h.detach())
        # closed form must agree with the reference exactly (same op order)
        agg_o = O.film_mean(h.detach(), O.edge_mlp(pose, p["w1"], p["b1"], p["w2"], p["b2"]), b["src"], b["dst"])
        print(f"  layer case {i}: max|ref-oracle| = {(agg - agg_o).abs().max().item():.3e}")
        sd = gcn.state_dict()
        npz_case = dict(ks=np.asarray(ks), src=b["src"], dst=b["dst"], h=h, pose=pose, dagg=dagg,
                        w1=p["w1"], b1=p["b1"], w2=p["w2"], b2=p["b2"], z=z, agg=agg, dh=h.grad,
                        dw1=gcn.edge_encoder.layers[0].weight.grad, db1=gcn.edge_encoder.layers[0].bias.grad,
                        dw2=gcn.edge_encoder.layers[2].weight.grad, db2=gcn.edge_encoder.layers[2].bias.grad)
        for k2, v in npz_case.items():
            out[f"c{i}_{k2}"] = v.detach().numpy() if isinstance(v, torch.Tensor) else v
        out[f"c{i}_state_keys"] = np.asarray(sorted(sd.keys()))
    npz("layer_cases.npz", **out)


# ----------------------------------------------------------------------------------------------
class _FakeEncoder(torch.nn.Module):
    """Stands in for the torchvision backbone (out of scope): returns the preset feature map."""
    def __init__(self, h):
        super().__init__()
        self.h = h

    def forward(self, image):
        return [self.h]


def golden_blocks():
    """GNN slice of multi_view_dgl_model.forward (models.py:180-189): compress / multi_gcn variants,
    intended ('images') and literal ('image', unmodified forward end-to-end) behaviour."""
    cfgs = [  # (seed, F, K, G, compress, multi)
        (1, 32, 5, 3, False, False),
        (2, 32, 5, 3, True, False),
        (3, 32, 5, 2, True, True),
        (1, 48, 3, 4, True, True),
    ]
    out = {"num_cases": len(cfgs)}
    for i, (seed, f, k, gN, comp, multi) in enumerate(cfgs):
        opt = types.SimpleNamespace(feature_dim=f, compress_gcn=comp, multi_gcn=multi, camera_num=k, image_size=32,
                                    task='depth', skip_level=False, backbone='resnet18', pretrained=False)
        # the decoder/encoder ctor are out of scope; build the class without running them
        orig_enc, orig_dec = ref.encoder, ref.decoder
        ref.encoder = lambda o: torch.nn.Identity()
        ref.decoder = lambda o, d: torch.nn.Identity()
        try:
            model = ref.multi_view_dgl_model(opt)
        finally:
            ref.encoder, ref.decoder = orig_enc, orig_dec
        gnn_keys = sorted(k2 for k2 in model.state_dict().keys())
        sb = O.synthetic_batch(gN, k, f, p=64, seed=seed)
        h0 = sb["h"].reshape(gN * k, f, 8, 8).clone().requires_grad_()
        pose, dy = sb["pose"], None
        layers = [O.init_layer_params(f, compress_gcn=comp, seed=seed * 10 + j) for j in range(2 if multi else 1)]
        load_layer(model.gcn1, model.conv1 if comp else None, layers[0])
        if multi:
            load_layer(model.gcn2, model.conv2, layers[1])
        g = ref_graph(sb)
        g.edata['pose'] = pose
        # -- intended: the forward's slice with GCN reading 'images'
        h = h0
        for j in range(len(layers)):
            gcn, conv = (model.gcn1, getattr(model, 'conv1', None)) if j == 0 else (model.gcn2, model.conv2)
            g_h = run_ref_layer(gcn, g, h, pose)
            h = torch.cat((h, g_h), dim=1)
            if comp:
                h = conv(h)
        y = h
        gen = torch.Generator().manual_seed(seed + 100)
        dy = torch.randn(y.shape, generator=gen)
        y.backward(dy)
        grads = {n_: p_.grad.clone() for n_, p_ in model.named_parameters() if p_.grad is not None}
        # -- literal: UNMODIFIED forward end to end, encoder/decoder stubbed
        model.encoder = _FakeEncoder(h0.detach())
        model.decoder = torch.nn.Identity()
        g.ndata['image'] = torch.zeros(gN * k, 3, 32, 32)
        with torch.no_grad():
            y_lit = model(g)
        rec = dict(cfg=np.asarray([seed, f, k, gN, int(comp), int(multi)]), src=sb["src"], dst=sb["dst"],
                   h=h0, pose=pose, dy=dy, y=y, y_literal=y_lit, dh=h0.grad,
                   state_keys=np.asarray(gnn_keys))
        for j, lyr in enumerate(layers):
            for k2, v in lyr.items():
                rec[f"l{j}_{k2}"] = v
        for n_, v in grads.items():
            rec["grad/" + n_] = v
        for k2, v in rec.items():
            out[f"c{i}_{k2}"] = v.detach().numpy() if isinstance(v, torch.Tensor) else v
    npz("block_cases.npz", **out)


# ----------------------------------------------------------------------------------------------
def golden_script_shape():
    """Script shapes (G=8, K=5, P=8x8; training.py:28, -camera_idx 01234) at F=512 / 1280.
    Inputs are regenerated from the seed by oracle.synthetic_batch; only checksums of the inputs and a
    strided sample of the reference outputs are stored (fixtures stay small)."""
    out = {}
    for f in (512, 1280):
        sb = O.synthetic_batch(8, 5, f, p=64, seed=1)
        p = O.init_layer_params(f, compress_gcn=True, seed=1)
        opt = types.SimpleNamespace(feature_dim=f)
        gcn = ref.GCN(opt)
        conv = torch.nn.Conv2d(2 * f, f, kernel_size=1)   # models.py:164
        load_layer(gcn, conv, p)
        g = ref_graph(sb)
        h = sb["h"].reshape(40, f, 8, 8).clone().requires_grad_()
        g_h = run_ref_layer(gcn, g, h, sb["pose"])
        y = conv(torch.cat((h, g_h), dim=1))
        y.backward(sb["dy"].reshape(y.shape))
        idx = np.arange(0, 40 * f * 64, 997)
        out[f"f{f}_idx"] = idx
        out[f"f{f}_h_sum"] = np.float64(sb["h"].double().sum().item())
        out[f"f{f}_pose_sum"] = np.float64(sb["pose"].double().sum().item())
        out[f"f{f}_agg_s"] = g_h.detach().reshape(-1).numpy()[idx]
        out[f"f{f}_y_s"] = y.detach().reshape(-1).numpy()[idx]
        out[f"f{f}_dh_s"] = h.grad.reshape(-1).numpy()[idx]
        out[f"f{f}_y_absmax"] = np.float32(y.detach().abs().max().item())
        out[f"f{f}_dh_absmax"] = np.float32(h.grad.abs().max().item())
        out[f"f{f}_dwc_s"] = conv.weight.grad.reshape(-1).numpy()[::1009]
        out[f"f{f}_dw2_s"] = gcn.edge_encoder.layers[2].weight.grad.reshape(-1).numpy()[::1009]
        out[f"f{f}_dw1"] = gcn.edge_encoder.layers[0].weight.grad.numpy()
        out[f"f{f}_db2"] = gcn.edge_encoder.layers[2].bias.grad.numpy()
        out[f"f{f}_dbc"] = conv.bias.grad.numpy()
    npz("script_shape_cases.npz", **out)


def golden_script_c1():
    """C1 as the script runs it (train_camera_warehouse_gcn.sh: mobilenetv2 F=1280, one layer, NO compress conv, fp32):
    G=8, K=5, P=8x8.  The block output is cat((h, g_h), 1) (models.py:182); strided samples + checksums only."""
    f = 1280
    sb = O.synthetic_batch(8, 5, f, p=64, seed=1)
    p = O.init_layer_params(f, compress_gcn=False, seed=1)
    gcn = ref.GCN(types.SimpleNamespace(feature_dim=f))
    load_layer(gcn, None, p)
    g = ref_graph(sb)
    h = sb["h"].reshape(40, f, 8, 8).clone().requires_grad_()
    g_h = run_ref_layer(gcn, g, h, sb["pose"])
    y = torch.cat((h, g_h), dim=1)                                    # models.py:182, compress_gcn off
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(101))
    y.backward(dy)
    idx = np.arange(0, 40 * 2 * f * 64, 997)
    idx_h = np.arange(0, 40 * f * 64, 997)
    l0, l2 = gcn.edge_encoder.layers[0], gcn.edge_encoder.layers[2]
    npz("script_c1_nocompress.npz", idx=idx, idx_h=idx_h, h_sum=np.float64(sb["h"].double().sum().item()),
        dy_sum=np.float64(dy.double().sum().item()),
        y_s=y.detach().reshape(-1).numpy()[idx], y_absmax=np.float32(y.detach().abs().max().item()),
        dh_s=h.grad.reshape(-1).numpy()[idx_h], dh_absmax=np.float32(h.grad.abs().max().item()),
        dw2_s=l2.weight.grad.reshape(-1).numpy()[::1009], dw2_absmax=np.float32(l2.weight.grad.abs().max().item()),
        dw1=l0.weight.grad.numpy(), db1=l0.bias.grad.numpy(), db2=l2.bias.grad.numpy())


def golden_pose():
    """dgl/utils.py:80-85 self-check inputs; expected values printed by the reference."""
    sys.path.insert(0, "/root/reference/dgl")
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_utils", "/root/reference/dgl/utils.py")
    ru = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ru)
    p1 = np.array([-5.902101516723632812e+01, 1.927141571044921875e+02, 5.996415138244628906e+00, -2.683681845664978027e-01,
                   2.737782299518585205e-01, 6.468879580497741699e-01, 6.592115163803100586e-01])
    p2 = np.array([-5.506656265258789062e+01, 1.961531982421875000e+02, 5.998308658599853516e+00, -4.198011383414268494e-02,
                   3.792373836040496826e-01, 1.026619002223014832e-01, 9.186278581619262695e-01])
    rng = np.random.RandomState(7)
    poses = rng.randn(5, 7)
    poses[:, 3:] /= np.linalg.norm(poses[:, 3:], axis=1, keepdims=True)
    edges = np.stack([ru.cal_relative_pose(poses[i], poses[j]) for i in range(5) for j in range(5) if i != j])
    npz("pose_cases.npz", p1=p1, p2=p2, rel12=ru.cal_relative_pose(p1, p2), rel11=ru.cal_relative_pose(p1, p1),
        poses=poses, edges=edges)


def golden_model_keys():
    """state_dict keys + shapes of the reference multi_view_dgl_model (pins the script-level drop-in, compat/model/models.py)."""
    import json
    out = {}
    for name, kw in {"resnet18_compress_multi": dict(backbone="resnet18", feature_dim=512, compress_gcn=True, multi_gcn=True, skip_level=False),
                     "mobilenetv2_plain_skip": dict(backbone="mobilenetv2", feature_dim=1280, compress_gcn=False, multi_gcn=False, skip_level=True)}.items():
        opt = types.SimpleNamespace(pretrained=False, camera_num=5, image_size=64, task="depth", output_dim=1, **kw)
        out[name] = {k: list(v.shape) for k, v in ref.multi_view_dgl_model(opt).state_dict().items()}
    json.dump(out, open(os.path.join(HERE, "model_state_keys.json"), "w"))
    print("wrote model_state_keys.json")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    todo = dict(csr=golden_csr, layers=golden_layers, blocks=golden_blocks, script=golden_script_shape, c1=golden_script_c1, pose=golden_pose, keys=golden_model_keys)
    for k, fn_ in todo.items():
        if a.only in (None, k):
            print("==", k)
            fn_()
