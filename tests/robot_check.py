"""Shared checker used by __graft_entry__.smoke() and the GPU tests: run the CUDA GNN block and the CPU oracle on the
same seeded inputs and compare outputs + every gradient (normalised max error)."""
import types

import torch


def make_block(R, O, f, comp, multi, seed=7):
    layers = [O.init_layer_params(f, compress_gcn=comp, seed=seed + j) for j in range(2 if multi else 1)]
    blk = R.GNNBlock(types.SimpleNamespace(feature_dim=f, compress_gcn=comp, multi_gcn=multi))
    sd = {}
    for j, l in enumerate(layers):
        pre = f"gcn{j + 1}.edge_encoder.layers."
        sd.update({pre + "0.weight": l["w1"], pre + "0.bias": l["b1"], pre + "2.weight": l["w2"], pre + "2.bias": l["b2"]})
        if comp:
            sd.update({f"conv{j + 1}.weight": l["wc"].reshape(f, 2 * f, 1, 1), f"conv{j + 1}.bias": l["bc"]})
    blk.load_state_dict(sd)
    return blk, layers


def nerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def check_block(R, O, g_, k, f, dtype, tol, comp=True, multi=True, seed=3):
    sb = O.synthetic_batch(g_, k, f, p=64, seed=seed)
    n = g_ * k
    blk, layers = make_block(R, O, f, comp, multi)
    dy = torch.randn((n, f if comp else 2 * f, 8, 8), generator=torch.Generator().manual_seed(11))
    h_o = sb["h"].reshape(n, f, 8, 8).double().requires_grad_()
    ls = [{kk: v.double().requires_grad_() for kk, v in l.items()} for l in layers]
    y_o = O.gnn_block(h_o, sb["pose"].double(), sb["src"], sb["dst"], ls, compress_gcn=comp)
    y_o.backward(dy.double())
    blk.cuda()
    g = R.robot_batch(g_, k, device="cuda")
    h = sb["h"].reshape(n, f, 8, 8).cuda().to(dtype).requires_grad_()
    y = blk(g, h, sb["pose"].cuda())
    y.backward(dy.cuda().to(dtype))
    errs = {"y": nerr(y.float(), y_o), "dh": nerr(h.grad.float(), h_o.grad)}
    for j, l in enumerate(ls):
        gcn = getattr(blk, f"gcn{j + 1}")
        errs[f"l{j}.dw2"] = nerr(gcn.edge_encoder.layers[2].weight.grad, l["w2"].grad)
        errs[f"l{j}.dw1"] = nerr(gcn.edge_encoder.layers[0].weight.grad, l["w1"].grad)
        if comp:
            errs[f"l{j}.dwc"] = nerr(getattr(blk, f"conv{j + 1}").weight.grad.reshape(f, 2 * f), l["wc"].grad)
    bad = {kk: v for kk, v in errs.items() if not v <= (tol if kk in ("y", "dh") else 2 * tol)}
    if bad:
        raise AssertionError(f"CUDA GNN block differs from the oracle ({dtype}): {bad}")
    return errs
