"""Micro-benchmark of the fused dX + FiLM-backward kernel (rgnn_tc_conv_dx_film_bwd) against the unfused pair, bench shape.
RGNN_DXF_DEBUG knock-outs locate the time: 1 = epilogue only drains TMEM, 2 = no per-edge dots, 4 = no dh part, 8 = no MMAs."""
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import robot_gnn_b200 as R
    from robot_gnn_b200 import functional as Fn
    L = R._lib
    G, K, F = int(os.environ.get("G", 4096)), int(os.environ.get("K", 5)), int(os.environ.get("F", 512))
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(1)
    N, E = G * K, G * K * (K - 1)
    bf = lambda *s: torch.randn(s, generator=gen, device=dev).bfloat16()
    h, dy = torch.relu(bf(N, F, 64)), bf(N, F, 64)
    z = torch.sigmoid(bf(E, 2 * F).float()).bfloat16()
    wc = (torch.rand((F, 2 * F), generator=gen, device=dev) * 2 - 1) / (2 * F) ** 0.5
    wt = wc.t().contiguous().bfloat16()
    wtp = wt.index_select(0, Fn._dx_film_perm(F, dev))
    dh, dz = torch.empty_like(h), torch.empty_like(z)
    db2 = torch.zeros(2 * F, device=dev)

    def run():
        L.check(L.lib().rgnn_tc_conv_dx_film_bwd(L.ptr(wtp), L.ptr(dy), L.ptr(h), L.ptr(z), 1, L.ptr(dh), L.ptr(dz), L.ptr(db2), G, K, F, F, L.stream()), "dxf")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    chk = (float(dh.float().abs().sum()), float(dz.float().abs().sum()), float(db2.abs().sum()))
    print(f"checksums dh {chk[0]:.6e} dz {chk[1]:.6e} db2 {chk[2]:.6e}")
    print(f"RGNN_DXF_DEBUG={os.environ.get('RGNN_DXF_DEBUG', '0')}: {ms:.3f} ms per launch, {ms * 1e3 / (G * F // 64 / 148):.2f} us per tile per SM")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sweep":
        for d in ("0", "1", "8", "9", "16", "17", "25"):
            subprocess.run([sys.executable, __file__], env=dict(os.environ, RGNN_DXF_DEBUG=d))
    else:
        main()
