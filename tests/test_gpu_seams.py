"""GPU: SURVEY 8(f) row 2 - the encoder -> GNN and GNN -> decoder seams (dgl/model/models.py:177-180,190-205; blocks.py:13-25):
channels-last / other-dtype feature maps cross into and out of the robot-graph kernels' NCHW layout in one fused pass each."""
import types

import pytest
import torch

from conftest import nerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,f,hw", [(10, 512, (8, 8)), (3, 48, (7, 7)), (5, 130, (4, 9)), (1, 64, (8, 8))])
@pytest.mark.parametrize("din,dout", [(torch.float32, torch.bfloat16), (torch.bfloat16, torch.float32), (torch.float32, torch.float32),
                                      (torch.bfloat16, torch.bfloat16)])
def test_layout_cast_is_exactly_permute_plus_cast(R, n, f, hw, din, dout):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((n, f) + hw, generator=g, device="cuda").to(din)
    xcl = x.contiguous(memory_format=torch.channels_last)
    y = R.to_gnn_layout(xcl, dout)                                   # channels-last -> NCHW, cast fused
    assert y.is_contiguous() and y.dtype == dout and torch.equal(y, x.to(dout))
    z = R.from_gnn_layout(x, dout, channels_last=True)               # NCHW -> channels-last, cast fused
    assert z.is_contiguous(memory_format=torch.channels_last) and z.shape == x.shape and torch.equal(z, x.to(dout))
    # gradients travel back through the mirror-image kernel
    xr = xcl.clone().requires_grad_()
    w = torch.randn_like(y)
    (R.to_gnn_layout(xr, dout).float() * w.float()).sum().backward()
    assert xr.grad.dtype == din and nerr(xr.grad.float(), w.float()) <= (1e-6 if din == torch.float32 else 8e-3)


def test_channels_last_fp32_backbone_feeds_the_bf16_block_and_a_channels_last_decoder(R):
    """An fp32 channels-last encoder output goes through the bf16 throughput mode and comes back fp32 channels-last: same numbers as
    the NCHW bf16 path, gradients included, and no torch transpose / cast copy on the way."""
    from robot_check import make_block
    from oracle import gnn_oracle as O
    G, K, F = 6, 5, 256
    blk, _ = make_block(R, O, F, True, True)
    blk.cuda()
    g = R.robot_batch(G, K, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(3)
    h32 = torch.rand((G * K, F, 8, 8), generator=gen, device="cuda")
    pose = torch.randn((G * K * (K - 1), 9), generator=gen, device="cuda")
    dy = torch.randn((G * K, F, 8, 8), generator=gen, device="cuda")
    # reference: explicit NCHW bf16 run
    hb = h32.bfloat16().requires_grad_()
    yb = R.gnn_slice(blk, g, hb, pose)
    yb.backward(dy.bfloat16())
    ref_grads = {n: p.grad.clone() for n, p in blk.named_parameters()}
    for p in blk.parameters():
        p.grad = None
    # seam run: channels-last fp32 in, bf16 compute, channels-last fp32 out
    hcl = h32.contiguous(memory_format=torch.channels_last).requires_grad_()
    y = R.gnn_slice(blk, g, hcl, pose, compute_dtype=torch.bfloat16)
    assert y.dtype == torch.float32 and y.is_contiguous(memory_format=torch.channels_last) and y.shape == h32.shape
    assert torch.equal(y, yb.detach().float())
    y.backward(dy.contiguous(memory_format=torch.channels_last))
    assert hcl.grad.shape == h32.shape and nerr(hcl.grad, hb.grad.float()) <= 1e-6
    for n, p in blk.named_parameters():
        assert nerr(p.grad, ref_grads[n]) <= 1e-5, n
