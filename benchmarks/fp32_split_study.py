"""Error study for putting the fp32 parity mode on the tensor cores (SURVEY §7 'fp32 parity vs tensor cores').

Question: can Linear(F,2F) (models.py:149-150) and Conv2d(2F,F,1) (models.py:164) of the fp32 mode run as split-bf16 products on
tcgen05 (fp32 accumulation in TMEM) and stay inside the 1e-5 normalised-max-error bound at the longest contraction (K = 2F = 2560)?

x (fp32, 24-bit significand) = x0 + x1 + x2 with three bf16 terms (8 bits each, round-to-nearest residuals: exact up to the last bit).
A.B ~ sum over plane pairs (i, j), i + j <= 2  (6 products; the three dropped ones are ~2^-24 relative).  The planes are concatenated
along K, so ONE bf16 GEMM with K' = 6 K computes it.  What can go wrong is the tensor core's fp32 accumulator (truncation when an
addend is aligned to a large running sum), so the study varies (a) the ORDER of the plane pairs (small terms first keeps the running
sum small while the corrections land) and (b) the accumulation length per accumulator (K-chunks summed afterwards in fp32 RN).

This script uses cuBLAS bf16 GEMMs with fp32 output (torch.mm(..., out_dtype=float32)) as the stand-in for "tcgen05, fp32 TMEM
accumulator"; it is a study tool, not part of the product path.  Output: one JSON line per (case, scheme).
"""
import json
import sys

import torch


def split3(x):
    x0 = x.bfloat16()
    r = x - x0.float()
    x1 = r.bfloat16()
    x2 = (r - x1.float()).bfloat16()
    return [x0, x1, x2]


def split_tf32(x):
    hi = (x.view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32)      # RN to 10 explicit mantissa bits
    return hi, x - hi


def mm32(a, b):
    return torch.mm(a, b.t(), out_dtype=torch.float32)


def nerr(c, ref):
    return float((c.double() - ref).abs().max() / ref.abs().max())


def run_case(name, a, b, out):
    """a [M, K], b [N, K] fp32 on cuda; C = a . b^T."""
    ref = a.double() @ b.double().t()
    K = a.shape[1]
    res = {}
    torch.backends.cuda.matmul.allow_tf32 = False
    res["fp32_cublas_simt"] = nerr(a @ b.t(), ref)
    torch.backends.cuda.matmul.allow_tf32 = True
    res["tf32_1pass"] = nerr(a @ b.t(), ref)
    ah, al = split_tf32(a)
    bh, bl = split_tf32(b)
    res["tf32_3pass_small_first"] = nerr(((al @ bh.t()) + (ah @ bl.t())) + (ah @ bh.t()), ref)
    torch.backends.cuda.matmul.allow_tf32 = False
    ap, bp = split3(a), split3(b)
    res["bf16_1pass"] = nerr(mm32(ap[0], bp[0]), ref)
    pairs_big_first = [(0, 0), (0, 1), (1, 0), (0, 2), (1, 1), (2, 0)]
    pairs_small_first = pairs_big_first[::-1]
    pairs3 = [(1, 0), (0, 1), (0, 0)]
    for tag, pairs in (("bf16x3_6prod_big_first", pairs_big_first), ("bf16x3_6prod_small_first", pairs_small_first),
                       ("bf16x2_3prod_small_first", pairs3)):
        A = torch.cat([ap[i] for i, _ in pairs], 1)
        B = torch.cat([bp[j] for _, j in pairs], 1)
        res[tag + "_one_acc"] = nerr(mm32(A, B), ref)
    # chunked: each accumulator sees at most `ck` elements of the original K per plane pair; partial sums added in fp32 (RN)
    for ck in (128, 256, 512):
        acc = torch.zeros_like(ref, dtype=torch.float32)
        for k0 in range(0, K, ck):
            A = torch.cat([ap[i][:, k0:k0 + ck] for i, _ in pairs_small_first], 1)
            B = torch.cat([bp[j][:, k0:k0 + ck] for _, j in pairs_small_first], 1)
            acc += mm32(A, B)
        res[f"bf16x3_6prod_small_first_chunk{ck}"] = nerr(acc, ref)
    # corrections in one accumulator, the main product chunked
    for ck in (256,):
        A = torch.cat([ap[i] for i, _ in pairs_small_first[:-1]], 1)
        B = torch.cat([bp[j] for _, j in pairs_small_first[:-1]], 1)
        acc = mm32(A, B)
        for k0 in range(0, K, ck):
            acc += mm32(ap[0][:, k0:k0 + ck], bp[0][:, k0:k0 + ck])
        res[f"bf16x3_corr_one_acc_main_chunk{ck}"] = nerr(acc, ref)
    for k, v in res.items():
        line = dict(case=name, M=a.shape[0], N=b.shape[0], K=K, scheme=k, nerr=v, within_1e5=bool(v <= 1e-5))
        print(json.dumps(line))
        out.append(line)


def main():
    assert torch.cuda.is_available()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(1)
    out = []
    for F in (512, 1280):
        K2 = 2 * F
        # compress conv (models.py:164): Wc [F, 2F] ~ U(+-1/sqrt(2F)); X = cat(h, g_h) >= 0 (post-ReLU features, FiLM mean of them)
        wc = (torch.rand((F, K2), generator=g, device=dev) * 2 - 1) / K2 ** 0.5
        x = torch.relu(torch.randn((4096, K2), generator=g, device=dev))
        run_case(f"conv_fwd_F{F}", wc, x, out)
        # its dX: K = F (out channels), operands W^T and dY ~ N(0,1)
        dy = torch.randn((4096, F), generator=g, device=dev)
        run_case(f"conv_dx_F{F}", wc.t().contiguous(), dy, out)
        # dW: contraction over nodes x pixels (long K): K = 64 * 512 robots
        if F == 512:
            run_case("conv_dw_K32768", torch.randn((256, 32768), generator=g, device=dev),
                     torch.relu(torch.randn((512, 32768), generator=g, device=dev)), out)
        # edge encoder layer 2 (models.py:149): A1 = relu(.) [E, F], W2 [2F, F] ~ U(+-1/sqrt(F))
        w2 = (torch.rand((K2, F), generator=g, device=dev) * 2 - 1) / F ** 0.5
        a1 = torch.relu(torch.randn((4096, F), generator=g, device=dev) * 0.5)
        run_case(f"mlp2_fwd_F{F}", a1, w2, out)
    worst = {}
    for l in out:
        worst[l["scheme"]] = max(worst.get(l["scheme"], 0.0), l["nerr"])
    print(json.dumps(dict(summary="worst normalised max error per scheme", worst=worst)))
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            for l in out:
                f.write(json.dumps(l) + "\n")
            f.write(json.dumps(dict(summary="worst", worst=worst)) + "\n")


if __name__ == "__main__":
    main()
