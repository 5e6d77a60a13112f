"""torch.autograd.Function wrappers over the C ABI of librobotgnn.so.

Every forward/backward here is a sequence of kernel launches on torch's current stream - no host sync,
no torch math on the data path (torch only allocates the output tensors).
Reference math: SURVEY.md App. A; reference code: dgl/model/models.py:142-155 (edge encoder),
:207-211 (FiLM message / mean), :182-184 (concat + 1x1 compress).
"""
import torch

from . import _lib
from ._lib import lib, check, ptr, stream, dtype_code

FOLD_IDENTITY_MIN_BETA = 0.25   # GCN2Conv: fold (1-beta) I + beta W into one bf16 operand only for beta >= this (layers 1-3 at lambda = 1)

_POST = {"none": _lib.POST_NONE, "sum": _lib.POST_NONE, "mean": _lib.POST_MEAN, "scale": _lib.POST_SCALE}


def _row_shape(x):
    f = x.shape[1] if x.dim() > 1 else 1
    p = 1
    for s in x.shape[2:]:
        p *= s
    return f, p


def _deg_power(ix, t):
    """Exponent q such that the index-side vector t is deg^-q (None: 0); None when t is not one of the index's degree norms."""
    if t is None:
        return 0.0
    for name, q in (("inv_in_deg", 1.0), ("inv_out_deg", 1.0), ("rsqrt_in", 0.5), ("rsqrt_out", 0.5)):
        ref = getattr(ix, name, None)
        if ref is not None and t.data_ptr() == ref.data_ptr() and t.shape == ref.shape:
            return q
    return None


def _uniform_mean(ix, nbr_scale, post, dst_scale, addend):
    if addend is not None or not (2 <= ix.complete_k <= 8) or ix.N != ix.G * ix.complete_k:
        return False
    a = _deg_power(ix, nbr_scale)
    b = {_lib.POST_NONE: 0.0, _lib.POST_MEAN: 1.0}.get(post, _deg_power(ix, dst_scale))
    return a is not None and b is not None and a + b == 1.0


def _aggregate(ix, x, z, use_beta, nbr_scale, post, dst_scale, addend, out, copy_to, transposed,
               ldx=None, ldadd=None, ldout=None, shape=None):
    """One rgnn_aggregate launch.  x [N, F, *] (node stride ldx); out node stride ldout."""
    f, p = shape if shape is not None else _row_shape(x)
    if z is None:            # no per-channel modulation: channels and pixels are interchangeable
        f, p = 1, f * p
        if _uniform_mean(ix, nbr_scale, post, dst_scale, addend):
            # complete K-robot graph: in-degree = out-degree = K-1 at every node, so this normalised sum (mean, GraphConv
            # norm both/left/right, GCN2Conv, and the transposed gathers of their backward passes) IS the plain mean over the
            # other K-1 robots, and the transposed graph is the graph itself -> the streaming complete-graph kernel
            nbr_scale, post, dst_scale, transposed, use_beta = None, _lib.POST_MEAN, None, False, True
            if (f * p) % 64 == 0:
                f, p = f * p // 64, 64
    row = f * p
    buckets = getattr(ix, "buckets", None)
    if (buckets and z is not None and not transposed and post == _lib.POST_MEAN and nbr_scale is None and dst_scale is None
            and addend is None and use_beta):
        # ragged batch of complete graphs: one compiled-topology launch per robot count instead of the CSR-walking kernels
        nbytes = x.element_size() * (ix.N * row * (2 + (copy_to is not None)) + ix.E * 2 * f)
        done = True
        with _lib.timed("aggregate_fwd", nbytes):
            for k, cnt, bases in buckets:
                rc = lib().rgnn_film_bucket_forward(dtype_code(x), ptr(x), row if ldx is None else ldx, ptr(z), ptr(out),
                                                    row if ldout is None else ldout, ptr(copy_to), ptr(bases), cnt, k, f, p, stream())
                if rc == -4:          # shape not covered (same answer for every bucket: nothing has been launched yet)
                    done = False
                    break
                check(rc, "rgnn_film_bucket_forward")
        if done:
            return out
    if transposed:
        ptr32, nbr, eid = ix.out_ptr32, ix.out_dst32, ix.out_eid32
    else:
        ptr32, nbr, eid = ix.in_ptr32, ix.in_src32, ix.in_eid32
    nbytes = x.element_size() * (ix.N * row * (2 + (addend is not None) + (copy_to is not None)) + (ix.E * 2 * f if z is not None else 0))
    with _lib.timed("aggregate_bwd" if transposed else "aggregate_fwd", nbytes):
      check(lib().rgnn_aggregate(dtype_code(x), ptr(x), row if ldx is None else ldx, ptr(z), int(use_beta), ptr(nbr_scale), post,
                               ptr(dst_scale), ptr(addend), row if ldadd is None else ldadd, ptr(out),
                               row if ldout is None else ldout, ptr(copy_to),
                               ptr(ptr32), ptr(nbr), ptr(eid), ptr(ix.noff32), ix.G, ix.max_nodes, ix.max_edges,
                               0 if transposed else ix.complete_k, ix.N, ix.E, f, p, stream()), "rgnn_aggregate")
    return out


class AggregateFn(torch.autograd.Function):
    """out[v] = post_v( sum_{e in in(v)} (gamma_e * ns[src_e] * x[src_e] + beta_e) ).

    z given  -> FiLM message + mean reduce of the reference GCN (edge_udf/node_udf, models.py:207-211).
    z None   -> update_all(copy_u, sum|mean) with optional GraphConv-style degree scales.
    """

    @staticmethod
    def forward(ctx, x, z, g, post, nbr_scale, dst_scale, use_beta=True):
        ix = g.index()
        _lib.require_cuda(x, z)
        x = x.contiguous()
        ctx.use_beta = bool(use_beta)
        if z is not None:
            if use_beta and (post != "mean" or nbr_scale is not None):
                raise ValueError("FiLM modulation is defined for the mean reduce without degree scales")
            if not use_beta and z.requires_grad:
                raise NotImplementedError("gradients with respect to edge weights are not implemented")
            z = z.contiguous()
            if z.dtype != x.dtype:
                raise TypeError("z and x must share a dtype")
            if z.shape != (ix.E, 2 * x.shape[1]):
                raise ValueError(f"z must be [E, 2F] = [{ix.E}, {2 * x.shape[1]}], got {tuple(z.shape)}")
        if x.shape[0] != ix.N:
            raise ValueError(f"feature rows ({x.shape[0]}) != graph nodes ({ix.N})")
        out = torch.empty_like(x)
        with _lib.on_device(x.device):
            _aggregate(ix, x, z, ctx.use_beta, nbr_scale, _POST[post], dst_scale, None, out, None, False)
        ctx.g, ctx.post = g, post
        ctx.save_for_backward(x, z, nbr_scale, dst_scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, z, nbr_scale, dst_scale = ctx.saved_tensors
        ix = ctx.g.index()
        dout = dout.contiguous()
        dx = dz = None
        with _lib.on_device(dout.device):
            if z is not None and ctx.use_beta and ctx.needs_input_grad[0] and ctx.needs_input_grad[1]:
                dx, dz = _film_backward(ix, x, dout, z, None, False)
                return dx, dz, None, None, None, None, None
            if ctx.needs_input_grad[0]:
                back_nbr = {"mean": ix.inv_in_deg, "scale": dst_scale}.get(ctx.post)
                dx = torch.empty_like(x)
                _aggregate(ix, dout, z, False, back_nbr, _lib.POST_SCALE if nbr_scale is not None else _lib.POST_NONE,
                           nbr_scale, None, dx, None, True)
            if z is not None and ctx.use_beta and ctx.needs_input_grad[1]:
                f, p = _row_shape(x)
                dz = torch.empty_like(z)
                with _lib.timed("film_edge_grad", x.element_size() * (2.0 * ix.E * f * p + 4.0 * ix.E * f)):
                    check(lib().rgnn_film_edge_grad(dtype_code(x), ptr(x), f * p, ptr(dout), f * p, ptr(z), 0, ptr(ix.inv_in_deg),
                                                    ptr(ix.src32), ptr(ix.dst32), ptr(dz), ix.E, f, p, stream()), "rgnn_film_edge_grad")
        return dx, dz, None, None, None, None, None


class FilmMeanCatFn(torch.autograd.Function):
    """torch.cat((h, g_h), dim=1) with g_h = FiLM-mean(h)  (dgl/model/models.py:181-182) in ONE kernel: the
    aggregate lands in the second channel half of the output and h is copied into the first half from the
    tile already staged in shared memory."""

    @staticmethod
    def forward(ctx, h, z, g):
        ix = g.index()
        _lib.require_cuda(h, z)
        h = h.contiguous()
        f, p = _row_shape(h)
        if z is not None:
            z = z.contiguous()
            if z.dtype != h.dtype or z.shape != (ix.E, 2 * f):
                raise ValueError("z must be [E, 2F] in the dtype of h")
        if h.shape[0] != ix.N:
            raise ValueError(f"feature rows ({h.shape[0]}) != graph nodes ({ix.N})")
        out = torch.empty((h.shape[0], 2 * f) + tuple(h.shape[2:]), dtype=h.dtype, device=h.device)
        with _lib.on_device(h.device):
            _aggregate(ix, h, z, True, None, _lib.POST_MEAN, None, None, out.view(-1)[f * p:], out, False,
                       ldout=2 * f * p, shape=(f, p))
        ctx.g = g
        ctx.save_for_backward(h, z)
        return out

    @staticmethod
    def backward(ctx, dout):
        h, z = ctx.saved_tensors
        ix = ctx.g.index()
        dout = dout.contiguous()
        f, p = _row_shape(h)
        dagg = dout.view(-1)[f * p:]          # second channel half, node stride 2FP
        dh = dz = None
        with _lib.on_device(dout.device):
            if z is not None and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
                dh, dz = _film_backward(ix, h, dagg, z, dout, False, ldd=2 * f * p, ldadd=2 * f * p, dh=torch.empty_like(h))
            elif ctx.needs_input_grad[0]:
                dh = torch.empty_like(h)
                _aggregate(ix, dagg, None, False, ix.inv_in_deg, _lib.POST_NONE, None, dout, dh, None, True,
                           ldx=2 * f * p, ldadd=2 * f * p, shape=(f, p))
        return dh, dz, None


def film_mean_cat(g, h, z):
    """cat(h, FiLM-mean(h)) along channels, [N, 2F, ...]."""
    return FilmMeanCatFn.apply(h, z, g)


def film_mean(g, h, z):
    """agg = mean over in-edges of (gamma_e * h[src_e] + beta_e); z = [E, 2F] interleaved (gamma, beta)."""
    return AggregateFn.apply(h, z, g, "mean", None, None)


def copy_u_reduce(g, x, op="sum", nbr_scale=None, dst_scale=None, edge_weight=None):
    """update_all(fn.copy_u, fn.sum|fn.mean) with optional per-source / per-destination scales; with edge_weight [E] the message is
    fn.u_mul_e (DGL GraphConv / GCN2Conv `edge_weight`): the scalar rides in the gamma slot of the per-edge modulation rows."""
    post = "scale" if (op == "sum" and dst_scale is not None) else op
    if edge_weight is None:
        return AggregateFn.apply(x, None, g, post, nbr_scale, dst_scale)
    ix = g.index()
    ew = edge_weight.reshape(-1)
    if ew.shape[0] != ix.E:
        raise ValueError(f"edge_weight must have one value per edge ({ix.E}), got {tuple(edge_weight.shape)}")
    if ew.requires_grad:
        raise NotImplementedError("gradients with respect to edge_weight are not implemented (the reference never passes one)")
    f = x.shape[1] if x.dim() > 1 else 1
    z = torch.zeros((ix.E, 2 * f), dtype=x.dtype, device=x.device)      # (gamma, beta) rows: gamma = w_e for every channel, beta unused
    z[:, 0::2] = ew.to(x.dtype).unsqueeze(1)
    return AggregateFn.apply(x, z, g, post, nbr_scale, dst_scale, False)


def _gemm(a, sa, b, sb, c, sc, m, n, k, batch=1, red=1, split=1, partial=None, epi=0, bias=None, aux=None, alpha=1.0, beta=0.0,
          tag="gemm"):
    with _lib.timed(tag, 2.0 * m * n * k * batch * red):
      check(lib().rgnn_gemm(dtype_code(a), ptr(a), *sa, dtype_code(b), ptr(b), *sb, dtype_code(c), ptr(c), *sc,
                          m, n, k, batch, red, split, ptr(partial), epi, ptr(bias), ptr(aux), alpha, beta, stream()), "rgnn_gemm")


def _cast_w(w, transpose=False, both=False, alpha=1.0, ident=0.0):
    """fp32 master weight [R, C] -> bf16 [R, C] and/or its transpose [C, R] (one small kernel per step); with alpha / ident the
    cast is of alpha * W + ident * I (GCN2Conv's identity mapping folded into the weight operand)."""
    r, c = w.shape
    o = torch.empty((r, c), dtype=torch.bfloat16, device=w.device) if (both or not transpose) else None
    ot = torch.empty((c, r), dtype=torch.bfloat16, device=w.device) if (both or transpose) else None
    if alpha == 1.0 and ident == 0.0:
        check(lib().rgnn_tc_cast_weights(ptr(w), r, c, ptr(o), ptr(ot), stream()), "rgnn_tc_cast_weights")
    else:
        check(lib().rgnn_tc_cast_weights_affine(ptr(w), r, c, float(alpha), float(ident), ptr(o), ptr(ot), stream()), "rgnn_tc_cast_weights_affine")
    return (o, ot) if both else (ot if transpose else o)


def _cached_casts(w, wf, split3=False):
    """bf16 copy AND bf16 transpose (or the split-bf16 planes of both) of a weight, cached on the parameter object `w` until the
    parameter changes (`_version`: optimizer steps and load_state_dict write in place and bump it).  One cast kernel produces both
    orientations, so a training step pays one launch per weight instead of one per use (forward + dX), and a loop over fixed
    weights (eval, the bench) pays none.  `wf` is the fp32 contiguous 2-D view the kernels read.  During CUDA-graph capture the
    cast is always re-issued into the cached buffers, so a replayed graph re-casts the weights it finds at replay time."""
    r, c = wf.shape
    key = "_rgnn_cast3" if split3 else "_rgnn_cast"
    ent = getattr(w, key, None)
    ver = w._version
    fresh = ent is not None and ent[0] == ver and ent[1].device == wf.device and ent[3] == wf.data_ptr() and tuple(ent[1].shape[-2:]) == (r, c)
    if fresh and not torch.cuda.is_current_stream_capturing():
        return ent[1], ent[2]
    if ent is not None and ent[1].device == wf.device and tuple(ent[1].shape[-2:]) == (r, c):
        o, ot = ent[1], ent[2]
    elif split3:
        o = torch.empty((3, r, c), dtype=torch.bfloat16, device=wf.device)
        ot = torch.empty((3, c, r), dtype=torch.bfloat16, device=wf.device)
    else:
        o = torch.empty((r, c), dtype=torch.bfloat16, device=wf.device)
        ot = torch.empty((c, r), dtype=torch.bfloat16, device=wf.device)
    if split3:
        check(lib().rgnn_tc_split3_weights(ptr(wf), r, c, ptr(o), ptr(ot), stream()), "rgnn_tc_split3_weights")
    else:
        check(lib().rgnn_tc_cast_weights(ptr(wf), r, c, ptr(o), ptr(ot), stream()), "rgnn_tc_cast_weights")
    try:
        setattr(w, key, (ver, o, ot, wf.data_ptr()))
    except Exception:
        pass
    return o, ot


_PERM = {}


def _dx_film_perm(f, device):
    """Row order of the permuted Wc^T operand of rgnn_tc_conv_dx_film_bwd (include/robot_gnn.h): per 64-channel tile and group of 8
    channels, the 8 direct rows then the 8 aggregated rows."""
    key = (f, str(device))
    if key not in _PERM:
        c = torch.arange(f).view(f // 64, 8, 1, 8)                     # [tile, group, 1, lane]
        _PERM[key] = torch.cat((c, c + f), dim=2).reshape(-1).to(device)   # [tile, group, {direct, agg}, lane]
    return _PERM[key]


def _cached_wt_perm(w, wt, f):
    """The permuted transposed weight for the fused dX + FiLM-backward kernel, cached on the parameter like its casts."""
    ent = getattr(w, "_rgnn_castp", None)
    ver = w._version
    if ent is not None and ent[0] == ver and ent[2] == wt.data_ptr() and not torch.cuda.is_current_stream_capturing():
        return ent[1]
    wp = ent[1] if (ent is not None and ent[1].shape == wt.shape and ent[1].device == wt.device) else torch.empty_like(wt)
    torch.index_select(wt, 0, _dx_film_perm(f, wt.device), out=wp)     # (a row gather of a 1 MB weight copy: host-side plumbing)
    try:
        w._rgnn_castp = (ver, wp, wt.data_ptr())
    except Exception:
        pass
    return wp


def _split3(x, n_valid=None, plane=None):
    """fp32 tensor -> bf16 planes [3, plane] with x = x0 + x1 + x2 (rgnn_tc_split3); elements [n_valid, plane) are zero."""
    n = x.numel() if n_valid is None else n_valid
    plane = n if plane is None else plane
    out = torch.empty((3, plane), dtype=torch.bfloat16, device=x.device)
    check(lib().rgnn_tc_split3(ptr(x), n, plane, ptr(out), stream()), "rgnn_tc_split3")
    return out


def _split3_w(w, transpose=False):
    """fp32 master weight [R, C] -> bf16 planes [3, R, C] (or of the transpose, [3, C, R])."""
    r, c = w.shape
    out = torch.empty((3, c, r) if transpose else (3, r, c), dtype=torch.bfloat16, device=w.device)
    check(lib().rgnn_tc_split3_weights(ptr(w), r, c, None if transpose else ptr(out), ptr(out) if transpose else None, stream()),
          "rgnn_tc_split3_weights")
    return out


SMALL_K_BLOCKS = 48     # weight-gradient GEMMs with no more k-blocks than this run as ONE split: the tiles are written straight to the
                        # result and the partial-sum launch disappears (the script batch of 8 graphs is launch-bound, not SM-bound)


def _tc_splits(tiles, items):
    if items <= SMALL_K_BLOCKS:
        return 1
    return max(1, min(items, 148 // max(1, tiles)))


TC3_MAX_K_PER_ACC = 4096      # split-bf16 products: elements of the original K one TMEM accumulator may sum (error study)


def _tc3_splits(tiles, kblocks, k_per_block=64):
    """Split count of a split-plane weight-gradient GEMM: fill the SMs AND keep every accumulator's reduction short."""
    need = (kblocks * k_per_block + TC3_MAX_K_PER_ACC - 1) // TC3_MAX_K_PER_ACC
    return max(1, min(kblocks, max(_tc_splits(tiles, kblocks), need)))


def _splits(tiles, work_items):
    """Enough split-K slices to fill 148 SMs about twice, at most one per work item."""
    s = max(1, (2 * 148 + tiles - 1) // tiles)
    return max(1, min(s, work_items))


def _splits_mn(m, n, work_items):
    """Split count for an [m, n] CUDA-core product: the large fp32 products run 128-wide tiles, two CTAs per SM (k_sgemm_v), so
    aim for ~4 CTAs per SM; small ones keep the 64 x 64 rule."""
    if m >= 64 and n >= 64:
        tiles = ((m + 127) // 128) * ((n + 127) // 128 if n > 64 else 1)
        return max(1, min((4 * 148 + tiles - 1) // tiles, work_items))
    return _splits(((m + 63) // 64) * ((n + 63) // 64), work_items)


def _f32c(t):
    """fp32 contiguous view of a parameter for the C ABI: the tensor itself when it already is one (the usual case; every no-op
    .detach() / .float() / .contiguous() still builds a tensor object, and the script batch is host-bound)."""
    return t if (t.dtype is torch.float32 and t.is_contiguous()) else t.detach().float().contiguous()


def _edge_mlp_forward(pose, w1, b1, w2, b2, out_dtype):
    _lib.require_cuda(pose, w1, b1, w2, b2)
    pose = _f32c(pose)                        # edge.float(), models.py:153
    e, hid, f2 = pose.shape[0], w1.shape[0], w2.shape[0]
    if pose.shape[1] != 9 or w1.shape != (hid, 9) or w2.shape != (f2, hid) or f2 % 2:
        raise ValueError("edge encoder expects pose [E,9], W1 [H,9], W2 [2F,H]")
    w2_param = w2
    w1, b1, w2, b2 = _f32c(w1), _f32c(b1), _f32c(w2), _f32c(b2)
    w2t = None                    # the K-major operand of dA1 = dZ . W2, produced by the same cast launch as the forward operand
    a1 = torch.empty((e, hid), dtype=out_dtype, device=pose.device)
    z = torch.empty((e, f2), dtype=out_dtype, device=pose.device)
    with _lib.on_device(pose.device):
        check(lib().rgnn_edge_mlp1_fwd(dtype_code(a1), ptr(pose), ptr(w1), ptr(b1), ptr(a1), e, hid, stream()), "rgnn_edge_mlp1_fwd")
        if _lib.USE_TC and out_dtype == torch.bfloat16 and hid % 64 == 0 and f2 % 128 == 0 and e > 0:
            w2b, w2t = _cached_casts(w2_param, w2)
            with _lib.timed("mlp2_fwd", 2.0 * e * f2 * hid):
                check(lib().rgnn_tc_gemm_tn(ptr(a1), hid, ptr(w2b), hid, ptr(z), f2, e, f2, hid,
                                            _lib.EPI_BIAS_N | _lib.EPI_SIGMOID, ptr(b2), None, 1.0, 0.0, stream()), "rgnn_tc_gemm_tn")
        elif _lib.USE_TC3 and out_dtype == torch.float32 and hid % 64 == 0 and f2 % 128 == 0 and e > 0:
            # fp32 parity mode on the tensor cores: split-bf16 planes, six plane-pair products, fp32 accumulate and output
            prow = (e + 63) // 64 * 64
            a1p = _split3(a1, e * hid, prow * hid)
            w2p, w2t = _cached_casts(w2_param, w2, split3=True)
            with _lib.timed("mlp2_fwd", 2.0 * e * f2 * hid):
                check(lib().rgnn_tc3_gemm_tn(ptr(a1p), hid, prow, ptr(w2p), hid, f2, ptr(z), f2, e, f2, hid,
                                             _lib.EPI_BIAS_N | _lib.EPI_SIGMOID, ptr(b2), 1.0, stream()), "rgnn_tc3_gemm_tn")
        else:
            _gemm(a1, (0, 0, hid, 1), w2, (0, 0, 1, hid), z, (0, f2, 1), e, f2, hid,
                  epi=_lib.EPI_BIAS_N | _lib.EPI_SIGMOID, bias=b2, tag="mlp2_fwd")
    return pose, w2, a1, z, w2t


class EdgeMLPFn(torch.autograd.Function):
    """z = sigmoid(relu(pose W1^T + b1) W2^T + b2)  - edge_encoder.layers (models.py:146-151); [E, 2F]."""

    @staticmethod
    def forward(ctx, pose, w1, b1, w2, b2, out_dtype):
        pose, w2, a1, z, ctx.w2t = _edge_mlp_forward(pose, w1, b1, w2, b2, out_dtype)
        ctx.save_for_backward(pose, w2, a1, z)
        return z

    @staticmethod
    def backward(ctx, dz):
        pose, w2, a1, z = ctx.saved_tensors
        dz = dz.contiguous()
        dev = dz.device
        with _lib.on_device(dev):
            dzp = torch.empty_like(z)
            check(lib().rgnn_sigmoid_grad(dtype_code(z), ptr(z), ptr(dz), ptr(dzp), z.numel(), stream()), "rgnn_sigmoid_grad")
            dw1, db1, dw2, db2 = edge_mlp_param_grads(pose, w2, a1, dzp, w2t=ctx.w2t)
        return None, dw1, db1, dw2, db2, None


def edge_mlp_param_grads(pose, w2, a1, dzp, w2t=None, db2=None):
    """Parameter gradients of the edge encoder from dzp = dL/d(pre-sigmoid) [E, 2F] (autograd of models.py:146-151).
    w2t: the transposed bf16 weight (or its split planes) the forward's cast already produced."""
    e, hid = a1.shape
    f2 = w2.shape[0]
    dev = a1.device
    f32 = dict(dtype=torch.float32, device=dev)
    da1 = torch.empty_like(a1)
    dw2 = torch.empty((f2, hid), **f32)
    if _lib.USE_TC and a1.dtype == torch.bfloat16 and hid % 128 == 0 and f2 % 128 == 0 and e > 0:
        if w2t is None or w2t.dim() != 2:
            w2t = _cast_w(w2, transpose=True)                   # [hid, f2], K-major for dA1 = dZ . W2
        with _lib.timed("mlp2_dx", 2.0 * e * f2 * hid):
            check(lib().rgnn_tc_gemm_tn(ptr(dzp), f2, ptr(w2t), f2, ptr(da1), hid, e, hid, f2, _lib.EPI_RELU_MASK, None, ptr(a1),
                                        1.0, 0.0, stream()), "rgnn_tc_gemm_tn")
        split = _tc_splits((f2 // 128) * (hid // (256 if hid % 256 == 0 else 128)), (e + 63) // 64)
        partial = torch.empty((split, f2, hid), **f32)
        with _lib.timed("mlp2_dw", 2.0 * e * f2 * hid):
            check(lib().rgnn_tc_gemm_nn(ptr(dzp), f2, ptr(a1), hid, ptr(dw2), hid, f2, hid, e, 1.0, ptr(partial), split, stream()),
                  "rgnn_tc_gemm_nn")
    elif _lib.USE_TC3 and a1.dtype == torch.float32 and dzp.dtype == torch.float32 and hid % 128 == 0 and f2 % 128 == 0 and e > 0:
        prow = (e + 63) // 64 * 64                              # rows beyond E are zero planes: they are the K tail of dW2 = dZ^T . A1
        dzp3, a13 = _split3(dzp, e * f2, prow * f2), _split3(a1, e * hid, prow * hid)
        w2t3 = w2t if (w2t is not None and w2t.dim() == 3) else _split3_w(w2, transpose=True)    # [3, hid, f2]: K-major for dA1 = dZ . W2
        with _lib.timed("mlp2_dx", 2.0 * e * f2 * hid):
            check(lib().rgnn_tc3_gemm_tn(ptr(dzp3), f2, prow, ptr(w2t3), f2, hid, ptr(da1), hid, e, hid, f2, 0, None, 1.0, stream()),
                  "rgnn_tc3_gemm_tn")
        check(lib().rgnn_relu_grad(dtype_code(a1), ptr(a1), ptr(da1), ptr(da1), a1.numel(), stream()), "rgnn_relu_grad")
        split = _tc3_splits((f2 // 128) * (hid // (256 if hid % 256 == 0 else 128)), prow // 64)
        partial = torch.empty((split, f2, hid), **f32)
        with _lib.timed("mlp2_dw", 2.0 * e * f2 * hid):
            check(lib().rgnn_tc3_gemm_nn(ptr(dzp3), f2, prow, ptr(a13), hid, prow, ptr(dw2), hid, f2, hid, e, 1.0, ptr(partial), split,
                                         stream()), "rgnn_tc3_gemm_nn")
    else:
        _gemm(dzp, (0, 0, f2, 1), w2, (0, 0, hid, 1), da1, (0, hid, 1), e, hid, f2, epi=_lib.EPI_RELU_MASK, aux=a1, tag="mlp2_dx")
        split = _splits_mn(f2, hid, max(1, (e + 63) // 64))
        partial = torch.empty((split, f2, hid), **f32) if split > 1 else None
        _gemm(dzp, (0, 0, 1, f2), a1, (0, 0, hid, 1), dw2, (0, hid, 1), f2, hid, e, split=split, partial=partial, tag="mlp2_dw")
    if db2 is None:       # (the fused dX + FiLM-backward kernel hands the column sums over already)
        db2 = torch.empty((f2,), **f32)
        check(lib().rgnn_reduce_sum(dtype_code(dzp), ptr(dzp), e, f2, f2, 1, 1, 1, ptr(db2), stream()), "rgnn_reduce_sum")
    dw1, db1 = torch.empty((hid, 9), **f32), torch.empty((hid,), **f32)
    check(lib().rgnn_edge_mlp1_bwd(dtype_code(da1), ptr(pose), ptr(da1), ptr(dw1), ptr(db1), e, hid, stream()), "rgnn_edge_mlp1_bwd")
    return dw1, db1, dw2, db2


def edge_mlp(pose, w1, b1, w2, b2, out_dtype=torch.float32):
    return EdgeMLPFn.apply(pose, w1, b1, w2, b2, out_dtype)


def _compress_forward(h, gh, wc, bc):
    """-> (y, wc2 fp32 [Fo, 2F], mode, planes, wt): mode 0 = CUDA-core products, 1 = bf16 tcgen05, 3 = fp32 as split-bf16 planes on
    tcgen05 (planes = (h3, g3), kept for the weight gradient; (None, None) otherwise); wt = the transposed weight operand of dX
    that the forward's cast launch produced alongside (None in mode 0)."""
    _lib.require_cuda(h, gh, wc, bc)
    n, f = h.shape[0], h.shape[1]
    _, p = _row_shape(h)
    wc2 = wc.detach().float().reshape(wc.shape[0], -1).contiguous()
    fo = wc2.shape[0]
    if wc2.shape[1] != 2 * f or gh.shape != h.shape or h.dtype != gh.dtype:
        raise ValueError("compress expects h, g_h [N,F,...] and weight [Fo, 2F(,1,1)]")
    bcf = _f32c(bc)
    y = torch.empty((n, fo) + tuple(h.shape[2:]), dtype=h.dtype, device=h.device)
    shape_ok = p == 64 and f % 64 == 0 and fo % 128 == 0 and n > 0
    tc = 1 if (_lib.USE_TC and h.dtype == torch.bfloat16 and shape_ok) else (3 if (_lib.USE_TC3 and h.dtype == torch.float32 and shape_ok) else 0)
    planes, wt = (None, None), None
    with _lib.on_device(h.device):
        if tc == 1:
            wb, wt = _cached_casts(wc, wc2)
            with _lib.timed("compress_fwd", 2.0 * fo * 2 * f * n * p):
                check(lib().rgnn_tc_conv1x1(ptr(wb), ptr(h), f, ptr(gh), f, ptr(y), fo, None, 0, ptr(bcf), n, stream()), "rgnn_tc_conv1x1")
        elif tc == 3:
            planes = (_split3(h), _split3(gh))
            w3, wt = _cached_casts(wc, wc2, split3=True)
            with _lib.timed("compress_fwd", 2.0 * fo * 2 * f * n * p):
                check(lib().rgnn_tc3_conv1x1(ptr(w3), ptr(planes[0]), f, ptr(planes[1]), f, n, ptr(y), fo, None, 0, ptr(bcf), n, stream()),
                      "rgnn_tc3_conv1x1")
        else:
            if p == 1:    # [N,F,1,1] maps: the nodes form the N dimension of one product
                sb, sc, bt, nn = (0, 0, 1, f), (0, 1, fo), 1, n
            else:         # one [Fo x P] product per node
                sb, sc, bt, nn = (f * p, 0, p, 1), (fo * p, p, 1), n, p
            _gemm(wc2, (0, 0, 2 * f, 1), h, sb, y, sc, fo, nn, f, batch=bt, epi=_lib.EPI_BIAS_M, bias=bcf, tag="compress_fwd")
            _gemm(wc2[:, f:], (0, 0, 2 * f, 1), gh, sb, y, sc, fo, nn, f, batch=bt, epi=_lib.EPI_ADD_AUX, aux=y, beta=1.0, tag="compress_fwd")
    return y, wc2, tc, planes, wt


def _compress_backward(dy, h, gh, wc2, tc, need_dh, need_dgh, need_dw, need_db, planes=(None, None), wt=None):
    """Autograd of the concat + 1x1 conv: dh (direct), dg_h, dWc [Fo, 2F] fp32, dbc [Fo] fp32."""
    n, f = h.shape[0], h.shape[1]
    _, p = _row_shape(h)
    fo = wc2.shape[0]
    dev = dy.device
    f32 = dict(dtype=torch.float32, device=dev)
    dh = dgh = dwc = dbc = None
    tc_dx = tc == 1 and f % 128 == 0 and fo % 64 == 0
    tc_dw = tc == 1 and f % 256 == 0 and fo % 128 == 0
    tc3_dx = tc == 3 and f % 128 == 0 and fo % 64 == 0
    tc3_dw = tc == 3 and f % 256 == 0 and fo % 128 == 0 and planes[0] is not None
    with _lib.on_device(dev):
        if (tc3_dx and (need_dh or need_dgh)) or (tc3_dw and need_dw):
            dy3 = _split3(dy)
        if tc3_dx and (need_dh or need_dgh):
            wt3 = wt if (wt is not None and wt.dim() == 3) else _split3_w(wc2, transpose=True)            # [3, 2F, Fo]
            dh, dgh = torch.empty_like(h), torch.empty_like(gh)
            with _lib.timed("compress_dx", 2.0 * fo * 2 * f * n * p):
                check(lib().rgnn_tc3_conv1x1(ptr(wt3), ptr(dy3), fo, None, 0, n, ptr(dh), f, ptr(dgh), f, None, n, stream()), "rgnn_tc3_conv1x1")
        if tc3_dw and need_dw:
            dwc = torch.empty((fo, 2 * f), **f32)
            split = _tc3_splits((fo // 128) * (2 * f // 256), n)
            partial = torch.empty((split * fo * 2 * f,), **f32)
            with _lib.timed("compress_dw", 2.0 * fo * 2 * f * n * p):
                check(lib().rgnn_tc3_conv1x1_wgrad(ptr(dy3), fo, ptr(planes[0]), f, ptr(planes[1]), f, n, ptr(dwc), 2 * f, ptr(partial), split, n,
                                                   stream()), "rgnn_tc3_conv1x1_wgrad")
        if tc_dx and (need_dh or need_dgh):
            if wt is None or wt.dim() != 2:
                wt = _cast_w(wc2, transpose=True)           # [2F, Fo] K-major: dX_v = Wc^T . dY_v
            dh, dgh = torch.empty_like(h), torch.empty_like(gh)
            with _lib.timed("compress_dx", 2.0 * fo * 2 * f * n * p):
                check(lib().rgnn_tc_conv1x1(ptr(wt), ptr(dy), fo, None, 0, ptr(dh), f, ptr(dgh), f, None, n, stream()), "rgnn_tc_conv1x1")
        if tc_dw and need_dw:
            dwc = torch.empty((fo, 2 * f), **f32)
            split = _tc_splits((fo // 128) * (2 * f // 256), n)
            partial = torch.empty((split * fo * (2 * f + 1),), **f32)
            if need_db:       # the bias gradient rides on the same pass over dy (idle epilogue warps row-sum the staged tiles)
                dbc = torch.empty((fo,), **f32)
                need_db = False
            with _lib.timed("compress_dw", 2.0 * fo * 2 * f * n * p):
                check(lib().rgnn_tc_conv1x1_wgrad(ptr(dy), fo, ptr(h), f, ptr(gh), f, ptr(dwc), 2 * f, ptr(dbc), ptr(partial), split, n,
                                                  stream()), "rgnn_tc_conv1x1_wgrad")
        if p == 1:
            sb, sc, bt, nn = (0, 0, 1, fo), (0, 1, f), 1, n
        else:
            sb, sc, bt, nn = (fo * p, 0, p, 1), (f * p, p, 1), n, p
        if need_dh and dh is None:
            dh = torch.empty_like(h)
            _gemm(wc2, (0, 0, 1, 2 * f), dy, sb, dh, sc, f, nn, fo, batch=bt, tag="compress_dx")
        if need_dgh and dgh is None:
            dgh = torch.empty_like(gh)
            _gemm(wc2[:, f:], (0, 0, 1, 2 * f), dy, sb, dgh, sc, f, nn, fo, batch=bt, tag="compress_dx")
        if need_dw and dwc is None:
            dwc = torch.empty((fo, 2 * f), **f32)
            split = _splits_mn(fo, f, n)
            partial = torch.empty((split, fo, f), **f32) if split > 1 else None
            for half, src in ((0, h), (1, gh)):
                # dWc[o, k2] = sum_v sum_p dy[v,o,p] * src[v,k2,p]: K = P per node, reduced over the N nodes
                _gemm(dy, (0, fo * p, p, 1), src, (0, f * p, 1, p), dwc[:, half * f:], (0, 2 * f, 1), fo, f, p,
                      red=n, split=split, partial=partial, tag="compress_dw")
        if need_db:
            dbc = torch.empty((fo,), **f32)
            with _lib.timed("bias_grad", dy.element_size() * float(dy.numel())):
                check(lib().rgnn_reduce_sum(dtype_code(dy), ptr(dy), n, fo * p, fo, p, p, 1, ptr(dbc), stream()), "rgnn_reduce_sum")
    return dh, dgh, dwc, dbc


class CompressFn(torch.autograd.Function):
    """y = Conv2d(2F, F, 1)(cat(h, g_h))  - dgl/model/models.py:182-184 - without materialising the concat:
    y_v = Wc[:, :F] h_v + Wc[:, F:] g_v + bc per node (a [F x 2F] . [2F x P] product per robot)."""

    @staticmethod
    def forward(ctx, h, gh, wc, bc):
        h, gh = h.contiguous(), gh.contiguous()
        y, wc2, ctx.tc, planes, ctx.wt = _compress_forward(h, gh, wc, bc)
        ctx.save_for_backward(h, gh, wc2, *[t for t in planes if t is not None])
        ctx.wshape = tuple(wc.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        h, gh, wc2 = ctx.saved_tensors[:3]
        planes = tuple(ctx.saved_tensors[3:5]) if len(ctx.saved_tensors) >= 5 else (None, None)
        dh, dgh, dwc, dbc = _compress_backward(dy.contiguous(), h, gh, wc2, ctx.tc, *ctx.needs_input_grad, planes=planes, wt=ctx.wt)
        return dh, dgh, (dwc.reshape(ctx.wshape) if dwc is not None else None), dbc


def _film_backward(ix, h, dagg, z, addend, fuse_sig, ldd=None, ldadd=None, dh=None):
    """dh (= addend + transposed FiLM gather of dagg) and dz; one fused staged kernel when the batch allows, else two.
    dagg / addend may be channel-halves of a wider buffer (node strides ldd / ldadd); dh defaults to in-place on addend."""
    f, p = _row_shape(h)
    ldd = f * p if ldd is None else ldd
    ldadd = f * p if ldadd is None else ldadd
    if dh is None:
        dh = addend if addend is not None else torch.empty_like(h)
    dz = torch.empty_like(z)
    nbytes = h.element_size() * (ix.N * f * p * (3 + (addend is not None)) + 2 * ix.E * 2 * f)
    buckets = getattr(ix, "buckets", None)
    if buckets:
        done = True
        with _lib.timed("film_backward", nbytes):
            for k, cnt, bases in buckets:
                rc = lib().rgnn_film_bucket_backward(dtype_code(h), ptr(h), f * p, ptr(dagg), ldd, ptr(z), int(fuse_sig), ptr(addend), ldadd,
                                                     ptr(dh), f * p, ptr(dz), ptr(bases), cnt, k, f, p, stream())
                if rc == -4:
                    done = False
                    break
                check(rc, "rgnn_film_bucket_backward")
        if done:
            return dh, dz
    with _lib.timed("film_backward", nbytes):
        rc = lib().rgnn_film_backward(dtype_code(h), ptr(h), f * p, ptr(dagg), ldd, ptr(z), int(fuse_sig), ptr(addend), ldadd,
                                      ptr(dh), f * p, ptr(dz), ptr(ix.inv_in_deg), ptr(ix.ptr32), ptr(ix.idx32), ptr(ix.noff32),
                                      ix.G, ix.max_nodes, ix.max_edges, ix.complete_k, ix.N, ix.E, f, p, stream())
    if rc == -4:      # RGNN_E_UNSUPPORTED: general graphs / ragged rows -> the two global-gather kernels
        _aggregate(ix, dagg, z, False, ix.inv_in_deg, _lib.POST_NONE, None, addend, dh, None, True, ldx=ldd, ldadd=ldadd, shape=(f, p))
        check(lib().rgnn_film_edge_grad(dtype_code(h), ptr(h), f * p, ptr(dagg), ldd, ptr(z), int(fuse_sig), ptr(ix.inv_in_deg),
                                        ptr(ix.src32), ptr(ix.dst32), ptr(dz), ix.E, f, p, stream()), "rgnn_film_edge_grad")
    else:
        check(rc, "rgnn_film_backward")
    return dh, dz


class GNNLayerFn(torch.autograd.Function):
    """One whole layer of the reference's GNN slice as a single autograd node (dgl/model/models.py:181-184 / :186-189):
        z = edge_encoder(pose);  g_h = FiLM-mean(h, z);  y = Conv2d(2F,F,1)(cat(h, g_h)).
    Fusing the node lets the backward (i) add the direct and the aggregated parts of dh inside the aggregation kernel,
    (ii) apply the sigmoid derivative where dgamma/dbeta are produced, (iii) keep every intermediate in its kernel dtype."""

    @staticmethod
    def forward(ctx, h, pose, w1, b1, w2, b2, wc, bc, g):
        ix = g.index()
        _lib.require_cuda(h, pose)
        h = h.contiguous()
        if h.shape[0] != ix.N:
            raise ValueError(f"feature rows ({h.shape[0]}) != graph nodes ({ix.N})")
        if pose.shape[0] != ix.E:
            raise ValueError(f"pose rows ({pose.shape[0]}) != graph edges ({ix.E})")
        pose, w2f, a1, z, ctx.w2t = _edge_mlp_forward(pose, w1, b1, w2, b2, h.dtype)
        if z.shape[1] != 2 * h.shape[1]:
            raise ValueError("edge encoder width does not match the feature maps")
        f = h.shape[1]
        _, p = _row_shape(h)
        fo = wc.shape[0]
        need_bwd = any(ctx.needs_input_grad)
        fused = bool(_lib.USE_FUSED and _lib.USE_TC and h.dtype == torch.bfloat16 and p == 64 and 2 <= ix.complete_k <= 8 and
                     f % 32 == 0 and fo % 128 == 0 and ix.N > 0)
        if fused:
            # aggregation in the GEMM prologue: g_h is produced in shared memory; it is written out only for the backward
            wc2 = wc.detach().float().reshape(fo, -1).contiguous()
            bcf = _f32c(bc)
            y = torch.empty((ix.N, fo) + tuple(h.shape[2:]), dtype=h.dtype, device=h.device)
            gh = torch.empty_like(h) if need_bwd else None
            with _lib.on_device(h.device):
                wb, ctx.wt = _cached_casts(wc, wc2)
                with _lib.timed("compress_film_fused_fwd", 2.0 * fo * 2 * f * ix.N * p):
                    check(lib().rgnn_tc_conv_film_fused(ptr(wb), ptr(h), ptr(z), ptr(bcf), ptr(y), ptr(gh), ix.G, ix.complete_k, f, fo,
                                                        stream()), "rgnn_tc_conv_film_fused")
            ctx.tc = 1
            planes = (None, None)
            if gh is None:
                gh = h.new_empty(0)
        else:
            gh = torch.empty_like(h)
            with _lib.on_device(h.device):
                _aggregate(ix, h, z, True, None, _lib.POST_MEAN, None, None, gh, None, False)
            y, wc2, ctx.tc, planes, ctx.wt = _compress_forward(h, gh, wc, bc)
        # backward plan: FiLM backward inside the dX GEMM's epilogue (dcat never written) when the shape allows
        ctx.fused_bwd = bool(_lib.USE_FUSED_BWD and ctx.tc == 1 and need_bwd and 2 <= ix.complete_k <= 6 and p == 64 and f % 64 == 0
                             and fo % 64 == 0 and max(f, fo) <= _lib.FUSED_BWD_MAX_F and ctx.wt is not None and ctx.wt.dim() == 2 and ix.N == ix.G * ix.complete_k)
        ctx.wtp = _cached_wt_perm(wc, ctx.wt, f) if ctx.fused_bwd else None
        ctx.g, ctx.wshape = g, tuple(wc.shape)
        ctx.save_for_backward(h, pose, w2f, a1, z, gh, wc2, *[t for t in planes if t is not None and need_bwd])
        return y

    @staticmethod
    def backward(ctx, dy):
        h, pose, w2f, a1, z, gh, wc2 = ctx.saved_tensors[:7]
        planes = tuple(ctx.saved_tensors[7:9]) if len(ctx.saved_tensors) >= 9 else (None, None)
        ix = ctx.g.index()
        need = ctx.needs_input_grad
        need_edge = any(need[2:6])
        dy = dy.contiguous()
        dh = dw1 = db1 = dw2 = db2 = dzp = None
        if getattr(ctx, "fused_bwd", False) and (need[0] or need_edge):
            f, fo = h.shape[1], wc2.shape[0]
            dh, dzp = torch.empty_like(h), torch.empty_like(z)
            # db2 (bias gradient of the edge encoder's Linear(F,2F)) = column sums of dz: accumulated inside the kernel
            db2 = torch.empty((2 * f,), dtype=torch.float32, device=dy.device) if (need_edge and f <= 2048) else None
            nbytes = 2.0 * (ix.N * 64 * (2 * f + fo) + 2 * ix.E * 2 * f)
            with _lib.on_device(dy.device), _lib.timed("dx_film_bwd", 2.0 * fo * 2 * f * ix.N * 64, nbytes):
                check(lib().rgnn_tc_conv_dx_film_bwd(ptr(ctx.wtp), ptr(dy), ptr(h), ptr(z), 1, ptr(dh), ptr(dzp), ptr(db2), ix.G, ix.complete_k,
                                                     f, fo, stream()), "rgnn_tc_conv_dx_film_bwd")
            _, _, dwc, dbc = _compress_backward(dy, h, gh, wc2, ctx.tc, False, False, need[6], need[7], planes=planes, wt=ctx.wt)
        else:
            dh_direct, dgh, dwc, dbc = _compress_backward(dy, h, gh, wc2, ctx.tc, need[0], need[0] or need_edge, need[6], need[7],
                                                          planes=planes, wt=ctx.wt)
            if need[0] or need_edge:
                with _lib.on_device(dy.device):
                    dh, dzp = _film_backward(ix, h, dgh, z, dh_direct, True)
        with _lib.on_device(dy.device):
            if need[0] or need_edge:
                if need_edge:
                    dw1, db1, dw2, db2 = edge_mlp_param_grads(pose, w2f, a1, dzp, w2t=ctx.w2t, db2=db2)
        return (dh if need[0] else None, None, dw1, db1, dw2, db2,
                dwc.reshape(ctx.wshape) if dwc is not None else None, dbc, None)


def gnn_layer(g, h, pose, w1, b1, w2, b2, wc, bc):
    return GNNLayerFn.apply(h, pose, w1, b1, w2, b2, wc, bc, g)


def compress(h, gh, wc, bc):
    return CompressFn.apply(h, gh, wc, bc)


class LinearFn(torch.autograd.Function):
    """y = alpha * (x @ W) + ident * x + bias on the last dim (W is [in, out] as in DGL).
    ident = 0: GraphConv weight transform; ident = 1 - beta, alpha = beta: GCN2Conv's identity mapping
    ((1-beta) I + beta W) - both fused in the GEMM epilogue."""

    @staticmethod
    def forward(ctx, x, w, bias, alpha, ident):
        _lib.require_cuda(x, w, bias)
        x = x.contiguous()
        din, dout = w.shape
        if x.shape[-1] != din or (ident != 0.0 and din != dout):
            raise ValueError("linear: feature size mismatch")
        rows = x.numel() // din
        wf = w.detach().float().contiguous()
        y = torch.empty(tuple(x.shape[:-1]) + (dout,), dtype=x.dtype, device=x.device)
        epi, bf = 0, None
        if bias is not None:
            epi |= _lib.EPI_BIAS_N
            bf = bias.detach().float().contiguous()
        if ident != 0.0:
            epi |= _lib.EPI_ADD_AUX
        ctx.tc = bool(_lib.USE_TC and x.dtype == torch.bfloat16 and din % 128 == 0 and dout % 128 == 0 and rows > 0)
        with _lib.on_device(x.device):
            # the identity mapping is folded into the bf16 weight operand only while beta (= alpha here) is large enough for
            # beta * w_ii to survive next to 1 - beta in a bf16 diagonal; deep GCNII layers (beta = log(lambda/l + 1) -> 0) keep the
            # residual as the fp32 epilogue term alpha * acc + ident * x, like the CUDA-core path and DGL do
            ctx.fold = bool(ident == 0.0 or alpha >= FOLD_IDENTITY_MIN_BETA)
            if ctx.tc and ctx.fold:
                # tcgen05 path: (alpha W + ident I)^T in bf16 is the K-major B operand - the identity mapping lives in the
                # weight matrix (one plain GEMM, no second operand to read in the epilogue); bias fused in the epilogue
                wt = _cast_w(wf, transpose=True, alpha=float(alpha), ident=float(ident))
                with _lib.timed("linear_fwd", 2.0 * rows * din * dout):
                    check(lib().rgnn_tc_gemm_tn(ptr(x), din, ptr(wt), din, ptr(y), dout, rows, dout, din, epi & ~_lib.EPI_ADD_AUX, ptr(bf),
                                                None, 1.0, 0.0, stream()), "rgnn_tc_gemm_tn")
            elif ctx.tc:
                wt = _cast_w(wf, transpose=True)
                with _lib.timed("linear_fwd", 2.0 * rows * din * dout):
                    check(lib().rgnn_tc_gemm_tn(ptr(x), din, ptr(wt), din, ptr(y), dout, rows, dout, din, epi, ptr(bf),
                                                ptr(x), float(alpha), float(ident), stream()), "rgnn_tc_gemm_tn")
            else:
                _gemm(x, (0, 0, din, 1), wf, (0, 0, dout, 1), y, (0, dout, 1), rows, dout, din, epi=epi, bias=bf,
                      aux=x if ident != 0.0 else None, alpha=float(alpha), beta=float(ident))
        ctx.save_for_backward(x, wf)
        ctx.alpha, ctx.ident, ctx.has_bias = float(alpha), float(ident), bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wf = ctx.saved_tensors
        dy = dy.contiguous()
        din, dout = wf.shape
        rows = x.numel() // din
        f32 = dict(dtype=torch.float32, device=dy.device)
        dx = dw = db = None
        with _lib.on_device(dy.device):
            if ctx.needs_input_grad[0]:
                dx = torch.empty_like(x)
                epi = _lib.EPI_ADD_AUX if ctx.ident != 0.0 else 0
                if ctx.tc and ctx.fold:     # dX = dY . (alpha W + ident I)^T: the [in, out] matrix in bf16 is already the K-major B operand
                    wb = _cast_w(wf, alpha=ctx.alpha, ident=ctx.ident)
                    with _lib.timed("linear_dx", 2.0 * rows * din * dout):
                        check(lib().rgnn_tc_gemm_tn(ptr(dy), dout, ptr(wb), dout, ptr(dx), din, rows, din, dout, 0, None,
                                                    None, 1.0, 0.0, stream()), "rgnn_tc_gemm_tn")
                elif ctx.tc:                # small beta: dX = alpha * dY . W^T + ident * dY with the residual in the fp32 epilogue
                    wb = _cast_w(wf)
                    with _lib.timed("linear_dx", 2.0 * rows * din * dout):
                        check(lib().rgnn_tc_gemm_tn(ptr(dy), dout, ptr(wb), dout, ptr(dx), din, rows, din, dout, epi, None,
                                                    ptr(dy), ctx.alpha, ctx.ident, stream()), "rgnn_tc_gemm_tn")
                else:
                    _gemm(dy, (0, 0, dout, 1), wf, (0, 0, 1, dout), dx, (0, din, 1), rows, din, dout,
                          epi=epi, aux=dy if ctx.ident != 0.0 else None, alpha=ctx.alpha, beta=ctx.ident)
            if ctx.needs_input_grad[1] and ctx.tc:
                dw = torch.empty((din, dout), **f32)
                split = _tc_splits((din // 128) * (dout // (256 if dout % 256 == 0 else 128)), (rows + 63) // 64)
                partial = torch.empty((split, din, dout), **f32)
                with _lib.timed("linear_dw", 2.0 * rows * din * dout):
                    check(lib().rgnn_tc_gemm_nn(ptr(x), din, ptr(dy), dout, ptr(dw), dout, din, dout, rows, ctx.alpha, ptr(partial), split,
                                                stream()), "rgnn_tc_gemm_nn")
            elif ctx.needs_input_grad[1]:
                dw = torch.empty((din, dout), **f32)
                split = _splits_mn(din, dout, max(1, (rows + 63) // 64))
                partial = torch.empty((split, din, dout), **f32) if split > 1 else None
                _gemm(x, (0, 0, 1, din), dy, (0, 0, dout, 1), dw, (0, dout, 1), din, dout, rows, split=split, partial=partial,
                      alpha=ctx.alpha)
            if ctx.has_bias and ctx.needs_input_grad[2]:
                db = torch.empty((dout,), **f32)
                check(lib().rgnn_reduce_sum(dtype_code(dy), ptr(dy), rows, dout, dout, 1, 1, 1, ptr(db), stream()), "rgnn_reduce_sum")
        return dx, dw, db, None, None


def linear(x, w, bias=None, alpha=1.0, ident=0.0):
    return LinearFn.apply(x, w, bias, alpha, ident)


class AxpbyFn(torch.autograd.Function):
    """out = a * x + b * y  (GCN2Conv's initial-residual mix (1-alpha) * agg + alpha * feat_0)."""

    @staticmethod
    def forward(ctx, x, y, a, b):
        _lib.require_cuda(x, y)
        x, y = x.contiguous(), y.contiguous()
        if x.shape != y.shape or x.dtype != y.dtype:
            raise ValueError("axpby operands must match")
        out = torch.empty_like(x)
        with _lib.on_device(x.device):
            check(lib().rgnn_axpby(dtype_code(x), float(a), ptr(x), float(b), ptr(y), ptr(out), x.numel(), stream()), "rgnn_axpby")
        ctx.a, ctx.b = float(a), float(b)
        return out

    @staticmethod
    def backward(ctx, d):
        d = d.contiguous()
        outs = []
        with _lib.on_device(d.device):
            for need, coef in ((ctx.needs_input_grad[0], ctx.a), (ctx.needs_input_grad[1], ctx.b)):
                if not need:
                    outs.append(None)
                    continue
                o = torch.empty_like(d)
                check(lib().rgnn_axpby(dtype_code(d), coef, ptr(d), 0.0, ptr(d), ptr(o), d.numel(), stream()), "rgnn_axpby")
                outs.append(o)
        return outs[0], outs[1], None, None


def axpby(x, y, a, b):
    return AxpbyFn.apply(x, y, a, b)


def _mean_mix(ix, x, addend, sa, sb):
    """sa * mean over the other robots of the graph + sb * addend in one pass; None when the kernel does not cover the shape."""
    row = x.numel() // max(1, x.shape[0])
    out = torch.empty_like(x)
    nbytes = x.element_size() * ix.N * row * (2 + (addend is not None))
    with _lib.timed("mean_mix", nbytes):
        rc = lib().rgnn_complete_mean_mix(dtype_code(x), ptr(x), row, ptr(addend), row, ptr(out), row, ix.G, ix.complete_k, row,
                                          float(sa), float(sb), stream())
    if rc == -4:          # RGNN_E_UNSUPPORTED
        return None
    check(rc, "rgnn_complete_mean_mix")
    return out


class MeanMixFn(torch.autograd.Function):
    """r = sa * A_hat x + sb * x0 on complete K-robot graphs, where A_hat = D^-1/2 A D^-1/2 is the mean over the other K-1 robots
    (GCN2Conv's propagation + initial residual) - one streaming kernel forward, one backward (A_hat is symmetric)."""

    @staticmethod
    def forward(ctx, x, x0, g, sa, sb):
        ix = g.index()
        _lib.require_cuda(x, x0)
        x, x0 = x.contiguous(), x0.contiguous()
        if x.shape != x0.shape or x.dtype != x0.dtype or x.shape[0] != ix.N:
            raise ValueError("mean_mix operands must be two [N, ...] tensors of one shape and dtype")
        with _lib.on_device(x.device):
            out = _mean_mix(ix, x, x0, sa, sb)
            if out is None:
                agg = torch.empty_like(x)
                _aggregate(ix, x, None, True, None, _lib.POST_MEAN, None, None, agg, None, False)
                out = torch.empty_like(x)
                check(lib().rgnn_axpby(dtype_code(x), float(sa), ptr(agg), float(sb), ptr(x0), ptr(out), x.numel(), stream()), "rgnn_axpby")
        ctx.g, ctx.sa, ctx.sb = g, float(sa), float(sb)
        return out

    @staticmethod
    def backward(ctx, d):
        ix = ctx.g.index()
        d = d.contiguous()
        dx = dx0 = None
        with _lib.on_device(d.device):
            if ctx.needs_input_grad[0]:
                dx = _mean_mix(ix, d, None, ctx.sa, 0.0)
                if dx is None:
                    agg = torch.empty_like(d)
                    _aggregate(ix, d, None, True, None, _lib.POST_MEAN, None, None, agg, None, False)
                    dx = torch.empty_like(d)
                    check(lib().rgnn_axpby(dtype_code(d), ctx.sa, ptr(agg), 0.0, ptr(agg), ptr(dx), d.numel(), stream()), "rgnn_axpby")
            if ctx.needs_input_grad[1]:
                dx0 = torch.empty_like(d)
                check(lib().rgnn_axpby(dtype_code(d), ctx.sb, ptr(d), 0.0, ptr(d), ptr(dx0), d.numel(), stream()), "rgnn_axpby")
        return dx, dx0, None, None, None


def mean_mix(g, x, x0, sa, sb):
    return MeanMixFn.apply(x, x0, g, sa, sb)


class BiasAddFn(torch.autograd.Function):
    """y = x + bias over the last dim (GraphConv bias when no weight transform follows the aggregation)."""

    @staticmethod
    def forward(ctx, x, bias):
        _lib.require_cuda(x, bias)
        x = x.contiguous()
        y = torch.empty_like(x)
        bf = bias.detach().float().contiguous()
        with _lib.on_device(x.device):
            check(lib().rgnn_bias_add(dtype_code(x), ptr(x), ptr(bf), ptr(y), x.numel() // bf.numel(), bf.numel(), stream()), "rgnn_bias_add")
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        cols = dy.shape[-1]
        db = None
        if ctx.needs_input_grad[1]:
            db = torch.empty((cols,), dtype=torch.float32, device=dy.device)
            with _lib.on_device(dy.device):
                check(lib().rgnn_reduce_sum(dtype_code(dy), ptr(dy), dy.numel() // cols, cols, cols, 1, 1, 1, ptr(db), stream()),
                      "rgnn_reduce_sum")
        return dy, db


def bias_add(x, bias):
    return BiasAddFn.apply(x, bias)


# ---- seams either side of the GNN slice ---------------------------------------------------------------------------
_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}


def _is_channels_last(x):
    return x.dim() == 4 and x.shape[1] > 1 and x.shape[2] * x.shape[3] > 1 and x.is_contiguous(memory_format=torch.channels_last) \
        and not x.is_contiguous()


class LayoutCastFn(torch.autograd.Function):
    """One-pass memory-format + dtype change of a [N, F, H, W] feature map (rgnn_layout_cast): channels-last <-> NCHW, fp32 <-> bf16.
    to_nchw=True : input physically [N, H*W, F] (channels-last)  -> contiguous NCHW tensor of `dtype`   (encoder -> GNN seam)
    to_nchw=False: input contiguous NCHW                          -> channels-last tensor of `dtype`    (GNN -> decoder seam)
    The backward is the mirror-image call, so a channels-last bf16/fp32 backbone or decoder never sees a torch transpose copy."""

    @staticmethod
    def forward(ctx, x, dtype, to_nchw):
        _lib.require_cuda(x)
        if x.dtype not in _DT or dtype not in _DT:
            raise TypeError("feature maps are float32 or bfloat16")
        n, f, hh, ww = x.shape
        ctx.in_dtype, ctx.to_nchw = x.dtype, to_nchw
        if to_nchw:
            if not x.is_contiguous(memory_format=torch.channels_last):
                raise ValueError("to_nchw expects a channels-last tensor")
            out = torch.empty((n, f, hh, ww), dtype=dtype, device=x.device)
            rows, cols = hh * ww, f
        else:
            x = x.contiguous()
            out = torch.empty((n, f, hh, ww), dtype=dtype, device=x.device, memory_format=torch.channels_last)
            rows, cols = f, hh * ww
        with _lib.on_device(x.device), _lib.timed("seam", x.numel() * (x.element_size() + out.element_size())):
            check(lib().rgnn_layout_cast(_DT[x.dtype], ptr(x), _DT[dtype], ptr(out), n, rows, cols, stream()), "rgnn_layout_cast")
        return out

    @staticmethod
    def backward(ctx, d):
        n, f, hh, ww = d.shape
        if ctx.to_nchw:       # forward produced NCHW from channels-last: the gradient goes back as channels-last in the input's dtype
            d = d.contiguous()
            out = torch.empty((n, f, hh, ww), dtype=ctx.in_dtype, device=d.device, memory_format=torch.channels_last)
            rows, cols = f, hh * ww
        else:
            if not d.is_contiguous(memory_format=torch.channels_last):
                d = d.contiguous(memory_format=torch.channels_last)
            out = torch.empty((n, f, hh, ww), dtype=ctx.in_dtype, device=d.device)
            rows, cols = hh * ww, f
        with _lib.on_device(d.device):
            check(lib().rgnn_layout_cast(_DT[d.dtype], ptr(d), _DT[ctx.in_dtype], ptr(out), n, rows, cols, stream()), "rgnn_layout_cast")
        return out, None, None


def to_gnn_layout(h, dtype=None):
    """Encoder -> GNN seam: whatever the backbone emits (NCHW or channels-last, fp32 or bf16) -> the contiguous NCHW tensor of `dtype`
    the robot-graph kernels consume; a channels-last input is transposed and cast in one pass."""
    dtype = h.dtype if dtype is None else dtype
    if _is_channels_last(h):
        return LayoutCastFn.apply(h, dtype, True)
    h = h.contiguous()
    return h if h.dtype == dtype else h.to(dtype)


def from_gnn_layout(y, dtype=None, channels_last=False):
    """GNN -> decoder seam: the block's NCHW output -> the layout / dtype the decoder's first conv wants, in one pass."""
    dtype = y.dtype if dtype is None else dtype
    if channels_last:
        return LayoutCastFn.apply(y, dtype, False)
    return y if y.dtype == dtype else y.to(dtype)
