"""ctypes binding of librobotgnn.so (include/robot_gnn.h).  No fallback: a missing library or a failing
call raises - the product path never computes on the CPU or through another backend."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RGNN_LIB") or os.path.join(_HERE, "librobotgnn.so")      # RGNN_LIB: an experimental build (development)

F32, BF16 = 0, 1
POST_NONE, POST_MEAN, POST_SCALE = 0, 1, 2
EPI_BIAS_M, EPI_BIAS_N, EPI_SIGMOID, EPI_RELU_MASK, EPI_ADD_AUX = 1, 2, 4, 8, 16

_i64, _i32, _f32, _vp = ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_void_p

# name -> argtypes; must list every symbol declared in include/robot_gnn.h (tests/test_cabi.py checks)
SIGNATURES = {
    "rgnn_version": [],
    "rgnn_arch": [],
    "rgnn_device_ok": [],
    "rgnn_batch_graphs": [_vp] * 4 + [_i64] * 2 + [_vp] * 5,
    "rgnn_build_batched_csr": [_vp, _vp, _i64, _i64, _vp, _vp, _i64] + [_vp] * 13,
    "rgnn_aggregate": [_i32, _vp, _i64, _vp, _i32, _vp, _i32, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp]
                      + [_i64] * 8 + [_vp],
    "rgnn_check_complete": [_vp, _vp, _i64, _i64, _i64, _vp, _vp],
    "rgnn_relative_pose": [_i32, _vp, _vp, _i64, _i64, _vp],
    "rgnn_film_edge_grad": [_i32, _vp, _i64, _vp, _i64, _vp, _i32, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp],
    "rgnn_film_backward": [_i32, _vp, _i64, _vp, _i64, _vp, _i32, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp] + [_i64] * 8 + [_vp],
    "rgnn_film_bucket_forward": [_i32, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _i64, _vp],
    "rgnn_film_bucket_backward": [_i32, _vp, _i64, _vp, _i64, _vp, _i32, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _i64, _vp],
    "rgnn_edge_mlp1_fwd": [_i32, _vp, _vp, _vp, _vp, _i64, _i64, _vp],
    "rgnn_edge_mlp1_bwd": [_i32, _vp, _vp, _vp, _vp, _i64, _i64, _vp],
    "rgnn_gemm": [_i32, _vp, _i64, _i64, _i64, _i64, _i32, _vp, _i64, _i64, _i64, _i64, _i32, _vp, _i64, _i64, _i64]
                 + [_i64] * 6 + [_vp, _i32, _vp, _vp, _f32, _f32, _vp],
    "rgnn_tc_cast_weights": [_vp, _i64, _i64, _vp, _vp, _vp],
    "rgnn_tc_cast_weights_affine": [_vp, _i64, _i64, _f32, _f32, _vp, _vp, _vp],
    "rgnn_tc_conv1x1": [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp],
    "rgnn_tc_conv_film_fused": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp],
    "rgnn_tc_conv_dx_film_bwd": [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp],
    "rgnn_tc_conv1x1_wgrad": [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _i64, _vp],
    "rgnn_tc_gemm_tn": [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _i32, _vp, _vp, _f32, _f32, _vp],
    "rgnn_tc_gemm_nn": [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _f32, _vp, _i64, _vp],
    "rgnn_tc_split3": [_vp, _i64, _i64, _vp, _vp],
    "rgnn_tc_split3_weights": [_vp, _i64, _i64, _vp, _vp, _vp],
    "rgnn_tc3_conv1x1": [_vp, _vp, _i64, _vp, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp],
    "rgnn_tc3_conv1x1_wgrad": [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _vp, _i64, _vp, _i64, _i64, _vp],
    "rgnn_tc3_gemm_tn": [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _i64, _i64, _i32, _vp, _f32, _vp],
    "rgnn_tc3_gemm_nn": [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _i64, _i64, _f32, _vp, _i64, _vp],
    "rgnn_layout_cast": [_i32, _vp, _i32, _vp, _i64, _i64, _i64, _vp],
    "rgnn_relu_grad": [_i32, _vp, _vp, _vp, _i64, _vp],
    "rgnn_sigmoid_grad": [_i32, _vp, _vp, _vp, _i64, _vp],
    "rgnn_axpby": [_i32, _f32, _vp, _f32, _vp, _vp, _i64, _vp],
    "rgnn_complete_mean_mix": [_i32, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _f32, _f32, _vp],
    "rgnn_bias_add": [_i32, _vp, _vp, _vp, _i64, _i64, _vp],
    "rgnn_reduce_sum": [_i32, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp],
}

_lib = None
USE_FUSED = False  # FiLM aggregation fused into the compress GEMM's prologue (rgnn_tc_conv_film_fused; complete graphs of 2..8 robots,
                   # bf16, P = 64).  Correct and tested, but measured SLOWER than aggregation kernel + GEMM at F = 512 (2.13 vs 1.71 ms
                   # with the aggregate stored for the backward, 1.74 vs 1.71 ms in eval), so it is opt-in.
USE_FUSED_BWD = os.environ.get("RGNN_FUSED_BWD", "1") != "0"   # FiLM backward in the epilogue of the compress conv's dX GEMM (rgnn_tc_conv_dx_film_bwd): dcat never reaches HBM.
                       # bf16, P = 64, complete graphs of 2..6 robots; anything else runs dX GEMM + FiLM backward kernel.
FUSED_BWD_MAX_F = int(os.environ.get("RGNN_FUSED_BWD_MAXF", "896"))   # widest layer that takes the fused backward: measured on B200 it wins at
                       # F = 512 (10.22 vs 11.06 ms/step) and F = 768 (20.07 vs 21.39 ms) and loses at F = 1280 (26.37 vs 25.42 ms) and F = 2048
                       # (30.75 vs 28.55 ms), where the CTA-pair dX GEMM (1.5 PFLOP/s) dominates the pair
USE_TC3 = True     # fp32 (parity mode) transforms as split-bf16 products on the tcgen05 path when the shape allows (rgnn_tc3_*); False:
                   # the CUDA-core FFMA GEMMs (rgnn_gemm)
USE_TC = True      # bf16 transforms on the tcgen05 path when the shape allows (tests flip it to compare with rgnn_gemm)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"librobotgnn.so not found at {LIB_PATH}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU / PyTorch fallback for the robot-graph layers.")
        l = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = ctypes.c_int
        l.rgnn_launch_count.argtypes = []
        l.rgnn_launch_count.restype = ctypes.c_longlong
        l.rgnn_last_error.argtypes = []
        l.rgnn_last_error.restype = ctypes.c_char_p
        _lib = l
    return _lib


# ---- optional per-call CUDA-event timing (bench.py turns it on around its timed region) ------------------
PROFILE = None      # None = off; else dict: name -> list of (start_event, end_event, work)


class _Timed:
    __slots__ = ("name", "work", "nbytes", "s")

    def __init__(self, name, work, nbytes):
        self.name, self.work, self.nbytes = name, work, nbytes      # nbytes: algorithmic HBM bytes of a kernel whose `work` is flops

    def __enter__(self):
        self.s = torch.cuda.Event(enable_timing=True)
        self.s.record()
        return self

    def __exit__(self, *a):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        if PROFILE is not None:
            PROFILE.setdefault(self.name, []).append((self.s, e, self.work, self.nbytes))
        return False


class _NoTimer:
    __slots__ = ()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NO_TIMER = _NoTimer()


def timed(name, work=0.0, nbytes=0.0):
    """with timed("name", work): <C-ABI call>   - records CUDA events on the current stream when PROFILE is on; otherwise one shared
    do-nothing context (no object per call: the script batch is host-bound)."""
    return _NO_TIMER if PROFILE is None else _Timed(name, work, nbytes)


def profile_summary():
    """name -> dict(calls, ms_total, ms_avg, work_total) after a torch.cuda.synchronize()."""
    out = {}
    for name, evs in (PROFILE or {}).items():
        ms = [ev[0].elapsed_time(ev[1]) for ev in evs]
        out[name] = dict(calls=len(ms), ms_total=sum(ms), ms_avg=sum(ms) / max(1, len(ms)), work_total=sum(ev[2] for ev in evs),
                         bytes_total=sum(ev[3] for ev in evs))
    return out


def check(rc, what):
    if rc != 0:
        msg = lib().rgnn_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    # a plain int converts to the c_void_p parameter by itself; building the ctypes object first costs as much as the call
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_cur_device = getattr(torch._C, "_cuda_getDevice", None)


def stream():
    """torch's current stream on the current device as a cudaStream_t.  The raw C accessors are ~20x cheaper than
    torch.cuda.current_stream().cuda_stream, and at the reference's batch (8 graphs) the step is host-bound: 30 such lookups."""
    if _raw_stream is not None and _cur_device is not None:
        return ctypes.c_void_p(_raw_stream(_cur_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class on_device:
    """`with on_device(t.device):` - torch.cuda.device() without its cost when the device is already current (the usual case:
    one process per GPU).  The launch-bound script batch enters this a dozen times per step."""
    __slots__ = ("idx", "prev")

    def __init__(self, device):
        self.idx = device.index

    def __enter__(self):
        self.prev = _cur_device() if _cur_device is not None else torch.cuda.current_device()
        if self.idx is not None and self.idx != self.prev:
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.idx is not None and self.idx != self.prev:
            torch.cuda.set_device(self.prev)
        return False


def dtype_code(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"robot-graph kernels support float32 and bfloat16, got {t.dtype}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("robot-graph layers run on CUDA (sm_100a) only - there is no CPU fallback; "
                               "move the graph and its features to a B200 device first")
