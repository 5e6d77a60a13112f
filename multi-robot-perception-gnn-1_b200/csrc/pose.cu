// the edge-feature producer.  Replaces the per-sample Python double loop over cal_relative_pose
// (dgl/dataloader.py:116-122, dgl/utils.py:54-77): for every ordered robot pair (i, j), i != j, of every sample,
//   feature = (t_j - t_i) || first 6 entries (column-major) of R(q_j * conj(q_i)),
// in the reference's src-major edge order.  Computed in fp64 with the reference's operation order and no FMA contraction,
// then rounded to fp32 exactly as `torch.from_numpy(...).float()` does - so the output is bit-identical to the reference.
#include "common.cuh"

namespace rgnn {

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }

template <typename TP>
__global__ void __launch_bounds__(256) k_relative_pose(const TP* __restrict__ poses, float* __restrict__ out, int64_t G, int K) {
  const int per = K * (K - 1);
  const int64_t total = G * per;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = e / per;
    const int el = (int)(e - g * per);
    const int i = el / (K - 1), jj = el % (K - 1);
    const int j = jj < i ? jj : jj + 1;
    const TP* p1 = poses + (g * K + i) * 7;
    const TP* p2 = poses + (g * K + j) * 7;
    // conj(q1) and q2, (x, y, z, w)                                                      utils.py:72-73
    const double x1 = -(double)p1[3], y1 = -(double)p1[4], z1 = -(double)p1[5], w1 = (double)p1[6];
    const double x2 = (double)p2[3], y2 = (double)p2[4], z2 = (double)p2[5], w2 = (double)p2[6];
    // q2 * q1^-1 in the reference's left-to-right evaluation order                        utils.py:75-76
    const double x = sub(add(add(mul(w1, x2), mul(x1, w2)), mul(y1, z2)), mul(z1, y2));
    const double y = add(add(sub(mul(w1, y2), mul(x1, z2)), mul(y1, w2)), mul(z1, x2));
    const double z = add(sub(add(mul(w1, z2), mul(x1, y2)), mul(y1, x2)), mul(z1, w2));
    const double w = sub(sub(sub(mul(w1, w2), mul(x1, x2)), mul(y1, y2)), mul(z1, z2));
    // quat_to_so3, first six entries a00 a10 a20 a01 a11 a21                               utils.py:54-66
    const double xx = mul(x, x), yy = mul(y, y), zz = mul(z, z);
    float* o = out + e * 9;
    o[0] = (float)sub((double)p2[0], (double)p1[0]);
    o[1] = (float)sub((double)p2[1], (double)p1[1]);
    o[2] = (float)sub((double)p2[2], (double)p1[2]);
    o[3] = (float)sub(sub(1.0, mul(2.0, yy)), mul(2.0, zz));
    o[4] = (float)add(mul(mul(2.0, x), y), mul(mul(2.0, z), w));
    o[5] = (float)sub(mul(mul(2.0, x), z), mul(mul(2.0, y), w));
    o[6] = (float)sub(mul(mul(2.0, x), y), mul(mul(2.0, z), w));
    o[7] = (float)sub(sub(1.0, mul(2.0, xx)), mul(2.0, zz));
    o[8] = (float)add(mul(mul(2.0, y), z), mul(mul(2.0, x), w));
  }
}

}  // namespace rgnn

using namespace rgnn;

extern "C" int rgnn_relative_pose(int pose_is_f64, const void* poses, float* edge_pose, int64_t num_graphs, int64_t robots, void* stream) {
  RGNN_REQUIRE(num_graphs >= 0 && robots >= 1 && robots < (1 << 15), RGNN_E_ARG, "rgnn_relative_pose: bad sizes");
  const int64_t total = num_graphs * robots * (robots - 1);
  if (total == 0) return 0;
  RGNN_REQUIRE(poses && edge_pose, RGNN_E_ARG, "rgnn_relative_pose: null pointer");
  int64_t blocks = (total + 255) / 256;
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  cudaStream_t st = as_stream(stream);
  if (pose_is_f64) k_relative_pose<double><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const double*>(poses), edge_pose, num_graphs, (int)robots);
  else k_relative_pose<float><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const float*>(poses), edge_pose, num_graphs, (int)robots);
  RGNN_LAUNCH_CHECK("k_relative_pose");
  return 0;
}


// ---- section 8(f) row 2: the seams either side of the GNN slice --------------------------------------------------------------
// A backbone run in channels-last (NHWC: what cuDNN's tensor-core convolutions want) hands over [N, H*W, F] rows; the robot-graph
// kernels consume NCHW [N, F, H*W] (a robot's [64 ch x 64 px] slab is one MN-major tensor-core operand block); the decoder's first
// conv (TransBlock 5x5, dgl/model/blocks.py:16) again prefers channels-last.  One pass per seam: tiled transpose with the dtype
// conversion fused (fp32 <-> bf16), instead of torch's .contiguous() followed by .to(dtype) (two passes each way).
namespace rgnn {
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) k_layout_cast(const TI* __restrict__ x, TO* __restrict__ out, int64_t nodes, int R, int C) {
  // per node: in [R, C] (C contiguous) -> out [C, R] (R contiguous); 64 x 64 tiles through shared memory, coalesced on both sides
  __shared__ float t[64][65];
  const int tiles_r = (R + 63) / 64, tiles_c = (C + 63) / 64;
  const int64_t total = nodes * tiles_r * tiles_c;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;       // 64 x 4
  for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
    const int64_t v = w / (tiles_r * tiles_c);
    const int rem = (int)(w - v * tiles_r * tiles_c);
    const int r0 = (rem / tiles_c) * 64, c0 = (rem % tiles_c) * 64;
    const TI* xi = x + v * (int64_t)R * C;
    TO* oo = out + v * (int64_t)R * C;
#pragma unroll 4
    for (int j = ty; j < 64; j += 4) {
      const int r = r0 + j, c = c0 + tx;
      t[j][tx] = (r < R && c < C) ? to_f32<TI>(xi[(int64_t)r * C + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int j = ty; j < 64; j += 4) {
      const int c = c0 + j, r = r0 + tx;
      if (r < R && c < C) oo[(int64_t)c * R + r] = from_f32<TO>(t[tx][j]);
    }
    __syncthreads();
  }
}
}  // namespace rgnn

// out[v, c, r] = (out_dtype) x[v, r, c]: per-node transpose [rows, cols] -> [cols, rows] with dtype conversion.
// channels-last -> NCHW: rows = H*W, cols = F;  NCHW -> channels-last: rows = F, cols = H*W.
extern "C" int rgnn_layout_cast(int in_dtype, const void* x, int out_dtype, void* out, int64_t nodes, int64_t rows, int64_t cols, void* stream) {
  using namespace rgnn;
  RGNN_REQUIRE((in_dtype == RGNN_F32 || in_dtype == RGNN_BF16) && (out_dtype == RGNN_F32 || out_dtype == RGNN_BF16), RGNN_E_ARG,
               "rgnn_layout_cast: bad dtype");
  RGNN_REQUIRE(nodes >= 0 && rows > 0 && cols > 0 && rows < (1 << 30) && cols < (1 << 30), RGNN_E_ARG, "rgnn_layout_cast: bad sizes");
  if (nodes == 0) return 0;
  RGNN_REQUIRE(x && out, RGNN_E_ARG, "rgnn_layout_cast: null pointer");
  const int64_t tiles = nodes * ((rows + 63) / 64) * ((cols + 63) / 64);
  int64_t blocks = tiles < 16 * num_sms() ? tiles : 16 * num_sms();
  cudaStream_t st = as_stream(stream);
  const int R = (int)rows, C = (int)cols;
  if (in_dtype == RGNN_F32 && out_dtype == RGNN_F32)
    k_layout_cast<float, float><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const float*>(x), static_cast<float*>(out), nodes, R, C);
  else if (in_dtype == RGNN_F32)
    k_layout_cast<float, __nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const float*>(x), static_cast<__nv_bfloat16*>(out), nodes, R, C);
  else if (out_dtype == RGNN_F32)
    k_layout_cast<__nv_bfloat16, float><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<float*>(out), nodes, R, C);
  else
    k_layout_cast<__nv_bfloat16, __nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), nodes, R, C);
  RGNN_LAUNCH_CHECK("k_layout_cast");
  return 0;
}
