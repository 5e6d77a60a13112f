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
