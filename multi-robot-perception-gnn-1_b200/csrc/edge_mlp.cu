// A4 (+A10): first layer of the FiLM edge encoder, nn.Linear(9,F)+ReLU (dgl/model/models.py:147-148) and its
// weight gradients.  K = 9 is far too thin for a tensor-core tile: a1 is an HBM-bound elementwise producer
// (E*F outputs from 9 inputs each), the backward a column reduction over edges.
#include "common.cuh"

namespace rgnn {

template <typename T>
__global__ void __launch_bounds__(256) k_mlp1_fwd(const float* __restrict__ pose, const float* __restrict__ w1,
                                                  const float* __restrict__ b1, T* __restrict__ a1, int64_t E, int F) {
  // block: 8 edges x F features; pose rows staged in shared memory
  __shared__ float sp[8][9];
  const int64_t e0 = (int64_t)blockIdx.x * 8;
  if (threadIdx.x < 72) {
    const int r = threadIdx.x / 9, j = threadIdx.x % 9;
    sp[r][j] = (e0 + r < E) ? pose[(e0 + r) * 9 + j] : 0.f;
  }
  __syncthreads();
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float w[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) w[j] = w1[f * 9 + j];
    const float b = b1[f];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (e0 + r < E) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 9; ++j) acc = fmaf(sp[r][j], w[j], acc);
        acc += b;
        a1[(e0 + r) * F + f] = from_f32<T>(fmaxf(acc, 0.f));
      }
    }
  }
}

// dw1[f,j] = sum_e da1[e,f] pose[e,j]; db1[f] = sum_e da1[e,f].  Thread = 16-byte vector of features (coalesced 128-bit row reads),
// streams its chunk of edges with the 9 pose values of each edge broadcast from shared memory; fp32 atomics at the end
// (outputs zeroed by the caller wrapper).  grid = (ceil(F/VEC/64), chunks over E).
template <typename T>
__global__ void __launch_bounds__(64) k_mlp1_bwd_vec(const float* __restrict__ pose, const T* __restrict__ da1,
                                                     float* __restrict__ dw1, float* __restrict__ db1, int64_t E, int F, int64_t chunk) {
  constexpr int VEC = Vec16<T>::N;
  const int fv = blockIdx.x * 64 + threadIdx.x;            // feature-vector index
  const bool live = fv * VEC < F;
  const int64_t e0 = (int64_t)blockIdx.y * chunk, e1 = min(E, e0 + chunk);
  __shared__ float sp[64][9];
  float acc[VEC][10];
#pragma unroll
  for (int q = 0; q < VEC; ++q)
#pragma unroll
    for (int j = 0; j < 10; ++j) acc[q][j] = 0.f;
  for (int64_t base = e0; base < e1; base += 64) {
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 9; i += 64) {
      const int64_t e = base + i / 9;
      sp[i / 9][i % 9] = e < e1 ? pose[e * 9 + i % 9] : 0.f;
    }
    __syncthreads();
    if (live) {
      const int n = (int)min((int64_t)64, e1 - base);
      const T* row = da1 + base * F + (int64_t)fv * VEC;
#pragma unroll 2
      for (int r = 0; r < n; ++r) {
        float d[VEC];
        Vec16<T>::load(row + (int64_t)r * F, d);
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
#pragma unroll
          for (int j = 0; j < 9; ++j) acc[q][j] = fmaf(d[q], sp[r][j], acc[q][j]);
          acc[q][9] += d[q];
        }
      }
    }
  }
  if (live) {
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      const int f = fv * VEC + q;
#pragma unroll
      for (int j = 0; j < 9; ++j) atomicAdd(dw1 + f * 9 + j, acc[q][j]);
      atomicAdd(db1 + f, acc[q][9]);
    }
  }
}

// scalar variant for feature counts that are not a multiple of the 16-byte vector
template <typename T>
__global__ void __launch_bounds__(64) k_mlp1_bwd(const float* __restrict__ pose, const T* __restrict__ da1,
                                                 float* __restrict__ dw1, float* __restrict__ db1, int64_t E, int F, int64_t chunk) {
  const int f = blockIdx.x * 64 + threadIdx.x;
  const int64_t e0 = (int64_t)blockIdx.y * chunk, e1 = min(E, e0 + chunk);
  __shared__ float sp[64][9];
  float acc[10];
#pragma unroll
  for (int j = 0; j < 10; ++j) acc[j] = 0.f;
  for (int64_t base = e0; base < e1; base += 64) {
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 9; i += 64) {
      const int64_t e = base + i / 9;
      sp[i / 9][i % 9] = e < e1 ? pose[e * 9 + i % 9] : 0.f;
    }
    __syncthreads();
    if (f < F) {
      const int n = (int)min((int64_t)64, e1 - base);
      for (int r = 0; r < n; ++r) {
        const float d = to_f32<T>(da1[(base + r) * F + f]);
#pragma unroll
        for (int j = 0; j < 9; ++j) acc[j] = fmaf(d, sp[r][j], acc[j]);
        acc[9] += d;
      }
    }
  }
  if (f < F) {
#pragma unroll
    for (int j = 0; j < 9; ++j) atomicAdd(dw1 + f * 9 + j, acc[j]);
    atomicAdd(db1 + f, acc[9]);
  }
}

}  // namespace rgnn

using namespace rgnn;

extern "C" int rgnn_edge_mlp1_fwd(int dtype, const float* pose, const float* w1, const float* b1, void* a1,
                                  int64_t E, int64_t F, void* stream) {
  RGNN_REQUIRE(dtype == RGNN_F32 || dtype == RGNN_BF16, RGNN_E_ARG, "rgnn_edge_mlp1_fwd: bad dtype %d", dtype);
  RGNN_REQUIRE(E >= 0 && F > 0 && F < (1 << 28), RGNN_E_ARG, "rgnn_edge_mlp1_fwd: bad sizes");
  if (E == 0) return 0;
  RGNN_REQUIRE(pose && w1 && b1 && a1, RGNN_E_ARG, "rgnn_edge_mlp1_fwd: null pointer");
  const int64_t blocks = (E + 7) / 8;
  RGNN_REQUIRE(blocks < (1ll << 31), RGNN_E_RANGE, "rgnn_edge_mlp1_fwd: too many edges");
  cudaStream_t st = as_stream(stream);
  if (dtype == RGNN_F32) k_mlp1_fwd<float><<<(unsigned)blocks, 256, 0, st>>>(pose, w1, b1, static_cast<float*>(a1), E, (int)F);
  else k_mlp1_fwd<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(pose, w1, b1, static_cast<__nv_bfloat16*>(a1), E, (int)F);
  RGNN_LAUNCH_CHECK("k_mlp1_fwd");
  return 0;
}

extern "C" int rgnn_edge_mlp1_bwd(int dtype, const float* pose, const void* da1, float* dw1, float* db1,
                                  int64_t E, int64_t F, void* stream) {
  RGNN_REQUIRE(dtype == RGNN_F32 || dtype == RGNN_BF16, RGNN_E_ARG, "rgnn_edge_mlp1_bwd: bad dtype %d", dtype);
  RGNN_REQUIRE(E >= 0 && F > 0 && F < (1 << 28), RGNN_E_ARG, "rgnn_edge_mlp1_bwd: bad sizes");
  RGNN_REQUIRE(dw1 && db1, RGNN_E_ARG, "rgnn_edge_mlp1_bwd: null output");
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(dw1, 0, sizeof(float) * 9 * F, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(db1, 0, sizeof(float) * F, st);
  if (e != cudaSuccess) return cuda_fail(e, "memset dw1/db1");
  if (E == 0) return 0;
  RGNN_REQUIRE(pose && da1, RGNN_E_ARG, "rgnn_edge_mlp1_bwd: null input");
  const int es = dtype == RGNN_F32 ? 4 : 2, vec = 16 / es;
  const bool vec_ok = F % vec == 0 && (reinterpret_cast<uintptr_t>(da1) & 15u) == 0;
  const int fb = vec_ok ? (int)((F / vec + 63) / 64) : (int)((F + 63) / 64);
  int64_t chunks = (8 * kNumSMs + fb - 1) / fb;
  const int64_t max_chunks = (E + 63) / 64;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  if (chunks > 65535) chunks = 65535;
  const int64_t chunk = ((E + chunks - 1) / chunks + 63) / 64 * 64;
  chunks = (E + chunk - 1) / chunk;
  dim3 grid(fb, (unsigned)chunks);
  if (vec_ok) {
    if (dtype == RGNN_F32) k_mlp1_bwd_vec<float><<<grid, 64, 0, st>>>(pose, static_cast<const float*>(da1), dw1, db1, E, (int)F, chunk);
    else k_mlp1_bwd_vec<__nv_bfloat16><<<grid, 64, 0, st>>>(pose, static_cast<const __nv_bfloat16*>(da1), dw1, db1, E, (int)F, chunk);
  } else {
    if (dtype == RGNN_F32) k_mlp1_bwd<float><<<grid, 64, 0, st>>>(pose, static_cast<const float*>(da1), dw1, db1, E, (int)F, chunk);
    else k_mlp1_bwd<__nv_bfloat16><<<grid, 64, 0, st>>>(pose, static_cast<const __nv_bfloat16*>(da1), dw1, db1, E, (int)F, chunk);
  }
  RGNN_LAUNCH_CHECK("k_mlp1_bwd");
  return 0;
}
