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

// Vector variant: thread = one 16-byte vector of features (weights and bias of its VEC features live in registers) x one of EL edge
// lanes; a block covers MLP1_EPB edges whose pose rows sit in shared memory (broadcast reads).  Stores are 16 bytes per thread,
// 512 contiguous bytes per warp - the kernel is a pure HBM writer (E*F outputs from 9 inputs each).
constexpr int MLP1_EPB = 64;
template <typename T>
__global__ void __launch_bounds__(256) k_mlp1_fwd_vec(const float* __restrict__ pose, const float* __restrict__ w1,
                                                      const float* __restrict__ b1, T* __restrict__ a1, int64_t E, int F, int cv) {
  constexpr int VEC = Vec16<T>::N;
  __shared__ float sp[MLP1_EPB][9];
  const int el = blockDim.x / cv;                            // edge lanes
  const int fv = blockIdx.x * cv + threadIdx.x % cv, lane_e = threadIdx.x / cv;
  const int64_t e0 = (int64_t)blockIdx.y * MLP1_EPB;
  for (int i = threadIdx.x; i < MLP1_EPB * 9; i += blockDim.x) {
    const int64_t e = e0 + i / 9;
    sp[i / 9][i % 9] = e < E ? pose[e * 9 + i % 9] : 0.f;
  }
  const bool live = fv * VEC < F;
  float w[VEC][9], b[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) {
#pragma unroll
    for (int j = 0; j < 9; ++j) w[q][j] = live ? w1[(fv * VEC + q) * 9 + j] : 0.f;
    b[q] = live ? b1[fv * VEC + q] : 0.f;
  }
  __syncthreads();
  if (!live) return;
  const int n = (int)min((int64_t)MLP1_EPB, E - e0);
  for (int r = lane_e; r < n; r += el) {
    float o[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 9; ++j) acc = fmaf(sp[r][j], w[q][j], acc);       // same chain as the scalar kernel
      o[q] = fmaxf(acc + b[q], 0.f);
    }
    Vec16<T>::store(a1 + (e0 + r) * F + (int64_t)fv * VEC, o);
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

// Second-generation variant of k_mlp1_bwd_vec: block = cv feature vectors x rg row groups (256 threads), every thread walks the
// rows rg, rg + RG, ... of the block's chunk with four 16-byte loads in flight; the row groups are summed through shared memory
// and ONE thread per feature vector issues the atomics - 4x more bytes in flight per SM and 4-8x fewer atomics per address
// than one 64-thread block per chunk (whose 1184 chunks serialised on the 5120 output addresses).
template <typename T>
__global__ void __launch_bounds__(256) k_mlp1_bwd_vec2(const float* __restrict__ pose, const T* __restrict__ da1,
                                                       float* __restrict__ dw1, float* __restrict__ db1, int64_t E, int F, int64_t chunk, int cv) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int TR = 64;                                     // rows per shared-memory pose tile
  __shared__ float sp[TR][9];
  __shared__ float red[256 * VEC];
  const int rgn = blockDim.x / cv;
  const int c = threadIdx.x % cv, rg = threadIdx.x / cv;
  const int fv = blockIdx.x * cv + c;
  const bool live = fv * VEC < F;
  const int64_t e0 = (int64_t)blockIdx.y * chunk, e1 = min(E, e0 + chunk);
  float acc[VEC][10];
#pragma unroll
  for (int q = 0; q < VEC; ++q)
#pragma unroll
    for (int j = 0; j < 10; ++j) acc[q][j] = 0.f;
  for (int64_t base = e0; base < e1; base += TR) {
    __syncthreads();
    for (int i = threadIdx.x; i < TR * 9; i += blockDim.x) {
      const int64_t e = base + i / 9;
      sp[i / 9][i % 9] = e < e1 ? pose[e * 9 + i % 9] : 0.f;
    }
    __syncthreads();
    if (live) {
      const int n = (int)min((int64_t)TR, e1 - base);
      const T* row = da1 + base * F + (int64_t)fv * VEC;
      int r = rg;
      for (; r + 3 * rgn < n; r += 4 * rgn) {
        uint4 u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = *reinterpret_cast<const uint4*>(row + (int64_t)(r + k * rgn) * F);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float d[VEC];
          Vec16<T>::unpack(u[k], d);
#pragma unroll
          for (int q = 0; q < VEC; ++q) {
#pragma unroll
            for (int j = 0; j < 9; ++j) acc[q][j] = fmaf(d[q], sp[r + k * rgn][j], acc[q][j]);
            acc[q][9] += d[q];
          }
        }
      }
      for (; r < n; r += rgn) {
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
  // sum the row groups (one of the 10 output columns at a time) and let row group 0 publish
#pragma unroll
  for (int j = 0; j < 10; ++j) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < VEC; ++q) red[threadIdx.x * VEC + q] = acc[q][j];
    __syncthreads();
    // red[(g2 * cv + c) * VEC + q] = red[g2 * nout + o] with o = c * VEC + q: one thread per output o sums the row groups
    // (consecutive threads, consecutive addresses), instead of the cv threads of row group 0 summing VEC outputs each
    const int nout = cv * VEC;
    for (int o = threadIdx.x; o < nout; o += blockDim.x) {
      float v = 0.f;
      for (int g2 = 0; g2 < rgn; ++g2) v += red[g2 * nout + o];
      const int f = blockIdx.x * nout + o;
      if (f < F) { if (j < 9) atomicAdd(dw1 + f * 9 + j, v); else atomicAdd(db1 + f, v); }
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
  {
    const int es = dtype == RGNN_F32 ? 4 : 2, vec = 16 / es;
    const int64_t eb = (E + MLP1_EPB - 1) / MLP1_EPB;
    if (F % vec == 0 && (reinterpret_cast<uintptr_t>(a1) & 15u) == 0 && eb <= 65535) {
      const int fvn = (int)(F / vec);
      int cv = 8;                                            // (bf16, 128 hidden features = 16 vectors: no idle half-warps)
      while (cv < fvn && cv < 256) cv <<= 1;                 // feature vectors per block (power of two), 256 / cv edge lanes
      dim3 grid((unsigned)((fvn + cv - 1) / cv), (unsigned)eb);
      if (dtype == RGNN_F32) k_mlp1_fwd_vec<float><<<grid, 256, 0, st>>>(pose, w1, b1, static_cast<float*>(a1), E, (int)F, cv);
      else k_mlp1_fwd_vec<__nv_bfloat16><<<grid, 256, 0, st>>>(pose, w1, b1, static_cast<__nv_bfloat16*>(a1), E, (int)F, cv);
      RGNN_LAUNCH_CHECK("k_mlp1_fwd_vec");
      return 0;
    }
  }
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
  if (vec_ok) {
    const int fvn = (int)(F / vec);
    int cv = 8;                                            // (bf16, 128 hidden features = 16 vectors: no idle half-warps)
    while (cv < fvn && cv < 256) cv <<= 1;
    const int fbx = (fvn + cv - 1) / cv;
    int64_t chunks2 = (4 * num_sms() + fbx - 1) / fbx;          // ~4 blocks of 256 threads per SM
    const int64_t maxc = (E + 63) / 64;
    if (chunks2 > maxc) chunks2 = maxc;
    if (chunks2 < 1) chunks2 = 1;
    if (chunks2 > 65535) chunks2 = 65535;
    const int64_t chunk2 = ((E + chunks2 - 1) / chunks2 + 63) / 64 * 64;
    chunks2 = (E + chunk2 - 1) / chunk2;
    dim3 grid2((unsigned)fbx, (unsigned)chunks2);
    if (dtype == RGNN_F32) k_mlp1_bwd_vec2<float><<<grid2, 256, 0, st>>>(pose, static_cast<const float*>(da1), dw1, db1, E, (int)F, chunk2, cv);
    else k_mlp1_bwd_vec2<__nv_bfloat16><<<grid2, 256, 0, st>>>(pose, static_cast<const __nv_bfloat16*>(da1), dw1, db1, E, (int)F, chunk2, cv);
    RGNN_LAUNCH_CHECK("k_mlp1_bwd_vec2");
    return 0;
  }
  const int fb = vec_ok ? (int)((F / vec + 63) / 64) : (int)((F + 63) / 64);
  int64_t chunks = (8 * num_sms() + fb - 1) / fb;
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
