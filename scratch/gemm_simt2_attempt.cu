// A4/A8 (+A10, A11): strided GEMM family on CUDA cores with fp32-exact products - the PARITY path
// (fp32 1e-5 cannot be met by a single-pass bf16/tf32 MMA on the K = 2F contraction) and the
// general-shape fallback of the tcgen05 path in gemm_tc.cu.
// Tile 64x64x16, 256 threads, 4x4 register micro-tile, operands converted to fp32 in shared memory.
#include <stdlib.h>
#include "common.cuh"

namespace rgnn {

struct GemmParams {
  const void* A; int64_t sAz, sAr, sAm, sAk;
  const void* B; int64_t sBz, sBr, sBk, sBn;
  void* C; int64_t sCz, sCm, sCn;
  int64_t M, N, K, batch, red, split;
  float* partial;
  int epi; const float* bias; const void* aux; float alpha, beta;
  int c_dtype;
  int64_t tiles_m, tiles_n;
  int fold;            // k_gemm_simt2 only: 2 = one 128-column tile spans two batch items of N = 64 columns (per-node conv products)
};

constexpr int BM = 64, BN = 64, BK = 16;

__device__ __forceinline__ float epilogue(int epi, float acc, int64_t m, int64_t n, const float* bias, float auxv, float alpha, float beta) {
  float v = alpha * acc;
  if (epi & RGNN_EPI_ADD_AUX) v = fmaf(beta, auxv, v);
  if (epi & RGNN_EPI_BIAS_M) v += bias[m];
  if (epi & RGNN_EPI_BIAS_N) v += bias[n];
  if (epi & RGNN_EPI_SIGMOID) v = 1.0f / (1.0f + expf(-v));
  if (epi & RGNN_EPI_RELU_MASK) v = auxv > 0.f ? v : 0.f;
  return v;
}

__device__ __forceinline__ float load_aux(int c_dtype, const void* aux, int64_t idx) {
  if (!aux) return 0.f;
  return c_dtype == RGNN_F32 ? static_cast<const float*>(aux)[idx] : __bfloat162float(static_cast<const __nv_bfloat16*>(aux)[idx]);
}
__device__ __forceinline__ void store_c(int c_dtype, void* C, int64_t idx, float v) {
  if (c_dtype == RGNN_F32) static_cast<float*>(C)[idx] = v;
  else static_cast<__nv_bfloat16*>(C)[idx] = __float2bfloat16_rn(v);
}

template <typename TA, typename TB>
__global__ void __launch_bounds__(256) k_gemm_simt(GemmParams p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  int64_t bid = blockIdx.x;
  const int64_t tn = bid % p.tiles_n; bid /= p.tiles_n;
  const int64_t tm = bid % p.tiles_m; bid /= p.tiles_m;
  const int64_t zs = bid;                       // batch index, or split index when split > 1
  const int64_t z = p.split > 1 ? 0 : zs;
  int64_t r0 = 0, r1 = p.red, kbeg = 0, kend = p.K;
  if (p.split > 1) {
    if (p.red > 1) {   // cut the reduction-batch loop
      const int64_t chunk = (p.red + p.split - 1) / p.split;
      r0 = zs * chunk; r1 = min(p.red, r0 + chunk);
    } else {           // cut K itself (in multiples of the k-tile)
      const int64_t chunk = ((p.K + p.split - 1) / p.split + BK - 1) / BK * BK;
      kbeg = min(p.K, zs * chunk); kend = min(p.K, kbeg + chunk);
    }
  }
  const int64_t m0 = tm * BM, n0 = tn * BN;
  const TA* __restrict__ A = static_cast<const TA*>(p.A) + z * p.sAz;
  const TB* __restrict__ B = static_cast<const TB*>(p.B) + z * p.sBz;
  const bool a_kfast = p.sAk == 1, b_nfast = p.sBn == 1;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t r = r0; r < r1; ++r) {
    const TA* Ar = A + r * p.sAr;
    const TB* Br = B + r * p.sBr;
    for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int i = tid + it * 256;
        int am, ak;
        if (a_kfast) { ak = i & 15; am = i >> 4; } else { am = i & 63; ak = i >> 6; }
        const int64_t gm = m0 + am, gk = k0 + ak;
        As[ak][am] = (gm < p.M && gk < kend) ? to_f32<TA>(Ar[gm * p.sAm + gk * p.sAk]) : 0.f;
        int bn, bk;
        if (b_nfast) { bn = i & 63; bk = i >> 6; } else { bk = i & 15; bn = i >> 4; }
        const int64_t gn = n0 + bn, gk2 = k0 + bk;
        Bs[bk][bn] = (gn < p.N && gk2 < kend) ? to_f32<TB>(Br[gk2 * p.sBk + gn * p.sBn]) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  if (p.split > 1) {
    float* P = p.partial + zs * p.M * p.N;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t m = m0 + ty * 4 + i;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t n = n0 + tx * 4 + j;
        if (m < p.M && n < p.N) P[m * p.N + n] = acc[i][j];
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (m < p.M && n < p.N) {
        const int64_t idx = z * p.sCz + m * p.sCm + n * p.sCn;
        store_c(p.c_dtype, p.C, idx, epilogue(p.epi, acc[i][j], m, n, p.bias, load_aux(p.c_dtype, p.aux, idx), p.alpha, p.beta));
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Second-generation CUDA-core kernel for the large fp32 products (same arithmetic as k_gemm_simt: one fmaf chain per output in
// ascending k, so results are bit-identical): 128 x BN2 x 16 tiles, 8 x (BN2/16) register micro-tile, packed FFMA2 (two fp32 FMAs per
// issue slot; the A operand is stored duplicated in shared memory so that (a, a) pairs load directly), double-buffered shared
// memory with the next tile's global loads in flight while the current one is multiplied.
constexpr int BM2 = 128, BK2 = 16;
constexpr int AS_LD = BM2 + 2;                 // float2 per k row (+2: 16-byte aligned rows, k-fast stores spread over banks)

template <typename TA, typename TB, int BN2>
__global__ void __launch_bounds__(256, 2) k_gemm_simt2(GemmParams p) {
  constexpr int TNV = BN2 / 64;                // float4 column chunks per thread (chunk c covers columns c*64 + tx*4 .. +3)
  constexpr int BS_LD = BN2 + 4;
  constexpr int NB = BN2 * BK2 / 256;          // B elements per thread per tile
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* As = reinterpret_cast<float2*>(smem_raw);                               // [2][BK2][AS_LD]
  float* Bs = reinterpret_cast<float*>(smem_raw + 2 * BK2 * AS_LD * sizeof(float2));   // [2][BK2][BS_LD]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  int64_t bid = blockIdx.x;
  const int64_t tn = bid % p.tiles_n; bid /= p.tiles_n;
  const int64_t tm = bid % p.tiles_m; bid /= p.tiles_m;
  const int64_t zs = bid;
  const int64_t z = p.split > 1 ? 0 : zs * p.fold;
  int64_t r0 = 0, r1 = p.red, kbeg = 0, kend = p.K;
  if (p.split > 1) {
    if (p.red > 1) {
      const int64_t chunk = (p.red + p.split - 1) / p.split;
      r0 = zs * chunk; r1 = min(p.red, r0 + chunk);
    } else {
      const int64_t chunk = ((p.K + p.split - 1) / p.split + BK - 1) / BK * BK;
      kbeg = min(p.K, zs * chunk); kend = min(p.K, kbeg + chunk);
    }
  }
  const int64_t m0 = tm * BM2, n0 = tn * BN2;
  const TA* __restrict__ A = static_cast<const TA*>(p.A) + z * p.sAz;
  const TB* __restrict__ B = static_cast<const TB*>(p.B) + z * p.sBz;
  const bool a_kfast = p.sAk == 1, b_nfast = p.sBn == 1;
  const bool fold = p.fold == 2;
  const int64_t tpr = (kend - kbeg + BK2 - 1) / BK2;          // k tiles per reduction item
  const int64_t ntiles = (r1 > r0 && kend > kbeg) ? (r1 - r0) * tpr : 0;

  float2 acc[8][2 * TNV];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 2 * TNV; ++j) acc[i][j] = make_float2(0.f, 0.f);

  float ra[8], rb[NB];
  auto gload = [&](int64_t t) {
    const int64_t r = r0 + t / tpr, k0 = kbeg + (t % tpr) * BK2;
    const TA* Ar = A + r * p.sAr;
    const TB* Br = B + r * p.sBr;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int i = tid + it * 256;
      int am, ak;
      if (a_kfast) { ak = i & 15; am = i >> 4; } else { am = i & 127; ak = i >> 7; }
      const int64_t gm = m0 + am, gk = k0 + ak;
      ra[it] = (gm < p.M && gk < kend) ? to_f32<TA>(Ar[gm * p.sAm + gk * p.sAk]) : 0.f;
    }
#pragma unroll
    for (int it = 0; it < NB; ++it) {
      const int i = tid + it * 256;
      int bn, bk;
      if (b_nfast) { bn = i % BN2; bk = i / BN2; } else { bk = i & 15; bn = i >> 4; }
      const int64_t gk = k0 + bk;
      int64_t gn = n0 + bn, zo = 0;
      if (fold) { zo = bn >> 6; gn = bn & 63; }
      rb[it] = (gn < p.N && gk < kend && (!fold || z + zo < p.batch)) ? to_f32<TB>(Br[zo * p.sBz + gk * p.sBk + gn * p.sBn]) : 0.f;
    }
  };
  auto sstore = [&](int buf) {
    float2* as = As + buf * BK2 * AS_LD;
    float* bs = Bs + buf * BK2 * BS_LD;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int i = tid + it * 256;
      int am, ak;
      if (a_kfast) { ak = i & 15; am = i >> 4; } else { am = i & 127; ak = i >> 7; }
      as[ak * AS_LD + am] = make_float2(ra[it], ra[it]);
    }
#pragma unroll
    for (int it = 0; it < NB; ++it) {
      const int i = tid + it * 256;
      int bn, bk;
      if (b_nfast) { bn = i % BN2; bk = i / BN2; } else { bk = i & 15; bn = i >> 4; }
      bs[bk * BS_LD + bn] = rb[it];
    }
  };

  if (ntiles > 0) { gload(0); sstore(0); }
  __syncthreads();
  for (int64_t t = 0; t < ntiles; ++t) {
    const int buf = (int)(t & 1);
    if (t + 1 < ntiles) gload(t + 1);
    const float2* as = As + buf * BK2 * AS_LD + ty * 8;
    const float* bs = Bs + buf * BK2 * BS_LD + tx * 4;
#pragma unroll
    for (int kk = 0; kk < BK2; ++kk) {
      float2 a2[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(as + kk * AS_LD + 2 * i);
        a2[2 * i] = make_float2(v.x, v.y); a2[2 * i + 1] = make_float2(v.z, v.w);
      }
      float2 b2[2 * TNV];
#pragma unroll
      for (int c = 0; c < TNV; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(bs + kk * BS_LD + c * 64);
        b2[2 * c] = make_float2(v.x, v.y); b2[2 * c + 1] = make_float2(v.z, v.w);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 2 * TNV; ++j) acc[i][j] = ffma2(a2[i], b2[j], acc[i][j]);
    }
    if (t + 1 < ntiles) sstore(buf ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty * 8 + i;
#pragma unroll
    for (int j = 0; j < 4 * TNV; ++j) {
      const int col = (j >> 2) * 64 + tx * 4 + (j & 3);           // column inside the tile
      const float v = (j & 1) ? acc[i][j >> 1].y : acc[i][j >> 1].x;
      int64_t n = n0 + col, zz = z;
      if (fold) { zz = z + (col >> 6); n = col & 63; }
      if (m < p.M && n < p.N && (!fold || zz < p.batch)) {
        if (p.split > 1) {
          p.partial[(zs * p.M + m) * p.N + n] = v;
        } else {
          const int64_t idx = zz * p.sCz + m * p.sCm + n * p.sCn;
          store_c(p.c_dtype, p.C, idx, epilogue(p.epi, v, m, n, p.bias, load_aux(p.c_dtype, p.aux, idx), p.alpha, p.beta));
        }
      }
    }
  }
}

template <typename TA, typename TB>
static int launch_simt2(GemmParams& p, cudaStream_t st) {
  const bool wide = p.fold == 2 || p.N > 64;
  const int bn2 = wide ? 128 : 64;
  p.tiles_m = (p.M + BM2 - 1) / BM2;
  p.tiles_n = p.fold == 2 ? 1 : (p.N + bn2 - 1) / bn2;
  const int64_t zcount = p.split > 1 ? p.split : (p.batch + p.fold - 1) / p.fold;
  const int64_t blocks = p.tiles_m * p.tiles_n * zcount;
  if (blocks >= (1ll << 31)) { set_error("rgnn_gemm: grid too large"); return RGNN_E_RANGE; }
  const size_t smem = 2 * BK2 * AS_LD * sizeof(float2) + 2 * BK2 * (bn2 + 4) * sizeof(float);
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm_simt2<TA, TB, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_gemm_simt2<TA, TB, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_gemm_simt2)");
    done = true;
  }
  if (wide) k_gemm_simt2<TA, TB, 128><<<(unsigned)blocks, 256, smem, st>>>(p);
  else k_gemm_simt2<TA, TB, 64><<<(unsigned)blocks, 256, smem, st>>>(p);
  RGNN_LAUNCH_CHECK("k_gemm_simt2");
  return 0;
}

__global__ void __launch_bounds__(256) k_splitk_reduce(GemmParams p) {
  const int64_t total = p.M * p.N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int64_t s = 0; s < p.split; ++s) acc += p.partial[s * total + i];   // fixed order -> deterministic
    const int64_t m = i / p.N, n = i % p.N;
    const int64_t idx = m * p.sCm + n * p.sCn;
    store_c(p.c_dtype, p.C, idx, epilogue(p.epi, acc, m, n, p.bias, load_aux(p.c_dtype, p.aux, idx), p.alpha, p.beta));
  }
}

// out[j] += sum over (i in chunk, k) x[i*s_outer + j*s_j + k*s_inner]
template <typename T>
__global__ void __launch_bounds__(256) k_reduce_sum(const T* __restrict__ x, int64_t outer, int64_t s_outer, int64_t len, int64_t s_j,
                                                    int64_t inner, int64_t s_inner, float* __restrict__ out, int64_t chunk) {
  const int64_t i0 = (int64_t)blockIdx.y * chunk, i1 = min(outer, i0 + chunk);
  if (s_j == 1 || inner == 1) {
    // j-contiguous: thread = j
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= len) return;
    float acc = 0.f;
    for (int64_t i = i0; i < i1; ++i)
      for (int64_t k = 0; k < inner; ++k) acc += to_f32<T>(x[i * s_outer + j * s_j + k * s_inner]);
    atomicAdd(out + j, acc);
  } else {
    // inner-contiguous (NCHW bias gradient): warp = j, lanes stride the inner run
    const int lane = threadIdx.x & 31;
    const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= len) return;
    float acc = 0.f;
    for (int64_t i = i0; i < i1; ++i)
      for (int64_t k = lane; k < inner; k += 32) acc += to_f32<T>(x[i * s_outer + j * s_j + k * s_inner]);
    acc = warp_sum(acc);
    if (lane == 0) atomicAdd(out + j, acc);
  }
}

// NCHW bias gradient at full bandwidth: x [outer, len, inner] contiguous, inner % VEC == 0.  Thread = one 16-byte vector of
// the (len*inner) row, streams its chunk of `outer`; the inner/VEC lanes of a channel combine by shuffle; one atomic per channel.
template <typename T>
__global__ void __launch_bounds__(256) k_reduce_rows_vec(const T* __restrict__ x, int64_t outer, int64_t len, int inner,
                                                         float* __restrict__ out, int64_t chunk) {
  constexpr int VEC = Vec16<T>::N;
  const int64_t row_vecs = len * inner / VEC;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lpc = inner / VEC;
  const int64_t i0 = (int64_t)blockIdx.y * chunk, i1 = min(outer, i0 + chunk);
  float acc = 0.f;
  if (t < row_vecs) {
    const T* px = x + t * VEC;
    const int64_t stride = len * inner;
    int64_t i = i0;
    for (; i + 4 <= i1; i += 4) {       // 4 independent 16-byte loads in flight
      float v0[VEC], v1[VEC], v2[VEC], v3[VEC];
      Vec16<T>::load(px + i * stride, v0); Vec16<T>::load(px + (i + 1) * stride, v1);
      Vec16<T>::load(px + (i + 2) * stride, v2); Vec16<T>::load(px + (i + 3) * stride, v3);
#pragma unroll
      for (int q = 0; q < VEC; ++q) acc += (v0[q] + v1[q]) + (v2[q] + v3[q]);
    }
    for (; i < i1; ++i) {
      float v0[VEC];
      Vec16<T>::load(px + i * stride, v0);
#pragma unroll
      for (int q = 0; q < VEC; ++q) acc += v0[q];
    }
  }
  for (int o = min(lpc, 32) >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (t < row_vecs && (lpc >= 32 ? (threadIdx.x & 31) == 0 : (t % lpc) == 0)) atomicAdd(out + (t * VEC) / inner, acc);
}

// column sums of a row-major [rows, cols] matrix (bias gradient of nn.Linear): thread = 16-byte vector of columns,
// streams its chunk of rows with 4 loads in flight, VEC atomics at the end.
template <typename T>
__global__ void __launch_bounds__(256) k_reduce_cols_vec(const T* __restrict__ x, int64_t rows, int64_t cols, float* __restrict__ out, int64_t chunk) {
  constexpr int VEC = Vec16<T>::N;
  const int64_t cv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cv * VEC >= cols) return;
  const int64_t i0 = (int64_t)blockIdx.y * chunk, i1 = min(rows, i0 + chunk);
  float acc[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) acc[q] = 0.f;
  const T* px = x + cv * VEC;
  int64_t i = i0;
  for (; i + 4 <= i1; i += 4) {
    float v0[VEC], v1[VEC], v2[VEC], v3[VEC];
    Vec16<T>::load(px + i * cols, v0); Vec16<T>::load(px + (i + 1) * cols, v1);
    Vec16<T>::load(px + (i + 2) * cols, v2); Vec16<T>::load(px + (i + 3) * cols, v3);
#pragma unroll
    for (int q = 0; q < VEC; ++q) acc[q] += (v0[q] + v1[q]) + (v2[q] + v3[q]);
  }
  for (; i < i1; ++i) {
    float v0[VEC];
    Vec16<T>::load(px + i * cols, v0);
#pragma unroll
    for (int q = 0; q < VEC; ++q) acc[q] += v0[q];
  }
#pragma unroll
  for (int q = 0; q < VEC; ++q) atomicAdd(out + cv * VEC + q, acc[q]);
}

}  // namespace rgnn

using namespace rgnn;

extern "C" int rgnn_gemm(int a_dtype, const void* A, int64_t sAz, int64_t sAr, int64_t sAm, int64_t sAk,
                         int b_dtype, const void* B, int64_t sBz, int64_t sBr, int64_t sBk, int64_t sBn,
                         int c_dtype, void* C, int64_t sCz, int64_t sCm, int64_t sCn,
                         int64_t M, int64_t N, int64_t K, int64_t batch, int64_t red, int64_t split, float* partial,
                         int epi, const float* bias, const void* aux, float alpha, float beta, void* stream) {
  RGNN_REQUIRE((a_dtype | 1) == 1 && (b_dtype | 1) == 1 && (c_dtype | 1) == 1, RGNN_E_ARG, "rgnn_gemm: bad dtype");
  RGNN_REQUIRE(M >= 0 && N >= 0 && K >= 0 && batch >= 0 && red >= 0 && split >= 1, RGNN_E_ARG, "rgnn_gemm: bad sizes");
  RGNN_REQUIRE(epi >= 0 && epi < 32, RGNN_E_ARG, "rgnn_gemm: bad epilogue %d", epi);
  RGNN_REQUIRE(!((epi & RGNN_EPI_BIAS_M) && (epi & RGNN_EPI_BIAS_N)), RGNN_E_ARG, "rgnn_gemm: one bias vector only");
  if (M == 0 || N == 0 || batch == 0) return 0;
  RGNN_REQUIRE(A && B && C, RGNN_E_ARG, "rgnn_gemm: null operand");
  RGNN_REQUIRE(!(epi & (RGNN_EPI_BIAS_M | RGNN_EPI_BIAS_N)) || bias, RGNN_E_ARG, "rgnn_gemm: epilogue needs bias");
  RGNN_REQUIRE(!(epi & (RGNN_EPI_RELU_MASK | RGNN_EPI_ADD_AUX)) || aux, RGNN_E_ARG, "rgnn_gemm: epilogue needs aux");
  RGNN_REQUIRE(split == 1 || (batch == 1 && partial), RGNN_E_ARG, "rgnn_gemm: split needs batch == 1 and a partial buffer");
  GemmParams p;
  p.A = A; p.sAz = sAz; p.sAr = sAr; p.sAm = sAm; p.sAk = sAk;
  p.B = B; p.sBz = sBz; p.sBr = sBr; p.sBk = sBk; p.sBn = sBn;
  p.C = C; p.sCz = sCz; p.sCm = sCm; p.sCn = sCn;
  p.M = M; p.N = N; p.K = K; p.batch = batch; p.red = red; p.split = split;
  if (red > 1 && p.split > red) p.split = red;
  if (red <= 1 && p.split > (K + BK - 1) / BK) p.split = (K + BK - 1) / BK;
  if (p.split < 1) p.split = 1;
  p.partial = partial; p.epi = epi; p.bias = bias; p.aux = aux; p.alpha = alpha; p.beta = beta; p.c_dtype = c_dtype;
  p.fold = 1;
  cudaStream_t st = as_stream(stream);
  // large all-fp32 products (the parity path's GEMMs): the FFMA2 kernel, unless the environment says RGNN_SIMT2=0
  static int simt2 = -1;
  if (simt2 < 0) { const char* e = getenv("RGNN_SIMT2"); simt2 = (e && e[0] == '0') ? 0 : 1; }
  if (simt2 && a_dtype == RGNN_F32 && b_dtype == RGNN_F32 && M >= 64 && (N >= 64) && (double)M * N * K * (red > 1 ? red : 1) * batch >= 1e7) {
    if (p.split == 1 && batch > 1 && N == 64 && sAz == 0 && red <= 1) p.fold = 2;
    const int rc = launch_simt2<float, float>(p, st);
    if (rc) return rc;
  } else {
  p.tiles_m = (M + BM - 1) / BM; p.tiles_n = (N + BN - 1) / BN;
  const int64_t blocks = p.tiles_m * p.tiles_n * (p.split > 1 ? p.split : batch);
  RGNN_REQUIRE(blocks < (1ll << 31), RGNN_E_RANGE, "rgnn_gemm: grid too large");
  if (a_dtype == RGNN_F32 && b_dtype == RGNN_F32) k_gemm_simt<float, float><<<(unsigned)blocks, 256, 0, st>>>(p);
  else if (a_dtype == RGNN_F32) k_gemm_simt<float, __nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(p);
  else if (b_dtype == RGNN_F32) k_gemm_simt<__nv_bfloat16, float><<<(unsigned)blocks, 256, 0, st>>>(p);
  else k_gemm_simt<__nv_bfloat16, __nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(p);
  RGNN_LAUNCH_CHECK("k_gemm_simt");
  }
  if (p.split > 1) {
    int64_t rb = (M * N + 255) / 256;
    if (rb > 16 * kNumSMs) rb = 16 * kNumSMs;
    k_splitk_reduce<<<(unsigned)rb, 256, 0, st>>>(p);
    RGNN_LAUNCH_CHECK("k_splitk_reduce");
  }
  return 0;
}

template <typename T>
__global__ void __launch_bounds__(256) k_sigmoid_grad(const T* __restrict__ z, const T* __restrict__ dz, T* __restrict__ o, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float zz = rgnn::to_f32<T>(z[i]);
    o[i] = rgnn::from_f32<T>(rgnn::to_f32<T>(dz[i]) * zz * (1.f - zz));
  }
}

extern "C" int rgnn_sigmoid_grad(int dtype, const void* z, const void* dz, void* dzpre, int64_t n, void* stream) {
  RGNN_REQUIRE(dtype == RGNN_F32 || dtype == RGNN_BF16, RGNN_E_ARG, "rgnn_sigmoid_grad: bad dtype %d", dtype);
  RGNN_REQUIRE(n >= 0, RGNN_E_ARG, "rgnn_sigmoid_grad: negative size");
  if (n == 0) return 0;
  RGNN_REQUIRE(z && dz && dzpre, RGNN_E_ARG, "rgnn_sigmoid_grad: null pointer");
  int64_t blocks = (n + 255) / 256;
  if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
  cudaStream_t st = as_stream(stream);
  if (dtype == RGNN_F32) k_sigmoid_grad<float><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const float*>(z), static_cast<const float*>(dz), static_cast<float*>(dzpre), n);
  else k_sigmoid_grad<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(z), static_cast<const __nv_bfloat16*>(dz), static_cast<__nv_bfloat16*>(dzpre), n);
  RGNN_LAUNCH_CHECK("k_sigmoid_grad");
  return 0;
}

template <typename T>
__global__ void __launch_bounds__(256) k_axpby(float a, const T* __restrict__ x, float b, const T* __restrict__ y, T* __restrict__ o, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = a * rgnn::to_f32<T>(x[i]);
    if (b != 0.f) v = fmaf(b, rgnn::to_f32<T>(y[i]), v);
    o[i] = rgnn::from_f32<T>(v);
  }
}

extern "C" int rgnn_axpby(int dtype, float a, const void* x, float b, const void* y, void* out, int64_t n, void* stream) {
  RGNN_REQUIRE(dtype == RGNN_F32 || dtype == RGNN_BF16, RGNN_E_ARG, "rgnn_axpby: bad dtype %d", dtype);
  RGNN_REQUIRE(n >= 0, RGNN_E_ARG, "rgnn_axpby: negative size");
  if (n == 0) return 0;
  RGNN_REQUIRE(x && y && out, RGNN_E_ARG, "rgnn_axpby: null pointer");
  int64_t blocks = (n + 255) / 256;
  if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
  cudaStream_t st = as_stream(stream);
  if (dtype == RGNN_F32) k_axpby<float><<<(unsigned)blocks, 256, 0, st>>>(a, static_cast<const float*>(x), b, static_cast<const float*>(y), static_cast<float*>(out), n);
  else k_axpby<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(a, static_cast<const __nv_bfloat16*>(x), b, static_cast<const __nv_bfloat16*>(y), static_cast<__nv_bfloat16*>(out), n);
  RGNN_LAUNCH_CHECK("k_axpby");
  return 0;
}

template <typename T>
__global__ void __launch_bounds__(256) k_bias_add(const T* __restrict__ x, const float* __restrict__ b, T* __restrict__ o, int64_t rows, int64_t cols) {
  const int64_t n = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    o[i] = rgnn::from_f32<T>(rgnn::to_f32<T>(x[i]) + b[i % cols]);
}

extern "C" int rgnn_bias_add(int dtype, const void* x, const float* bias, void* out, int64_t rows, int64_t cols, void* stream) {
  RGNN_REQUIRE(dtype == RGNN_F32 || dtype == RGNN_BF16, RGNN_E_ARG, "rgnn_bias_add: bad dtype %d", dtype);
  RGNN_REQUIRE(rows >= 0 && cols >= 0, RGNN_E_ARG, "rgnn_bias_add: negative size");
  if (rows == 0 || cols == 0) return 0;
  RGNN_REQUIRE(x && bias && out, RGNN_E_ARG, "rgnn_bias_add: null pointer");
  int64_t blocks = (rows * cols + 255) / 256;
  if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
  cudaStream_t st = as_stream(stream);
  if (dtype == RGNN_F32) k_bias_add<float><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const float*>(x), bias, static_cast<float*>(out), rows, cols);
  else k_bias_add<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), bias, static_cast<__nv_bfloat16*>(out), rows, cols);
  RGNN_LAUNCH_CHECK("k_bias_add");
  return 0;
}

extern "C" int rgnn_reduce_sum(int dtype, const void* x, int64_t outer, int64_t s_outer, int64_t len, int64_t s_j,
                               int64_t inner, int64_t s_inner, float* out, void* stream) {
  RGNN_REQUIRE(dtype == RGNN_F32 || dtype == RGNN_BF16, RGNN_E_ARG, "rgnn_reduce_sum: bad dtype %d", dtype);
  RGNN_REQUIRE(outer >= 0 && len >= 0 && inner >= 0, RGNN_E_ARG, "rgnn_reduce_sum: bad sizes");
  if (len == 0) return 0;
  RGNN_REQUIRE(out, RGNN_E_ARG, "rgnn_reduce_sum: null out");
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * len, st);
  if (e != cudaSuccess) return cuda_fail(e, "memset reduce out");
  if (outer == 0 || inner == 0) return 0;
  RGNN_REQUIRE(x, RGNN_E_ARG, "rgnn_reduce_sum: null x");
  {  // contiguous [outer, len, inner] with 16-byte vectors inside a channel: the streaming kernel
    const int es = dtype == RGNN_F32 ? 4 : 2, vec = 16 / es;
    const int64_t lpc = inner / vec;
    if (s_inner == 1 && s_j == inner && s_outer == len * inner && inner % vec == 0 && (lpc & (lpc - 1)) == 0 &&
        (reinterpret_cast<uintptr_t>(x) & 15u) == 0 && inner < (1 << 20)) {
      const int64_t row_vecs = len * inner / vec;
      const int64_t bx = (row_vecs + 255) / 256;
      int64_t chunks = (8 * kNumSMs + bx - 1) / bx;
      if (chunks > outer) chunks = outer;
      if (chunks > 65535) chunks = 65535;
      const int64_t chunk = (outer + chunks - 1) / chunks;
      chunks = (outer + chunk - 1) / chunk;
      RGNN_REQUIRE(bx < (1ll << 31), RGNN_E_RANGE, "rgnn_reduce_sum: too long");
      dim3 grid((unsigned)bx, (unsigned)chunks);
      if (dtype == RGNN_F32) k_reduce_rows_vec<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), outer, len, (int)inner, out, chunk);
      else k_reduce_rows_vec<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), outer, len, (int)inner, out, chunk);
      RGNN_LAUNCH_CHECK("k_reduce_rows_vec");
      return 0;
    }
  }
  {  // contiguous row-major [outer, len] column sums with 16-byte vectors
    const int es = dtype == RGNN_F32 ? 4 : 2, vec = 16 / es;
    if (inner == 1 && s_j == 1 && s_outer == len && len % vec == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0) {
      const int64_t cvs = len / vec;
      const int threads = 64;                      // few columns, many rows: small blocks, many row chunks
      const int64_t bx = (cvs + threads - 1) / threads;
      int64_t chunks = (16 * kNumSMs + bx - 1) / bx;
      if (chunks > (outer + 15) / 16) chunks = (outer + 15) / 16;
      if (chunks > 65535) chunks = 65535;
      if (chunks < 1) chunks = 1;
      const int64_t chunk = (outer + chunks - 1) / chunks;
      chunks = (outer + chunk - 1) / chunk;
      dim3 grid((unsigned)bx, (unsigned)chunks);
      if (dtype == RGNN_F32) k_reduce_cols_vec<float><<<grid, threads, 0, st>>>(static_cast<const float*>(x), outer, len, out, chunk);
      else k_reduce_cols_vec<__nv_bfloat16><<<grid, threads, 0, st>>>(static_cast<const __nv_bfloat16*>(x), outer, len, out, chunk);
      RGNN_LAUNCH_CHECK("k_reduce_cols_vec");
      return 0;
    }
  }
  const bool jfast = (s_j == 1 || inner == 1);
  const int64_t bx = jfast ? (len + 255) / 256 : (len + 7) / 8;
  int64_t chunks = (8 * kNumSMs + bx - 1) / bx;
  if (chunks > outer) chunks = outer;
  if (chunks > 65535) chunks = 65535;
  if (chunks < 1) chunks = 1;
  const int64_t chunk = (outer + chunks - 1) / chunks;
  chunks = (outer + chunk - 1) / chunk;
  RGNN_REQUIRE(bx < (1ll << 31), RGNN_E_RANGE, "rgnn_reduce_sum: too long");
  dim3 grid((unsigned)bx, (unsigned)chunks);
  if (dtype == RGNN_F32) k_reduce_sum<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), outer, s_outer, len, s_j, inner, s_inner, out, chunk);
  else k_reduce_sum<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), outer, s_outer, len, s_j, inner, s_inner, out, chunk);
  RGNN_LAUNCH_CHECK("k_reduce_sum");
  return 0;
}
