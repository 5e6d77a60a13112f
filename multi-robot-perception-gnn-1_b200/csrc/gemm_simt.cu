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
  int fold;            // k_sgemm_v only: 2 = one 128-column tile spans two batch items of N = 64 columns (per-node conv products)
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
// Vectorised fp32 kernel for the large, aligned all-fp32 products of the parity mode (same arithmetic as k_gemm_simt: one
// fmaf chain per output in ascending k and the same split-K chunking, so results are bit-identical).  128 x BN2 x 16 tiles,
// 8 x (BN2/16) register micro-tile, packed FFMA2; the A operand is kept duplicated in shared memory so that (a, a) pairs load
// directly; 16-byte global loads through per-thread pointers that advance by a constant per tile; the next tile's loads are
// in flight while the current one is multiplied (register staging, double-buffered shared memory, one barrier per tile).
constexpr int VM = 128, VK = 16;
constexpr int VA_LD = VM + 2;                  // float2 per k row (+2: rows stay 16-byte aligned, k-fast stores spread over banks)

template <bool A_KFAST, bool B_NFAST, int BN2>
__global__ void __launch_bounds__(256, 2) k_sgemm_v(GemmParams p) {
  constexpr int TNV = BN2 / 64;                // float4 column chunks per thread (chunk c = columns c*64 + tx*4 .. +3)
  constexpr int VB_LD = BN2 + 4;
  constexpr int NBV = BN2 * VK / 4 / 256;      // float4 of B per thread per tile (2 or 1)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* As = reinterpret_cast<float2*>(smem_raw);                                   // [2][VK][VA_LD]
  float* Bs = reinterpret_cast<float*>(smem_raw + 2 * VK * VA_LD * sizeof(float2));    // [2][VK][VB_LD]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  int64_t bid = blockIdx.x;
  const int64_t tn = bid % p.tiles_n; bid /= p.tiles_n;
  const int64_t tm = bid % p.tiles_m; bid /= p.tiles_m;
  const int64_t zs = bid;
  const bool fold = p.fold == 2;
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
  const int64_t m0 = tm * VM, n0 = tn * BN2;
  const float* __restrict__ A = static_cast<const float*>(p.A) + z * p.sAz;
  const float* __restrict__ B = static_cast<const float*>(p.B) + z * p.sBz;
  const int tpr = (int)((kend - kbeg + VK - 1) / VK);            // k tiles per reduction item
  const int64_t ntiles = (r1 > r0 && kend > kbeg) ? (r1 - r0) * tpr : 0;

  // ---- per-thread load slots: element offsets inside the operand at (r = 0, k = kbeg), validity along the non-k direction
  int a_k[2], a_m[2], b_k[NBV], b_n[NBV];
  int64_t a_off[2], b_off[NBV];
  bool a_ok[2], b_ok[NBV];
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int q = tid + it * 256;
    if (A_KFAST) { a_k[it] = (q & 3) * 4; a_m[it] = q >> 2; } else { a_m[it] = (q & 31) * 4; a_k[it] = q >> 5; }
    a_ok[it] = m0 + a_m[it] < p.M;           // M % 4 == 0 when vectors run along m
    a_off[it] = (m0 + a_m[it]) * p.sAm + (kbeg + a_k[it]) * p.sAk;
  }
#pragma unroll
  for (int it = 0; it < NBV; ++it) {
    const int q = tid + it * 256;
    if (B_NFAST) { b_n[it] = (q % (BN2 / 4)) * 4; b_k[it] = q / (BN2 / 4); } else { b_k[it] = (q & 3) * 4; b_n[it] = q >> 2; }
    int64_t gn = n0 + b_n[it], zo = 0;
    if (fold) { zo = b_n[it] >> 6; gn = b_n[it] & 63; }
    b_ok[it] = gn < p.N && (!fold || z + zo < p.batch);
    b_off[it] = zo * p.sBz + gn * p.sBn + (kbeg + b_k[it]) * p.sBk;
  }

  float2 acc[8][2 * TNV];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 2 * TNV; ++j) acc[i][j] = make_float2(0.f, 0.f);

  float4 ra[2], rb[NBV];
  int64_t r = r0; int kt = 0;                  // the tile the NEXT gload fetches: reduction item r, k tile kt
  auto gload = [&]() {
    const int64_t k0 = kbeg + (int64_t)kt * VK;
    const float* Ar = A + r * p.sAr + (int64_t)kt * VK * p.sAk;
    const float* Br = B + r * p.sBr + (int64_t)kt * VK * p.sBk;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const bool ok = a_ok[it] && k0 + a_k[it] < kend;          // K % 4 == 0 when vectors run along k
      ra[it] = ok ? __ldg(reinterpret_cast<const float4*>(Ar + a_off[it])) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int it = 0; it < NBV; ++it) {
      const bool ok = b_ok[it] && k0 + b_k[it] < kend;
      rb[it] = ok ? __ldg(reinterpret_cast<const float4*>(Br + b_off[it])) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (++kt == tpr) { kt = 0; ++r; }
  };
  auto sstore = [&](int buf) {
    float2* as = As + buf * VK * VA_LD;
    float* bs = Bs + buf * VK * VB_LD;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      if (A_KFAST) {
        float2* d = as + a_k[it] * VA_LD + a_m[it];
        d[0] = make_float2(ra[it].x, ra[it].x); d[VA_LD] = make_float2(ra[it].y, ra[it].y);
        d[2 * VA_LD] = make_float2(ra[it].z, ra[it].z); d[3 * VA_LD] = make_float2(ra[it].w, ra[it].w);
      } else {
        float4* d = reinterpret_cast<float4*>(as + a_k[it] * VA_LD + a_m[it]);
        d[0] = make_float4(ra[it].x, ra[it].x, ra[it].y, ra[it].y);
        d[1] = make_float4(ra[it].z, ra[it].z, ra[it].w, ra[it].w);
      }
    }
#pragma unroll
    for (int it = 0; it < NBV; ++it) {
      if (B_NFAST) {
        *reinterpret_cast<float4*>(bs + b_k[it] * VB_LD + b_n[it]) = rb[it];
      } else {
        float* d = bs + b_k[it] * VB_LD + b_n[it];
        d[0] = rb[it].x; d[VB_LD] = rb[it].y; d[2 * VB_LD] = rb[it].z; d[3 * VB_LD] = rb[it].w;
      }
    }
  };

  if (ntiles > 0) { gload(); sstore(0); }
  __syncthreads();
  for (int64_t t = 0; t < ntiles; ++t) {
    const int buf = (int)(t & 1);
    if (t + 1 < ntiles) gload();
    const float2* as = As + buf * VK * VA_LD + ty * 8;
    const float* bs = Bs + buf * VK * VB_LD + tx * 4;
#pragma unroll
    for (int kk = 0; kk < VK; ++kk) {
      float2 b2[2 * TNV];
#pragma unroll
      for (int c = 0; c < TNV; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(bs + kk * VB_LD + c * 64);
        b2[2 * c] = make_float2(v.x, v.y); b2[2 * c + 1] = make_float2(v.z, v.w);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(as + kk * VA_LD + 2 * i);
        const float2 a0 = make_float2(v.x, v.y), a1 = make_float2(v.z, v.w);
#pragma unroll
        for (int j = 0; j < 2 * TNV; ++j) {
          acc[2 * i][j] = ffma2(a0, b2[j], acc[2 * i][j]);
          acc[2 * i + 1][j] = ffma2(a1, b2[j], acc[2 * i + 1][j]);
        }
      }
    }
    if (t + 1 < ntiles) sstore(buf ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty * 8 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int c = 0; c < TNV; ++c) {
      const int col = c * 64 + tx * 4;
      int64_t n = n0 + col, zz = z;
      if (fold) { zz = z + c; n = tx * 4; }
      if (fold && zz >= p.batch) continue;
      const float v[4] = {acc[i][2 * c].x, acc[i][2 * c].y, acc[i][2 * c + 1].x, acc[i][2 * c + 1].y};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n + j >= p.N) continue;
        if (p.split > 1) {
          p.partial[(zs * p.M + m) * p.N + n + j] = v[j];
        } else {
          const int64_t idx = zz * p.sCz + m * p.sCm + (n + j) * p.sCn;
          store_c(p.c_dtype, p.C, idx, epilogue(p.epi, v[j], m, n + j, p.bias, load_aux(p.c_dtype, p.aux, idx), p.alpha, p.beta));
        }
      }
    }
  }
}

static bool al16p(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

// true when k_sgemm_v can take the product: all-fp32 operands, unit stride along one direction of A and of B, every vector
// 16-byte aligned and entirely inside or outside the operand
static bool sgemm_v_ok(const GemmParams& p, int a_dtype, int b_dtype) {
  if (a_dtype != RGNN_F32 || b_dtype != RGNN_F32) return false;
  if (p.M < 64 || p.N < 64) return false;
  if ((double)p.M * p.N * p.K * (p.red > 1 ? p.red : 1) * p.batch < 2e7) return false;
  const bool a_k = p.sAk == 1, a_m = p.sAm == 1, b_n = p.sBn == 1, b_k = p.sBk == 1;
  if (!(a_k || a_m) || !(b_n || b_k)) return false;
  if (!al16p(p.A) || !al16p(p.B)) return false;
  auto m4 = [](int64_t v) { return (v & 3) == 0; };
  if (!m4(p.sAz) || !m4(p.sAr) || !m4(p.sBz) || !m4(p.sBr)) return false;
  if (a_k) { if (!m4(p.sAm) || !m4(p.K)) return false; } else { if (!m4(p.sAk) || !m4(p.M)) return false; }
  if (b_n) { if (!m4(p.sBk) || !m4(p.N)) return false; } else { if (!m4(p.sBn) || !m4(p.K)) return false; }
  return true;
}

static int launch_sgemm_v(GemmParams& p, cudaStream_t st) {
  const bool a_k = p.sAk == 1, b_n = p.sBn == 1;
  if (p.split == 1 && p.batch > 1 && p.N == 64 && p.sAz == 0 && p.red <= 1 && b_n) p.fold = 2;
  const bool wide = p.fold == 2 || p.N > 64;
  const int bn2 = wide ? 128 : 64;
  p.tiles_m = (p.M + VM - 1) / VM;
  p.tiles_n = p.fold == 2 ? 1 : (p.N + bn2 - 1) / bn2;
  const int64_t zcount = p.split > 1 ? p.split : (p.batch + p.fold - 1) / p.fold;
  const int64_t blocks = p.tiles_m * p.tiles_n * zcount;
  if (blocks >= (1ll << 31)) { set_error("rgnn_gemm: grid too large"); return RGNN_E_RANGE; }
  const size_t smem = 2 * VK * VA_LD * sizeof(float2) + 2 * VK * (bn2 + 4) * sizeof(float);
#define RGNN_SG(AK, BNF, W)                                                                                            \
  {                                                                                                                    \
    static DeviceOnce once;                                                                                          \
    if (once.need()) {                                                                                                       \
      cudaError_t e = cudaFuncSetAttribute(k_sgemm_v<AK, BNF, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); \
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_sgemm_v)");                                    \
      once.mark();                                                                                                     \
    }                                                                                                                  \
    k_sgemm_v<AK, BNF, W><<<(unsigned)blocks, 256, smem, st>>>(p);                                                     \
  }
  if (wide) {
    if (a_k && b_n) RGNN_SG(true, true, 128) else if (a_k) RGNN_SG(true, false, 128) else if (b_n) RGNN_SG(false, true, 128) else RGNN_SG(false, false, 128)
  } else {
    if (a_k && b_n) RGNN_SG(true, true, 64) else if (a_k) RGNN_SG(true, false, 64) else if (b_n) RGNN_SG(false, true, 64) else RGNN_SG(false, false, 64)
  }
#undef RGNN_SG
  RGNN_LAUNCH_CHECK("k_sgemm_v");
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

// Same sums with 256-thread blocks = cv column vectors x rg row groups: every thread walks rows rg, rg + RG, ... of the block's
// chunk with eight 16-byte loads in flight, the row groups meet in shared memory and one thread per column vector publishes -
// an order of magnitude fewer atomics per output address than one 64-thread block per 70-row chunk, and the HBM pipe stays full.
template <typename T>
__global__ void __launch_bounds__(256) k_reduce_cols_vec2(const T* __restrict__ x, int64_t rows, int64_t cols, float* __restrict__ out, int64_t chunk,
                                                          int cv) {
  constexpr int VEC = Vec16<T>::N;
  __shared__ float red[256 * VEC];
  const int rgn = blockDim.x / cv;
  const int c = threadIdx.x % cv, rg = threadIdx.x / cv;
  const int64_t col = ((int64_t)blockIdx.x * cv + c) * VEC;
  const bool live = col < cols;
  const int64_t i0 = (int64_t)blockIdx.y * chunk, i1 = min(rows, i0 + chunk);
  float acc[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) acc[q] = 0.f;
  if (live) {
    const T* px = x + col;
    int64_t i = i0 + rg;
    for (; i + 7 * rgn < i1; i += 8 * rgn) {
      uint4 u[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) u[k] = *reinterpret_cast<const uint4*>(px + (i + (int64_t)k * rgn) * cols);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float v[VEC];
        Vec16<T>::unpack(u[k], v);
#pragma unroll
        for (int q = 0; q < VEC; ++q) acc[q] += v[q];
      }
    }
    for (; i < i1; i += rgn) {
      float v[VEC];
      Vec16<T>::load(px + i * cols, v);
#pragma unroll
      for (int q = 0; q < VEC; ++q) acc[q] += v[q];
    }
  }
#pragma unroll
  for (int q = 0; q < VEC; ++q) red[threadIdx.x * VEC + q] = acc[q];
  __syncthreads();
  if (rg == 0 && live) {
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      float v = 0.f;
      for (int g2 = 0; g2 < rgn; ++g2) v += red[(g2 * cv + c) * VEC + q];
      atomicAdd(out + col + q, v);
    }
  }
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
  static int use_v = -1;
  if (use_v < 0) { const char* e = getenv("RGNN_SGEMM_V"); use_v = (e && e[0] == '0') ? 0 : 1; }   // RGNN_SGEMM_V=0: general kernel only
  if (use_v && sgemm_v_ok(p, a_dtype, b_dtype)) {
    const int rc = launch_sgemm_v(p, st);
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
    if (rb > 16 * num_sms()) rb = 16 * num_sms();
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
  if (blocks > 16 * num_sms()) blocks = 16 * num_sms();
  cudaStream_t st = as_stream(stream);
  if (dtype == RGNN_F32) k_sigmoid_grad<float><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const float*>(z), static_cast<const float*>(dz), static_cast<float*>(dzpre), n);
  else k_sigmoid_grad<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(z), static_cast<const __nv_bfloat16*>(dz), static_cast<__nv_bfloat16*>(dzpre), n);
  RGNN_LAUNCH_CHECK("k_sigmoid_grad");
  return 0;
}

template <typename T>
__global__ void __launch_bounds__(256) k_relu_grad(const T* __restrict__ a, const T* __restrict__ da, T* __restrict__ o, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    o[i] = rgnn::to_f32<T>(a[i]) > 0.f ? da[i] : rgnn::from_f32<T>(0.f);
}

extern "C" int rgnn_relu_grad(int dtype, const void* a, const void* da, void* out, int64_t n, void* stream) {
  RGNN_REQUIRE(dtype == RGNN_F32 || dtype == RGNN_BF16, RGNN_E_ARG, "rgnn_relu_grad: bad dtype %d", dtype);
  RGNN_REQUIRE(n >= 0, RGNN_E_ARG, "rgnn_relu_grad: negative size");
  if (n == 0) return 0;
  RGNN_REQUIRE(a && da && out, RGNN_E_ARG, "rgnn_relu_grad: null pointer");
  int64_t blocks = (n + 255) / 256;
  if (blocks > 16 * num_sms()) blocks = 16 * num_sms();
  cudaStream_t st = as_stream(stream);
  if (dtype == RGNN_F32) k_relu_grad<float><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const float*>(a), static_cast<const float*>(da), static_cast<float*>(out), n);
  else k_relu_grad<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(da), static_cast<__nv_bfloat16*>(out), n);
  RGNN_LAUNCH_CHECK("k_relu_grad");
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
  if (blocks > 16 * num_sms()) blocks = 16 * num_sms();
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
  if (blocks > 16 * num_sms()) blocks = 16 * num_sms();
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
      int64_t chunks = (8 * num_sms() + bx - 1) / bx;
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
      if (outer >= 4096) {                         // tall matrices (bias gradients over the edges / pixels): row-group blocks
        int cv = 32;
        while (cv < cvs && cv < 256) cv <<= 1;
        const int64_t bx2 = (cvs + cv - 1) / cv;
        int64_t chunks2 = (4 * num_sms() + bx2 - 1) / bx2;
        if (chunks2 > 65535) chunks2 = 65535;
        if (chunks2 < 1) chunks2 = 1;
        const int64_t chunk2 = (outer + chunks2 - 1) / chunks2;
        chunks2 = (outer + chunk2 - 1) / chunk2;
        dim3 grid2((unsigned)bx2, (unsigned)chunks2);
        if (dtype == RGNN_F32) k_reduce_cols_vec2<float><<<grid2, 256, 0, st>>>(static_cast<const float*>(x), outer, len, out, chunk2, cv);
        else k_reduce_cols_vec2<__nv_bfloat16><<<grid2, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), outer, len, out, chunk2, cv);
        RGNN_LAUNCH_CHECK("k_reduce_cols_vec2");
        return 0;
      }
      const int threads = 64;                      // few columns, many rows: small blocks, many row chunks
      const int64_t bx = (cvs + threads - 1) / threads;
      int64_t chunks = (16 * num_sms() + bx - 1) / bx;
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
  int64_t chunks = (8 * num_sms() + bx - 1) / bx;
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
