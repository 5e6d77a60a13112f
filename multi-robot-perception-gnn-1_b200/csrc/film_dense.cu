// A5/A6/A10 for complete K-robot graphs of ANY K (BASELINE configs[4] sweeps K up to 64): where K > 8 no longer fits the
// register-tiled kernels of film_complete.cu, one CTA owns (graph, 64-vector channel tile), stages the K node slices in
// shared memory with the bulk-copy engine (one contiguous cp.async.bulk per robot) and register-tiles over R destinations:
// every staged 16-byte vector is read from shared memory once per R outputs, gamma/beta come straight from global
// (each pair is needed by exactly one lane group, so staging them would only add a copy), and the edge id is computed from
// (u, a) - no CSR walk.  fp32 keeps the reference's rounding order (ascending source, mul then add, true division).
//
// Thread map (256 threads): lane group = 64 vectors of the tile (2 warps), 4 destination groups of R robots each per pass.
#include "common.cuh"

namespace rgnn {

namespace dense {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
template <typename T> __device__ __forceinline__ float2 ldg_pair(const T* p);
template <> __device__ __forceinline__ float2 ldg_pair<float>(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
template <> __device__ __forceinline__ float2 ldg_pair<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p));
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <typename T> __device__ __forceinline__ void st_pair(T* p, float a, float b);
template <> __device__ __forceinline__ void st_pair<float>(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
template <> __device__ __forceinline__ void st_pair<__nv_bfloat16>(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}
__device__ __forceinline__ int edge_of(int K, int u, int a) { return u * (K - 1) + (a < u ? a : a - 1); }

constexpr int NVEC = 64;    // 16-byte vectors per node in a tile (CT = NVEC / LPC channels)
constexpr int NGRP = 4;     // destination groups per CTA

// Stage `nt` node slices (tile of `ct` channels at channel c0) of tensors a (and b, if non-null) into shared memory.
template <typename T>
__device__ __forceinline__ void stage_tiles(uint64_t* bar, T* sa, const T* a, int64_t lda, T* sb, const T* b, int64_t ldb,
                                            int64_t n0, int K, int c0, int ct, int P) {
  const uint32_t bytes = (uint32_t)(ct * P) * sizeof(T);
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) mbar_expect_tx(bar, bytes * (uint32_t)K * (b ? 2u : 1u));
    __syncwarp();
    for (int u = threadIdx.x; u < K; u += 32) {
      bulk_g2s(sa + (size_t)u * ct * P, a + (n0 + u) * lda + (int64_t)c0 * P, bytes, bar);
      if (b) bulk_g2s(sb + (size_t)u * ct * P, b + (n0 + u) * ldb + (int64_t)c0 * P, bytes, bar);
    }
  }
}

// Stage the gamma/beta pairs of the edges {(u, a): u in [0,K), a in [a_base, a_base + AP)} (transposed = false) or
// {(u, a): u in [u_base, u_base + AP), a in [0,K)} (transposed = true) for the tile's CT channels into shared memory:
// zs[outer][inner_local][2*CT].  Each (edge, tile) row is 2*CT*sizeof(T) = 32 bytes = two independent 16-byte loads, all
// issued up front by the whole CTA (the inner loops then touch shared memory only - no exposed global latency).
template <typename T, int CT>
__device__ __forceinline__ void stage_z(T* zs, const T* __restrict__ zc0, int64_t zrow, int K, int base, int AP, bool transposed) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int HALVES = 2 * CT / VEC;            // 16-byte pieces per row (2 for both dtypes)
  const int total = K * AP * HALVES;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int hf = i % HALVES, al = (i / HALVES) % AP, o = i / (HALVES * AP);
    const int u = transposed ? base + al : o, a = transposed ? o : base + al;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (u < K && a < K && u != a) v = __ldg(reinterpret_cast<const uint4*>(zc0 + (int64_t)edge_of(K, u, a) * zrow + hf * VEC));
    *reinterpret_cast<uint4*>(zs + ((size_t)o * AP + al) * 2 * CT + hf * VEC) = v;
  }
}
template <typename T> __device__ __forceinline__ float2 lds_pair(const T* p);
template <> __device__ __forceinline__ float2 lds_pair<float>(const float* p) { return *reinterpret_cast<const float2*>(p); }
template <> __device__ __forceinline__ float2 lds_pair<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint32_t w = *reinterpret_cast<const uint32_t*>(p);
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

// ---------------------------------------------------------------------------------------------------------------------
// forward: out[a] = (1/(K-1)) sum_{u != a} (gamma[e(u,a)] * x[u] + beta[e(u,a)])     (HAS_Z = false: plain mean)
template <typename T, int LPC, int R, bool HAS_Z>
__global__ void __launch_bounds__(256) k_film_fwd_dense(const T* __restrict__ x, int64_t ldx, const T* __restrict__ z, T* __restrict__ out,
                                                        int64_t ldout, T* __restrict__ copy_to, int K, int F, int P) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int CT = NVEC / LPC;
  constexpr int AP = NGRP * R;                 // destinations per pass
  constexpr bool EXACT = sizeof(T) == 4;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  T* xs = reinterpret_cast<T*>(smem + 128);
  T* zs = xs + (size_t)K * CT * P;
  const int tiles_c = F / CT;                  // host guarantees F % CT == 0
  const int64_t g = blockIdx.x / tiles_c;
  const int c0 = (int)(blockIdx.x % tiles_c) * CT;
  const int64_t n0 = g * K, e0 = g * (int64_t)K * (K - 1);
  const int slice = CT * P;
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  stage_tiles<T>(bar, xs, x, ldx, nullptr, nullptr, 0, n0, K, c0, CT, P);
  const int j = threadIdx.x & (NVEC - 1), grp = threadIdx.x / NVEC;
  const int off = j * VEC, cl = off / P;
  const int64_t zrow = 2 * (int64_t)F;
  const T* zc0 = z + e0 * zrow + 2 * c0;
  if (HAS_Z) stage_z<T, CT>(zs, zc0, zrow, K, 0, AP, false);
  mbar_wait(bar, 0);
  if (copy_to && grp == 0) {
    for (int u = 0; u < K; ++u)
      *reinterpret_cast<uint4*>(copy_to + (n0 + u) * ldout + (int64_t)c0 * P + off) = *reinterpret_cast<const uint4*>(xs + (size_t)u * slice + off);
  }
  const float inv = 1.0f / (K - 1);
  for (int a_base = 0; a_base < K; a_base += AP) {
    if (HAS_Z && a_base > 0) { __syncthreads(); stage_z<T, CT>(zs, zc0, zrow, K, a_base, AP, false); }
    __syncthreads();
    const int a0 = a_base + grp * R;
    if (a0 >= K) continue;
    float acc[R][VEC], bsum[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      bsum[r] = 0.f;
#pragma unroll
      for (int q = 0; q < VEC; ++q) acc[r][q] = 0.f;
    }
    for (int u = 0; u < K; ++u) {        // ascending source == ascending edge id among the in-edges of every destination
      float xv[VEC];
      Vec16<T>::load(xs + (size_t)u * slice + off, xv);
      const T* zu = zs + ((size_t)u * AP + grp * R) * 2 * CT + 2 * cl;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int a = a0 + r;
        if (a >= K || a == u) continue;
        float2 gb = make_float2(1.f, 0.f);
        if (HAS_Z) gb = lds_pair<T>(zu + r * 2 * CT);
        if (HAS_Z && !EXACT) bsum[r] += gb.y;
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
          if (!HAS_Z) acc[r][q] = __fadd_rn(acc[r][q], xv[q]);
          else if (EXACT) acc[r][q] = __fadd_rn(acc[r][q], __fadd_rn(__fmul_rn(gb.x, xv[q]), gb.y));
          else acc[r][q] = fmaf(gb.x, xv[q], acc[r][q]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int a = a0 + r;
      if (a >= K) continue;
#pragma unroll
      for (int q = 0; q < VEC; ++q) acc[r][q] = EXACT ? __fdiv_rn(acc[r][q], (float)(K - 1)) : (acc[r][q] + bsum[r]) * inv;
      Vec16<T>::store(out + (n0 + a) * ldout + (int64_t)c0 * P + off, acc[r]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward: dh[u] = addend[u] + w sum_{a != u} gamma[e(u,a)] dagg[a];  dz[e(u,a)] = sig'(z) w (sum_p h[u] dagg[a], sum_p dagg[a])
template <typename T, int LPC, int R>
__global__ void __launch_bounds__(256) k_film_bwd_dense(const T* __restrict__ h, int64_t ldh, const T* __restrict__ dagg, int64_t ldd,
                                                        const T* __restrict__ z, const T* addend, int64_t ldadd, T* dh, int64_t lddh,
                                                        T* __restrict__ dz, int K, int F, int P, int fuse_sig) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int CT = NVEC / LPC;
  constexpr int AP = NGRP * R;                 // sources per pass
  static_assert(R == LPC, "the reduce-scatter below hands one finished source per lane of the channel group");
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  const int tiles_c = F / CT;
  const int64_t g = blockIdx.x / tiles_c;
  const int c0 = (int)(blockIdx.x % tiles_c) * CT;
  const int64_t n0 = g * K, e0 = g * (int64_t)K * (K - 1);
  const int slice = CT * P;
  T* hs = reinterpret_cast<T*>(smem + 128);
  T* ds = hs + (size_t)K * slice;
  T* zs = ds + (size_t)K * slice;
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  stage_tiles<T>(bar, hs, h, ldh, ds, dagg, ldd, n0, K, c0, CT, P);
  const int j = threadIdx.x & (NVEC - 1), grp = threadIdx.x / NVEC;
  const int l = j % LPC;
  const int off = j * VEC, cl = off / P;
  const int64_t zrow = 2 * (int64_t)F;
  const T* zc0 = z + e0 * zrow + 2 * c0;
  T* dzc = dz + e0 * zrow + 2 * (c0 + cl);
  const float w = 1.0f / (K - 1);
  const int64_t go = (int64_t)c0 * P + off;
  stage_z<T, CT>(zs, zc0, zrow, K, 0, AP, true);      // zs[a][u_local][2*CT] for the first chunk of sources
  mbar_wait(bar, 0);

  for (int u_base = 0; u_base < K; u_base += AP) {
    if (u_base > 0) { __syncthreads(); stage_z<T, CT>(zs, zc0, zrow, K, u_base, AP, true); }
    __syncthreads();
    const int u0 = u_base + grp * R;
    if (u0 >= K) continue;                      // warp-uniform: the warps of a destination group leave together
    // ---- dh for sources u0 .. u0+R-1: every destination's staged gradient vector is read once per R sources
    {
      float acc[R][VEC];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int q = 0; q < VEC; ++q) acc[r][q] = 0.f;
      for (int a = 0; a < K; ++a) {
        float dv[VEC];
        Vec16<T>::load(ds + (size_t)a * slice + off, dv);
        const T* za = zs + ((size_t)a * AP + grp * R) * 2 * CT + 2 * cl;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float gam = lds_pair<T>(za + r * 2 * CT).x;     // rows of non-edges (u == a, u >= K) were staged as zeros
#pragma unroll
          for (int q = 0; q < VEC; ++q) acc[r][q] = fmaf(gam, dv[q], acc[r][q]);
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int u = u0 + r;
        if (u >= K) continue;
        if (addend) {
          float av[VEC];
          Vec16<T>::load(addend + (n0 + u) * ldadd + go, av);
#pragma unroll
          for (int q = 0; q < VEC; ++q) acc[r][q] = fmaf(acc[r][q], w, av[q]);
        } else {
#pragma unroll
          for (int q = 0; q < VEC; ++q) acc[r][q] *= w;
        }
        Vec16<T>::store(dh + (n0 + u) * lddh + go, acc[r]);
      }
    }
    // ---- dz for the same sources: R = LPC source vectors in registers; per destination the R partial pixel-dots are
    //      reduce-scattered over the LPC lanes of the channel, so lane l finishes (and stores) the edge of source u0 + l.
    {
      float hv[R][VEC];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (u0 + r < K) Vec16<T>::load(hs + (size_t)(u0 + r) * slice + off, hv[r]);
        else {
#pragma unroll
          for (int q = 0; q < VEC; ++q) hv[r][q] = 0.f;
        }
      }
      const int my_u = u0 + l;
      for (int a = 0; a < K; ++a) {
        float dv[VEC];
        Vec16<T>::load(ds + (size_t)a * slice + off, dv);
        float sd = 0.f;
#pragma unroll
        for (int q = 0; q < VEC; ++q) sd += dv[q];
        float v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float s = 0.f;
#pragma unroll
          for (int q = 0; q < VEC; ++q) s = fmaf(hv[r][q], dv[q], s);
          v[r] = s;
        }
#pragma unroll
        for (int o = LPC / 2; o > 0; o >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);
        // reduce-scatter: halves are exchanged high bit first, so the surviving index is exactly the lane's own number
        {
          int n = R;
#pragma unroll
          for (int o = LPC / 2; o > 0; o >>= 1) {
            n >>= 1;
            const bool up = (l & o) != 0;
#pragma unroll
            for (int i = 0; i < R / 2; ++i) {
              if (i < n) {
                const float send = up ? v[i] : v[i + n];
                const float keep = up ? v[i + n] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
              }
            }
          }
        }
        if (my_u < K && my_u != a) {
          float dg = v[0] * w, db = sd * w;
          if (fuse_sig) {
            const float2 gb = lds_pair<T>(zs + ((size_t)a * AP + grp * R + l) * 2 * CT + 2 * cl);
            dg *= gb.x * (1.f - gb.x); db *= gb.y * (1.f - gb.y);
          }
          st_pair<T>(dzc + (int64_t)edge_of(K, my_u, a) * zrow, dg, db);
        }
      }
    }
  }
}

}  // namespace dense

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T, int LPC>
static int fwd_dense_t(int K, const T* x, int64_t ldx, const T* z, T* out, int64_t ldout, T* copy_to, int64_t G, int F, int P, cudaStream_t st) {
  constexpr int CT = dense::NVEC / LPC;
  if (F % CT != 0) return RGNN_E_UNSUPPORTED;
  const int Rf = K <= 16 ? 4 : 8;
  const size_t smem = 128 + (size_t)K * CT * P * sizeof(T) + (z ? (size_t)K * dense::NGRP * Rf * 2 * CT * sizeof(T) : 0);
  if (smem > 200 * 1024) return RGNN_E_UNSUPPORTED;
  const int64_t blocks = G * (F / CT);
  if (blocks >= (1ll << 31)) return RGNN_E_RANGE;
#define RGNN_GO(RR, ZZ)                                                                                                    \
  do {                                                                                                                     \
    static DeviceOnce once;                                                                                              \
    if (once.need()) {                                                                                                           \
      cudaError_t e = cudaFuncSetAttribute(dense::k_film_fwd_dense<T, LPC, RR, ZZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_film_fwd_dense)");                                 \
      once.mark();                                                                                                         \
    }                                                                                                                      \
    dense::k_film_fwd_dense<T, LPC, RR, ZZ><<<(unsigned)blocks, 256, smem, st>>>(x, ldx, z, out, ldout, copy_to, K, F, P);  \
  } while (0)
  if (K <= 16) { if (z) RGNN_GO(4, true); else RGNN_GO(4, false); }
  else { if (z) RGNN_GO(8, true); else RGNN_GO(8, false); }
#undef RGNN_GO
  RGNN_LAUNCH_CHECK("k_film_fwd_dense");
  return 0;
}

template <typename T, int LPC>
static int bwd_dense_t(int K, const T* h, int64_t ldh, const T* dagg, int64_t ldd, const T* z, const T* addend, int64_t ldadd, T* dh,
                       int64_t lddh, T* dz, int64_t G, int F, int P, int fuse_sig, cudaStream_t st) {
  constexpr int CT = dense::NVEC / LPC;
  if (F % CT != 0) return RGNN_E_UNSUPPORTED;
  const size_t smem = 128 + 2 * (size_t)K * CT * P * sizeof(T) + (size_t)K * dense::NGRP * LPC * 2 * CT * sizeof(T);
  if (smem > 200 * 1024) return RGNN_E_UNSUPPORTED;
  const int64_t blocks = G * (F / CT);
  if (blocks >= (1ll << 31)) return RGNN_E_RANGE;
  static DeviceOnce once;
  if (once.need()) {
    cudaError_t e = cudaFuncSetAttribute(dense::k_film_bwd_dense<T, LPC, LPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_film_bwd_dense)");
    once.mark();
  }
  dense::k_film_bwd_dense<T, LPC, LPC><<<(unsigned)blocks, 256, smem, st>>>(h, ldh, dagg, ldd, z, addend, ldadd, dh, lddh, dz, K, F, P, fuse_sig);
  RGNN_LAUNCH_CHECK("k_film_bwd_dense");
  return 0;
}

// tensor-core variants for bf16, K in {16, 32, 48, 64} (film_mma.cu); RGNN_E_UNSUPPORTED otherwise
int film_fwd_mma(int dtype, int K, const void* x, int64_t ldx, const void* z, void* out, int64_t ldout, void* copy_to, int64_t G, int64_t F,
                 int64_t P, cudaStream_t st);
int film_bwd_mma(int dtype, int K, const void* h, int64_t ldh, const void* dagg, int64_t ldd, const void* z, const void* addend, int64_t ldadd,
                 void* dh, int64_t lddh, void* dz, int64_t G, int64_t F, int64_t P, int fuse_sig, cudaStream_t st);

// Entry points (aggregate.cu dispatches here for complete graphs with K > 8; P/VEC must be 8 (bf16, P=64) or 16 (fp32, P=64)).
int film_fwd_dense(int dtype, int K, const void* x, int64_t ldx, const void* z, void* out, int64_t ldout, void* copy_to,
                   int64_t G, int64_t F, int64_t P, cudaStream_t st) {
  {
    const int rc = film_fwd_mma(dtype, K, x, ldx, z, out, ldout, copy_to, G, F, P, st);
    if (rc != RGNN_E_UNSUPPORTED) return rc;
  }
  const int es = dtype == RGNN_F32 ? 4 : 2, vec = 16 / es;
  if (P % vec != 0 || (P * es) % 16 != 0 || !al16(x) || !al16(out) || (copy_to && !al16(copy_to)) || ldx % vec || ldout % vec) return RGNN_E_UNSUPPORTED;
  if (dtype == RGNN_F32 && P / vec == 16)
    return fwd_dense_t<float, 16>(K, static_cast<const float*>(x), ldx, static_cast<const float*>(z), static_cast<float*>(out), ldout,
                                  static_cast<float*>(copy_to), G, (int)F, (int)P, st);
  if (dtype == RGNN_BF16 && P / vec == 8)
    return fwd_dense_t<__nv_bfloat16, 8>(K, static_cast<const __nv_bfloat16*>(x), ldx, static_cast<const __nv_bfloat16*>(z),
                                         static_cast<__nv_bfloat16*>(out), ldout, static_cast<__nv_bfloat16*>(copy_to), G, (int)F, (int)P, st);
  return RGNN_E_UNSUPPORTED;
}

int film_bwd_dense(int dtype, int K, const void* h, int64_t ldh, const void* dagg, int64_t ldd, const void* z, const void* addend,
                   int64_t ldadd, void* dh, int64_t lddh, void* dz, int64_t G, int64_t F, int64_t P, int fuse_sig, cudaStream_t st) {
  {
    const int rc = film_bwd_mma(dtype, K, h, ldh, dagg, ldd, z, addend, ldadd, dh, lddh, dz, G, F, P, fuse_sig, st);
    if (rc != RGNN_E_UNSUPPORTED) return rc;
  }
  const int es = dtype == RGNN_F32 ? 4 : 2, vec = 16 / es;
  if (P % vec != 0) return RGNN_E_UNSUPPORTED;
  if (dtype == RGNN_F32 && P / vec == 16)
    return bwd_dense_t<float, 16>(K, static_cast<const float*>(h), ldh, static_cast<const float*>(dagg), ldd, static_cast<const float*>(z),
                                  static_cast<const float*>(addend), ldadd, static_cast<float*>(dh), lddh, static_cast<float*>(dz), G, (int)F, (int)P, fuse_sig, st);
  if (dtype == RGNN_BF16 && P / vec == 8)
    return bwd_dense_t<__nv_bfloat16, 8>(K, static_cast<const __nv_bfloat16*>(h), ldh, static_cast<const __nv_bfloat16*>(dagg), ldd,
                                         static_cast<const __nv_bfloat16*>(z), static_cast<const __nv_bfloat16*>(addend), ldadd,
                                         static_cast<__nv_bfloat16*>(dh), lddh, static_cast<__nv_bfloat16*>(dz), G, (int)F, (int)P, fuse_sig, st);
  return RGNN_E_UNSUPPORTED;
}

}  // namespace rgnn
