// A5/A6/A10 on complete K-robot graphs (dgl/model/models.py:207-211,223 and their autograd), streaming variant for P = 8x8:
// a producer warp keeps a ring of shared-memory stages full with the bulk-copy engine (cp.async.bulk + mbarrier, no tensor
// map: every piece is a contiguous run), consumer warps read their vectors from shared memory, do the FiLM arithmetic with
// packed fp32 (FFMA2) and store straight to global memory.  Compared with the register-prefetch kernels of film_complete.cu
// the bytes in flight per SM no longer depend on occupancy x registers (NST-1 whole stages are always outstanding), no LSU
// instruction is spent on global loads, and gamma/beta come from the same stage instead of 20 scattered 4-byte loads.
//
// Work item = (graph g, block of CH channels): K pieces of CH*P elements (4 KB each, one per robot; kinds x K for the
// backward) plus K(K-1) pieces of 2*CH interleaved gamma/beta values (128 B each).  CH * P = TPB * 16 B / sizeof(T).
#include <stdlib.h>
#include "common.cuh"
#include "film_common.cuh"

namespace rgnn {

constexpr int RING_P = 64;        // pixels per channel this variant is built for (8x8 maps: image_size 256 / 32, models.py:41)
constexpr int RING_TPB = 256;     // consumer threads = 16-byte vectors per robot per item

template <typename T, int K> struct RingShape {
  static constexpr int VEC = Vec16<T>::N;
  static constexpr int NE = K * (K - 1);
  static constexpr int LPC = RING_P / VEC;                 // lanes per channel
  static constexpr int CH = RING_TPB / LPC;                // channels per item: 32 (bf16) / 16 (fp32)
  static constexpr int PIECE = RING_TPB * 16;              // 4 KB
  static constexpr int ZPIECE = 2 * CH * (int)sizeof(T);   // 128 B
};

__device__ __forceinline__ float2 unpack_pair(uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }

template <typename T> __device__ __forceinline__ float2 lds_pair(const unsigned char* p);
template <> __device__ __forceinline__ float2 lds_pair<float>(const unsigned char* p) { return *reinterpret_cast<const float2*>(p); }
template <> __device__ __forceinline__ float2 lds_pair<__nv_bfloat16>(const unsigned char* p) { return unpack_pair(*reinterpret_cast<const uint32_t*>(p)); }

// ---------------------------------------------------------------------------------------------------------------------
// forward: out[a] = mean_{u != a} (gamma[e(u,a)] * x[u] + beta[e(u,a)])       (HAS_Z = false: plain mean, copy_u)
template <typename T, int K, bool HAS_Z, int NST>
__global__ void __launch_bounds__(RING_TPB + 32) k_film_fwd_ring(const T* __restrict__ x, int64_t ldx, const T* __restrict__ z,
                                                                 T* __restrict__ out, int64_t ldout, T* __restrict__ copy_to,
                                                                 int64_t G, int F, const int2* __restrict__ gb) {
  // gb (optional): per-graph (first node row, first edge row) of the graphs this launch covers - a bucket of equal-K graphs
  // inside a ragged batch; null = graph g owns rows [g K, (g+1) K) and edges [g NE, (g+1) NE)
  using S = RingShape<T, K>;
  constexpr int VEC = S::VEC, NE = S::NE, NP = VEC / 2;
  constexpr bool EXACT = sizeof(T) == 4;       // fp32: the reference's rounding order; bf16: fused multiply-add
  constexpr int STAGE = K * S::PIECE + (HAS_Z ? NE * S::ZPIECE : 0);
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + NST;
  unsigned char* ring = smem + 128;
  const int tid = threadIdx.x, lane = tid & 31;
  const int ipg = F / S::CH;                               // items per graph
  const int64_t items = G * ipg;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) { tma::mbar_init(full + s, 1); tma::mbar_init(empty + s, RING_TPB / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (tid >= RING_TPB) {                                   // ---- producer warp
    int s = 0; uint32_t ph = 0;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
      const int64_t g = it / ipg;
      const int cb = (int)(it - g * ipg);
      const int64_t nb = gb ? (int64_t)gb[g].x : g * K, eb = gb ? (int64_t)gb[g].y : g * NE;
      tma::mbar_wait(empty + s, ph ^ 1);                   // first lap: passes immediately (phase bit starts at 0)
      unsigned char* st = ring + (size_t)s * STAGE;
      if (lane == 0) tma::mbar_expect_tx(full + s, STAGE);
      __syncwarp();
      for (int i = lane; i < K + (HAS_Z ? NE : 0); i += 32) {
        if (i < K) tma::bulk_g2s(st + i * S::PIECE, x + (nb + i) * ldx + (int64_t)cb * S::CH * RING_P, S::PIECE, full + s);
        else tma::bulk_g2s(st + K * S::PIECE + (i - K) * S::ZPIECE, z + (eb + (i - K)) * 2 * (int64_t)F + 2 * cb * S::CH, S::ZPIECE, full + s);
      }
      if (++s == NST) { s = 0; ph ^= 1; }
    }
    return;
  }

  // ---- consumers
  const int cl = tid / S::LPC;                             // channel of this thread inside the item
  const float w = 1.0f / (K - 1);
  int s = 0; uint32_t ph = 0;
  for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
    const int64_t g = it / ipg;
    const int cb = (int)(it - g * ipg);
    const int64_t off = (int64_t)cb * S::CH * RING_P + tid * VEC;
    const int64_t n0 = gb ? (int64_t)gb[g].x : g * K;
    const unsigned char* st = ring + (size_t)s * STAGE;
    tma::mbar_wait(full + s, ph);
    float2 xv[K][NP];
#pragma unroll
    for (int u = 0; u < K; ++u) {
      const uint4 cx = *reinterpret_cast<const uint4*>(st + u * S::PIECE + tid * 16);
      if (copy_to) *reinterpret_cast<uint4*>(copy_to + (n0 + u) * ldout + off) = cx;
      if constexpr (EXACT) {
        xv[u][0] = make_float2(__uint_as_float(cx.x), __uint_as_float(cx.y));
        xv[u][1] = make_float2(__uint_as_float(cx.z), __uint_as_float(cx.w));
      } else {
        xv[u][0] = unpack_pair(cx.x); xv[u][1] = unpack_pair(cx.y); xv[u][2] = unpack_pair(cx.z); xv[u][3] = unpack_pair(cx.w);
      }
    }
    const unsigned char* zs = st + K * S::PIECE + cl * 2 * (int)sizeof(T);
#pragma unroll
    for (int a = 0; a < K; ++a) {
      float2 acc[NP];
#pragma unroll
      for (int q = 0; q < NP; ++q) acc[q] = make_float2(0.f, 0.f);
      float bsum = 0.f;
#pragma unroll
      for (int u = 0; u < K; ++u) {      // ascending u == ascending edge id among the in-edges of a
        if (u == a) continue;
        if constexpr (!HAS_Z) {
#pragma unroll
          for (int q = 0; q < NP; ++q) acc[q] = fadd2(acc[q], xv[u][q]);
        } else {
          const float2 gb = lds_pair<T>(zs + edge_id(K, u, a) * S::ZPIECE);
          const float2 g2 = make_float2(gb.x, gb.x);
          if constexpr (EXACT) {      // scalar on purpose: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (one rounding)
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              acc[q].x = __fadd_rn(acc[q].x, __fadd_rn(__fmul_rn(gb.x, xv[u][q].x), gb.y));
              acc[q].y = __fadd_rn(acc[q].y, __fadd_rn(__fmul_rn(gb.x, xv[u][q].y), gb.y));
            }
          } else {
            bsum += gb.y;
#pragma unroll
            for (int q = 0; q < NP; ++q) acc[q] = ffma2(g2, xv[u][q], acc[q]);
          }
        }
      }
      float o[VEC];
      if constexpr (EXACT) {
#pragma unroll
        for (int q = 0; q < NP; ++q) { o[2 * q] = __fdiv_rn(acc[q].x, (float)(K - 1)); o[2 * q + 1] = __fdiv_rn(acc[q].y, (float)(K - 1)); }
      } else {
        const float2 b2 = make_float2(bsum, bsum), w2 = make_float2(w, w);
#pragma unroll
        for (int q = 0; q < NP; ++q) { const float2 r = fmul2(fadd2(acc[q], b2), w2); o[2 * q] = r.x; o[2 * q + 1] = r.y; }
      }
      Vec16<T>::store(out + (n0 + a) * ldout + off, o);
    }
    __syncwarp();
    if (lane == 0) tma::mbar_arrive(empty + s);
    if (++s == NST) { s = 0; ph ^= 1; }
  }
}

template <int K> constexpr int ring_fwd_stages() { return K <= 3 ? 8 : (K <= 5 ? 4 : 3); }

template <typename T, int K, bool HAS_Z>
static int launch_fwd_ring_k(const T* x, int64_t ldx, const T* z, T* out, int64_t ldout, T* copy_to, int64_t G, int F, cudaStream_t st, const int2* gb) {
  using S = RingShape<T, K>;
  constexpr int NST = ring_fwd_stages<K>();
  constexpr size_t smem = 128 + (size_t)NST * (K * S::PIECE + (HAS_Z ? S::NE * S::ZPIECE : 0));
  auto kern = k_film_fwd_ring<T, K, HAS_Z, NST>;
  static DeviceOnce once;
  static int per_sm = 1;
  if (once.need()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_film_fwd_ring)");
    int occ = 1;          // (same answer on every B200: concurrent per-device threads store the same value)
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, RING_TPB + 32, smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor(k_film_fwd_ring)");
    per_sm = occ < 1 ? 1 : occ;
    once.mark();
  }
  const int64_t items = G * (F / S::CH);
  int64_t blocks = (int64_t)per_sm * num_sms();
  if (blocks > items) blocks = items;
  if (blocks < 1) return 0;
  kern<<<(unsigned)blocks, RING_TPB + 32, smem, st>>>(x, ldx, z, out, ldout, copy_to, G, F, gb);
  RGNN_LAUNCH_CHECK("k_film_fwd_ring");
  return 0;
}

template <typename T, bool HAS_Z>
static int launch_fwd_ring(int K, const T* x, int64_t ldx, const T* z, T* out, int64_t ldout, T* copy_to, int64_t G, int F, cudaStream_t st, const int2* gb) {
#define RGNN_CASE(KK) case KK: return launch_fwd_ring_k<T, KK, HAS_Z>(x, ldx, z, out, ldout, copy_to, G, F, st, gb);
  switch (K) { RGNN_CASE(2) RGNN_CASE(3) RGNN_CASE(4) RGNN_CASE(5) RGNN_CASE(6) RGNN_CASE(7) RGNN_CASE(8) default: return RGNN_E_UNSUPPORTED; }
#undef RGNN_CASE
}

template <typename T> __device__ __forceinline__ void unpack2(const uint4& v, float2 (&o)[Vec16<T>::N / 2]);
template <> __device__ __forceinline__ void unpack2<float>(const uint4& v, float2 (&o)[2]) {
  o[0] = make_float2(__uint_as_float(v.x), __uint_as_float(v.y)); o[1] = make_float2(__uint_as_float(v.z), __uint_as_float(v.w));
}
template <> __device__ __forceinline__ void unpack2<__nv_bfloat16>(const uint4& v, float2 (&o)[4]) {
  o[0] = unpack_pair(v.x); o[1] = unpack_pair(v.y); o[2] = unpack_pair(v.z); o[3] = unpack_pair(v.w);
}

// ---------------------------------------------------------------------------------------------------------------------
// GCNII initial residual on complete graphs (dgl GCN2Conv): out[a] = sa * mean_{u != a} x[u] + sb * addend[a].
// The symmetric normalisation D^-1/2 A D^-1/2 of a complete K-graph is the mean over the other K-1 robots, and the alpha mix
// with the initial features rides on the same pass (one kernel instead of aggregation + axpby: 4 instead of 7 tensor passes).
template <typename T, int K, bool HAS_ADD, int NST>
__global__ void __launch_bounds__(RING_TPB + 32) k_mean_mix_ring(const T* __restrict__ x, int64_t ldx, const T* __restrict__ addend, int64_t ldadd,
                                                                 T* __restrict__ out, int64_t ldout, int64_t G, int ipg, float sa, float sb) {
  using S = RingShape<T, K>;
  constexpr int VEC = S::VEC, NP = VEC / 2;
  constexpr int STAGE = (HAS_ADD ? 2 : 1) * K * S::PIECE;
  constexpr int ROWV = RING_TPB * VEC;                     // elements of one node row per item
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + NST;
  unsigned char* ring = smem + 128;
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t items = G * ipg;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) { tma::mbar_init(full + s, 1); tma::mbar_init(empty + s, RING_TPB / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid >= RING_TPB) {
    int s = 0; uint32_t ph = 0;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
      const int64_t g = it / ipg;
      const int64_t coff = (it - g * ipg) * ROWV;
      tma::mbar_wait(empty + s, ph ^ 1);
      unsigned char* st = ring + (size_t)s * STAGE;
      if (lane == 0) tma::mbar_expect_tx(full + s, STAGE);
      __syncwarp();
      if (lane < K) tma::bulk_g2s(st + lane * S::PIECE, x + (g * K + lane) * ldx + coff, S::PIECE, full + s);
      else if (HAS_ADD && lane < 2 * K) tma::bulk_g2s(st + lane * S::PIECE, addend + (g * K + lane - K) * ldadd + coff, S::PIECE, full + s);
      if (++s == NST) { s = 0; ph ^= 1; }
    }
    return;
  }
  const float wa = sa / (K - 1);
  const float2 wa2 = make_float2(wa, wa), sb2 = make_float2(sb, sb);
  int s = 0; uint32_t ph = 0;
  for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
    const int64_t g = it / ipg;
    const int64_t off = (it - g * ipg) * ROWV + tid * VEC;
    const unsigned char* st = ring + (size_t)s * STAGE;
    tma::mbar_wait(full + s, ph);
    float2 xv[K][NP];
#pragma unroll
    for (int u = 0; u < K; ++u) unpack2<T>(*reinterpret_cast<const uint4*>(st + u * S::PIECE + tid * 16), xv[u]);
#pragma unroll
    for (int a = 0; a < K; ++a) {
      float2 acc[NP];
#pragma unroll
      for (int q = 0; q < NP; ++q) acc[q] = make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < K; ++u) {
        if (u == a) continue;
#pragma unroll
        for (int q = 0; q < NP; ++q) acc[q] = fadd2(acc[q], xv[u][q]);
      }
      float o[VEC];
      if constexpr (HAS_ADD) {
        float2 av[NP];
        unpack2<T>(*reinterpret_cast<const uint4*>(st + (K + a) * S::PIECE + tid * 16), av);
#pragma unroll
        for (int q = 0; q < NP; ++q) { const float2 r = ffma2(acc[q], wa2, fmul2(av[q], sb2)); o[2 * q] = r.x; o[2 * q + 1] = r.y; }
      } else {
#pragma unroll
        for (int q = 0; q < NP; ++q) { const float2 r = fmul2(acc[q], wa2); o[2 * q] = r.x; o[2 * q + 1] = r.y; }
      }
      Vec16<T>::store(out + (g * K + a) * ldout + off, o);
    }
    __syncwarp();
    if (lane == 0) tma::mbar_arrive(empty + s);
    if (++s == NST) { s = 0; ph ^= 1; }
  }
}

template <typename T, int K, bool HAS_ADD>
static int launch_mean_mix_k(const T* x, int64_t ldx, const T* addend, int64_t ldadd, T* out, int64_t ldout, int64_t G, int64_t row, float sa,
                             float sb, cudaStream_t st) {
  using S = RingShape<T, K>;
  constexpr int STAGE = (HAS_ADD ? 2 : 1) * K * S::PIECE;
  constexpr int NST = (96 * 1024) / STAGE >= 6 ? 6 : ((96 * 1024) / STAGE >= 3 ? (96 * 1024) / STAGE : 3);
  constexpr size_t smem = 128 + (size_t)NST * STAGE;
  auto kern = k_mean_mix_ring<T, K, HAS_ADD, NST>;
  static DeviceOnce once;
  static int per_sm = 1;
  if (once.need()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_mean_mix_ring)");
    int occ = 1;          // (same answer on every B200: concurrent per-device threads store the same value)
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, RING_TPB + 32, smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor(k_mean_mix_ring)");
    per_sm = occ < 1 ? 1 : occ;
    once.mark();
  }
  const int64_t ipg = row / (RING_TPB * S::VEC);
  const int64_t items = G * ipg;
  int64_t blocks = (int64_t)per_sm * num_sms();
  if (blocks > items) blocks = items;
  if (blocks < 1) return 0;
  kern<<<(unsigned)blocks, RING_TPB + 32, smem, st>>>(x, ldx, addend, ldadd, out, ldout, G, (int)ipg, sa, sb);
  RGNN_LAUNCH_CHECK("k_mean_mix_ring");
  return 0;
}

template <typename T>
static int launch_mean_mix(int K, const T* x, int64_t ldx, const T* addend, int64_t ldadd, T* out, int64_t ldout, int64_t G, int64_t row,
                           float sa, float sb, cudaStream_t st) {
#define RGNN_CASE(KK)                                                                                              \
  case KK:                                                                                                         \
    return addend ? launch_mean_mix_k<T, KK, true>(x, ldx, addend, ldadd, out, ldout, G, row, sa, sb, st)          \
                  : launch_mean_mix_k<T, KK, false>(x, ldx, addend, ldadd, out, ldout, G, row, sa, sb, st);
  switch (K) { RGNN_CASE(2) RGNN_CASE(3) RGNN_CASE(4) RGNN_CASE(5) RGNN_CASE(6) RGNN_CASE(7) RGNN_CASE(8) default: return RGNN_E_UNSUPPORTED; }
#undef RGNN_CASE
}

// ---------------------------------------------------------------------------------------------------------------------
// backward:  dh[u] = addend[u] + w * sum_{a != u} gamma[e(u,a)] * dagg[a]
//            dz[e(u,a)] = (sig') * w * (sum_p h[u] dagg[a], sum_p dagg[a]),   w = 1/(K-1)
// Two consumer groups of TPB threads share every stage: group A produces dh (reads dagg, gamma, addend), group B produces
// dz (reads h, dagg; per-lane pixel partials, one reduce-scatter over the LPC lanes of a channel, sigmoid' in the epilogue).
// Splitting the work instead of the items halves the registers either group needs and doubles the warps that hide latency.
template <typename T, int K, int TPB> struct RingBwd {
  static constexpr int VEC = Vec16<T>::N;
  static constexpr int NE = K * (K - 1);
  static constexpr int LPC = RING_P / VEC;
  static constexpr int CH = TPB / LPC;
  static constexpr int PIECE = TPB * 16;
  static constexpr int ZPIECE = 2 * CH * (int)sizeof(T);
  static constexpr int NVP = (NE + LPC - 1) / LPC * LPC;
  __host__ __device__ static constexpr int stage(int nkind) { return nkind * K * PIECE + NE * ZPIECE; }
  static constexpr int HEAD = 128 + (8 * CH * K * 4 + 127) / 128 * 128;     // barriers + sum exchange (<= 8 stages)
};

template <typename T, int K, int TPB, int NST>
__global__ void __launch_bounds__(2 * TPB + 32, 1) k_film_bwd_ring(const T* __restrict__ h, int64_t ldh, const T* __restrict__ dagg, int64_t ldd,
                                                                   const T* __restrict__ z, const T* addend, int64_t ldadd,
                                                                   T* dh, int64_t lddh, T* __restrict__ dz, int64_t G, int F, int fuse_sig, const int2* __restrict__ gb) {
  using S = RingBwd<T, K, TPB>;
  constexpr int VEC = S::VEC, NE = S::NE, NP = VEC / 2, LPC = S::LPC, NVP = S::NVP;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + NST;
  uint64_t* sdfull = empty + NST;                          // group A -> group B: the per-channel sums of dagg are in sdx
  float* sdx = reinterpret_cast<float*>(smem + 128);        // [NST][CH][K]
  unsigned char* ring = smem + S::HEAD;
  const int tid = threadIdx.x, lane = tid & 31;
  const bool has_add = addend != nullptr;
  const int stage_bytes = S::stage(has_add ? 3 : 2);
  const int ipg = F / S::CH;
  const int64_t items = G * ipg;
  constexpr int OFF_D = K * S::PIECE, OFF_Z = 2 * K * S::PIECE, OFF_A = OFF_Z + NE * S::ZPIECE;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) { tma::mbar_init(full + s, 1); tma::mbar_init(empty + s, 2 * TPB / 32); tma::mbar_init(sdfull + s, TPB / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (tid >= 2 * TPB) {                                    // ---- producer warp
    int s = 0; uint32_t ph = 0;
    const int ncopy = (has_add ? 3 : 2) * K + NE;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
      const int64_t g = it / ipg;
      const int64_t nb = gb ? (int64_t)gb[g].x : g * K, eb = gb ? (int64_t)gb[g].y : g * NE;
      const int cb = (int)(it - g * ipg);
      const int64_t coff = (int64_t)cb * S::CH * RING_P;
      tma::mbar_wait(empty + s, ph ^ 1);
      unsigned char* st = ring + (size_t)s * stage_bytes;
      if (lane == 0) tma::mbar_expect_tx(full + s, stage_bytes);
      __syncwarp();
      for (int i = lane; i < ncopy; i += 32) {
        if (i < K) tma::bulk_g2s(st + i * S::PIECE, h + (nb + i) * ldh + coff, S::PIECE, full + s);
        else if (i < 2 * K) tma::bulk_g2s(st + OFF_D + (i - K) * S::PIECE, dagg + (nb + (i - K)) * ldd + coff, S::PIECE, full + s);
        else if (i < 2 * K + NE) tma::bulk_g2s(st + OFF_Z + (i - 2 * K) * S::ZPIECE, z + (eb + (i - 2 * K)) * 2 * (int64_t)F + 2 * cb * S::CH, S::ZPIECE, full + s);
        else tma::bulk_g2s(st + OFF_A + (i - 2 * K - NE) * S::PIECE, addend + (nb + (i - 2 * K - NE)) * ldadd + coff, S::PIECE, full + s);
      }
      if (++s == NST) { s = 0; ph ^= 1; }
    }
    return;
  }

  const float w = 1.0f / (K - 1);
  const float2 w2 = make_float2(w, w);
  const bool grp_b = tid >= TPB;
  const int t = grp_b ? tid - TPB : tid;                    // vector of the item this thread owns
  const int cl = t / LPC;
  int s = 0; uint32_t ph = 0;

  if (!grp_b) {                                            // ---- group A: dh
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
      const int64_t g = it / ipg;
      const int64_t nb = gb ? (int64_t)gb[g].x : g * K, eb = gb ? (int64_t)gb[g].y : g * NE;
      const int cb = (int)(it - g * ipg);
      const int64_t off = (int64_t)cb * S::CH * RING_P + t * VEC;
      const unsigned char* st = ring + (size_t)s * stage_bytes;
      tma::mbar_wait(full + s, ph);
      float2 dv[K][NP];
#pragma unroll
      for (int a = 0; a < K; ++a) unpack2<T>(*reinterpret_cast<const uint4*>(st + OFF_D + a * S::PIECE + t * 16), dv[a]);
      {   // pixel sums of dagg per (channel, destination) for group B's dbeta: lane partial -> LPC-lane reduction -> shared memory
        float sd[K];
#pragma unroll
        for (int a = 0; a < K; ++a) {
          float2 s2 = dv[a][0];
#pragma unroll
          for (int q = 1; q < NP; ++q) s2 = fadd2(s2, dv[a][q]);
          sd[a] = s2.x + s2.y;
        }
#pragma unroll
        for (int o = LPC / 2; o > 0; o >>= 1) {
#pragma unroll
          for (int a = 0; a < K; ++a) sd[a] += __shfl_xor_sync(0xffffffffu, sd[a], o);
        }
        if (t % LPC == 0) {
#pragma unroll
          for (int a = 0; a < K; ++a) sdx[(s * S::CH + cl) * K + a] = sd[a];
        }
        __syncwarp();
        if (lane == 0) tma::mbar_arrive(sdfull + s);
      }
      const unsigned char* zs = st + OFF_Z + cl * 2 * (int)sizeof(T);
#pragma unroll
      for (int u = 0; u < K; ++u) {
        float2 acc[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) acc[q] = make_float2(0.f, 0.f);
#pragma unroll
        for (int a = 0; a < K; ++a) {
          if (a == u) continue;
          const float gam = lds_pair<T>(zs + edge_id(K, u, a) * S::ZPIECE).x;
          const float2 g2 = make_float2(gam, gam);
#pragma unroll
          for (int q = 0; q < NP; ++q) acc[q] = ffma2(g2, dv[a][q], acc[q]);
        }
        if (has_add) {
          float2 av[NP];
          unpack2<T>(*reinterpret_cast<const uint4*>(st + OFF_A + u * S::PIECE + t * 16), av);
#pragma unroll
          for (int q = 0; q < NP; ++q) acc[q] = ffma2(acc[q], w2, av[q]);
        } else {
#pragma unroll
          for (int q = 0; q < NP; ++q) acc[q] = fmul2(acc[q], w2);
        }
        float o[VEC];
#pragma unroll
        for (int q = 0; q < NP; ++q) { o[2 * q] = acc[q].x; o[2 * q + 1] = acc[q].y; }
        Vec16<T>::store(dh + (nb + u) * lddh + off, o);
      }
      __syncwarp();
      if (lane == 0) tma::mbar_arrive(empty + s);
      if (++s == NST) { s = 0; ph ^= 1; }
    }
    return;
  }

  // ---- group B: dz
  const int l = t % LPC;
  int my_e[NVP / LPC], my_a[NVP / LPC];
#pragma unroll
  for (int i = 0; i < NVP / LPC; ++i) {
    const int e = scattered_id<NVP, LPC>(i, l);
    const int u = e / (K - 1), jj = e % (K - 1);
    my_e[i] = e < NE ? e : -1;
    my_a[i] = jj < u ? jj : jj + 1;
  }
  for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
    const int64_t g = it / ipg;
      const int64_t nb = gb ? (int64_t)gb[g].x : g * K, eb = gb ? (int64_t)gb[g].y : g * NE;
    const int cb = (int)(it - g * ipg);
    const unsigned char* st = ring + (size_t)s * stage_bytes;
    tma::mbar_wait(full + s, ph);
    float2 dv[K][NP];
#pragma unroll
    for (int a = 0; a < K; ++a) unpack2<T>(*reinterpret_cast<const uint4*>(st + OFF_D + a * S::PIECE + t * 16), dv[a]);
    float pg[NVP];
#pragma unroll
    for (int e = 0; e < NVP; ++e) pg[e] = 0.f;
#pragma unroll
    for (int u = 0; u < K; ++u) {
      float2 hv[NP];
      unpack2<T>(*reinterpret_cast<const uint4*>(st + u * S::PIECE + t * 16), hv);
#pragma unroll
      for (int a = 0; a < K; ++a) {
        if (a == u) continue;
        float2 s2 = fmul2(hv[0], dv[a][0]);
#pragma unroll
        for (int q = 1; q < NP; ++q) s2 = ffma2(hv[q], dv[a][q], s2);
        pg[edge_id(K, u, a)] = s2.x + s2.y;
      }
    }
    reduce_scatter<NVP, LPC>(pg, l);
    tma::mbar_wait(sdfull + s, ph);
    const float* sdc = sdx + (s * S::CH + cl) * K;
    const unsigned char* zs = st + OFF_Z + cl * 2 * (int)sizeof(T);
    T* dzc = dz + eb * 2 * (int64_t)F + 2 * (cb * S::CH + cl);
#pragma unroll
    for (int i = 0; i < NVP / LPC; ++i) {
      if (my_e[i] >= 0) {
        float dg = pg[i] * w, db = sdc[my_a[i]] * w;
        if (fuse_sig) {
          const float2 gb = lds_pair<T>(zs + my_e[i] * S::ZPIECE);
          dg *= gb.x * (1.f - gb.x); db *= gb.y * (1.f - gb.y);
        }
        st_pair<T>(dzc + my_e[i] * 2 * (int64_t)F, dg, db);
      }
    }
    __syncwarp();
    if (lane == 0) tma::mbar_arrive(empty + s);
    if (++s == NST) { s = 0; ph ^= 1; }
  }
}

template <typename T, int K>
static int launch_bwd_ring_k(const T* h, int64_t ldh, const T* dagg, int64_t ldd, const T* z, const T* addend, int64_t ldadd, T* dh, int64_t lddh,
                             T* dz, int64_t G, int F, int fuse_sig, cudaStream_t st, const int2* gb) {
  constexpr int TPB = K <= 5 ? 256 : 128;
  using S = RingBwd<T, K, TPB>;
  constexpr int NST = (200 * 1024) / S::stage(3) >= 4 ? 4 : ((200 * 1024) / S::stage(3) >= 3 ? 3 : 2);
  const size_t smem = S::HEAD + (size_t)NST * S::stage(addend ? 3 : 2);
  auto kern = k_film_bwd_ring<T, K, TPB, NST>;
  static DeviceOnce once;
  if (once.need()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::HEAD + NST * S::stage(3));
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_film_bwd_ring)");
    once.mark();
  }
  if (F % S::CH) return RGNN_E_UNSUPPORTED;
  const int64_t items = G * (F / S::CH);
  int64_t blocks = num_sms();
  if (blocks > items) blocks = items;
  if (blocks < 1) return 0;
  kern<<<(unsigned)blocks, 2 * TPB + 32, smem, st>>>(h, ldh, dagg, ldd, z, addend, ldadd, dh, lddh, dz, G, F, fuse_sig, gb);
  RGNN_LAUNCH_CHECK("k_film_bwd_ring");
  return 0;
}

template <typename T>
static int launch_bwd_ring(int K, const T* h, int64_t ldh, const T* dagg, int64_t ldd, const T* z, const T* addend, int64_t ldadd, T* dh,
                           int64_t lddh, T* dz, int64_t G, int F, int fuse_sig, cudaStream_t st, const int2* gb) {
#define RGNN_CASE(KK) case KK: return launch_bwd_ring_k<T, KK>(h, ldh, dagg, ldd, z, addend, ldadd, dh, lddh, dz, G, F, fuse_sig, st, gb);
  switch (K) { RGNN_CASE(2) RGNN_CASE(3) RGNN_CASE(4) RGNN_CASE(5) RGNN_CASE(6) RGNN_CASE(7) RGNN_CASE(8) default: return RGNN_E_UNSUPPORTED; }
#undef RGNN_CASE
}

static int g_ring = -1;
static bool ring_enabled() {
  if (g_ring < 0) { const char* e = getenv("RGNN_FILM_RING"); g_ring = (e && e[0] == '0') ? 0 : 1; }   // RGNN_FILM_RING=0: register-prefetch kernels
  return g_ring == 1;
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// true when the streaming variant can take the call (the caller falls back to film_complete.cu otherwise)
bool film_fwd_ring_ok(int dtype, int K, const void* x, int64_t ldx, const void* z, int64_t F, int64_t P) {
  const int es = dtype == RGNN_F32 ? 4 : 2;
  const int ch = dtype == RGNN_F32 ? RingShape<float, 2>::CH : RingShape<__nv_bfloat16, 2>::CH;
  return ring_enabled() && K >= 2 && K <= 8 && P == RING_P && F % ch == 0 && (ldx * es) % 16 == 0 && aligned16(x) && (!z || aligned16(z));
}

int film_fwd_ring(int dtype, int K, const void* x, int64_t ldx, const void* z, void* out, int64_t ldout, void* copy_to, int64_t G,
                  int64_t F, cudaStream_t st, const int2* gb) {
  if (dtype == RGNN_F32) {
    auto X = static_cast<const float*>(x); auto Z = static_cast<const float*>(z);
    return z ? launch_fwd_ring<float, true>(K, X, ldx, Z, static_cast<float*>(out), ldout, static_cast<float*>(copy_to), G, (int)F, st, gb)
             : launch_fwd_ring<float, false>(K, X, ldx, Z, static_cast<float*>(out), ldout, static_cast<float*>(copy_to), G, (int)F, st, gb);
  }
  auto X = static_cast<const __nv_bfloat16*>(x); auto Z = static_cast<const __nv_bfloat16*>(z);
  return z ? launch_fwd_ring<__nv_bfloat16, true>(K, X, ldx, Z, static_cast<__nv_bfloat16*>(out), ldout, static_cast<__nv_bfloat16*>(copy_to), G, (int)F, st, gb)
           : launch_fwd_ring<__nv_bfloat16, false>(K, X, ldx, Z, static_cast<__nv_bfloat16*>(out), ldout, static_cast<__nv_bfloat16*>(copy_to), G, (int)F, st, gb);
}

// Backward; returns RGNN_E_UNSUPPORTED when the shape / alignment is not covered (caller falls back to film_complete.cu).
int film_bwd_ring(int dtype, int K, const void* h, int64_t ldh, const void* dagg, int64_t ldd, const void* z, const void* addend,
                  int64_t ldadd, void* dh, int64_t lddh, void* dz, int64_t G, int64_t F, int64_t P, int fuse_sig, cudaStream_t st, const int2* gb) {
  const int es = dtype == RGNN_F32 ? 4 : 2;
  if (!ring_enabled() || K < 2 || K > 8 || P != RING_P || !aligned16(h) || !aligned16(dagg) || !aligned16(z) || (addend && !aligned16(addend)) ||
      (ldh * es) % 16 || (ldd * es) % 16 || (addend && (ldadd * es) % 16))
    return RGNN_E_UNSUPPORTED;
  if (dtype == RGNN_F32)
    return launch_bwd_ring<float>(K, static_cast<const float*>(h), ldh, static_cast<const float*>(dagg), ldd, static_cast<const float*>(z),
                                  static_cast<const float*>(addend), ldadd, static_cast<float*>(dh), lddh, static_cast<float*>(dz), G, (int)F, fuse_sig, st, gb);
  return launch_bwd_ring<__nv_bfloat16>(K, static_cast<const __nv_bfloat16*>(h), ldh, static_cast<const __nv_bfloat16*>(dagg), ldd,
                                        static_cast<const __nv_bfloat16*>(z), static_cast<const __nv_bfloat16*>(addend), ldadd,
                                        static_cast<__nv_bfloat16*>(dh), lddh, static_cast<__nv_bfloat16*>(dz), G, (int)F, fuse_sig, st, gb);
}

}  // namespace rgnn

// ---- ragged batches of complete graphs (a dropped camera in some samples; dgl/dataloader.py:88-99 builds each sample's graph from
// the cameras present): the batch is bucketed by robot count on the host and each bucket runs the compiled-topology ring kernel
// over ITS graphs, addressed through `bases` = int32 pairs (first node row, first edge row) - instead of sending the whole batch
// to the CSR-walking staged kernels (0.29 of HBM peak).
extern "C" int rgnn_film_bucket_forward(int dtype, const void* x, int64_t ldx, const void* z, void* out, int64_t ldout, void* copy_x_to,
                                        const int32_t* bases, int64_t bucket_graphs, int64_t robots, int64_t feat, int64_t pix, void* stream) {
  using namespace rgnn;
  RGNN_REQUIRE(dtype == RGNN_F32 || dtype == RGNN_BF16, RGNN_E_ARG, "rgnn_film_bucket_forward: bad dtype %d", dtype);
  RGNN_REQUIRE(bucket_graphs >= 0 && feat > 0 && pix > 0, RGNN_E_ARG, "rgnn_film_bucket_forward: bad sizes");
  if (bucket_graphs == 0) return 0;
  RGNN_REQUIRE(x && out && bases, RGNN_E_ARG, "rgnn_film_bucket_forward: null pointer");
  RGNN_REQUIRE((reinterpret_cast<uintptr_t>(bases) & 7u) == 0, RGNN_E_ALIGN, "rgnn_film_bucket_forward: bases must be 8-byte aligned");
  const int es = dtype == RGNN_F32 ? 4 : 2;
  if (!film_fwd_ring_ok(dtype, (int)robots, x, ldx, z, feat, pix) || !aligned16(out) || (ldout * es) % 16 || (copy_x_to && !aligned16(copy_x_to)))
    return RGNN_E_UNSUPPORTED;
  return film_fwd_ring(dtype, (int)robots, x, ldx, z, out, ldout, copy_x_to, bucket_graphs, feat, as_stream(stream),
                       reinterpret_cast<const int2*>(bases));
}

extern "C" int rgnn_film_bucket_backward(int dtype, const void* h, int64_t ldh, const void* dagg, int64_t lddagg, const void* z, int fuse_sigmoid_grad,
                                         const void* addend, int64_t ldadd, void* dh, int64_t lddh, void* dz, const int32_t* bases,
                                         int64_t bucket_graphs, int64_t robots, int64_t feat, int64_t pix, void* stream) {
  using namespace rgnn;
  RGNN_REQUIRE(dtype == RGNN_F32 || dtype == RGNN_BF16, RGNN_E_ARG, "rgnn_film_bucket_backward: bad dtype %d", dtype);
  RGNN_REQUIRE(bucket_graphs >= 0 && feat > 0 && pix > 0, RGNN_E_ARG, "rgnn_film_bucket_backward: bad sizes");
  if (bucket_graphs == 0) return 0;
  RGNN_REQUIRE(h && dagg && z && dh && dz && bases, RGNN_E_ARG, "rgnn_film_bucket_backward: null pointer");
  RGNN_REQUIRE((reinterpret_cast<uintptr_t>(bases) & 7u) == 0, RGNN_E_ALIGN, "rgnn_film_bucket_backward: bases must be 8-byte aligned");
  const int es = dtype == RGNN_F32 ? 4 : 2;
  if (!aligned16(dh) || (lddh * es) % 16 || !aligned16(dz)) return RGNN_E_UNSUPPORTED;
  return film_bwd_ring(dtype, (int)robots, h, ldh, dagg, lddagg, z, addend, ldadd, dh, lddh, dz, bucket_graphs, feat, pix, fuse_sigmoid_grad,
                       as_stream(stream), reinterpret_cast<const int2*>(bases));
}

extern "C" int rgnn_complete_mean_mix(int dtype, const void* x, int64_t ldx, const void* addend, int64_t ldadd, void* out, int64_t ldout,
                                      int64_t num_graphs, int64_t robots, int64_t row, float sa, float sb, void* stream) {
  using namespace rgnn;
  RGNN_REQUIRE(dtype == RGNN_F32 || dtype == RGNN_BF16, RGNN_E_ARG, "rgnn_complete_mean_mix: bad dtype %d", dtype);
  RGNN_REQUIRE(num_graphs >= 0 && robots >= 2 && row > 0, RGNN_E_ARG, "rgnn_complete_mean_mix: bad sizes");
  if (num_graphs == 0) return 0;
  RGNN_REQUIRE(x && out, RGNN_E_ARG, "rgnn_complete_mean_mix: null pointer");
  const int es = dtype == RGNN_F32 ? 4 : 2;
  const int64_t item = (int64_t)RING_TPB * (16 / es);
  if (robots > 8 || row % item || !aligned16(x) || !aligned16(out) || (addend && !aligned16(addend)) || (ldx * es) % 16 || (ldout * es) % 16 ||
      (addend && (ldadd * es) % 16) || row / item >= (1ll << 31)) {
    set_error("rgnn_complete_mean_mix: shape / alignment not covered (robots <= 8, row a multiple of %lld elements, 16-byte rows)", (long long)item);
    return RGNN_E_UNSUPPORTED;
  }
  cudaStream_t st = as_stream(stream);
  if (dtype == RGNN_F32)
    return launch_mean_mix<float>((int)robots, static_cast<const float*>(x), ldx, static_cast<const float*>(addend), ldadd, static_cast<float*>(out), ldout,
                                  num_graphs, row, sa, sb, st);
  return launch_mean_mix<__nv_bfloat16>((int)robots, static_cast<const __nv_bfloat16*>(x), ldx, static_cast<const __nv_bfloat16*>(addend), ldadd,
                                        static_cast<__nv_bfloat16*>(out), ldout, num_graphs, row, sa, sb, st);
}
