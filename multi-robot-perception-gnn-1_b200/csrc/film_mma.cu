// A5/A6/A10 for complete graphs with MANY robots (BASELINE configs[4]: K = 16 .. 64), bf16, P = 8x8: the per-channel K x K FiLM
// products on the tensor cores.
//
// For one graph and one channel c the FiLM-mean forward is a small dense product
//     out_c[a, p] = 1/(K-1) * ( sum_u Gamma_c[a, u] * x_c[u, p]  +  sum_u Beta_c[a, u] ),   Gamma_c[a, u] = gamma_{u->a, c}  (0 on the diagonal)
// i.e. [K x K] . [K x 64] per (graph, channel) - 63 FMAs per feature element at K = 64, which on CUDA cores is issue-bound
// (k_film_fwd_dense: 5.6 ms per launch at G = 128, F = 512 where the HBM traffic is worth 0.3 ms).  The backward is two more
// products of the same size:  dh_c = addend + 1/(K-1) Gamma_c^T . dAgg_c  and  dGamma_c = 1/(K-1) dAgg_c . x_c^T.
// These are tiny, differently-weighted GEMMs (one K x K matrix per channel, gathered from the edge-major z), far below a tcgen05
// tile, so they run as warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate): one CTA = (graph, 8 channels), one WARP per
// channel; the K robots' [8 ch x 64 px] slices are bulk-copied to shared memory (one 1 KB cp.async.bulk per robot, rows padded
// to 1040 B so the 8 rows of an ldmatrix hit 8 different bank groups), Gamma_c is scattered into a [K x K] bf16 matrix per channel
// (row pitch 2K + 16 B) while the beta row sums are accumulated with shared-memory atomics; fragments come from ldmatrix(.trans).
// fp32 batches keep the CUDA-core kernels of film_dense.cu (parity mode: the reference's rounding order).
#include <stdlib.h>
#include "common.cuh"
#include "film_common.cuh"

namespace rgnn {
namespace fmma {

constexpr int CT = 8;            // channels per CTA = warps per CTA
constexpr int XP = 1040;         // bytes per robot row in shared memory: 8 ch x 128 B + 16 B pad (65 x 16: odd)
constexpr int P = 64;

__host__ __device__ constexpr int gp_of(int K) { return 2 * K + 16; }      // bytes per Gamma row: (K/8 + 1) x 16, odd multiple for K % 16 == 0

__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// edge rows a thread keeps in flight: ALL of its rows at K = 64 (16 rows = 128 registers, free at that point of the kernels: the
// accumulators are not live yet / any more).  One CTA per SM and 8 warps leave nothing else to hide a gather's latency behind:
// ncu showed 54 % of all warp samples waiting on these loads with four rows in flight.
__host__ __device__ constexpr int rows_in_flight(int per) { return per >= 16 ? 16 : (per >= 8 ? 8 : 4); }

// Gamma_c[a][u] (bf16, zero diagonal) for the CTA's 8 channels from the edge-major z; optionally the beta row sums bs[c][a].
// thread = (destination a, segment of sources u); the edge rows (32 B per edge for the 8 channels) are fetched FOUR AT A TIME before
// any of them is scattered (a loop with one row in flight is latency-bound: 16 L2 round trips per CTA at K = 64), and the 8 beta
// sums stay in registers and reach shared memory with ONE atomic per (thread, channel).
template <bool WITH_BETA, int K>
__device__ __forceinline__ void build_gamma(unsigned char* gs, float* bs, const __nv_bfloat16* __restrict__ zc0, int64_t zrow) {
  constexpr int GP = gp_of(K), NSEG = 256 / K, PER = (K + NSEG - 1) / NSEG, NB = rows_in_flight(PER);
  for (int i = threadIdx.x; i < CT * K; i += blockDim.x) {
    const int c = i / K, a = i - c * K;
    *reinterpret_cast<__nv_bfloat16*>(gs + (size_t)c * K * GP + a * GP + a * 2) = __float2bfloat16_rn(0.f);
    if (WITH_BETA) bs[i] = 0.f;
  }
  __syncthreads();
  if constexpr (PER < 4) {           // few edges per thread (K = 16): one edge per thread and iteration, atomics per edge
    constexpr int NE = K * (K - 1);
    for (int e = threadIdx.x; e < NE; e += blockDim.x) {
      const int u = e / (K - 1), r = e - u * (K - 1), a = r < u ? r : r + 1;
      const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(zc0 + (int64_t)e * zrow));
      const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(zc0 + (int64_t)e * zrow + 8));
      const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        *reinterpret_cast<uint16_t*>(gs + (size_t)c * K * GP + a * GP + u * 2) = (uint16_t)(w[c] & 0xffffu);
        if (WITH_BETA) atomicAdd(bs + c * K + a, __uint_as_float(w[c] & 0xffff0000u));
      }
    }
    return;
  }
  const int a = threadIdx.x % K, seg = threadIdx.x / K;
  if (seg < NSEG) {
    float bsum[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) bsum[c] = 0.f;
    const int u0 = seg * PER;
#pragma unroll
    for (int b = 0; b < PER; b += NB) {
      uint4 v[NB][2];
      bool ok[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int u = u0 + b + j;
        ok[j] = (b + j < PER) && u < K && u != a;
        if (ok[j]) {
          const int64_t e = edge_id(K, u, a);
          v[j][0] = __ldg(reinterpret_cast<const uint4*>(zc0 + e * zrow));
          v[j][1] = __ldg(reinterpret_cast<const uint4*>(zc0 + e * zrow + 8));
        }
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        if (!ok[j]) continue;
        const int u = u0 + b + j;
        const uint32_t w[8] = {v[j][0].x, v[j][0].y, v[j][0].z, v[j][0].w, v[j][1].x, v[j][1].y, v[j][1].z, v[j][1].w};
#pragma unroll
        for (int c = 0; c < CT; ++c) {
          *reinterpret_cast<uint16_t*>(gs + (size_t)c * K * GP + a * GP + u * 2) = (uint16_t)(w[c] & 0xffffu);
          if (WITH_BETA) bsum[c] += __uint_as_float(w[c] & 0xffff0000u);
        }
      }
    }
    if (WITH_BETA) {
#pragma unroll
      for (int c = 0; c < CT; ++c) atomicAdd(bs + c * K + a, bsum[c]);
    }
  }
}

__device__ __forceinline__ void stage_rows(uint64_t* bar, unsigned char* dst, const __nv_bfloat16* src, int64_t ld, int64_t n0, int c0, int K) {
  for (int u = threadIdx.x; u < K; u += 32) tma::bulk_g2s(dst + (size_t)u * XP, src + (n0 + u) * ld + (int64_t)c0 * P, CT * P * 2, bar);
}

// ---------------------------------------------------------------------------------------------------------------------
template <int KT>      // K = 16 KT robots
__global__ void __launch_bounds__(256) k_film_fwd_mma(const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ z,
                                                      __nv_bfloat16* __restrict__ out, int64_t ldout, int F) {
  constexpr int K = 16 * KT, GP = gp_of(K);
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  unsigned char* xs = smem + 128;
  unsigned char* gs = xs + (size_t)K * XP;
  float* bs = reinterpret_cast<float*>(gs + (size_t)CT * K * GP);
  const int tiles_c = F / CT;
  const int64_t g = blockIdx.x / tiles_c;
  const int c0 = (int)(blockIdx.x % tiles_c) * CT;
  const int64_t n0 = g * K, e0 = g * (int64_t)K * (K - 1);
  if (threadIdx.x == 0) {
    tma::mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma::mbar_expect_tx(bar, (uint32_t)K * CT * P * 2);
  }
  __syncthreads();
  if (threadIdx.x < 32) stage_rows(bar, xs, x, ldx, n0, c0, K);
  build_gamma<true, K>(gs, bs, z + e0 * 2 * (int64_t)F + 2 * c0, 2 * (int64_t)F);
  __syncthreads();
  tma::mbar_wait(bar, 0);

  const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;         // warp = channel
  const uint32_t gbase = tma::smem_u32(gs + (size_t)c * K * GP), xbase = tma::smem_u32(xs + c * 128);
  float acc[KT][8][4];
#pragma unroll
  for (int mt = 0; mt < KT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
#pragma unroll
  for (int kt = 0; kt < KT; ++kt) {
    uint32_t b[4][4];                                                // B = x_c[u][p]: rows u (k), 8 p per 16 bytes -> ldmatrix.trans
    const int ub = kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) ldsm4t(xbase + ub * XP + (2 * j + (lane >> 4)) * 16, b[j]);
#pragma unroll
    for (int mt = 0; mt < KT; ++mt) {
      uint32_t a[4];                                                 // A = Gamma_c[a][u]: rows a, u contiguous
      ldsm4(gbase + (mt * 16 + (lane & 15)) * GP + (kt * 16 + (lane >> 4) * 8) * 2, a);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        mma16816(acc[mt][2 * j], a, b[j][0], b[j][1]);
        mma16816(acc[mt][2 * j + 1], a, b[j][2], b[j][3]);
      }
    }
  }
  const float w = 1.0f / (K - 1);
#pragma unroll
  for (int mt = 0; mt < KT; ++mt) {
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const int a = mt * 16 + (lane >> 2) + hrow * 8;
      const float bsum = bs[c * K + a];
      __nv_bfloat16* o = out + (n0 + a) * ldout + (int64_t)(c0 + c) * P + (lane & 3) * 2;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        *reinterpret_cast<__nv_bfloat162*>(o + nt * 8) = __floats2bfloat162_rn((acc[mt][nt][2 * hrow] + bsum) * w, (acc[mt][nt][2 * hrow + 1] + bsum) * w);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward: dh[u] = addend[u] + 1/(K-1) sum_a gamma_{u->a} dAgg[a];  dz_{u->a} = 1/(K-1) (sum_p h[u,p] dAgg[a,p], sum_p dAgg[a,p]) [* sigmoid']
template <int KT>
__global__ void __launch_bounds__(256) k_film_bwd_mma(const __nv_bfloat16* __restrict__ h, int64_t ldh, const __nv_bfloat16* __restrict__ dagg, int64_t ldd,
                                                      const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ addend, int64_t ldadd,
                                                      __nv_bfloat16* __restrict__ dh, int64_t lddh, __nv_bfloat16* __restrict__ dz, int F, int fuse_sig) {
  constexpr int K = 16 * KT, GP = gp_of(K);
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  unsigned char* xs = smem + 128;
  unsigned char* ds = xs + (size_t)K * XP;
  unsigned char* gs = ds + (size_t)K * XP;
  float* dbs = reinterpret_cast<float*>(gs + (size_t)CT * K * GP);
  const int tiles_c = F / CT;
  const int64_t g = blockIdx.x / tiles_c;
  const int c0 = (int)(blockIdx.x % tiles_c) * CT;
  const int64_t n0 = g * K, e0 = g * (int64_t)K * (K - 1);
  if (threadIdx.x == 0) {
    tma::mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma::mbar_expect_tx(bar, 2u * K * CT * P * 2);
  }
  __syncthreads();
  if (threadIdx.x < 32) { stage_rows(bar, xs, h, ldh, n0, c0, K); stage_rows(bar, ds, dagg, ldd, n0, c0, K); }
  const __nv_bfloat16* zc0 = z + e0 * 2 * (int64_t)F + 2 * c0;
  build_gamma<false, K>(gs, nullptr, zc0, 2 * (int64_t)F);
  __syncthreads();
  tma::mbar_wait(bar, 0);

  const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t gbase = tma::smem_u32(gs + (size_t)c * K * GP), xbase = tma::smem_u32(xs + c * 128), dbase = tma::smem_u32(ds + c * 128);
  const float w = 1.0f / (K - 1);
  // ---- dbeta_a = sum_p dAgg[a, p]: lane sums rows a = lane, lane + 32
  for (int a = lane; a < K; a += 32) {
    const uint4* row = reinterpret_cast<const uint4*>(ds + (size_t)a * XP + c * 128);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 v = row[j];
      const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) s += __uint_as_float(ws[i] << 16) + __uint_as_float(ws[i] & 0xffff0000u);
    }
    dbs[c * K + a] = s;
  }
  __syncwarp();
  // ---- dh_c[u][p] = addend + w * sum_a Gamma_c[a][u] dAgg_c[a][p]:  A = Gamma_c^T (ldmatrix.trans), B = dAgg_c (ldmatrix.trans)
  {
    float acc[KT][8][4];
#pragma unroll
    for (int mt = 0; mt < KT; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      uint32_t b[4][4];
      const int ab = kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
      for (int j = 0; j < 4; ++j) ldsm4t(dbase + ab * XP + (2 * j + (lane >> 4)) * 16, b[j]);
#pragma unroll
      for (int mt = 0; mt < KT; ++mt) {
        uint32_t a[4];       // A[m = u][k = a] = Gamma[a][u]: memory rows a, u contiguous -> transposed 8x8 blocks
        ldsm4t(gbase + (kt * 16 + (lane & 7) + ((lane >> 4) & 1) * 8) * GP + (mt * 16 + ((lane >> 3) & 1) * 8) * 2, a);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          mma16816(acc[mt][2 * j], a, b[j][0], b[j][1]);
          mma16816(acc[mt][2 * j + 1], a, b[j][2], b[j][3]);
        }
      }
    }
#pragma unroll
    for (int mt = 0; mt < KT; ++mt) {
      const int64_t col = (int64_t)(c0 + c) * P + (lane & 3) * 2;
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        const int u = mt * 16 + (lane >> 2) + hrow * 8;
        __nv_bfloat16* o = dh + (n0 + u) * lddh + col;
        const __nv_bfloat16* ad = addend ? addend + (n0 + u) * ldadd + col : nullptr;
        float2 t[8];                                                   // the row's 8 addend pairs are asked for together
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) t[nt] = ad ? ldg_pair<__nv_bfloat16>(ad + nt * 8) : make_float2(0.f, 0.f);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
          *reinterpret_cast<__nv_bfloat162*>(o + nt * 8) =
              __floats2bfloat162_rn(acc[mt][nt][2 * hrow] * w + t[nt].x, acc[mt][nt][2 * hrow + 1] * w + t[nt].y);
      }
    }
  }
  // ---- dGamma_c[a][u] = w * sum_p dAgg_c[a][p] x_c[u][p]:  A = dAgg_c (rows a, p contiguous), B[k = p][n = u] = x_c[u][p] (rows u, p contiguous)
  {
    float acc[KT][2 * KT][4];
#pragma unroll
    for (int mt = 0; mt < KT; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2 * KT; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {                                  // 64 pixels = 4 k-tiles
      uint32_t b[KT][4];                                               // b[j]: n-tiles 2j, 2j+1 (robots u)
#pragma unroll
      for (int j = 0; j < KT; ++j) ldsm4(xbase + ((2 * j + (lane >> 4)) * 8 + (lane & 7)) * XP + (kt * 16 + ((lane >> 3) & 1) * 8) * 2, b[j]);
#pragma unroll
      for (int mt = 0; mt < KT; ++mt) {
        uint32_t a[4];
        ldsm4(dbase + (mt * 16 + (lane & 15)) * XP + (kt * 16 + (lane >> 4) * 8) * 2, a);
#pragma unroll
        for (int j = 0; j < KT; ++j) {
          mma16816(acc[mt][2 * j], a, b[j][0], b[j][1]);
          mma16816(acc[mt][2 * j + 1], a, b[j][2], b[j][3]);
        }
      }
    }
    // dgamma (with its sigmoid' factor) replaces gamma IN PLACE in the channel's shared-memory matrix (the thread that holds the
    // accumulator element reads and rewrites the same bf16 cell); the edge rows then leave below as whole 32-byte sectors.
    // (Writing the 4-byte (gamma, beta) pairs straight from the fragments costs 8 partial-sector requests per edge row.)
    unsigned char* gc = gs + (size_t)c * K * GP;
#pragma unroll
    for (int mt = 0; mt < KT; ++mt) {
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        const int a = mt * 16 + (lane >> 2) + hrow * 8;
#pragma unroll
        for (int nt = 0; nt < 2 * KT; ++nt) {
          const int u0 = nt * 8 + (lane & 3) * 2;
          uint32_t* cell = reinterpret_cast<uint32_t*>(gc + a * GP + u0 * 2);          // (u0, u0 + 1): one aligned 4-byte word
          const uint32_t gw2 = *cell;
          float d0 = acc[mt][nt][2 * hrow] * w, d1 = acc[mt][nt][2 * hrow + 1] * w;
          if (fuse_sig) {
            const float g0 = __uint_as_float(gw2 << 16), g1 = __uint_as_float(gw2 & 0xffff0000u);
            d0 *= g0 * (1.f - g0);
            d1 *= g1 * (1.f - g1);
          }
          __nv_bfloat162 o2 = __floats2bfloat162_rn(d0, d1);
          *cell = *reinterpret_cast<uint32_t*>(&o2);
        }
      }
    }
  }
  __syncthreads();
  {
    __nv_bfloat16* dz0 = dz + e0 * 2 * (int64_t)F + 2 * c0;
    const int64_t zrow = 2 * (int64_t)F;
    constexpr int NSEG = 256 / K, PER = (K + NSEG - 1) / NSEG, NB = rows_in_flight(PER);
    if constexpr (PER < 4) {
      constexpr int NE = K * (K - 1);
      for (int e = threadIdx.x; e < NE; e += blockDim.x) {
        const int u = e / (K - 1), r = e - u * (K - 1), a = r < u ? r : r + 1;
        uint32_t bw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (fuse_sig) {
          const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(zc0 + (int64_t)e * zrow));
          const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(zc0 + (int64_t)e * zrow + 8));
          bw[0] = v0.x; bw[1] = v0.y; bw[2] = v0.z; bw[3] = v0.w; bw[4] = v1.x; bw[5] = v1.y; bw[6] = v1.z; bw[7] = v1.w;
        }
        uint32_t ow[8];
#pragma unroll
        for (int cc = 0; cc < CT; ++cc) {
          const uint32_t dgb = *reinterpret_cast<const uint16_t*>(gs + (size_t)cc * K * GP + a * GP + u * 2);
          float dbt = dbs[cc * K + a] * w;
          if (fuse_sig) {
            const float bet = __uint_as_float(bw[cc] & 0xffff0000u);
            dbt *= bet * (1.f - bet);
          }
          __nv_bfloat16 db16 = __float2bfloat16_rn(dbt);
          ow[cc] = dgb | ((uint32_t)(*reinterpret_cast<uint16_t*>(&db16)) << 16);
        }
        uint4* dst = reinterpret_cast<uint4*>(dz0 + (int64_t)e * zrow);
        dst[0] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        dst[1] = make_uint4(ow[4], ow[5], ow[6], ow[7]);
      }
      return;
    }
    const int a = threadIdx.x % K, seg = threadIdx.x / K;
    if (seg < NSEG) {
      const int u0 = seg * PER;
      float dbw[CT];
#pragma unroll
      for (int cc = 0; cc < CT; ++cc) dbw[cc] = dbs[cc * K + a] * w;
#pragma unroll
      for (int b = 0; b < PER; b += NB) {          // as in build_gamma
        uint4 v[NB][2];
        bool ok[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const int u = u0 + b + j;
          ok[j] = (b + j < PER) && u < K && u != a;
          if (ok[j] && fuse_sig) {
            const int64_t e = edge_id(K, u, a);
            v[j][0] = __ldg(reinterpret_cast<const uint4*>(zc0 + e * zrow));
            v[j][1] = __ldg(reinterpret_cast<const uint4*>(zc0 + e * zrow + 8));
          }
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          if (!ok[j]) continue;
          const int u = u0 + b + j;
          const int64_t e = edge_id(K, u, a);
          const uint32_t bw[8] = {v[j][0].x, v[j][0].y, v[j][0].z, v[j][0].w, v[j][1].x, v[j][1].y, v[j][1].z, v[j][1].w};
          uint32_t ow[8];
#pragma unroll
          for (int cc = 0; cc < CT; ++cc) {
            const uint32_t dgb = *reinterpret_cast<const uint16_t*>(gs + (size_t)cc * K * GP + a * GP + u * 2);
            float dbt = dbw[cc];
            if (fuse_sig) {
              const float bet = __uint_as_float(bw[cc] & 0xffff0000u);
              dbt *= bet * (1.f - bet);
            }
            __nv_bfloat16 db16 = __float2bfloat16_rn(dbt);
            ow[cc] = dgb | ((uint32_t)(*reinterpret_cast<uint16_t*>(&db16)) << 16);
          }
          uint4* dst = reinterpret_cast<uint4*>(dz0 + e * zrow);
          dst[0] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
          dst[1] = make_uint4(ow[4], ow[5], ow[6], ow[7]);
        }
      }
    }
  }
}

static bool enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("RGNN_FILM_MMA"); on = (e && e[0] == '0') ? 0 : 1; }
  return on == 1;
}
static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <int KT>
static int fwd_t(const __nv_bfloat16* x, int64_t ldx, const __nv_bfloat16* z, __nv_bfloat16* out, int64_t ldout, int64_t G, int F, cudaStream_t st) {
  constexpr int K = 16 * KT;
  const size_t smem = 128 + (size_t)K * XP + (size_t)CT * K * gp_of(K) + (size_t)CT * K * 4;
  static DeviceOnce once;
  if (once.need()) {
    cudaError_t e = cudaFuncSetAttribute(k_film_fwd_mma<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_film_fwd_mma)");
    once.mark();
  }
  k_film_fwd_mma<KT><<<(unsigned)(G * (F / CT)), 256, smem, st>>>(x, ldx, z, out, ldout, F);
  RGNN_LAUNCH_CHECK("k_film_fwd_mma");
  return 0;
}

template <int KT>
static int bwd_t(const __nv_bfloat16* h, int64_t ldh, const __nv_bfloat16* dagg, int64_t ldd, const __nv_bfloat16* z, const __nv_bfloat16* addend,
                 int64_t ldadd, __nv_bfloat16* dh, int64_t lddh, __nv_bfloat16* dz, int64_t G, int F, int fuse_sig, cudaStream_t st) {
  constexpr int K = 16 * KT;
  const size_t smem = 128 + 2 * (size_t)K * XP + (size_t)CT * K * gp_of(K) + (size_t)CT * K * 4;
  static DeviceOnce once;
  if (once.need()) {
    cudaError_t e = cudaFuncSetAttribute(k_film_bwd_mma<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_film_bwd_mma)");
    once.mark();
  }
  k_film_bwd_mma<KT><<<(unsigned)(G * (F / CT)), 256, smem, st>>>(h, ldh, dagg, ldd, z, addend, ldadd, dh, lddh, dz, F, fuse_sig);
  RGNN_LAUNCH_CHECK("k_film_bwd_mma");
  return 0;
}

}  // namespace fmma

// bf16, P = 64, K in {16, 32, 48, 64}, FiLM (z given), 16-byte addressable rows; anything else: RGNN_E_UNSUPPORTED (CUDA-core kernels)
int film_fwd_mma(int dtype, int K, const void* x, int64_t ldx, const void* z, void* out, int64_t ldout, void* copy_to, int64_t G, int64_t F,
                 int64_t P, cudaStream_t st) {
  using namespace fmma;
  if (!enabled() || dtype != RGNN_BF16 || P != 64 || K % 16 != 0 || K < 16 || K > 64 || !z || copy_to || F % CT != 0 || ldx % 8 != 0 || ldout % 2 != 0 ||
      !al16(x) || !al16(z) || (reinterpret_cast<uintptr_t>(out) & 3u) || G * (F / CT) >= (1ll << 31) || G < 1)
    return RGNN_E_UNSUPPORTED;
  auto X = static_cast<const __nv_bfloat16*>(x); auto Z = static_cast<const __nv_bfloat16*>(z); auto O = static_cast<__nv_bfloat16*>(out);
  switch (K / 16) {
    case 1: return fwd_t<1>(X, ldx, Z, O, ldout, G, (int)F, st);
    case 2: return fwd_t<2>(X, ldx, Z, O, ldout, G, (int)F, st);
    case 3: return fwd_t<3>(X, ldx, Z, O, ldout, G, (int)F, st);
    default: return fwd_t<4>(X, ldx, Z, O, ldout, G, (int)F, st);
  }
}

int film_bwd_mma(int dtype, int K, const void* h, int64_t ldh, const void* dagg, int64_t ldd, const void* z, const void* addend, int64_t ldadd,
                 void* dh, int64_t lddh, void* dz, int64_t G, int64_t F, int64_t P, int fuse_sig, cudaStream_t st) {
  using namespace fmma;
  if (!enabled() || dtype != RGNN_BF16 || P != 64 || K % 16 != 0 || K < 16 || K > 64 || F % CT != 0 || ldh % 8 != 0 || ldd % 8 != 0 ||
      (addend && ldadd % 2 != 0) || lddh % 2 != 0 || !al16(h) || !al16(dagg) || !al16(z) || (reinterpret_cast<uintptr_t>(dh) & 3u) ||
      !al16(dz) || (addend && (reinterpret_cast<uintptr_t>(addend) & 3u)) || G * (F / CT) >= (1ll << 31) || G < 1)
    return RGNN_E_UNSUPPORTED;
  auto H = static_cast<const __nv_bfloat16*>(h); auto D = static_cast<const __nv_bfloat16*>(dagg); auto Z = static_cast<const __nv_bfloat16*>(z);
  auto A = static_cast<const __nv_bfloat16*>(addend); auto DH = static_cast<__nv_bfloat16*>(dh); auto DZ = static_cast<__nv_bfloat16*>(dz);
  switch (K / 16) {
    case 1: return bwd_t<1>(H, ldh, D, ldd, Z, A, ldadd, DH, lddh, DZ, G, (int)F, fuse_sig, st);
    case 2: return bwd_t<2>(H, ldh, D, ldd, Z, A, ldadd, DH, lddh, DZ, G, (int)F, fuse_sig, st);
    case 3: return bwd_t<3>(H, ldh, D, ldd, Z, A, ldadd, DH, lddh, DZ, G, (int)F, fuse_sig, st);
    default: return bwd_t<4>(H, ldh, D, ldd, Z, A, ldadd, DH, lddh, DZ, G, (int)F, fuse_sig, st);
  }
}

}  // namespace rgnn
