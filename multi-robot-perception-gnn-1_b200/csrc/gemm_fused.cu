// A5/A6 + A8 fused (BASELINE.json north_star (c)): FiLM-mean aggregation in the PROLOGUE of the compress GEMM.
//
//   y_v = Wc[:, :F] . h_v + Wc[:, F:] . g_v + b ,   g_a = mean_{u != a} (gamma_{u->a} * h_u + beta_{u->a})
//   (dgl/model/models.py:181-184: gcn1(g) -> cat -> conv1)
//
// One CTA owns (graph g, 128 output channels).  Per k-block of 32 input channels the TMA producer stages
//   Wc[:, k-block] and Wc[:, F + k-block]   (two K-major SWIZZLE_64B A tiles, 8 KB each),
//   h[g's K robots, k-block, :]            (one MN-major SWIZZLE_128B B tile, K x 4 KB: a robot's [32 ch x 64 px] slab),
//   z[g's K(K-1) edges, k-block]           (gamma/beta pairs, K(K-1) x 128 B),
// the MMA issuer immediately multiplies the h half (tcgen05.mma into TMEM, N = 64 K columns), while 8 FiLM warps compute the
// aggregated tile g from the SAME staged h tile, writing it into shared memory in exactly the swizzled operand layout TMA would
// have produced (the FiLM map is element-wise across robots, so source and destination offsets inside a robot's slab are
// identical - the swizzle is transparent).  The issuer then multiplies the g half from that tile.  The aggregated features
// reach HBM only if the caller needs them for the backward (one TMA store per tile from the operand buffer); in eval they
// never leave the SM.  HBM traffic per layer forward: h once + y once (+ g once when training) instead of
// h twice + g write + g read + y.
//
// 512 threads: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-11 FiLM producers, warps 12-15 epilogue.
#include "tc_common.cuh"

namespace rgnn {
namespace fused {

using namespace rgnn::tc;

constexpr int BM = 128;         // output channels per CTA tile
constexpr int BKF = 32;         // input channels per stage
constexpr int NTHR = 512;

struct FusedParams {
  int tiles_m;                  // Fo / 128
  int64_t graphs;
  int F, Fo;                    // input channels per half, output channels
  int kblocks;                  // F / 32
  const __nv_bfloat16* z;       // unused on device (z arrives through its tensor map); kept for debugging
  const float* bias;
  int store_g;                  // also write the aggregated features (needed by the backward)
};

template <int KR> struct Cfg {
  static constexpr int NE = KR * (KR - 1);
  static constexpr int NCOL = 64 * KR;                                   // accumulator columns
  static constexpr int N1 = NCOL <= 256 ? NCOL : (KR == 5 || KR == 6 ? 192 : 256);   // first MMA's N (multiple of 64)
  static constexpr int N2 = NCOL - N1;                                   // second MMA's N (0, 128, 192 or 256)
  static constexpr uint32_t A_BYTES = BM * BKF * 2;                      // 8 KB, SWIZZLE_64B
  static constexpr uint32_t H_BYTES = KR * BKF * 128;                    // K x 4 KB, SWIZZLE_128B
  static constexpr uint32_t Z_BYTES = ((NE * 128 + 1023) / 1024) * 1024; // keeps the next stage 1024-aligned
  static constexpr uint32_t STAGE = 2 * A_BYTES + 2 * H_BYTES + Z_BYTES;
  static constexpr int NST = (3 * STAGE + 2 * 16384 + 2048 <= 227 * 1024) ? 3 : 2;
  static constexpr uint32_t TX = 2 * A_BYTES + H_BYTES + NE * 128;       // bytes the producer's TMA loads deliver per stage
  static constexpr size_t SMEM = 1024 + (size_t)NST * STAGE + 2 * 16384 + 256;
};

template <int KR>
__global__ void __launch_bounds__(NTHR, 1)
k_conv_film_fused(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH,
                  const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmY,
                  const __grid_constant__ CUtensorMap tmG, const FusedParams p) {
  using C = Cfg<KR>;
  constexpr int NST = C::NST;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t base = smem_u32(smem);
  const uint32_t out_stage = base + NST * C::STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NST * C::STAGE + 2 * 16384);
  const uint32_t bar_full = smem_u32(bars), bar_gfull = smem_u32(bars + NST), bar_empty = smem_u32(bars + 2 * NST);
  const uint32_t bar_tfull = smem_u32(bars + 3 * NST), bar_tempty = smem_u32(bars + 3 * NST + 1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * NST + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmH) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmZ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_gfull + 8 * i, 8); mbar_init(bar_empty + 8 * i, 1); }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tempty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const int64_t work_total = p.graphs * p.tiles_m;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int64_t w = blockIdx.x; w < work_total; w += gridDim.x) {
        const int tm = (int)(w % p.tiles_m);
        const int64_t g = w / p.tiles_m;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t fb = bar_full + 8 * stage;
          const uint32_t s0 = base + stage * C::STAGE;
          mbar_expect_tx(fb, C::TX);
          tma_2d(s0, &tmW, fb, kb * BKF, tm * BM);                               // Wc[:, k-block]
          tma_2d(s0 + C::A_BYTES, &tmW, fb, p.F + kb * BKF, tm * BM);            // Wc[:, F + k-block]
          tma_3d(s0 + 2 * C::A_BYTES, &tmH, fb, 0, kb * BKF, (int)(g * KR));     // h of the graph's robots
          tma_2d(s0 + 2 * C::A_BYTES + 2 * C::H_BYTES, &tmZ, fb, 2 * kb * BKF, (int)(g * C::NE));   // gamma/beta of its edges
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(C::N1 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)((C::N2 > 0 ? C::N2 : 64) >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    int stage = 0; uint32_t phase = 0; uint32_t tphase = 0;
    for (int64_t w = blockIdx.x; w < work_total; w += gridDim.x) {
      mbar_wait(bar_tempty, tphase ^ 1);                 // the single accumulator has been drained
      tphase ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int kb = 0; kb < p.kblocks; ++kb) {
        const uint32_t s0 = base + stage * C::STAGE;
        mbar_wait(bar_full + 8 * stage, phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BKF / 16; ++k) {           // h half:  D += Wc[:, kb] . h
            const uint64_t da = make_desc(s0 + k * 32, 0, 512, 4);
            const uint32_t hb = s0 + 2 * C::A_BYTES + k * 2048;
            umma(tmem_base, da, make_desc(hb, BKF * 128, 1024), idesc1, (kb > 0 || k > 0) ? 1u : 0u);
            if (C::N2 > 0) umma(tmem_base + C::N1, da, make_desc(hb + (C::N1 / 64) * (BKF * 128), BKF * 128, 1024), idesc2, (kb > 0 || k > 0) ? 1u : 0u);
          }
        }
        __syncwarp();
        mbar_wait(bar_gfull + 8 * stage, phase);         // the FiLM warps have written the aggregated tile
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BKF / 16; ++k) {           // aggregated half:  D += Wc[:, F + kb] . g
            const uint64_t da = make_desc(s0 + C::A_BYTES + k * 32, 0, 512, 4);
            const uint32_t gb = s0 + 2 * C::A_BYTES + C::H_BYTES + k * 2048;
            umma(tmem_base, da, make_desc(gb, BKF * 128, 1024), idesc1, 1u);
            if (C::N2 > 0) umma(tmem_base + C::N1, da, make_desc(gb + (C::N1 / 64) * (BKF * 128), BKF * 128, 1024), idesc2, 1u);
          }
          umma_commit(bar_empty + 8 * stage);
          if (kb == p.kblocks - 1) umma_commit(bar_tfull);
        }
        __syncwarp();
        if (++stage == NST) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ================================ FiLM producers (256 threads) ================================
    const int ft = threadIdx.x - 128;                    // 0..255: one (channel, 16-byte pixel chunk) position of the 32 x 8 grid
    const int c = ft >> 3, q = ft & 7;
    const uint32_t voff = c * 128 + ((q ^ (c & 7)) << 4);   // swizzled offset inside a robot's [32 ch x 128 B] slab
    const bool issuer = ft == 0;
    const float inv = 1.0f / (KR - 1);
    int stage = 0; uint32_t phase = 0;
    for (int64_t w = blockIdx.x; w < work_total; w += gridDim.x) {
      const int tm = (int)(w % p.tiles_m);
      const int64_t g = w / p.tiles_m;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        const uint32_t s0 = base + stage * C::STAGE;
        const uint32_t hs = s0 + 2 * C::A_BYTES, gs = hs + C::H_BYTES, zs = gs + C::H_BYTES;
        if (p.store_g) {
          // the aggregated tile written NST k-blocks ago into this buffer must have been read by its TMA store
          if (issuer) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NST - 1) : "memory");
          named_bar_sync(3, 256);
        }
        mbar_wait(bar_full + 8 * stage, phase);
        float xv[KR][8];
#pragma unroll
        for (int u = 0; u < KR; ++u) {
          uint32_t a0, a1, a2, a3;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(hs + u * (BKF * 128) + voff));
          const uint32_t ws[4] = {a0, a1, a2, a3};
#pragma unroll
          for (int i = 0; i < 4; ++i) { xv[u][2 * i] = __uint_as_float(ws[i] << 16); xv[u][2 * i + 1] = __uint_as_float(ws[i] & 0xffff0000u); }
        }
#pragma unroll
        for (int a = 0; a < KR; ++a) {
          float acc[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = 0.f;
          float bsum = 0.f;
#pragma unroll
          for (int u = 0; u < KR; ++u) {
            if (u == a) continue;
            const int e = u * (KR - 1) + (a < u ? a : a - 1);
            uint32_t gw;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(gw) : "r"(zs + e * 128 + c * 4));
            const float gam = __uint_as_float(gw << 16), bet = __uint_as_float(gw & 0xffff0000u);
            bsum += bet;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(gam, xv[u][i], acc[i]);
          }
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = pack_bf16((acc[2 * i] + bsum) * inv, (acc[2 * i + 1] + bsum) * inv);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(gs + a * (BKF * 128) + voff), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to tcgen05 / TMA
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_gfull + 8 * stage);
        if (p.store_g && tm == 0) {                       // every M-tile computes the same g: one of them writes it out
          named_bar_sync(3, 256);
          if (issuer) {
            tma_store_3d(&tmG, gs, 0, kb * BKF, (int)(g * KR));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        } else if (p.store_g && issuer) {
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");       // keep the group count in step with the wait above
        }
        if (++stage == NST) { stage = 0; phase ^= 1; }
      }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  } else if (warp >= 12) {
    // ================================ epilogue (128 threads) ================================
    const int ew = warp & 3;
    const int row = ew * 32 + lane;
    uint32_t tphase = 0;
    for (int64_t w = blockIdx.x; w < work_total; w += gridDim.x) {
      const int tm = (int)(w % p.tiles_m);
      const int64_t g = w / p.tiles_m;
      mbar_wait(bar_tfull, tphase);
      tphase ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tacc = tmem_base + ((uint32_t)(ew * 32) << 16);
      const float bias_m = p.bias ? p.bias[tm * BM + row] : 0.f;
#pragma unroll 1
      for (int j = 0; j < KR; ++j) {
        const uint32_t stg = out_stage + (j & 1) * 16384;
        if (ew == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        named_bar_sync(2, 128);
        uint32_t r0[32], r1[32];
        tmem_ld32_issue(tacc + (2 * j) * 32, r0);
        tmem_ld32_issue(tacc + (2 * j + 1) * 32, r1);
        tmem_ld_wait();
        if (j == KR - 1) {                               // accumulator fully read: the issuer may start the next tile
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty);
        }
#pragma unroll
        for (int qq = 0; qq < 8; ++qq) {
          const uint32_t* r = qq < 4 ? r0 : r1;
          const int b = (qq & 3) * 8;
          const uint32_t w0 = pack_bf16(__uint_as_float(r[b + 0]) + bias_m, __uint_as_float(r[b + 1]) + bias_m);
          const uint32_t w1 = pack_bf16(__uint_as_float(r[b + 2]) + bias_m, __uint_as_float(r[b + 3]) + bias_m);
          const uint32_t w2 = pack_bf16(__uint_as_float(r[b + 4]) + bias_m, __uint_as_float(r[b + 5]) + bias_m);
          const uint32_t w3 = pack_bf16(__uint_as_float(r[b + 6]) + bias_m, __uint_as_float(r[b + 7]) + bias_m);
          const uint32_t addr = stg + row * 128 + ((qq ^ (row & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        named_bar_sync(2, 128);
        if (ew == 0 && lane == 0) {
          tma_store_2d(&tmY, stg, 0, (int)((g * KR + j) * p.Fo + tm * BM));
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (ew == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

template <int KR>
static int launch_fused(const void* W, const void* h, const void* z, const float* bias, void* y, void* gout, int64_t graphs, int64_t F,
                        int64_t Fo, cudaStream_t st) {
  using C = Cfg<KR>;
  const int64_t nodes = graphs * KR, edges = graphs * C::NE;
  CUtensorMap tw, th, tz, ty, tg;
  {
    const uint64_t d[2] = {(uint64_t)(2 * F), (uint64_t)Fo}, s[1] = {(uint64_t)(2 * F) * 2};
    const uint32_t b[2] = {BKF, BM};
    int rc = make_map(&tw, W, 2, d, s, b, CU_TENSOR_MAP_SWIZZLE_64B); if (rc) return rc;
  }
  {
    const uint64_t d[3] = {64, (uint64_t)F, (uint64_t)nodes}, s[2] = {128, (uint64_t)F * 128};
    const uint32_t b[3] = {64, BKF, KR};
    int rc = make_map(&th, h, 3, d, s, b); if (rc) return rc;
    rc = make_map(&tg, gout ? gout : h, 3, d, s, b); if (rc) return rc;
  }
  {
    const uint64_t d[2] = {(uint64_t)(2 * F), (uint64_t)edges}, s[1] = {(uint64_t)(2 * F) * 2};
    const uint32_t b[2] = {2 * BKF, (uint32_t)C::NE};
    int rc = make_map(&tz, z, 2, d, s, b, CU_TENSOR_MAP_SWIZZLE_NONE); if (rc) return rc;
  }
  {
    const uint64_t d[2] = {64, (uint64_t)(nodes * Fo)}, s[1] = {128};
    const uint32_t b[2] = {64, 128};
    int rc = make_map(&ty, y, 2, d, s, b); if (rc) return rc;
  }
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_film_fused<KR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_conv_film_fused)");
    attr_done = true;
  }
  FusedParams p = {};
  p.tiles_m = (int)(Fo / BM); p.graphs = graphs; p.F = (int)F; p.Fo = (int)Fo; p.kblocks = (int)(F / BKF);
  p.bias = bias; p.store_g = gout != nullptr;
  const int64_t work = graphs * p.tiles_m;
  const int grid = (int)(work < kNumSMs ? work : kNumSMs);
  if (grid < 1) return 0;
  k_conv_film_fused<KR><<<grid, NTHR, C::SMEM, st>>>(tw, th, tz, ty, tg, p);
  RGNN_LAUNCH_CHECK("k_conv_film_fused");
  return 0;
}

}  // namespace fused
}  // namespace rgnn

using namespace rgnn;

// y[v] = W[:, :F] h[v] + W[:, F:] g[v] + bias,  g = FiLM-mean(h, z) over G complete `robots`-node graphs (reference edge order);
// h [G*K, F, 64], z [G*K(K-1), 2F], W [Fo, 2F], y [G*K, Fo, 64] all bf16; g_out (nullable) receives the aggregated features.
extern "C" int rgnn_tc_conv_film_fused(const void* W, const void* h, const void* z, const float* bias, void* y, void* g_out,
                                       int64_t graphs, int64_t robots, int64_t F, int64_t Fo, void* stream) {
  RGNN_REQUIRE(W && h && z && y && graphs >= 0 && F > 0 && Fo > 0, RGNN_E_ARG, "rgnn_tc_conv_film_fused: bad arguments");
  RGNN_REQUIRE(robots >= 2 && robots <= 8, RGNN_E_UNSUPPORTED, "rgnn_tc_conv_film_fused: 2..8 robots per graph");
  RGNN_REQUIRE(F % 32 == 0 && Fo % 128 == 0, RGNN_E_UNSUPPORTED, "rgnn_tc_conv_film_fused: F %% 32 and Fo %% 128 required");
  RGNN_REQUIRE(tc::al16(W) && tc::al16(h) && tc::al16(z) && tc::al16(y) && (!g_out || tc::al16(g_out)), RGNN_E_ALIGN, "rgnn_tc_conv_film_fused: 16-byte alignment");
  RGNN_REQUIRE(graphs * robots * (Fo > F ? Fo : F) < (1ll << 31) && graphs * robots * (robots - 1) < (1ll << 31), RGNN_E_RANGE,
               "rgnn_tc_conv_film_fused: too many rows for a tensor-map coordinate");
  if (graphs == 0) return 0;
  cudaStream_t st = as_stream(stream);
  switch (robots) {
    case 2: return fused::launch_fused<2>(W, h, z, bias, y, g_out, graphs, F, Fo, st);
    case 3: return fused::launch_fused<3>(W, h, z, bias, y, g_out, graphs, F, Fo, st);
    case 4: return fused::launch_fused<4>(W, h, z, bias, y, g_out, graphs, F, Fo, st);
    case 5: return fused::launch_fused<5>(W, h, z, bias, y, g_out, graphs, F, Fo, st);
    case 6: return fused::launch_fused<6>(W, h, z, bias, y, g_out, graphs, F, Fo, st);
    case 7: return fused::launch_fused<7>(W, h, z, bias, y, g_out, graphs, F, Fo, st);
    default: return fused::launch_fused<8>(W, h, z, bias, y, g_out, graphs, F, Fo, st);
  }
}
