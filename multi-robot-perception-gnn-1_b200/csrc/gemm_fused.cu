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
  int debug;                    // RGNN_FUSED_DEBUG bit mask (experiments only): 1 = FiLM warps skip the arithmetic, 2 = epilogue skips stores
};

template <int KR> struct Cfg {
  static constexpr int NE = KR * (KR - 1);
  static constexpr int NCOL = 64 * KR;                                   // accumulator columns
  static constexpr int N1 = NCOL <= 256 ? NCOL : (KR == 5 || KR == 6 ? 192 : 256);   // first MMA's N (multiple of 64)
  static constexpr int N2 = NCOL - N1;                                   // second MMA's N (0, 128, 192 or 256)
  static constexpr uint32_t A_BYTES = BM * BKF * 2;                      // 8 KB, SWIZZLE_64B
  static constexpr uint32_t H_BYTES = KR * BKF * 128;                    // K x 4 KB, SWIZZLE_128B
  static constexpr uint32_t Z_BYTES = ((NE * 128 + 1023) / 1024) * 1024; // keeps the next stage 1024-aligned
  // Operand ring (filled by TMA, deep: HBM/L2 latency) and aggregated-tile ring (filled by the FiLM warps, shallow: on-chip latency)
  static constexpr uint32_t STAGE = 2 * A_BYTES + H_BYTES + Z_BYTES;
  static constexpr int NG = 2;
  static constexpr uint32_t OUT_BYTES = 16384;                           // one bf16 output slab
  static constexpr int NST_FIT = (227 * 1024 - 1024 - 512 - NG * (int)H_BYTES - (int)OUT_BYTES) / (int)STAGE;
  static constexpr int NST = NST_FIT > 6 ? 6 : NST_FIT;
  static_assert(NST >= 2, "shared memory budget");
  static constexpr uint32_t TX = 2 * A_BYTES + H_BYTES + NE * 128;       // bytes the producer's TMA loads deliver per stage
  static constexpr uint32_t G_OFF = NST * STAGE;
  static constexpr uint32_t OUT_OFF = G_OFF + NG * H_BYTES;
  static constexpr uint32_t BAR_OFF = OUT_OFF + OUT_BYTES;
  static constexpr size_t SMEM = 1024 + (size_t)BAR_OFF + 512;
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
  const uint32_t g_ring = base + C::G_OFF, out_stage = base + C::OUT_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + NST);
  const uint32_t bar_gfull = smem_u32(bars + 2 * NST), bar_gempty = smem_u32(bars + 2 * NST + C::NG);
  const uint32_t bar_tfull = smem_u32(bars + 2 * NST + 2 * C::NG), bar_tempty = smem_u32(bars + 2 * NST + 2 * C::NG + 1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 2 * C::NG + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmH) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmZ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < C::NG; ++i) { mbar_init(bar_gfull + 8 * i, 8); mbar_init(bar_gempty + 8 * i, 1); }
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
          tma_2d(s0 + 2 * C::A_BYTES + C::H_BYTES, &tmZ, fb, 2 * kb * BKF, (int)(g * C::NE));   // gamma/beta of its edges
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // ONE thread runs the whole issue loop (tcgen05.mma / commit are single-thread instructions; eight small MMAs per 32-channel
    // stage make the issuer's own instruction latency the critical path, so descriptors are precomputed and only their 14-bit
    // address field is bumped).  Skewed by one k-block: the h half of block i is issued as soon as its operands land, the
    // aggregated half of block i-1 right after it - the tensor pipe always has queued work while the FiLM warps process block i.
    if (lane == 0) {
      const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(C::N1 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)((C::N2 > 0 ? C::N2 : 64) >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint64_t dA = make_desc(base, 0, 512, 4);                    // K-major SWIZZLE_64B template (address field = base)
      const uint64_t dB = make_desc(base, BKF * 128, 1024);              // MN-major SWIZZLE_128B template
      auto at = [](uint64_t d, uint32_t byte_off) { return d + (uint64_t)(byte_off >> 4); };   // stays inside the 14-bit field (< 256 KB)
      int stage = 0; uint32_t phase = 0;         // operand ring position of the h half being issued
      int pstage = 0;                            // operand stage of the pending aggregated half
      int gb = 0; uint32_t gphase = 0;           // aggregated-tile ring position of the pending aggregated half
      uint32_t tphase = 0;
      auto issue_agg = [&](bool last_of_tile) {
        mbar_wait(bar_gfull + 8 * gb, gphase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_off = pstage * C::STAGE + C::A_BYTES, g_off = C::G_OFF + gb * C::H_BYTES;
#pragma unroll
        for (int k = 0; k < BKF / 16; ++k) {             // aggregated half:  D += Wc[:, F + kb] . g
          umma(tmem_base, at(dA, a_off + k * 32), at(dB, g_off + k * 2048), idesc1, 1u);
          if (C::N2 > 0) umma(tmem_base + C::N1, at(dA, a_off + k * 32), at(dB, g_off + k * 2048 + (C::N1 / 64) * (BKF * 128)), idesc2, 1u);
        }
        umma_commit(bar_empty + 8 * pstage);             // operand stage free (its h tile was also read by the FiLM warps: done before gfull)
        umma_commit(bar_gempty + 8 * gb);                // aggregated tile buffer free
        if (last_of_tile) umma_commit(bar_tfull);
        if (++gb == C::NG) { gb = 0; gphase ^= 1; }
      };
      for (int64_t w = blockIdx.x; w < work_total; w += gridDim.x) {
        mbar_wait(bar_tempty, tphase ^ 1);               // the single accumulator has been drained
        tphase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_off = stage * C::STAGE, h_off = a_off + 2 * C::A_BYTES;
          const uint32_t acc0 = kb > 0 ? 1u : 0u;
#pragma unroll
          for (int k = 0; k < BKF / 16; ++k) {           // h half:  D += Wc[:, kb] . h
            umma(tmem_base, at(dA, a_off + k * 32), at(dB, h_off + k * 2048), idesc1, k > 0 ? 1u : acc0);
            if (C::N2 > 0) umma(tmem_base + C::N1, at(dA, a_off + k * 32), at(dB, h_off + k * 2048 + (C::N1 / 64) * (BKF * 128)), idesc2, k > 0 ? 1u : acc0);
          }
          if (kb > 0) issue_agg(false);                  // aggregated half of the previous k-block
          pstage = stage;
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
        issue_agg(true);                                 // aggregated half of the tile's last k-block
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
    int gb = 0; uint32_t gphase = 0;
    for (int64_t w = blockIdx.x; w < work_total; w += gridDim.x) {
      const int tm = (int)(w % p.tiles_m);
      const int64_t g = w / p.tiles_m;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        const uint32_t s0 = base + stage * C::STAGE;
        const uint32_t hs = s0 + 2 * C::A_BYTES, zs = hs + C::H_BYTES;
        const uint32_t gs = g_ring + gb * C::H_BYTES;
        mbar_wait(bar_full + 8 * stage, phase);
        float2 xv[KR][4];
        const unsigned char* hsp = smem + (hs - base);
#pragma unroll
        for (int u = 0; u < KR; ++u) {      // plain shared loads (not volatile asm): the compiler is free to hoist / interleave them
          const uint4 t = *reinterpret_cast<const uint4*>(hsp + u * (BKF * 128) + voff);
          const uint32_t ws[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) xv[u][i] = make_float2(__uint_as_float(ws[i] << 16), __uint_as_float(ws[i] & 0xffff0000u));
        }
        // the KR(KR-1) gamma/beta pairs of this channel, fetched up front (plain loads: the compiler may schedule them freely)
        const uint32_t* zrow = reinterpret_cast<const uint32_t*>(smem + (zs - base)) + c;
        uint32_t gw[C::NE];
#pragma unroll
        for (int e = 0; e < C::NE; ++e) gw[e] = zrow[e * 32];
        // the aggregated-tile buffer: its previous tensor-core reader (and, when g is stored, its TMA store) must be done
        mbar_wait(bar_gempty + 8 * gb, gphase ^ 1);
        if (p.store_g) {
          if (issuer) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(C::NG - 1) : "memory");
          named_bar_sync(3, 256);
        }
        unsigned char* gsp = smem + (gs - base);
        uint4 outv[KR];
        if (!(p.debug & 1)) {
#pragma unroll
          for (int a = 0; a < KR; ++a) {
            float2 acc[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = make_float2(0.f, 0.f);
            float bsum = 0.f;
#pragma unroll
            for (int u = 0; u < KR; ++u) {
              if (u == a) continue;
              const uint32_t g1 = gw[u * (KR - 1) + (a < u ? a : a - 1)];
              const float gam = __uint_as_float(g1 << 16);
              bsum += __uint_as_float(g1 & 0xffff0000u);
              const float2 g2 = make_float2(gam, gam);
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[i] = ffma2(g2, xv[u][i], acc[i]);      // FFMA2: two fp32 FMAs per instruction
            }
            const float2 b2 = make_float2(bsum, bsum), i2 = make_float2(inv, inv);
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 t = fmul2(fadd2(acc[i], b2), i2);
              o[i] = pack_bf16(t.x, t.y);
            }
            outv[a] = make_uint4(o[0], o[1], o[2], o[3]);
          }
#pragma unroll
          for (int a = 0; a < KR; ++a) *reinterpret_cast<uint4*>(gsp + a * (BKF * 128) + voff) = outv[a];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to tcgen05 / TMA
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_gfull + 8 * gb);
        if (p.store_g && tm == (int)(g % p.tiles_m)) {    // every M-tile computes the same g: one of them (rotating with g, so
                                                          // that the duty is spread over all CTAs) writes it out
          named_bar_sync(3, 256);
          if (issuer) {
            tma_store_3d(&tmG, gs, 0, kb * BKF, (int)(g * KR));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        } else if (p.store_g && issuer) {
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");       // keep the group count in step with the wait above
        }
        if (++stage == NST) { stage = 0; phase ^= 1; }
        if (++gb == C::NG) { gb = 0; gphase ^= 1; }
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
        const uint32_t stg = out_stage;
        if (ew == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the slab's previous store has read it
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
          if (!(p.debug & 2)) tma_store_2d(&tmY, stg, 0, (int)((g * KR + j) * p.Fo + tm * BM));
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
  static DeviceOnce once;
  if (once.need()) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_film_fused<KR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_conv_film_fused)");
    once.mark();
  }
  FusedParams p = {};
  p.tiles_m = (int)(Fo / BM); p.graphs = graphs; p.F = (int)F; p.Fo = (int)Fo; p.kblocks = (int)(F / BKF);
  p.bias = bias; p.store_g = gout != nullptr;
  { const char* e = getenv("RGNN_FUSED_DEBUG"); p.debug = e ? atoi(e) : 0; }
  const int64_t work = graphs * p.tiles_m;
  const int grid = (int)(work < num_sms() ? work : num_sms());
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
