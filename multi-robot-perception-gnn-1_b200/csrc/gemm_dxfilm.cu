// A10 fused (BASELINE.json north_star (c)/(d)): the FiLM-mean BACKWARD in the EPILOGUE of the compress conv's dX GEMM, so the
// gradient of the concat, dcat = [dh_direct ; dg_h] (2 x N F P values), never exists in HBM.
//
//   dcat_v = Wc^T . dY_v                                        (autograd of Conv2d(2F,F,1) on cat(h, g_h), dgl/model/models.py:182-184)
//   dh[u]  = dcat[u, :F] + (1/(K-1)) sum_{a != u} gamma_{u->a} * dcat[a, F:]          (autograd of edge_udf / node_udf, models.py:207-211)
//   dgamma_{u->a} = (1/(K-1)) sum_p h[u,:,p] * dcat[a, F:, p] ,  dbeta_{u->a} = (1/(K-1)) sum_p dcat[a, F:, p] ;  dz = d(gamma,beta) * sigmoid'
//
// One CTA owns (graph g, 64 channels c0..c0+63): accumulator rows = the 64 "direct" rows Wc[:, c]^T and the 64 "aggregated" rows
// Wc[:, F + c]^T of the SAME channels, columns = ALL robots of the graph - exactly the values the FiLM backward of those channels
// needs.  The weight operand is row-permuted on the host: inside every group of 16 rows the first 8 are direct and the last 8 the
// aggregated rows of the same 8 channels, so that the 16-lane x 256-bit TMEM load hands one thread BOTH rows of a channel.
// A tile is processed as two half-tiles of 32 pixels: accumulator = 128 rows x (K robots x 32 px) = 160 columns at K = 5, three of
// them fit in TMEM, and the FiLM epilogue of half-tile i runs under the loads and MMAs of half-tiles i+1, i+2.
//
//   warp 0  TMA producer : ring of {A 128 x 64 (permuted Wc^T, K-major SWIZZLE_128B), B K x [64 o x 32 px] (dY, NCHW = MN-major,
//                          64-byte rows, SWIZZLE_64B)}; per half-tile the graph's h rows (double-buffered), per tile its gamma/beta rows
//   warp 1  MMA issuer   : tcgen05.mma cta_group::1, M = 128, N = 32 K, fp32 accumulate in TMEM
//   warp 2  TMEM alloc
//   warps 4-11 epilogue  : two warps per TMEM lane quadrant split its channels; tcgen05.ld.16x256b, FFMA2 math on all lanes, 4-byte
//                          dh stores, per-edge sums completed by a reduce-scatter over a channel's four threads, dz as (gamma, beta)
//                          pairs with sigmoid', dz column sums (db2 of the edge encoder) in shared memory -> one atomic per cell
// HBM traffic per layer: dY + h read, dh written (3 tensor passes) instead of dX (1 + 2) and the FiLM backward (3 + 1).
// History, knock-out timings and the variants that were measured and removed (single 320-column accumulator with cluster multicast,
// CTA pairs, weight tile in tensor memory, L2 prefetch): profiles/r02i_fused_bwd_ab.txt, DESIGN.md 4.1.
#include "tc_common.cuh"
#include "film_common.cuh"

namespace rgnn {
namespace dxf {

using namespace rgnn::tc;

constexpr int BM = 128, BK = 64, CT = 64;     // accumulator rows, k-block, channels per tile
constexpr int NTHR = 384;                     // warps 0-3: TMA / MMA / TMEM alloc / spare; warps 4-11: epilogue (2 per TMEM lane quadrant)

struct Params {
  int tiles;            // F / 64 channel tiles per graph
  int64_t graphs;
  int F, Fo, kblocks;   // kblocks = Fo / 64
  __nv_bfloat16* dh;    // [G K, F, 64]
  __nv_bfloat16* dz;    // [G K (K-1), 2 F]
  float* db2;           // optional fp32 [2 F]: column sums of dz (the bias gradient of the edge encoder's last Linear) are accumulated here (zeroed by the launcher)
  int fuse_sig;
  int debug;            // RGNN_DXF_DEBUG (experiments): 1 = epilogue only drains TMEM, 2 = no per-edge dots, 4 = no dh part, 8 = MMA issuer skips the MMAs
};

// mbarrier wait for the single-lane producer / issuer loops: back off between polls.  These warps share their SM sub-partitions
// with the FiLM epilogue warps, which are issue-bound (ncu: 28 % of all executed instructions were try_wait / branch pairs of
// spinning warps); a sleeping poller leaves the issue slots to the arithmetic.
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    __nanosleep(64);
  }
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// 16 lanes x 256 bits, twice (16 columns): registers {row t/4: col 2(t%4), +1; row t/4 + 8: same cols} per 8-column block - the
// mma.sync accumulator fragment layout
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// Shared-memory / tensor-memory budget of the pixel-half design (a single K x 64-column accumulator, the first version, left no room
// for a second one in TMEM: its epilogue could not overlap the next tile's MMAs, 10.5 us per tile).
template <int KR> struct CfgR {
  static constexpr int NE = KR * (KR - 1);
  static constexpr int NCOL = 32 * KR;                        // accumulator columns of a half-tile (UMMA N: multiple of 16)
  static constexpr int NBUF = 512 / NCOL >= 4 ? 4 : 512 / NCOL;
  static constexpr uint32_t A_BYTES = BM * BK * 2;            // 16 KB per k-block
  static constexpr uint32_t B_BYTES = KR * BK * 64;           // K x 4 KB (SWIZZLE_64B rows)
  static constexpr uint32_t STAGE = A_BYTES + B_BYTES;
  static constexpr uint32_t H_BYTES = KR * CT * 64;           // K x 4 KB
  static constexpr uint32_t Z_BYTES = ((NE * CT * 4 + 1023) / 1024) * 1024;
  static constexpr int NST_FIT = (227 * 1024 - 1024 - 512 - 2 * (int)H_BYTES - 2 * (int)Z_BYTES) / (int)STAGE;
  static constexpr int NST = NST_FIT > 6 ? 6 : NST_FIT;
  static_assert(NST >= 2 && NBUF >= 2, "shared memory / tensor memory budget");
  static constexpr uint32_t H_OFF = NST * STAGE;              // two h half-tile buffers
  static constexpr uint32_t Z_OFF = H_OFF + 2 * H_BYTES;
  static constexpr uint32_t BAR_OFF = Z_OFF + 2 * Z_BYTES;   // two gamma/beta buffers (tile i+1's rows arrive under tile i's epilogue)
  static constexpr size_t SMEM = 1024 + (size_t)BAR_OFF + 512;
};

template <int KR>
__global__ void __launch_bounds__(NTHR, 1)
k_dx_film_bwd_res(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmH,
                  const __grid_constant__ CUtensorMap tmZ, const Params p) {
  using C = CfgR<KR>;
  constexpr int NST = C::NST, NE = C::NE, NBUF = C::NBUF;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t base = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + NST);
  const uint32_t bar_tfull = smem_u32(bars + 2 * NST), bar_tempty = smem_u32(bars + 2 * NST + NBUF);
  const uint32_t bar_xfull = smem_u32(bars + 2 * NST + 2 * NBUF), bar_xempty = smem_u32(bars + 2 * NST + 2 * NBUF + 2);
  const uint32_t bar_zfull = smem_u32(bars + 2 * NST + 2 * NBUF + 4), bar_zempty = smem_u32(bars + 2 * NST + 2 * NBUF + 6);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 2 * NBUF + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDY) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmH) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmZ) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < NBUF; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 8); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_xfull + 8 * i, 1); mbar_init(bar_xempty + 8 * i, 8); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_zfull + 8 * i, 1); mbar_init(bar_zempty + 8 * i, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // per-CTA column sums of dz (one cell per (channel, gamma|beta); a cell is only ever touched by the thread that owns that channel)
  float* dbacc = reinterpret_cast<float*>(smem + C::BAR_OFF + 512);
  if (p.db2) for (int i = threadIdx.x; i < 2 * p.F; i += NTHR) dbacc[i] = 0.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // work item = (graph, channel tile), two half-tiles each; consecutive CTAs take the channel tiles of one graph (its dY is shared in L2)
  const int64_t work_total = p.graphs * p.tiles;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0, zphase = 0;
      int xb = 0, zb = 0; uint32_t xphase = 0;
      for (int64_t w = blockIdx.x; w < work_total; w += gridDim.x) {
        const int t = (int)(w % p.tiles);
        const int64_t g = w / p.tiles;
        for (int ph = 0; ph < 2; ++ph) {
          // the half-tile's h rows (and, with the first half, the tile's gamma/beta rows): double-buffered, asked for up front
          mbar_wait_backoff(bar_xempty + 8 * xb, xphase ^ 1);
          if (ph == 0) {
            mbar_wait_backoff(bar_zempty + 8 * zb, zphase ^ 1);
            mbar_expect_tx(bar_zfull + 8 * zb, NE * CT * 4);
            tma_2d(base + C::Z_OFF + zb * C::Z_BYTES, &tmZ, bar_zfull + 8 * zb, 2 * t * CT, (int)(g * NE));
            if (++zb == 2) { zb = 0; zphase ^= 1; }
          }
          mbar_expect_tx(bar_xfull + 8 * xb, C::H_BYTES);
          tma_3d(base + C::H_OFF + xb * C::H_BYTES, &tmH, bar_xfull + 8 * xb, ph * 32, t * CT, (int)(g * KR));
          if (++xb == 2) { xb = 0; xphase ^= 1; }
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait_backoff(bar_empty + 8 * stage, phase ^ 1);
            const uint32_t fb = bar_full + 8 * stage;
            const uint32_t s0 = base + stage * C::STAGE;
            const bool skip_a = (p.debug & 16) != 0;                                                      // knock-out: what an A operand kept on chip would save
            mbar_expect_tx(fb, skip_a ? C::B_BYTES : C::STAGE);
            if (!skip_a) tma_2d(s0, &tmW, fb, kb * BK, t * BM);                                          // permuted Wc^T rows of the channel tile
            tma_3d(s0 + C::A_BYTES, &tmDY, fb, ph * 32, kb * BK, (int)(g * KR));                         // dY[g's robots, 64 o, 32 px]
            if (++stage == NST) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      // D fp32, A/B bf16, A K-major (SWIZZLE_128B), B MN-major (SWIZZLE_64B: 32 pixels = 64-byte rows), N = 32 K, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(C::NCOL >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int tb = 0; uint32_t tphase = 0;
      for (int64_t w = blockIdx.x; w < work_total; w += gridDim.x) {
        for (int ph = 0; ph < 2; ++ph) {
          mbar_wait_backoff(bar_tempty + 8 * tb, tphase ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tacc = tmem_base + tb * C::NCOL;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait_backoff(bar_full + 8 * stage, phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = base + stage * C::STAGE, sb = sa + C::A_BYTES;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              if (p.debug & 8) continue;
              // B: atoms of 32 px (64 B) x 8 o-rows; robots (MN atoms) are BK * 64 bytes apart, 8-row groups 512 bytes, 16 o-rows 1024 bytes
              umma(tacc, make_desc(sa + k * 32, 0, 1024), make_desc(sb + k * 1024, BK * 64, 512, 4), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(bar_empty + 8 * stage);
            if (kb == p.kblocks - 1) umma_commit(bar_tfull + 8 * tb);
            if (++stage == NST) { stage = 0; phase ^= 1; }
          }
          if (++tb == NBUF) { tb = 0; tphase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue: FiLM backward ================================
    // The accumulator is read with the 16-lane x 256-bit TMEM load shape: thread t of a warp receives, for accumulator rows t/4 and
    // t/4 + 8 of a 16-lane group, columns 2 (t % 4) + {0, 1} of every 8-column block.  With the host's row permutation (8 direct rows
    // then the 8 aggregated rows of the same channels) a thread therefore holds BOTH rows of its channel for two pixels: no lane idles
    // in either half of the math, and the direct and aggregated values meet without a shuffle.  The two warps of a TMEM lane quadrant
    // split its channels (16 lanes = 8 channels each) and each covers all 32 pixels; the per-edge sums are completed by a butterfly
    // over the four threads of a channel once per tile.
    const int q = warp & 3;                                     // TMEM lane quadrant
    const int lg = (warp - 4) >> 2;                             // which 16 of its lanes: 8 of the quadrant's 16 channels
    const int pq = lane & 3;                                    // pixel pair inside an 8-pixel block
    const int ci = 8 * (2 * q + lg) + (lane >> 2);              // channel inside the tile (see the row permutation, host side)
    const int hsw = (ci >> 1) & 3;                              // SWIZZLE_64B: 16-byte chunk index ^ ((row >> 1) & 3)
    const float wmean = 1.0f / (KR - 1);
    int tb = 0; uint32_t tphase = 0;
    int xb = 0, zb = 0; uint32_t xphase = 0, zphase = 0;
    for (int64_t w = blockIdx.x; w < work_total; w += gridDim.x) {
      const int t = (int)(w % p.tiles);
      const int64_t g = w / p.tiles;
      const int c = t * CT + ci;
      float gw[NE];
      float2 dg[NE];
      float2 db[KR];
#pragma unroll
      for (int e = 0; e < NE; ++e) dg[e] = make_float2(0.f, 0.f);
#pragma unroll
      for (int a = 0; a < KR; ++a) db[a] = make_float2(0.f, 0.f);
      __nv_bfloat16* dh_row = p.dh + ((g * KR) * (int64_t)p.F + c) * 64 + 2 * pq;
      const uint32_t zs = base + C::Z_OFF + zb * C::Z_BYTES + ci * 4;
      mbar_wait_backoff(bar_zfull + 8 * zb, zphase);
#pragma unroll
      for (int e = 0; e < NE; ++e) gw[e] = __uint_as_float(lds_u32(zs + e * (CT * 4)) << 16) * wmean;
#pragma unroll 1
      for (int ph = 0; ph < 2; ++ph) {
        mbar_wait_backoff(bar_xfull + 8 * xb, xphase);
        mbar_wait_backoff(bar_tfull + 8 * tb, tphase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + tb * C::NCOL + ((uint32_t)(q * 32 + lg * 16) << 16);
        const uint32_t hs = base + C::H_OFF + xb * C::H_BYTES + ci * 64 + 4 * pq;      // (shared-window address: plain pointers compile to generic loads)
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {                           // 16 pixels of every robot per pass: 4 per thread
          // r[u]: {direct px, px+1, aggregated px, px+1} for the first 8-pixel block, then the same for the second
          uint32_t r[KR][8];
#pragma unroll
          for (int u = 0; u < KR; ++u) tmem_ld_16x256b_x2(tacc + u * 32 + j * 16, r[u]);
          tmem_ld_wait();
          if (j == 1) {                                         // this warp has read its part of the accumulator
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * tb);
          }
          if (p.debug & 1) continue;
          if (!(p.debug & 4)) {
#pragma unroll
            for (int u = 0; u < KR; ++u) {
              float2 s0 = make_float2(__uint_as_float(r[u][0]), __uint_as_float(r[u][1]));
              float2 s1 = make_float2(__uint_as_float(r[u][4]), __uint_as_float(r[u][5]));
#pragma unroll
              for (int a = 0; a < KR; ++a) {
                if (a == u) continue;
                const float gm = gw[edge_id(KR, u, a)];
                const float2 g2 = make_float2(gm, gm);
                s0 = ffma2(g2, make_float2(__uint_as_float(r[a][2]), __uint_as_float(r[a][3])), s0);
                s1 = ffma2(g2, make_float2(__uint_as_float(r[a][6]), __uint_as_float(r[a][7])), s1);
              }
              __nv_bfloat16* o = dh_row + (int64_t)u * p.F * 64 + ph * 32 + j * 16;
              *reinterpret_cast<uint32_t*>(o) = pack_bf16(s0.x, s0.y);
              *reinterpret_cast<uint32_t*>(o + 8) = pack_bf16(s1.x, s1.y);
            }
          }
          if (!(p.debug & 2)) {
#pragma unroll
            for (int a = 0; a < KR; ++a) {
              db[a] = fadd2(db[a], make_float2(__uint_as_float(r[a][2]), __uint_as_float(r[a][3])));
              db[a] = fadd2(db[a], make_float2(__uint_as_float(r[a][6]), __uint_as_float(r[a][7])));
            }
#pragma unroll
            for (int u = 0; u < KR; ++u) {
              // h half-tile: [robot][channel][64-byte row]; this thread's two pixel pairs of the pass
              const uint32_t h0 = lds_u32(hs + u * (CT * 64) + (((2 * j) ^ hsw) << 4));
              const uint32_t h1 = lds_u32(hs + u * (CT * 64) + (((2 * j + 1) ^ hsw) << 4));
              const float2 hv0 = make_float2(__uint_as_float(h0 << 16), __uint_as_float(h0 & 0xffff0000u));
              const float2 hv1 = make_float2(__uint_as_float(h1 << 16), __uint_as_float(h1 & 0xffff0000u));
#pragma unroll
              for (int a = 0; a < KR; ++a) {
                if (a == u) continue;
                float2 acc = dg[edge_id(KR, u, a)];
                acc = ffma2(hv0, make_float2(__uint_as_float(r[a][2]), __uint_as_float(r[a][3])), acc);
                acc = ffma2(hv1, make_float2(__uint_as_float(r[a][6]), __uint_as_float(r[a][7])), acc);
                dg[edge_id(KR, u, a)] = acc;
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_xempty + 8 * xb);        // this h half-tile has been read
        if (++xb == 2) { xb = 0; xphase ^= 1; }
        if (++tb == NBUF) { tb = 0; tphase ^= 1; }
      }
      // ---- both halves done: the four threads of a channel hold the sums over their 16 pixels each.  Reduce-scatter over the quad:
      // every thread ends up with the complete sums of a quarter of the channel's edges and writes those (a butterfly would leave all
      // four threads with all sums and three of four lanes idle through the write-out: 940 of 2100 warp instructions per tile).
      constexpr int NEP = (NE + 3) / 4 * 4;
      float v[NEP];
#pragma unroll
      for (int e = 0; e < NEP; ++e) v[e] = e < NE ? dg[e].x + dg[e].y : 0.f;
      reduce_scatter<NEP, 4>(v, pq);                             // v[i], i < NEP / 4: edge scattered_id<NEP, 4>(i, pq)
      float dbt[KR];                                             // the K dbeta pixel sums: every thread needs all of them
#pragma unroll
      for (int a = 0; a < KR; ++a) {
        float t = db[a].x + db[a].y;
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        dbt[a] = t * wmean;
      }
      {
        __nv_bfloat16* dzc = p.dz + (g * NE) * 2 * (int64_t)p.F + 2 * c;
        float sg = 0.f, sb = 0.f;
#pragma unroll
        for (int i = 0; i < NEP / 4; ++i) {
          const int e = scattered_id<NEP, 4>(i, pq);
          if (NE % 4 != 0 && e >= NE) continue;
          const int u = e / (KR - 1), rr = e - u * (KR - 1), a = rr < u ? rr : rr + 1;       // edge e = u -> a
          float dbt_a = dbt[0];
#pragma unroll
          for (int k = 1; k < KR; ++k) dbt_a = a == k ? dbt[k] : dbt_a;
          const uint32_t zw = lds_u32(zs + e * (CT * 4));
          const float gam = __uint_as_float(zw << 16), bet = __uint_as_float(zw & 0xffff0000u);
          float dgm = v[i] * wmean, dbe = dbt_a;
          if (p.fuse_sig) { dgm *= gam * (1.f - gam); dbe *= bet * (1.f - bet); }
          *reinterpret_cast<uint32_t*>(dzc + (int64_t)e * 2 * p.F) = pack_bf16(dgm, dbe);
          sg += dgm; sb += dbe;
        }
        if (p.db2) {
          sg += __shfl_xor_sync(0xffffffffu, sg, 1); sg += __shfl_xor_sync(0xffffffffu, sg, 2);
          sb += __shfl_xor_sync(0xffffffffu, sb, 1); sb += __shfl_xor_sync(0xffffffffu, sb, 2);
          if (pq == 0) { dbacc[2 * c] += sg; dbacc[2 * c + 1] += sb; }
        }
      }
      __syncwarp();                                             // the gamma/beta rows have been read
      if (lane == 0) mbar_arrive(bar_zempty + 8 * zb);
      if (++zb == 2) { zb = 0; zphase ^= 1; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (p.db2) for (int i = threadIdx.x; i < 2 * p.F; i += NTHR) atomicAdd(p.db2 + i, dbacc[i]);     // one atomic per cell per CTA
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

template <int KR>
static int launch_res(const void* Wp, const void* dy, const void* h, const void* z, void* dh, void* dz, float* db2, int64_t graphs, int64_t F,
                      int64_t Fo, int fuse_sig, cudaStream_t st) {
  using C = CfgR<KR>;
  const int64_t nodes = graphs * KR, edges = graphs * C::NE;
  CUtensorMap tw, tdy, th, tz;
  {
    const uint64_t d[2] = {(uint64_t)Fo, (uint64_t)(2 * F)}, s[1] = {(uint64_t)Fo * 2};
    const uint32_t b[2] = {BK, BM};
    int rc = make_map(&tw, Wp, 2, d, s, b); if (rc) return rc;
  }
  {
    const uint64_t d[3] = {64, (uint64_t)Fo, (uint64_t)nodes}, s[2] = {128, (uint64_t)Fo * 128};
    const uint32_t b[3] = {32, BK, KR};
    int rc = make_map(&tdy, dy, 3, d, s, b, CU_TENSOR_MAP_SWIZZLE_64B); if (rc) return rc;
  }
  {
    const uint64_t d[3] = {64, (uint64_t)F, (uint64_t)nodes}, s[2] = {128, (uint64_t)F * 128};
    const uint32_t b[3] = {32, CT, KR};
    int rc = make_map(&th, h, 3, d, s, b, CU_TENSOR_MAP_SWIZZLE_64B); if (rc) return rc;
  }
  {
    const uint64_t d[2] = {(uint64_t)(2 * F), (uint64_t)edges}, s[1] = {(uint64_t)(2 * F) * 2};
    const uint32_t b[2] = {2 * CT, (uint32_t)C::NE};
    int rc = make_map(&tz, z, 2, d, s, b, CU_TENSOR_MAP_SWIZZLE_NONE); if (rc) return rc;
  }
  const size_t smem = C::SMEM + (db2 ? (size_t)(2 * F * 4 + 127) / 128 * 128 : 0);
  if (smem > 227 * 1024) return RGNN_E_UNSUPPORTED;
  static DeviceOnce once;
  if (once.need()) {
    cudaError_t e = cudaFuncSetAttribute(k_dx_film_bwd_res<KR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_dx_film_bwd_res)");
    once.mark();
  }
  Params p = {};
  p.tiles = (int)(F / CT); p.graphs = graphs; p.F = (int)F; p.Fo = (int)Fo; p.kblocks = (int)(Fo / BK);
  p.dh = static_cast<__nv_bfloat16*>(dh); p.dz = static_cast<__nv_bfloat16*>(dz); p.db2 = db2; p.fuse_sig = fuse_sig;
  { const char* e = getenv("RGNN_DXF_DEBUG"); p.debug = e ? atoi(e) : 0; }
  const int64_t work = graphs * p.tiles;
  const int grid = (int)(work < num_sms() ? work : num_sms());
  if (grid < 1) return 0;
  k_dx_film_bwd_res<KR><<<grid, NTHR, smem, st>>>(tw, tdy, th, tz, p);
  RGNN_LAUNCH_CHECK("k_dx_film_bwd_res");
  return 0;
}

}  // namespace dxf
}  // namespace rgnn

using namespace rgnn;

// dh, dz of one whole layer from dY in ONE kernel (see the header of this file).  W_perm: bf16 [2F, Fo] = Wc^T with its rows permuted so
// that, for every channel tile t (64 channels) and j < 8, rows 128 t + 16 j + (0..7) are Wc[:, 64 t + 8 j + (0..7)]^T (direct half) and rows
// 128 t + 16 j + 8 + (0..7) are Wc[:, F + 64 t + 8 j + (0..7)]^T (aggregated half).  dy [G K, Fo, 64], h [G K, F, 64], z [G K (K-1), 2F]
// (gamma, beta interleaved), dh [G K, F, 64], dz [G K (K-1), 2F], all bf16; G complete `robots`-node graphs (2..6) in reference edge order.
// dbias2 (nullable): fp32 [2F], receives the column sums of dz (overwritten): db2 of the edge encoder without a pass over dz.
extern "C" int rgnn_tc_conv_dx_film_bwd(const void* W_perm, const void* dy, const void* h, const void* z, int fuse_sigmoid_grad,
                                        void* dh, void* dz, float* dbias2, int64_t graphs, int64_t robots, int64_t F, int64_t Fo, void* stream) {
  RGNN_REQUIRE(W_perm && dy && h && z && dh && dz && graphs >= 0 && F > 0 && Fo > 0, RGNN_E_ARG, "rgnn_tc_conv_dx_film_bwd: bad arguments");
  RGNN_REQUIRE(robots >= 2 && robots <= 6, RGNN_E_UNSUPPORTED, "rgnn_tc_conv_dx_film_bwd: 2..6 robots per graph");
  RGNN_REQUIRE(F % 64 == 0 && Fo % 64 == 0, RGNN_E_UNSUPPORTED, "rgnn_tc_conv_dx_film_bwd: F %% 64 and Fo %% 64 required");
  RGNN_REQUIRE(tc::al16(W_perm) && tc::al16(dy) && tc::al16(h) && tc::al16(z) && tc::al16(dh) && tc::al16(dz), RGNN_E_ALIGN,
               "rgnn_tc_conv_dx_film_bwd: 16-byte alignment");
  RGNN_REQUIRE(graphs * robots * (Fo > F ? Fo : F) < (1ll << 31) && graphs * robots * (robots - 1) < (1ll << 31), RGNN_E_RANGE,
               "rgnn_tc_conv_dx_film_bwd: too many rows for a tensor-map coordinate");
  cudaStream_t st = as_stream(stream);
  if (dbias2) {       // the CTAs add their partial column sums with atomics: start from zero
    cudaError_t e = cudaMemsetAsync(dbias2, 0, sizeof(float) * 2 * (size_t)F, st);
    if (e != cudaSuccess) return cuda_fail(e, "memset dbias2");
  }
  if (graphs == 0) return 0;
  switch (robots) {
    case 2: return dxf::launch_res<2>(W_perm, dy, h, z, dh, dz, dbias2, graphs, F, Fo, fuse_sigmoid_grad, st);
    case 3: return dxf::launch_res<3>(W_perm, dy, h, z, dh, dz, dbias2, graphs, F, Fo, fuse_sigmoid_grad, st);
    case 4: return dxf::launch_res<4>(W_perm, dy, h, z, dh, dz, dbias2, graphs, F, Fo, fuse_sigmoid_grad, st);
    case 5: return dxf::launch_res<5>(W_perm, dy, h, z, dh, dz, dbias2, graphs, F, Fo, fuse_sigmoid_grad, st);
    default: return dxf::launch_res<6>(W_perm, dy, h, z, dh, dz, dbias2, graphs, F, Fo, fuse_sigmoid_grad, st);
  }
}
