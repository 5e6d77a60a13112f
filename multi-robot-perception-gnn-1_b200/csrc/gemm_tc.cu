// A4/A8 (+A10): the weight transforms of the robot-graph layer as bf16 GEMMs on the 5th-gen tensor cores
// (tcgen05.mma, fp32 accumulators in TMEM, operands staged by TMA into 128B-swizzled shared memory).
//
// Reference ops replaced (file:line under /root/reference):
//   Conv2d(2F,F,1)(cat(h,g_h))  dgl/model/models.py:164,182-184  -> conv1x1  : Y_v = Wc . [h_v ; g_v] + b      (per robot v)
//   its autograd                 (training.py:209)                -> conv1x1  : dX_v = Wc^T . dY_v   (two outputs: dh | dagg)
//                                                                 -> wgrad    : dWc = sum_v dY_v . X_v^T  (split over robots)
//   nn.Linear(F,2F)+Sigmoid      dgl/model/models.py:149-150      -> gemm_tn  : Z = sigmoid(A1 . W2^T + b2)
//   its autograd                                                  -> gemm_tn  : dA1 = (dZ . W2) * [A1>0] ;  gemm_nn : dW2 = dZ^T . A1
//
// One persistent, warp-specialised kernel (1 CTA / SM, 384 threads), launched as clusters of two CTAs = one tcgen05 CTA pair:
//   warp 0   TMA producer   (6-stage ring of {A 128x64, own half of B (BN/2)x64} bf16 tiles; all bytes counted on the leader's mbarrier)
//   warp 1   MMA issuer     (leader CTA only, one elected lane; 4 x tcgen05.mma.cta_group::2 256xBNx16 per stage over both SMs;
//                            multicast tcgen05.commit frees the stage in both CTAs)
//   warp 2   TMEM allocator (2 x BN fp32 columns = double-buffered accumulator, epilogue overlaps the next tile)
//   (RGNN_TC_2CTA=0: cta_group::1 MMAs per CTA, 4-stage ring of {A, whole B}, the B halves TMA-multicast)
//   warps 4-11 epilogue     (two warpgroups, one per column half: tcgen05.ld 32x32b.x32 -> bias / sigmoid / relu-mask ->
//                            bf16 slabs staged in swizzled shared memory and written by TMA stores, or fp32 split partials)
// NCHW feature maps are consumed where they lie: a [64 ch x 64 px] slab of one robot is exactly one MN-major
// SWIZZLE_128B operand block (row = channel, 128 B = the 64 contiguous pixels), so the concat of models.py:182 is
// just "k-blocks 0..F/64 from h, the rest from g_h" (two tensor maps) and is never materialised.
#include "tc_common.cuh"

namespace rgnn {
namespace tc {

constexpr int BM = 128, BK = 64, NSTAGE = 4;
constexpr bool PREFETCH_L2 = false; // measured on B200: prefetching next-tile operands into L2 SLOWS the step (1.19 -> 1.43 ms
                                    // compress fwd): the extra L2 requests cost more than the HBM misses they hide
constexpr int PF_AHEAD = 12;        // WGRAD: k-blocks (robots) of L2 lead
constexpr int NTHREADS = 384;       // warps 0-3: TMA / MMA / TMEM alloc / spare; warps 4-11: two epilogue warpgroups (column halves)
constexpr int NOUT_STAGE = 2;       // bf16 output slabs (one per epilogue warpgroup; 4 x 48 KB operand stages leave room for no more)
constexpr int LOAD_TN = 0, LOAD_CONV = 1, LOAD_WGRAD = 2, LOAD_NN = 3;
constexpr int EPI_NCHW = 0, EPI_ROW = 1, EPI_PARTIAL = 2;

struct TcParams {
  int load_mode, epi_mode, epi_flags;
  int tiles_m, tiles_n, splits;
  int kblocks;          // k-blocks (of 64) in the full reduction
  int kb_per_split;     // k-blocks per split (== kblocks when splits == 1)
  int kb_b0;            // CONV: k-blocks taken from tensor map B0 (the rest from B1)
  int n_b0;             // WGRAD: output columns [0, n_b0) come from B0 (then B1)
  int rows_a;           // WGRAD: rows of A per robot (= channels of dY)
  int rows_b0, rows_b1; // WGRAD: rows per robot of B0 / B1
  int nodes;            // CONV: robots; output guard
  int64_t M, N;         // logical output extent (guards)
  void* out0; void* out1; int m_split;   // rows >= m_split go to out1 (row index - m_split)
  int64_t ld0, ld1;     // NCHW: elements per robot in out0/out1; ROW/PARTIAL: ld0 = row stride
  const float* bias; const void* aux;
  float alpha, beta;    // ROW epilogue: v = alpha * acc [+ beta * aux[m,n] with RGNN_EPI_ADD_AUX] (GCN2Conv identity mapping)
  int nchw_split;       // NCHW epilogue: 1 = the two warpgroups take two slabs each (one staging buffer per warpgroup)
  int n_fast;           // tile order: 1 = the N-tiles of one M-tile group are consecutive work items (A is the big operand: linear layers)
  float* rowsum;        // WGRAD: if non-null, fp32 [splits, M] partial row sums of A (the bias gradient of the 1x1 conv) are written here
  // ---- fp32 operands as three bf16 planes (x = x0 + x1 + x2, 8 significand bits each; profiles/r02_fp32_split_study.jsonl) ----
  // The reduction then runs `nprod` times over the k-blocks, product j taking plane pa[j] of A and plane pb[j] of B, the
  // 2^-16 terms first and x0.y0 last, all into ONE fp32 TMEM accumulator: A.B = sum_{i+j<=2} A_i.B_j to ~3e-7 (fp32 products),
  // provided a single accumulator sees no more than ~4096 elements of the original K (the accumulator truncates on alignment).
  int nprod;            // 0 / 1 = plain bf16 operands
  int pa[6], pb[6];
  int plane_a, plane_b0, plane_b1;   // distance between planes in the operand's OUTER tensor-map coordinate (rows / robots)
  int out_f32;          // NCHW / ROW epilogues write fp32 (two 128-byte slabs per 64 accumulator columns; exact sigmoid)
};

// CL = 2: the two CTAs of a cluster work on the two M-tiles that share one B tile (same tn); each loads HALF of the B tile
// and TMA-multicasts it into both CTAs' shared memory, cutting L2->SM operand traffic per CTA from 48 KB to 32 KB per
// k-block.  Every stage's empty barrier then collects the tcgen05.commit of BOTH CTAs (multicast commit) before either
// producer may overwrite the stage.  The MMAs themselves stay cta_group::1 (each CTA owns its accumulator).
// TWO (with CL = 2): the cluster is a tcgen05 CTA PAIR - one cta_group::2 MMA of M = 256 issued by the leader (rank 0) over both SMs.
// Each CTA then keeps only ITS half of the B tile: 32 KB instead of 48 KB land in shared memory per k-block and each MMA reads
// 8 KB instead of 12 KB, i.e. ~125 B/clk of shared-memory traffic instead of ~187 B/clk against a 128 B/clk port (the multicast
// variant caps at ~1.55 PFLOP/s with its stores compiled out), and the ring deepens from 4 to 6 stages.
template <int BN, int CL, bool AUX, bool TWO>      // AUX: ROW epilogue with a second operand (identity mapping / ReLU mask)
__global__ void __launch_bounds__(NTHREADS, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
          const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmO0,
          const __grid_constant__ CUtensorMap tmO1, const TcParams p) {
  static_assert(!TWO || CL == 2, "a CTA pair is a cluster of two");
  constexpr int NSTAGE = TWO ? 6 : tc::NSTAGE;
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = (TWO ? BN / 2 : BN) * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN;    // 256 or 512: power of two
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve: [tiles (1024-aligned)] [barriers]
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t tiles = smem_u32(smem);
  constexpr uint32_t OUT_STAGE_BYTES = 128 * 128;      // one robot's [128 ch x 64 px] bf16 output slab
  const uint32_t out_stage = tiles + NSTAGE * STAGE_BYTES;   // NOUT_STAGE slabs, 1024-aligned (TMA-store staging)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES + NOUT_STAGE * OUT_STAGE_BYTES);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + NSTAGE);
  const uint32_t bar_tfull = smem_u32(bars + 2 * NSTAGE), bar_tempty = smem_u32(bars + 2 * NSTAGE + 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 4);
  // CTA pair + row sums: the peer never sees its own operand bytes land (they are counted on the leader's barrier), so the row-sum
  // warps of BOTH CTAs read the A tile after the leader's MMAs on it have retired (multicast commit -> mdone) and only their
  // four arrivals free the stage.
  const uint32_t bar_mdone = smem_u32(bars + 2 * NSTAGE + 5);
  const bool rs_after_mma = TWO && p.rowsum != nullptr;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
  constexpr uint16_t MC_MASK = (uint16_t)((1u << CL) - 1);
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB1) : "memory");
    if (p.epi_mode != EPI_PARTIAL) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO0) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO1) : "memory");
    }
  }
  if (warp == 1 && lane == 0) {
    // a stage is free once every CTA's MMAs retired (+ this CTA's 4 epilogue warps when they also read the A tile: rowsum mode)
    // (CTA pair: only the leader commits, to both CTAs; the leader's accumulator-empty barrier collects both CTAs' epilogue warps)
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, rs_after_mma ? 4 : (TWO ? 1 : CL) + (p.rowsum ? 4 : 0));
      mbar_init(bar_mdone + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, TWO ? 16 : 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if constexpr (TWO) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CL > 1) cluster_sync_all();       // the peer's barriers exist before any multicast / remote arrive targets them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_mc = (p.tiles_m + CL - 1) / CL;     // M-tile groups: the CTAs of a cluster take consecutive M-tiles
  const int work_total = tiles_mc * p.tiles_n * p.splits;
  const int w_first = blockIdx.x / CL, w_step = gridDim.x / CL;
  const int nprod = p.nprod > 1 ? p.nprod : 1;
  const bool a_mn = p.load_mode == LOAD_NN;
  const bool b_mn = p.load_mode == LOAD_CONV || p.load_mode == LOAD_NN;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int w = w_first; w < work_total; w += w_step) {
        const int tm = (p.n_fast ? (w / p.tiles_n) % tiles_mc : w % tiles_mc) * CL + crank, tn = p.n_fast ? w % p.tiles_n : (w / tiles_mc) % p.tiles_n,
                  sp = w / (tiles_mc * p.tiles_n);
        const int kb0 = sp * p.kb_per_split, kb1 = min(p.kblocks, kb0 + p.kb_per_split);
        for (int pj = 0; pj < nprod; ++pj) {
        // split-plane products: the same k-blocks again with another (A plane, B plane) pair - only an outer coordinate moves
        const int oa = p.nprod > 1 ? p.pa[pj] * p.plane_a : 0, ob0 = p.nprod > 1 ? p.pb[pj] * p.plane_b0 : 0,
                  ob1 = p.nprod > 1 ? p.pb[pj] * p.plane_b1 : 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_sleep(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t fb = bar_full + 8 * stage;
          const uint32_t sa = tiles + stage * STAGE_BYTES, sb = sa + A_BYTES;
          if constexpr (TWO) {
            // CTA pair: own A rows + own half of B into this CTA's stage; every byte is counted on the LEADER's barrier
            if (crank == 0) mbar_expect_tx(fb, 2 * STAGE_BYTES);
            if (p.load_mode == LOAD_TN) {
              tma_2d_2sm(sa, &tmA, fb, kb * BK, tm * BM + oa);
              tma_2d_2sm(sb, &tmB0, fb, kb * BK, tn * BN + crank * (BN / 2) + ob0);
            } else if (p.load_mode == LOAD_CONV) {
              tma_2d_2sm(sa, &tmA, fb, kb * BK, tm * BM + oa);
              const CUtensorMap* mb = kb < p.kb_b0 ? &tmB0 : &tmB1;
              const int kc = (kb < p.kb_b0 ? kb : kb - p.kb_b0) * BK;
              tma_3d_2sm(sb, mb, fb, 0, kc, tn * (BN / 64) + crank * (BN / 128) + (kb < p.kb_b0 ? ob0 : ob1));
            } else if (p.load_mode == LOAD_WGRAD) {
              tma_2d_2sm(sa, &tmA, fb, 0, kb * p.rows_a + tm * BM + oa);
              const int n0 = tn * BN;
              const CUtensorMap* mb = n0 < p.n_b0 ? &tmB0 : &tmB1;
              const int rowb = n0 < p.n_b0 ? kb * p.rows_b0 + n0 + ob0 : kb * p.rows_b1 + (n0 - p.n_b0) + ob1;
              tma_2d_2sm(sb, mb, fb, 0, rowb + crank * (BN / 2));
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_2d_2sm(sa + j * (BK * 128), &tmA, fb, tm * BM + j * 64, kb * BK + oa);
#pragma unroll
              for (int j = 0; j < BN / 128; ++j) tma_2d_2sm(sb + j * (BK * 128), &tmB0, fb, tn * BN + (crank * (BN / 128) + j) * 64, kb * BK + ob0);
            }
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_expect_tx(fb, STAGE_BYTES);
          // A: this CTA's own 128 rows.  B: tensor maps carry HALF-tile boxes; a lone CTA fetches both halves, a cluster CTA
          // fetches half `crank` and multicasts it to every CTA of the cluster.
          constexpr uint32_t HALF_B = B_BYTES / 2;
          const int h0 = CL > 1 ? crank : 0, h1 = CL > 1 ? crank + 1 : 2;
          if (p.load_mode == LOAD_TN) {
            tma_2d(sa, &tmA, fb, kb * BK, tm * BM + oa);
            for (int hb = h0; hb < h1; ++hb) {
              if (CL > 1) tma_2d_mc(sb + hb * HALF_B, &tmB0, fb, kb * BK, tn * BN + hb * (BN / 2) + ob0, MC_MASK);
              else tma_2d(sb + hb * HALF_B, &tmB0, fb, kb * BK, tn * BN + hb * (BN / 2) + ob0);
            }
          } else if (p.load_mode == LOAD_CONV) {
            tma_2d(sa, &tmA, fb, kb * BK, tm * BM + oa);
            const CUtensorMap* mb = kb < p.kb_b0 ? &tmB0 : &tmB1;
            const int kc = (kb < p.kb_b0 ? kb : kb - p.kb_b0) * BK;
            const int ob = kb < p.kb_b0 ? ob0 : ob1;
            if (PREFETCH_L2 && crank == 0 && (w % tiles_mc) == 0 && w + w_step < work_total) {
              // the activation tile of this CTA's NEXT work item (one whole tile of lead time); the M-tile siblings reuse it from L2
              const int tn2 = ((w + w_step) / tiles_mc) % p.tiles_n;
              tma_prefetch_3d(mb, 0, kc, tn2 * (BN / 64));
              tma_prefetch_3d(mb, 0, kc, tn2 * (BN / 64) + BN / 128);
            }
            for (int hb = h0; hb < h1; ++hb) {
              if (CL > 1) tma_3d_mc(sb + hb * HALF_B, mb, fb, 0, kc, tn * (BN / 64) + hb * (BN / 128) + ob, MC_MASK);
              else tma_3d(sb + hb * HALF_B, mb, fb, 0, kc, tn * (BN / 64) + hb * (BN / 128) + ob);
            }
          } else if (p.load_mode == LOAD_WGRAD) {   // k-block == robot kb; both operands K-major (pixels contiguous)
            tma_2d(sa, &tmA, fb, 0, kb * p.rows_a + tm * BM + oa);
            const int n0 = tn * BN;
            const CUtensorMap* mb = n0 < p.n_b0 ? &tmB0 : &tmB1;
            const int rowb = n0 < p.n_b0 ? kb * p.rows_b0 + n0 + ob0 : kb * p.rows_b1 + (n0 - p.n_b0) + ob1;
            if (PREFETCH_L2 && kb + PF_AHEAD < kb1) {     // robots further down this split: dY rows by the tn == 0 CTAs, X rows by tm == 0
              if (tn == 0) tma_prefetch_2d(&tmA, 0, (kb + PF_AHEAD) * p.rows_a + tm * BM);
              if (tm == 0) {
                const int rb2 = n0 < p.n_b0 ? (kb + PF_AHEAD) * p.rows_b0 + n0 : (kb + PF_AHEAD) * p.rows_b1 + (n0 - p.n_b0);
                tma_prefetch_2d(mb, 0, rb2);
                tma_prefetch_2d(mb, 0, rb2 + BN / 2);
              }
            }
            for (int hb = h0; hb < h1; ++hb) {
              if (CL > 1) tma_2d_mc(sb + hb * HALF_B, mb, fb, 0, rowb + hb * (BN / 2), MC_MASK);
              else tma_2d(sb + hb * HALF_B, mb, fb, 0, rowb + hb * (BN / 2));
            }
          } else {                                   // LOAD_NN: both operands MN-major, 64-wide boxes
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_2d(sa + j * (BK * 128), &tmA, fb, tm * BM + j * 64, kb * BK + oa);
            for (int hb = h0; hb < h1; ++hb)
              for (int j = hb * (BN / 128); j < (hb + 1) * (BN / 128); ++j) {
                if (CL > 1) tma_2d_mc(sb + j * (BK * 128), &tmB0, fb, tn * BN + j * 64, kb * BK + ob0, MC_MASK);
                else tma_2d(sb + j * (BK * 128), &tmB0, fb, tn * BN + j * 64, kb * BK + ob0);
              }
          }
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        }
      }
    }
  } else if (warp == 1 && !(TWO && crank != 0)) {
    // ================================ MMA issuer (CTA pair: the leader only) ================================
    // instruction descriptor: D fp32, A/B bf16, majors, N>>3, M>>4 (cute::UMMA::InstrDescriptor)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((TWO ? 2 * BM : BM) >> 4) << 24);
    const uint32_t a_lbo = a_mn ? BK * 128 : 0, b_lbo = b_mn ? BK * 128 : 0;
    const uint32_t a_kstep = a_mn ? 2048 : 32, b_kstep = b_mn ? 2048 : 32;   // bytes per 16 k
    int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
    for (int w = w_first; w < work_total; w += w_step) {
      const int sp = w / (tiles_mc * p.tiles_n);
      const int kb0 = sp * p.kb_per_split, kb1 = min(p.kblocks, kb0 + p.kb_per_split);
      mbar_wait(bar_tempty + 8 * as, aphase ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tacc = tmem_base + as * BN;
      const int nit = nprod * (kb1 - kb0);                 // split-plane products: the k-block range is walked nprod times
      for (int it = 0; it < nit; ++it) {
        mbar_wait(bar_full + 8 * stage, phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t sa = tiles + stage * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = make_desc(sa + k * a_kstep, a_lbo, 1024);
            const uint64_t db = make_desc(sb + k * b_kstep, b_lbo, 1024);
            if constexpr (TWO) umma2(tacc, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
            else umma(tacc, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
          if constexpr (TWO) {
            umma_commit2_mc((rs_after_mma ? bar_mdone : bar_empty) + 8 * stage, MC_MASK);   // the stage is free (or ready for the row sums) in both CTAs
            if (it == nit - 1) umma_commit2_mc(bar_tfull + 8 * as, MC_MASK);     // both halves of the accumulator are complete
          } else {
          if (CL > 1) umma_commit_mc(bar_empty + 8 * stage, MC_MASK);   // both CTAs' producers wait for both consumers
          else umma_commit(bar_empty + 8 * stage);            // stage reusable once these MMAs retire
          if (it == nit - 1) umma_commit(bar_tfull + 8 * as); // accumulator complete
          }
        }
        __syncwarp();
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
      if (kb1 <= kb0 && elect_one()) {                                  // empty split: hand over (epilogue writes zeros)
        if constexpr (TWO) umma_commit2_mc(bar_tfull + 8 * as, MC_MASK); else umma_commit(bar_tfull + 8 * as);
      }
      __syncwarp();
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ================================
    const int ew = (warp - 4) & 3;              // == warp % 4: TMEM lanes [32*ew, 32*ew+32)
    const int eh = (warp - 4) >> 2;             // epilogue warpgroup: 0 = left half of the tile's columns, 1 = right half
    const int row = ew * 32 + lane;             // row inside the 128-row tile
    const int bar_id = 1 + eh;                  // named barrier of this warpgroup
    int as = 0; uint32_t aphase = 0;
    int rstage = 0; uint32_t rphase = 0;        // rowsum mode: this warp's own walk over the operand ring
    for (int w = w_first; w < work_total; w += w_step) {
      const int tm = (p.n_fast ? (w / p.tiles_n) % tiles_mc : w % tiles_mc) * CL + crank, tn = p.n_fast ? w % p.tiles_n : (w / tiles_mc) % p.tiles_n,
                sp = w / (tiles_mc * p.tiles_n);
      const int kb0 = sp * p.kb_per_split, kb1 = min(p.kblocks, kb0 + p.kb_per_split);
      const bool empty = kb1 <= kb0;
      if (p.rowsum && eh == 0) {
        // The bias gradient db[o] = sum_{v,p} dY[v,o,p] rides on the WGRAD mainloop: the A tile of every k-block IS dY[v, o-tile, :]
        // in shared memory, and these four warps are idle until the accumulator is complete.  Thread = one tile row (128 B);
        // the 16-byte chunks are visited in a lane-rotated order (the 128B swizzle only permutes chunks inside a row, which a
        // row sum does not care about) so the 8 lanes of a quarter-warp hit 8 different bank groups.
        float rs = 0.f;
        const bool mine = tn == 0;              // every N-tile of an M-tile loads the same A: only one of them sums it
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait((rs_after_mma ? bar_mdone : bar_full) + 8 * rstage, rphase);
          if (mine) {
            const uint32_t rowaddr = tiles + rstage * STAGE_BYTES + row * 128;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              uint32_t a0, a1, a2, a3;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(rowaddr + (((i + lane) & 7) << 4)));
              rs += (__uint_as_float(a0 << 16) + __uint_as_float(a0 & 0xffff0000u)) + (__uint_as_float(a1 << 16) + __uint_as_float(a1 & 0xffff0000u)) +
                    (__uint_as_float(a2 << 16) + __uint_as_float(a2 & 0xffff0000u)) + (__uint_as_float(a3 << 16) + __uint_as_float(a3 & 0xffff0000u));
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_empty + 8 * rstage);
          if (++rstage == NSTAGE) { rstage = 0; rphase ^= 1; }
        }
        const int64_t mr = (int64_t)tm * BM + row;
        if (mine && mr < p.M) p.rowsum[(int64_t)sp * p.M + mr] = rs;
      }
      // ROW epilogues that read a second operand (identity mapping / ReLU mask): each aux slab [128 rows x 128 B] is fetched
      // COALESCED (8 lanes per 128-byte row, 4 rows per warp load) into registers one slab ahead - the tile's first slab while
      // this warp would only wait for the accumulator, slab j+1 during the arithmetic of slab j - and parked in the warpgroup's
      // staging buffer at the swizzled positions the results will overwrite; a warp's 32 rows belong to its own 32 threads, so
      // __syncwarp is the only synchronisation.  (One 16-byte load per thread-row instead: 32 sectors per request, twice the
      // L2 traffic, and the TN GEMM ran at half speed.)
      constexpr bool use_aux = AUX;
      uint4 ua[8];
      auto aux_fetch = [&](int j) {
        const int64_t c0 = (int64_t)tn * BN + j * 64 + (lane & 7) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t mr = (int64_t)tm * BM + ew * 32 + i * 4 + (lane >> 3);
          ua[i] = mr < p.M ? __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.aux) + mr * p.ld0 + c0)) : make_uint4(0, 0, 0, 0);
        }
      };
      if (use_aux && !empty) aux_fetch(eh * (BN / 128));
      mbar_wait_sleep(bar_tfull + 8 * as, aphase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tacc = tmem_base + as * BN + ((uint32_t)(ew * 32) << 16);
      const int64_t m = (int64_t)tm * BM + row;
      const float bias_m = (p.bias && (p.epi_flags & RGNN_EPI_BIAS_M) && m < p.M) ? p.bias[m] : 0.f;
      if (p.epi_mode == EPI_NCHW || p.epi_mode == EPI_ROW) {
        // bf16 outputs leave through shared memory: each 128-row x 64-column (128 B) slab is staged with the 128B swizzle
        // (conflict-free 16-byte stores) and written by ONE TMA store - full 128 B lines, rows beyond M clipped by the map.
        //   NCHW: row = output channel, slab j = robot j of the tile: the slab is a contiguous 16 KB piece of [N, F, 8, 8].
        //   ROW : row = output row, slab j = columns [64 j, 64 j + 64) of a row-major matrix.
        const bool nchw = p.epi_mode == EPI_NCHW;
        const bool second = nchw && m >= p.m_split;
        const int row0 = tm * BM - (second ? p.m_split : 0);            // first channel of this tile inside its output tensor
        const int rows_per_node = second ? (int)(p.ld1 / 64) : (int)(p.ld0 / 64);
        // NCHW (conv) tiles are limited by the output write traffic, not by these warps: one warpgroup walks all slabs with both
        // staging buffers (store of slab j overlaps the staging of slab j+1).  ROW tiles (edge-MLP, short K, sigmoid) are limited
        // by epilogue issue: the two warpgroups take one column half each, one buffer per warpgroup.
        const bool split_wg = !nchw || p.nchw_split;
        bool f32_done = false;
        if constexpr (!AUX) {
          if (p.out_f32) {
            // fp32 outputs (split-plane products of the fp32 parity mode): the same staging slab (128 rows x 128 B, 128B swizzle)
            // now holds 32 fp32 columns, so a 64-column group (one robot's 64 pixels / 64 matrix columns) leaves as two slabs;
            // the output tensor maps describe the fp32 tensor as pairs of 16-bit elements (TMA moves bytes, it does no arithmetic).
            f32_done = true;
            constexpr int NSLAB = BN / 32;
            const int f_lo = split_wg ? eh * (NSLAB / 2) : 0, f_hi = split_wg ? (eh + 1) * (NSLAB / 2) : (eh == 0 ? NSLAB : 0);
#pragma unroll 1
            for (int j = f_lo; j < f_hi; ++j) {
              const int node = tn * (BN / 64) + (j >> 1);
              const uint32_t stg = out_stage + (split_wg ? eh : (j & 1)) * OUT_STAGE_BYTES;
              const int64_t n0 = (int64_t)tn * BN + j * 32;          // first fp32 column of the slab (ROW)
              uint32_t r0[32];
              tmem_ld32_issue(tacc + j * 32, r0);
              tmem_ld_wait();
              float v[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r0[i]) + bias_m;
              if (!nchw) {
                if (p.alpha != 1.0f) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) v[i] *= p.alpha;
                }
                if (p.epi_flags & RGNN_EPI_BIAS_N) {
#pragma unroll
                  for (int q = 0; q < 8; ++q) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + q * 4));
                    v[4 * q] += b0.x; v[4 * q + 1] += b0.y; v[4 * q + 2] += b0.z; v[4 * q + 3] += b0.w;
                  }
                }
                if (p.epi_flags & RGNN_EPI_SIGMOID) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) v[i] = __fdividef(1.0f, 1.0f + __expf(-v[i]));   // parity mode: ~2 ulp, not the bf16-grade tanh.approx
                }
              }
              if (ew == 0 && lane == 0) {
                if (!split_wg) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              }
              epi_bar_sync(bar_id);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const uint32_t addr = stg + row * 128 + ((q ^ (row & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(__float_as_uint(v[4 * q])), "r"(__float_as_uint(v[4 * q + 1])),
                             "r"(__float_as_uint(v[4 * q + 2])), "r"(__float_as_uint(v[4 * q + 3])) : "memory");
              }
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              epi_bar_sync(bar_id);
              if (ew == 0 && lane == 0) {
                if (nchw) { if (node < p.nodes && (int64_t)tm * BM < p.M) tma_store_2d(second ? &tmO1 : &tmO0, stg, (j & 1) * 64, node * rows_per_node + row0); }
                else if (n0 < p.N) tma_store_2d(&tmO0, stg, (int)(2 * n0), tm * BM);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
            }
          }
        }
        const int j_lo = f32_done ? 0 : (split_wg ? eh * (BN / 128) : 0), j_hi = f32_done ? 0 : (split_wg ? (eh + 1) * (BN / 128) : (eh == 0 ? BN / 64 : 0));
#pragma unroll 1
        for (int j = j_lo; j < j_hi; ++j) {
          const int node = tn * (BN / 64) + j;
          const uint32_t stg = out_stage + (split_wg ? eh : (j & 1)) * OUT_STAGE_BYTES;
          const int64_t n0 = (int64_t)tn * BN + j * 64;
          if (use_aux) {
            if (ew == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            epi_bar_sync(bar_id);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rl = ew * 32 + i * 4 + (lane >> 3);
              const uint32_t addr = stg + rl * 128 + (((lane & 7) ^ (rl & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(ua[i].x), "r"(ua[i].y), "r"(ua[i].z), "r"(ua[i].w) : "memory");
            }
            __syncwarp();
            if (j + 1 < j_hi) aux_fetch(j + 1);
          }
          uint32_t r0[32], r1[AUX ? 1 : 32];
          uint32_t wq[8][4];                      // the slab row of this thread, packed; staged after the buffer is known to be free
          if constexpr (!AUX) {
            tmem_ld32_issue(tacc + (2 * j) * 32, r0);
            tmem_ld32_issue(tacc + (2 * j + 1) * 32, r1);
            tmem_ld_wait();
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if constexpr (AUX) {
              if ((q & 3) == 0) {                 // 32 accumulator columns at a time (the registers are needed for the aux prefetch)
                tmem_ld32_issue(tacc + (2 * j + (q >> 2)) * 32, r0);
                tmem_ld_wait();
              }
            }
            const uint32_t* r = (AUX || q < 4) ? r0 : r1;
            uint4 u = make_uint4(0, 0, 0, 0);     // this thread's aux chunk q (from the staged slab)
            if (use_aux)
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(stg + row * 128 + ((q ^ (row & 7)) << 4)));
            const int b = (q & 3) * 8;
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[b + i]) + bias_m;
            if (!nchw) {
              if (p.alpha != 1.0f) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] *= p.alpha;
              }
              if ((p.epi_flags & RGNN_EPI_ADD_AUX) && m < p.M) {
                const uint32_t ws[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  v[2 * i] = fmaf(p.beta, __uint_as_float(ws[i] << 16), v[2 * i]);
                  v[2 * i + 1] = fmaf(p.beta, __uint_as_float(ws[i] & 0xffff0000u), v[2 * i + 1]);
                }
              }
              if (p.epi_flags & RGNN_EPI_BIAS_N) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + q * 8));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + q * 8 + 4));
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
              }
              if (p.epi_flags & RGNN_EPI_SIGMOID) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {     // sigmoid(x) = 0.5 tanh(x/2) + 0.5: one MUFU instead of ex2 + rcp (error 2^-11 << bf16 ulp)
                  float th;
                  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * v[i]));
                  v[i] = fmaf(th, 0.5f, 0.5f);
                }
              }
              if ((p.epi_flags & RGNN_EPI_RELU_MASK) && m < p.M) {
                const uint32_t ws[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {     // bf16 > 0  <=>  sign clear and magnitude non-zero
                  if (!((ws[i] & 0x7fffu) != 0 && !(ws[i] & 0x8000u))) v[2 * i] = 0.f;
                  if (!(((ws[i] >> 16) & 0x7fffu) != 0 && !((ws[i] >> 16) & 0x8000u))) v[2 * i + 1] = 0.f;
                }
              }
            }
            wq[q][0] = pack_bf16(v[0], v[1]); wq[q][1] = pack_bf16(v[2], v[3]); wq[q][2] = pack_bf16(v[4], v[5]); wq[q][3] = pack_bf16(v[6], v[7]);
          }
          // only now wait for the staging buffer: the previous store's shared-memory read overlapped the TMEM load and the arithmetic
          if (!use_aux) {                         // (with an aux operand the buffer was claimed before it was parked there)
            if (ew == 0 && lane == 0) {
              if (!split_wg) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            epi_bar_sync(bar_id);
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint32_t addr = stg + row * 128 + ((q ^ (row & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(wq[q][0]), "r"(wq[q][1]), "r"(wq[q][2]), "r"(wq[q][3]) : "memory");
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          epi_bar_sync(bar_id);
          if (ew == 0 && lane == 0) {
            if (nchw) { if (node < p.nodes && (int64_t)tm * BM < p.M) tma_store_2d(second ? &tmO1 : &tmO0, stg, 0, node * rows_per_node + row0); }
            else if (n0 < p.N) tma_store_2d(&tmO0, stg, (int)n0, tm * BM);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      } else {
#pragma unroll 1
      for (int c = eh * (BN / 64); c < (eh + 1) * (BN / 64); ++c) {
        uint32_t r[32];
        tmem_ld32_issue(tacc + c * 32, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = empty ? 0.f : __uint_as_float(r[i]);
        {
          const int64_t n0 = (int64_t)tn * BN + c * 32;
          if (m < p.M && n0 < p.N) {
            float* o = static_cast<float*>(p.out0) + ((int64_t)sp * p.M + m) * p.ld0 + n0;
#pragma unroll
            for (int q = 0; q < 8; ++q) reinterpret_cast<float4*>(o)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        }
      }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (TWO && crank != 0) mbar_arrive_cluster(bar_tempty + 8 * as, 0);  // the leader's MMA warp waits for both CTAs' epilogues
        else mbar_arrive(bar_tempty + 8 * as);
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (ew == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging must outlive its readers (both warpgroups)
  }
  // ================================ teardown ================================
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CL > 1) cluster_sync_all();       // nobody leaves while a peer may still multicast into it or signal its barriers
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if constexpr (TWO) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// out[i] = alpha * sum_s partial[s][i]   (fixed order: deterministic); optionally, in the same launch, a second (small) segment:
// out2[j] = sum_s part2[s][j], j < n2 (the conv bias gradient's row-sum partials ride on the weight gradient's reduction)
__global__ void __launch_bounds__(256) k_sum_partials(const float* __restrict__ part, int splits, int64_t n, float* __restrict__ out,
                                                      int64_t ld_out, int64_t cols, float alpha, const float* __restrict__ part2 = nullptr,
                                                      int64_t n2 = 0, float* __restrict__ out2 = nullptr) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n + n2; i += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    if (i < n) {
      for (int s = 0; s < splits; ++s) acc += part[(int64_t)s * n + i];
      out[(i / cols) * ld_out + (i % cols)] = alpha * acc;
    } else {
      const int64_t j = i - n;
      for (int s = 0; s < splits; ++s) acc += part2[(int64_t)s * n2 + j];
      out2[j] = acc;
    }
  }
}

// fp32 [R, C] -> bf16 [R, C] and (optionally) its transpose [C, R]: per-step cast of the fp32 master weights
// (alpha * w + ident * I: GCN2Conv's identity mapping (1-beta) I + beta W folded into the weight matrix itself)
__global__ void __launch_bounds__(256) k_cast_weights(const float* __restrict__ w, int R, int C, __nv_bfloat16* __restrict__ o,
                                                      __nv_bfloat16* __restrict__ ot, float alpha, float ident) {
  __shared__ float t[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    const float v = (r < R && c < C) ? fmaf(alpha, w[(int64_t)r * C + c], r == c ? ident : 0.f) : 0.f;
    t[j][tx] = v;
    if (o && r < R && c < C) o[(int64_t)r * C + c] = __float2bfloat16_rn(v);
  }
  __syncthreads();
  if (ot) {
    for (int j = ty; j < 32; j += 8) {
      const int c = c0 + j, r = r0 + tx;
      if (r < R && c < C) ot[(int64_t)c * R + r] = __float2bfloat16_rn(t[tx][j]);
    }
  }
}

// fp32 -> three bf16 planes with x = x0 + x1 + x2 (each term the round-to-nearest bf16 of what is left: 3 x 8 significand bits
// cover fp32's 24).  Plane p of element i lives at out[p * plane + i]; elements [n, plane) of every plane are zeroed (K tails of
// MN-major operands must contribute nothing).  4 elements per thread: one 16-byte load, three 8-byte stores.
__device__ __forceinline__ void split3(float x, __nv_bfloat16& a, __nv_bfloat16& b, __nv_bfloat16& c) {
  a = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(a);        // exact: a keeps the leading bits of x
  b = __float2bfloat16_rn(r1);
  c = __float2bfloat16_rn(r1 - __bfloat162float(b));
}
__global__ void __launch_bounds__(256) k_split3(const float* __restrict__ x, int64_t n4, int64_t plane4, __nv_bfloat16* __restrict__ out) {
  // n4 / plane4: counts in units of 4 elements
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane4; i += (int64_t)gridDim.x * blockDim.x) {
    uint2 w[3] = {make_uint2(0, 0), make_uint2(0, 0), make_uint2(0, 0)};
    if (i < n4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
      const float f[4] = {v.x, v.y, v.z, v.w};
      __nv_bfloat16 h[3][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) split3(f[k], h[0][k], h[1][k], h[2][k]);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        __nv_bfloat162 lo = __halves2bfloat162(h[q][0], h[q][1]), hi = __halves2bfloat162(h[q][2], h[q][3]);
        w[q] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      }
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) reinterpret_cast<uint2*>(out + q * plane4 * 4)[i] = w[q];
  }
}
// weights [R, C] fp32 -> planes [3, R, C] and/or planes of the transpose [3, C, R]
__global__ void __launch_bounds__(256) k_split3_weights(const float* __restrict__ w, int R, int C, __nv_bfloat16* __restrict__ o,
                                                        __nv_bfloat16* __restrict__ ot) {
  __shared__ float t[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t plane = (int64_t)R * C;
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    const float v = (r < R && c < C) ? w[(int64_t)r * C + c] : 0.f;
    t[j][tx] = v;
    if (o && r < R && c < C) {
      __nv_bfloat16 a, b, d;
      split3(v, a, b, d);
      const int64_t i = (int64_t)r * C + c;
      o[i] = a; o[plane + i] = b; o[2 * plane + i] = d;
    }
  }
  __syncthreads();
  if (ot) {
    for (int j = ty; j < 32; j += 8) {
      const int c = c0 + j, r = r0 + tx;
      if (r < R && c < C) {
        __nv_bfloat16 a, b, d;
        split3(t[tx][j], a, b, d);
        const int64_t i = (int64_t)c * R + r;
        ot[i] = a; ot[plane + i] = b; ot[2 * plane + i] = d;
      }
    }
  }
}

// Operand planes of a split-bf16 product (host side of TcParams::nprod / pa / pb / plane_*): nullptr = plain bf16 operands.
struct Split3 {
  int64_t plane_a, plane_b0, plane_b1;    // outer-coordinate distance between the three planes of A / B0 / B1
};
static void set_split3(TcParams& p, const Split3* s3) {
  if (!s3) { p.nprod = 1; return; }
  // plane pairs (i, j), i + j <= 2, smallest terms first, x0.y0 last (profiles/r02_fp32_split_study.jsonl: 2.5e-6 at K = 2560 in this
  // order against 2.6e-5 largest-first)
  static const int PA[6] = {2, 1, 0, 1, 0, 0}, PB[6] = {0, 1, 2, 0, 1, 0};
  p.nprod = 6;
  for (int j = 0; j < 6; ++j) { p.pa[j] = PA[j]; p.pb[j] = PB[j]; }
  p.plane_a = (int)s3->plane_a; p.plane_b0 = (int)s3->plane_b0; p.plane_b1 = (int)s3->plane_b1;
  p.out_f32 = 1;
}

// ---------------------------------------------------------------------------------------------------------
static int g_cluster = -1;    // RGNN_TC_CLUSTER=1 disables the 2-CTA multicast clusters (debug / A-B comparison)

template <int BN, int CL, bool AUX, bool TWO = false>
static int launch_cl(const CUtensorMap& a, const CUtensorMap& b0, const CUtensorMap& b1, const TcParams& p, cudaStream_t st,
                     const CUtensorMap* o0, const CUtensorMap* o1) {
  constexpr size_t smem = 1024 + (TWO ? (size_t)6 * (BM * BK * 2 + BN / 2 * BK * 2) : (size_t)NSTAGE * (BM * BK * 2 + BN * BK * 2)) + NOUT_STAGE * 128 * 128 + 384;
  static DeviceOnce once;
  if (once.need()) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm_tc<BN, CL, AUX, TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_gemm_tc)");
    once.mark();
  }
  const int64_t work = (int64_t)((p.tiles_m + CL - 1) / CL) * p.tiles_n * p.splits;
  int64_t groups = work < num_sms() / CL ? work : num_sms() / CL;
  if (groups < 1) return 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(groups * CL));
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  count_launch();
  cudaError_t e = cudaLaunchKernelEx(&cfg, k_gemm_tc<BN, CL, AUX, TWO>, a, b0, b1, o0 ? *o0 : a, o1 ? *o1 : (o0 ? *o0 : a), p);
  if (e != cudaSuccess) return cuda_fail(e, "k_gemm_tc launch");
  return 0;
}

template <int BN>
static int launch(const CUtensorMap& a, const CUtensorMap& b0, const CUtensorMap& b1, const TcParams& p, cudaStream_t st,
                  const CUtensorMap* o0 = nullptr, const CUtensorMap* o1 = nullptr) {
  if (g_cluster < 0) {
    const char* e = getenv("RGNN_TC_CLUSTER");
    g_cluster = (e && e[0] == '1') ? 1 : 2;
  }
  // clusters only pay when there are at least two M-tiles to pair up
  const bool aux = p.epi_mode == EPI_ROW && (p.epi_flags & (RGNN_EPI_ADD_AUX | RGNN_EPI_RELU_MASK)) != 0;
  static int two = -1;          // cta_group::2 CTA pairs for the 256-wide tiles without an in-kernel row sum; RGNN_TC_2CTA=0: multicast pairs
  if (two < 0) { const char* e = getenv("RGNN_TC_2CTA"); two = (e && e[0] == '0') ? 0 : 1; }
  if constexpr (BN == 256) {
    if (two == 1 && g_cluster == 2 && p.tiles_m >= 2)
      return aux ? launch_cl<BN, 2, true, true>(a, b0, b1, p, st, o0, o1) : launch_cl<BN, 2, false, true>(a, b0, b1, p, st, o0, o1);
  }
  if (g_cluster == 2 && p.tiles_m >= 2) return aux ? launch_cl<BN, 2, true>(a, b0, b1, p, st, o0, o1) : launch_cl<BN, 2, false>(a, b0, b1, p, st, o0, o1);
  return aux ? launch_cl<BN, 1, true>(a, b0, b1, p, st, o0, o1) : launch_cl<BN, 1, false>(a, b0, b1, p, st, o0, o1);
}

}  // namespace tc
}  // namespace rgnn

using namespace rgnn;
using namespace rgnn::tc;

extern "C" int rgnn_tc_cast_weights(const float* w, int64_t rows, int64_t cols, void* out_bf16, void* out_t_bf16, void* stream) {
  RGNN_REQUIRE(w && (out_bf16 || out_t_bf16) && rows > 0 && cols > 0 && rows < (1 << 30) && cols < (1 << 30), RGNN_E_ARG,
               "rgnn_tc_cast_weights: bad arguments");
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
  k_cast_weights<<<grid, 256, 0, as_stream(stream)>>>(w, (int)rows, (int)cols, static_cast<__nv_bfloat16*>(out_bf16),
                                                      static_cast<__nv_bfloat16*>(out_t_bf16), 1.0f, 0.0f);
  RGNN_LAUNCH_CHECK("k_cast_weights");
  return 0;
}

extern "C" int rgnn_tc_cast_weights_affine(const float* w, int64_t rows, int64_t cols, float alpha, float ident, void* out_bf16, void* out_t_bf16,
                                           void* stream) {
  RGNN_REQUIRE(w && (out_bf16 || out_t_bf16) && rows > 0 && cols > 0 && rows < (1 << 30) && cols < (1 << 30), RGNN_E_ARG,
               "rgnn_tc_cast_weights_affine: bad arguments");
  RGNN_REQUIRE(ident == 0.0f || rows == cols, RGNN_E_ARG, "rgnn_tc_cast_weights_affine: the identity term needs a square matrix");
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
  k_cast_weights<<<grid, 256, 0, as_stream(stream)>>>(w, (int)rows, (int)cols, static_cast<__nv_bfloat16*>(out_bf16),
                                                      static_cast<__nv_bfloat16*>(out_t_bf16), alpha, ident);
  RGNN_LAUNCH_CHECK("k_cast_weights");
  return 0;
}

// out[v, m, p] = sum_{k<c0} W[m,k] x0[v,k,p] + sum_{k<c1} W[m,c0+k] x1[v,k,p] (+ bias[m]);  rows m < m0 -> out0 [nodes, m0, 64],
// rows >= m0 -> out1 [nodes, m1, 64].  All bf16, P = 64 pixels.
static int conv1x1_impl(const void* W, const void* x0, int64_t c0, const void* x1, int64_t c1,
                        void* out0, int64_t m0, void* out1, int64_t m1, const float* bias, int64_t nodes, void* stream, const Split3* s3) {
  RGNN_REQUIRE(W && x0 && out0 && c0 > 0 && m0 > 0 && nodes >= 0 && c1 >= 0 && m1 >= 0, RGNN_E_ARG, "rgnn_tc_conv1x1: bad arguments");
  RGNN_REQUIRE((c1 == 0 || x1) && (m1 == 0 || out1), RGNN_E_ARG, "rgnn_tc_conv1x1: missing second operand/output");
  RGNN_REQUIRE(c0 % 64 == 0 && c1 % 64 == 0 && m0 % 128 == 0 && m1 % 128 == 0, RGNN_E_UNSUPPORTED,
               "rgnn_tc_conv1x1: channels must be multiples of 64 (inputs) / 128 (outputs)");
  RGNN_REQUIRE(al16(W) && al16(x0) && al16(out0) && (!x1 || al16(x1)) && (!out1 || al16(out1)), RGNN_E_ALIGN, "rgnn_tc_conv1x1: 16-byte alignment");
  RGNN_REQUIRE(nodes < (1ll << 30), RGNN_E_RANGE, "rgnn_tc_conv1x1: too many robots");
  if (nodes == 0) return 0;
  const int64_t K = c0 + c1, M = m0 + m1;
  constexpr int BN = 256;
  // split-plane operands: W is [3, M, K]; x0 / x1 are [3, plane_b, c, 64] (plane_b >= nodes robots per plane)
  const int64_t np = s3 ? 3 : 1;
  const int64_t nodes_b0 = s3 ? 2 * s3->plane_b0 + nodes : nodes, nodes_b1 = s3 ? 2 * s3->plane_b1 + nodes : nodes;
  RGNN_REQUIRE(!s3 || (s3->plane_a == M && s3->plane_b0 >= nodes && s3->plane_b1 >= nodes && 3 * s3->plane_b0 < (1ll << 30) && 3 * M < (1ll << 30)),
               RGNN_E_ARG, "rgnn_tc3_conv1x1: bad plane strides");
  CUtensorMap ta, tb0, tb1;
  {
    const uint64_t d[2] = {(uint64_t)K, (uint64_t)(np * M)}, s[1] = {(uint64_t)K * 2};
    const uint32_t b[2] = {64, 128};
    int rc = make_map(&ta, W, 2, d, s, b); if (rc) return rc;
  }
  {
    const uint64_t d[3] = {64, (uint64_t)c0, (uint64_t)nodes_b0}, s[2] = {128, (uint64_t)c0 * 128};
    const uint32_t b[3] = {64, 64, BN / 128};        // half of the tile's robots per box (see k_gemm_tc, CL)
    int rc = make_map(&tb0, x0, 3, d, s, b); if (rc) return rc;
  }
  if (c1 > 0) {
    const uint64_t d[3] = {64, (uint64_t)c1, (uint64_t)nodes_b1}, s[2] = {128, (uint64_t)c1 * 128};
    const uint32_t b[3] = {64, 64, BN / 128};
    int rc = make_map(&tb1, x1, 3, d, s, b); if (rc) return rc;
  } else tb1 = tb0;
  TcParams p = {};
  set_split3(p, s3);
  p.load_mode = LOAD_CONV; p.epi_mode = EPI_NCHW;
  // short reductions (dX at F = 512: 8 k-blocks per tile) are paced by the epilogue, not the MMAs: both warpgroups then take two
  // slabs each (dX 1.25 -> 1.12 ms); long ones keep one warpgroup with both staging buffers (the split costs the forward 0.05 ms).
  // RGNN_TC_NCHW_SPLIT=0|1 forces either.
  { static int sp = -2; if (sp == -2) { const char* e = getenv("RGNN_TC_NCHW_SPLIT"); sp = e ? atoi(e) : -1; }
    p.nchw_split = sp >= 0 ? sp : ((K / 64) * (s3 ? 6 : 1) <= 12); } p.epi_flags = bias ? RGNN_EPI_BIAS_M : 0; p.alpha = 1.0f;
  p.tiles_m = (int)(M / 128); p.tiles_n = (int)((nodes + BN / 64 - 1) / (BN / 64)); p.splits = 1;
  p.kblocks = (int)(K / 64); p.kb_per_split = p.kblocks; p.kb_b0 = (int)(c0 / 64);
  p.nodes = (int)nodes; p.M = M; p.N = nodes * 64;
  p.out0 = out0; p.out1 = out1; p.m_split = (int)m0; p.ld0 = m0 * 64; p.ld1 = m1 * 64; p.bias = bias;
  RGNN_REQUIRE((int64_t)p.tiles_m * p.tiles_n < (1ll << 31), RGNN_E_RANGE, "rgnn_tc_conv1x1: too many tiles");
  RGNN_REQUIRE(nodes * (m0 > m1 ? m0 : m1) < (1ll << 31), RGNN_E_RANGE, "rgnn_tc_conv1x1: too many output rows for a tensor map coordinate");
  CUtensorMap to0, to1;
  const uint64_t orow = s3 ? 128 : 64;      // fp32 outputs: a 64-pixel row is 128 16-bit units (256 B), stored as two 128-byte boxes
  {
    const uint64_t d[2] = {orow, (uint64_t)(nodes * m0)}, s[1] = {orow * 2};
    const uint32_t b[2] = {64, 128};
    int rc = make_map(&to0, out0, 2, d, s, b); if (rc) return rc;
  }
  if (m1 > 0) {
    const uint64_t d[2] = {orow, (uint64_t)(nodes * m1)}, s[1] = {orow * 2};
    const uint32_t b[2] = {64, 128};
    int rc = make_map(&to1, out1, 2, d, s, b); if (rc) return rc;
  } else to1 = to0;
  return launch<BN>(ta, tb0, tb1, p, as_stream(stream), &to0, &to1);
}

extern "C" int rgnn_tc_conv1x1(const void* W, const void* x0, int64_t c0, const void* x1, int64_t c1,
                               void* out0, int64_t m0, void* out1, int64_t m1, const float* bias, int64_t nodes, void* stream) {
  return conv1x1_impl(W, x0, c0, x1, c1, out0, m0, out1, m1, bias, nodes, stream, nullptr);
}

extern "C" int rgnn_tc3_conv1x1(const void* W3, const void* x0_3, int64_t c0, const void* x1_3, int64_t c1, int64_t plane_nodes,
                                float* out0, int64_t m0, float* out1, int64_t m1, const float* bias, int64_t nodes, void* stream) {
  const Split3 s3 = {m0 + m1, plane_nodes, plane_nodes};
  return conv1x1_impl(W3, x0_3, c0, x1_3, c1, out0, m0, out1, m1, bias, nodes, stream, &s3);
}

// dW[m, k] = sum_v sum_p dy[v,m,p] * x[v,k,p], x = x0 (k < c0) | x1;  dW fp32 [M, c0+c1] (row stride ldw), partial fp32 [splits, M, c0+c1].
static int wgrad_impl(const void* dy, int64_t M, const void* x0, int64_t c0, const void* x1, int64_t c1,
                      float* dW, int64_t ldw, float* dbias, float* partial, int64_t splits, int64_t nodes, void* stream, const Split3* s3) {
  RGNN_REQUIRE(dy && x0 && dW && partial && M > 0 && c0 > 0 && c1 >= 0 && nodes >= 0 && splits >= 1, RGNN_E_ARG, "rgnn_tc_conv1x1_wgrad: bad arguments");
  RGNN_REQUIRE(c1 == 0 || x1, RGNN_E_ARG, "rgnn_tc_conv1x1_wgrad: missing x1");
  constexpr int BN = 256;
  RGNN_REQUIRE(M % 128 == 0 && c0 % BN == 0 && c1 % BN == 0, RGNN_E_UNSUPPORTED, "rgnn_tc_conv1x1_wgrad: channel counts must be multiples of 128 / 256");
  RGNN_REQUIRE(al16(dy) && al16(x0) && (!x1 || al16(x1)) && al16(partial), RGNN_E_ALIGN, "rgnn_tc_conv1x1_wgrad: 16-byte alignment");
  RGNN_REQUIRE(nodes * M < (1ll << 31) && nodes * (c0 > c1 ? c0 : c1) < (1ll << 31), RGNN_E_RANGE, "rgnn_tc_conv1x1_wgrad: too many rows for a tensor map coordinate");
  const int64_t Nn = c0 + c1;
  cudaStream_t st = as_stream(stream);
  if (nodes == 0) {
    cudaError_t e = cudaMemset2DAsync(dW, ldw * 4, 0, Nn * 4, M, st);
    if (e == cudaSuccess && dbias) e = cudaMemsetAsync(dbias, 0, M * 4, st);
    return e == cudaSuccess ? 0 : cuda_fail(e, "memset dW");
  }
  // split-plane operands: dy is [3, plane_a robots, M, 64], x0 / x1 are [3, plane_b robots, c, 64]; the plane strides handed to the
  // kernel are in ROWS of the 2-D maps (robots x channels)
  RGNN_REQUIRE(!s3 || (!dbias && s3->plane_a >= nodes && s3->plane_b0 >= nodes && s3->plane_b1 >= nodes), RGNN_E_ARG, "rgnn_tc3_conv1x1_wgrad: bad plane strides");
  RGNN_REQUIRE(!s3 || (3 * s3->plane_a * M < (1ll << 31) && 3 * s3->plane_b0 * (c0 > c1 ? c0 : c1) < (1ll << 31)), RGNN_E_RANGE,
               "rgnn_tc3_conv1x1_wgrad: too many rows for a tensor map coordinate");
  const int64_t na = s3 ? 2 * s3->plane_a + nodes : nodes, nb0 = s3 ? 2 * s3->plane_b0 + nodes : nodes, nb1 = s3 ? 2 * s3->plane_b1 + nodes : nodes;
  CUtensorMap ta, tb0, tb1;
  {
    const uint64_t d[2] = {64, (uint64_t)(na * M)}, s[1] = {128};
    const uint32_t b[2] = {64, 128};
    int rc = make_map(&ta, dy, 2, d, s, b); if (rc) return rc;
  }
  {
    const uint64_t d[2] = {64, (uint64_t)(nb0 * c0)}, s[1] = {128};
    const uint32_t b[2] = {64, BN / 2};
    int rc = make_map(&tb0, x0, 2, d, s, b); if (rc) return rc;
  }
  if (c1 > 0) {
    const uint64_t d[2] = {64, (uint64_t)(nb1 * c1)}, s[1] = {128};
    const uint32_t b[2] = {64, BN / 2};
    int rc = make_map(&tb1, x1, 2, d, s, b); if (rc) return rc;
  } else tb1 = tb0;
  if (splits > nodes) splits = nodes;
  TcParams p = {};
  if (s3) {
    const Split3 rows = {s3->plane_a * M, s3->plane_b0 * c0, s3->plane_b1 * c1};
    set_split3(p, &rows);
    p.out_f32 = 0;                 // partials are fp32 anyway
  } else p.nprod = 1;
  p.load_mode = LOAD_WGRAD; p.epi_mode = EPI_PARTIAL;
  p.tiles_m = (int)(M / 128); p.tiles_n = (int)(Nn / BN); p.splits = (int)splits;
  p.kblocks = (int)nodes; p.kb_per_split = (int)((nodes + splits - 1) / splits);
  splits = (nodes + p.kb_per_split - 1) / p.kb_per_split;   // no empty split
  p.splits = (int)splits;
  p.n_b0 = (int)c0; p.rows_a = (int)M; p.rows_b0 = (int)c0; p.rows_b1 = (int)c1;
  p.M = M; p.N = Nn; p.out0 = partial; p.ld0 = Nn;
  p.rowsum = dbias ? partial + splits * M * Nn : nullptr;     // [splits, M] behind the tile partials
  if (splits == 1 && ldw % 4 == 0 && (reinterpret_cast<uintptr_t>(dW) & 15u) == 0) {
    // one split (small batches, launch-bound): the tiles ARE the result - no partial buffer, no reduction launch
    p.out0 = dW; p.ld0 = ldw; p.rowsum = dbias;
    return launch<BN>(ta, tb0, tb1, p, st);
  }
  int rc = launch<BN>(ta, tb0, tb1, p, st);
  if (rc) return rc;
  const int64_t n = M * Nn;
  int64_t blocks = (n + (dbias ? M : 0) + 255) / 256;
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  k_sum_partials<<<(unsigned)blocks, 256, 0, st>>>(partial, (int)splits, n, dW, ldw, Nn, 1.0f, dbias ? p.rowsum : nullptr, dbias ? M : 0, dbias);
  RGNN_LAUNCH_CHECK("k_sum_partials");
  return 0;
}

extern "C" int rgnn_tc_conv1x1_wgrad(const void* dy, int64_t M, const void* x0, int64_t c0, const void* x1, int64_t c1,
                                     float* dW, int64_t ldw, float* dbias, float* partial, int64_t splits, int64_t nodes, void* stream) {
  return wgrad_impl(dy, M, x0, c0, x1, c1, dW, ldw, dbias, partial, splits, nodes, stream, nullptr);
}

extern "C" int rgnn_tc3_conv1x1_wgrad(const void* dy3, int64_t M, const void* x0_3, int64_t c0, const void* x1_3, int64_t c1, int64_t plane_nodes,
                                      float* dW, int64_t ldw, float* partial, int64_t splits, int64_t nodes, void* stream) {
  const Split3 s3 = {plane_nodes, plane_nodes, plane_nodes};
  return wgrad_impl(dy3, M, x0_3, c0, x1_3, c1, dW, ldw, nullptr, partial, splits, nodes, stream, &s3);
}

// C[M,N] = epi(A[M,K] . B[N,K]^T): A, B bf16 K-major (row strides lda, ldb elements); C bf16 row-major (ldc); bias fp32 [N];
// aux bf16 [M, ldc] for RGNN_EPI_RELU_MASK.  epi: bit-or of RGNN_EPI_BIAS_N | RGNN_EPI_SIGMOID | RGNN_EPI_RELU_MASK.
static int gemm_tn_impl(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                        int64_t M, int64_t N, int64_t K, int epi, const float* bias, const void* aux, float alpha, float beta,
                        void* stream, const Split3* s3) {
  RGNN_REQUIRE(A && B && C && M >= 0 && N > 0 && K > 0, RGNN_E_ARG, "rgnn_tc_gemm_tn: bad arguments");
  RGNN_REQUIRE((epi & ~(RGNN_EPI_BIAS_N | RGNN_EPI_SIGMOID | RGNN_EPI_RELU_MASK | RGNN_EPI_ADD_AUX)) == 0, RGNN_E_UNSUPPORTED, "rgnn_tc_gemm_tn: unsupported epilogue");
  RGNN_REQUIRE(!((epi & RGNN_EPI_RELU_MASK) && (epi & RGNN_EPI_ADD_AUX)), RGNN_E_ARG, "rgnn_tc_gemm_tn: aux is either the ReLU mask or the addend");
  RGNN_REQUIRE(!(epi & RGNN_EPI_BIAS_N) || bias, RGNN_E_ARG, "rgnn_tc_gemm_tn: bias missing");
  RGNN_REQUIRE(!(epi & (RGNN_EPI_RELU_MASK | RGNN_EPI_ADD_AUX)) || aux, RGNN_E_ARG, "rgnn_tc_gemm_tn: aux missing");
  RGNN_REQUIRE(K % 64 == 0 && N % 128 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldc % 8 == 0, RGNN_E_UNSUPPORTED,
               "rgnn_tc_gemm_tn: K %% 64, N %% 128 and 16-byte row strides required");
  RGNN_REQUIRE(al16(A) && al16(B) && al16(C) && (!aux || al16(aux)), RGNN_E_ALIGN, "rgnn_tc_gemm_tn: 16-byte alignment");
  RGNN_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), RGNN_E_RANGE, "rgnn_tc_gemm_tn: extent too large");
  if (M == 0) return 0;
  const bool wide = N % 256 == 0;
  const int BN = wide ? 256 : 128;
  // split-plane operands: A is [3, plane_a rows, lda], B is [3, plane_b0 rows, ldb]; C fp32
  RGNN_REQUIRE(!s3 || (!(epi & (RGNN_EPI_RELU_MASK | RGNN_EPI_ADD_AUX)) && s3->plane_a >= M && s3->plane_b0 >= N && 3 * s3->plane_a < (1ll << 31) && ldc % 4 == 0),
               RGNN_E_ARG, "rgnn_tc3_gemm_tn: bad plane strides / unsupported epilogue");
  CUtensorMap ta, tb;
  {
    const uint64_t d[2] = {(uint64_t)K, (uint64_t)(s3 ? 2 * s3->plane_a + M : M)}, s[1] = {(uint64_t)lda * 2};
    const uint32_t b[2] = {64, 128};
    int rc = make_map(&ta, A, 2, d, s, b); if (rc) return rc;
  }
  {
    const uint64_t d[2] = {(uint64_t)K, (uint64_t)(s3 ? 2 * s3->plane_b0 + N : N)}, s[1] = {(uint64_t)ldb * 2};
    const uint32_t b[2] = {64, (uint32_t)BN / 2};
    int rc = make_map(&tb, B, 2, d, s, b); if (rc) return rc;
  }
  TcParams p = {};
  set_split3(p, s3);
  p.load_mode = LOAD_TN; p.epi_mode = EPI_ROW; p.epi_flags = epi;
  p.n_fast = 1;          // activations (A, rows = edges / pixels) are the big operand: its N-tile siblings run back to back and hit L2
  p.tiles_m = (int)((M + 127) / 128); p.tiles_n = (int)(N / BN); p.splits = 1;
  p.kblocks = (int)(K / 64); p.kb_per_split = p.kblocks;
  p.M = M; p.N = N; p.out0 = C; p.ld0 = ldc; p.bias = bias; p.aux = aux; p.alpha = alpha; p.beta = beta;
  RGNN_REQUIRE(!bias || (reinterpret_cast<uintptr_t>(bias) & 15u) == 0, RGNN_E_ALIGN, "rgnn_tc_gemm_tn: bias must be 16-byte aligned");
  CUtensorMap to;
  {
    const uint64_t ou = s3 ? 2 : 1;         // fp32 C described as pairs of 16-bit units
    const uint64_t d[2] = {(uint64_t)N * ou, (uint64_t)M}, s[1] = {(uint64_t)ldc * 2 * ou};
    const uint32_t b[2] = {64, 128};
    int rc = make_map(&to, C, 2, d, s, b); if (rc) return rc;
  }
  return wide ? launch<256>(ta, tb, tb, p, as_stream(stream), &to, &to) : launch<128>(ta, tb, tb, p, as_stream(stream), &to, &to);
}

extern "C" int rgnn_tc_gemm_tn(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                               int64_t M, int64_t N, int64_t K, int epi, const float* bias, const void* aux, float alpha, float beta,
                               void* stream) {
  return gemm_tn_impl(A, lda, B, ldb, C, ldc, M, N, K, epi, bias, aux, alpha, beta, stream, nullptr);
}

extern "C" int rgnn_tc3_gemm_tn(const void* A3, int64_t lda, int64_t plane_a, const void* B3, int64_t ldb, int64_t plane_b, float* C, int64_t ldc,
                                int64_t M, int64_t N, int64_t K, int epi, const float* bias, float alpha, void* stream) {
  const Split3 s3 = {plane_a, plane_b, plane_b};
  return gemm_tn_impl(A3, lda, B3, ldb, C, ldc, M, N, K, epi, bias, nullptr, alpha, 0.0f, stream, &s3);
}

// C[M,N] fp32 = A^T . B with A bf16 [K, M] (row stride lda), B bf16 [K, N] (row stride ldb): both operands MN-major.
// Split over K; partial fp32 [splits, M, N].  (dW2 = dZ^T . A1 of the edge encoder.)
static int gemm_nn_impl(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc,
                        int64_t M, int64_t N, int64_t K, float alpha, float* partial, int64_t splits, void* stream, const Split3* s3) {
  RGNN_REQUIRE(A && B && C && partial && M > 0 && N > 0 && K >= 0 && splits >= 1, RGNN_E_ARG, "rgnn_tc_gemm_nn: bad arguments");
  RGNN_REQUIRE(M % 128 == 0 && N % 128 == 0 && lda % 8 == 0 && ldb % 8 == 0, RGNN_E_UNSUPPORTED, "rgnn_tc_gemm_nn: M, N %% 128 and 16-byte row strides required");
  RGNN_REQUIRE(al16(A) && al16(B) && al16(partial), RGNN_E_ALIGN, "rgnn_tc_gemm_nn: 16-byte alignment");
  RGNN_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), RGNN_E_RANGE, "rgnn_tc_gemm_nn: extent too large");
  cudaStream_t st = as_stream(stream);
  if (K == 0) {
    cudaError_t e = cudaMemset2DAsync(C, ldc * 4, 0, N * 4, M, st);
    return e == cudaSuccess ? 0 : cuda_fail(e, "memset C");
  }
  const bool wide = N % 256 == 0;
  const int BN = wide ? 256 : 128;
  // split-plane operands: A is [3, plane_a rows (K, zero beyond K up to a multiple of 64), lda], B likewise
  RGNN_REQUIRE(!s3 || (s3->plane_a >= ((K + 63) / 64) * 64 && s3->plane_b0 >= ((K + 63) / 64) * 64 && 3 * s3->plane_a < (1ll << 31) && 3 * s3->plane_b0 < (1ll << 31)),
               RGNN_E_ARG, "rgnn_tc3_gemm_nn: planes must be padded (zero-filled) to a multiple of 64 rows");
  CUtensorMap ta, tb;
  {
    const uint64_t d[2] = {(uint64_t)M, (uint64_t)(s3 ? 3 * s3->plane_a : K)}, s[1] = {(uint64_t)lda * 2};
    const uint32_t b[2] = {64, 64};
    int rc = make_map(&ta, A, 2, d, s, b); if (rc) return rc;
  }
  {
    const uint64_t d[2] = {(uint64_t)N, (uint64_t)(s3 ? 3 * s3->plane_b0 : K)}, s[1] = {(uint64_t)ldb * 2};
    const uint32_t b[2] = {64, 64};
    int rc = make_map(&tb, B, 2, d, s, b); if (rc) return rc;
  }
  const int64_t kblocks = (K + 63) / 64;
  if (splits > kblocks) splits = kblocks;
  TcParams p = {};
  set_split3(p, s3);
  p.out_f32 = 0;
  p.load_mode = LOAD_NN; p.epi_mode = EPI_PARTIAL;
  p.tiles_m = (int)(M / 128); p.tiles_n = (int)(N / BN); p.splits = (int)splits;
  p.kblocks = (int)kblocks; p.kb_per_split = (int)((kblocks + splits - 1) / splits);
  splits = (kblocks + p.kb_per_split - 1) / p.kb_per_split;   // no empty split
  p.splits = (int)splits;
  p.M = M; p.N = N; p.out0 = partial; p.ld0 = N;
  if (splits == 1 && alpha == 1.0f && ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(C) & 15u) == 0) {      // as in the weight-gradient GEMM above
    p.out0 = C; p.ld0 = ldc;
    return wide ? launch<256>(ta, tb, tb, p, st) : launch<128>(ta, tb, tb, p, st);
  }
  int rc = wide ? launch<256>(ta, tb, tb, p, st) : launch<128>(ta, tb, tb, p, st);
  if (rc) return rc;
  const int64_t n = M * N;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  k_sum_partials<<<(unsigned)blocks, 256, 0, st>>>(partial, (int)splits, n, C, ldc, N, alpha);
  RGNN_LAUNCH_CHECK("k_sum_partials");
  return 0;
}

extern "C" int rgnn_tc_gemm_nn(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc,
                               int64_t M, int64_t N, int64_t K, float alpha, float* partial, int64_t splits, void* stream) {
  return gemm_nn_impl(A, lda, B, ldb, C, ldc, M, N, K, alpha, partial, splits, stream, nullptr);
}

extern "C" int rgnn_tc3_gemm_nn(const void* A3, int64_t lda, int64_t plane_a, const void* B3, int64_t ldb, int64_t plane_b, float* C, int64_t ldc,
                                int64_t M, int64_t N, int64_t K, float alpha, float* partial, int64_t splits, void* stream) {
  const Split3 s3 = {plane_a, plane_b, plane_b};
  return gemm_nn_impl(A3, lda, B3, ldb, C, ldc, M, N, K, alpha, partial, splits, stream, &s3);
}

// x fp32 [n] -> bf16 planes [3, plane] with x = x0 + x1 + x2; elements [n, plane) of every plane are zero.  n %% 4 == 0, plane %% 8 == 0.
extern "C" int rgnn_tc_split3(const float* x, int64_t n, int64_t plane, void* out_bf16, void* stream) {
  RGNN_REQUIRE(x && out_bf16 && n >= 0 && plane >= n, RGNN_E_ARG, "rgnn_tc_split3: bad arguments");
  RGNN_REQUIRE(n % 4 == 0 && plane % 8 == 0 && al16(x) && al16(out_bf16), RGNN_E_ALIGN, "rgnn_tc_split3: n %% 4, plane %% 8 and 16-byte alignment required");
  if (plane == 0) return 0;
  int64_t blocks = (plane / 4 + 255) / 256;
  if (blocks > 16 * num_sms()) blocks = 16 * num_sms();
  k_split3<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, n / 4, plane / 4, static_cast<__nv_bfloat16*>(out_bf16));
  RGNN_LAUNCH_CHECK("k_split3");
  return 0;
}

// fp32 master weights [rows, cols] -> bf16 planes [3, rows, cols] and/or planes of the transpose [3, cols, rows] (either may be NULL)
extern "C" int rgnn_tc_split3_weights(const float* w, int64_t rows, int64_t cols, void* out3, void* out3_t, void* stream) {
  RGNN_REQUIRE(w && (out3 || out3_t) && rows > 0 && cols > 0 && rows < (1 << 30) && cols < (1 << 30), RGNN_E_ARG,
               "rgnn_tc_split3_weights: bad arguments");
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
  k_split3_weights<<<grid, 256, 0, as_stream(stream)>>>(w, (int)rows, (int)cols, static_cast<__nv_bfloat16*>(out3), static_cast<__nv_bfloat16*>(out3_t));
  RGNN_LAUNCH_CHECK("k_split3_weights");
  return 0;
}
