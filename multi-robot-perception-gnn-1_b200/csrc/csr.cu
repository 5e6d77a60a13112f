// A2/A3: batched-graph index work - replaces dgl.batch (dgl/training.py:57-58) and the
// in_degrees / in_edges(form='eid') / degree-norm work DGL does inside update_all
// (call site dgl/model/models.py:223).  Integer outputs are bit-exact with DGL's conventions:
// stable order (ascending edge id inside a node), int64 ids.
//
// Layout choice: a robot batch is G small independent graphs whose nodes and edges are contiguous,
// so ONE WARP OWNS ONE GRAPH.  Its counters live in the output arrays themselves, nobody else touches
// them, and no global scan / atomic / sort is needed: the CSR slice of graph g starts at edge_offsets[g].
// Stability comes from walking the edge list 32 edges at a time in order and ranking equal keys inside a
// chunk with __match_any_sync.  Grid = multiple of 148 SMs, warps stride over graphs.
#include "common.cuh"

namespace rgnn {

// exclusive scan of two int64 arrays of length G into [G+1] outputs, single CTA (G <= a few 100k).
__global__ void __launch_bounds__(1024) k_batch_offsets(const int64_t* __restrict__ bnn, const int64_t* __restrict__ bne,
                                                        int64_t G, int64_t* __restrict__ noff, int64_t* __restrict__ eoff) {
  __shared__ long long wsum[2][32];
  __shared__ long long carry[2];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) { carry[0] = 0; carry[1] = 0; noff[0] = 0; eoff[0] = 0; }
  __syncthreads();
  for (int64_t base = 0; base < G; base += blockDim.x) {
    const int64_t i = base + tid;
    long long v[2] = {i < G ? bnn[i] : 0, i < G ? bne[i] : 0};
    long long inc[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      long long x = v[a];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        long long y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      inc[a] = x;
      if (lane == 31) wsum[a][wid] = x;
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        long long x = lane < (blockDim.x >> 5) ? wsum[a][lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          long long y = __shfl_up_sync(0xffffffffu, x, o);
          if (lane >= o) x += y;
        }
        wsum[a][lane] = x;  // inclusive over warps
      }
    }
    __syncthreads();
    const long long c0 = carry[0] + (wid ? wsum[0][wid - 1] : 0);
    const long long c1 = carry[1] + (wid ? wsum[1][wid - 1] : 0);
    if (i < G) { noff[i + 1] = c0 + inc[0]; eoff[i + 1] = c1 + inc[1]; }
    __syncthreads();
    if (tid == 0) { carry[0] += wsum[0][(blockDim.x >> 5) - 1]; carry[1] += wsum[1][(blockDim.x >> 5) - 1]; }
    __syncthreads();
  }
}

// global id = local id + node offset of the owning graph; one warp per graph, coalesced over its edges.
__global__ void __launch_bounds__(256) k_shift_ids(const int64_t* __restrict__ lsrc, const int64_t* __restrict__ ldst,
                                                   const int64_t* __restrict__ noff, const int64_t* __restrict__ eoff,
                                                   int64_t G, int64_t* __restrict__ src, int64_t* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t g = warp; g < G; g += nwarps) {
    const int64_t off = noff[g], e0 = eoff[g], e1 = eoff[g + 1];
    for (int64_t e = e0 + lane; e < e1; e += 32) {
      src[e] = lsrc[e] + off;
      dst[e] = ldst[e] + off;
    }
  }
}

struct CsrOut {
  int64_t *in_deg, *out_deg, *in_ptr, *in_src, *in_eid, *out_ptr, *out_dst, *out_eid;
  int32_t *in_ptr32, *out_ptr32, *noff32, *in_src32, *in_eid32, *out_dst32, *out_eid32, *src32, *dst32;
  float *inv_in, *rsqrt_in, *rsqrt_out;
  int32_t* status;
};

// One warp per graph.  Phase 1: degrees (cursor arrays zeroed, then counted with match_any ranks).
// Phase 2: local exclusive scan -> ptr (= edge offset of the graph + local prefix), norms, cursors := ptr.
// Phase 3: stable placement.  The int32 ptr copies double as the running cursors and are rewritten at the end.
__global__ void __launch_bounds__(128) k_build_csr(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                                   const int64_t* __restrict__ noff, const int64_t* __restrict__ eoff,
                                                   int64_t G, int64_t N, int64_t E, CsrOut o) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  int32_t* cin = o.in_ptr32;    // cursor[v] for v in graph (entry N is written last, by graph G-1's warp)
  int32_t* cout = o.out_ptr32;
  for (int64_t g = warp; g < G; g += nwarps) {
    const int64_t n0 = noff[g], n1 = noff[g + 1], e0 = eoff[g], e1 = eoff[g + 1];
    if (lane == 0) { o.noff32[g] = (int32_t)n0; if (g == G - 1) o.noff32[G] = (int32_t)n1; }
    // ---- phase 1: degrees
    for (int64_t v = n0 + lane; v < n1; v += 32) { cin[v] = 0; cout[v] = 0; }
    __syncwarp();
    bool bad = false;
    for (int64_t base = e0; base < e1; base += 32) {
      const int64_t e = base + lane;
      const bool act = e < e1;
      int64_t s = act ? src[e] : -1, d = act ? dst[e] : -1;
      if (act && (s < n0 || s >= n1 || d < n0 || d >= n1)) { bad = true; s = d = -1; }
      const unsigned ms = __match_any_sync(0xffffffffu, s), md = __match_any_sync(0xffffffffu, d);
      if (s >= 0 && (ms & lt) == 0) cout[s] += __popc(ms);   // leader of each key adds the chunk count
      if (d >= 0 && (md & lt) == 0) cin[d] += __popc(md);
      __syncwarp();
    }
    const bool any_bad = __any_sync(0xffffffffu, bad);
    if (any_bad && lane == 0) atomicOr(o.status, 1);
    // ---- phase 2: exclusive scan over the graph's nodes
    int carry_in = (int)e0, carry_out = (int)e0;
    for (int64_t base = n0; base < n1; base += 32) {
      const int64_t v = base + lane;
      const int di = v < n1 ? cin[v] : 0, dq = v < n1 ? cout[v] : 0;
      int xi = di, xo = dq;
#pragma unroll
      for (int k = 1; k < 32; k <<= 1) {
        const int yi = __shfl_up_sync(0xffffffffu, xi, k), yo = __shfl_up_sync(0xffffffffu, xo, k);
        if (lane >= k) { xi += yi; xo += yo; }
      }
      if (v < n1) {
        const int pi = carry_in + xi - di, po = carry_out + xo - dq;
        cin[v] = pi; cout[v] = po;
        if (o.in_deg) o.in_deg[v] = di;
        if (o.out_deg) o.out_deg[v] = dq;
        if (o.in_ptr) o.in_ptr[v] = pi;
        if (o.out_ptr) o.out_ptr[v] = po;
        o.inv_in[v] = di > 0 ? __fdiv_rn(1.0f, (float)di) : 0.0f;
        o.rsqrt_in[v] = __fdiv_rn(1.0f, __fsqrt_rn((float)max(di, 1)));
        o.rsqrt_out[v] = __fdiv_rn(1.0f, __fsqrt_rn((float)max(dq, 1)));
      }
      carry_in += __shfl_sync(0xffffffffu, xi, 31);
      carry_out += __shfl_sync(0xffffffffu, xo, 31);
    }
    __syncwarp();
    // ---- phase 3: stable placement (cursors advance; restored below)
    for (int64_t base = e0; base < e1; base += 32) {
      const int64_t e = base + lane;
      const bool act = e < e1;
      int64_t s = act ? src[e] : -1, d = act ? dst[e] : -1;
      const bool out_of_graph = act && (s < n0 || s >= n1 || d < n0 || d >= n1);
      // an edge that leaves its graph (status bit 0; DGL raises at construction) is kept IN BOUNDS for the kernels that walk the
      // edge list: both endpoints collapse onto the graph's first node, so nothing is gathered from another graph or past N
      if (act) { o.src32[e] = (int32_t)(out_of_graph ? (n1 > n0 ? n0 : 0) : s); o.dst32[e] = (int32_t)(out_of_graph ? (n1 > n0 ? n0 : 0) : d); }
      if (out_of_graph) s = d = -1;
      const unsigned ms = __match_any_sync(0xffffffffu, s), md = __match_any_sync(0xffffffffu, d);
      if (s >= 0) {
        const int pos = cout[s] + __popc(ms & lt);
        o.out_dst32[pos] = (int32_t)d; o.out_eid32[pos] = (int32_t)e;
        if (o.out_dst) o.out_dst[pos] = d;
        if (o.out_eid) o.out_eid[pos] = e;
      }
      if (d >= 0) {
        const int pos = cin[d] + __popc(md & lt);
        o.in_src32[pos] = (int32_t)s; o.in_eid32[pos] = (int32_t)e;
        if (o.in_src) o.in_src[pos] = s;
        if (o.in_eid) o.in_eid[pos] = e;
      }
      __syncwarp();
      if (s >= 0 && (ms & lt) == 0) cout[s] += __popc(ms);
      if (d >= 0 && (md & lt) == 0) cin[d] += __popc(md);
      __syncwarp();
    }
    if (any_bad && n1 > n0) {
      // dropped edges leave unplaced slots at the end of this graph's CSR slice, and the last node's range runs up to the next
      // graph's first slot: fill them with (the last node itself, the graph's first edge) - in-bounds by construction instead
      // of uninitialised memory.  The batch is still flagged invalid (status bit 0, RobotGraph.validate()).
      __syncwarp();
      const int end_in = cin[n1 - 1], end_out = cout[n1 - 1];
      for (int64_t pos = end_in + lane; pos < e1; pos += 32) {
        o.in_src32[pos] = (int32_t)(n1 - 1); o.in_eid32[pos] = (int32_t)e0;
        if (o.in_src) o.in_src[pos] = n1 - 1;
        if (o.in_eid) o.in_eid[pos] = e0;
      }
      for (int64_t pos = end_out + lane; pos < e1; pos += 32) {
        o.out_dst32[pos] = (int32_t)(n1 - 1); o.out_eid32[pos] = (int32_t)e0;
        if (o.out_dst) o.out_dst[pos] = n1 - 1;
        if (o.out_eid) o.out_eid[pos] = e0;
      }
      __syncwarp();
    }
    // ---- restore ptr32 from the advanced cursors: cursor[v] is now ptr[v+1]; shift right by one node
    int prev_in = (int)e0, prev_out = (int)e0;
    for (int64_t base = n0; base < n1; base += 32) {
      const int64_t v = base + lane;
      const int ci = v < n1 ? cin[v] : 0, co = v < n1 ? cout[v] : 0;
      int pi = __shfl_up_sync(0xffffffffu, ci, 1), po = __shfl_up_sync(0xffffffffu, co, 1);
      if (lane == 0) { pi = prev_in; po = prev_out; }
      prev_in = __shfl_sync(0xffffffffu, ci, 31);
      prev_out = __shfl_sync(0xffffffffu, co, 31);
      __syncwarp();
      if (v < n1) { cin[v] = pi; cout[v] = po; }
    }
    if (g == G - 1 && lane == 0) {
      cin[N] = (int32_t)E; cout[N] = (int32_t)E;
      if (o.in_ptr) o.in_ptr[N] = E;
      if (o.out_ptr) o.out_ptr[N] = E;
    }
  }
}

}  // namespace rgnn

using namespace rgnn;

extern "C" int rgnn_batch_graphs(const int64_t* local_src, const int64_t* local_dst, const int64_t* bnn, const int64_t* bne,
                                 int64_t G, int64_t E, int64_t* noff, int64_t* eoff, int64_t* src, int64_t* dst, void* stream) {
  RGNN_REQUIRE(G >= 0 && E >= 0, RGNN_E_ARG, "rgnn_batch_graphs: negative size");
  RGNN_REQUIRE(bnn && bne && noff && eoff, RGNN_E_ARG, "rgnn_batch_graphs: null offsets/count pointer");
  RGNN_REQUIRE(E == 0 || (local_src && local_dst && src && dst), RGNN_E_ARG, "rgnn_batch_graphs: null edge pointer");
  cudaStream_t st = as_stream(stream);
  k_batch_offsets<<<1, 1024, 0, st>>>(bnn, bne, G, noff, eoff);
  RGNN_LAUNCH_CHECK("k_batch_offsets");
  if (E > 0 && G > 0) {
    const int wpb = 8;
    int64_t blocks = (G + wpb - 1) / wpb;
    if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
    k_shift_ids<<<(unsigned)blocks, wpb * 32, 0, st>>>(local_src, local_dst, noff, eoff, G, src, dst);
    RGNN_LAUNCH_CHECK("k_shift_ids");
  }
  return 0;
}

extern "C" int rgnn_build_batched_csr(const int64_t* src, const int64_t* dst, int64_t N, int64_t E,
                                      const int64_t* noff, const int64_t* eoff, int64_t G,
                                      int64_t* in_deg, int64_t* out_deg, int64_t* in_ptr, int64_t* in_src, int64_t* in_eid,
                                      int64_t* out_ptr, int64_t* out_dst, int64_t* out_eid,
                                      int32_t* ptr32, int32_t* idx32, float* norm, int32_t* status, void* stream) {
  RGNN_REQUIRE(N >= 0 && E >= 0 && G >= 0, RGNN_E_ARG, "rgnn_build_batched_csr: negative size");
  RGNN_REQUIRE(N < (1ll << 31) - 1 && E < (1ll << 31) - 1, RGNN_E_RANGE, "rgnn_build_batched_csr: N or E >= 2^31");
  RGNN_REQUIRE(ptr32 && norm && status && noff && eoff, RGNN_E_ARG, "rgnn_build_batched_csr: null required pointer");
  RGNN_REQUIRE(E == 0 || (src && dst && idx32), RGNN_E_ARG, "rgnn_build_batched_csr: null edge pointer");
  RGNN_REQUIRE(G > 0 || N == 0, RGNN_E_ARG, "rgnn_build_batched_csr: nodes without graphs");
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(status, 0, sizeof(int32_t), st);
  if (e != cudaSuccess) return cuda_fail(e, "memset status");
  CsrOut o;
  o.in_deg = in_deg; o.out_deg = out_deg; o.in_ptr = in_ptr; o.in_src = in_src; o.in_eid = in_eid;
  o.out_ptr = out_ptr; o.out_dst = out_dst; o.out_eid = out_eid;
  o.in_ptr32 = ptr32; o.out_ptr32 = ptr32 + (N + 1); o.noff32 = ptr32 + 2 * (N + 1);
  o.in_src32 = idx32; o.in_eid32 = idx32 + E; o.out_dst32 = idx32 + 2 * E; o.out_eid32 = idx32 + 3 * E;
  o.src32 = idx32 + 4 * E; o.dst32 = idx32 + 5 * E;
  o.inv_in = norm; o.rsqrt_in = norm + N; o.rsqrt_out = norm + 2 * N;
  o.status = status;
  if (G == 0) {  // empty batch: ptr = {0}
    e = cudaMemsetAsync(ptr32, 0, 3 * sizeof(int32_t), st);
    if (e != cudaSuccess) return cuda_fail(e, "memset ptr32");
    if (in_ptr) cudaMemsetAsync(in_ptr, 0, sizeof(int64_t), st);
    if (out_ptr) cudaMemsetAsync(out_ptr, 0, sizeof(int64_t), st);
    return 0;
  }
  const int wpb = 4;
  int64_t blocks = (G + wpb - 1) / wpb;
  if (blocks > 16 * num_sms()) blocks = 16 * num_sms();
  k_build_csr<<<(unsigned)blocks, wpb * 32, 0, st>>>(src, dst, noff, eoff, G, N, E, o);
  RGNN_LAUNCH_CHECK("k_build_csr");
  return 0;
}
