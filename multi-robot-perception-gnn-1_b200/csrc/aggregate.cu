// A5/A6 (+A10, A11): fused gather -> FiLM / normalise -> sum|mean aggregation.
// Replaces edge_udf + node_udf + DGL degree bucketing (dgl/model/models.py:207-211,223) and
// update_all(copy_u, sum|mean) (dgl/model/dgl_models.py:126).  The [E,F,P] message tensor of the reference
// (and its mailbox copy) is never materialised: HBM traffic is  read x once + read z once + write out once.
//
// Two kernels:
//  * k_agg_staged  - the robot-batch path.  One CTA owns (graph g, channel tile): the K node slices
//    x[n0..n1, c0:c0+CT, :] are K contiguous runs of CT*P*s bytes, fetched ONCE by the bulk-copy engine
//    (cp.async.bulk -> shared, mbarrier completion), and every destination of the graph reduces from shared
//    memory.  128-bit shared reads / global writes, consecutive lanes -> consecutive 16 B.
//  * k_agg_generic - any graph (huge K, ragged P): destination-centric gather straight from global / L2.
// Both keep the reference's arithmetic order (ascending edge id; gamma*x rounded, +beta rounded, sum, divide) so
// fp32 output is bit-identical to the reference's mailbox.mean(1).
#include "common.cuh"

namespace rgnn {

struct AggParams {
  const void* x; int64_t ldx;
  const void* z; int use_beta;
  const float* nbr_scale; int post; const float* dst_scale;
  const void* addend; int64_t ldadd;
  void* out; int64_t ldout; void* copy_to;
  const int32_t *ptr, *nbr, *eid, *noff;
  int64_t G, N, E; int F, P;
};

template <typename T> __device__ __forceinline__ float2 load_gb(const T* z, int64_t idx);
template <> __device__ __forceinline__ float2 load_gb<float>(const float* z, int64_t idx) {
  return __ldg(reinterpret_cast<const float2*>(z + idx));
}
template <> __device__ __forceinline__ float2 load_gb<__nv_bfloat16>(const __nv_bfloat16* z, int64_t idx) {
  const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(z + idx));
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

template <int V> __device__ __forceinline__ void film_acc(float (&acc)[V], const float (&xv)[V], bool has_z, float2 gb,
                                                          bool use_beta, float ns, bool has_ns) {
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float t = xv[i];
    if (has_ns) t = __fmul_rn(t, ns);
    if (has_z) {
      t = __fmul_rn(gb.x, t);
      if (use_beta) t = __fadd_rn(t, gb.y);
    }
    acc[i] = __fadd_rn(acc[i], t);
  }
}

template <int V> __device__ __forceinline__ void film_post(float (&acc)[V], int post, int deg, float ds) {
  if (post == RGNN_POST_MEAN) {
    if (deg > 0) {
      const float d = (float)deg;
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = __fdiv_rn(acc[i], d);
    }
  } else if (post == RGNN_POST_SCALE) {
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = __fmul_rn(acc[i], ds);
  }
}

// ------------------------------------------------------------------------------------------------
// generic: one thread per (node, 16-byte vector) - VEC = Vec16<T>::N - or per element (VEC = 1).
template <typename T, int VEC>
__global__ void __launch_bounds__(256) k_agg_generic(AggParams p) {
  const T* __restrict__ x = static_cast<const T*>(p.x);
  const T* __restrict__ z = static_cast<const T*>(p.z);
  const T* __restrict__ addend = static_cast<const T*>(p.addend);
  T* __restrict__ out = static_cast<T*>(p.out);
  T* __restrict__ copy_to = static_cast<T*>(p.copy_to);
  const int64_t row = (int64_t)p.F * p.P;
  const int64_t vec_per_row = row / VEC;
  const int64_t total = p.N * vec_per_row;
  const bool has_z = z != nullptr, has_ns = p.nbr_scale != nullptr;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = w / vec_per_row;
    const int64_t off = (w - v * vec_per_row) * VEC;
    const int c = (int)(off / p.P);
    const int i0 = p.ptr[v], i1 = p.ptr[v + 1];
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
    for (int i = i0; i < i1; ++i) {
      const int u = p.nbr[i];
      float xv[VEC];
      if constexpr (VEC == 1) xv[0] = to_f32<T>(x[(int64_t)u * p.ldx + off]);
      else Vec16<T>::load(x + (int64_t)u * p.ldx + off, xv);
      float2 gb = make_float2(1.f, 0.f);
      if (has_z) gb = load_gb<T>(z, (int64_t)p.eid[i] * 2 * p.F + 2 * c);
      film_acc<VEC>(acc, xv, has_z, gb, p.use_beta != 0, has_ns ? p.nbr_scale[u] : 1.f, has_ns);
    }
    film_post<VEC>(acc, p.post, i1 - i0, p.post == RGNN_POST_SCALE ? p.dst_scale[v] : 1.f);
    if (addend) {
      float av[VEC];
      if constexpr (VEC == 1) av[0] = to_f32<T>(addend[v * p.ldadd + off]);
      else Vec16<T>::load(addend + v * p.ldadd + off, av);
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = __fadd_rn(acc[i], av[i]);
    }
    if constexpr (VEC == 1) out[v * p.ldout + off] = from_f32<T>(acc[0]);
    else Vec16<T>::store(out + v * p.ldout + off, acc);
    if (copy_to) {
      if constexpr (VEC == 1) copy_to[v * p.ldout + off] = x[v * p.ldx + off];
      else *reinterpret_cast<uint4*>(copy_to + v * p.ldout + off) = *reinterpret_cast<const uint4*>(x + v * p.ldx + off);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// staged: CTA = (graph, channel tile).  Dynamic shared memory:
//   [0, 16)                      mbarrier
//   [16, ...)                    int32 s_ptr[maxK+1], s_nbr[maxE], s_eid[maxE]   (padded to 16 B)
//   [tile_off, ...)              T tile[K][ct*P]
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

// Issue `n` bulk copies of `bytes` each, spread over the lanes of one warp; lane 0 arms the barrier first.
// src(i) / dst(i) are callables.  Must be called by a full warp.
template <typename FS, typename FD>
__device__ __forceinline__ void warp_bulk_rows(uint64_t* bar, int n, uint32_t bytes, FS src, FD dst) {
  const int lane = threadIdx.x & 31;
  for (int i = lane; i < n; i += 32) bulk_g2s(dst(i), src(i), bytes, bar);
}

template <typename T> __device__ __forceinline__ float2 lds_gb(const T* zt, int idx);
template <> __device__ __forceinline__ float2 lds_gb<float>(const float* zt, int idx) { return *reinterpret_cast<const float2*>(zt + idx); }
template <> __device__ __forceinline__ float2 lds_gb<__nv_bfloat16>(const __nv_bfloat16* zt, int idx) {
  const uint32_t w = *reinterpret_cast<const uint32_t*>(zt + idx);
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

// zmode: 0 = no FiLM, 1 = gamma/beta read from global, 2 = the graph's z slice [Eg][2*ct] staged in shared memory too
template <typename T>
__global__ void __launch_bounds__(256) k_agg_staged(AggParams p, int CT, int maxK, int maxE, int tile_off, int zmode) {
  constexpr int VEC = Vec16<T>::N;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  int32_t* s_ptr = reinterpret_cast<int32_t*>(smem + 16);
  int32_t* s_nbr = s_ptr + (maxK + 1);
  int32_t* s_eid = s_nbr + maxE;
  T* tile = reinterpret_cast<T*>(smem + tile_off);
  T* ztile = tile + (size_t)maxK * CT * p.P;

  const T* __restrict__ x = static_cast<const T*>(p.x);
  const T* __restrict__ z = static_cast<const T*>(p.z);
  const T* __restrict__ addend = static_cast<const T*>(p.addend);
  T* __restrict__ out = static_cast<T*>(p.out);
  T* __restrict__ copy_to = static_cast<T*>(p.copy_to);

  const int tiles_c = (p.F + CT - 1) / CT;
  const int64_t g = blockIdx.x / tiles_c;
  const int c0 = (int)(blockIdx.x % tiles_c) * CT;
  const int ct = min(CT, p.F - c0);
  const int n0 = p.noff[g], K = p.noff[g + 1] - n0;
  if (K == 0) return;
  const int slice = ct * p.P;                 // elements per node in this tile (contiguous in HBM)
  const uint32_t slice_bytes = (uint32_t)slice * sizeof(T);
  const int e0 = p.ptr[n0], eg = p.ptr[n0 + K] - e0;   // a batched graph's CSR slice == its edge-id range
  const uint32_t zrow_bytes = (uint32_t)(2 * ct) * sizeof(T);

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) mbar_expect_tx(bar, slice_bytes * (uint32_t)K + (zmode == 2 ? zrow_bytes * (uint32_t)eg : 0u));
    __syncwarp();
    warp_bulk_rows(bar, K, slice_bytes,
                   [&](int a) { return x + (int64_t)(n0 + a) * p.ldx + (int64_t)c0 * p.P; },
                   [&](int a) { return tile + (size_t)a * slice; });
    if (zmode == 2)
      warp_bulk_rows(bar, eg, zrow_bytes,
                     [&](int i) { return z + (int64_t)(e0 + i) * 2 * p.F + 2 * c0; },
                     [&](int i) { return ztile + (size_t)i * 2 * ct; });
  }
  // while the copy engine works: this graph's CSR slice -> shared
  for (int i = threadIdx.x; i <= K; i += blockDim.x) s_ptr[i] = p.ptr[n0 + i] - e0;
  for (int i = threadIdx.x; i < eg; i += blockDim.x) { s_nbr[i] = p.nbr[e0 + i] - n0; s_eid[i] = zmode ? p.eid[e0 + i] - e0 : 0; }
  __syncthreads();
  mbar_wait(bar, 0);

  const bool has_z = zmode != 0, has_ns = p.nbr_scale != nullptr, use_beta = p.use_beta != 0;
  const int nvec = slice / VEC;
  for (int j = threadIdx.x; j < nvec; j += blockDim.x) {
    const int off = j * VEC;
    const int cl = off / p.P;                  // channel inside the tile
    const int64_t goff = (int64_t)c0 * p.P + off;
    for (int a = 0; a < K; ++a) {
      const int i0 = s_ptr[a], i1 = s_ptr[a + 1];
      float acc[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
      for (int i = i0; i < i1; ++i) {
        const int u = s_nbr[i];
        float xv[VEC];
        Vec16<T>::load(tile + (size_t)u * slice + off, xv);
        float2 gb = make_float2(1.f, 0.f);
        if (zmode == 2) gb = lds_gb<T>(ztile, s_eid[i] * 2 * ct + 2 * cl);
        else if (zmode == 1) gb = load_gb<T>(z, (int64_t)(e0 + s_eid[i]) * 2 * p.F + 2 * (c0 + cl));
        film_acc<VEC>(acc, xv, has_z, gb, use_beta, has_ns ? p.nbr_scale[n0 + u] : 1.f, has_ns);
      }
      const int64_t v = n0 + a;
      film_post<VEC>(acc, p.post, i1 - i0, p.post == RGNN_POST_SCALE ? p.dst_scale[v] : 1.f);
      if (addend) {
        float av[VEC];
        Vec16<T>::load(addend + v * p.ldadd + goff, av);
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = __fadd_rn(acc[i], av[i]);
      }
      Vec16<T>::store(out + v * p.ldout + goff, acc);
      if (copy_to)
        *reinterpret_cast<uint4*>(copy_to + v * p.ldout + goff) = *reinterpret_cast<const uint4*>(tile + (size_t)a * slice + off);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// A10 fused: the whole backward of the FiLM-mean aggregation for one (graph, channel tile) from ONE staging of the
// h and dAgg tiles:   dh[u]   = addend[u] + sum_{e: src=u} gamma_e * w_dst * dAgg[dst_e]
//                     dz[e,:] = sigmoid'(z) * w_dst * ( sum_p h[src_e] * dAgg[dst_e] ,  sum_p dAgg[dst_e] )
// HBM traffic: read h, dAgg, z (+addend) once, write dh, dz once (the reference keeps three [E,F,P] activations alive).
struct FilmBwdParams {
  const void* h; int64_t ldh;
  const void* dagg; int64_t ldd;
  const void* z;
  const void* addend; int64_t ldadd;
  void* dh; int64_t lddh;
  void* dz;
  const float* inv_in_deg;
  const int32_t *in_ptr, *in_src, *in_eid, *out_ptr, *out_dst, *out_eid, *noff;
  int F, P, fuse_sig;
};

template <typename T>
__global__ void __launch_bounds__(256) k_film_bwd_staged(FilmBwdParams p, int CT, int maxK, int maxE, int tile_off) {
  constexpr int VEC = Vec16<T>::N;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  int32_t* s_iptr = reinterpret_cast<int32_t*>(smem + 16);
  int32_t* s_optr = s_iptr + (maxK + 1);
  int32_t* s_isrc = s_optr + (maxK + 1);
  int32_t* s_ieid = s_isrc + maxE;
  int32_t* s_odst = s_ieid + maxE;
  int32_t* s_oeid = s_odst + maxE;
  float* s_w = reinterpret_cast<float*>(s_oeid + maxE);
  T* htile = reinterpret_cast<T*>(smem + tile_off);
  T* dtile = htile + (size_t)maxK * CT * p.P;
  T* ztile = dtile + (size_t)maxK * CT * p.P;
  T* dztile = ztile + (size_t)maxE * 2 * CT;

  const T* __restrict__ h = static_cast<const T*>(p.h);
  const T* __restrict__ dagg = static_cast<const T*>(p.dagg);
  const T* __restrict__ z = static_cast<const T*>(p.z);
  const T* __restrict__ addend = static_cast<const T*>(p.addend);
  T* __restrict__ dh = static_cast<T*>(p.dh);
  T* __restrict__ dz = static_cast<T*>(p.dz);

  const int tiles_c = (p.F + CT - 1) / CT;
  const int64_t g = blockIdx.x / tiles_c;
  const int c0 = (int)(blockIdx.x % tiles_c) * CT;
  const int ct = min(CT, p.F - c0);
  const int n0 = p.noff[g], K = p.noff[g + 1] - n0;
  if (K == 0) return;
  const int slice = ct * p.P;
  const uint32_t slice_bytes = (uint32_t)slice * sizeof(T);
  const int e0 = p.in_ptr[n0], eg = p.in_ptr[n0 + K] - e0;
  const uint32_t zrow_bytes = (uint32_t)(2 * ct) * sizeof(T);

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) mbar_expect_tx(bar, 2u * slice_bytes * (uint32_t)K + zrow_bytes * (uint32_t)eg);
    __syncwarp();
    warp_bulk_rows(bar, K, slice_bytes,
                   [&](int a) { return h + (int64_t)(n0 + a) * p.ldh + (int64_t)c0 * p.P; },
                   [&](int a) { return htile + (size_t)a * slice; });
    warp_bulk_rows(bar, K, slice_bytes,
                   [&](int a) { return dagg + (int64_t)(n0 + a) * p.ldd + (int64_t)c0 * p.P; },
                   [&](int a) { return dtile + (size_t)a * slice; });
    warp_bulk_rows(bar, eg, zrow_bytes,
                   [&](int i) { return z + (int64_t)(e0 + i) * 2 * p.F + 2 * c0; },
                   [&](int i) { return ztile + (size_t)i * 2 * ct; });
  }
  for (int i = threadIdx.x; i <= K; i += blockDim.x) { s_iptr[i] = p.in_ptr[n0 + i] - e0; s_optr[i] = p.out_ptr[n0 + i] - e0; }
  for (int i = threadIdx.x; i < K; i += blockDim.x) s_w[i] = p.inv_in_deg[n0 + i];
  for (int i = threadIdx.x; i < eg; i += blockDim.x) {
    s_isrc[i] = p.in_src[e0 + i] - n0; s_ieid[i] = p.in_eid[e0 + i] - e0;
    s_odst[i] = p.out_dst[e0 + i] - n0; s_oeid[i] = p.out_eid[e0 + i] - e0;
  }
  __syncthreads();
  mbar_wait(bar, 0);

  const int lpc = p.P / VEC;                  // lanes sharing one channel (power of two <= 32, host-checked)
  const int nvec = slice / VEC;
  const int nvec_pad = (nvec + 31) & ~31;     // whole warps iterate together (shuffles)
  for (int j = threadIdx.x; j < nvec_pad; j += blockDim.x) {
    const bool live = j < nvec;
    const int off = live ? j * VEC : 0;
    const int cl = off / p.P;
    const int64_t goff = (int64_t)c0 * p.P + off;
    // ---- dh: transposed gather over the out-edges of every source u
    if (live) {
      for (int u = 0; u < K; ++u) {
        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
        for (int i = s_optr[u]; i < s_optr[u + 1]; ++i) {
          const int v = s_odst[i];
          float dv[VEC];
          Vec16<T>::load(dtile + (size_t)v * slice + off, dv);
          const float gw = lds_gb<T>(ztile, s_oeid[i] * 2 * ct + 2 * cl).x * s_w[v];
#pragma unroll
          for (int q = 0; q < VEC; ++q) acc[q] = fmaf(gw, dv[q], acc[q]);
        }
        const int64_t gu = n0 + u;
        if (addend) {
          float av[VEC];
          Vec16<T>::load(addend + gu * p.ldadd + goff, av);
#pragma unroll
          for (int q = 0; q < VEC; ++q) acc[q] += av[q];
        }
        Vec16<T>::store(dh + gu * p.lddh + goff, acc);
      }
    }
    // ---- dz: per destination keep dAgg in registers, dot with every in-neighbour's h; lane-group reduction over pixels
    for (int a = 0; a < K; ++a) {
      float dv[VEC];
      float sd = 0.f;
      if (live) {
        Vec16<T>::load(dtile + (size_t)a * slice + off, dv);
#pragma unroll
        for (int q = 0; q < VEC; ++q) sd += dv[q];
      } else {
#pragma unroll
        for (int q = 0; q < VEC; ++q) dv[q] = 0.f;
      }
      for (int o = lpc >> 1; o > 0; o >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);
      const float wa = s_w[a];
      for (int i = s_iptr[a]; i < s_iptr[a + 1]; ++i) {
        float sxd = 0.f;
        if (live) {
          float hv[VEC];
          Vec16<T>::load(htile + (size_t)s_isrc[i] * slice + off, hv);
#pragma unroll
          for (int q = 0; q < VEC; ++q) sxd = fmaf(hv[q], dv[q], sxd);
        }
        for (int o = lpc >> 1; o > 0; o >>= 1) sxd += __shfl_xor_sync(0xffffffffu, sxd, o);
        if (live && (j % lpc) == 0) {
          const int zi = s_ieid[i] * 2 * ct + 2 * cl;
          const float2 gb = lds_gb<T>(ztile, zi);
          float dg = sxd * wa, db = sd * wa;
          if (p.fuse_sig) { dg *= gb.x * (1.f - gb.x); db *= gb.y * (1.f - gb.y); }
          if constexpr (sizeof(T) == 4) *reinterpret_cast<float2*>(dztile + zi) = make_float2(dg, db);
          else *reinterpret_cast<__nv_bfloat162*>(dztile + zi) = __floats2bfloat162_rn(dg, db);
        }
      }
    }
  }
  __syncthreads();
  // ---- dz tile -> global, 16-byte coalesced rows
  const int zvec_row = (2 * ct) / VEC;
  for (int t = threadIdx.x; t < eg * zvec_row; t += blockDim.x) {
    const int e = t / zvec_row, q = t % zvec_row;
    *reinterpret_cast<uint4*>(dz + (int64_t)(e0 + e) * 2 * p.F + 2 * c0 + q * VEC) = *reinterpret_cast<const uint4*>(dztile + (size_t)e * 2 * ct + q * VEC);
  }
}

// ------------------------------------------------------------------------------------------------
// A10: per-edge FiLM gradients.  One warp per edge; LPC lanes share a channel (P/VEC vectors), 32/LPC channels
// in flight per warp; shuffle reduction inside the lane group; sigmoid' fused.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) k_film_edge_grad(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dagg, int64_t ldd,
                                                        const T* __restrict__ z, const float* __restrict__ inv_in_deg,
                                                        const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                                        T* __restrict__ dz, int64_t E, int F, int P, int LPC, int fuse_sig) {
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPC, grp = lane / LPC, cpw = 32 / LPC;  // channels per warp pass
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int cblocks = (F + cpw - 1) / cpw;
  for (int64_t w = warp; w < E * cblocks; w += nwarps) {
    const int64_t e = w / cblocks;
    const int c = (int)(w % cblocks) * cpw + grp;
    const int u = src[e], v = dst[e];
    float sxd = 0.f, sd = 0.f;
    if (c < F) {
      const T* xr = x + (int64_t)u * ldx + (int64_t)c * P;
      const T* dr = dagg + (int64_t)v * ldd + (int64_t)c * P;
      for (int pp = sub * VEC; pp < P; pp += LPC * VEC) {
        float xv[VEC], dv[VEC];
        if constexpr (VEC == 1) { xv[0] = to_f32<T>(xr[pp]); dv[0] = to_f32<T>(dr[pp]); }
        else { Vec16<T>::load(xr + pp, xv); Vec16<T>::load(dr + pp, dv); }
#pragma unroll
        for (int i = 0; i < VEC; ++i) { sxd = fmaf(xv[i], dv[i], sxd); sd += dv[i]; }
      }
    }
    for (int o = LPC >> 1; o > 0; o >>= 1) {
      sxd += __shfl_xor_sync(0xffffffffu, sxd, o);
      sd += __shfl_xor_sync(0xffffffffu, sd, o);
    }
    if (sub == 0 && c < F) {
      const float wgt = inv_in_deg[v];
      const int64_t zi = e * 2 * F + 2 * c;
      const float2 gb = load_gb<T>(z, zi);
      float dg = sxd * wgt, db = sd * wgt;
      if (fuse_sig) { dg *= gb.x * (1.f - gb.x); db *= gb.y * (1.f - gb.y); }
      if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float2*>(dz + zi) = make_float2(dg, db);
      } else {
        __nv_bfloat162 b = __floats2bfloat162_rn(dg, db);
        *reinterpret_cast<__nv_bfloat162*>(dz + zi) = b;
      }
    }
  }
}

}  // namespace rgnn

using namespace rgnn;

namespace rgnn {
int film_fwd_complete(int dtype, int K, const void* x, int64_t ldx, const void* z, void* out, int64_t ldout, void* copy_to,
                      int64_t G, int64_t F, int64_t P, cudaStream_t st);
int film_bwd_complete(int dtype, int K, const void* h, int64_t ldh, const void* dagg, int64_t ldd, const void* z, const void* addend,
                      int64_t ldadd, void* dh, int64_t lddh, void* dz, int64_t G, int64_t F, int64_t P, int fuse_sig, cudaStream_t st);
int film_fwd_dense(int dtype, int K, const void* x, int64_t ldx, const void* z, void* out, int64_t ldout, void* copy_to,
                   int64_t G, int64_t F, int64_t P, cudaStream_t st);
int film_bwd_dense(int dtype, int K, const void* h, int64_t ldh, const void* dagg, int64_t ldd, const void* z, const void* addend,
                   int64_t ldadd, void* dh, int64_t lddh, void* dz, int64_t G, int64_t F, int64_t P, int fuse_sig, cudaStream_t st);
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" int rgnn_aggregate(int dtype, const void* x, int64_t ldx, const void* z, int use_beta,
                              const float* nbr_scale, int post, const float* dst_scale,
                              const void* addend, int64_t ldadd, void* out, int64_t ldout, void* copy_x_to,
                              const int32_t* ptr32, const int32_t* nbr32, const int32_t* eid32,
                              const int32_t* noff32, int64_t G, int64_t max_graph_nodes, int64_t max_graph_edges, int64_t complete_k,
                              int64_t N, int64_t E, int64_t F, int64_t P, void* stream) {
  RGNN_REQUIRE(dtype == RGNN_F32 || dtype == RGNN_BF16, RGNN_E_ARG, "rgnn_aggregate: bad dtype %d", dtype);
  RGNN_REQUIRE(N >= 0 && E >= 0 && F > 0 && P > 0, RGNN_E_ARG, "rgnn_aggregate: bad sizes N=%lld E=%lld F=%lld P=%lld",
               (long long)N, (long long)E, (long long)F, (long long)P);
  RGNN_REQUIRE(post >= 0 && post <= 2, RGNN_E_ARG, "rgnn_aggregate: bad post mode %d", post);
  if (N == 0) return 0;
  RGNN_REQUIRE(x && out && ptr32, RGNN_E_ARG, "rgnn_aggregate: null x/out/ptr");
  RGNN_REQUIRE(E == 0 || (nbr32 && (eid32 || !z)), RGNN_E_ARG, "rgnn_aggregate: null CSR arrays");
  RGNN_REQUIRE(post != RGNN_POST_SCALE || dst_scale, RGNN_E_ARG, "rgnn_aggregate: POST_SCALE needs dst_scale");
  RGNN_REQUIRE(F * P < (1ll << 31) && F < (1 << 30), RGNN_E_RANGE, "rgnn_aggregate: row too large");
  const int es = dtype == RGNN_F32 ? 4 : 2;
  const int vec = 16 / es;
  AggParams p;
  p.x = x; p.ldx = ldx; p.z = z; p.use_beta = use_beta; p.nbr_scale = nbr_scale; p.post = post; p.dst_scale = dst_scale;
  p.addend = addend; p.ldadd = ldadd; p.out = out; p.ldout = ldout; p.copy_to = copy_x_to;
  p.ptr = ptr32; p.nbr = nbr32; p.eid = eid32; p.noff = noff32; p.G = G; p.N = N; p.E = E; p.F = (int)F; p.P = (int)P;
  cudaStream_t st = as_stream(stream);
  const bool vec_ok = (P % vec == 0) && aligned16(x) && aligned16(out) && (ldx % vec == 0) && (ldout % vec == 0) &&
                      (!addend || (aligned16(addend) && ldadd % vec == 0)) && (!copy_x_to || aligned16(copy_x_to)) &&
                      (!z || (reinterpret_cast<uintptr_t>(z) & 7u) == 0);
  if (!vec_ok) {
    RGNN_REQUIRE(!z || (reinterpret_cast<uintptr_t>(z) & (2 * es - 1)) == 0, RGNN_E_ALIGN, "rgnn_aggregate: z not pair-aligned");
    const int64_t total = N * F * P;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 32 * num_sms()) blocks = 32 * num_sms();
    if (dtype == RGNN_F32) k_agg_generic<float, 1><<<(unsigned)blocks, 256, 0, st>>>(p);
    else k_agg_generic<__nv_bfloat16, 1><<<(unsigned)blocks, 256, 0, st>>>(p);
    RGNN_LAUNCH_CHECK("k_agg_generic<1>");
    return 0;
  }
  // ---- complete K-robot graphs (the reference's topology): register-tiled kernel, no CSR walk
  if (complete_k >= 2 && complete_k <= 8 && post == RGNN_POST_MEAN && !nbr_scale && !addend && N == G * complete_k &&
      E == G * complete_k * (complete_k - 1) && (!z || (reinterpret_cast<uintptr_t>(z) & (2 * es - 1)) == 0) && use_beta) {
    return film_fwd_complete(dtype, (int)complete_k, x, ldx, z, out, ldout, copy_x_to, G, F, P, st);
  }
  // ---- complete graphs with more robots than the register kernels hold: shared-memory staged, register-tiled over destinations
  if (complete_k > 8 && complete_k < 4096 && vec_ok && post == RGNN_POST_MEAN && !nbr_scale && !addend && N == G * complete_k &&
      E == G * complete_k * (complete_k - 1) && use_beta) {
    const int rc = film_fwd_dense(dtype, (int)complete_k, x, ldx, z, out, ldout, copy_x_to, G, F, P, st);
    if (rc != RGNN_E_UNSUPPORTED) return rc;
  }
  // ---- staged path: needs the batch layout and every graph's tile + CSR slice to fit shared memory
  const int maxK = (int)max_graph_nodes, maxE = (int)max_graph_edges;
  if (noff32 && G > 0 && maxK > 0 && max_graph_nodes < 4096 && max_graph_edges < (1 << 20)) {
    const int64_t row_bytes_per_ch = P * es;                 // bytes of one channel of one node
    const int idx_bytes = (int)(((maxK + 1 + 2 * (int64_t)maxE) * 4 + 15) / 16 * 16);
    const int tile_off = ((16 + idx_bytes) + 127) / 128 * 128;
    const int64_t budget = 44 * 1024 - tile_off;           // ~5 CTAs / SM
    // gamma/beta slice in shared memory as well when its rows are 16-byte addressable
    const int zalign = 8 / es;                               // CT multiple making 2*CT*es a 16-byte multiple
    const bool z_stage_ok = z && aligned16(z) && (2 * F * es) % 16 == 0 && F % zalign == 0;
    int zmode = z ? (z_stage_ok ? 2 : 1) : 0;
    int64_t CT = 0;
    for (int attempt = 0; attempt < 2 && CT < 1; ++attempt) {
      const int64_t per_ch = maxK * row_bytes_per_ch + (zmode == 2 ? (int64_t)maxE * 2 * es : 0);
      CT = budget > 0 ? budget / per_ch : 0;
      if (CT > F) CT = F;
      if (zmode == 2) CT -= CT % zalign;
      // bulk copies need 16-byte multiples: ct*P*es % 16 == 0 for every tile incl. the tail
      while (CT > 0 && ((CT * row_bytes_per_ch) % 16 != 0 || ((F % CT) * row_bytes_per_ch) % 16 != 0 ||
                        (zmode == 2 && (F % CT) % zalign != 0))) CT -= (zmode == 2 ? zalign : 1);
      if (CT < 1 && zmode == 2) zmode = 1;                   // retry without staging z
    }
    if (CT >= 1 && (int64_t)CT * P >= vec) {
      const int64_t tiles_c = (F + CT - 1) / CT;
      const int64_t blocks = G * tiles_c;
      if (blocks < (1ll << 31)) {
        const size_t smem = (size_t)tile_off + (size_t)maxK * CT * row_bytes_per_ch + (zmode == 2 ? (size_t)maxE * 2 * CT * es : 0);
        if (dtype == RGNN_F32) {
          k_agg_staged<float><<<(unsigned)blocks, 256, smem, st>>>(p, (int)CT, maxK, maxE, tile_off, zmode);
        } else {
          k_agg_staged<__nv_bfloat16><<<(unsigned)blocks, 256, smem, st>>>(p, (int)CT, maxK, maxE, tile_off, zmode);
        }
        RGNN_LAUNCH_CHECK("k_agg_staged");
        return 0;
      }
    }
  }
  const int64_t total = N * (F * P / vec);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 32 * num_sms()) blocks = 32 * num_sms();
  if (dtype == RGNN_F32) k_agg_generic<float, 4><<<(unsigned)blocks, 256, 0, st>>>(p);
  else k_agg_generic<__nv_bfloat16, 8><<<(unsigned)blocks, 256, 0, st>>>(p);
  RGNN_LAUNCH_CHECK("k_agg_generic");
  return 0;
}

extern "C" int rgnn_film_edge_grad(int dtype, const void* x, int64_t ldx, const void* dagg, int64_t ldd,
                                   const void* z, int fuse_sigmoid_grad, const float* inv_in_deg, const int32_t* src32, const int32_t* dst32,
                                   void* dz, int64_t E, int64_t F, int64_t P, void* stream) {
  RGNN_REQUIRE(dtype == RGNN_F32 || dtype == RGNN_BF16, RGNN_E_ARG, "rgnn_film_edge_grad: bad dtype %d", dtype);
  RGNN_REQUIRE(E >= 0 && F > 0 && P > 0, RGNN_E_ARG, "rgnn_film_edge_grad: bad sizes");
  if (E == 0) return 0;
  RGNN_REQUIRE(x && dagg && z && inv_in_deg && src32 && dst32 && dz, RGNN_E_ARG, "rgnn_film_edge_grad: null pointer");
  RGNN_REQUIRE(F * P < (1ll << 31), RGNN_E_RANGE, "rgnn_film_edge_grad: row too large");
  const int es = dtype == RGNN_F32 ? 4 : 2;
  RGNN_REQUIRE((reinterpret_cast<uintptr_t>(z) & (2 * es - 1)) == 0 && (reinterpret_cast<uintptr_t>(dz) & (2 * es - 1)) == 0,
               RGNN_E_ALIGN, "rgnn_film_edge_grad: z/dz not pair-aligned");
  const int vec = 16 / es;
  const bool vec_ok = (P % vec == 0) && aligned16(x) && aligned16(dagg) && (ldx % vec == 0) && (ldd % vec == 0);
  const int64_t per = vec_ok ? (P / vec) : P;
  int lpc = 1;
  while (lpc < 32 && lpc < per) lpc <<= 1;
  const int cpw = 32 / lpc;
  const int64_t warps = E * ((F + cpw - 1) / cpw);
  int64_t blocks = (warps + 7) / 8;
  if (blocks > 64 * num_sms()) blocks = 64 * num_sms();
  cudaStream_t st = as_stream(stream);
#define LAUNCH(T, V)                                                                                              \
  k_film_edge_grad<T, V><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const T*>(x), ldx, static_cast<const T*>(dagg), ldd, \
                                                           static_cast<const T*>(z), inv_in_deg, src32, dst32,    \
                                                           static_cast<T*>(dz), E, (int)F, (int)P, lpc, fuse_sigmoid_grad)
  if (dtype == RGNN_F32) { if (vec_ok) LAUNCH(float, 4); else LAUNCH(float, 1); }
  else { if (vec_ok) LAUNCH(__nv_bfloat16, 8); else LAUNCH(__nv_bfloat16, 1); }
#undef LAUNCH
  RGNN_LAUNCH_CHECK("k_film_edge_grad");
  return 0;
}

// Fused backward of the FiLM-mean aggregation (A10) for batched robot graphs.  Returns RGNN_E_UNSUPPORTED when the
// batch does not fit the staged kernel (caller then uses rgnn_aggregate on the out-CSR + rgnn_film_edge_grad).
extern "C" int rgnn_film_backward(int dtype, const void* h, int64_t ldh, const void* dagg, int64_t lddagg, const void* z,
                                  int fuse_sigmoid_grad, const void* addend, int64_t ldadd, void* dh, int64_t lddh, void* dz,
                                  const float* inv_in_deg, const int32_t* ptr32, const int32_t* idx32, const int32_t* noff32,
                                  int64_t G, int64_t max_graph_nodes, int64_t max_graph_edges, int64_t complete_k,
                                  int64_t N, int64_t E, int64_t F, int64_t P, void* stream) {
  RGNN_REQUIRE(dtype == RGNN_F32 || dtype == RGNN_BF16, RGNN_E_ARG, "rgnn_film_backward: bad dtype %d", dtype);
  RGNN_REQUIRE(N >= 0 && E >= 0 && G >= 0 && F > 0 && P > 0, RGNN_E_ARG, "rgnn_film_backward: bad sizes");
  if (N == 0) return 0;
  RGNN_REQUIRE(h && dagg && z && dh && dz && inv_in_deg && ptr32 && idx32 && noff32, RGNN_E_ARG, "rgnn_film_backward: null pointer");
  const int es = dtype == RGNN_F32 ? 4 : 2;
  const int vec = 16 / es;
  const int maxK = (int)max_graph_nodes, maxE = (int)max_graph_edges;
  const int64_t lpc = P / vec;
  const bool ok = G > 0 && maxK > 0 && maxK < 4096 && maxE < (1 << 20) && P % vec == 0 && lpc >= 1 && lpc <= 32 && (lpc & (lpc - 1)) == 0 &&
                  aligned16(h) && aligned16(dagg) && aligned16(z) && aligned16(dh) && aligned16(dz) && (!addend || aligned16(addend)) &&
                  ldh % vec == 0 && lddagg % vec == 0 && lddh % vec == 0 && (!addend || ldadd % vec == 0) &&
                  (2 * F * es) % 16 == 0 && F % (8 / es) == 0 && F * P < (1ll << 31);
  RGNN_REQUIRE(ok, RGNN_E_UNSUPPORTED, "rgnn_film_backward: shape/alignment not covered by the staged kernel");
  if (complete_k >= 2 && complete_k <= 8 && N == G * complete_k && E == G * complete_k * (complete_k - 1)) {
    const int rc = film_bwd_complete(dtype, (int)complete_k, h, ldh, dagg, lddagg, z, addend, ldadd, dh, lddh, dz, G, F, P,
                                     fuse_sigmoid_grad, as_stream(stream));
    if (rc != RGNN_E_UNSUPPORTED) return rc;
  }
  if (complete_k > 8 && complete_k < 4096 && N == G * complete_k && E == G * complete_k * (complete_k - 1)) {
    const int rc = film_bwd_dense(dtype, (int)complete_k, h, ldh, dagg, lddagg, z, addend, ldadd, dh, lddh, dz, G, F, P,
                                  fuse_sigmoid_grad, as_stream(stream));
    if (rc != RGNN_E_UNSUPPORTED) return rc;
  }
  const int zalign = 8 / es;
  const int64_t row_bytes_per_ch = P * es;
  const int idx_bytes = (int)(((2 * (maxK + 1) + 4 * (int64_t)maxE + maxK) * 4 + 15) / 16 * 16);
  const int tile_off = ((16 + idx_bytes) + 127) / 128 * 128;
  const int64_t budget = 56 * 1024 - tile_off;             // 4 CTAs / SM
  const int64_t per_ch = 2 * maxK * row_bytes_per_ch + (int64_t)maxE * 4 * es;
  int64_t CT = budget > 0 ? budget / per_ch : 0;
  if (CT > F) CT = F;
  CT -= CT % zalign;
  while (CT > 0 && ((CT * row_bytes_per_ch) % 16 != 0 || ((F % CT) * row_bytes_per_ch) % 16 != 0 || (F % CT) % zalign != 0)) CT -= zalign;
  RGNN_REQUIRE(CT >= 1, RGNN_E_UNSUPPORTED, "rgnn_film_backward: a graph's tile does not fit shared memory");
  const int64_t tiles_c = (F + CT - 1) / CT;
  const int64_t blocks = G * tiles_c;
  RGNN_REQUIRE(blocks < (1ll << 31), RGNN_E_RANGE, "rgnn_film_backward: grid too large");
  FilmBwdParams p;
  p.h = h; p.ldh = ldh; p.dagg = dagg; p.ldd = lddagg; p.z = z; p.addend = addend; p.ldadd = ldadd; p.dh = dh; p.lddh = lddh; p.dz = dz;
  p.inv_in_deg = inv_in_deg;
  p.in_ptr = ptr32; p.out_ptr = ptr32 + (N + 1);
  p.in_src = idx32; p.in_eid = idx32 + E; p.out_dst = idx32 + 2 * E; p.out_eid = idx32 + 3 * E;
  p.noff = noff32; p.F = (int)F; p.P = (int)P; p.fuse_sig = fuse_sigmoid_grad;
  const size_t smem = (size_t)tile_off + 2 * (size_t)maxK * CT * row_bytes_per_ch + (size_t)maxE * 4 * CT * es;
  cudaStream_t st = as_stream(stream);
  static DeviceOnce once[2];
  if (once[dtype].need()) {
    cudaError_t e = dtype == RGNN_F32
        ? cudaFuncSetAttribute(k_film_bwd_staged<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)
        : cudaFuncSetAttribute(k_film_bwd_staged<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_film_bwd_staged)");
    once[dtype].mark();
  }
  if (dtype == RGNN_F32) k_film_bwd_staged<float><<<(unsigned)blocks, 256, smem, st>>>(p, (int)CT, maxK, maxE, tile_off);
  else k_film_bwd_staged<__nv_bfloat16><<<(unsigned)blocks, 256, smem, st>>>(p, (int)CT, maxK, maxE, tile_off);
  RGNN_LAUNCH_CHECK("k_film_bwd_staged");
  return 0;
}
