// Helpers shared by the complete-graph FiLM kernels (film_complete.cu, film_ring.cu).
#pragma once
#include "common.cuh"

namespace rgnn {

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

__host__ __device__ constexpr int edge_id(int K, int u, int a) { return u * (K - 1) + (a < u ? a : a - 1); }

// Reduce-scatter of NV per-lane partial sums over the LPC lanes that share a channel: after log2(LPC) steps lane l holds the
// fully reduced values with ids  i + sum_s bit_s(l) * (NVP >> (s+1)),  i < NVP/LPC.  NVP/2 + NVP/4 + ... shuffles instead of
// NV*log2(LPC), every step's shuffles independent of each other (latency hidden by ILP, not by occupancy).
template <int NVP, int LPC>
__device__ __forceinline__ void reduce_scatter(float (&a)[NVP], int l) {
  static_assert(NVP % LPC == 0, "pad the value count to a multiple of the lane group");
  if constexpr (LPC > 1) {
    constexpr int H = NVP / 2;
    const bool up = (l & (LPC / 2)) != 0;
    float b[H];
#pragma unroll
    for (int i = 0; i < H; ++i) {
      const float send = up ? a[i] : a[i + H];
      const float keep = up ? a[i + H] : a[i];
      b[i] = keep + __shfl_xor_sync(0xffffffffu, send, LPC / 2);
    }
    reduce_scatter<H, LPC / 2>(b, l);
#pragma unroll
    for (int i = 0; i < H / (LPC / 2); ++i) a[i] = b[i];
  }
}
template <int NVP, int LPC> __device__ __forceinline__ int scattered_id(int i, int l) {
  int id = i, h = NVP / 2;
#pragma unroll
  for (int o = LPC / 2; o > 0; o >>= 1) { if (l & o) id += h; h >>= 1; }
  return id;
}

// bulk-copy engine (TMA without a tensor map) + mbarrier
namespace tma {
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
}  // namespace tma

}  // namespace rgnn
