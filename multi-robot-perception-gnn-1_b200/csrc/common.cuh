// Shared helpers for librobotgnn (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/robot_gnn.h"

namespace rgnn {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define RGNN_REQUIRE(cond, code, ...)            \
  do {                                           \
    if (!(cond)) {                               \
      rgnn::set_error(__VA_ARGS__);              \
      return (code);                             \
    }                                            \
  } while (0)

void count_launch();

#define RGNN_LAUNCH_CHECK(what)                              \
  do {                                                       \
    rgnn::count_launch();                                    \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return rgnn::cuda_fail(e__, what); \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// SM count of the CURRENT device (148 on B200), cached per device: grids are sized in multiples of it.
int num_sms();

// One-time per-DEVICE setup guard.  cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and occupancy answers belong to a device /
// context, and autograd runs backward on one thread per device: a process that drives several GPUs must opt in on each of them
// (a process-wide `static bool` would skip the second GPU and every launch above 48 KB would fail there).  Bit d = device d done.
struct DeviceOnce {
  std::atomic<unsigned long long> mask{0};
  bool need() const {
    int d = 0;
    cudaGetDevice(&d);
    return d < 0 || d >= 64 || !((mask.load(std::memory_order_acquire) >> d) & 1ull);
  }
  void mark() {
    int d = 0;
    cudaGetDevice(&d);
    if (d >= 0 && d < 64) mask.fetch_or(1ull << d, std::memory_order_release);
  }
};

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 16-byte vector of T (4 floats or 8 bf16) <-> fp32 registers
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void unpack(const uint4& t, float (&v)[4]) {
    v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
  }
  __device__ static __forceinline__ void load(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ static __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void unpack(const uint4& t, float (&v)[8]) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&b);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// Blackwell packed fp32: one FFMA2 / FADD2 / FMUL2 instruction does two IEEE fp32 operations (each rounded on its own, so results
// equal the scalar forms); regular 3-register FFMA issues every other cycle per SM sub-partition, the packed form restores full rate.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{ .reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n mov.b64 rc, {%6, %7};\n fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0, %1}, rd; }"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{ .reg .b64 ra, rb, rd;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n add.rn.f32x2 rd, ra, rb;\n mov.b64 {%0, %1}, rd; }"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("{ .reg .b64 ra, rb, rd;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n mul.rn.f32x2 rd, ra, rb;\n mov.b64 {%0, %1}, rd; }"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace rgnn
