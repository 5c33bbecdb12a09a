// common.cuh — shared device helpers for libadd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <mutex>
#include "../../include/add_b200.h"

#define ADD_CHECK_ARG(cond) do { if (!(cond)) return ADD_ERR_BAD_ARG; } while (0)
#define ADD_CHECK_SUP(cond) do { if (!(cond)) return ADD_ERR_UNSUPPORTED; } while (0)
// last CUDA runtime/driver error seen by this thread's launches (api.cu); add_last_cuda_error() reports it
extern thread_local int g_add_last_cuda_error;
extern int g_add_grid_pct;  // api.cu: persistent-grid scale (tuning)
extern int g_add_conv_grid_pct;   // api.cu: same, small convs of the persistent conv kernel only
extern int g_add_pdl;      // api.cu: 1 = launch the tcgen05 kernels with programmatic stream serialization
#define ADD_RETURN_LAUNCH() do { cudaError_t e_ = cudaGetLastError(); if (e_ == cudaSuccess) return ADD_OK; \
    g_add_last_cuda_error = (int)e_; return ADD_ERR_CUDA; } while (0)

typedef __nv_bfloat16 bf16;

// cudaFuncSetAttribute (dynamic shared memory size, carve-out) is a PER-DEVICE setting: a process that drives several
// GPUs (the reference's nn.DataParallel) must apply it on each device it launches on, not once per process.
struct PerDeviceOnce { std::mutex mu; bool done[64] = {}; };
template <class F> static inline void once_per_device(PerDeviceOnce& o, F fn) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) { fn(); return; }
  std::lock_guard<std::mutex> g(o.mu);
  if (!o.done[dev]) { fn(); o.done[dev] = true; }
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- element access, fp32 <-> storage dtype ------------------------------------------------
__device__ __forceinline__ float ld1(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld1(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// 4 consecutive channels (pointer must be 16 B aligned for float, 8 B for bf16)
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
  uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}
__device__ __forceinline__ float4 relu4(float4 v) {
  return make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
}

static inline size_t dtype_size(int dt) { return dt == ADD_BF16 ? 2 : 4; }

static inline bool tensor_ok(const add_tensor_t* t) {
  return t && t->ptr && t->n > 0 && t->h > 0 && t->w > 0 && t->c > 0 && t->pix_stride >= t->c &&
         (t->dtype == ADD_F32 || t->dtype == ADD_BF16);
}
// 4-channel vector access needs c, pix_stride and the base pointer aligned to 4 elements.
static inline bool tensor_vec4_ok(const add_tensor_t* t) {
  size_t a = 4 * dtype_size(t->dtype);
  return (t->c % 4 == 0) && (t->pix_stride % 4 == 0) && ((uintptr_t)t->ptr % a == 0);
}

// Output extent `out` is acceptable for a conv over `in` pixels when the last output's window
// [(out-1)*stride - pad, ... + dil*(k-1)] starts less than one stride past the image (the reference's
// FactorizedReduce pads one zero row/column on the far side, operations.py:98) and ends at or after 0.
static inline bool conv_extent_ok(int in, int out, int k, int stride, int pad, int dil) {
  if (out <= 0) return false;
  long long first = (long long)(out - 1) * stride - pad;
  return first < (long long)in + stride && first + (long long)dil * (k - 1) >= 0;
}

// PyTorch's area_pixel_compute_source_index (align_corners=False, bilinear) in fp32.
__device__ __forceinline__ void bilinear_src(int dst, float scale, int in_size, int& i0, int& i1,
                                             float& l0, float& l1) {
  // one fused multiply-add, like the reference's own builds: ATen's CPU kernels are compiled with FP
  // contraction on (and nvcc contracts the CUDA ones), so `scale*(dst+0.5)-0.5` is a single FMA there.
  float src = fmaf(scale, static_cast<float>(dst) + 0.5f, -0.5f);
  src = src < 0.f ? 0.f : src;
  i0 = static_cast<int>(src);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = (i1 == i0) ? 0.f : src - static_cast<float>(i0);   // clamped edge / 1-pixel source: exact copy
  l0 = 1.f - l1;
}
