// pool.cu — the non-convolutional OPS primitives (operations.py:7-11): avg_pool_3x3 (count_include_pad=False),
// max_pool_3x3, skip_connect (Identity) and none (Zero: x[::s, ::s] * 0).  NHWC, 4-channel vectors, HBM-bound;
// ACCUMULATE makes them usable as cell edges (node sum as accumulate-into-slice, ADD.py:108).
#include "common.cuh"

namespace {

// mode 0 = average over the in-image taps, 1 = max (out-of-image taps are -inf)
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
pool3x3_kernel(const TI* __restrict__ x, TO* __restrict__ y, int H, int W, int C, int xs, int Ho, int Wo, int ys,
               int stride, int mode, uint32_t flags) {
  const unsigned cv = (unsigned)C >> 2, row = (unsigned)Wo * cv, total = (unsigned)Ho * row;
  const int n = blockIdx.y;
  const TI* xn = x + (size_t)n * H * W * xs;
  TO* yn = y + (size_t)n * Ho * Wo * ys;
  for (unsigned idx = blockIdx.x * 256u + threadIdx.x; idx < total; idx += gridDim.x * 256u) {
    const unsigned oy = idx / row, rem = idx - oy * row, ox = rem / cv, c = (rem - ox * cv) * 4;
    float4 acc = mode ? make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY) : make_float4(0.f, 0.f, 0.f, 0.f);
    int cnt = 0;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = (int)oy * stride - 1 + ky;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = (int)ox * stride - 1 + kx;
        if (ix < 0 || ix >= W) continue;
        float4 v = ld4(xn + ((size_t)iy * W + ix) * xs + c);
        if (flags & ADD_RELU_IN) v = relu4(v);
        if (mode) { acc.x = fmaxf(acc.x, v.x); acc.y = fmaxf(acc.y, v.y); acc.z = fmaxf(acc.z, v.z); acc.w = fmaxf(acc.w, v.w); }
        else { acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        ++cnt;
      }
    }
    if (!mode) { const float f = (float)cnt; acc.x /= f; acc.y /= f; acc.z /= f; acc.w /= f; }   // divisor = taps inside the image
    TO* dst = yn + ((size_t)oy * Wo + ox) * ys + c;
    if (flags & ADD_ACCUMULATE) { float4 o = ld4(const_cast<const TO*>(dst)); acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
    if (flags & ADD_RELU_OUT) acc = relu4(acc);
    st4(dst, acc);
  }
}

// y (+)= scale * x[:, ::stride, ::stride, :]   (Identity: scale 1, stride 1; Zero: scale 0 — IEEE x*0, like x.mul(0.))
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
scale_kernel(const TI* __restrict__ x, TO* __restrict__ y, int H, int W, int C, int xs, int Ho, int Wo, int ys,
             int stride, float scale, uint32_t flags) {
  const unsigned cv = (unsigned)C >> 2, row = (unsigned)Wo * cv, total = (unsigned)Ho * row;
  const int n = blockIdx.y;
  const TI* xn = x + (size_t)n * H * W * xs;
  TO* yn = y + (size_t)n * Ho * Wo * ys;
  for (unsigned idx = blockIdx.x * 256u + threadIdx.x; idx < total; idx += gridDim.x * 256u) {
    const unsigned oy = idx / row, rem = idx - oy * row, ox = rem / cv, c = (rem - ox * cv) * 4;
    float4 v = ld4(xn + ((size_t)(oy * stride) * W + ox * stride) * xs + c);
    if (flags & ADD_RELU_IN) v = relu4(v);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    TO* dst = yn + ((size_t)oy * Wo + ox) * ys + c;
    if (flags & ADD_ACCUMULATE) { float4 o = ld4(const_cast<const TO*>(dst)); v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
    if (flags & ADD_RELU_OUT) v = relu4(v);
    st4(dst, v);
  }
}

// Stand-alone depthwise KxK (stride 1, pad K/2), fp32 accumulate: the first half of a SepConv half when the channel
// count is wider than the fused tensor-core kernel takes (C > 256: BASELINE config 5, C = 320 / 640); the pointwise
// 1x1 + BN then runs as a tcgen05 conv over this kernel's bf16 output — the same bf16 rounding point as the fused kernel.
// Thread = one output pixel x 4 channels; the K*K neighbouring loads hit L1/L2 (memory-bound, small deep-stride maps).
template <typename TI, typename TO, int K>
__global__ void __launch_bounds__(256)
depthwise_kernel(const TI* __restrict__ x, TO* __restrict__ y, const float* __restrict__ w, int H, int W, int C, int xs, int ys,
                 uint32_t flags) {
  const unsigned cv = (unsigned)C >> 2, row = (unsigned)W * cv, total = (unsigned)H * row;
  const int n = blockIdx.y;
  const TI* xn = x + (size_t)n * H * W * xs;
  TO* yn = y + (size_t)n * H * W * ys;
  for (unsigned idx = blockIdx.x * 256u + threadIdx.x; idx < total; idx += gridDim.x * 256u) {
    const unsigned oy = idx / row, rem = idx - oy * row, ox = rem / cv, c = (rem - ox * cv) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      const int iy = (int)oy - K / 2 + ky;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const int ix = (int)ox - K / 2 + kx;
        if (ix < 0 || ix >= W) continue;
        float4 v = ld4(xn + ((size_t)iy * W + ix) * xs + c);
        if (flags & ADD_RELU_IN) v = relu4(v);
        const float4 wk = __ldg(reinterpret_cast<const float4*>(w + (size_t)(ky * K + kx) * C + c));
        acc.x = fmaf(wk.x, v.x, acc.x); acc.y = fmaf(wk.y, v.y, acc.y); acc.z = fmaf(wk.z, v.z, acc.z); acc.w = fmaf(wk.w, v.w, acc.w);
      }
    }
    if (flags & ADD_RELU_OUT) acc = relu4(acc);
    st4(yn + ((size_t)oy * W + ox) * ys + c, acc);
  }
}

inline dim3 ew_grid(const add_tensor_t* y) {
  long long total = (long long)y->h * y->w * (y->c / 4);
  long long bx = (total + 255) / 256;
  const long long cap = (148ll * 32 + y->n - 1) / y->n;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return dim3((unsigned)bx, (unsigned)y->n);
}

}  // namespace

extern "C" int add_pool3x3_fwd(const add_tensor_t* x, const add_tensor_t* y, int mode, int stride, uint32_t flags,
                               void* stream) {
  ADD_CHECK_ARG(tensor_ok(x) && tensor_ok(y) && (mode == 0 || mode == 1) && stride >= 1);
  ADD_CHECK_ARG(x->n == y->n && x->c == y->c && y->h == (x->h - 1) / stride + 1 && y->w == (x->w - 1) / stride + 1);
  ADD_CHECK_SUP(tensor_vec4_ok(x) && tensor_vec4_ok(y) && (long long)y->h * y->w * (y->c / 4) < (1ll << 31) && y->n < 65536);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  dim3 grid = ew_grid(y);
#define PL(TI, TO) pool3x3_kernel<TI, TO><<<grid, 256, 0, s>>>((const TI*)x->ptr, (TO*)y->ptr, x->h, x->w, x->c, \
    x->pix_stride, y->h, y->w, y->pix_stride, stride, mode, flags)
  if (x->dtype == ADD_F32 && y->dtype == ADD_F32) PL(float, float);
  else if (x->dtype == ADD_BF16 && y->dtype == ADD_BF16) PL(bf16, bf16);
  else if (x->dtype == ADD_F32 && y->dtype == ADD_BF16) PL(float, bf16);
  else if (x->dtype == ADD_BF16 && y->dtype == ADD_F32) PL(bf16, float);
  else return ADD_ERR_UNSUPPORTED;
#undef PL
  ADD_RETURN_LAUNCH();
}

extern "C" int add_scale_fwd(const add_tensor_t* x, const add_tensor_t* y, float scale, int stride, uint32_t flags,
                             void* stream) {
  ADD_CHECK_ARG(tensor_ok(x) && tensor_ok(y) && stride >= 1);
  ADD_CHECK_ARG(x->n == y->n && x->c == y->c && y->h == (x->h - 1) / stride + 1 && y->w == (x->w - 1) / stride + 1);
  ADD_CHECK_SUP(tensor_vec4_ok(x) && tensor_vec4_ok(y) && (long long)y->h * y->w * (y->c / 4) < (1ll << 31) && y->n < 65536);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  dim3 grid = ew_grid(y);
#define SC(TI, TO) scale_kernel<TI, TO><<<grid, 256, 0, s>>>((const TI*)x->ptr, (TO*)y->ptr, x->h, x->w, x->c, \
    x->pix_stride, y->h, y->w, y->pix_stride, stride, scale, flags)
  if (x->dtype == ADD_F32 && y->dtype == ADD_F32) SC(float, float);
  else if (x->dtype == ADD_BF16 && y->dtype == ADD_BF16) SC(bf16, bf16);
  else if (x->dtype == ADD_F32 && y->dtype == ADD_BF16) SC(float, bf16);
  else if (x->dtype == ADD_BF16 && y->dtype == ADD_F32) SC(bf16, float);
  else return ADD_ERR_UNSUPPORTED;
#undef SC
  ADD_RETURN_LAUNCH();
}

extern "C" int add_depthwise_fwd(const add_tensor_t* x, const add_tensor_t* y, const float* w_dw, int k, uint32_t flags,
                                 void* stream) {
  ADD_CHECK_ARG(tensor_ok(x) && tensor_ok(y) && w_dw && (k == 3 || k == 5));
  ADD_CHECK_ARG(x->n == y->n && x->c == y->c && x->h == y->h && x->w == y->w && !(flags & ADD_ACCUMULATE));
  ADD_CHECK_SUP(tensor_vec4_ok(x) && tensor_vec4_ok(y) && ((uintptr_t)w_dw % 16) == 0 &&
                (long long)y->h * y->w * (y->c / 4) < (1ll << 31) && y->n < 65536);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  dim3 grid = ew_grid(y);
#define DW(TI, TO, K_) depthwise_kernel<TI, TO, K_><<<grid, 256, 0, s>>>((const TI*)x->ptr, (TO*)y->ptr, w_dw, x->h, x->w, x->c, \
    x->pix_stride, y->pix_stride, flags)
  if (x->dtype == ADD_BF16 && y->dtype == ADD_BF16) { if (k == 3) DW(bf16, bf16, 3); else DW(bf16, bf16, 5); }
  else if (x->dtype == ADD_F32 && y->dtype == ADD_F32) { if (k == 3) DW(float, float, 3); else DW(float, float, 5); }
  else return ADD_ERR_UNSUPPORTED;
#undef DW
  ADD_RETURN_LAUNCH();
}
