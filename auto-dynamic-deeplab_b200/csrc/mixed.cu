// mixed.cu — the supernet edge (reference modeling/cell_level_search.py:10-29 MixedOp:
// sum_k w_k * op_k(x) over the eight PRIMITIVES) and the architecture-weight softmax (model_net_search.py:294-310).
//
//   add_mixed_light_fwd     ONE kernel for the four parameter-free primitives of an edge and their weights:
//                           y (+)= w_skip * x + w_max * BN(maxpool3x3(x)) + w_avg * BN(avgpool3x3(x)) + w_none * (x * 0),
//                           x read once (a 3x3 neighbourhood per pixel, from L1/L2), BN(affine=False) as per-channel
//                           (mean, inv_std) — eval-mode running statistics or the batch statistics of a training step.
//                           The four conv primitives (sep / dil convs) add w_k * op_k(x) in their own epilogues.
//   add_weighted_sum_fwd    y = sum_k w[k] * y_k with the K weights read ON THE DEVICE (the softmaxed alphas never
//                           visit the host), every y_k read once
//   add_weighted_sum_bwd    dy_k = w[k] * dy (for the k that need it) and dw[k] = <dy, y_k> (deterministic two-stage)
//   add_softmax_rows_fwd/bwd  softmax over the last dimension of a small [R, K] matrix (alphas [20, 8], betas [12, 4, 3])
// fp32 / bf16 activations as NHWC views; weights fp32.
#include "common.cuh"

namespace {

constexpr int MX_THREADS = 256;
constexpr int MX_MAXK = 8;

struct LightParams {
  const void* x; void* y; int H, W, C, xs, ys, n;
  const float* w;                 // device [8] in PRIMITIVES order: none, max_pool, avg_pool, skip, sep3, sep5, dil3, dil5
  const float* max_mean; const float* max_inv; const float* avg_mean; const float* avg_inv;   // [C] each (BN affine=False)
  uint32_t flags;
};

template <typename T>
__global__ void __launch_bounds__(MX_THREADS)
mixed_light_kernel(const LightParams p) {
  const unsigned cv = (unsigned)p.C >> 2, row = (unsigned)p.W * cv, total = (unsigned)p.H * row;
  const int n = blockIdx.y;
  const T* xn = static_cast<const T*>(p.x) + (size_t)n * p.H * p.W * p.xs;
  T* yn = static_cast<T*>(p.y) + (size_t)n * p.H * p.W * p.ys;
  const float w_none = __ldg(p.w + 0), w_max = __ldg(p.w + 1), w_avg = __ldg(p.w + 2), w_skip = __ldg(p.w + 3);
  for (unsigned idx = blockIdx.x * (unsigned)MX_THREADS + threadIdx.x; idx < total; idx += gridDim.x * (unsigned)MX_THREADS) {
    const unsigned oy = idx / row, rem = idx - oy * row, ox = rem / cv, c = (rem - ox * cv) * 4;
    float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), sm = make_float4(0.f, 0.f, 0.f, 0.f), ctr = sm;
    int cnt = 0;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = (int)oy - 1 + ky;
      if (iy < 0 || iy >= p.H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = (int)ox - 1 + kx;
        if (ix < 0 || ix >= p.W) continue;
        const float4 v = ld4(xn + ((size_t)iy * p.W + ix) * p.xs + c);
        mx.x = fmaxf(mx.x, v.x); mx.y = fmaxf(mx.y, v.y); mx.z = fmaxf(mx.z, v.z); mx.w = fmaxf(mx.w, v.w);
        sm.x += v.x; sm.y += v.y; sm.z += v.z; sm.w += v.w;
        if (ky == 1 && kx == 1) ctr = v;
        ++cnt;
      }
    }
    const float inv_cnt = 1.f / (float)cnt;                       // count_include_pad=False
    const float4 mm = __ldg(reinterpret_cast<const float4*>(p.max_mean + c)), mi = __ldg(reinterpret_cast<const float4*>(p.max_inv + c));
    const float4 am = __ldg(reinterpret_cast<const float4*>(p.avg_mean + c)), ai = __ldg(reinterpret_cast<const float4*>(p.avg_inv + c));
    float4 o;
    // same association as the reference's sum(w * op(x)) in PRIMITIVES order: none, max_pool, avg_pool, skip_connect
    o.x = w_none * (ctr.x * 0.f) + w_max * ((mx.x - mm.x) * mi.x) + w_avg * ((sm.x * inv_cnt - am.x) * ai.x) + w_skip * ctr.x;
    o.y = w_none * (ctr.y * 0.f) + w_max * ((mx.y - mm.y) * mi.y) + w_avg * ((sm.y * inv_cnt - am.y) * ai.y) + w_skip * ctr.y;
    o.z = w_none * (ctr.z * 0.f) + w_max * ((mx.z - mm.z) * mi.z) + w_avg * ((sm.z * inv_cnt - am.z) * ai.z) + w_skip * ctr.z;
    o.w = w_none * (ctr.w * 0.f) + w_max * ((mx.w - mm.w) * mi.w) + w_avg * ((sm.w * inv_cnt - am.w) * ai.w) + w_skip * ctr.w;
    T* dst = yn + ((size_t)oy * p.W + ox) * p.ys + c;
    if (p.flags & ADD_ACCUMULATE) { const float4 old = ld4(const_cast<const T*>(dst)); o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
    st4(dst, o);
  }
}

struct WsumParams {
  const float* y[MX_MAXK]; int ys[MX_MAXK];        // y_k base pointers / pixel strides (nullptr: primitive skipped)
  float* dyk[MX_MAXK]; int dys[MX_MAXK];           // backward outputs (nullptr: not needed)
  const float* w; int K, C; long long P;
};
__global__ void __launch_bounds__(MX_THREADS)
weighted_sum_fwd_kernel(const WsumParams p, float* __restrict__ out, int os) {
  const int c4n = p.C / 4;
  const long long total = p.P * c4n;
  float w[MX_MAXK];
#pragma unroll
  for (int k = 0; k < MX_MAXK; ++k) w[k] = k < p.K ? __ldg(p.w + k) : 0.f;
  for (long long i = blockIdx.x * (long long)MX_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * MX_THREADS) {
    const long long q = i / c4n; const int v = (int)(i - q * c4n);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < MX_MAXK; ++k) {
      if (k >= p.K || !p.y[k]) continue;
      const float4 t = __ldg(reinterpret_cast<const float4*>(p.y[k] + (size_t)q * p.ys[k]) + v);
      acc.x = fmaf(w[k], t.x, acc.x); acc.y = fmaf(w[k], t.y, acc.y); acc.z = fmaf(w[k], t.z, acc.z); acc.w = fmaf(w[k], t.w, acc.w);
    }
    *(reinterpret_cast<float4*>(out + (size_t)q * os) + v) = acc;
  }
}
// dy_k = w[k] * dy;  partial[block][k] = sum over the block's elements of dy * y_k   (fixed-order reductions)
__global__ void __launch_bounds__(MX_THREADS)
weighted_sum_bwd_kernel(const WsumParams p, const float* __restrict__ dy, int ds, double* __restrict__ part) {
  __shared__ double red[MX_THREADS / 32][MX_MAXK];
  const int c4n = p.C / 4;
  const long long total = p.P * c4n;
  float w[MX_MAXK]; double dot[MX_MAXK];
#pragma unroll
  for (int k = 0; k < MX_MAXK; ++k) { w[k] = k < p.K ? __ldg(p.w + k) : 0.f; dot[k] = 0.0; }
  for (long long i = blockIdx.x * (long long)MX_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * MX_THREADS) {
    const long long q = i / c4n; const int v = (int)(i - q * c4n);
    const float4 g = __ldg(reinterpret_cast<const float4*>(dy + (size_t)q * ds) + v);
#pragma unroll
    for (int k = 0; k < MX_MAXK; ++k) {
      if (k >= p.K) continue;
      if (p.y[k]) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p.y[k] + (size_t)q * p.ys[k]) + v);
        dot[k] += (double)(g.x * t.x + g.y * t.y + g.z * t.z + g.w * t.w);
      }
      if (p.dyk[k])
        *(reinterpret_cast<float4*>(p.dyk[k] + (size_t)q * p.dys[k]) + v) = make_float4(w[k] * g.x, w[k] * g.y, w[k] * g.z, w[k] * g.w);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < MX_MAXK; ++k) {
    double d = dot[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0) red[warp][k] = d;
  }
  __syncthreads();
  if (threadIdx.x < MX_MAXK) {
    double s = 0.0;
    for (int wv = 0; wv < MX_THREADS / 32; ++wv) s += red[wv][threadIdx.x];
    part[(size_t)blockIdx.x * MX_MAXK + threadIdx.x] = s;
  }
}
__global__ void weighted_sum_dw_kernel(const double* __restrict__ part, int blocks, int K, float* __restrict__ dw) {
  const int k = threadIdx.x;
  if (k < K) {
    double s = 0.0;
    for (int b = 0; b < blocks; ++b) s += part[(size_t)b * MX_MAXK + k];
    dw[k] = (float)s;
  }
}

// softmax over the last dimension of [R, K] (K <= 32): one warp per row
__global__ void softmax_rows_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int R, int K) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= R) return;
  float v = lane < K ? x[(size_t)r * K + lane] : -INFINITY;
  float m = v;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float e = lane < K ? expf(v - m) : 0.f, s = e;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane < K) y[(size_t)r * K + lane] = e / s;
}
// dx = y * (dy - sum_j dy_j y_j)
__global__ void softmax_rows_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx, int R, int K) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= R) return;
  const float yv = lane < K ? y[(size_t)r * K + lane] : 0.f, g = lane < K ? dy[(size_t)r * K + lane] : 0.f;
  float s = yv * g;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane < K) dx[(size_t)r * K + lane] = yv * (g - s);
}

inline unsigned mx_blocks(long long items) {
  long long b = (items + MX_THREADS - 1) / MX_THREADS;
  return (unsigned)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}
inline bool mx_f32_vec4(const add_tensor_t* t) {
  return t->dtype == ADD_F32 && t->c % 4 == 0 && t->pix_stride % 4 == 0 && ((uintptr_t)t->ptr % 16) == 0;
}

}  // namespace

extern "C" int add_mixed_light_fwd(const add_tensor_t* x, const add_tensor_t* y, const float* w8_dev, const float* max_mean,
                                   const float* max_inv_std, const float* avg_mean, const float* avg_inv_std, uint32_t flags,
                                   void* stream) {
  ADD_CHECK_ARG(tensor_ok(x) && tensor_ok(y) && w8_dev && max_mean && max_inv_std && avg_mean && avg_inv_std);
  ADD_CHECK_ARG(x->n == y->n && x->h == y->h && x->w == y->w && x->c == y->c);
  ADD_CHECK_SUP(tensor_vec4_ok(x) && tensor_vec4_ok(y) && x->dtype == y->dtype && x->n < 65536 &&
                (long long)x->h * x->w * (x->c / 4) < (1ll << 31));
  LightParams p;
  p.x = x->ptr; p.y = y->ptr; p.H = x->h; p.W = x->w; p.C = x->c; p.xs = x->pix_stride; p.ys = y->pix_stride; p.n = x->n;
  p.w = w8_dev; p.max_mean = max_mean; p.max_inv = max_inv_std; p.avg_mean = avg_mean; p.avg_inv = avg_inv_std; p.flags = flags;
  long long bx = ((long long)x->h * x->w * (x->c / 4) + MX_THREADS - 1) / MX_THREADS;
  const long long cap = (148ll * 8 + x->n - 1) / x->n;
  if (bx > cap) bx = cap;
  dim3 grid((unsigned)(bx < 1 ? 1 : bx), (unsigned)x->n);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (x->dtype == ADD_F32) mixed_light_kernel<float><<<grid, MX_THREADS, 0, s>>>(p);
  else mixed_light_kernel<bf16><<<grid, MX_THREADS, 0, s>>>(p);
  ADD_RETURN_LAUNCH();
}

static int fill_wsum(WsumParams& p, const add_tensor_t* const* ys, const add_tensor_t* const* dys, int K, const add_tensor_t* ref) {
  for (int k = 0; k < MX_MAXK; ++k) { p.y[k] = nullptr; p.ys[k] = 0; p.dyk[k] = nullptr; p.dys[k] = 0; }
  for (int k = 0; k < K; ++k) {
    if (ys && ys[k]) {
      if (!(tensor_ok(ys[k]) && mx_f32_vec4(ys[k]) && ys[k]->n == ref->n && ys[k]->h == ref->h && ys[k]->w == ref->w && ys[k]->c == ref->c))
        return ADD_ERR_UNSUPPORTED;
      p.y[k] = (const float*)ys[k]->ptr; p.ys[k] = ys[k]->pix_stride;
    }
    if (dys && dys[k]) {
      if (!(tensor_ok(dys[k]) && mx_f32_vec4(dys[k]) && dys[k]->n == ref->n && dys[k]->h == ref->h && dys[k]->w == ref->w && dys[k]->c == ref->c))
        return ADD_ERR_UNSUPPORTED;
      p.dyk[k] = (float*)dys[k]->ptr; p.dys[k] = dys[k]->pix_stride;
    }
  }
  p.K = K; p.C = ref->c; p.P = (long long)ref->n * ref->h * ref->w;
  return ADD_OK;
}

extern "C" int add_weighted_sum_fwd(const add_tensor_t* const* ys, int k, const float* w_dev, const add_tensor_t* out, void* stream) {
  ADD_CHECK_ARG(ys && w_dev && tensor_ok(out) && k > 0 && k <= MX_MAXK);
  ADD_CHECK_SUP(mx_f32_vec4(out));
  WsumParams p;
  int rc = fill_wsum(p, ys, nullptr, k, out);
  if (rc != ADD_OK) return rc;
  p.w = w_dev;
  weighted_sum_fwd_kernel<<<mx_blocks(p.P * (p.C / 4)), MX_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p, (float*)out->ptr, out->pix_stride);
  ADD_RETURN_LAUNCH();
}

extern "C" int64_t add_weighted_sum_workspace_bytes(int n, int h, int w, int c) {
  if (n <= 0 || h <= 0 || w <= 0 || c <= 0) return ADD_ERR_BAD_ARG;
  return (int64_t)mx_blocks((long long)n * h * w * (c / 4)) * MX_MAXK * sizeof(double);
}

extern "C" int add_weighted_sum_bwd(const add_tensor_t* dy, const add_tensor_t* const* ys, const add_tensor_t* const* dys, int k,
                                    const float* w_dev, float* dw_dev, void* workspace, int64_t workspace_bytes, void* stream) {
  ADD_CHECK_ARG(tensor_ok(dy) && w_dev && dw_dev && workspace && k > 0 && k <= MX_MAXK);
  ADD_CHECK_SUP(mx_f32_vec4(dy));
  if (workspace_bytes < add_weighted_sum_workspace_bytes(dy->n, dy->h, dy->w, dy->c)) return ADD_ERR_WORKSPACE;
  WsumParams p;
  int rc = fill_wsum(p, ys, dys, k, dy);
  if (rc != ADD_OK) return rc;
  p.w = w_dev;
  const unsigned blocks = mx_blocks(p.P * (p.C / 4));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  weighted_sum_bwd_kernel<<<blocks, MX_THREADS, 0, s>>>(p, (const float*)dy->ptr, dy->pix_stride, (double*)workspace);
  weighted_sum_dw_kernel<<<1, 32, 0, s>>>((const double*)workspace, (int)blocks, k, dw_dev);
  ADD_RETURN_LAUNCH();
}

extern "C" int add_softmax_rows_fwd(const float* x, float* y, int rows, int k, void* stream) {
  ADD_CHECK_ARG(x && y && rows > 0 && k > 0);
  ADD_CHECK_SUP(k <= 32);
  softmax_rows_fwd_kernel<<<ceil_div(rows, 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(x, y, rows, k);
  ADD_RETURN_LAUNCH();
}
extern "C" int add_softmax_rows_bwd(const float* y, const float* dy, float* dx, int rows, int k, void* stream) {
  ADD_CHECK_ARG(y && dy && dx && rows > 0 && k > 0);
  ADD_CHECK_SUP(k <= 32);
  softmax_rows_bwd_kernel<<<ceil_div(rows, 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(y, dy, dx, rows, k);
  ADD_RETURN_LAUNCH();
}
