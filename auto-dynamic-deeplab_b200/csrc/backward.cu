// backward.cu — the training step's backward kernels (train.py:216-247 forward -> mean-over-exits
// CE -> backward -> SGD-nesterov).  fp32 NHWC throughout (the parity path); every reduction is deterministic
// (fixed-order partial sums, no atomics), so a step is bit-reproducible run to run.
//
//   add_relu_mask_bwd          dx *= (x > 0)                       ReLU in front of every conv (operations.py:21,33,47)
//   add_conv2d_wgrad           dW[ky][kx][ci][co] = sum_p relu?(x)[p*s - pad + k*dil][ci] * dy[p][co]
//   add_conv2d_dgrad           dx[q][ci] = sum_{k,co} dy[(q + pad - k*dil)/s][co] * w[k][ci][co]      (any stride)
//   add_depthwise_wgrad        dw[ky][kx][c] = sum_p relu?(x)[p + k - pad][c] * dy[p][c]
//   add_bn_bwd_reduce / _apply training-mode BatchNorm backward (F.batch_norm / SynchronizedBatchNorm2d,
//                              sync_batchnorm/batchnorm.py:59-75,113-125 differentiated): per-channel [sum dy, sum dy*xhat],
//                              then dx = gamma*inv_std*(dy - sum1/M - xhat*sum2/M)
//   add_bilinear_bwd           adjoint of F.interpolate(bilinear, align_corners=False) as a deterministic gather
//   add_pool3x3_bwd            avg (count_include_pad=False) / max (first maximum) 3x3 pool backward, gather form
//   add_ce_loss_fwd_bwd        nn.CrossEntropyLoss(weight, ignore_index) on NCHW logits: loss and d(loss)/d(logits)
//   add_sgd_nesterov           torch.optim.SGD(momentum, weight_decay, nesterov) over a table of parameter tensors
// The dgrad of a stride-1 conv / depthwise conv is the FORWARD kernel with flipped weights (host side, training.py).
#include "common.cuh"

namespace {

constexpr int BW_THREADS = 256;

__device__ __forceinline__ float relu_if(float v, bool on) { return on ? fmaxf(v, 0.f) : v; }

// ---- dx *= (x > 0) ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(BW_THREADS)
relu_mask_kernel(const float* __restrict__ x, int xs, float* __restrict__ dx, int ds, long long pixels, int c4) {
  const long long total = pixels * c4;
  for (long long i = blockIdx.x * (long long)BW_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * BW_THREADS) {
    const long long p = i / c4; const int v = (int)(i - p * c4);
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + p * xs) + v);
    float4* d = reinterpret_cast<float4*>(dx + p * ds) + v;
    float4 g = *d;
    g.x = a.x > 0.f ? g.x : 0.f; g.y = a.y > 0.f ? g.y : 0.f; g.z = a.z > 0.f ? g.z : 0.f; g.w = a.w > 0.f ? g.w : 0.f;
    *d = g;
  }
}

// ---- conv2d wgrad ------------------------------------------------------------------------------------
// GEMM view: dW_tap[ci][co] = A_tap^T [Cin x P] * dY [P x Cout], A_tap = the input pixels tap `tap` pairs with each
// output pixel.  grid = (ci tiles x co tiles, taps, splits over the pixel range); block = 256 threads, each a 4x4
// register tile of a 64x64 (ci, co) tile; 16 pixels per shared-memory stage.  Partials [split][tap][Cin][Cout] are
// summed in fixed order by wgrad_reduce_kernel.
constexpr int WG_T = 64, WG_P = 16;
struct WgradParams {
  const float* x; const float* dy; float* part;
  int N, H, W, Cin, xs, Ho, Wo, Cout, ys;
  int kw, stride, pad, dil, relu_in;
  int ci_tiles, co_tiles; long long P, per;      // P = N*Ho*Wo output pixels, per = pixels per split
};
__global__ void __launch_bounds__(BW_THREADS)
conv2d_wgrad_kernel(const WgradParams p) {
  __shared__ float a_s[WG_P][WG_T + 4];
  __shared__ float d_s[WG_P][WG_T + 4];
  const int tile = blockIdx.x, tap = blockIdx.y, split = blockIdx.z;
  const int ci0 = (tile / p.co_tiles) * WG_T, co0 = (tile % p.co_tiles) * WG_T;
  const int ky = tap / p.kw, kx = tap % p.kw;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;          // thread -> co group tx, ci group ty
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const long long p0 = split * p.per, p1 = min(p.P, p0 + p.per);
  const int lp = threadIdx.x >> 4, lc = (threadIdx.x & 15) * 4;     // loader: pixel lp of the stage, channels lc..lc+3
  for (long long base = p0; base < p1; base += WG_P) {
    const long long q = base + lp;
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), dv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < p1) {
      const int ox = (int)(q % p.Wo); const long long t = q / p.Wo; const int oy = (int)(t % p.Ho); const int n = (int)(t / p.Ho);
      const int iy = oy * p.stride - p.pad + ky * p.dil, ix = ox * p.stride - p.pad + kx * p.dil;
      if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
        const float* xp = p.x + (((size_t)n * p.H + iy) * p.W + ix) * p.xs + ci0 + lc;
        if (ci0 + lc + 3 < p.Cin) av = __ldg(reinterpret_cast<const float4*>(xp));
        else {
          if (ci0 + lc < p.Cin) av.x = __ldg(xp);
          if (ci0 + lc + 1 < p.Cin) av.y = __ldg(xp + 1);
          if (ci0 + lc + 2 < p.Cin) av.z = __ldg(xp + 2);
        }
        if (p.relu_in) { av.x = fmaxf(av.x, 0.f); av.y = fmaxf(av.y, 0.f); av.z = fmaxf(av.z, 0.f); av.w = fmaxf(av.w, 0.f); }
      }
      const float* dp = p.dy + (size_t)q * p.ys + co0 + lc;
      if (co0 + lc + 3 < p.Cout) dv = __ldg(reinterpret_cast<const float4*>(dp));
      else {
        if (co0 + lc < p.Cout) dv.x = __ldg(dp);
        if (co0 + lc + 1 < p.Cout) dv.y = __ldg(dp + 1);
        if (co0 + lc + 2 < p.Cout) dv.z = __ldg(dp + 2);
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&a_s[lp][lc]) = av;
    *reinterpret_cast<float4*>(&d_s[lp][lc]) = dv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < WG_P; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&a_s[k][ty * 4]);
      const float4 d4 = *reinterpret_cast<const float4*>(&d_s[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, d[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], d[j], acc[i][j]);
    }
  }
  float* out = p.part + ((size_t)split * gridDim.y + tap) * p.Cin * p.Cout;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = ci0 + ty * 4 + i;
    if (ci >= p.Cin) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co < p.Cout) out[(size_t)ci * p.Cout + co] = acc[i][j];
    }
  }
}
// dw[i] (+)= sum over splits, fixed order
__global__ void __launch_bounds__(BW_THREADS)
wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, long long n, int splits, int accumulate) {
  for (long long i = blockIdx.x * (long long)BW_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * BW_THREADS) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += part[(size_t)k * n + i];
    dw[i] = accumulate ? dw[i] + s : s;
  }
}

inline int wgrad_splits(long long P, int tiles, int taps) {
  long long s = (148ll * 4 + (long long)tiles * taps - 1) / ((long long)tiles * taps);
  const long long max_by_work = (P + 255) / 256;
  if (s > max_by_work) s = max_by_work;
  return (int)(s < 1 ? 1 : (s > 1024 ? 1024 : s));
}

// ---- conv2d dgrad, any stride (gather form) -------------------------------------------------------------
// thread = (input pixel, 4 input channels); weights [ky][kx][ci][co] read through the read-only cache.
struct DgradParams {
  const float* dy; const float* w; float* dx;
  int N, H, W, Cin, ds, Ho, Wo, Cout, ys, kh, kw, stride, pad, dil, accumulate;
};
__global__ void __launch_bounds__(BW_THREADS)
conv2d_dgrad_kernel(const DgradParams p) {
  const int c4n = (p.Cin + 3) / 4;
  const long long total = (long long)p.N * p.H * p.W * c4n;
  for (long long i = blockIdx.x * (long long)BW_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * BW_THREADS) {
    const int cv = (int)(i % c4n); const long long q = i / c4n;
    const int ix = (int)(q % p.W); const long long t = q / p.W; const int iy = (int)(t % p.H); const int n = (int)(t / p.H);
    const int ci = cv * 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ky = 0; ky < p.kh; ++ky) {
      const int ny = iy + p.pad - ky * p.dil;
      if (ny < 0 || ny % p.stride) continue;
      const int oy = ny / p.stride;
      if (oy >= p.Ho) continue;
      for (int kx = 0; kx < p.kw; ++kx) {
        const int nx = ix + p.pad - kx * p.dil;
        if (nx < 0 || nx % p.stride) continue;
        const int ox = nx / p.stride;
        if (ox >= p.Wo) continue;
        const float* dyp = p.dy + (((size_t)n * p.Ho + oy) * p.Wo + ox) * p.ys;
        const float* wp = p.w + ((size_t)(ky * p.kw + kx) * p.Cin + ci) * p.Cout;
        for (int co = 0; co < p.Cout; ++co) {
          const float g = __ldg(dyp + co);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (ci + u < p.Cin) acc[u] = fmaf(g, __ldg(wp + (size_t)u * p.Cout + co), acc[u]);
        }
      }
    }
    float* d = p.dx + (size_t)q * p.ds + ci;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (ci + u < p.Cin) d[u] = p.accumulate ? d[u] + acc[u] : acc[u];
  }
}

// ---- depthwise wgrad -----------------------------------------------------------------------------------------
// grid (splits); thread = (channel, pixel lane); K*K register accumulators per thread; lanes meet in shared memory in
// fixed order; partial [split][K*K][C]
template <int K>
__global__ void __launch_bounds__(BW_THREADS)
depthwise_wgrad_kernel(const float* __restrict__ x, int xs, const float* __restrict__ dy, int ys, float* __restrict__ part,
                       int N, int H, int W, int C, long long per, int relu_in) {
  extern __shared__ float red[];                   // [lanes][K*K][C] would be large: reduce tap by tap through [lanes][C]
  const int lanes = BW_THREADS / C > 0 ? BW_THREADS / C : 1;
  const int c = threadIdx.x % C, l = threadIdx.x / C;
  const long long P = (long long)N * H * W;
  const long long p0 = blockIdx.x * per, p1 = min(P, p0 + per);
  float acc[K * K];
#pragma unroll
  for (int i = 0; i < K * K; ++i) acc[i] = 0.f;
  if (l < lanes && c < C) {
    for (long long q = p0 + l; q < p1; q += lanes) {
      const int ox = (int)(q % W); const long long t = q / W; const int oy = (int)(t % H); const int n = (int)(t / H);
      const float g = __ldg(dy + (size_t)q * ys + c);
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        const int iy = oy + ky - K / 2;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const int ix = ox + kx - K / 2;
          if (ix < 0 || ix >= W) continue;
          float v = __ldg(x + (((size_t)n * H + iy) * W + ix) * xs + c);
          if (relu_in) v = fmaxf(v, 0.f);
          acc[ky * K + kx] = fmaf(v, g, acc[ky * K + kx]);
        }
      }
    }
  }
  for (int tap = 0; tap < K * K; ++tap) {
    __syncthreads();
    if (l < lanes && c < C) red[l * C + c] = acc[tap];
    __syncthreads();
    if (threadIdx.x < C) {
      float s = 0.f;
      for (int r = 0; r < lanes; ++r) s += red[r * C + threadIdx.x];
      part[((size_t)blockIdx.x * K * K + tap) * C + threadIdx.x] = s;
    }
  }
}

// ---- BatchNorm backward --------------------------------------------------------------------------------------
// y = (x - mean) * inv_std * gamma + beta [-> ReLU].  Pass 1: per-channel sum1 = sum dy', sum2 = sum dy' * xhat with
// dy' = dy masked by (y > 0) when the ReLU was fused.  Same thread mapping and fixed-order merge as the forward
// statistics kernel (batchnorm.cu).
__global__ void __launch_bounds__(BW_THREADS)
bn_bwd_partial_kernel(const float* __restrict__ dy, int dys, const float* __restrict__ x, int xs, const float* __restrict__ mean,
                      const float* __restrict__ inv_std, const float* __restrict__ gamma, const float* __restrict__ beta,
                      long long P, int C, long long per, int relu_out, float* __restrict__ part /* [blocks][2][C] */) {
  extern __shared__ float red[];                   // [lanes][2][C4*4]
  const int c4n = C / 4;
  const int lanes = BW_THREADS / c4n;
  const int v = threadIdx.x % c4n, l = threadIdx.x / c4n;
  const long long p0 = blockIdx.x * per, p1 = min(P, p0 + per);
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  if (l < lanes) {
    const float4 m4 = __ldg(reinterpret_cast<const float4*>(mean) + v), i4 = __ldg(reinterpret_cast<const float4*>(inv_std) + v);
    float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gamma) g4 = __ldg(reinterpret_cast<const float4*>(gamma) + v);
    if (beta) b4 = __ldg(reinterpret_cast<const float4*>(beta) + v);
    const float m[4] = {m4.x, m4.y, m4.z, m4.w}, is[4] = {i4.x, i4.y, i4.z, i4.w};
    const float g[4] = {g4.x, g4.y, g4.z, g4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
    for (long long q = p0 + l; q < p1; q += lanes) {
      const float4 d4 = __ldg(reinterpret_cast<const float4*>(dy + (size_t)q * dys) + v);
      const float4 x4 = __ldg(reinterpret_cast<const float4*>(x + (size_t)q * xs) + v);
      const float d[4] = {d4.x, d4.y, d4.z, d4.w}, xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float xh = (xv[u] - m[u]) * is[u];
        float dd = d[u];
        if (relu_out && !(xh * g[u] + b[u] > 0.f)) dd = 0.f;
        s1[u] += dd; s2[u] = fmaf(dd, xh, s2[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { red[(l * 2 + 0) * C + v * 4 + u] = s1[u]; red[(l * 2 + 1) * C + v * 4 + u] = s2[u]; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += BW_THREADS) {
    const int which = i / C, c = i % C;
    float s = 0.f;
    for (int r = 0; r < lanes; ++r) s += red[(r * 2 + which) * C + c];
    part[((size_t)blockIdx.x * 2 + which) * C + c] = s;
  }
}
__global__ void __launch_bounds__(BW_THREADS)
bn_bwd_finalize_kernel(const float* __restrict__ part, int blocks, int C, double* __restrict__ sums /* [2][C] */) {
  for (int i = blockIdx.x * BW_THREADS + threadIdx.x; i < 2 * C; i += gridDim.x * BW_THREADS) {
    double s = 0.0;
    for (int b = 0; b < blocks; ++b) s += (double)part[(size_t)b * 2 * C + i];
    sums[i] = s;
  }
}
// dx = gamma * inv_std * (dy' - sum1/M - xhat * sum2/M * var_term[c]); var_term = 0 where the synchronised path
// clamped the variance at eps (batchnorm.py:125: clamp has no gradient there), else 1
__global__ void __launch_bounds__(BW_THREADS)
bn_bwd_apply_kernel(const float* __restrict__ dy, int dys, const float* __restrict__ x, int xs, const float* __restrict__ mean,
                    const float* __restrict__ inv_std, const float* __restrict__ gamma, const float* __restrict__ beta,
                    const double* __restrict__ sums, double inv_count, const float* __restrict__ count_dev,
                    const float* __restrict__ var_term, long long P, int C, int relu_out, float* __restrict__ dx, int dxs) {
  if (count_dev) inv_count = 1.0 / (double)__ldg(count_dev);
  const int c4n = C / 4;
  const long long total = P * c4n;
  for (long long i = blockIdx.x * (long long)BW_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * BW_THREADS) {
    const long long q = i / c4n; const int v = (int)(i - q * c4n);
    const float4 d4 = __ldg(reinterpret_cast<const float4*>(dy + (size_t)q * dys) + v);
    const float4 x4 = __ldg(reinterpret_cast<const float4*>(x + (size_t)q * xs) + v);
    const float d[4] = {d4.x, d4.y, d4.z, d4.w}, xv[4] = {x4.x, x4.y, x4.z, x4.w};
    float o[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = v * 4 + u;
      const float m = __ldg(mean + c), is = __ldg(inv_std + c);
      const float g = gamma ? __ldg(gamma + c) : 1.f, b = beta ? __ldg(beta + c) : 0.f;
      const float xh = (xv[u] - m) * is;
      float dd = d[u];
      if (relu_out && !(xh * g + b > 0.f)) dd = 0.f;
      const float m1 = (float)(sums[c] * inv_count), m2 = (float)(sums[C + c] * inv_count) * (var_term ? __ldg(var_term + c) : 1.f);
      o[u] = g * is * (dd - m1 - xh * m2);
    }
    *(reinterpret_cast<float4*>(dx + (size_t)q * dxs) + v) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// ---- bilinear backward ---------------------------------------------------------------------------------------------
// tables per axis (built by bilinear_tables_kernel with the forward's own source-index arithmetic): for output index o:
// i0[o], i1[o], l0[o], l1[o]; for input index i: the contiguous output range [lo[i], hi[i]] that reads it.
__global__ void bilinear_tables_kernel(int in_size, int out_size, float scale, int* __restrict__ i0, int* __restrict__ i1,
                                       float* __restrict__ l0, float* __restrict__ l1) {
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < out_size; o += gridDim.x * blockDim.x) {
    int a, b; float w0, w1;
    bilinear_src(o, scale, in_size, a, b, w0, w1);
    i0[o] = a; i1[o] = b; l0[o] = w0; l1[o] = w1;
  }
}
__global__ void bilinear_ranges_kernel(int in_size, int out_size, const int* __restrict__ i0, const int* __restrict__ i1,
                                       int* __restrict__ lo, int* __restrict__ hi) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < in_size; i += gridDim.x * blockDim.x) {
    int a = out_size, b = -1;
    for (int o = 0; o < out_size; ++o)
      if (i0[o] == i || i1[o] == i) { a = min(a, o); b = max(b, o); }
    lo[i] = a; hi[i] = b;
  }
}
struct BilTab { const int* i0; const int* i1; const float* l0; const float* l1; const int* lo; const int* hi; };
__global__ void __launch_bounds__(BW_THREADS)
bilinear_bwd_kernel(const float* __restrict__ dy, int dys, int Ho, int Wo, float* __restrict__ dx, int dxs, int N, int H, int W,
                    int C, BilTab ty, BilTab tx, int accumulate) {
  const int c4n = C / 4;
  const long long total = (long long)N * H * W * c4n;
  for (long long i = blockIdx.x * (long long)BW_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * BW_THREADS) {
    const int v = (int)(i % c4n); const long long q = i / c4n;
    const int ix = (int)(q % W); const long long t = q / W; const int iy = (int)(t % H); const int n = (int)(t / H);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int y_lo = ty.lo[iy], y_hi = ty.hi[iy], x_lo = tx.lo[ix], x_hi = tx.hi[ix];
    for (int oy = y_lo; oy <= y_hi; ++oy) {
      const float wy = (ty.i0[oy] == iy ? ty.l0[oy] : 0.f) + ((ty.i1[oy] == iy && ty.i1[oy] != ty.i0[oy]) ? ty.l1[oy] : 0.f);
      if (wy == 0.f) continue;
      const float* row = dy + ((size_t)n * Ho + oy) * Wo * dys;
      for (int ox = x_lo; ox <= x_hi; ++ox) {
        const float wx = (tx.i0[ox] == ix ? tx.l0[ox] : 0.f) + ((tx.i1[ox] == ix && tx.i1[ox] != tx.i0[ox]) ? tx.l1[ox] : 0.f);
        if (wx == 0.f) continue;
        const float4 g = __ldg(reinterpret_cast<const float4*>(row + (size_t)ox * dys) + v);
        const float wgt = wy * wx;
        acc.x = fmaf(wgt, g.x, acc.x); acc.y = fmaf(wgt, g.y, acc.y); acc.z = fmaf(wgt, g.z, acc.z); acc.w = fmaf(wgt, g.w, acc.w);
      }
    }
    float4* d = reinterpret_cast<float4*>(dx + (size_t)q * dxs) + v;
    if (accumulate) { const float4 o = *d; acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
    *d = acc;
  }
}

// ---- 3x3 pool backward (stride 1 or 2, padding 1), gather form ---------------------------------------------------
// mode 0: avg, count_include_pad=False: dx[q] = sum over windows containing q of dy[o] / count(o)
// mode 1: max: dx[q] = sum over windows containing q whose FIRST maximum (row-major scan, like ATen) is q of dy[o]
__global__ void __launch_bounds__(BW_THREADS)
pool3x3_bwd_kernel(const float* __restrict__ x, int xs, const float* __restrict__ dy, int dys, float* __restrict__ dx, int dxs,
                   int N, int H, int W, int C, int Ho, int Wo, int mode, int stride, int accumulate) {
  const long long total = (long long)N * H * W * C;
  for (long long i = blockIdx.x * (long long)BW_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * BW_THREADS) {
    const int c = (int)(i % C); const long long q = i / C;
    const int ix = (int)(q % W); const long long t = q / W; const int iy = (int)(t % H); const int n = (int)(t / H);
    float acc = 0.f;
    for (int oy = 0; oy < Ho; ++oy) {
      const int wy0 = oy * stride - 1;
      if (iy < wy0 || iy > wy0 + 2) continue;
      for (int ox = 0; ox < Wo; ++ox) {
        const int wx0 = ox * stride - 1;
        if (ix < wx0 || ix > wx0 + 2) continue;
        const float g = __ldg(dy + (((size_t)n * Ho + oy) * Wo + ox) * dys + c);
        const int ya = max(wy0, 0), yb = min(wy0 + 2, H - 1), xa = max(wx0, 0), xb = min(wx0 + 2, W - 1);
        if (mode == 0) {
          acc += g / (float)((yb - ya + 1) * (xb - xa + 1));
        } else {
          float best = -INFINITY; int by = -1, bx = -1;
          for (int yy = ya; yy <= yb; ++yy)
            for (int xx = xa; xx <= xb; ++xx) {
              const float vv = __ldg(x + (((size_t)n * H + yy) * W + xx) * xs + c);
              if (vv > best || vv != vv) { best = vv; by = yy; bx = xx; }
            }
          if (by == iy && bx == ix) acc += g;
        }
      }
    }
    float* d = dx + (size_t)q * dxs + c;
    *d = accumulate ? *d + acc : acc;
  }
}

// ---- cross entropy on NCHW fp32 logits --------------------------------------------------------------------------------
// nn.CrossEntropyLoss(weight=w, ignore_index): loss = sum_p w[t_p] * (lse_p - z_p[t_p]) / sum_p w[t_p] over valid pixels.
// Pass 1: per-block partial (loss numerator, weight sum) in double; pass 2 (after the host-free finalize): gradient
// dlogits[p][c] = scale * w[t_p] * (softmax_c - [c == t_p]) / wsum, 0 at ignored pixels.
__global__ void __launch_bounds__(BW_THREADS)
ce_partial_kernel(const float* __restrict__ logits, const long long* __restrict__ target, int N, int C, long long HW,
                  long long sn, long long sc, long long sp, long long ignore_index, const float* __restrict__ cw,
                  double* __restrict__ part /* [blocks][2] */) {
  __shared__ double red[2][BW_THREADS / 32];
  double num = 0.0, den = 0.0;
  const long long total = (long long)N * HW;
  for (long long i = blockIdx.x * (long long)BW_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * BW_THREADS) {
    const long long t = target[i];
    if (t == ignore_index || t < 0 || t >= C) continue;
    const int n = (int)(i / HW); const long long pix = i - (long long)n * HW;
    const float* z = logits + (size_t)n * sn + (size_t)pix * sp;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, __ldg(z + (size_t)c * sc));
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(__ldg(z + (size_t)c * sc) - mx);
    const float lse = mx + logf(s);
    const float w = cw ? __ldg(cw + t) : 1.f;
    num += (double)(w * (lse - __ldg(z + (size_t)t * sc)));
    den += (double)w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { num += __shfl_xor_sync(0xffffffffu, num, o); den += __shfl_xor_sync(0xffffffffu, den, o); }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = num; red[1][warp] = den; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k < BW_THREADS / 32; ++k) { a += red[0][k]; b += red[1][k]; }
    part[2 * blockIdx.x] = a; part[2 * blockIdx.x + 1] = b;
  }
}
__global__ void ce_finalize_kernel(const double* __restrict__ part, int blocks, float* __restrict__ out2 /* loss, wsum */) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k < blocks; ++k) { a += part[2 * k]; b += part[2 * k + 1]; }
    out2[0] = (float)(a / b); out2[1] = (float)b;
  }
}
__global__ void __launch_bounds__(BW_THREADS)
ce_grad_kernel(const float* __restrict__ logits, const long long* __restrict__ target, int N, int C, int Cstore, long long HW,
               long long sn, long long sc, long long sp, long long ignore_index, const float* __restrict__ cw,
               const float* __restrict__ loss_wsum, float scale, float* __restrict__ dlogits) {
  const long long total = (long long)N * HW;
  const float inv = scale / loss_wsum[1];
  for (long long i = blockIdx.x * (long long)BW_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * BW_THREADS) {
    const long long t = target[i];
    const int n = (int)(i / HW); const long long pix = i - (long long)n * HW;
    const float* z = logits + (size_t)n * sn + (size_t)pix * sp;
    float* g = dlogits + (size_t)n * sn + (size_t)pix * sp;
    for (int c = C; c < Cstore; ++c) g[(size_t)c * sc] = 0.f;           // padding channels carry no gradient
    if (t == ignore_index || t < 0 || t >= C) {
      for (int c = 0; c < C; ++c) g[(size_t)c * sc] = 0.f;
      continue;
    }
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, __ldg(z + (size_t)c * sc));
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(__ldg(z + (size_t)c * sc) - mx);
    const float w = (cw ? __ldg(cw + t) : 1.f) * inv, rs = 1.f / s;
    for (int c = 0; c < C; ++c) {
      const float pr = expf(__ldg(z + (size_t)c * sc) - mx) * rs;
      g[(size_t)c * sc] = w * (pr - (c == t ? 1.f : 0.f));
    }
  }
}

// ---- SGD with momentum / weight decay / nesterov over a table of tensors ---------------------------------------
// torch.optim.SGD: g = grad + wd * p; buf = momentum * buf + g (buf = g on the first step); g = g + momentum * buf
// (nesterov) or buf; p -= lr * g.  One launch for all parameters: blockIdx.y = tensor.
struct SgdEntry { float* p; const float* g; float* buf; long long n; };
__global__ void __launch_bounds__(BW_THREADS)
sgd_kernel(const SgdEntry* __restrict__ tab, float lr, const float* __restrict__ lr_dev, float momentum, float wd, int nesterov,
           int first_step) {
  if (lr_dev) lr = __ldg(lr_dev);              // the poly schedule changes lr every iteration: a captured graph reads it here
  const SgdEntry e = tab[blockIdx.y];
  for (long long i = blockIdx.x * (long long)BW_THREADS + threadIdx.x; i < e.n; i += (long long)gridDim.x * BW_THREADS) {
    float g = e.g[i] + wd * e.p[i];
    if (momentum != 0.f) {
      const float b = first_step ? g : momentum * e.buf[i] + g;
      e.buf[i] = b;
      g = nesterov ? g + momentum * b : b;
    }
    e.p[i] -= lr * g;
  }
}

inline unsigned bw_blocks(long long items) {
  long long b = (items + BW_THREADS - 1) / BW_THREADS;
  return (unsigned)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}
inline bool f32_vec4(const add_tensor_t* t) {
  return t->dtype == ADD_F32 && t->c % 4 == 0 && t->pix_stride % 4 == 0 && ((uintptr_t)t->ptr % 16) == 0;
}

}  // namespace

extern "C" int add_relu_mask_bwd(const add_tensor_t* x, const add_tensor_t* dx, void* stream) {
  ADD_CHECK_ARG(tensor_ok(x) && tensor_ok(dx) && x->n == dx->n && x->h == dx->h && x->w == dx->w && x->c == dx->c);
  ADD_CHECK_SUP(f32_vec4(x) && f32_vec4(dx));
  const long long pixels = (long long)x->n * x->h * x->w;
  relu_mask_kernel<<<bw_blocks(pixels * (x->c / 4)), BW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      (const float*)x->ptr, x->pix_stride, (float*)dx->ptr, dx->pix_stride, pixels, x->c / 4);
  ADD_RETURN_LAUNCH();
}

extern "C" int64_t add_conv2d_wgrad_workspace_bytes(int n, int ho, int wo, int cin, int cout, int kh, int kw) {
  if (n <= 0 || ho <= 0 || wo <= 0 || cin <= 0 || cout <= 0 || kh <= 0 || kw <= 0) return ADD_ERR_BAD_ARG;
  const int tiles = ceil_div(cin, WG_T) * ceil_div(cout, WG_T);
  const int s = wgrad_splits((long long)n * ho * wo, tiles, kh * kw);
  return (int64_t)s * kh * kw * cin * cout * sizeof(float);
}

extern "C" int add_conv2d_wgrad(const add_tensor_t* x, const add_tensor_t* dy, float* dw, int kh, int kw, int stride, int pad,
                                int dil, uint32_t flags, void* workspace, int64_t workspace_bytes, void* stream) {
  ADD_CHECK_ARG(tensor_ok(x) && tensor_ok(dy) && dw && workspace && x->n == dy->n && kh > 0 && kw > 0 && stride > 0 && dil > 0);
  ADD_CHECK_SUP(x->dtype == ADD_F32 && dy->dtype == ADD_F32 && x->pix_stride % 4 == 0 && dy->pix_stride % 4 == 0 &&
                ((uintptr_t)x->ptr % 16) == 0 && ((uintptr_t)dy->ptr % 16) == 0);
  if (workspace_bytes < add_conv2d_wgrad_workspace_bytes(x->n, dy->h, dy->w, x->c, dy->c, kh, kw)) return ADD_ERR_WORKSPACE;
  WgradParams p;
  p.x = (const float*)x->ptr; p.dy = (const float*)dy->ptr; p.part = (float*)workspace;
  p.N = x->n; p.H = x->h; p.W = x->w; p.Cin = x->c; p.xs = x->pix_stride;
  p.Ho = dy->h; p.Wo = dy->w; p.Cout = dy->c; p.ys = dy->pix_stride;
  p.kw = kw; p.stride = stride; p.pad = pad; p.dil = dil; p.relu_in = (flags & ADD_RELU_IN) ? 1 : 0;
  p.ci_tiles = ceil_div(x->c, WG_T); p.co_tiles = ceil_div(dy->c, WG_T);
  p.P = (long long)x->n * dy->h * dy->w;
  const int tiles = p.ci_tiles * p.co_tiles, taps = kh * kw;
  const int S = wgrad_splits(p.P, tiles, taps);
  p.per = (p.P + S - 1) / S;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  conv2d_wgrad_kernel<<<dim3(tiles, taps, S), BW_THREADS, 0, s>>>(p);
  const long long nw = (long long)taps * x->c * dy->c;
  wgrad_reduce_kernel<<<bw_blocks(nw), BW_THREADS, 0, s>>>((const float*)workspace, dw, nw, S, (flags & ADD_ACCUMULATE) ? 1 : 0);
  ADD_RETURN_LAUNCH();
}

extern "C" int add_conv2d_dgrad(const add_tensor_t* dy, const float* w, const add_tensor_t* dx, int kh, int kw, int stride,
                                int pad, int dil, uint32_t flags, void* stream) {
  ADD_CHECK_ARG(tensor_ok(dy) && tensor_ok(dx) && w && dy->n == dx->n && kh > 0 && kw > 0 && stride > 0 && dil > 0);
  ADD_CHECK_SUP(dy->dtype == ADD_F32 && dx->dtype == ADD_F32);
  DgradParams p;
  p.dy = (const float*)dy->ptr; p.w = w; p.dx = (float*)dx->ptr;
  p.N = dx->n; p.H = dx->h; p.W = dx->w; p.Cin = dx->c; p.ds = dx->pix_stride;
  p.Ho = dy->h; p.Wo = dy->w; p.Cout = dy->c; p.ys = dy->pix_stride;
  p.kh = kh; p.kw = kw; p.stride = stride; p.pad = pad; p.dil = dil; p.accumulate = (flags & ADD_ACCUMULATE) ? 1 : 0;
  const long long total = (long long)dx->n * dx->h * dx->w * ((dx->c + 3) / 4);
  conv2d_dgrad_kernel<<<bw_blocks(total), BW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
  ADD_RETURN_LAUNCH();
}

static inline int dw_wgrad_blocks(long long P) {
  long long b = (P + 1023) / 1024;
  return (int)(b < 1 ? 1 : (b > 148 * 4 ? 148 * 4 : b));
}
extern "C" int64_t add_depthwise_wgrad_workspace_bytes(int n, int h, int w, int c, int k) {
  if (n <= 0 || h <= 0 || w <= 0 || c <= 0 || k <= 0) return ADD_ERR_BAD_ARG;
  return (int64_t)dw_wgrad_blocks((long long)n * h * w) * k * k * c * sizeof(float);
}
extern "C" int add_depthwise_wgrad(const add_tensor_t* x, const add_tensor_t* dy, float* dw, int k, uint32_t flags,
                                   void* workspace, int64_t workspace_bytes, void* stream) {
  ADD_CHECK_ARG(tensor_ok(x) && tensor_ok(dy) && dw && workspace && x->n == dy->n && x->h == dy->h && x->w == dy->w && x->c == dy->c);
  ADD_CHECK_SUP((k == 3 || k == 5) && x->dtype == ADD_F32 && dy->dtype == ADD_F32 && x->c <= BW_THREADS);
  if (workspace_bytes < add_depthwise_wgrad_workspace_bytes(x->n, x->h, x->w, x->c, k)) return ADD_ERR_WORKSPACE;
  const long long P = (long long)x->n * x->h * x->w;
  const int B = dw_wgrad_blocks(P);
  const long long per = (P + B - 1) / B;
  const int lanes = BW_THREADS / x->c > 0 ? BW_THREADS / x->c : 1;
  const size_t smem = (size_t)lanes * x->c * sizeof(float);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int relu = (flags & ADD_RELU_IN) ? 1 : 0;
  if (k == 3)
    depthwise_wgrad_kernel<3><<<B, BW_THREADS, smem, s>>>((const float*)x->ptr, x->pix_stride, (const float*)dy->ptr, dy->pix_stride,
                                                          (float*)workspace, x->n, x->h, x->w, x->c, per, relu);
  else
    depthwise_wgrad_kernel<5><<<B, BW_THREADS, smem, s>>>((const float*)x->ptr, x->pix_stride, (const float*)dy->ptr, dy->pix_stride,
                                                          (float*)workspace, x->n, x->h, x->w, x->c, per, relu);
  const long long nw = (long long)k * k * x->c;
  wgrad_reduce_kernel<<<bw_blocks(nw), BW_THREADS, 0, s>>>((const float*)workspace, dw, nw, B, (flags & ADD_ACCUMULATE) ? 1 : 0);
  ADD_RETURN_LAUNCH();
}

static inline int bn_bwd_blocks(long long P) {
  long long b = (P + 511) / 512;
  return (int)(b < 1 ? 1 : (b > 148 * 4 ? 148 * 4 : b));
}
extern "C" int64_t add_bn_bwd_workspace_bytes(int n, int h, int w, int c) {
  if (n <= 0 || h <= 0 || w <= 0 || c <= 0) return ADD_ERR_BAD_ARG;
  return (int64_t)bn_bwd_blocks((long long)n * h * w) * 2 * c * sizeof(float);
}
/* sums: double [2][C] = per-rank [sum dy', sum dy'*xhat] (the caller all-reduces them for SynchronizedBatchNorm2d) */
extern "C" int add_bn_bwd_reduce(const add_tensor_t* dy, const add_tensor_t* x, const float* mean, const float* inv_std,
                                 const float* gamma, const float* beta, uint32_t flags, double* sums, void* workspace,
                                 int64_t workspace_bytes, void* stream) {
  ADD_CHECK_ARG(tensor_ok(dy) && tensor_ok(x) && mean && inv_std && sums && workspace);
  ADD_CHECK_ARG(dy->n == x->n && dy->h == x->h && dy->w == x->w && dy->c == x->c);
  ADD_CHECK_SUP(f32_vec4(dy) && f32_vec4(x) && x->c / 4 <= BW_THREADS);
  if (workspace_bytes < add_bn_bwd_workspace_bytes(x->n, x->h, x->w, x->c)) return ADD_ERR_WORKSPACE;
  const long long P = (long long)x->n * x->h * x->w;
  const int B = bn_bwd_blocks(P);
  const long long per = (P + B - 1) / B;
  const int lanes = BW_THREADS / (x->c / 4);
  const size_t smem = (size_t)lanes * 2 * x->c * sizeof(float);
  ADD_CHECK_SUP(smem <= 48 * 1024);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bn_bwd_partial_kernel<<<B, BW_THREADS, smem, s>>>((const float*)dy->ptr, dy->pix_stride, (const float*)x->ptr, x->pix_stride, mean,
                                                    inv_std, gamma, beta, P, x->c, per, (flags & ADD_RELU_OUT) ? 1 : 0,
                                                    (float*)workspace);
  bn_bwd_finalize_kernel<<<ceil_div(2 * x->c, BW_THREADS), BW_THREADS, 0, s>>>((const float*)workspace, B, x->c, sums);
  ADD_RETURN_LAUNCH();
}
extern "C" int add_bn_bwd_apply(const add_tensor_t* dy, const add_tensor_t* x, const float* mean, const float* inv_std,
                                const float* gamma, const float* beta, const double* sums, double inv_count,
                                const float* count_dev, const float* var_term, uint32_t flags, const add_tensor_t* dx,
                                void* stream) {
  ADD_CHECK_ARG(tensor_ok(dy) && tensor_ok(x) && tensor_ok(dx) && mean && inv_std && sums);
  ADD_CHECK_ARG(dy->n == x->n && dy->h == x->h && dy->w == x->w && dy->c == x->c && dx->c == x->c && dx->n == x->n);
  ADD_CHECK_SUP(f32_vec4(dy) && f32_vec4(x) && f32_vec4(dx));
  const long long P = (long long)x->n * x->h * x->w;
  bn_bwd_apply_kernel<<<bw_blocks(P * (x->c / 4)), BW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      (const float*)dy->ptr, dy->pix_stride, (const float*)x->ptr, x->pix_stride, mean, inv_std, gamma, beta, sums, inv_count,
      count_dev, var_term, P, x->c, (flags & ADD_RELU_OUT) ? 1 : 0, (float*)dx->ptr, dx->pix_stride);
  ADD_RETURN_LAUNCH();
}

/* tables for one axis: int32 i0[out], i1[out], float l0[out], l1[out], int32 lo[in], hi[in] (device buffers) */
extern "C" int add_bilinear_bwd_tables(int in_size, int out_size, int32_t* i0, int32_t* i1, float* l0, float* l1, int32_t* lo,
                                       int32_t* hi, void* stream) {
  ADD_CHECK_ARG(in_size > 0 && out_size > 0 && i0 && i1 && l0 && l1 && lo && hi);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bilinear_tables_kernel<<<ceil_div(out_size, 128), 128, 0, s>>>(in_size, out_size, (float)in_size / (float)out_size, i0, i1, l0, l1);
  bilinear_ranges_kernel<<<ceil_div(in_size, 128), 128, 0, s>>>(in_size, out_size, i0, i1, lo, hi);
  ADD_RETURN_LAUNCH();
}
extern "C" int add_bilinear_bwd(const add_tensor_t* dy, const add_tensor_t* dx, const int32_t* yi0, const int32_t* yi1,
                                const float* yl0, const float* yl1, const int32_t* ylo, const int32_t* yhi, const int32_t* xi0,
                                const int32_t* xi1, const float* xl0, const float* xl1, const int32_t* xlo, const int32_t* xhi,
                                uint32_t flags, void* stream) {
  ADD_CHECK_ARG(tensor_ok(dy) && tensor_ok(dx) && dy->n == dx->n && dy->c == dx->c && yi0 && xi0);
  ADD_CHECK_SUP(f32_vec4(dy) && f32_vec4(dx));
  BilTab ty{yi0, yi1, yl0, yl1, ylo, yhi}, tx{xi0, xi1, xl0, xl1, xlo, xhi};
  const long long total = (long long)dx->n * dx->h * dx->w * (dx->c / 4);
  bilinear_bwd_kernel<<<bw_blocks(total), BW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      (const float*)dy->ptr, dy->pix_stride, dy->h, dy->w, (float*)dx->ptr, dx->pix_stride, dx->n, dx->h, dx->w, dx->c, ty, tx,
      (flags & ADD_ACCUMULATE) ? 1 : 0);
  ADD_RETURN_LAUNCH();
}

extern "C" int add_pool3x3_bwd(const add_tensor_t* x, const add_tensor_t* dy, const add_tensor_t* dx, int mode, int stride,
                               uint32_t flags, void* stream) {
  ADD_CHECK_ARG(tensor_ok(x) && tensor_ok(dy) && tensor_ok(dx) && (mode == 0 || mode == 1) && (stride == 1 || stride == 2));
  ADD_CHECK_ARG(x->n == dy->n && x->c == dy->c && dx->n == x->n && dx->h == x->h && dx->w == x->w && dx->c == x->c);
  ADD_CHECK_SUP(x->dtype == ADD_F32 && dy->dtype == ADD_F32 && dx->dtype == ADD_F32);
  const long long total = (long long)x->n * x->h * x->w * x->c;
  pool3x3_bwd_kernel<<<bw_blocks(total), BW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      (const float*)x->ptr, x->pix_stride, (const float*)dy->ptr, dy->pix_stride, (float*)dx->ptr, dx->pix_stride, x->n, x->h,
      x->w, x->c, dy->h, dy->w, mode, stride, (flags & ADD_ACCUMULATE) ? 1 : 0);
  ADD_RETURN_LAUNCH();
}

static inline int ce_blocks(long long total) {
  long long b = (total + 1023) / 1024;
  return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}
extern "C" int64_t add_ce_loss_workspace_bytes(int n, int h, int w) {
  if (n <= 0 || h <= 0 || w <= 0) return ADD_ERR_BAD_ARG;
  return (int64_t)ce_blocks((long long)n * h * w) * 2 * sizeof(double);
}
/* logits element (n, c, pixel) lives at n*stride_n + c*stride_c + pixel*stride_pix (NCHW: HW*C, HW, 1; NHWC with padded
 * channels: HW*Cs, 1, Cs); channels [num_class, c_store) are padding: ignored, zero gradient.
 * loss_wsum: float[2] = (mean loss, sum of the valid pixels' class weights).  dlogits NULL: forward only (same strides).
 * grad_scale multiplies the gradient (1/C for the mean over the C exits, train.py:233). */
extern "C" int add_ce_loss_fwd_bwd(const float* logits, const int64_t* target, int n, int num_class, int h, int w,
                                   int64_t stride_n, int64_t stride_c, int64_t stride_pix, int c_store,
                                   int64_t ignore_index, const float* class_weight, float grad_scale, float* loss_wsum,
                                   float* dlogits, void* workspace, int64_t workspace_bytes, void* stream) {
  ADD_CHECK_ARG(logits && target && loss_wsum && workspace && n > 0 && num_class > 0 && h > 0 && w > 0);
  if (workspace_bytes < add_ce_loss_workspace_bytes(n, h, w)) return ADD_ERR_WORKSPACE;
  const long long HW = (long long)h * w;
  const int B = ce_blocks(HW * n);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ADD_CHECK_ARG(c_store >= num_class && stride_c > 0 && stride_pix > 0 && stride_n > 0);
  ce_partial_kernel<<<B, BW_THREADS, 0, s>>>(logits, (const long long*)target, n, num_class, HW, stride_n, stride_c, stride_pix,
                                             ignore_index, class_weight, (double*)workspace);
  ce_finalize_kernel<<<1, 32, 0, s>>>((const double*)workspace, B, loss_wsum);
  if (dlogits)
    ce_grad_kernel<<<bw_blocks(HW * n), BW_THREADS, 0, s>>>(logits, (const long long*)target, n, num_class, c_store, HW, stride_n,
                                                            stride_c, stride_pix, ignore_index, class_weight, loss_wsum,
                                                            grad_scale, dlogits);
  ADD_RETURN_LAUNCH();
}

/* table: device array of n_tensors {float* param, const float* grad, float* momentum_buf, int64 numel} (32 bytes each) */
extern "C" int add_sgd_nesterov(const void* table_dev, int n_tensors, int64_t max_numel, float lr, const float* lr_dev,
                                float momentum, float weight_decay, int nesterov, int first_step, void* stream) {
  ADD_CHECK_ARG(table_dev && n_tensors > 0 && max_numel > 0);
  ADD_CHECK_SUP(n_tensors < 65536);
  long long bx = (max_numel + BW_THREADS * 4 - 1) / (BW_THREADS * 4);
  if (bx > 64) bx = 64;
  sgd_kernel<<<dim3((unsigned)bx, (unsigned)n_tensors), BW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      (const SgdEntry*)table_dev, lr, lr_dev, momentum, weight_decay, nesterov, first_step);
  ADD_RETURN_LAUNCH();
}
