// resize.cu — memory-bound NHWC helpers: bilinear resize, NCHW<->NHWC edges, global average
// pool, EDM MLP tail.  All HBM-bound: vectorised 4-channel accesses, grid sized to the data.
#include "common.cuh"

namespace {

// ---- bilinear, align_corners=False -----------------------------------------------------------
// Thread = one output pixel x V consecutive channels (V = 8 for 16-byte-aligned bf16 views: four 16 B loads,
// one 16 B store; V = 4 otherwise).  grid = (blocks over Ho*Wo*C/V, N); 32-bit index arithmetic.
template <int V> struct VecIO;
template <> struct VecIO<4> {
  template <typename T> static __device__ __forceinline__ void load(const T* p, float (&v)[4]) {
    float4 t = ld4(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  template <typename T> static __device__ __forceinline__ void store(T* p, const float (&v)[4]) {
    st4(p, make_float4(v[0], v[1], v[2], v[3]));
  }
};
template <> struct VecIO<8> {
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[8]) {
    uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(u[i] << 16); v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u); }
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[8]) {
    uint4 r;
    uint32_t* u = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); u[i] = *reinterpret_cast<uint32_t*>(&b); }
    *reinterpret_cast<uint4*>(p) = r;
  }
};

// one output pixel x V channels (the whole generic kernel; also the border path of the integer-upscale kernel)
template <typename TI, typename TO, int V>
__device__ __forceinline__ void bilinear_pixel(const TI* __restrict__ xn, TO* __restrict__ yn, int oy, int ox, int c, int H, int W,
                                               int xs, int Wo, int ys, float sh, float sw, uint32_t flags) {
  int y0, y1, x0, x1; float hl0, hl1, wl0, wl1;
  bilinear_src(oy, sh, H, y0, y1, hl0, hl1);
  bilinear_src(ox, sw, W, x0, x1, wl0, wl1);
  float v00[V], v01[V], v10[V], v11[V], r[V];
  VecIO<V>::load(xn + ((size_t)y0 * W + x0) * xs + c, v00);
  VecIO<V>::load(xn + ((size_t)y0 * W + x1) * xs + c, v01);
  VecIO<V>::load(xn + ((size_t)y1 * W + x0) * xs + c, v10);
  VecIO<V>::load(xn + ((size_t)y1 * W + x1) * xs + c, v11);
  const bool relu_in = flags & ADD_RELU_IN, relu_out = flags & ADD_RELU_OUT;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float a = v00[i], b = v01[i], cc = v10[i], d = v11[i];
    if (relu_in) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); cc = fmaxf(cc, 0.f); d = fmaxf(d, 0.f); }
    float o = hl0 * (wl0 * a + wl1 * b) + hl1 * (wl0 * cc + wl1 * d);
    r[i] = relu_out ? fmaxf(o, 0.f) : o;
  }
  VecIO<V>::store(yn + ((size_t)oy * Wo + ox) * ys + c, r);
}

template <typename TI, typename TO, int V>
__global__ void __launch_bounds__(256)
bilinear_kernel(const TI* __restrict__ x, TO* __restrict__ y, int H, int W, int C, int xs,
                int Ho, int Wo, int ys, float sh, float sw, uint32_t flags) {
  const unsigned cv = (unsigned)C / V, row = (unsigned)Wo * cv, total = (unsigned)Ho * row;
  const int n = blockIdx.y;
  const TI* xn = x + (size_t)n * H * W * xs;
  TO* yn = y + (size_t)n * Ho * Wo * ys;
  for (unsigned idx = blockIdx.x * 256u + threadIdx.x; idx < total; idx += gridDim.x * 256u) {
    const unsigned oy = idx / row, rem = idx - oy * row, ox = rem / cv, c = (rem - ox * cv) * V;
    bilinear_pixel<TI, TO, V>(xn, yn, (int)oy, (int)ox, (int)c, H, W, xs, Wo, ys, sh, sw, flags);
  }
}

// ---- upscale, bf16, 8 channels per thread: one thread per SOURCE cell -----------------------------------
// The generic kernel is ISSUE-bound, not HBM-bound, when it upsamples (ncu r01y: issue slots 67-71 % busy at 54-70 %
// of DRAM peak; the ~x4 exit resize 63x127 -> 256x512 ran at 1.75 TB/s): every output pixel pays two integer
// divisions, two source-index computations, four 16-byte gathers, 32 bf16->fp32 conversions and 48 FLOPs for the 16
// bytes it stores.  All outputs whose source index pair is (y0, x0) read the same four source pixels, so here a thread
// owns a source cell (y0, x0) for 8 channels: sources loaded and converted once, the horizontal lerp of a column
// shared by the cell's rows — for a x4 upscale ~4x fewer instructions per output.  Which outputs belong to a cell, and
// their weights, come from bilinear_src itself (the index is monotone in the output coordinate, so a cell's outputs
// are a contiguous range found by walking from an under-estimate); every output is the same expression as in
// bilinear_pixel.  Results are bit-identical to the generic kernel's, borders (clamped sources) included.
__device__ __forceinline__ int bilinear_first_out(int cell, float scale, int in_size, int out_size) {
  // smallest output index whose source index i0 is >= cell.  Cell 0 also owns the clamped outputs (source coordinate
  // < 0 -> index 0), so it starts at output 0 whatever the scale (the estimate below is only an under-estimate for
  // cell >= 1: with scale factors above ~6 it would skip cell 0's first outputs)
  if (cell <= 0) return 0;
  int e = (int)floorf((static_cast<float>(cell) + 0.5f) / scale - 0.5f) - 1;
  e = e < 0 ? 0 : e;
  while (e < out_size) {
    int i0, i1; float l0, l1;
    bilinear_src(e, scale, in_size, i0, i1, l0, l1);
    if (i0 >= cell) break;
    ++e;
  }
  return e;
}

__global__ void __launch_bounds__(256)
bilinear_up_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int H, int W, int C, int xs,
                   int Ho, int Wo, int ys, float sh, float sw, uint32_t flags) {
  const unsigned cv = (unsigned)C / 8, row = (unsigned)W * cv, total = (unsigned)H * row;
  const int n = blockIdx.y;
  const bf16* xn = x + (size_t)n * H * W * xs;
  bf16* yn = y + (size_t)n * Ho * Wo * ys;
  const bool relu_in = flags & ADD_RELU_IN, relu_out = flags & ADD_RELU_OUT;
  for (unsigned idx = blockIdx.x * 256u + threadIdx.x; idx < total; idx += gridDim.x * 256u) {
    const unsigned byu = idx / row, rem = idx - byu * row, bxu = rem / cv;
    const int c = (int)(rem - bxu * cv) * 8;
    const int by = (int)byu, bx = (int)bxu;                      // source cell: rows (by, by1), columns (bx, bx1)
    const int by1 = by + (by < H - 1 ? 1 : 0), bx1 = bx + (bx < W - 1 ? 1 : 0);
    const int oy_lo = bilinear_first_out(by, sh, H, Ho), ox_lo = bilinear_first_out(bx, sw, W, Wo);
    float a[8], b[8], cc[8], d[8];
    VecIO<8>::load(xn + ((size_t)by * W + bx) * xs + c, a);
    VecIO<8>::load(xn + ((size_t)by * W + bx1) * xs + c, b);
    VecIO<8>::load(xn + ((size_t)by1 * W + bx) * xs + c, cc);
    VecIO<8>::load(xn + ((size_t)by1 * W + bx1) * xs + c, d);
    if (relu_in) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] = fmaxf(a[i], 0.f); b[i] = fmaxf(b[i], 0.f); cc[i] = fmaxf(cc[i], 0.f); d[i] = fmaxf(d[i], 0.f); }
    }
    for (int ox = ox_lo; ox < Wo; ++ox) {
      int x0, x1; float wl0, wl1;
      bilinear_src(ox, sw, W, x0, x1, wl0, wl1);
      if (x0 != bx) break;
      float top[8], bot[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { top[i] = wl0 * a[i] + wl1 * b[i]; bot[i] = wl0 * cc[i] + wl1 * d[i]; }
      for (int oy = oy_lo; oy < Ho; ++oy) {
        int y0, y1; float hl0, hl1;
        bilinear_src(oy, sh, H, y0, y1, hl0, hl1);
        if (y0 != by) break;
        float r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float o = hl0 * top[i] + hl1 * bot[i];
          r[i] = relu_out ? fmaxf(o, 0.f) : o;
        }
        VecIO<8>::store(yn + ((size_t)oy * Wo + ox) * ys + c, r);
      }
    }
  }
}

// ---- NCHW fp32 -> NHWC (tile transpose through smem) -------------------------------------------
template <typename TO>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ src, TO* __restrict__ dst, int C, int HW, int Cdst, int ys) {
  __shared__ float tile[32][33];
  int n = blockIdx.z;
  int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    int c = c0 + j, p = p0 + tx;
    tile[j][tx] = (c < C && p < HW) ? __ldg(src + ((size_t)n * C + c) * HW + p) : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    int p = p0 + j, c = c0 + tx;
    if (p < HW && c < Cdst) st1(dst + ((size_t)n * HW + p) * ys + c, tile[tx][j]);
  }
}

template <typename TI>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const TI* __restrict__ src, float* __restrict__ dst, int C, int HW, int xs) {
  __shared__ float tile[32][33];
  int n = blockIdx.z;
  int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    int p = p0 + j, c = c0 + tx;
    tile[j][tx] = (p < HW && c < C) ? ld1(src + ((size_t)n * HW + p) * xs + c) : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    int c = c0 + j, p = p0 + tx;
    if (c < C && p < HW) dst[((size_t)n * C + c) * HW + p] = tile[tx][j];
  }
}

// ---- global average pool: two deterministic stages ------------------------------------------------
// stage 1: grid (splits, n); each block sums a contiguous pixel range for ALL channels (thread = channel vector x
// pixel lane, 4 independent loads in flight per thread), block-reduces over pixel lanes in smem and writes
// partial[n][split][C].  stage 2: one block per image sums the splits in fixed order.
// V = channels per thread: 8 for 16-byte-aligned bf16 views (16 B loads), else 4.
constexpr int GAP_THREADS = 256;
__host__ __device__ inline int gap_splits(int HW, int n) {
  int s = (HW + 255) / 256;
  int cap = (148 * 8 + n - 1) / n;
  return s < 1 ? 1 : (s > cap ? cap : s);
}

template <typename TI, int V>
__global__ void __launch_bounds__(GAP_THREADS)
gap_partial_kernel(const TI* __restrict__ x, float* __restrict__ part, int HW, int C, int xs, uint32_t flags) {
  extern __shared__ float gap_red[];                 // [lanes][Cg]
  // blockIdx.z = channel group of up to GAP_THREADS * V channels (wide inputs: the ASPP image pool at Cin = 3200)
  const int c0 = blockIdx.z * (GAP_THREADS * V);
  const int Cg = min(C - c0, GAP_THREADS * V);
  const int cv = Cg / V;
  const int lanes = GAP_THREADS / cv;                // pixel lanes (>= 1)
  const int n = blockIdx.y, S = gridDim.x;
  const int per = (HW + S - 1) / S;
  const int p0 = blockIdx.x * per, p1 = min(HW, p0 + per);
  const int v = threadIdx.x % cv, l = threadIdx.x / cv;
  const bool relu = flags & ADD_RELU_IN;
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  if (l < lanes) {
    const TI* xn = x + (size_t)n * HW * xs + c0 + v * V;
    int p = p0 + l;
    for (; p + 3 * lanes < p1; p += 4 * lanes) {
      float t[4][V];
#pragma unroll
      for (int u = 0; u < 4; ++u) VecIO<V>::load(xn + (size_t)(p + u * lanes) * xs, t[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] += relu ? fmaxf(t[u][i], 0.f) : t[u][i];
    }
    for (; p < p1; p += lanes) {
      float t[V];
      VecIO<V>::load(xn + (size_t)p * xs, t);
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] += relu ? fmaxf(t[i], 0.f) : t[i];
    }
#pragma unroll
    for (int i = 0; i < V; ++i) gap_red[l * Cg + v * V + i] = acc[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cg; c += GAP_THREADS) {
    float tot = gap_red[c];
    for (int r = 1; r < lanes; ++r) tot += gap_red[r * Cg + c];
    part[((size_t)n * S + blockIdx.x) * C + c0 + c] = tot;
  }
}

__global__ void __launch_bounds__(GAP_THREADS)
gap_finalize_kernel(const float* __restrict__ part, float* __restrict__ out, int S, int C, float inv_hw) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += GAP_THREADS) {
    float s = 0.f;
    for (int i = 0; i < S; ++i) s += part[((size_t)n * S + i) * C + c];
    out[(size_t)n * C + c] = s * inv_hw;
  }
}

// ---- EDM MLP tail: one block (128 threads) per image -------------------------------------------
__global__ void __launch_bounds__(128)
edm_mlp_kernel(const float* __restrict__ pooled, const float* __restrict__ w0, const float* __restrict__ b0,
               const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
               const float* __restrict__ b2, float* __restrict__ out) {
  __shared__ float a[128], h0[64], h1[32];
  int n = blockIdx.x, t = threadIdx.x;
  a[t] = pooled[(size_t)n * 128 + t];
  __syncthreads();
  if (t < 64) {
    float s = b0[t];
    for (int k = 0; k < 128; ++k) s = fmaf(w0[t * 128 + k], a[k], s);
    h0[t] = fmaxf(s, 0.f);
  }
  __syncthreads();
  if (t < 32) {
    float s = b1[t];
    for (int k = 0; k < 64; ++k) s = fmaf(w1[t * 64 + k], h0[k], s);
    h1[t] = fmaxf(s, 0.f);
  }
  __syncthreads();
  if (t == 0) {
    float s = b2[0];
    for (int k = 0; k < 32; ++k) s = fmaf(w2[k], h1[k], s);
    out[n] = s;
  }
}

int g_bilinear_up = 1;      // 1 = source-cell kernel for >= 1.5x bf16 upscales (default), 0 = generic kernel everywhere (A/B)

}  // namespace

/* 1 = source-cell kernel for >= 1.5x bf16 upscales (default); 0 = generic per-pixel kernel everywhere (A/B runs). */
extern "C" int add_bilinear_set_mode(int mode) { g_bilinear_up = mode ? 1 : 0; return ADD_OK; }

extern "C" int add_bilinear_fwd(const add_tensor_t* x, const add_tensor_t* y, uint32_t flags, void* stream) {
  ADD_CHECK_ARG(tensor_ok(x) && tensor_ok(y));
  ADD_CHECK_ARG(x->n == y->n && x->c == y->c);
  ADD_CHECK_SUP(tensor_vec4_ok(x) && tensor_vec4_ok(y));
  float sh = (float)x->h / (float)y->h, sw = (float)x->w / (float)y->w;
  ADD_CHECK_SUP((long long)y->h * y->w * (y->c / 4) < (1ll << 31) && y->n < 65536);
  const bool v8 = x->dtype == ADD_BF16 && y->dtype == ADD_BF16 && x->c % 8 == 0 && x->pix_stride % 8 == 0 &&
                  y->pix_stride % 8 == 0 && ((uintptr_t)x->ptr % 16) == 0 && ((uintptr_t)y->ptr % 16) == 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // upscales by >= 1.5x in both directions (bf16, 16-byte vectors): one thread per source cell
  if (v8 && g_bilinear_up && 2 * y->h >= 3 * x->h && 2 * y->w >= 3 * x->w) {
    const long long tot = (long long)x->h * x->w * (x->c / 8);
    long long gb = (tot + 255) / 256;
    const long long capu = (148ll * 16 + y->n - 1) / y->n;
    if (gb > capu) gb = capu;
    dim3 gridu((unsigned)gb, (unsigned)y->n);
    bilinear_up_kernel<<<gridu, 256, 0, s>>>((const bf16*)x->ptr, (bf16*)y->ptr, x->h, x->w, x->c, x->pix_stride, y->h, y->w,
                                            y->pix_stride, sh, sw, flags);
    ADD_RETURN_LAUNCH();
  }
  const long long total = (long long)y->h * y->w * (y->c / (v8 ? 8 : 4));
  long long bx = (total + 255) / 256;
  const long long cap = (148ll * 32 + y->n - 1) / y->n;
  if (bx > cap) bx = cap;
  dim3 grid((unsigned)bx, (unsigned)y->n);
#define BL(TI, TO, V) bilinear_kernel<TI, TO, V><<<grid, 256, 0, s>>>((const TI*)x->ptr, (TO*)y->ptr, x->h, \
    x->w, x->c, x->pix_stride, y->h, y->w, y->pix_stride, sh, sw, flags)
  if (v8) BL(bf16, bf16, 8);
  else if (x->dtype == ADD_F32 && y->dtype == ADD_F32) BL(float, float, 4);
  else if (x->dtype == ADD_BF16 && y->dtype == ADD_BF16) BL(bf16, bf16, 4);
  else if (x->dtype == ADD_F32 && y->dtype == ADD_BF16) BL(float, bf16, 4);
  else if (x->dtype == ADD_BF16 && y->dtype == ADD_F32) BL(bf16, float, 4);
  else return ADD_ERR_UNSUPPORTED;
#undef BL
  ADD_RETURN_LAUNCH();
}

extern "C" int add_nchw_to_nhwc(const float* src, int c_src, const add_tensor_t* y, void* stream) {
  ADD_CHECK_ARG(src && tensor_ok(y) && c_src > 0 && c_src <= y->c);
  int HW = y->h * y->w;
  dim3 grid(ceil_div(HW, 32), ceil_div(y->c, 32), y->n);
  ADD_CHECK_SUP(grid.y < 65536 && grid.z < 65536);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (y->dtype == ADD_F32)
    nchw_to_nhwc_kernel<float><<<grid, 256, 0, s>>>(src, (float*)y->ptr, c_src, HW, y->c, y->pix_stride);
  else
    nchw_to_nhwc_kernel<bf16><<<grid, 256, 0, s>>>(src, (bf16*)y->ptr, c_src, HW, y->c, y->pix_stride);
  ADD_RETURN_LAUNCH();
}

extern "C" int add_nhwc_to_nchw(const add_tensor_t* x, float* dst, void* stream) {
  ADD_CHECK_ARG(dst && tensor_ok(x));
  int HW = x->h * x->w;
  dim3 grid(ceil_div(HW, 32), ceil_div(x->c, 32), x->n);
  ADD_CHECK_SUP(grid.y < 65536 && grid.z < 65536);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (x->dtype == ADD_F32)
    nhwc_to_nchw_kernel<float><<<grid, 256, 0, s>>>((const float*)x->ptr, dst, x->c, HW, x->pix_stride);
  else
    nhwc_to_nchw_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)x->ptr, dst, x->c, HW, x->pix_stride);
  ADD_RETURN_LAUNCH();
}

extern "C" int64_t add_global_avgpool_workspace_bytes(int n, int h, int w, int c) {
  if (n <= 0 || h <= 0 || w <= 0 || c <= 0) return ADD_ERR_BAD_ARG;
  return (int64_t)n * gap_splits(h * w, n) * c * sizeof(float);
}

extern "C" int add_global_avgpool_fwd(const add_tensor_t* x, float* out, uint32_t flags, void* workspace,
                                      int64_t workspace_bytes, void* stream) {
  ADD_CHECK_ARG(tensor_ok(x) && out && workspace);
  const bool v8 = x->dtype == ADD_BF16 && x->c % 8 == 0 && x->pix_stride % 8 == 0 && ((uintptr_t)x->ptr % 16) == 0;
  ADD_CHECK_SUP(tensor_vec4_ok(x));
  if (workspace_bytes < add_global_avgpool_workspace_bytes(x->n, x->h, x->w, x->c)) return ADD_ERR_WORKSPACE;
  const int HW = x->h * x->w, S = gap_splits(HW, x->n);
  const int V = v8 ? 8 : 4;
  dim3 grid(S, x->n, ceil_div(x->c, GAP_THREADS * V));          // z: channel groups of GAP_THREADS * V channels
  ADD_CHECK_SUP(grid.z < 65536);
  size_t smem = (size_t)GAP_THREADS * V * sizeof(float);       // lanes * Cg <= GAP_THREADS * V
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (v8)
    gap_partial_kernel<bf16, 8><<<grid, GAP_THREADS, smem, s>>>((const bf16*)x->ptr, (float*)workspace, HW, x->c, x->pix_stride, flags);
  else if (x->dtype == ADD_F32)
    gap_partial_kernel<float, 4><<<grid, GAP_THREADS, smem, s>>>((const float*)x->ptr, (float*)workspace, HW, x->c, x->pix_stride, flags);
  else
    gap_partial_kernel<bf16, 4><<<grid, GAP_THREADS, smem, s>>>((const bf16*)x->ptr, (float*)workspace, HW, x->c, x->pix_stride, flags);
  gap_finalize_kernel<<<x->n, GAP_THREADS, 0, s>>>((const float*)workspace, out, S, x->c, 1.f / (float)HW);
  ADD_RETURN_LAUNCH();
}

extern "C" int add_edm_mlp_fwd(const float* pooled, int n, const float* w0, const float* b0, const float* w1,
                               const float* b1, const float* w2, const float* b2, float* out, void* stream) {
  ADD_CHECK_ARG(pooled && w0 && b0 && w1 && b1 && w2 && b2 && out && n > 0);
  edm_mlp_kernel<<<n, 128, 0, static_cast<cudaStream_t>(stream)>>>(pooled, w0, b0, w1, b1, w2, b2, out);
  ADD_RETURN_LAUNCH();
}

// ---- ASPP image-pool branch folded into a per-image bias (aspp_train.py:49-57) -------------------------------
// The pooled branch is constant over the image, so after the 1x1 over the concatenation it is just a per-image
// vector: bias_n[co] = b_out[co] + sum_d Wout_pool[d][co] * relu(b5[d] + sum_ci W5[ci][d] * pooled[n][ci]).
// One block per image, fp32 throughout; replaces the GAP->1x1->broadcast->concat-slice round trip (268 MB written
// and re-read per 4 images at 256x512).
namespace {
constexpr int PB_THREADS = 1024, PB_KSPLIT = 4;
// thread (o, q): output o (< 256), reduction slice q (k = q, q+4, ...); partial sums meet in shared memory in
// fixed order (deterministic)
__device__ __forceinline__ void pb_gemv(const float* __restrict__ w, const float* __restrict__ bias, const float* vin, int kdim,
                                        int odim, float* part, float* vout, bool relu) {
  for (int o0 = 0; o0 < odim; o0 += PB_THREADS / PB_KSPLIT) {
    const int o = o0 + (threadIdx.x % (PB_THREADS / PB_KSPLIT)), q = threadIdx.x / (PB_THREADS / PB_KSPLIT);
    float s = 0.f;
    if (o < odim) {
#pragma unroll 8
      for (int k = q; k < kdim; k += PB_KSPLIT) s = fmaf(__ldg(w + (size_t)k * odim + o), vin[k], s);
    }
    part[threadIdx.x] = s;
    __syncthreads();
    if (q == 0 && o < odim) {
      float t = bias ? bias[o] : 0.f;
#pragma unroll
      for (int j = 0; j < PB_KSPLIT; ++j) t += part[j * (PB_THREADS / PB_KSPLIT) + (threadIdx.x % (PB_THREADS / PB_KSPLIT))];
      vout[o] = relu ? fmaxf(t, 0.f) : t;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(PB_THREADS)
aspp_pool_bias_kernel(const float* __restrict__ pooled, int cin, const float* __restrict__ w5, const float* __restrict__ b5,
                      int depth, const float* __restrict__ w_out_pool, const float* __restrict__ b_out, int cout,
                      float* __restrict__ bias_out) {
  extern __shared__ float sm[];                 // [cin] pooled, [depth] p5, [PB_THREADS] partials
  float* pin = sm; float* p5 = sm + cin; float* part = p5 + depth;
  const int n = blockIdx.x;
  for (int i = threadIdx.x; i < cin; i += PB_THREADS) pin[i] = pooled[(size_t)n * cin + i];
  __syncthreads();
  pb_gemv(w5, b5, pin, cin, depth, part, p5, true);
  pb_gemv(w_out_pool, b_out, p5, depth, cout, part, bias_out + (size_t)n * cout, false);
}
}  // namespace

extern "C" int add_aspp_pool_bias_fwd(const float* pooled, int n, int cin, const float* w5, const float* b5, int depth,
                                      const float* w_out_pool, const float* b_out, int cout, float* bias_out, void* stream) {
  ADD_CHECK_ARG(pooled && w5 && w_out_pool && bias_out && n > 0 && cin > 0 && depth > 0 && cout > 0);
  ADD_CHECK_SUP((size_t)(cin + depth + PB_THREADS) * sizeof(float) <= 48 * 1024);
  aspp_pool_bias_kernel<<<n, PB_THREADS, (size_t)(cin + depth + PB_THREADS) * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      pooled, cin, w5, b5, depth, w_out_pool, b_out, cout, bias_out);
  ADD_RETURN_LAUNCH();
}

// ---- label widening: uint8 labels (Cityscapes PNG depth; 255 = ignore) -> the int64 the Evaluator path consumes ----
// Lets the host ship 1 byte per pixel instead of 8 (utils/metrics.py:35 masks on 0 <= gt < num_class, so 255 stays ignored).
namespace {
__global__ void __launch_bounds__(256)
widen_labels_kernel(const uint8_t* __restrict__ src, long long* __restrict__ dst, long long n) {
  for (long long i = (blockIdx.x * 256ll + threadIdx.x) * 16; i < n; i += (long long)gridDim.x * 256 * 16) {
    if (i + 16 <= n) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + i));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        longlong2 o;
        o.x = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
        o.y = (w[(j + 1) >> 2] >> (8 * ((j + 1) & 3))) & 0xffu;
        *reinterpret_cast<longlong2*>(dst + i + j) = o;
      }
    } else {
      for (long long k = i; k < n; ++k) dst[k] = src[k];
    }
  }
}
}  // namespace

// ---- loader edge: uint8 HWC image -> normalised fp32 NCHW (dataloaders/custom_transforms.py:17-24 + :39) -------
// Cityscapes images are uint8 PNGs; the reference normalises them on the host (img /= 255.0 in float32, then
// img -= mean and img /= std, which numpy evaluates in float64 because mean / std are float64 arrays, rounding to
// float32 after each) and ships 12 bytes per pixel to the GPU.  Here the 3 bytes travel and the same arithmetic —
// same operation order and roundings, so bit-identical — runs on the device: 4x fewer PCIe / host-DRAM bytes.
namespace {
__global__ void __launch_bounds__(256)
normalize_u8_hwc_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, long long hw, double m0, double m1, double m2,
                        double s0, double s1, double s2) {
  const int n = blockIdx.y;
  const uint8_t* in = src + (size_t)n * hw * 3;
  float* out = dst + (size_t)n * hw * 3;
  for (long long p = blockIdx.x * 256ll + threadIdx.x; p < hw; p += (long long)gridDim.x * 256) {
    const float r = __fdiv_rn((float)in[3 * p], 255.0f), g = __fdiv_rn((float)in[3 * p + 1], 255.0f), b = __fdiv_rn((float)in[3 * p + 2], 255.0f);
    const float r1 = (float)((double)r - m0), g1 = (float)((double)g - m1), b1 = (float)((double)b - m2);
    out[p] = (float)((double)r1 / s0);
    out[hw + p] = (float)((double)g1 / s1);
    out[2 * hw + p] = (float)((double)b1 / s2);
  }
}
}  // namespace

extern "C" int add_normalize_u8_hwc_to_nchw(const uint8_t* src, float* dst, int n, int h, int w, double mean0, double mean1,
                                            double mean2, double std0, double std1, double std2, void* stream) {
  ADD_CHECK_ARG(src && dst && n > 0 && h > 0 && w > 0 && std0 != 0.0 && std1 != 0.0 && std2 != 0.0);
  ADD_CHECK_SUP(n < 65536);
  const long long hw = (long long)h * w;
  long long blocks = (hw + 255) / 256;
  const long long cap = (148ll * 16 + n - 1) / n;
  if (blocks > cap) blocks = cap;
  normalize_u8_hwc_kernel<<<dim3((unsigned)blocks, (unsigned)n), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, hw, mean0,
                                                                                                        mean1, mean2, std0, std1, std2);
  ADD_RETURN_LAUNCH();
}

extern "C" int add_widen_labels_u8(const uint8_t* src, int64_t* dst, int64_t n, void* stream) {
  ADD_CHECK_ARG(src && dst && n >= 0);
  ADD_CHECK_SUP(((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 16) == 0);
  if (n == 0) return ADD_OK;
  long long blocks = (n / 16 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  widen_labels_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, (long long*)dst, n);
  ADD_RETURN_LAUNCH();
}

// ---- image gather: dst[j] = src[idx[j]] for whole per-image slabs (early-exit batch compaction) ----
namespace {
template <typename V>
__global__ void __launch_bounds__(256)
gather_images_kernel(const V* __restrict__ src, V* __restrict__ dst, const int* __restrict__ idx,
                     long long vec_per_image) {
  const int j = blockIdx.y;
  const V* s = src + (size_t)idx[j] * vec_per_image;
  V* d = dst + (size_t)j * vec_per_image;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < vec_per_image; i += (long long)gridDim.x * 256)
    d[i] = __ldg(s + i);
}
template <typename V>
void launch_gather(const void* src, void* dst, const int32_t* idx, int count, int64_t bytes, cudaStream_t s) {
  long long vec = bytes / (long long)sizeof(V);
  long long bx = (vec + 255) / 256;
  if (bx > 148 * 4) bx = 148 * 4;
  dim3 grid((unsigned)bx, (unsigned)count);
  gather_images_kernel<V><<<grid, 256, 0, s>>>((const V*)src, (V*)dst, idx, vec);
}
}  // namespace

// channel-slice variant: dst[j][p][0..c) = src[idx[j]][p][0..c) for NHWC views (possibly slices of wider buffers) —
// gathers only the live channels (the 48-channel low-level slot of the decoder's 304-channel concat buffer)
namespace {
__global__ void __launch_bounds__(256)
gather_view_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, const int* __restrict__ idx, long long pixels,
                   int vec_per_pix, long long src_img_vec, long long dst_img_vec, int src_pix_vec, int dst_pix_vec) {
  const int j = blockIdx.y;
  const uint4* s = src + (size_t)idx[j] * src_img_vec;
  uint4* d = dst + (size_t)j * dst_img_vec;
  const long long total = pixels * vec_per_pix;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long p = i / vec_per_pix;
    const int v = (int)(i - p * vec_per_pix);
    d[p * dst_pix_vec + v] = __ldg(s + p * src_pix_vec + v);
  }
}
}  // namespace

extern "C" int add_gather_images_view(const add_tensor_t* src, const add_tensor_t* dst, const int32_t* idx_dev, void* stream) {
  ADD_CHECK_ARG(tensor_ok(src) && tensor_ok(dst) && idx_dev);
  ADD_CHECK_ARG(src->h == dst->h && src->w == dst->w && src->c == dst->c && src->dtype == dst->dtype);
  const size_t e = dtype_size(src->dtype);
  ADD_CHECK_SUP((src->c * e) % 16 == 0 && (src->pix_stride * e) % 16 == 0 && (dst->pix_stride * e) % 16 == 0 &&
                ((uintptr_t)src->ptr % 16) == 0 && ((uintptr_t)dst->ptr % 16) == 0 && dst->n < 65536);
  const long long pixels = (long long)src->h * src->w;
  const int vpp = (int)(src->c * e / 16);
  long long bx = (pixels * vpp + 255) / 256;
  if (bx > 148 * 4) bx = 148 * 4;
  gather_view_kernel<<<dim3((unsigned)bx, (unsigned)dst->n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (const uint4*)src->ptr, (uint4*)dst->ptr, idx_dev, pixels, vpp, pixels * (long long)(src->pix_stride * e / 16),
      pixels * (long long)(dst->pix_stride * e / 16), (int)(src->pix_stride * e / 16), (int)(dst->pix_stride * e / 16));
  ADD_RETURN_LAUNCH();
}

extern "C" int add_gather_images(const void* src, void* dst, const int32_t* idx_dev, int count,
                                 int64_t bytes_per_image, void* stream) {
  ADD_CHECK_ARG(src && dst && idx_dev && count > 0 && bytes_per_image > 0);
  ADD_CHECK_SUP(count < 65536);
  uintptr_t a = (uintptr_t)src | (uintptr_t)dst | (uintptr_t)bytes_per_image;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a % 16 == 0) launch_gather<uint4>(src, dst, idx_dev, count, bytes_per_image, s);
  else if (a % 8 == 0) launch_gather<uint2>(src, dst, idx_dev, count, bytes_per_image, s);
  else if (a % 4 == 0) launch_gather<uint32_t>(src, dst, idx_dev, count, bytes_per_image, s);
  else launch_gather<uint8_t>(src, dst, idx_dev, count, bytes_per_image, s);      // uint8 label slabs of odd-sized images
  ADD_RETURN_LAUNCH();
}
