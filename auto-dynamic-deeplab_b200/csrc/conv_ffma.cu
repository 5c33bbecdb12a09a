// conv_ffma.cu — CUDA-core (fp32 FFMA) implicit-GEMM convolution and the fused SepConv half.
// This is the exact-fp32 parity path (and the small-shape path); the bf16 tensor-core path lives
// in conv_tc.cu.  NHWC activations, fp32 weights [kh][kw][Cin][Cout] with BN scale folded.
#include "common.cuh"

namespace {

struct ConvParams {
  const void* x; void* y; const float* w; const float* bias;
  long long bias_img_stride;   // floats between images' bias vectors (0 = shared)
  int N, H, W, Cin, xs;        // input view
  int Ho, Wo, Cout, ys;        // output view
  int kh, kw, stride, pad, dil;
  uint32_t flags;
  int M;                       // N*Ho*Wo
};

constexpr int CV_BM = 256;     // output pixels per block
constexpr int CV_BK = 16;      // reduction slice
constexpr int CV_TM = 8;       // pixels per thread
constexpr int CV_LDA = CV_BM + 4;

// Block: 256 threads = 32 pixel-groups (ty) x 8 cout-lanes (tx).  Thread computes CV_TM pixels
// (ty*8..ty*8+7) x TN couts (n0 + j*8 + tx).
template <typename TI, typename TO, int TN>
__global__ void __launch_bounds__(256)
conv2d_ffma_kernel(const ConvParams p) {
  constexpr int BN = 8 * TN;
  __shared__ __align__(16) float As[CV_BK][CV_LDA];
  __shared__ __align__(16) float Bs[CV_BK][BN];

  const int t = threadIdx.x;
  const int tx = t & 7, ty = t >> 3;
  const int m0 = blockIdx.x * CV_BM;
  const int n0 = blockIdx.y * BN;
  const TI* __restrict__ x = static_cast<const TI*>(p.x);
  const bool relu_in = p.flags & ADD_RELU_IN;

  // A-load mapping: 4 rounds; pixel = (t>>2) + 64*r, 4-channel chunk = t&3
  const int chunk = t & 3;
  int a_iy0[4], a_ix0[4];
  long long a_nbase[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int m = m0 + (t >> 2) + 64 * r;
    if (m < p.M) {
      int ox = m % p.Wo; int tmp = m / p.Wo; int oy = tmp % p.Ho; int n = tmp / p.Ho;
      a_iy0[r] = oy * p.stride - p.pad;
      a_ix0[r] = ox * p.stride - p.pad;
      a_nbase[r] = (long long)n * p.H * p.W;
    } else {
      a_iy0[r] = -(1 << 28); a_ix0[r] = 0; a_nbase[r] = 0;
    }
  }

  float acc[CV_TM][TN];
#pragma unroll
  for (int i = 0; i < CV_TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int ky = 0; ky < p.kh; ++ky) {
    for (int kx = 0; kx < p.kw; ++kx) {
      const float* __restrict__ wt = p.w + (size_t)(ky * p.kw + kx) * p.Cin * p.Cout;
      for (int c0 = 0; c0 < p.Cin; c0 += CV_BK) {
        // ---- load A tile (pixels x 16 channels of this tap) ----
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          int iy = a_iy0[r] + ky * p.dil, ix = a_ix0[r] + kx * p.dil;
          int c = c0 + chunk * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W && c < p.Cin) {
            v = ld4(x + (a_nbase[r] + (long long)iy * p.W + ix) * p.xs + c);
            if (relu_in) v = relu4(v);
          }
          int pm = (t >> 2) + 64 * r;
          As[chunk * 4 + 0][pm] = v.x; As[chunk * 4 + 1][pm] = v.y;
          As[chunk * 4 + 2][pm] = v.z; As[chunk * 4 + 3][pm] = v.w;
        }
        // ---- load B tile (16 cin rows x BN couts) ----
        for (int e = t; e < CV_BK * BN; e += 256) {
          int k = e / BN, n = e % BN;
          int c = c0 + k, co = n0 + n;
          Bs[k][n] = (c < p.Cin && co < p.Cout) ? __ldg(wt + (size_t)c * p.Cout + co) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CV_BK; ++k) {
          float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * CV_TM]);
          float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * CV_TM + 4]);
          float a[CV_TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
          float b[TN];
#pragma unroll
          for (int j = 0; j < TN; ++j) b[j] = Bs[k][j * 8 + tx];
#pragma unroll
          for (int i = 0; i < CV_TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
      }
    }
  }

  // ---- epilogue: + bias, (+= y), ReLU, store ----
  TO* __restrict__ y = static_cast<TO*>(p.y);
  const bool relu_out = p.flags & ADD_RELU_OUT, accum = p.flags & ADD_ACCUMULATE;
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    int co = n0 + j * 8 + tx;
    if (co >= p.Cout) continue;
    const float bv0 = (p.bias && p.bias_img_stride == 0) ? __ldg(p.bias + co) : 0.f;
#pragma unroll
    for (int i = 0; i < CV_TM; ++i) {
      int m = m0 + ty * CV_TM + i;
      if (m >= p.M) continue;
      TO* dst = y + (size_t)m * p.ys + co;
      const float bv = (p.bias && p.bias_img_stride != 0) ? __ldg(p.bias + (long long)(m / (p.Ho * p.Wo)) * p.bias_img_stride + co) : bv0;
      float v = acc[i][j] + bv;
      if (accum) v += ld1(const_cast<const TO*>(dst));
      if (relu_out) v = fmaxf(v, 0.f);
      st1(dst, v);
    }
  }
}

template <typename TI, typename TO>
int launch_conv(const ConvParams& p, cudaStream_t s) {
  // pick TN so that 8*TN covers Cout with the least padding waste
  int cout = p.Cout;
  int best_tn = 8; double best_cost = 1e30;
  const int cands[] = {3, 5, 6, 8, 10};
  for (int tn : cands) {
    int bn = 8 * tn;
    int tiles = ceil_div(cout, bn);
    double cost = (double)tiles * bn + 2.0 * tiles;  // padded width + small per-tile overhead
    if (cost < best_cost) { best_cost = cost; best_tn = tn; }
  }
  dim3 grid(ceil_div(p.M, CV_BM), 1), block(256);
#define LAUNCH_TN(TNV) case TNV: grid.y = ceil_div(cout, 8 * TNV); \
    conv2d_ffma_kernel<TI, TO, TNV><<<grid, block, 0, s>>>(p); break;
  switch (best_tn) {
    LAUNCH_TN(3) LAUNCH_TN(5) LAUNCH_TN(6) LAUNCH_TN(8) LAUNCH_TN(10)
    default: return ADD_ERR_UNSUPPORTED;
  }
#undef LAUNCH_TN
  ADD_RETURN_LAUNCH();
}

// ----------------------------------------------------------------------------------------------
// SepConv half: ReLU -> depthwise KxK (stride 1, pad K/2) -> pointwise 1x1 -> +bias (-> ReLU)
// ----------------------------------------------------------------------------------------------
struct SepParams {
  const void* x; void* y; const float* w_dw; const float* w_pw; const float* bias;
  int N, H, W, C, xs, Cout, ys;
  uint32_t flags;
  int M;
};

constexpr int SP_BM = 64;
constexpr int SP_LDD = SP_BM + 4;
constexpr int SP_BK = 16;

// dynamic smem: D[C][SP_LDD] floats + Ws[SP_BK][16*TN] floats
template <typename TI, typename TO, int K, int TN>
__global__ void __launch_bounds__(256)
sepconv_half_kernel(const SepParams p) {
  extern __shared__ __align__(16) float smem[];
  float* D = smem;                                  // [C][SP_LDD]
  float* Ws = smem + (size_t)p.C * SP_LDD;          // [SP_BK][16*TN]
  constexpr int BN = 16 * TN;
  const int t = threadIdx.x;
  const int m0 = blockIdx.x * SP_BM;
  const TI* __restrict__ x = static_cast<const TI*>(p.x);
  const bool relu_in = p.flags & ADD_RELU_IN;
  const int cv = p.C >> 2;                          // 4-channel vectors per pixel

  // ---- phase 1: depthwise into smem D[c][pixel] ----
  for (int item = t; item < SP_BM * cv; item += 256) {
    int pm = item / cv, c = (item % cv) * 4;
    int m = m0 + pm;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < p.M) {
      int ox = m % p.W; int tmp = m / p.W; int oy = tmp % p.H; int n = tmp / p.H;
      const TI* xn = x + (size_t)n * p.H * p.W * p.xs + c;
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        int iy = oy + ky - K / 2;
        if (iy < 0 || iy >= p.H) continue;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          int ix = ox + kx - K / 2;
          if (ix < 0 || ix >= p.W) continue;
          float4 v = ld4(xn + ((size_t)iy * p.W + ix) * p.xs);
          if (relu_in) v = relu4(v);
          float4 wv = __ldg(reinterpret_cast<const float4*>(p.w_dw + (size_t)(ky * K + kx) * p.C + c));
          a.x = fmaf(v.x, wv.x, a.x); a.y = fmaf(v.y, wv.y, a.y);
          a.z = fmaf(v.z, wv.z, a.z); a.w = fmaf(v.w, wv.w, a.w);
        }
      }
    }
    D[(c + 0) * SP_LDD + pm] = a.x; D[(c + 1) * SP_LDD + pm] = a.y;
    D[(c + 2) * SP_LDD + pm] = a.z; D[(c + 3) * SP_LDD + pm] = a.w;
  }

  // ---- phase 2: pointwise GEMM  out[pm][co] = sum_c D[c][pm] * Wpw[c][co] ----
  const int tx = t & 15, ty = t >> 4;               // 16 cout lanes x 16 pixel groups of 4
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int c0 = 0; c0 < p.C; c0 += SP_BK) {
    __syncthreads();   // D complete (first iter) / Ws consumed (later iters)
    for (int e = t; e < SP_BK * BN; e += 256) {
      int k = e / BN, n = e % BN;
      int c = c0 + k;
      Ws[e] = (c < p.C && n < p.Cout) ? __ldg(p.w_pw + (size_t)c * p.Cout + n) : 0.f;
    }
    __syncthreads();
    int kmax = min(SP_BK, p.C - c0);
    for (int k = 0; k < kmax; ++k) {
      float4 a4 = *reinterpret_cast<const float4*>(&D[(c0 + k) * SP_LDD + ty * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        float b = Ws[k * BN + j * 16 + tx];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][j] = fmaf(a[i], b, acc[i][j]);
      }
    }
  }

  TO* __restrict__ y = static_cast<TO*>(p.y);
  const bool relu_out = p.flags & ADD_RELU_OUT, accum = p.flags & ADD_ACCUMULATE;
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    int co = j * 16 + tx;
    if (co >= p.Cout) continue;
    float bv = p.bias ? __ldg(p.bias + co) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int m = m0 + ty * 4 + i;
      if (m >= p.M) continue;
      TO* dst = y + (size_t)m * p.ys + co;
      float v = acc[i][j] + bv;
      if (accum) v += ld1(const_cast<const TO*>(dst));
      if (relu_out) v = fmaxf(v, 0.f);
      st1(dst, v);
    }
  }
}

template <typename TI, typename TO, int K>
int launch_sep(const SepParams& p, cudaStream_t s) {
  int tn = ceil_div(p.Cout, 16);
  size_t smem = ((size_t)p.C * SP_LDD + (size_t)SP_BK * 16 * tn) * sizeof(float);
  dim3 grid(ceil_div(p.M, SP_BM)), block(256);
#define LAUNCH_TN(TNV) case TNV: { \
    auto kfn = sepconv_half_kernel<TI, TO, K, TNV>; \
    if (smem > 48 * 1024) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    kfn<<<grid, block, smem, s>>>(p); } break;
  switch (tn) {
    LAUNCH_TN(1) LAUNCH_TN(2) LAUNCH_TN(3) LAUNCH_TN(4) LAUNCH_TN(5) LAUNCH_TN(6)
    LAUNCH_TN(8) LAUNCH_TN(10)
    default: return ADD_ERR_UNSUPPORTED;
  }
#undef LAUNCH_TN
  ADD_RETURN_LAUNCH();
}

}  // namespace

extern "C" int add_conv2d_fwd(const add_tensor_t* x, const add_tensor_t* y, const float* w,
                              const float* bias, int64_t bias_image_stride, int kh, int kw, int stride, int pad, int dil,
                              uint32_t flags, void* stream) {
  ADD_CHECK_ARG(tensor_ok(x) && tensor_ok(y) && w);
  ADD_CHECK_ARG(kh > 0 && kw > 0 && stride > 0 && dil > 0);
  ADD_CHECK_ARG(x->n == y->n);
  // `pad` is the top/left padding (may be negative); the output extent is the caller's: taps that
  // fall outside the image read zero, so bottom/right padding is implicit.  Only require that the
  // last output's tap window still touches the image.
  ADD_CHECK_ARG(conv_extent_ok(x->h, y->h, kh, stride, pad, dil) && conv_extent_ok(x->w, y->w, kw, stride, pad, dil));
  // 4-channel vector loads on the input side
  ADD_CHECK_SUP(tensor_vec4_ok(x));
  ADD_CHECK_SUP((long long)y->n * y->h * y->w < (1ll << 31));
  ConvParams p;
  p.x = x->ptr; p.y = y->ptr; p.w = w; p.bias = bias; p.bias_img_stride = bias ? bias_image_stride : 0;
  p.N = x->n; p.H = x->h; p.W = x->w; p.Cin = x->c; p.xs = x->pix_stride;
  p.Ho = y->h; p.Wo = y->w; p.Cout = y->c; p.ys = y->pix_stride;
  p.kh = kh; p.kw = kw; p.stride = stride; p.pad = pad; p.dil = dil; p.flags = flags;
  p.M = y->n * y->h * y->w;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (x->dtype == ADD_F32 && y->dtype == ADD_F32) return launch_conv<float, float>(p, s);
  if (x->dtype == ADD_BF16 && y->dtype == ADD_BF16) return launch_conv<bf16, bf16>(p, s);
  if (x->dtype == ADD_BF16 && y->dtype == ADD_F32) return launch_conv<bf16, float>(p, s);
  if (x->dtype == ADD_F32 && y->dtype == ADD_BF16) return launch_conv<float, bf16>(p, s);
  return ADD_ERR_UNSUPPORTED;
}

extern "C" int add_sepconv_half_fwd(const add_tensor_t* x, const add_tensor_t* y, const float* w_dw,
                                    const float* w_pw, const float* bias, int k, uint32_t flags,
                                    void* stream) {
  ADD_CHECK_ARG(tensor_ok(x) && tensor_ok(y) && w_dw && w_pw);
  ADD_CHECK_ARG(x->n == y->n && x->h == y->h && x->w == y->w);
  ADD_CHECK_SUP(k == 3 || k == 5);
  ADD_CHECK_SUP(tensor_vec4_ok(x) && x->c <= 320 && y->c <= 160);
  ADD_CHECK_SUP(((uintptr_t)w_dw % 16) == 0);
  ADD_CHECK_SUP((long long)y->n * y->h * y->w < (1ll << 31));
  SepParams p;
  p.x = x->ptr; p.y = y->ptr; p.w_dw = w_dw; p.w_pw = w_pw; p.bias = bias;
  p.N = x->n; p.H = x->h; p.W = x->w; p.C = x->c; p.xs = x->pix_stride;
  p.Cout = y->c; p.ys = y->pix_stride; p.flags = flags; p.M = x->n * x->h * x->w;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define DISPATCH(TI, TO) (k == 3 ? launch_sep<TI, TO, 3>(p, s) : launch_sep<TI, TO, 5>(p, s))
  if (x->dtype == ADD_F32 && y->dtype == ADD_F32) return DISPATCH(float, float);
  if (x->dtype == ADD_BF16 && y->dtype == ADD_BF16) return DISPATCH(bf16, bf16);
#undef DISPATCH
  return ADD_ERR_UNSUPPORTED;
}
