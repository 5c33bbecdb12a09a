// head.cu — exit-head tail and Evaluator: final bilinear upsample fused with argmax / confusion
// histogram / entropy, the standalone int64 confusion matrix, and the confidence scalars.
// All HBM-bound integer/elementwise work.  The histogram is atomics-free: per-warp privatised
// shared-memory counters, intra-warp collisions resolved with match.any (leader adds popcount),
// deterministic block merge, per-block partials reduced by a second deterministic kernel.
#include "common.cuh"

namespace {

constexpr int HD_THREADS = 256;
constexpr int HD_WARPS = HD_THREADS / 32;
constexpr int MAX_CLASS = 32;               // num_class <= 32 (Cityscapes: 19)
constexpr int MAX_BINS = 361 + 7;           // supports num_class <= 19 for the histogram

// blocks per image: enough 2048-pixel chunks to fill the GPU (n*B ~ 4 CTAs per SM), few enough that the
// deterministic second-stage merge of the per-block histograms stays short
__host__ __device__ inline int head_blocks_per_image(long long hw, int n) {
  long long b = (hw + 2047) / 2048;
  long long cap = (148 * 4 + n - 1) / n;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

// One warp-synchronous histogram update.  `bin` < 0 means "no sample".  All 32 lanes must call.
__device__ __forceinline__ void warp_hist_add(unsigned int* hist, int bin, int lane) {
  unsigned peers = __match_any_sync(0xffffffffu, bin);
  if (bin >= 0 && lane == (__ffs(peers) - 1)) hist[bin] += __popc(peers);
  __syncwarp();
}

struct HeadParams {
  const float* x; int n, h, w, c, xs;      // low-res logits (fp32 NHWC)
  int H, W; float sh, sw;
  const long long* gt; const uint8_t* gt8; long long* pred;      // labels: int64 (metrics.py:34-39) or uint8 (the PNG bytes, 255 = ignore)
  unsigned int* part_hist;                  // [n][B][bins]
  double* part_ent;                         // [n][B]
  int bins; int B; int want_ent;
};

__global__ void __launch_bounds__(HD_THREADS)
upsample_argmax_kernel(const HeadParams p) {
  __shared__ unsigned int hist[HD_WARPS][MAX_BINS];
  __shared__ double ent_red[HD_WARPS];
  const int n = blockIdx.y, b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool do_hist = p.part_hist != nullptr;
  if (do_hist)
    for (int i = threadIdx.x; i < HD_WARPS * MAX_BINS; i += HD_THREADS) (&hist[0][0])[i] = 0u;
  __syncthreads();
  const long long HW = (long long)p.H * p.W;
  const long long per = (HW + p.B - 1) / p.B;
  const long long start = b * per, end = (start + per < HW) ? start + per : HW;
  const float* xn = p.x + (size_t)n * p.h * p.w * p.xs;
  const float inv_logc = 1.f / logf((float)p.c);
  double ent = 0.0;
  // uniform trip count per warp so match.any sees all 32 lanes
  for (long long base = start + warp * 32; base < end; base += HD_THREADS) {
    long long pix = base + lane;
    int bin = -1;
    if (pix < end) {
      int ox = (int)(pix % p.W), oy = (int)(pix / p.W);
      int y0, y1, x0, x1; float hl0, hl1, wl0, wl1;
      bilinear_src(oy, p.sh, p.h, y0, y1, hl0, hl1);
      bilinear_src(ox, p.sw, p.w, x0, x1, wl0, wl1);
      const float* p00 = xn + ((size_t)y0 * p.w + x0) * p.xs;
      const float* p01 = xn + ((size_t)y0 * p.w + x1) * p.xs;
      const float* p10 = xn + ((size_t)y1 * p.w + x0) * p.xs;
      const float* p11 = xn + ((size_t)y1 * p.w + x1) * p.xs;
      float v[MAX_CLASS];
      float best = -INFINITY; int arg = 0;
#pragma unroll
      for (int c = 0; c < MAX_CLASS; ++c) {
        if (c < p.c) {
          float val = hl0 * (wl0 * __ldg(p00 + c) + wl1 * __ldg(p01 + c)) +
                      hl1 * (wl0 * __ldg(p10 + c) + wl1 * __ldg(p11 + c));
          v[c] = val;
          if (val > best) { best = val; arg = c; }   // first maximum wins, like torch.argmax
        }
      }
      if (p.want_ent) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < MAX_CLASS; ++c) if (c < p.c) s += expf(v[c] - best);
        float logs = logf(s), e = 0.f;
#pragma unroll
        for (int c = 0; c < MAX_CLASS; ++c) if (c < p.c) {
          float lp = v[c] - best - logs;
          e += expf(lp) * lp;
        }
        ent += (double)(-e * inv_logc);
      }
      if (p.pred) p.pred[(size_t)n * HW + pix] = arg;
      if (do_hist && (p.gt || p.gt8)) {
        long long g = p.gt8 ? (long long)p.gt8[(size_t)n * HW + pix] : p.gt[(size_t)n * HW + pix];
        if (g >= 0 && g < p.c) bin = (int)g * p.c + arg;
      }
    }
    if (do_hist) warp_hist_add(hist[warp], bin, lane);
  }
  if (p.want_ent) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ent += __shfl_xor_sync(0xffffffffu, ent, o);
    if (lane == 0) ent_red[warp] = ent;
  }
  __syncthreads();
  if (do_hist) {
    unsigned int* out = p.part_hist + ((size_t)n * p.B + b) * p.bins;
    for (int i = threadIdx.x; i < p.bins; i += HD_THREADS) {
      unsigned int s = 0;
#pragma unroll
      for (int wv = 0; wv < HD_WARPS; ++wv) s += hist[wv][i];
      out[i] = s;
    }
  }
  if (p.want_ent && threadIdx.x == 0) {
    double s = 0.0;
    for (int wv = 0; wv < HD_WARPS; ++wv) s += ent_red[wv];
    p.part_ent[(size_t)n * p.B + b] = s;
  }
}

// Fast path for a compile-time class count: a thread owns a run of 8 consecutive output pixels of one row,
// [8r-4, 8r+4) — for the decoder's x8 upsample (decoder.py:28) the whole run interpolates between the SAME two
// source columns, so the 4 x NC source logits are loaded once per run and kept in registers (reloaded whenever
// the source columns change, so any size ratio is handled).  Same arithmetic order as upsample_logits_nchw.
template <int NC, bool WANT_ENT>
__global__ void __launch_bounds__(HD_THREADS)
upsample_argmax_runs_kernel(const HeadParams p) {
  __shared__ unsigned int hist[HD_WARPS][MAX_BINS];
  __shared__ double ent_red[HD_WARPS];
  const int n = blockIdx.y, b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool do_hist = p.part_hist != nullptr;
  if (do_hist)
    for (int i = threadIdx.x; i < HD_WARPS * MAX_BINS; i += HD_THREADS) (&hist[0][0])[i] = 0u;
  __syncthreads();
  const int runs_per_row = (p.W + 4 + 7) / 8;
  const long long n_runs = (long long)p.H * runs_per_row;
  const long long per = (n_runs + p.B - 1) / p.B;
  const long long start = b * per, end = (start + per < n_runs) ? start + per : n_runs;
  const long long HW = (long long)p.H * p.W;
  const float* xn = p.x + (size_t)n * p.h * p.w * p.xs;
  const float inv_logc = 1.f / logf((float)NC);
  constexpr int NCV = (NC + 3) / 4;
  const bool vec4 = (p.xs % 4 == 0) && (p.xs >= 4 * NCV) && (((uintptr_t)p.x & 15) == 0);
  const bool want_gt = do_hist && (p.gt || p.gt8);
  double ent = 0.0;
  for (long long base = start + warp * 32; base < end; base += HD_THREADS) {      // uniform trip count per warp
    const long long run = base + lane;
    const bool live = run < end;
    const int oy = live ? (int)(run / runs_per_row) : 0;
    const int ox_first = live ? ((int)(run % runs_per_row) * 8 - 4) : 0;
    const size_t row_pix = (size_t)n * HW + (size_t)oy * p.W;
    // the run's 8 labels first, all loads in flight together: inside the pixel loop each one was a dependent global
    // round trip in front of that pixel's histogram update (ncu r4: the kernel sat at 0.07 of HBM with ~half its issue
    // slots idle, stalled on these loads — not on the 19-class interpolation)
    // (labels outside [0, NC) -> 255 = no sample; the 8 of them packed into one 64-bit register)
    unsigned long long g8 = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ox = ox_first + j;
      long long g = -1;
      if (want_gt && live && ox >= 0 && ox < p.W)
        g = p.gt8 ? (long long)__ldg(p.gt8 + row_pix + ox) : __ldg(p.gt + row_pix + ox);
      g8 |= (unsigned long long)((g >= 0 && g < NC) ? (unsigned)g : 255u) << (8 * j);
    }
    int y0, y1; float hl0, hl1;
    bilinear_src(oy, p.sh, p.h, y0, y1, hl0, hl1);
    const float* row0 = xn + (size_t)y0 * p.w * p.xs;
    const float* row1 = xn + (size_t)y1 * p.w * p.xs;
    float v00[NC], v01[NC], v10[NC], v11[NC];
    int cx0 = -1, cx1 = -1;
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      const int ox = ox_first + j;
      int bin = -1;
      if (live && ox >= 0 && ox < p.W) {
        int x0, x1; float wl0, wl1;
        bilinear_src(ox, p.sw, p.w, x0, x1, wl0, wl1);
        if (x0 != cx0 || x1 != cx1) {
          const float* a0 = row0 + (size_t)x0 * p.xs; const float* a1 = row0 + (size_t)x1 * p.xs;
          const float* b0 = row1 + (size_t)x0 * p.xs; const float* b1 = row1 + (size_t)x1 * p.xs;
          if (vec4) {      // 16-byte loads: a pixel's NC logits + padding are NCV = ceil(NC / 4) float4 (pix_stride >= 4 * NCV)
#pragma unroll
            for (int c4 = 0; c4 < NCV; ++c4) {
              const float4 t0 = __ldg(reinterpret_cast<const float4*>(a0) + c4), t1 = __ldg(reinterpret_cast<const float4*>(a1) + c4);
              const float4 t2 = __ldg(reinterpret_cast<const float4*>(b0) + c4), t3 = __ldg(reinterpret_cast<const float4*>(b1) + c4);
              const float e0[4] = {t0.x, t0.y, t0.z, t0.w}, e1[4] = {t1.x, t1.y, t1.z, t1.w};
              const float e2[4] = {t2.x, t2.y, t2.z, t2.w}, e3[4] = {t3.x, t3.y, t3.z, t3.w};
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (4 * c4 + u < NC) { v00[4 * c4 + u] = e0[u]; v01[4 * c4 + u] = e1[u]; v10[4 * c4 + u] = e2[u]; v11[4 * c4 + u] = e3[u]; }
            }
          } else {
#pragma unroll
            for (int c = 0; c < NC; ++c) { v00[c] = __ldg(a0 + c); v01[c] = __ldg(a1 + c); v10[c] = __ldg(b0 + c); v11[c] = __ldg(b1 + c); }
          }
          cx0 = x0; cx1 = x1;
        }
        int arg = 0;
        {
          float v[NC];
          float best = -INFINITY;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const float val = hl0 * (wl0 * v00[c] + wl1 * v01[c]) + hl1 * (wl0 * v10[c] + wl1 * v11[c]);
            v[c] = val;
            if (val > best) { best = val; arg = c; }    // first maximum wins, like torch.argmax
          }
          if (WANT_ENT) {
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < NC; ++c) s += expf(v[c] - best);
            const float logs = logf(s);
            float e = 0.f;
#pragma unroll
            for (int c = 0; c < NC; ++c) { const float lp = v[c] - best - logs; e += expf(lp) * lp; }
            ent += (double)(-e * inv_logc);
          }
        }
        if (p.pred) p.pred[row_pix + ox] = arg;
        const int lab = (int)((g8 >> (8 * j)) & 0xffull);
        if (lab < NC) bin = lab * NC + arg;
      }
      if (do_hist) warp_hist_add(hist[warp], bin, lane);
    }
  }
  if (WANT_ENT) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ent += __shfl_xor_sync(0xffffffffu, ent, o);
    if (lane == 0) ent_red[warp] = ent;
  }
  __syncthreads();
  if (do_hist) {
    unsigned int* out = p.part_hist + ((size_t)n * p.B + b) * p.bins;
    for (int i = threadIdx.x; i < p.bins; i += HD_THREADS) {
      unsigned int s = 0;
#pragma unroll
      for (int wv = 0; wv < HD_WARPS; ++wv) s += hist[wv][i];
      out[i] = s;
    }
  }
  if (WANT_ENT && threadIdx.x == 0) {
    double s = 0.0;
    for (int wv = 0; wv < HD_WARPS; ++wv) s += ent_red[wv];
    p.part_ent[(size_t)n * p.B + b] = s;
  }
}

// deterministic second stage: one block per image
__global__ void __launch_bounds__(HD_THREADS)
head_finalize_kernel(const unsigned int* __restrict__ part_hist, const double* __restrict__ part_ent,
                     int B, int bins, long long* __restrict__ cm_out, float* __restrict__ ent_out,
                     double inv_hw, const int* __restrict__ cm_row) {
  // warp w sums the per-block histograms b = w, w+8, ... (coalesced rows, independent loads), then the eight
  // per-warp sums are added in fixed order: integer arithmetic, deterministic.
  __shared__ long long red[HD_WARPS][MAX_BINS];
  const int n = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (cm_out) {
    for (int i0 = 0; i0 < bins; i0 += 32 * 4) {
      long long s[4] = {0, 0, 0, 0};
      // four partial rows per trip: 16 independent loads in flight per thread (the merge is a chain of L2 round trips —
      // one row per trip made this kernel 40 us for 2 x 148 rows of 361 counters; integer sums, any order is exact)
      for (int b = warp; b < B; b += 4 * HD_WARPS) {
        unsigned int v[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int br = b + r * HD_WARPS;
          const unsigned int* row = part_hist + ((size_t)n * B + (br < B ? br : b)) * bins;
#pragma unroll
          for (int u = 0; u < 4; ++u) { const int i = i0 + u * 32 + lane; v[r][u] = (br < B && i < bins) ? __ldg(row + i) : 0u; }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int u = 0; u < 4; ++u) s[u] += v[r][u];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int i = i0 + u * 32 + lane; if (i < bins) red[warp][i] = s[u]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += HD_THREADS) {
      long long s = 0;
#pragma unroll
      for (int wv = 0; wv < HD_WARPS; ++wv) s += red[wv][i];
      cm_out[(size_t)(cm_row ? cm_row[n] : n) * bins + i] = s;   // cm_row: scatter to the image's ORIGINAL batch row
    }
  }
  if (ent_out && threadIdx.x == 0) {
    double s = 0.0;
    for (int b = 0; b < B; ++b) s += part_ent[(size_t)n * B + b];
    ent_out[n] = (float)(s * inv_hw);
  }
}

// ---- materialise NCHW fp32 logits at full resolution -------------------------------------------
__global__ void __launch_bounds__(256)
upsample_logits_nchw_kernel(const float* __restrict__ x, int n_img, int h, int w, int c, int xs,
                            float* __restrict__ dst, int H, int W, float sh, float sw) {
  long long HW = (long long)H * W, total = HW * n_img;
  for (long long idx = blockIdx.x * 256ll + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    int n = (int)(idx / HW); long long pix = idx % HW;
    int ox = (int)(pix % W), oy = (int)(pix / W);
    int y0, y1, x0, x1; float hl0, hl1, wl0, wl1;
    bilinear_src(oy, sh, h, y0, y1, hl0, hl1);
    bilinear_src(ox, sw, w, x0, x1, wl0, wl1);
    const float* xn = x + (size_t)n * h * w * xs;
    const float* p00 = xn + ((size_t)y0 * w + x0) * xs; const float* p01 = xn + ((size_t)y0 * w + x1) * xs;
    const float* p10 = xn + ((size_t)y1 * w + x0) * xs; const float* p11 = xn + ((size_t)y1 * w + x1) * xs;
    float* d = dst + (size_t)n * c * HW + pix;
    for (int ch = 0; ch < c; ++ch) {
      float val = hl0 * (wl0 * __ldg(p00 + ch) + wl1 * __ldg(p01 + ch)) +
                  hl1 * (wl0 * __ldg(p10 + ch) + wl1 * __ldg(p11 + ch));
      __stcs(d + (size_t)ch * HW, val);     // streaming store: 159 MB/image/exit never re-read here
    }
  }
}

// ---- Evaluator._generate_matrix: int64 gt/pred -> int64 [nc*nc] -------------------------------
__global__ void __launch_bounds__(HD_THREADS)
confusion_kernel(const long long* __restrict__ gt, const long long* __restrict__ pred, long long n_pix,
                 int nc, unsigned int* __restrict__ part, int vec_ok) {
  __shared__ unsigned int hist[HD_WARPS][MAX_BINS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < HD_WARPS * MAX_BINS; i += HD_THREADS) (&hist[0][0])[i] = 0u;
  __syncthreads();
  const long long per = ((n_pix + gridDim.x - 1) / gridDim.x + 1) & ~1ll;   // even -> 16 B aligned pairs
  const long long start = blockIdx.x * per, end = (start + per < n_pix) ? start + per : n_pix;
  // Two pixels per lane and step, 16-byte loads when both bases are 16-byte aligned (vec_ok), else two 8-byte loads
  // each (a per-image slice of an odd-sized map, e.g. target[i] of 1025x2049 labels, starts 8 bytes off).  UNR steps'
  // loads are issued before the first histogram update, so a warp keeps 4 KB in flight instead of 1 KB (r4: with one
  // step per round trip the kernel ran at 0.17 of HBM, stalled on these loads).
  constexpr int UNR = 4;
  for (long long base = start + warp * 64; base < end; base += (long long)HD_THREADS * 2 * UNR) {
    long long g[UNR][2], q[UNR][2];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long pix = base + (long long)u * HD_THREADS * 2 + lane * 2;
      g[u][0] = g[u][1] = -1; q[u][0] = q[u][1] = 0;
      if (vec_ok && pix + 1 < end) {
        const longlong2 gv = __ldcs(reinterpret_cast<const longlong2*>(gt + pix));
        const longlong2 qv = __ldcs(reinterpret_cast<const longlong2*>(pred + pix));
        g[u][0] = gv.x; g[u][1] = gv.y; q[u][0] = qv.x; q[u][1] = qv.y;
      } else {
        if (pix < end) { g[u][0] = __ldcs(gt + pix); q[u][0] = __ldcs(pred + pix); }
        if (pix + 1 < end) { g[u][1] = __ldcs(gt + pix + 1); q[u][1] = __ldcs(pred + pix + 1); }
      }
    }
    // a prediction outside [0,nc) has no cell in the matrix (the reference's bincount raises on a negative label and
    // fails its reshape on a too large one, metrics.py:37-38): such pixels are dropped, never aliased into a valid bin
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (base + (long long)u * HD_THREADS * 2 >= end) break;            // warp-uniform: no lane of this step is in range
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        int bin = -1;
        if (g[u][e] >= 0 && g[u][e] < nc && q[u][e] >= 0 && q[u][e] < nc) bin = (int)(g[u][e] * nc + q[u][e]);
        warp_hist_add(hist[warp], bin, lane);
      }
    }
  }
  __syncthreads();
  int bins = nc * nc;
  for (int i = threadIdx.x; i < bins; i += HD_THREADS) {
    unsigned int s = 0;
#pragma unroll
    for (int wv = 0; wv < HD_WARPS; ++wv) s += hist[wv][i];
    part[(size_t)blockIdx.x * bins + i] = s;
  }
}

__global__ void __launch_bounds__(HD_THREADS)
confusion_finalize_kernel(const unsigned int* __restrict__ part, int B, int bins, long long* __restrict__ out) {
  for (int i = threadIdx.x; i < bins; i += HD_THREADS) {
    long long s = 0;
    for (int b = 0; b < B; ++b) s += part[(size_t)b * bins + i];
    out[i] = s;
  }
}

// ---- Evaluator._generate_matrix, variant B (opt-in: add_confusion_set_impl(1)) ---------------------------------
// Thread-private histograms.  Variant A above privatises per WARP and resolves intra-warp collisions with match.any
// (ncu r4i: 114 us for 268 MB = DRAM 29 %, bound by the MATCH + leader-add chain per pixel and 2.5 M bank conflicts;
// its one-block finalize walks up to 1184 partial rows one dependent L2 round trip at a time and takes as long again).
// Here every THREAD owns a private histogram in shared memory, so an update is a plain load / add / store that no other
// thread can touch — no atomics, no match.any, no election.  Layout: word [bin pair p][thread t] holds bins 2p and 2p+1
// as two 16-bit halves -> the 32 lanes of a warp hit 32 different banks whatever their bins (bank = t % 32), and 256
// threads x <= 184 pairs x 4 B <= 184 KB fit one CTA per SM.  16-bit counters bound the pixels per thread: the grid is
// sized so that no thread sees more than CF_MAX_PER_THREAD.  The block merge reads the rows rotated by the pair index
// (conflict-free); the per-block partials are summed by a second kernel in fixed order with all loads of a thread
// independent.  Integer arithmetic throughout: bit-exact and deterministic, like variant A.
// Written after round 2's GPU budget was spent: OFF by default until tests/test_zz_gpu_late.py has pinned it bit-exact
// against variant A on a B200 (bench.py reports both, `candidates.confusion_matrix_private`).
constexpr int CF_THREADS = 256;
constexpr int CF_MAX_PAIRS = (MAX_BINS + 1) / 2;
constexpr long long CF_MAX_PER_THREAD = 60000;      // < 65536: a 16-bit private counter cannot wrap
constexpr int CF_UNR = 8;

__global__ void __launch_bounds__(CF_THREADS, 1)
confusion_private_kernel(const long long* __restrict__ gt, const long long* __restrict__ pred, long long n_pix,
                         int nc, unsigned int* __restrict__ part, int vec_ok) {
  extern __shared__ unsigned int cf_words[];          // [pairs][CF_THREADS]
  const int tid = threadIdx.x;
  const int bins = nc * nc, pairs = (bins + 1) >> 1;
  for (int p = 0; p < pairs; ++p) cf_words[p * CF_THREADS + tid] = 0u;   // own column only: no barrier needed before use
  const long long per = ((n_pix + gridDim.x - 1) / gridDim.x + 1) & ~1ll;   // even -> 16 B aligned pairs
  const long long start = blockIdx.x * per, end = (start + per < n_pix) ? start + per : n_pix;
  // Two pixels per thread and step (one 16-byte load of gt and one of pred when both bases are 16-byte aligned, else
  // 8-byte loads: a per-image slice of an odd-sized map starts 8 bytes off); CF_UNR steps' loads are issued before the
  // first update, so a warp keeps 8 KB in flight.
  for (long long base = start + (long long)tid * 2; base < end; base += (long long)CF_THREADS * 2 * CF_UNR) {
    long long g[CF_UNR][2], q[CF_UNR][2];
#pragma unroll
    for (int u = 0; u < CF_UNR; ++u) {
      const long long pix = base + (long long)u * CF_THREADS * 2;
      g[u][0] = g[u][1] = -1; q[u][0] = q[u][1] = 0;
      if (vec_ok && pix + 1 < end) {
        const longlong2 gv = __ldcs(reinterpret_cast<const longlong2*>(gt + pix));
        const longlong2 qv = __ldcs(reinterpret_cast<const longlong2*>(pred + pix));
        g[u][0] = gv.x; g[u][1] = gv.y; q[u][0] = qv.x; q[u][1] = qv.y;
      } else {
        if (pix < end) { g[u][0] = __ldcs(gt + pix); q[u][0] = __ldcs(pred + pix); }
        if (pix + 1 < end) { g[u][1] = __ldcs(gt + pix + 1); q[u][1] = __ldcs(pred + pix + 1); }
      }
    }
    // labels / predictions outside [0,nc) have no cell in the matrix (metrics.py:35 masks the labels; a prediction out of
    // range is dropped, never aliased into a valid bin): one unsigned compare each
#pragma unroll
    for (int u = 0; u < CF_UNR; ++u) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const unsigned long long gu = (unsigned long long)g[u][e], qu = (unsigned long long)q[u][e];
        if (gu < (unsigned long long)nc && qu < (unsigned long long)nc) {
          const int bin = (int)gu * nc + (int)qu;
          unsigned int* w = cf_words + (bin >> 1) * CF_THREADS + tid;
          *w += 1u << ((bin & 1) << 4);
        }
      }
    }
  }
  __syncthreads();
  // merge: thread p sums pair p over the 256 private columns, starting at column p (bank = (p + j) % 32: no conflicts),
  // low and high halves separately (<= 256 x 65535 each: fits 32 bits)
  for (int p = tid; p < pairs; p += CF_THREADS) {
    unsigned int lo = 0, hi = 0;
    for (int j = 0; j < CF_THREADS; ++j) {
      const unsigned int w = cf_words[p * CF_THREADS + ((j + p) & (CF_THREADS - 1))];
      lo += w & 0xffffu; hi += w >> 16;
    }
    part[(size_t)blockIdx.x * bins + 2 * p] = lo;
    if (2 * p + 1 < bins) part[(size_t)blockIdx.x * bins + 2 * p + 1] = hi;
  }
}

// Sum of the B per-block rows, int64 out.  Block j sums bins [128 j, 128 j + 128); its 32 warps take rows w, w + 32, ...
// four rows x four bins per thread and trip = 16 independent loads in flight (variant A's finalize walks the rows one
// dependent L2 round trip at a time in a single block); the warps meet in shared memory in fixed order.
constexpr int CFF_WARPS = 32;
__global__ void __launch_bounds__(CFF_WARPS * 32)
confusion_finalize_wide_kernel(const unsigned int* __restrict__ part, int B, int bins, long long* __restrict__ out) {
  __shared__ unsigned long long red[CFF_WARPS][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i0 = blockIdx.x * 128;
  unsigned long long s[4] = {0, 0, 0, 0};
  for (int b = warp; b < B; b += 4 * CFF_WARPS) {
    unsigned int v[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int br = b + r * CFF_WARPS;
      const unsigned int* row = part + (size_t)(br < B ? br : b) * bins;
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int i = i0 + u * 32 + lane; v[r][u] = (br < B && i < bins) ? __ldg(row + i) : 0u; }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int u = 0; u < 4; ++u) s[u] += v[r][u];
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) red[warp][u * 32 + lane] = s[u];
  __syncthreads();
  if (threadIdx.x < 128 && i0 + (int)threadIdx.x < bins) {
    unsigned long long t = 0;
#pragma unroll
    for (int wv = 0; wv < CFF_WARPS; ++wv) t += red[wv][threadIdx.x];
    out[i0 + threadIdx.x] = (long long)t;
  }
}

// ---- confidence scalars on NCHW fp32 logits -----------------------------------------------------
__global__ void __launch_bounds__(256)
confidence_kernel(const float* __restrict__ logits, int n_img, int c, long long HW, float thr,
                  double* __restrict__ part) {
  __shared__ double red[2][8];
  long long total = HW * n_img;
  double ent = 0.0, cnt = 0.0;
  const float inv_logc = 1.f / logf((float)c);
  for (long long idx = blockIdx.x * 256ll + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    int n = (int)(idx / HW); long long pix = idx % HW;
    const float* p = logits + (size_t)n * c * HW + pix;
    float v[MAX_CLASS]; float best = -INFINITY;
#pragma unroll
    for (int ch = 0; ch < MAX_CLASS; ++ch) if (ch < c) { v[ch] = __ldg(p + (size_t)ch * HW); best = fmaxf(best, v[ch]); }
    float s = 0.f;
#pragma unroll
    for (int ch = 0; ch < MAX_CLASS; ++ch) if (ch < c) s += expf(v[ch] - best);
    float logs = logf(s), e = 0.f;
#pragma unroll
    for (int ch = 0; ch < MAX_CLASS; ++ch) if (ch < c) { float lp = v[ch] - best - logs; e += expf(lp) * lp; }
    ent += (double)(-e * inv_logc);
    if (1.f / s > thr) cnt += 1.0;          // max softmax prob = exp(0)/s
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ent += __shfl_xor_sync(0xffffffffu, ent, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = ent; red[1][warp] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < 8; ++i) { a += red[0][i]; b += red[1][i]; }
    part[2 * blockIdx.x] = a; part[2 * blockIdx.x + 1] = b;
  }
}

__global__ void confidence_finalize_kernel(const double* __restrict__ part, int B, double inv_hw, float* out2) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < B; ++i) { a += part[2 * i]; b += part[2 * i + 1]; }
    out2[0] = (float)(a * inv_hw); out2[1] = (float)(b * inv_hw);
  }
}

inline int confusion_blocks(long long n_pix) {
  long long b = (n_pix + 4095) / 4096;
  return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}
// variant B: one CTA per SM (184 KB of private histograms each); more only when a thread would otherwise see more pixels
// than its 16-bit counters can hold
inline int confusion_private_blocks(long long n_pix) {
  long long b = (n_pix + 4095) / 4096;
  if (b > 148) b = 148;
  const long long need = (n_pix + CF_THREADS * CF_MAX_PER_THREAD - 1) / (CF_THREADS * CF_MAX_PER_THREAD);
  if (b < need) b = need;
  return (int)(b < 1 ? 1 : b);
}
int g_confusion_impl = 0;     // 0 = variant A (per-warp, match.any; the measured default), 1 = variant B (thread-private) +
                              // wide finalize, 2 = variant A's histogram kernel + the wide finalize
inline int confidence_blocks(long long total) {
  long long b = (total + 1023) / 1024;
  return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}

}  // namespace

extern "C" int add_upsample_logits_nchw(const add_tensor_t* x, float* dst, int H, int W, void* stream) {
  ADD_CHECK_ARG(tensor_ok(x) && dst && H > 0 && W > 0);
  ADD_CHECK_SUP(x->dtype == ADD_F32);
  long long total = (long long)x->n * H * W;
  int blocks = (int)((total + 255) / 256 < 148ll * 32 ? (total + 255) / 256 : 148ll * 32);
  upsample_logits_nchw_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (const float*)x->ptr, x->n, x->h, x->w, x->c, x->pix_stride, dst, H, W,
      (float)x->h / (float)H, (float)x->w / (float)W);
  ADD_RETURN_LAUNCH();
}

extern "C" int64_t add_head_workspace_bytes(int n, int H, int W, int num_class) {
  if (n <= 0 || H <= 0 || W <= 0 || num_class <= 0) return ADD_ERR_BAD_ARG;
  int B = head_blocks_per_image((long long)H * W, n);
  int64_t hist = (int64_t)n * B * num_class * num_class * sizeof(unsigned int);
  hist = (hist + 15) & ~15ll;
  return hist + (int64_t)n * B * sizeof(double);
}

static int upsample_argmax_impl(const add_tensor_t* x, int H, int W, const int64_t* gt, const uint8_t* gt8,
                                int64_t* pred_out, int64_t* cm_out, const int32_t* cm_row_index, float* entropy_out,
                                void* workspace, int64_t workspace_bytes, void* stream);

extern "C" int add_upsample_argmax_fwd(const add_tensor_t* x, int H, int W, const int64_t* gt,
                                       int64_t* pred_out, int64_t* cm_out, const int32_t* cm_row_index, float* entropy_out,
                                       void* workspace, int64_t workspace_bytes, void* stream) {
  return upsample_argmax_impl(x, H, W, gt, nullptr, pred_out, cm_out, cm_row_index, entropy_out, workspace, workspace_bytes, stream);
}

extern "C" int add_upsample_argmax_u8_fwd(const add_tensor_t* x, int H, int W, const uint8_t* gt_u8,
                                          int64_t* pred_out, int64_t* cm_out, const int32_t* cm_row_index, float* entropy_out,
                                          void* workspace, int64_t workspace_bytes, void* stream) {
  return upsample_argmax_impl(x, H, W, nullptr, gt_u8, pred_out, cm_out, cm_row_index, entropy_out, workspace, workspace_bytes, stream);
}

static int upsample_argmax_impl(const add_tensor_t* x, int H, int W, const int64_t* gt, const uint8_t* gt8,
                                int64_t* pred_out, int64_t* cm_out, const int32_t* cm_row_index, float* entropy_out,
                                void* workspace, int64_t workspace_bytes, void* stream) {
  ADD_CHECK_ARG(tensor_ok(x) && H > 0 && W > 0 && workspace);
  ADD_CHECK_ARG(!(cm_out && !gt && !gt8));
  ADD_CHECK_SUP(x->dtype == ADD_F32 && x->c <= MAX_CLASS);
  ADD_CHECK_SUP(!cm_out || x->c * x->c <= MAX_BINS);
  if (workspace_bytes < add_head_workspace_bytes(x->n, H, W, x->c)) return ADD_ERR_WORKSPACE;
  HeadParams p;
  p.x = (const float*)x->ptr; p.n = x->n; p.h = x->h; p.w = x->w; p.c = x->c; p.xs = x->pix_stride;
  p.H = H; p.W = W; p.sh = (float)x->h / (float)H; p.sw = (float)x->w / (float)W;
  p.gt = (const long long*)gt; p.gt8 = gt8; p.pred = (long long*)pred_out;
  p.bins = x->c * x->c; p.B = head_blocks_per_image((long long)H * W, x->n);
  int64_t hist_bytes = ((int64_t)x->n * p.B * p.bins * sizeof(unsigned int) + 15) & ~15ll;
  p.part_hist = cm_out ? (unsigned int*)workspace : nullptr;
  p.part_ent = (double*)((char*)workspace + hist_bytes);
  p.want_ent = entropy_out != nullptr;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  dim3 grid(p.B, x->n);
  if (x->c == 19) {            // Cityscapes: the specialised run kernel
    if (p.want_ent) upsample_argmax_runs_kernel<19, true><<<grid, HD_THREADS, 0, s>>>(p);
    else upsample_argmax_runs_kernel<19, false><<<grid, HD_THREADS, 0, s>>>(p);
  } else {
    upsample_argmax_kernel<<<grid, HD_THREADS, 0, s>>>(p);
  }
  if (cm_out || entropy_out)
    head_finalize_kernel<<<x->n, HD_THREADS, 0, s>>>(p.part_hist, p.part_ent, p.B, p.bins,
                                                      (long long*)cm_out, entropy_out, 1.0 / ((double)H * W), cm_row_index);
  ADD_RETURN_LAUNCH();
}

extern "C" int64_t add_confusion_workspace_bytes(int64_t n_pixels, int num_class) {
  if (n_pixels < 0 || num_class <= 0) return ADD_ERR_BAD_ARG;
  const int a = confusion_blocks(n_pixels), b = confusion_private_blocks(n_pixels);      // enough for either variant
  return (int64_t)(a > b ? a : b) * num_class * num_class * sizeof(unsigned int);
}

/* 0 = per-warp privatised histogram with match.any + one-block finalize (default), 1 = thread-private histograms + wide
 * finalize, 2 = the default histogram kernel + wide finalize (1 and 2 opt-in, see head.cu). */
extern "C" int add_confusion_set_impl(int impl) {
  if (impl < 0 || impl > 2) return ADD_ERR_BAD_ARG;
  g_confusion_impl = impl;
  return ADD_OK;
}

extern "C" int add_confusion_matrix(const int64_t* gt, const int64_t* pred, int64_t n_pixels, int num_class,
                                    int64_t* cm_out, void* workspace, int64_t workspace_bytes, void* stream) {
  ADD_CHECK_ARG(cm_out && workspace && n_pixels >= 0 && num_class > 0);
  ADD_CHECK_ARG(n_pixels == 0 || (gt && pred));
  ADD_CHECK_SUP(num_class * num_class <= MAX_BINS);
  ADD_CHECK_SUP(((uintptr_t)gt % 8 == 0) && ((uintptr_t)pred % 8 == 0));
  const int vec_ok = (((uintptr_t)gt % 16 == 0) && ((uintptr_t)pred % 16 == 0)) ? 1 : 0;
  if (workspace_bytes < add_confusion_workspace_bytes(n_pixels, num_class)) return ADD_ERR_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (g_confusion_impl == 1) {
    const int Bp = confusion_private_blocks(n_pixels);
    const size_t smem = (size_t)((num_class * num_class + 1) / 2) * CF_THREADS * sizeof(unsigned int);
    static PerDeviceOnce once;
    once_per_device(once, [] {
      cudaFuncSetAttribute(confusion_private_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           CF_MAX_PAIRS * CF_THREADS * (int)sizeof(unsigned int));
    });
    confusion_private_kernel<<<Bp, CF_THREADS, smem, s>>>((const long long*)gt, (const long long*)pred, n_pixels, num_class,
                                                          (unsigned int*)workspace, vec_ok);
    confusion_finalize_wide_kernel<<<(num_class * num_class + 127) / 128, CFF_WARPS * 32, 0, s>>>(
        (const unsigned int*)workspace, Bp, num_class * num_class, (long long*)cm_out);
    ADD_RETURN_LAUNCH();
  }
  int B = confusion_blocks(n_pixels);
  confusion_kernel<<<B, HD_THREADS, 0, s>>>((const long long*)gt, (const long long*)pred, n_pixels, num_class,
                                            (unsigned int*)workspace, vec_ok);
  if (g_confusion_impl == 2)
    confusion_finalize_wide_kernel<<<(num_class * num_class + 127) / 128, CFF_WARPS * 32, 0, s>>>(
        (const unsigned int*)workspace, B, num_class * num_class, (long long*)cm_out);
  else
    confusion_finalize_kernel<<<1, HD_THREADS, 0, s>>>((const unsigned int*)workspace, B, num_class * num_class,
                                                       (long long*)cm_out);
  ADD_RETURN_LAUNCH();
}

extern "C" int64_t add_confidence_workspace_bytes(int n, int H, int W) {
  if (n <= 0 || H <= 0 || W <= 0) return ADD_ERR_BAD_ARG;
  return (int64_t)confidence_blocks((long long)n * H * W) * 2 * sizeof(double);
}

extern "C" int add_confidence_nchw(const float* logits, int n, int num_class, int H, int W, float threshold,
                                   float* out2, void* workspace, int64_t workspace_bytes, void* stream) {
  ADD_CHECK_ARG(logits && out2 && workspace && n > 0 && H > 0 && W > 0 && num_class > 0);
  ADD_CHECK_SUP(num_class <= MAX_CLASS);
  if (workspace_bytes < add_confidence_workspace_bytes(n, H, W)) return ADD_ERR_WORKSPACE;
  long long HW = (long long)H * W;
  int B = confidence_blocks(HW * n);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  confidence_kernel<<<B, 256, 0, s>>>(logits, n, num_class, HW, threshold, (double*)workspace);
  confidence_finalize_kernel<<<1, 32, 0, s>>>((const double*)workspace, B, 1.0 / (double)HW, out2);
  ADD_RETURN_LAUNCH();
}
