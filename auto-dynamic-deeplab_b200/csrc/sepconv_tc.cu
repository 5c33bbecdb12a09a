// sepconv_tc.cu — one SepConv half (ReLU -> depthwise KxK -> pointwise 1x1 -> folded BN [-> ReLU] [+= y])
// for bf16 NHWC activations on sm_100a: CUDA-core depthwise feeding a tcgen05 pointwise GEMM.
//
// Replaces operations.py:51-54 / :55-58 (nn.Conv2d(groups=C) + nn.Conv2d(C,C,1) + BatchNorm [+ ReLU]).
//
// One CTA = one 8x16 patch of output pixels (128 = UMMA M) of one image, all channels.
//  1. ONE TMA box {C, 16+K-1, 8+K-1, 1} brings the input halo into shared memory; pixels outside the
//     image are zero-filled by the TMA unit (= the depthwise conv's zero padding).  The pointwise
//     weights (bf16, UMMA K-major SWIZZLE_128B image, packed by add_conv2d_tc_pack) arrive by TMA too.
//  2. ReLU-on-load is one in-place sweep over the halo.
//  3. Depthwise on the CUDA cores: a thread owns (channel pair, column pair) and slides down the
//     8+K-1 halo rows with K rotating fp32x2 accumulators per column — packed FFMA2 (fma.rn.f32x2),
//     K+1 shared-memory loads per 2*K*K FFMA2.  Each finished output row is rounded to bf16 and stored
//     straight into the UMMA A-operand tile (K-major, SWIZZLE_128B) in shared memory.
//  4. One thread issues tcgen05.mma (M=128, N=Cout padded to 16, K=C) into a TMEM accumulator.
//  5. Epilogue (shared with conv_tc.cu): tcgen05.ld -> +bias -> (+= y) -> ReLU -> bf16/fp32 stores
//     into the (possibly channel-sliced) NHWC output view.
// Several CTAs are resident per SM (C=40: 3-5), so one CTA's TMA / MMA / epilogue overlaps another's
// depthwise phase without an intra-CTA pipeline.
#include "tc_common.cuh"

namespace {

constexpr int SC_TH = 8, SC_TW = 16;           // output patch: 128 pixels, pixel m = row*16 + col
constexpr int SC_MAX_CTHREADS = 320;           // compute threads (plus one control warp)

struct ScParams {
  TcParams e;                // what epilogue_store reads (y, bias, Ho, Wo, Cout, ys, bw_log2 = 4, n_pad, flags)
  const float* w_dw;         // [K][K][C] fp32
  int C, kchunks, n_cthreads, n_items;
  uint32_t halo_bytes;
  int merged;                // 1: map_x is the merged {W*C, H, N} map in 8-byte elements (contiguous input)
};

__device__ __forceinline__ float2 bf16x2_to_float2(uint32_t v) {
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
__device__ __forceinline__ uint32_t float2_to_bf16x2(float2 v) {
  __nv_bfloat162 b = __floats2bfloat162_rn(v.x, v.y);
  return *reinterpret_cast<uint32_t*>(&b);
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

template <int K>
__global__ void __launch_bounds__(SC_MAX_CTHREADS + 32, 2)
sepconv_half_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const ScParams p) {
  constexpr int HR = SC_TH + K - 1, HC = SC_TW + K - 1;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];            // [0] TMA landed, [1] accumulator complete
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float bias_s[TC_MAX_NPAD];
  stage_bias(bias_s, p.e, threadIdx.x, blockDim.x);

  const int C = p.C;
  const uint32_t a_base = (smem_u32(smem_raw) + 1023u) & ~1023u;            // kchunks x 16 KB A tiles
  const uint32_t b_base = a_base + (uint32_t)p.kchunks * TC_A_BYTES;        // kchunks x (n_pad x 128 B)
  const uint32_t halo = b_base + (uint32_t)p.kchunks * p.e.b_bytes;         // [HR][HC][C] bf16
  const uint32_t wdw_s = halo + p.halo_bytes;                               // [K*K][C] fp32
  const uint32_t bar_in = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ctrl_warp = p.n_cthreads >> 5;
  const bool is_ctrl = warp == ctrl_warp;

  int t = blockIdx.x;
  const int tx = t % p.e.tiles_x; t /= p.e.tiles_x;
  const int ty = t % p.e.tiles_y; const int n = t / p.e.tiles_y;
  const int x0 = tx * SC_TW, y0 = ty * SC_TH;

  if (is_ctrl) {
    if (lane == 0) {
      mbar_init(bar_in, 1);
      mbar_init(bar_mma, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_smem)), "r"((uint32_t)p.e.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  pdl_launch_dependents();
  pdl_wait();                     // the previous kernel's activations are complete and visible from here on

  if (is_ctrl) {
    if (elect_one()) {
      mbar_expect_tx(bar_in, p.halo_bytes + (uint32_t)p.kchunks * p.e.b_bytes);
      if (p.merged) tma_load_3d(halo, &map_x, bar_in, (x0 - K / 2) * (C >> 2), y0 - K / 2, n);
      else tma_load_4d(halo, &map_x, bar_in, 0, x0 - K / 2, y0 - K / 2, n);
      for (int kc = 0; kc < p.kchunks; ++kc) tma_load_3d(b_base + kc * p.e.b_bytes, &map_w, bar_in, 0, 0, kc);
    }
  } else {
    // ---- stage the depthwise weights; zero the K-padding columns of the last A chunk ----
    for (int i = tid; i < K * K * C / 4; i += p.n_cthreads) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.w_dw) + i);
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(wdw_s + 16u * i), "f"(w4.x), "f"(w4.y), "f"(w4.z), "f"(w4.w) : "memory");
    }
    {
      const int c_last = C - (p.kchunks - 1) * TC_BK;            // channels in the last chunk
      const int u0 = c_last >> 3, u1 = ((c_last + 15) >> 4) << 1;   // 16-byte units [u0, u1) are padding read by the MMA
      const uint32_t a_last = a_base + (uint32_t)(p.kchunks - 1) * TC_A_BYTES;
      for (int i = tid; i < TC_BM * (u1 - u0); i += p.n_cthreads) {
        const int m = i / (u1 - u0), u = u0 + i % (u1 - u0);
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a_last + m * 128 + ((u ^ (m & 7)) << 4)), "r"(0u) : "memory");
      }
    }
    mbar_wait(bar_in, 0);
    if (p.e.flags & ADD_RELU_IN) {
      for (uint32_t off = tid * 16; off < p.halo_bytes; off += p.n_cthreads * 16) {
        const uint32_t addr = halo + off;
        uint32_t v0, v1, v2, v3;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(addr));
        asm("max.bf16x2 %0, %0, %1;" : "+r"(v0) : "r"(0u));
        asm("max.bf16x2 %0, %0, %1;" : "+r"(v1) : "r"(0u));
        asm("max.bf16x2 %0, %0, %1;" : "+r"(v2) : "r"(0u));
        asm("max.bf16x2 %0, %0, %1;" : "+r"(v3) : "r"(0u));
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
      }
    }
    asm volatile("bar.sync 1, %0;" ::"r"(p.n_cthreads) : "memory");   // weights staged + ReLU sweep done

    // ---- depthwise: item = (column pair xp, channel pair cp), sliding down the halo rows ----
    const int cpairs = C >> 1;
    for (int item = tid; item < p.n_items; item += p.n_cthreads) {
      const int cp = item % cpairs, xp = item / cpairs;
      float2 w[K * K];
#pragma unroll
      for (int i = 0; i < K * K; ++i) {
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(w[i].x), "=f"(w[i].y) : "r"(wdw_s + (uint32_t)(i * C + 2 * cp) * 4u));
      }
      float2 acc0[K], acc1[K];
#pragma unroll
      for (int i = 0; i < K; ++i) acc0[i] = acc1[i] = make_float2(0.f, 0.f);
      const uint32_t hsrc = halo + (uint32_t)((2 * xp) * C + 2 * cp) * 2u;
      const int kc = cp >> 5;                                     // 64-channel chunk of the A operand
      const uint32_t a_tile = a_base + (uint32_t)kc * TC_A_BYTES;
      const uint32_t cbyte = (uint32_t)(cp & 31) * 4u;             // byte inside the 128-byte row
#pragma unroll
      for (int r = 0; r < HR; ++r) {
        float2 in[K + 1];
#pragma unroll
        for (int j = 0; j <= K; ++j) in[j] = bf16x2_to_float2(lds32(hsrc + (uint32_t)((r * HC + j) * C) * 2u));
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          const int o = r - ky;
          if (o < 0 || o >= SC_TH) continue;
#pragma unroll
          for (int kx = 0; kx < K; ++kx) {
            acc0[o % K] = __ffma2_rn(w[ky * K + kx], in[kx], acc0[o % K]);
            acc1[o % K] = __ffma2_rn(w[ky * K + kx], in[kx + 1], acc1[o % K]);
          }
        }
        const int oc = r - (K - 1);                                // this output row is complete
        if (oc >= 0) {
          const int m0 = oc * SC_TW + 2 * xp, m1 = m0 + 1;
          sts32(a_tile + m0 * 128 + ((((cbyte >> 4) ^ (m0 & 7)) << 4) | (cbyte & 15u)), float2_to_bf16x2(acc0[oc % K]));
          sts32(a_tile + m1 * 128 + ((((cbyte >> 4) ^ (m1 & 7)) << 4) | (cbyte & 15u)), float2_to_bf16x2(acc1[oc % K]));
          acc0[oc % K] = acc1[oc % K] = make_float2(0.f, 0.f);
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // A-tile stores -> visible to the UMMA (async proxy)
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (is_ctrl) {
    if (elect_one()) {
      mbar_wait(bar_in, 0);                                         // pointwise weights landed
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.e.n_pad >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      for (int kc = 0; kc < p.kchunks; ++kc) {
        const int krem = C - kc * TC_BK;
        const int ksteps = krem >= TC_BK ? TC_BK / 16 : (krem + 15) / 16;
        const uint64_t adesc = make_kmajor_sw128_desc(a_base + kc * TC_A_BYTES);
        const uint64_t bdesc = make_kmajor_sw128_desc(b_base + kc * p.e.b_bytes);
        for (int k = 0; k < ksteps; ++k) umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kc > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit(bar_mma);
    }
  } else if (warp < 4) {
    mbar_wait(bar_mma, 0);
    epilogue_store(p.e, tmem_base, bias_s, warp, lane, n, y0, x0);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (is_ctrl) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.e.tmem_cols) : "memory");
  }
}

// ---- persistent, warp-specialised version (C <= 80) --------------------------------------------------
// The one-tile-per-CTA kernel above spends most of a CTA's life in its latency chain (TMEM alloc -> TMA
// round trip -> depthwise -> MMA -> epilogue: ~7 us for ~0.5 us of depthwise work, ncu r01d: issue-active
// 27-32 %).  Here a CTA loops over tiles with every stage double-buffered, so the chain is paid once:
//   warp 4      producer   TMA halo of tile i+1 / i+2 into a 2-slot ring (halo_full / halo_empty)
//   warps 6..   depthwise  thread = (channel pair, column pair), weights live in REGISTERS for the whole
//                          kernel; ReLU-on-load is a max.bf16x2 per loaded word; output -> A[i&1]
//   warp 5      MMA        tcgen05.mma A[i&1] x Wpw -> TMEM[i&1]; commit frees A and publishes TMEM
//   warps 0..3  epilogue   TMEM[i&1] -> +bias -> (+= y) -> ReLU -> global, overlapped with tile i+1's depthwise
constexpr int SP_MAX_S = 6;                   // max halo ring depth (runtime p.stages <= this)
constexpr int SP_FIRST_DW_WARP = 6;

struct SpParams {
  TcParams e;
  const float* w_dw;
  int C, kchunks, n_dw_warps, n_tiles;
  uint32_t halo_bytes;
  int merged;
  int stages;                // halo ring depth: how many tiles ahead the TMA producer runs
};

template <int K, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
sepconv_half_tc_persistent_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const SpParams p) {
  constexpr int HR = SC_TH + K - 1, HC = SC_TW + K - 1;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * SP_MAX_S + 8 + 1];
  __shared__ __align__(16) float bias_s[TC_MAX_NPAD];    // staged by the epilogue warps (warps 0..3)
  const int SP_S = p.stages;
  __shared__ uint32_t tmem_base_smem;

  const int C = p.C;
  const uint32_t a_bytes = (uint32_t)p.kchunks * TC_A_BYTES;
  const uint32_t a_base = (smem_u32(smem_raw) + 1023u) & ~1023u;            // 2 x kchunks x 16 KB
  const uint32_t b_base = a_base + 2u * a_bytes;                            // kchunks x (n_pad x 128 B)
  const uint32_t halo_base = b_base + (uint32_t)p.kchunks * p.e.b_bytes;    // SP_S x [HR][HC][C] bf16
  const uint32_t halo_stride = (p.halo_bytes + 127u) & ~127u;
  const uint32_t bar_hfull = smem_u32(&bars[0]), bar_hempty = smem_u32(&bars[SP_MAX_S]);
  const uint32_t bar_afull = smem_u32(&bars[2 * SP_MAX_S]), bar_aempty = smem_u32(&bars[2 * SP_MAX_S + 2]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * SP_MAX_S + 4]), bar_tempty = smem_u32(&bars[2 * SP_MAX_S + 6]);
  const uint32_t bar_b = smem_u32(&bars[2 * SP_MAX_S + 8]);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t tmem_cols = (uint32_t)p.e.tmem_cols;                        // per accumulator buffer

  if (warp == 4 && lane == 0) {
    for (int s = 0; s < SP_S; ++s) { mbar_init(bar_hfull + 8 * s, 1); mbar_init(bar_hempty + 8 * s, p.n_dw_warps); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_afull + 8 * b, p.n_dw_warps); mbar_init(bar_aempty + 8 * b, 1);
      mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 4);
    }
    mbar_init(bar_b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_smem)), "r"(2u * tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  pdl_launch_dependents();
  // programmatic dependent launch: only the threads that touch activations wait for the previous kernel (the halo
  // producer before its first load, the epilogue warps before their first y access); weight loads, the depthwise
  // warps' register weights and the bias staging overlap the predecessor's tail
  const int tiles_per_img = p.e.tiles_x * p.e.tiles_y;

  if (warp == 4) {
    // ===== producer =====
    if (elect_one()) {
      mbar_expect_tx(bar_b, (uint32_t)p.kchunks * p.e.b_bytes);
      for (int kc = 0; kc < p.kchunks; ++kc) tma_load_3d(b_base + kc * p.e.b_bytes, &map_w, bar_b, 0, 0, kc);
      pdl_wait();
      int it = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        const int s = it % SP_S; const uint32_t ph = (uint32_t)(it / SP_S) & 1u;
        const int n = tile / tiles_per_img, r = tile - n * tiles_per_img;
        const int ty = r / p.e.tiles_x, tx = r - ty * p.e.tiles_x;
        mbar_wait_relaxed(bar_hempty + 8 * s, ph ^ 1u);
        mbar_expect_tx(bar_hfull + 8 * s, p.halo_bytes);
        if (p.merged) tma_load_3d(halo_base + s * halo_stride, &map_x, bar_hfull + 8 * s, (tx * SC_TW - K / 2) * (C >> 2), ty * SC_TH - K / 2, n);
        else tma_load_4d(halo_base + s * halo_stride, &map_x, bar_hfull + 8 * s, 0, tx * SC_TW - K / 2, ty * SC_TH - K / 2, n);
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer =====
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.e.n_pad >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      mbar_wait(bar_b, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        const int b = it & 1; const uint32_t ph = (uint32_t)(it >> 1) & 1u;
        // tight spin: polling with a sleep (mbar_wait_relaxed) frees ~12 % of the SM's issue slots (ncu r3f) but puts up to
        // 128 ns between the depthwise warps' arrival and the MMA of every tile; measured neutral in the bench and the
        // microbench and +4 us per launch in the cold-cache ncu list (r03a), so the spin stays
        mbar_wait(bar_afull + 8 * b, ph);
        mbar_wait(bar_tempty + 8 * b, ph ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kc = 0; kc < p.kchunks; ++kc) {
          const int krem = C - kc * TC_BK;
          const int ksteps = krem >= TC_BK ? TC_BK / 16 : (krem + 15) / 16;
          const uint64_t adesc = make_kmajor_sw128_desc(a_base + b * a_bytes + kc * TC_A_BYTES);
          const uint64_t bdesc = make_kmajor_sw128_desc(b_base + kc * p.e.b_bytes);
          for (int k = 0; k < ksteps; ++k)
            umma_bf16(tmem_base + b * tmem_cols, adesc + 2 * k, bdesc + 2 * k, idesc, (kc > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(bar_aempty + 8 * b);      // A[b] may be overwritten once these MMAs have read it
        umma_commit(bar_tfull + 8 * b);       // accumulator b complete
      }
    }
  } else if (warp < 4) {
    // ===== epilogue =====
    stage_bias(bias_s, p.e, threadIdx.x, 128);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    pdl_wait();
    const bool pre_ok = epilogue_prefetchable(p.e);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int b = it & 1; const uint32_t ph = (uint32_t)(it >> 1) & 1u;
      const int n = tile / tiles_per_img, r = tile - n * tiles_per_img;
      const int ty = r / p.e.tiles_x, tx = r - ty * p.e.tiles_x;
      if (pre_ok) {         // += : this pixel's old y words are fetched while the depthwise / MMA of the tile still run
        uint4 old[EPI_PRE_MAX];
        epilogue_prefetch_old(p.e, warp, lane, n, ty * SC_TH, tx * SC_TW, old);
        mbar_wait_relaxed(bar_tfull + 8 * b, ph);
        epilogue_store(p.e, tmem_base + b * tmem_cols, bias_s, warp, lane, n, ty * SC_TH, tx * SC_TW, &old);
      } else {
        mbar_wait_relaxed(bar_tfull + 8 * b, ph);
        epilogue_store(p.e, tmem_base + b * tmem_cols, bias_s, warp, lane, n, ty * SC_TH, tx * SC_TW);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * b);
    }
  } else if (warp < SP_FIRST_DW_WARP + p.n_dw_warps) {
    // ===== depthwise =====
    const int dt = tid - SP_FIRST_DW_WARP * 32;
    const int cpairs = C >> 1;
    const int cp = dt % cpairs, xp = dt / cpairs;
    const bool relu_in = (p.e.flags & ADD_RELU_IN) != 0;
    float2 w[K * K];
#pragma unroll
    for (int i = 0; i < K * K; ++i) w[i] = __ldg(reinterpret_cast<const float2*>(p.w_dw + (size_t)i * C) + cp);
    {  // zero the K-padding columns of the last chunk of both A buffers (read by the MMA, never written below)
      const int c_last = C - (p.kchunks - 1) * TC_BK;
      const int u0 = c_last >> 3, u1 = ((c_last + 15) >> 4) << 1;
      const int nthr = p.n_dw_warps * 32;
      for (int i = dt; i < 2 * TC_BM * (u1 - u0); i += nthr) {
        const int b = i / (TC_BM * (u1 - u0)), j = i - b * TC_BM * (u1 - u0);
        const int m = j / (u1 - u0), u = u0 + j % (u1 - u0);
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};"
                     ::"r"(a_base + b * a_bytes + (uint32_t)(p.kchunks - 1) * TC_A_BYTES + m * 128 + ((u ^ (m & 7)) << 4)), "r"(0u) : "memory");
      }
    }
    const uint32_t hoff = (uint32_t)((2 * xp) * C + 2 * cp) * 2u;
    const uint32_t a_off = (uint32_t)(cp >> 5) * TC_A_BYTES;
    const uint32_t cbyte = (uint32_t)(cp & 31) * 4u;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int s = it % SP_S; const uint32_t hph = (uint32_t)(it / SP_S) & 1u;
      const int b = it & 1; const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(bar_hfull + 8 * s, hph);
      mbar_wait(bar_aempty + 8 * b, aph ^ 1u);
      const uint32_t hsrc = halo_base + s * halo_stride + hoff;
      const uint32_t a_tile = a_base + b * a_bytes + a_off;
      float2 acc0[K], acc1[K];
#pragma unroll
      for (int i = 0; i < K; ++i) acc0[i] = acc1[i] = make_float2(0.f, 0.f);
#pragma unroll
      for (int r = 0; r < HR; ++r) {
        float2 in[K + 1];
#pragma unroll
        for (int j = 0; j <= K; ++j) {
          uint32_t v = lds32(hsrc + (uint32_t)((r * HC + j) * C) * 2u);
          if (relu_in) asm("max.bf16x2 %0, %0, %1;" : "+r"(v) : "r"(0u));
          in[j] = bf16x2_to_float2(v);
        }
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          const int o = r - ky;
          if (o < 0 || o >= SC_TH) continue;
#pragma unroll
          for (int kx = 0; kx < K; ++kx) {
            acc0[o % K] = __ffma2_rn(w[ky * K + kx], in[kx], acc0[o % K]);
            acc1[o % K] = __ffma2_rn(w[ky * K + kx], in[kx + 1], acc1[o % K]);
          }
        }
        const int oc = r - (K - 1);
        if (oc >= 0) {
          const int m0 = oc * SC_TW + 2 * xp, m1 = m0 + 1;
          sts32(a_tile + m0 * 128 + ((((cbyte >> 4) ^ (m0 & 7)) << 4) | (cbyte & 15u)), float2_to_bf16x2(acc0[oc % K]));
          sts32(a_tile + m1 * 128 + ((((cbyte >> 4) ^ (m1 & 7)) << 4) | (cbyte & 15u)), float2_to_bf16x2(acc1[oc % K]));
          acc0[oc % K] = acc1[oc % K] = make_float2(0.f, 0.f);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar_afull + 8 * b); mbar_arrive(bar_hempty + 8 * s); }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 5) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * tmem_cols) : "memory");
  }
}

template <int K, int MAXT, int MINB>
int launch_sepconv_tc_persistent(const CUtensorMap& map_x, const CUtensorMap& map_w, const SpParams& p, int grid, int threads,
                                 size_t smem, cudaStream_t s) {
  static PerDeviceOnce once;
  once_per_device(once, [] {
    cudaFuncSetAttribute(sepconv_half_tc_persistent_kernel<K, MAXT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(sepconv_half_tc_persistent_kernel<K, MAXT, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  });
  launch_kernel(sepconv_half_tc_persistent_kernel<K, MAXT, MINB>, dim3(grid), dim3(threads), smem, s, map_x, map_w, p);
  ADD_RETURN_LAUNCH();
}

int g_sepconv_merged = 1;  // merged {W*C} halo tensor map for contiguous inputs (A/B switch: mode bit 1)
int g_sepconv_three = 0;   // 1 = try three CTAs per SM for the 3x3 (tuning; mode bit 2)
int g_sepconv_stages = 0;  // 0 = auto
int g_sepconv_mode = 1;    // 1 = persistent pipeline where it fits (default), 0 = one tile per CTA

template <int K>
int launch_sepconv_tc(const CUtensorMap& map_x, const CUtensorMap& map_w, const ScParams& p, long long grid, size_t smem, cudaStream_t s) {
  static PerDeviceOnce once;
  once_per_device(once, [] {
    cudaFuncSetAttribute(sepconv_half_tc_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(sepconv_half_tc_kernel<K>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  });
  launch_kernel(sepconv_half_tc_kernel<K>, dim3((unsigned)grid), dim3(p.n_cthreads + 32), smem, s, map_x, map_w, p);
  ADD_RETURN_LAUNCH();
}

}  // namespace

extern "C" int add_sepconv_half_tc_fwd(const add_tensor_t* x, const add_tensor_t* y, const float* w_dw,
                                       const void* w_pw_packed, const float* bias, int k, uint32_t flags,
                                       void* stream) {
  ADD_CHECK_ARG(tensor_ok(x) && tensor_ok(y) && w_dw && w_pw_packed);
  ADD_CHECK_ARG(x->n == y->n && x->h == y->h && x->w == y->w);
  ADD_CHECK_SUP(k == 3 || k == 5);
  ADD_CHECK_SUP(x->dtype == ADD_BF16 && x->c % 8 == 0 && x->c <= 256 && y->c <= 256);
  ADD_CHECK_SUP(x->pix_stride % 8 == 0 && ((uintptr_t)x->ptr % 16) == 0 && ((uintptr_t)w_pw_packed % 16) == 0 &&
                ((uintptr_t)w_dw % 16) == 0);
  if (y->dtype == ADD_BF16) ADD_CHECK_SUP(y->pix_stride % 8 == 0 && ((uintptr_t)y->ptr % 16) == 0);
  else ADD_CHECK_SUP(y->pix_stride % 4 == 0 && ((uintptr_t)y->ptr % 16) == 0);
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) { g_add_last_cuda_error = (int)cudaErrorSymbolNotFound; return ADD_ERR_CUDA; }

  const int C = x->c, HR = SC_TH + k - 1, HC = SC_TW + k - 1;
  ScParams p;
  std::memset(&p, 0, sizeof(p));
  p.e.y = y->ptr; p.e.bias = bias; p.e.Ho = y->h; p.e.Wo = y->w; p.e.Cout = y->c; p.e.ys = y->pix_stride;
  p.e.y_is_f32 = (y->dtype == ADD_F32);
  p.e.bw_log2 = 4;
  p.e.tiles_x = ceil_div(y->w, SC_TW); p.e.tiles_y = ceil_div(y->h, SC_TH);
  p.e.n_pad = npad_of(y->c);
  p.e.tmem_cols = 32; while (p.e.tmem_cols < p.e.n_pad) p.e.tmem_cols <<= 1;
  p.e.b_bytes = (uint32_t)p.e.n_pad * 128u;
  p.e.flags = flags;
  p.w_dw = w_dw; p.C = C; p.kchunks = kchunks_of(C);
  p.n_items = (SC_TW / 2) * (C / 2);
  int nct = round_up(p.n_items, 32);
  if (nct > SC_MAX_CTHREADS) nct = SC_MAX_CTHREADS;
  if (nct < 128) nct = 128;                      // the four epilogue warps
  p.n_cthreads = nct;
  p.halo_bytes = (uint32_t)HR * HC * C * 2u;
  const size_t smem = (size_t)p.kchunks * TC_A_BYTES + (size_t)p.kchunks * p.e.b_bytes + p.halo_bytes +
                      (size_t)k * k * C * 4 + 1024;
  ADD_CHECK_SUP(smem <= 220u * 1024u);

  CUtensorMap map_x, map_w;
  // Contiguous input (not a channel slice): merge W and C into one dimension of 8-byte elements, so a halo row is
  // ONE contiguous box row (HC*C*2 bytes) instead of HC separate C*2-byte rows — the TMA unit's cost is per box row.
  // Out-of-image columns/rows are still zero-filled (they are out of bounds of the merged dimension too).
  p.merged = (g_sepconv_merged && x->pix_stride == C && HC * C / 4 <= 256) ? 1 : 0;
  if (p.merged) {
    cuuint64_t dims[3] = {(cuuint64_t)x->w * C / 4, (cuuint64_t)x->h, (cuuint64_t)x->n};
    cuuint64_t strides[2] = {(cuuint64_t)x->w * C * 2, (cuuint64_t)x->h * x->w * C * 2};
    cuuint32_t box[3] = {(cuuint32_t)(HC * C / 4), (cuuint32_t)HR, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (encode(&map_x, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, x->ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ADD_ERR_UNSUPPORTED;
  } else {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)x->w, (cuuint64_t)x->h, (cuuint64_t)x->n};
    cuuint64_t strides[3] = {(cuuint64_t)x->pix_stride * 2, (cuuint64_t)x->w * x->pix_stride * 2,
                             (cuuint64_t)x->h * x->w * x->pix_stride * 2};
    cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)HC, (cuuint32_t)HR, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (encode(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x->ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ADD_ERR_UNSUPPORTED;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)TC_BK, (cuuint64_t)p.e.n_pad, (cuuint64_t)p.kchunks};
    cuuint64_t strides[2] = {(cuuint64_t)TC_BK * 2, (cuuint64_t)p.e.n_pad * TC_BK * 2};
    cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)p.e.n_pad, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (encode(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w_pw_packed), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ADD_ERR_UNSUPPORTED;
  }
  const long long grid = (long long)p.e.tiles_x * p.e.tiles_y * y->n;
  ADD_CHECK_SUP(grid < (1ll << 31));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (g_sepconv_mode == 1 && C <= 80) {
    SpParams q;
    std::memset(&q, 0, sizeof(q));
    q.e = p.e; q.w_dw = w_dw; q.C = C; q.kchunks = p.kchunks; q.n_dw_warps = C / 8; q.n_tiles = (int)grid;
    q.halo_bytes = p.halo_bytes; q.merged = p.merged;
    const size_t hstride = (q.halo_bytes + 127u) & ~127u;
    const size_t fixed = 2 * (size_t)q.kchunks * TC_A_BYTES + (size_t)q.kchunks * q.e.b_bytes + 1024;
    const int threads = (SP_FIRST_DW_WARP + q.n_dw_warps) * 32;
    // ring depth: as deep as fits two CTAs per SM (<= 110 KB each) when the CTA is small enough for that,
    // else as deep as fits one CTA (<= 200 KB); at least 2
    const bool small = threads <= 352 && 4 * q.e.tmem_cols <= 512;
    const size_t budget = (small && fixed + 2 * hstride <= 110u * 1024u) ? 110u * 1024u : 200u * 1024u;
    int stages = (int)((budget - fixed) / hstride);
    if (g_sepconv_stages > 0) stages = g_sepconv_stages;
    if (stages > SP_MAX_S) stages = SP_MAX_S;
    if (stages < 2) stages = 2;
    q.stages = stages;
    const size_t psmem = fixed + (size_t)stages * hstride;
    int sms = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    if (psmem <= 200u * 1024u && 2 * q.e.tmem_cols <= 512) {
      const bool two = threads <= 352 && psmem <= 110u * 1024u && 4 * q.e.tmem_cols <= 512;
      long long g = (long long)sms * (two ? 2 : 1) * g_add_grid_pct / 100;
      if (g > grid) g = grid;
      if (g < 1) g = 1;
      if (two && k == 3 && g_sepconv_three) {
        // 3x3 at C <= 40: three CTAs per SM (15 depthwise warps) when a 2-deep halo ring fits a third of the SM
        const size_t psmem3 = fixed + 2 * hstride;
        if (psmem3 <= 74u * 1024u && 6 * q.e.tmem_cols <= 512) {
          q.stages = 2;
          long long g3 = (long long)sms * 3 * g_add_grid_pct / 100;
          if (g3 > grid) g3 = grid;
          return launch_sepconv_tc_persistent<3, 352, 3>(map_x, map_w, q, (int)g3, threads, psmem3, s);
        }
      }
      if (two) return k == 3 ? launch_sepconv_tc_persistent<3, 352, 2>(map_x, map_w, q, (int)g, threads, psmem, s)
                             : launch_sepconv_tc_persistent<5, 352, 2>(map_x, map_w, q, (int)g, threads, psmem, s);
      return k == 3 ? launch_sepconv_tc_persistent<3, 512, 1>(map_x, map_w, q, (int)g, threads, psmem, s)
                    : launch_sepconv_tc_persistent<5, 512, 1>(map_x, map_w, q, (int)g, threads, psmem, s);
    }
  }
  return k == 3 ? launch_sepconv_tc<3>(map_x, map_w, p, grid, smem, s) : launch_sepconv_tc<5>(map_x, map_w, p, grid, smem, s);
}

/* 1 = persistent warp-specialised pipeline where it fits (default); 0 = one tile per CTA (kept for A/B runs). */
extern "C" int add_sepconv_tc_set_mode(int mode) {
  if (mode >= 16) { g_sepconv_stages = (mode >> 4) & 15; mode &= 15; }      // bits 4..7: forced halo ring depth (tuning)
  if (mode < 0 || mode > 7) return ADD_ERR_BAD_ARG;
  g_sepconv_three = (mode & 4) ? 1 : 0;
  g_sepconv_mode = mode & 1;
  g_sepconv_merged = (mode & 2) ? 0 : 1;      // bit 1 set = per-pixel halo rows even for contiguous inputs
  return ADD_OK;
}
