// conv_tc.cu — tcgen05 / TMEM implicit-GEMM convolution for bf16 NHWC activations (sm_100a).
//
//   y[n,oy,ox,co] (+)= act( bias[co] + sum_{ky,kx,ci} W[ky][kx][ci][co] *
//                           relu?(x[n, oy*s - pad + ky*dil, ox*s - pad + kx*dil, ci]) )
//
// GEMM view: M = output pixels, N = Cout (padded to 16, one N tile: Cout <= 256), K = taps x Cin.
// One CTA computes a 128-pixel spatial patch (BH x BW, BW a power of two) of one image.
//  * A (activations): for every tap the 128 x 64-channel operand tile is ONE 4-D TMA box
//    {64 ch, BW, BH, 1} of the NHWC tensor at the tap-shifted coordinate; out-of-image pixels and
//    channels past Cin are zero-filled by the TMA unit (= the conv's zero padding, for free), the
//    box lands in shared memory already in the UMMA K-major SWIZZLE_128B layout.  Strided convs
//    use the tensor map's element strides.
//  * B (weights): pre-packed bf16 [tap*kchunk][N_pad][64] (K-major), one 3-D TMA box per K chunk.
//  * D: fp32 accumulator in TMEM (N_pad columns x 128 lanes), tcgen05.mma issued by one thread.
//  * ReLU-on-load (the reference's ReLU -> conv order) is an in-place pass over the landed A tile
//    by the four otherwise idle epilogue warps, fenced to the async proxy before the MMA reads it.
//  * Epilogue: tcgen05.ld -> +bias (folded BN shift) -> (+= y, cell node sum) -> ReLU -> bf16/fp32
//    stores into the (possibly channel-sliced) NHWC output view.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer,
// warps 2..5 = ReLU pass + epilogue.  mbarrier ring of `stages` {A,B} slots; two CTAs per SM
// overlap one tile's epilogue with the other's main loop.
#include "tc_common.cuh"

namespace {

// ---- the kernel ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS)
conv2d_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3 * TC_MAX_STAGES + 1];   // full[s], empty[s], relu[s], accum
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float bias_s[TC_MAX_NPAD];
  // a per-image bias (ASPP pool branch) is ACTIVATION data written by the previous kernel: wait for it first
  if (p.bias_img_stride != 0) pdl_wait();
  stage_bias(bias_s, p, threadIdx.x, TC_THREADS, (int)(blockIdx.x / (unsigned)(p.tiles_x * p.tiles_y)));

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B tiles: 1024-B aligned
  const uint32_t stage_bytes = TC_A_BYTES + p.b_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool relu_in = (p.flags & ADD_RELU_IN) != 0;

  const uint32_t bar_full = smem_u32(&bars[0]);
  const uint32_t bar_empty = smem_u32(&bars[TC_MAX_STAGES]);
  const uint32_t bar_relu = smem_u32(&bars[2 * TC_MAX_STAGES]);
  const uint32_t bar_accum = smem_u32(&bars[3 * TC_MAX_STAGES]);

  // tile coordinates
  int t = blockIdx.x;
  const int tx = t % p.tiles_x; t /= p.tiles_x;
  const int ty = t % p.tiles_y; const int n = t / p.tiles_y;
  const int BW = 1 << p.bw_log2;
  const int x0 = tx * BW, y0 = ty * (TC_BM >> p.bw_log2);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_relu + 8 * s, 128);
    }
    mbar_init(bar_accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_smem)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  pdl_launch_dependents();
  pdl_wait();                     // the previous kernel's activations are complete and visible from here on

  const int iters = p.taps * p.kchunks;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        const int tap = it / p.kchunks, kc = it - tap * p.kchunks;
        const int ky = tap / p.taps_w, kx = tap - ky * p.taps_w;
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        const uint32_t a_dst = smem_base + s * stage_bytes, b_dst = a_dst + TC_A_BYTES;
        mbar_expect_tx(bar_full + 8 * s, stage_bytes);
        tma_load_4d(a_dst, &map_x, bar_full + 8 * s, kc * TC_BK, x0 * p.stride - p.pad + kx * p.dil,
                    y0 * p.stride - p.pad + ky * p.dil, n);
        tma_load_3d(b_dst, &map_w, bar_full + 8 * s, 0, 0, it);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_pad >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        const int kc = it % p.kchunks;
        const int krem = p.Cin - kc * TC_BK;
        const int ksteps = krem >= TC_BK ? TC_BK / 16 : (krem + 15) / 16;
        mbar_wait((relu_in ? bar_relu : bar_full) + 8 * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_src = smem_base + s * stage_bytes, b_src = a_src + TC_A_BYTES;
        const uint64_t adesc = make_kmajor_sw128_desc(a_src), bdesc = make_kmajor_sw128_desc(b_src);
        for (int k = 0; k < ksteps; ++k)   // +32 B along K inside the swizzle row = +2 in the address field
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
        umma_commit(bar_empty + 8 * s);     // frees the slot when these MMAs have read it
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      umma_commit(bar_accum);               // accumulator complete
    }
  } else {
    // ===== ReLU pass (optional) + epilogue: warps 2..5 =====
    const int et = threadIdx.x - 64;        // 0..127
    if (relu_in) {
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(bar_full + 8 * s, ph);
        const uint32_t a_src = smem_base + s * stage_bytes;
        relu_sweep(a_src, TC_A_BYTES, et);
        mbar_arrive(bar_relu + 8 * s);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
    // ---- epilogue ----
    mbar_wait(bar_accum, 0);
    epilogue_store(p, tmem_base, bias_s, warp, lane, n, y0, x0);
  }

  // ---- teardown ----
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// ---- persistent, warp-specialised version of the kernel above ---------------------------------------
// One tile per CTA pays the CTA's whole latency chain (TMEM alloc, barrier init, bias staging, tile-index math
// by every thread, TMA round trip, epilogue drain) for 128 pixels; for the network's many small convs that
// chain — not HBM or the tensor pipe — was the cost (ncu r01g: 4.2 k warp instructions per 128-pixel 1x1 tile,
// ~0.8 k of them essential).  Here a CTA loops over tiles: the {A,B} ring keeps streaming across tile
// boundaries, the accumulator is double-buffered in TMEM so tile i's epilogue overlaps tile i+1's MMAs, and
// for small K the weights are loaded ONCE and stay resident.
//   warp 0 TMA producer | warp 1 MMA issuer (+TMEM alloc) | warps 2..5 ReLU sweep | warps 6..9 epilogue
constexpr int TCP_THREADS = 320;
constexpr int TCP_MAX_STAGES = 8;

__device__ __forceinline__ void issue_tap_ring_dispatch(int ks, uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc_first);

// MT = M-tiles (128 pixels each) that share every weight tile: with N_pad = 256 the weight tile (32 KB) is 2/3 of a
// stage's L2->smem traffic and the ASPP / decoder 3x3s ran AT the L2 bandwidth ceiling (1085 TF/s x 48 KB per
// 128x256x64 MACs = 12.4 TB/s); MT = 2 computes two adjacent pixel tiles per weight tile (1.5x less traffic).
// CL = thread-block cluster size (1 or 2).  CL = 2: the two CTAs of a cluster walk the SAME sequence of weight tiles
// in lockstep (different pixel units); each fetches HALF of every weight tile and TMA-multicasts it into both CTAs'
// ring slot, so a weight tile crosses L2 -> SM once per cluster instead of once per CTA (ncu r01y: the ASPP 3x3s
// moved 10.6 TB/s from L2, ~85 % of the measured ~12.4 TB/s L2 ceiling, tensor pipe 61 % busy).  A slot is recycled
// when BOTH CTAs' MMAs have consumed it: the empty barrier counts CL arrivals, the MMA commit is multicast.
template <int MT, int CL>
__global__ void __launch_bounds__(TCP_THREADS)
conv2d_tc_persistent_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3 * TCP_MAX_STAGES + 5];   // full[s], empty[s], relu[s], tfull[2], tempty[2], bres
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float bias_s[TC_MAX_NPAD];    // staged by the epilogue warps themselves (off the prologue's critical path)

  const int iters = p.taps * p.kchunks;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bres_base = smem_base;                                                  // resident weights (if any)
  const uint32_t ring_base = smem_base + (p.b_resident ? (uint32_t)iters * p.b_bytes : 0u);
  const uint32_t stage_bytes = MT * TC_A_BYTES + (p.b_resident ? 0u : p.b_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool relu_in = (p.flags & ADD_RELU_IN) != 0;
  const uint32_t tmem_cols = (uint32_t)p.tmem_cols;
  const int nbuf = (2u * MT * tmem_cols <= 512u) ? 2 : 1;          // accumulator sets: double-buffered when TMEM allows

  const uint32_t bar_full = smem_u32(&bars[0]);
  const uint32_t bar_empty = smem_u32(&bars[TCP_MAX_STAGES]);
  const uint32_t bar_relu = smem_u32(&bars[2 * TCP_MAX_STAGES]);
  const uint32_t bar_tfull = smem_u32(&bars[3 * TCP_MAX_STAGES]);
  const uint32_t bar_tempty = smem_u32(&bars[3 * TCP_MAX_STAGES + 2]);
  const uint32_t bar_bres = smem_u32(&bars[3 * TCP_MAX_STAGES + 4]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, CL);
      mbar_init(bar_relu + 8 * s, 128);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 4); }
    mbar_init(bar_bres, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_smem)), "r"((uint32_t)nbuf * MT * tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CL > 1) cluster_sync_all();          // the peer's barriers are initialised before anyone arrives on them remotely
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  pdl_launch_dependents();        // (waits for the previous kernel: producer before its first activation load,
                                  //  epilogue warps before their first y access — see pdl_wait() below)
  const int BW = 1 << p.bw_log2, BH = TC_BM >> p.bw_log2;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int n_units = (p.n_tiles + MT - 1) / MT;                   // a unit = MT consecutive tiles
  const int n_img = p.n_tiles / tiles_per_img;
  // every CTA runs the same number of units (lockstep within a cluster); surplus units are computed on zero-filled
  // tiles and never stored
  const int units_per_cta = (n_units + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      if (p.b_resident) {
        mbar_expect_tx(bar_bres, (uint32_t)iters * p.b_bytes);
        for (int it = 0; it < iters; ++it) tma_load_3d(bres_base + it * p.b_bytes, &map_w, bar_bres, 0, 0, it);
      }
      pdl_wait();
      int s = 0; uint32_t ph = 0;
      for (int uj = 0; uj < units_per_cta; ++uj) {
        const int u = (int)blockIdx.x + uj * (int)gridDim.x;
        int tn[MT], xin[MT], yin[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const int tile = u * MT + m;
          const int n = tile / tiles_per_img, r = tile - n * tiles_per_img;
          const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
          tn[m] = tile < p.n_tiles ? n : n_img;                   // past the end: an image index out of bounds -> zero fill
          xin[m] = tx * BW * p.stride - p.pad; yin[m] = ty * BH * p.stride - p.pad;
        }
        int ky = 0, kx = 0, kc = 0;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          const uint32_t a_dst = ring_base + s * stage_bytes;
          mbar_expect_tx(bar_full + 8 * s, stage_bytes);
#pragma unroll
          for (int m = 0; m < MT; ++m)
            tma_load_4d(a_dst + m * TC_A_BYTES, &map_x, bar_full + 8 * s, kc * TC_BK, xin[m] + kx * p.dil, yin[m] + ky * p.dil, tn[m]);
          if (!p.b_resident) {
            if (CL > 1)      // my half of the weight rows, into both CTAs' slot s
              tma_load_3d_multicast(a_dst + MT * TC_A_BYTES + crank * (p.b_bytes / CL), &map_w, bar_full + 8 * s, 0,
                                    (int)crank * (p.n_pad / CL), it, (uint16_t)((1u << CL) - 1u));
            else
              tma_load_3d(a_dst + MT * TC_A_BYTES, &map_w, bar_full + 8 * s, 0, 0, it);
          }
          if (++kc == p.kchunks) { kc = 0; if (++kx == p.taps_w) { kx = 0; ++ky; } }
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_pad >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      if (p.b_resident) mbar_wait(bar_bres, 0);
      int s = 0; uint32_t ph = 0; int ti = 0;
      for (int uj = 0; uj < units_per_cta; ++uj, ++ti) {
        const int ab = nbuf == 2 ? (ti & 1) : 0;
        const uint32_t tph = nbuf == 2 ? ((uint32_t)(ti >> 1) & 1u) : ((uint32_t)ti & 1u);
        mbar_wait(bar_tempty + 8 * ab, tph ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        int kc = 0;
        for (int it = 0; it < iters; ++it) {
          const int krem = p.Cin - kc * TC_BK;
          const int ksteps = krem >= TC_BK ? TC_BK / 16 : (krem + 15) / 16;
          mbar_wait((relu_in ? bar_relu : bar_full) + 8 * s, ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_src = ring_base + s * stage_bytes;
          const uint32_t b_src = p.b_resident ? bres_base + it * p.b_bytes : a_src + MT * TC_A_BYTES;
          const uint64_t bdesc = make_kmajor_sw128_desc(b_src);
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            const uint64_t adesc = make_kmajor_sw128_desc(a_src + m * TC_A_BYTES);
            issue_tap_ring_dispatch(ksteps, tmem_base + (uint32_t)(ab * MT + m) * tmem_cols, adesc, bdesc, idesc, it > 0 ? 1u : 0u);
          }
          if (CL > 1) umma_commit_multicast(bar_empty + 8 * s, (uint16_t)((1u << CL) - 1u));
          else umma_commit(bar_empty + 8 * s);
          if (++kc == p.kchunks) kc = 0;
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        umma_commit(bar_tfull + 8 * ab);
      }
    }
  } else if (warp < 6) {
    // ===== ReLU-on-load sweep: warps 2..5 =====
    if (relu_in) {
      const int et = threadIdx.x - 64;
      int s = 0; uint32_t ph = 0;
      for (int uj = 0; uj < units_per_cta; ++uj) {
        for (int it = 0; it < iters; ++it) {
          mbar_wait(bar_full + 8 * s, ph);
          relu_sweep(ring_base + s * stage_bytes, MT * TC_A_BYTES, et);
          mbar_arrive(bar_relu + 8 * s);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ===== epilogue: warps 6..9 (TMEM lane quadrants 2,3,0,1) =====
    int cur_n = ((int)blockIdx.x * MT) / tiles_per_img;
    if (p.bias_img_stride != 0) pdl_wait();   // per-image bias = activation data of the previous kernel
    stage_bias(bias_s, p, threadIdx.x - 192, 128, cur_n);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (p.bias_img_stride == 0) pdl_wait();
    int ti = 0;
    for (int uj = 0; uj < units_per_cta; ++uj, ++ti) {
      const int u = (int)blockIdx.x + uj * (int)gridDim.x;
      const int ab = nbuf == 2 ? (ti & 1) : 0;
      const uint32_t tph = nbuf == 2 ? ((uint32_t)(ti >> 1) & 1u) : ((uint32_t)ti & 1u);
      mbar_wait_relaxed(bar_tfull + 8 * ab, tph);
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const int tile = u * MT + m;
        if (tile >= p.n_tiles) break;
        const int n = tile / tiles_per_img, r = tile - n * tiles_per_img;
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        if (p.bias_img_stride != 0 && n != cur_n) {              // per-image bias: restage when the image changes
          asm volatile("bar.sync 1, 128;" ::: "memory");
          stage_bias(bias_s, p, threadIdx.x - 192, 128, n);
          asm volatile("bar.sync 1, 128;" ::: "memory");
          cur_n = n;
        }
        epilogue_store(p, tmem_base + (uint32_t)(ab * MT + m) * tmem_cols, bias_s, warp, lane, n, ty * BH, tx * BW);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * ab);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CL > 1) cluster_sync_all();          // no CTA leaves while its peer may still multicast into it / arrive on its barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)nbuf * MT * tmem_cols) : "memory");
  }
}

// ---- MMA issue for one 64-channel chunk of a halo tile ---------------------------------------------
// The issuing thread is the pipeline's metronome.  With run-time tap / k-step loops it spent ~16 instructions
// (~160 cycles) per tcgen05.mma while a 128xNx16 MMA with N <= 96 occupies the tensor pipe for only 24-48 cycles
// (ncu r01p: tensor pipe 19 % busy on stem1, all samples of the issuing warp inside the loop, none in waits).
// Fully unrolled per (KH, KW, KSTEPS) the descriptor arithmetic is a handful of independent uniform adds.
template <int KH, int KW, int KS>
__device__ __forceinline__ void issue_chunk_resident(uint32_t tmem_d, uint64_t adesc0, uint64_t bdesc0, uint32_t row_step,
                                                     uint32_t tap_step, uint32_t bidx_step, uint32_t idesc, uint32_t acc_first) {
#pragma unroll
  for (int ky = 0; ky < KH; ++ky) {
#pragma unroll
    for (int kx = 0; kx < KW; ++kx) {
      const uint64_t a = adesc0 + (uint32_t)(ky * row_step + kx * tap_step);
      const uint64_t b = bdesc0 + (uint32_t)((ky * KW + kx) * bidx_step);
#pragma unroll
      for (int k = 0; k < KS; ++k)
        umma_bf16(tmem_d, a + 2 * k, b + 2 * k, idesc, (ky == 0 && kx == 0 && k == 0) ? acc_first : 1u);
    }
  }
}

// returns false when (kh, kw, ksteps) has no unrolled instance (caller falls back to the generic loop)
__device__ __forceinline__ bool issue_chunk_resident_dispatch(int kh, int kw, int ks, uint32_t tmem_d, uint64_t adesc0,
                                                              uint64_t bdesc0, uint32_t row_step, uint32_t tap_step,
                                                              uint32_t bidx_step, uint32_t idesc, uint32_t acc_first) {
#define ICR(KH_, KW_, KS_) if (kh == KH_ && kw == KW_ && ks == KS_) { \
    issue_chunk_resident<KH_, KW_, KS_>(tmem_d, adesc0, bdesc0, row_step, tap_step, bidx_step, idesc, acc_first); return true; }
  ICR(3, 3, 4) ICR(3, 3, 3) ICR(3, 3, 2) ICR(3, 3, 1)
#undef ICR
  return false;
}

// streamed weights: one ring slot per tap; k-steps unrolled
template <int KS>
__device__ __forceinline__ void issue_tap_ring(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc_first) {
#pragma unroll
  for (int k = 0; k < KS; ++k) umma_bf16(tmem_d, a + 2 * k, b + 2 * k, idesc, k == 0 ? acc_first : 1u);
}
__device__ __forceinline__ void issue_tap_ring_dispatch(int ks, uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc_first) {
  switch (ks) {
    case 4: issue_tap_ring<4>(tmem_d, a, b, idesc, acc_first); break;
    case 3: issue_tap_ring<3>(tmem_d, a, b, idesc, acc_first); break;
    case 2: issue_tap_ring<2>(tmem_d, a, b, idesc, acc_first); break;
    default: issue_tap_ring<1>(tmem_d, a, b, idesc, acc_first); break;
  }
}

// ---- halo-resident kernel (stride 1, BW = 128, BH = 1) ---------------------------------------------
// The per-tap kernel above re-fetches the 128 x 64 A tile from L2 once per tap (9x / 25x).  Here the
// kh halo rows of the current 64-channel chunk are brought into shared memory ONCE (one TMA box per
// row, zero-filled outside the image = the conv padding) and every tap's A operand is the same
// buffer addressed through a shifted UMMA descriptor: tap (ky,kx) starts (ky*pitch + kx*dil) pixels
// (128 B each) into the buffer.  TMA's SWIZZLE_128B is a function of the shared-memory address bits,
// which is also what the UMMA read side applies, so a 128-byte-granular shift stays consistent.
// ReLU-on-load is one in-place sweep per halo chunk instead of one per tap.
// Warps (224 threads): 0 = halo producer, 1 = MMA issuer (+TMEM alloc), 2 = weight producer,
// 3..6 = ReLU sweep + epilogue.
constexpr int TCH_THREADS = 224;

__device__ __forceinline__ uint64_t make_kmajor_sw128_desc_shifted(uint32_t saddr, int mode) {
  uint64_t d = make_kmajor_sw128_desc(saddr);
  if (mode == 1) d |= (uint64_t)((saddr >> 7) & 7u) << 49;    // base_offset: phase of the 1024-B swizzle pattern
  return d;
}

__global__ void __launch_bounds__(TCH_THREADS)
conv2d_tc_halo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[6 + 2 * TC_MAX_STAGES + 1];   // halo full/empty/relu [2], b_full[s], b_empty[s], accum
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float bias_s[TC_MAX_NPAD];
  // a per-image bias (ASPP pool branch) is ACTIVATION data written by the previous kernel: wait for it first
  if (p.bias_img_stride != 0) pdl_wait();
  stage_bias(bias_s, p, threadIdx.x, TCH_THREADS, (int)(blockIdx.x / (unsigned)(p.tiles_x * p.tiles_y)));

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t halo_bytes = (uint32_t)(p.taps / p.taps_w) * p.halo_pitch * 128u;    // kh rows
  const uint32_t b_base = smem_base + p.halo_bufs * halo_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool relu_in = (p.flags & ADD_RELU_IN) != 0;
  const int kh = p.taps / p.taps_w;

  const uint32_t bar_hfull = smem_u32(&bars[0]);
  const uint32_t bar_hempty = smem_u32(&bars[2]);
  const uint32_t bar_hrelu = smem_u32(&bars[4]);
  const uint32_t bar_bfull = smem_u32(&bars[6]);
  const uint32_t bar_bempty = smem_u32(&bars[6 + TC_MAX_STAGES]);
  const uint32_t bar_accum = smem_u32(&bars[6 + 2 * TC_MAX_STAGES]);

  int t = blockIdx.x;
  const int tx = t % p.tiles_x; t /= p.tiles_x;
  const int ty = t % p.tiles_y; const int n = t / p.tiles_y;
  const int x0 = tx * TC_BM, y0 = ty;

  if (threadIdx.x == 0) {
    for (int h = 0; h < 2; ++h) {
      mbar_init(bar_hfull + 8 * h, 1);
      mbar_init(bar_hempty + 8 * h, 1);
      mbar_init(bar_hrelu + 8 * h, 128);
    }
    for (int s2 = 0; s2 < p.stages; ++s2) {
      mbar_init(bar_bfull + 8 * s2, 1);
      mbar_init(bar_bempty + 8 * s2, 1);
    }
    mbar_init(bar_accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_smem)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  pdl_launch_dependents();
  pdl_wait();                     // the previous kernel's activations are complete and visible from here on

  if (warp == 0) {
    // ===== halo producer: kh row boxes per 64-channel chunk =====
    if (elect_one()) {
      int hb = 0; uint32_t ph = 0;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(bar_hempty + 8 * hb, ph ^ 1);
        mbar_expect_tx(bar_hfull + 8 * hb, halo_bytes);
        const uint32_t dst = smem_base + hb * halo_bytes;
        for (int ky = 0; ky < kh; ++ky)
          tma_load_4d(dst + ky * p.halo_pitch * 128, &map_x, bar_hfull + 8 * hb, kc * TC_BK, x0 - p.pad,
                      y0 - p.pad + ky * p.dil, n);
        if (++hb == p.halo_bufs) { hb = 0; ph ^= 1; }
      }
    }
  } else if (warp == 2) {
    // ===== weight producer: one [n_pad x 64] K-major tile per (chunk, tap) =====
    if (elect_one()) {
      int s2 = 0; uint32_t ph = 0;
      for (int kc = 0; kc < p.kchunks; ++kc)
        for (int tap = 0; tap < p.taps; ++tap) {
          mbar_wait(bar_bempty + 8 * s2, ph ^ 1);
          mbar_expect_tx(bar_bfull + 8 * s2, p.b_bytes);
          tma_load_3d(b_base + s2 * p.b_bytes, &map_w, bar_bfull + 8 * s2, 0, 0, tap * p.kchunks + kc);
          if (++s2 == p.stages) { s2 = 0; ph ^= 1; }
        }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_pad >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int hb = 0; uint32_t hph = 0; int s2 = 0; uint32_t bph = 0;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        const int krem = p.Cin - kc * TC_BK;
        const int ksteps = krem >= TC_BK ? TC_BK / 16 : (krem + 15) / 16;
        mbar_wait((relu_in ? bar_hrelu : bar_hfull) + 8 * hb, hph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t halo = smem_base + hb * halo_bytes;
        const uint64_t adesc0 = make_kmajor_sw128_desc_shifted(halo, p.base_off_mode);
        const uint64_t bdesc0 = make_kmajor_sw128_desc(b_base);
        const uint32_t row_step = (uint32_t)p.halo_pitch * 8u, tap_step = (uint32_t)p.dil * 8u, bstep = p.b_bytes >> 4;
        uint32_t acc = kc > 0 ? 1u : 0u;
        uint32_t a_row = 0;
        for (int ky = 0; ky < kh; ++ky, a_row += row_step) {
          uint32_t a_off = a_row;
          for (int kx = 0; kx < p.taps_w; ++kx, a_off += tap_step) {
            mbar_wait(bar_bfull + 8 * s2, bph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue_tap_ring_dispatch(ksteps, tmem_base, adesc0 + a_off, bdesc0 + (uint32_t)s2 * bstep, idesc, acc);
            acc = 1u;
            umma_commit(bar_bempty + 8 * s2);
            if (++s2 == p.stages) { s2 = 0; bph ^= 1; }
          }
        }
        umma_commit(bar_hempty + 8 * hb);     // halo buffer free once this chunk's MMAs have read it
        if (++hb == p.halo_bufs) { hb = 0; hph ^= 1; }
      }
      umma_commit(bar_accum);
    }
  } else if (warp >= 3) {
    // ===== ReLU sweep (one per halo chunk) + epilogue: warps 3..6 =====
    const int et = threadIdx.x - 96;        // 0..127
    if (relu_in) {
      int hb = 0; uint32_t ph = 0;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(bar_hfull + 8 * hb, ph);
        relu_sweep(smem_base + hb * halo_bytes, halo_bytes, et);
        mbar_arrive(bar_hrelu + 8 * hb);
        if (++hb == p.halo_bufs) { hb = 0; ph ^= 1; }
      }
    }
    mbar_wait(bar_accum, 0);
    epilogue_store(p, tmem_base, bias_s, warp, lane, n, y0, x0);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// ---- persistent halo-resident kernel -----------------------------------------------------------------
// Same A-operand scheme as conv2d_tc_halo_kernel (halo rows resident, taps through shifted descriptors), but a
// CTA loops over tiles: two halo buffers form a ring over (tile, channel chunk) so the next tile's halo streams in
// while the current one is multiplied, the accumulator is double-buffered in TMEM (epilogue of tile i overlaps the
// MMAs of tile i+1), and when all taps' weights fit they are loaded ONCE and stay resident — the one-tile-per-CTA
// kernel re-streamed every weight tile from L2 for each 128-pixel row (25 x 6 KB for a 5x5 at C=40).
//   warp 0 halo producer | warp 1 MMA issuer (+TMEM alloc) | warp 2 weight producer | warps 3..6 ReLU sweep |
//   warps 7..10 epilogue (TMEM lane quadrants 3,0,1,2)
constexpr int TCHP_THREADS = 352;
constexpr int TCHP_MAX_STAGES = 16;          // weight ring depth when the taps do not fit resident

// Row pairs (p.mt == 2): a unit is two output rows r and r + dil of one 128-pixel column strip.  They share kh-1 of
// their kh halo rows (kh + 1 rows are loaded instead of 2 kh) and EVERY weight tile: for the 5x5s the weight stream
// (25 x 6 KB per tile through a ring that smem limits to ~36 KB in flight = 37 GB/s per SM, ncu r01u) was the
// bound, so two tiles per weight tile doubles the useful work per streamed byte.
// Row quads (p.mt == 4, r02e): rows r, r+dil, r+2dil, r+3dil share kh+3 halo rows and every weight tile (a quarter of
// the weight stream per output row, (kh+3)/4 halo rows per output row instead of kh).
// Row-granular halo barriers (r02e): with streamed weights only ONE halo buffer fits next to the ring, and with one
// full/empty barrier per buffer the next unit's halo load and ReLU sweep could not start before the last MMA of the
// current unit had retired (load -> sweep -> MMAs strictly in series: 11 us per row pair of a 5x5 at C=40, of which
// ~3.5 us MMA).  Each halo ROW now has its own full / relu / empty barriers: the issuer releases row ky as soon as the
// taps of kernel row ky are issued (no later tap reads it), the producer refills it for the next unit and the sweep
// follows row by row, so load, sweep and MMAs of consecutive units overlap inside a single buffer.
constexpr int TCHP_MAX_ROWS = 8;             // halo rows per buffer: kh + mt - 1 <= 8
__device__ __forceinline__ void halo_unit_rows(int u, int units_per_col, int tiles_x, int dil, int mt, int& n, int& tx, int& r0) {
  // u -> (image n, column strip tx, first output row r0); mt > 1: rows are grouped (r0, r0 + dil, .., r0 + (mt-1) dil),
  // r0 = blk*mt*dil + off
  const int per_img = units_per_col * tiles_x;
  n = u / per_img;
  const int r = u - n * per_img;
  const int j = r / tiles_x;
  tx = r - j * tiles_x;
  r0 = mt > 1 ? (j / dil) * mt * dil + (j % dil) : j;
}

// One (unit, channel chunk) of the persistent halo kernel: every tap of the kh x kw kernel for MT output rows, with
// streamed or resident weights and row-group halo barriers.  Templated on (MT, KS) so the MT x KS MMAs of a tap are
// straight-line code: the single issuing thread is the pipeline's metronome, and with the run-time `switch (ksteps)`
// plus a run-time row loop per tap it spent ~87 cycles per tcgen05.mma that occupies the tensor pipe for 24
// (ncu r3d: tensor pipe 24 % busy on the 5x5 at C=40, issuer warp sampled in issue code, not in waits).
struct HaloIssueState { int s2; uint32_t bph; };
template <int MT, int KS>
__device__ __forceinline__ void halo_issue_chunk(const TcParams& p, int kh, int G, int hrows, uint32_t bar_ready, uint32_t bar_done,
                                                 uint32_t hph, uint32_t bar_bfull, uint32_t bar_bempty, HaloIssueState& st,
                                                 uint32_t tmem_d, uint32_t tmem_cols, uint64_t adesc0, uint64_t bdesc0,
                                                 uint32_t bidx, uint32_t bidx_step, uint32_t bstep, uint32_t row_step,
                                                 uint32_t tap_step, uint32_t idesc, uint32_t acc) {
  int ready = 0;                                // halo row groups [0, ready) of this buffer are loaded (and swept)
  uint32_t a_row = 0;
  const int kw = p.taps_w;
  const bool resident = p.b_resident != 0;
  for (int ky = 0; ky < kh; ++ky, a_row += row_step) {
    const int need = G == 1 ? ky + MT : 1;      // kernel row ky reads halo rows ky .. ky + MT - 1
    if (ready < need) {
      for (; ready < need; ++ready) mbar_wait(bar_ready + 8 * ready, hph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    uint32_t a_off = a_row;
    for (int kx = 0; kx < kw; ++kx, a_off += tap_step) {
      uint64_t bdesc;
      if (resident) {
        bdesc = bdesc0 + bidx; bidx += bidx_step;
      } else {
        mbar_wait(bar_bfull + 8 * st.s2, st.bph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        bdesc = bdesc0 + (uint32_t)st.s2 * bstep;
      }
      const uint64_t a = adesc0 + a_off;
#pragma unroll
      for (int k = 0; k < KS; ++k) {
#pragma unroll
        for (int m = 0; m < MT; ++m)            // output rows share the weight k-slice: halo rows m further down
          umma_bf16(tmem_d + (uint32_t)m * tmem_cols, a + (uint32_t)m * row_step + 2 * k, bdesc + 2 * k, idesc, k == 0 ? acc : 1u);
      }
      acc = 1u;
      if (!resident) {
        umma_commit(bar_bempty + 8 * st.s2);
        if (++st.s2 == p.stages) { st.s2 = 0; st.bph ^= 1; }
      }
    }
    // no later tap reads halo row ky: free it once the MMAs issued so far have retired (the last kernel row frees the
    // remaining MT - 1 rows too); a whole-buffer group is freed after the last kernel row
    if (G == 1) {
      umma_commit(bar_done + 8 * ky);
      if (ky == kh - 1)
        for (int j = kh; j < hrows; ++j) umma_commit(bar_done + 8 * j);
    } else if (ky == kh - 1) {
      umma_commit(bar_done);
    }
  }
}

__global__ void __launch_bounds__(TCHP_THREADS, 1)
conv2d_tc_halo_persistent_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // halo full/empty/relu [2][rows], b_full[s], b_empty[s], tfull[2], tempty[2], bres
  constexpr int HB = 2 * TCHP_MAX_ROWS;
  __shared__ __align__(8) uint64_t bars[3 * HB + 2 * TCHP_MAX_STAGES + 5];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float bias_s[TC_MAX_NPAD];

  const int kh = p.taps / p.taps_w;
  const int MT = p.mt;
  const int hrows = kh + (MT - 1);                                         // halo rows per (unit, chunk)
  const int iters = p.taps * p.kchunks;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t row_bytes = (uint32_t)p.halo_pitch * 128u;
  const uint32_t halo_bytes = (uint32_t)hrows * row_bytes;
  // barrier granularity: one group of G halo rows per full/relu/empty barrier.  Two halo buffers already overlap
  // load / sweep / MMAs of consecutive units, so the buffer is one group (one wait and one commit per chunk, as
  // before); a single buffer is released row by row (G = 1).
  const int G = p.halo_bufs >= 2 ? hrows : 1;
  const int ngroups = hrows / G;
  const uint32_t group_bytes = (uint32_t)G * row_bytes;
  const uint32_t b_base = smem_base + (uint32_t)p.halo_bufs * halo_bytes;  // ring slots, or the resident image
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool relu_in = (p.flags & ADD_RELU_IN) != 0;
  const uint32_t tmem_cols = (uint32_t)p.tmem_cols;
  const int nbuf = (2u * MT * tmem_cols <= 512u) ? 2 : 1;                  // accumulator sets

  // per-row halo barriers: index (buffer hb, row j) -> 8 * (hb * TCHP_MAX_ROWS + j)
  const uint32_t bar_hfull = smem_u32(&bars[0]);
  const uint32_t bar_hempty = smem_u32(&bars[HB]);
  const uint32_t bar_hrelu = smem_u32(&bars[2 * HB]);
  const uint32_t bar_bfull = smem_u32(&bars[3 * HB]);
  const uint32_t bar_bempty = smem_u32(&bars[3 * HB + TCHP_MAX_STAGES]);
  const uint32_t bar_tfull = smem_u32(&bars[3 * HB + 2 * TCHP_MAX_STAGES]);
  const uint32_t bar_tempty = smem_u32(&bars[3 * HB + 2 * TCHP_MAX_STAGES + 2]);
  const uint32_t bar_bres = smem_u32(&bars[3 * HB + 2 * TCHP_MAX_STAGES + 4]);

  if (threadIdx.x == 0) {
    for (int h = 0; h < HB; ++h) {
      mbar_init(bar_hfull + 8 * h, 1);
      mbar_init(bar_hempty + 8 * h, 1);
      mbar_init(bar_hrelu + 8 * h, 128);
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(bar_tfull + 8 * h, 1);
      mbar_init(bar_tempty + 8 * h, 4);
    }
    for (int s2 = 0; s2 < p.stages; ++s2) {
      mbar_init(bar_bfull + 8 * s2, 1);
      mbar_init(bar_bempty + 8 * s2, 1);
    }
    mbar_init(bar_bres, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_smem)), "r"((uint32_t)nbuf * MT * tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  pdl_launch_dependents();
  // units: (image, column strip, row or row pair); p.n_tiles holds the unit count, p.tiles_y the units per column
  const int n_units = p.n_tiles, upc = p.tiles_y;

  if (warp == 0) {
    // ===== halo producer: hrows row boxes per (unit, 64-channel chunk) =====
    if (elect_one()) {
      pdl_wait();
      int hb = 0; uint32_t ph = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        int n, tx, r0;
        halo_unit_rows(u, upc, p.tiles_x, p.dil, MT, n, tx, r0);
        for (int kc = 0; kc < p.kchunks; ++kc) {
          const uint32_t dst = smem_base + hb * halo_bytes;
          const uint32_t bo = 8u * (uint32_t)(hb * TCHP_MAX_ROWS);
          for (int g = 0; g < ngroups; ++g) {
            mbar_wait(bar_hempty + bo + 8 * g, ph ^ 1);
            mbar_expect_tx(bar_hfull + bo + 8 * g, group_bytes);
            for (int j = g * G; j < (g + 1) * G; ++j)
              tma_load_4d(dst + j * row_bytes, &map_x, bar_hfull + bo + 8 * g, kc * TC_BK, tx * TC_BM - p.pad,
                          r0 - p.pad + j * p.dil, n);
          }
          if (++hb == p.halo_bufs) { hb = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ===== weight producer: resident image once, or one [n_pad x 64] tile per (unit, chunk, tap) through the ring =====
    if (elect_one()) {
      if (p.b_resident) {
        mbar_expect_tx(bar_bres, (uint32_t)iters * p.b_bytes);
        for (int it = 0; it < iters; ++it) tma_load_3d(b_base + it * p.b_bytes, &map_w, bar_bres, 0, 0, it);
      } else {
        int s2 = 0; uint32_t ph = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x)
          for (int kc = 0; kc < p.kchunks; ++kc)
            for (int tap = 0; tap < p.taps; ++tap) {
              mbar_wait(bar_bempty + 8 * s2, ph ^ 1);
              mbar_expect_tx(bar_bfull + 8 * s2, p.b_bytes);
              tma_load_3d(b_base + s2 * p.b_bytes, &map_w, bar_bfull + 8 * s2, 0, 0, tap * p.kchunks + kc);
              if (++s2 == p.stages) { s2 = 0; ph ^= 1; }
            }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_pad >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      if (p.b_resident) mbar_wait(bar_bres, 0);
      int hb = 0; uint32_t hph = 0; int ti = 0;
      HaloIssueState ist; ist.s2 = 0; ist.bph = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++ti) {
        const int ab = nbuf == 2 ? (ti & 1) : 0;
        const uint32_t tph = nbuf == 2 ? ((uint32_t)(ti >> 1) & 1u) : ((uint32_t)ti & 1u);
        mbar_wait(bar_tempty + 8 * ab, tph ^ 1u);
        for (int kc = 0; kc < p.kchunks; ++kc) {
          const int krem = p.Cin - kc * TC_BK;
          const int ksteps = krem >= TC_BK ? TC_BK / 16 : (krem + 15) / 16;
          const uint32_t bar_ready = (relu_in ? bar_hrelu : bar_hfull) + 8u * (uint32_t)(hb * TCHP_MAX_ROWS);
          const uint32_t bar_done = bar_hempty + 8u * (uint32_t)(hb * TCHP_MAX_ROWS);
          // The single issuing thread is the pipeline's metronome: keep its per-MMA instruction count minimal.
          // Descriptors differ only in their 14-bit address field.
          const uint32_t halo = smem_base + hb * halo_bytes;
          const uint64_t adesc0 = make_kmajor_sw128_desc_shifted(halo, p.base_off_mode);
          const uint64_t bdesc0 = make_kmajor_sw128_desc(b_base);
          const uint32_t row_step = (uint32_t)p.halo_pitch * 8u, tap_step = (uint32_t)p.dil * 8u;   // in 16-byte units
          const uint32_t bstep = p.b_bytes >> 4;
          uint32_t bidx = p.b_resident ? (uint32_t)kc * bstep : 0u;              // resident: tile (tap*kchunks+kc)
          const uint32_t bidx_step = (uint32_t)p.kchunks * bstep;
          uint32_t acc = kc > 0 ? 1u : 0u;
          const uint32_t tmem_d = tmem_base + (uint32_t)(ab * MT) * tmem_cols;
          const bool unrolled = MT == 1 && p.b_resident && kh == 3 && p.taps_w == 3;
          if (unrolled) {
            for (int g = 0; g < ngroups; ++g) mbar_wait(bar_ready + 8 * g, hph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          if (unrolled && issue_chunk_resident_dispatch(kh, p.taps_w, ksteps, tmem_d, adesc0, bdesc0 + bidx, row_step,
                                                        tap_step, bidx_step, idesc, acc)) {
            // fully unrolled instance issued
            for (int g = 0; g < ngroups; ++g) umma_commit(bar_done + 8 * g);
          } else {
#define HIC(MT_, KS_) case MT_ * 8 + KS_: halo_issue_chunk<MT_, KS_>(p, kh, G, hrows, bar_ready, bar_done, hph, bar_bfull, bar_bempty, \
                                                          ist, tmem_d, tmem_cols, adesc0, bdesc0, bidx, bidx_step, bstep, row_step, tap_step, idesc, acc); break;
            switch (MT * 8 + ksteps) {
              HIC(1, 1) HIC(1, 2) HIC(1, 3) HIC(1, 4)
              HIC(2, 1) HIC(2, 2) HIC(2, 3) HIC(2, 4)
              HIC(4, 1) HIC(4, 2) HIC(4, 3) HIC(4, 4)
              default: break;
            }
#undef HIC
          }
          if (++hb == p.halo_bufs) { hb = 0; hph ^= 1; }
        }
        umma_commit(bar_tfull + 8 * ab);
      }
    }
  } else if (warp < 7) {
    // ===== ReLU sweep (one per halo chunk): warps 3..6 =====
    if (relu_in) {
      const int et = threadIdx.x - 96;        // 0..127
      int hb = 0; uint32_t ph = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x)
        for (int kc = 0; kc < p.kchunks; ++kc) {
          const uint32_t bo = 8u * (uint32_t)(hb * TCHP_MAX_ROWS);
          for (int g = 0; g < ngroups; ++g) {
            mbar_wait(bar_hfull + bo + 8 * g, ph);
            relu_sweep(smem_base + hb * halo_bytes + g * group_bytes, group_bytes, et);
            mbar_arrive(bar_hrelu + bo + 8 * g);
          }
          if (++hb == p.halo_bufs) { hb = 0; ph ^= 1; }
        }
    }
  } else {
    // ===== epilogue: warps 7..10 =====
    int cur_n = 0;
    { int tx0, r00; halo_unit_rows((int)blockIdx.x < n_units ? (int)blockIdx.x : 0, upc, p.tiles_x, p.dil, MT, cur_n, tx0, r00); }
    if (p.bias_img_stride != 0) pdl_wait();   // per-image bias = activation data of the previous kernel
    stage_bias(bias_s, p, threadIdx.x - 224, 128, cur_n);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (p.bias_img_stride == 0) pdl_wait();                               // += y reads and y writes must follow the previous kernel
    int ti = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++ti) {
      const int ab = nbuf == 2 ? (ti & 1) : 0;
      const uint32_t tph = nbuf == 2 ? ((uint32_t)(ti >> 1) & 1u) : ((uint32_t)ti & 1u);
      int n, tx, r0;
      halo_unit_rows(u, upc, p.tiles_x, p.dil, MT, n, tx, r0);
      if (p.bias_img_stride != 0 && n != cur_n) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        stage_bias(bias_s, p, threadIdx.x - 224, 128, n);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        cur_n = n;
      }
      mbar_wait_relaxed(bar_tfull + 8 * ab, tph);
      for (int m = 0; m < MT; ++m) {
        const int row = r0 + m * p.dil;
        if (row < p.Ho)                        // rows past the image (odd remainder of the pairing) computed on zeros, not stored
          epilogue_store(p, tmem_base + (uint32_t)(ab * MT + m) * tmem_cols, bias_s, warp, lane, n, row, tx * TC_BM);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * ab);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)nbuf * MT * tmem_cols) : "memory");
  }
}

// ---- host side -----------------------------------------------------------------------------------
int g_conv_cluster = 1;             // 1 = cluster-of-2 weight multicast for the N_pad = 256 convs (mode bit 8 clears)
int g_halo_pairs = 1;               // 1 = row pairs in the persistent halo kernel (mode bit 7 clears)
int g_halo_quads = 1;               // 1 = row quads for the streamed-weight 5x5s (mode bit 9 clears)
int g_conv_mt2 = 1;                 // 1 = two pixel tiles per weight tile for the weight-heavy convs (mode bit 6 clears)
int g_conv_flatten_1x1 = 1;         // 1 = stride-1 1x1 convs tile the flattened N*H*W pixel range (mode bit 11 clears)
int g_conv_mt2_small = 0;           // 1 = two pixel tiles per stage also for the resident-weight 1x1s (mode bit 10 sets; experiment)
int g_halo_stream_persistent = 0;   // 1 = use the persistent halo kernel also when the weights stream through a ring (tuning)
int g_persistent = 1;  // 1 = persistent warp-specialised kernel for the non-halo path (default), 0 = one tile per CTA
int g_halo_mode = 1;   // 0 = per-tap A tiles only, 1 = halo-resident A (measured on B200: base_offset must stay 0 — the
                       // swizzle is applied on absolute shared-memory address bits; mode 2 (base_offset = phase) is WRONG
                       // and kept only as the recorded experiment)


}  // namespace

extern "C" int64_t add_conv2d_tc_packed_bytes(int cin, int cout, int kh, int kw) {
  if (cin <= 0 || cout <= 0 || kh <= 0 || kw <= 0) return ADD_ERR_BAD_ARG;
  if (cout > 256) return ADD_ERR_UNSUPPORTED;
  return (int64_t)kh * kw * kchunks_of(cin) * npad_of(cout) * TC_BK * 2;
}

// w_hwio: fp32 [kh][kw][cin][cout] (BN scale folded) -> bf16 [kh*kw*kchunks][n_pad][64], zero padded.
extern "C" int add_conv2d_tc_pack(const float* w_hwio, int cin, int cout, int kh, int kw, void* packed_host) {
  ADD_CHECK_ARG(w_hwio && packed_host && cin > 0 && cout > 0 && kh > 0 && kw > 0);
  ADD_CHECK_SUP(cout <= 256);
  const int kc_n = kchunks_of(cin), n_pad = npad_of(cout);
  uint16_t* out = static_cast<uint16_t*>(packed_host);
  std::memset(out, 0, (size_t)kh * kw * kc_n * n_pad * TC_BK * 2);
  for (int tap = 0; tap < kh * kw; ++tap)
    for (int ci = 0; ci < cin; ++ci) {
      const int kc = ci / TC_BK, k = ci % TC_BK;
      const float* src = w_hwio + ((size_t)tap * cin + ci) * cout;
      uint16_t* dst = out + ((size_t)(tap * kc_n + kc) * n_pad) * TC_BK + k;
      for (int co = 0; co < cout; ++co) {
        uint32_t u; std::memcpy(&u, &src[co], 4);
        uint32_t rnd = 0x7FFFu + ((u >> 16) & 1u);           // round to nearest even
        dst[(size_t)co * TC_BK] = (uint16_t)((u + rnd) >> 16);
      }
    }
  return ADD_OK;
}

extern "C" int add_conv2d_tc_fwd(const add_tensor_t* x, const add_tensor_t* y, const void* w_packed,
                                 const float* bias, int64_t bias_image_stride, int kh, int kw, int stride, int pad, int dil,
                                 uint32_t flags, void* stream) {
  ADD_CHECK_ARG(tensor_ok(x) && tensor_ok(y) && w_packed);
  ADD_CHECK_ARG(kh > 0 && kw > 0 && stride > 0 && dil > 0 && x->n == y->n);
  ADD_CHECK_ARG(conv_extent_ok(x->h, y->h, kh, stride, pad, dil) && conv_extent_ok(x->w, y->w, kw, stride, pad, dil));
  ADD_CHECK_SUP(x->dtype == ADD_BF16 && y->c <= 256 && stride <= 2);
  // TMA: 16-byte aligned base and strides
  ADD_CHECK_SUP(x->pix_stride % 8 == 0 && ((uintptr_t)x->ptr % 16) == 0 && ((uintptr_t)w_packed % 16) == 0);
  if (y->dtype == ADD_BF16) ADD_CHECK_SUP(y->pix_stride % 8 == 0 && ((uintptr_t)y->ptr % 16) == 0);
  else ADD_CHECK_SUP(y->pix_stride % 4 == 0 && ((uintptr_t)y->ptr % 16) == 0);
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) { g_add_last_cuda_error = (int)cudaErrorSymbolNotFound; return ADD_ERR_CUDA; }

  // A stride-1 1x1 conv does not care where rows or images end: it is a GEMM over the N*H*W pixels.  Described as ONE
  // row of N*H*W pixels, every 128-pixel tile is full except the last — with (row, 128-column) tiles a 129-wide map
  // (the stride-8 level of the authors' 1025x2049 evaluation size) wastes half of every second tile.
  add_tensor_t xf, yf;
  if (g_conv_flatten_1x1 && kh == 1 && kw == 1 && stride == 1 && pad == 0 && !(bias && bias_image_stride != 0) &&
      x->h == y->h && x->w == y->w && (long long)x->n * x->h * x->w < (1ll << 31)) {
    xf = *x; yf = *y;
    xf.w = yf.w = x->n * x->h * x->w;
    xf.n = yf.n = 1; xf.h = yf.h = 1;
    x = &xf; y = &yf;
  }

  TcParams p;
  p.bias_img_stride = bias ? bias_image_stride : 0;
  p.y = y->ptr; p.bias = bias; p.Ho = y->h; p.Wo = y->w; p.Cout = y->c; p.ys = y->pix_stride;
  p.y_is_f32 = (y->dtype == ADD_F32);
  int bw_log2 = 3;                                   // BW = smallest power of two >= Wo, clamped to [8, 128]
  while ((1 << bw_log2) < y->w && bw_log2 < 7) ++bw_log2;
  p.bw_log2 = bw_log2;
  const int BW = 1 << bw_log2, BH = TC_BM / BW;
  p.tiles_x = ceil_div(y->w, BW); p.tiles_y = ceil_div(y->h, BH);
  p.taps_w = kw; p.taps = kh * kw; p.kchunks = kchunks_of(x->c); p.Cin = x->c;
  p.stride = stride; p.pad = pad; p.dil = dil;
  p.n_pad = npad_of(y->c);
  p.tmem_cols = 32; while (p.tmem_cols < p.n_pad) p.tmem_cols <<= 1;
  p.b_bytes = (uint32_t)p.n_pad * 128u;
  const uint32_t stage_bytes = TC_A_BYTES + p.b_bytes;
  int stages = (int)((100u * 1024u) / stage_bytes);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages < 2) stages = 2;
  if (stages > p.taps * p.kchunks) stages = p.taps * p.kchunks;
  if (stages < 1) stages = 1;
  p.stages = stages; p.flags = flags;

  // halo-resident mode: stride-1 multi-tap convs on 128-pixel row tiles (the large feature maps)
  int halo_mode = g_halo_mode;
  // (N_pad > 160: the per-tap pipeline is already MMA-bound and keeps two CTAs per SM)
  const bool halo = halo_mode > 0 && stride == 1 && p.taps > 1 && bw_log2 == 7 && p.n_pad <= 160;
  p.halo_pitch = 0; p.halo_bufs = 0; p.base_off_mode = halo_mode == 2 ? 1 : 0;
  size_t smem = 0;
  if (halo) {
    p.halo_pitch = round_up(TC_BM + (kw - 1) * dil, 8);
    const uint32_t halo_bytes = (uint32_t)kh * p.halo_pitch * 128u;
    p.halo_bufs = (p.kchunks > 1 && 2 * halo_bytes + 2 * p.b_bytes <= 110u * 1024u) ? 2 : 1;
    int sb = (int)((110u * 1024u - (uint32_t)p.halo_bufs * halo_bytes) / p.b_bytes);
    if (sb > TC_MAX_STAGES) sb = TC_MAX_STAGES;
    if (sb < 2) sb = 2;
    p.stages = sb;
    smem = (size_t)p.halo_bufs * halo_bytes + (size_t)sb * p.b_bytes + 1024;
    if (p.halo_pitch > 256 || smem > 200u * 1024u) return ADD_ERR_UNSUPPORTED;
  } else {
    smem = (size_t)stages * stage_bytes + 1024;     // + alignment slack
  }

  CUtensorMap map_x, map_w;
  {
    cuuint64_t dims[4] = {(cuuint64_t)x->c, (cuuint64_t)x->w, (cuuint64_t)x->h, (cuuint64_t)x->n};
    cuuint64_t strides[3] = {(cuuint64_t)x->pix_stride * 2, (cuuint64_t)x->w * x->pix_stride * 2,
                             (cuuint64_t)x->h * x->w * x->pix_stride * 2};
    cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)(BW * stride), (cuuint32_t)(BH * stride), 1};
    if (halo) { box[1] = (cuuint32_t)p.halo_pitch; box[2] = 1; }
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    if (encode(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x->ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ADD_ERR_UNSUPPORTED;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)TC_BK, (cuuint64_t)p.n_pad, (cuuint64_t)(p.taps * p.kchunks)};
    cuuint64_t strides[2] = {(cuuint64_t)TC_BK * 2, (cuuint64_t)p.n_pad * TC_BK * 2};
    cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)p.n_pad, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (encode(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w_packed), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ADD_ERR_UNSUPPORTED;
  }
  static PerDeviceOnce attr_once;
  once_per_device(attr_once, [] {
    cudaFuncSetAttribute(conv2d_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);   // + static < 227 KB
    cudaFuncSetAttribute(conv2d_tc_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  });
  const long long grid = (long long)p.tiles_x * p.tiles_y * y->n;
  ADD_CHECK_SUP(grid < (1ll << 31));
  if (!halo && g_persistent) {
    const int iters = p.taps * p.kchunks;
    p.n_tiles = (int)grid;
    p.b_resident = ((size_t)iters * p.b_bytes <= 72u * 1024u) ? 1 : 0;
    const size_t fixed = 1024 + (p.b_resident ? (size_t)iters * p.b_bytes : 0);
    int sms = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    // MT = 2 (two pixel tiles per weight tile) for the weight-heavy big convs: streamed weights, N_pad >= 128, many tiles
    int mt = (g_conv_mt2 && !p.b_resident && p.n_pad >= 128 && grid >= 4ll * sms) ? 2 : 1;
    if (g_conv_mt2_small && p.b_resident && p.taps == 1 && grid >= 4ll * sms && 4 * p.tmem_cols <= 512) mt = 2;   // experiment (mode bit 10)
    const size_t sbytes = (size_t)mt * TC_A_BYTES + (p.b_resident ? 0 : p.b_bytes);
    // two CTAs per SM when a 4-deep ring fits in half the shared memory and TMEM (4 accumulators) allows it
    const bool two = mt == 1 && fixed + 4 * sbytes <= 100u * 1024u && 4 * p.tmem_cols <= 512;
    const size_t budget = two ? 100u * 1024u : 200u * 1024u;
    int st = (int)((budget - fixed) / sbytes);
    if (st > TCP_MAX_STAGES) st = TCP_MAX_STAGES;
    if (st >= 2) {
      p.stages = st;
      const size_t psmem = fixed + (size_t)st * sbytes;
      static PerDeviceOnce ponce;
      once_per_device(ponce, [] {
        cudaFuncSetAttribute(conv2d_tc_persistent_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(conv2d_tc_persistent_kernel<1, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaFuncSetAttribute(conv2d_tc_persistent_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(conv2d_tc_persistent_kernel<2, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaFuncSetAttribute(conv2d_tc_persistent_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(conv2d_tc_persistent_kernel<2, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      });
      const long long units = (grid + mt - 1) / mt;
      long long g = (long long)sms * (two ? 2 : 1) * g_add_grid_pct / 100;
      if (units <= 16ll * sms) g = g * g_add_conv_grid_pct / 100;      // small problems: leave SMs to the other stream's kernel
      if (g > units) g = units;
      if (g < 1) g = 1;
      // cluster of 2 with multicast weight halves: the N_pad = 256 convs (ASPP / decoder 3x3) that sit at the L2 ceiling
      const bool cl2 = g_conv_cluster && mt == 2 && p.n_pad == 256 && g >= 2 && units >= 4ll * sms;
      if (cl2) {
        g &= ~1ll;                                        // whole clusters
        CUtensorMap map_wh;
        cuuint64_t dims[3] = {(cuuint64_t)TC_BK, (cuuint64_t)p.n_pad, (cuuint64_t)(p.taps * p.kchunks)};
        cuuint64_t strides[2] = {(cuuint64_t)TC_BK * 2, (cuuint64_t)p.n_pad * TC_BK * 2};
        cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)(p.n_pad / 2), 1};
        cuuint32_t estr[3] = {1, 1, 1};
        if (encode(&map_wh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w_packed), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
          return ADD_ERR_UNSUPPORTED;
        launch_kernel_cluster(conv2d_tc_persistent_kernel<2, 2>, dim3((unsigned)g), dim3(TCP_THREADS), psmem,
                              static_cast<cudaStream_t>(stream), 2, map_x, map_wh, p);
      } else if (mt == 2) {
        launch_kernel(conv2d_tc_persistent_kernel<2, 1>, dim3((unsigned)g), dim3(TCP_THREADS), psmem, static_cast<cudaStream_t>(stream), map_x, map_w, p);
      } else {
        launch_kernel(conv2d_tc_persistent_kernel<1, 1>, dim3((unsigned)g), dim3(TCP_THREADS), psmem, static_cast<cudaStream_t>(stream), map_x, map_w, p);
      }
      ADD_RETURN_LAUNCH();
    }
  }
  if (halo && g_persistent) {
    const int iters = p.taps * p.kchunks;
    const size_t budget = 222u * 1024u;
    int sms = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    TcParams q = p;
    // row pairs (r, r + dil) share halo rows and every weight tile: used when the weights have to stream (5x5s)
    // or the map is tall enough that pairing leaves every SM busy
    const bool resident1 = 2 * (size_t)kh * p.halo_pitch * 128u + (size_t)iters * p.b_bytes + 1024 <= budget;
    q.mt = (g_halo_pairs && 2 * p.tmem_cols <= 512 && (!resident1 || grid >= 4ll * sms)) ? 2 : 1;
    // row quads when the weights have to stream anyway (5x5s): a quarter of the weight stream per output row, as long
    // as the quads still cover most of the SMs, the quad's halo leaves room for a >= 8-deep weight ring and four
    // accumulators fit in TMEM
    if (g_halo_pairs && g_halo_quads && q.mt == 2 && kh + 3 <= TCHP_MAX_ROWS && 4 * p.tmem_cols <= 512 &&
        2 * (size_t)(kh + 1) * p.halo_pitch * 128u + (size_t)iters * p.b_bytes + 1024 > budget &&
        (size_t)(kh + 3) * p.halo_pitch * 128u + 8 * (size_t)p.b_bytes + 1024 <= budget &&
        (long long)ceil_div(y->h, 4 * dil) * dil * p.tiles_x * y->n >= (long long)sms * 4 / 5)
      q.mt = 4;
    if (kh + q.mt - 1 > TCHP_MAX_ROWS) return ADD_ERR_UNSUPPORTED;
    const size_t hbytes = (size_t)(kh + q.mt - 1) * p.halo_pitch * 128u;
    const int upc = q.mt > 1 ? ceil_div(y->h, q.mt * dil) * dil : y->h;           // units per column strip
    q.tiles_y = upc;
    const long long units = (long long)upc * p.tiles_x * y->n;
    q.n_tiles = (int)units;
    size_t psmem = 0;
    bool ok = true;
    if (2 * hbytes + (size_t)iters * p.b_bytes + 1024 <= budget) {
      // every tap's weights resident + two halo buffers (next unit's halo streams in under this unit's MMAs)
      q.b_resident = 1; q.halo_bufs = 2;
      psmem = 2 * hbytes + (size_t)iters * p.b_bytes + 1024;
    } else {
      // weights stream through a ring: what limits it is bytes in flight (ring depth x tile size over the TMA round
      // trip), so the ring gets everything a single halo buffer leaves
      q.b_resident = 0;
      q.halo_bufs = (2 * hbytes + 12 * (size_t)p.b_bytes + 1024 <= budget) ? 2 : 1;
      long long sb = ((long long)budget - 1024 - (long long)q.halo_bufs * (long long)hbytes) / (long long)p.b_bytes;
      if (sb > TCHP_MAX_STAGES) sb = TCHP_MAX_STAGES;
      if (sb < 2) ok = false;
      q.stages = (int)sb;
      psmem = (size_t)q.halo_bufs * hbytes + (size_t)(sb > 0 ? sb : 0) * p.b_bytes + 1024;
    }
    if (ok && (q.b_resident || q.mt >= 2 || g_halo_stream_persistent)) {
      static PerDeviceOnce hponce;
      once_per_device(hponce, [] {
        cudaFuncSetAttribute(conv2d_tc_halo_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024);
        cudaFuncSetAttribute(conv2d_tc_halo_persistent_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      });
      long long g = sms < units ? sms : units;
      launch_kernel(conv2d_tc_halo_persistent_kernel, dim3((unsigned)g), dim3(TCHP_THREADS), psmem, static_cast<cudaStream_t>(stream), map_x, map_w, q);
      ADD_RETURN_LAUNCH();
    }
  }
  if (halo)
    launch_kernel(conv2d_tc_halo_kernel, dim3((unsigned)grid), dim3(TCH_THREADS), smem, static_cast<cudaStream_t>(stream), map_x, map_w, p);
  else
    launch_kernel(conv2d_tc_kernel, dim3((unsigned)grid), dim3(TC_THREADS), smem, static_cast<cudaStream_t>(stream), map_x, map_w, p);
  ADD_RETURN_LAUNCH();
}

/* Tuning / experiment switch for the halo-resident A path (see conv2d_tc_halo_kernel). */
extern "C" int add_conv2d_tc_set_halo_mode(int mode) {
  g_persistent = (mode & 16) ? 0 : 1;                 // bit 4 set = one tile per CTA (A/B runs)
  g_conv_cluster = (mode & 256) ? 0 : 1;              // bit 8 set = no clusters / multicast
  g_halo_pairs = (mode & 128) ? 0 : 1;                // bit 7 set = no row pairs in the persistent halo kernel
  g_halo_quads = (mode & 512) ? 0 : 1;                // bit 9 set = no row quads (pairs only)
  g_conv_mt2_small = (mode & 1024) ? 1 : 0;           // bit 10 set = two pixel tiles per stage for the resident 1x1s
  g_conv_flatten_1x1 = (mode & 2048) ? 0 : 1;         // bit 11 set = (row, column) tiles for the 1x1s too
  g_conv_mt2 = (mode & 64) ? 0 : 1;                   // bit 6 set = one pixel tile per weight tile everywhere
  g_halo_stream_persistent = (mode & 32) ? 1 : 0;     // bit 5 set = persistent halo kernel with streamed weights
  mode &= 15;
  if (mode < 0 || mode > 2) return ADD_ERR_BAD_ARG;
  g_halo_mode = mode;
  return ADD_OK;
}
