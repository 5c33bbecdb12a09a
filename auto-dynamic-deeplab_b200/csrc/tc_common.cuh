// tc_common.cuh — shared pieces of the tcgen05/TMEM kernels (conv_tc.cu, sepconv_tc.cu): PTX wrappers
// (mbarrier, TMA, tcgen05.mma/commit/ld), UMMA descriptors, the fused epilogue, tensor-map encoding.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <cstring>
#include <mutex>
#include <utility>

namespace {


constexpr int TC_BM = 128;                 // pixels per CTA tile (UMMA M)
constexpr int TC_BK = 64;                  // channels per K chunk (128 B of bf16 = one swizzle row)
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;
constexpr int TC_THREADS = 192;
constexpr int TC_MAX_STAGES = 6;

struct TcParams {
  void* y; const float* bias;
  int Ho, Wo, Cout, ys, y_is_f32;
  int tiles_x, tiles_y;                    // spatial tiles per image
  int bw_log2;                             // BW = 1 << bw_log2, BH = 128 >> bw_log2
  int taps_w, taps, kchunks, Cin;
  int stride, pad, dil;
  int n_pad;                               // UMMA N
  int tmem_cols;
  int stages;
  uint32_t b_bytes;                        // n_pad * 128
  uint32_t flags;
  // halo-resident mode (stride 1, 128-pixel row tiles)
  int halo_pitch;                          // pixels per halo row in smem (multiple of 8)
  int halo_bufs;                           // 1 or 2 halo buffers
  int base_off_mode;                       // descriptor base_offset: 0 = always 0, 1 = (addr >> 7) & 7
  // persistent kernel
  long long bias_img_stride;               // floats between consecutive images' bias vectors (0 = one shared vector)
  int n_tiles;                             // tiles_x * tiles_y * N
  int b_resident;                          // 1: all weight chunks stay in shared memory for the whole kernel
  int mt;                                  // halo kernel: output rows per unit (1, or 2 = rows r and r+dil sharing halo rows and weights)
};

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// Kernels launched through launch_kernel() with PDL on may START before the previous kernel of their stream has
// finished: everything up to pdl_wait() (barrier init, TMEM alloc, bias staging, tensor-map prefetch — nothing
// that reads activations) overlaps the predecessor's tail; pdl_wait() returns once the predecessor grid has
// completed and its writes are visible.  pdl_launch_dependents() lets OUR successor start its own prologue.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }


template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster,
                                         Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g_add_pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g_add_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// One thread of a converged warp, chosen with elect.sync: unlike `lane == 0` the compiler then KNOWS the region is
// single-threaded and emits each tcgen05.mma / TMA once instead of wrapping it in an elect-and-loop-over-active-
// threads sequence (4 extra instructions and a branch per MMA in the issuing thread's critical loop).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// ---- thread-block cluster helpers (TMA multicast of a tile that every CTA of the cluster needs) -----------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one 3-D box fetched from L2 ONCE and written to the same shared-memory offset of every CTA in `mask`; each
// destination CTA's mbarrier (same offset) receives the complete_tx for the bytes that landed in ITS memory
__device__ __forceinline__ void tma_load_3d_multicast(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                                      uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask) : "memory");
}
// arrive on the mbarrier at this offset in every CTA of `mask` once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_multicast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format): 128-byte rows, 8-row groups
// 1024 B apart (SBO), LBO unused for swizzled K-major layouts.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address
  d |= (uint64_t)1 << 16;                          // leading byte offset (ignored) = 1
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}

// ---- shared pieces of the two kernels -----------------------------------------------------------
// In-place ReLU over `bytes` of bf16 in shared memory by the 128 ReLU/epilogue threads (et = 0..127).
__device__ __forceinline__ void relu_sweep(uint32_t base, uint32_t bytes, int et) {
  // batches of 4 x 16 B per thread: all loads of a batch are issued before the first max/store, so one batch costs
  // one shared-memory round trip instead of four back-to-back ones (the sweep sits on the MMA's critical path)
  constexpr uint32_t STEP = 128 * 16;
  uint32_t off = et * 16;
  for (; off + 3 * STEP < bytes; off += 4 * STEP) {
    uint32_t v[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[i][0]), "=r"(v[i][1]), "=r"(v[i][2]), "=r"(v[i][3])
                   : "r"(base + off + i * STEP));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) asm("max.bf16x2 %0, %0, %1;" : "+r"(v[i][j]) : "r"(0u));
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + off + i * STEP), "r"(v[i][0]), "r"(v[i][1]), "r"(v[i][2]), "r"(v[i][3]) : "memory");
    }
  }
  for (; off < bytes; off += STEP) {
    const uint32_t addr = base + off;
    uint32_t v0, v1, v2, v3;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(addr));
    asm("max.bf16x2 %0, %0, %1;" : "+r"(v0) : "r"(0u));
    asm("max.bf16x2 %0, %0, %1;" : "+r"(v1) : "r"(0u));
    asm("max.bf16x2 %0, %0, %1;" : "+r"(v2) : "r"(0u));
    asm("max.bf16x2 %0, %0, %1;" : "+r"(v3) : "r"(0u));
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to UMMA
}

// Folded-BN bias staged once per CTA in shared memory (zero past Cout up to n_pad), so the epilogue reads it with
// broadcast 16-byte loads instead of one predicated global load per output element (ncu r01g: that single line
// was 16 % of all warp instructions of the SepConv kernel — more than its FFMA2s).
constexpr int TC_MAX_NPAD = 256;
__device__ __forceinline__ void stage_bias(float* bias_s, const TcParams& p, int tid, int nthreads, int n = 0) {
  const float* b = p.bias ? p.bias + (long long)n * p.bias_img_stride : nullptr;      // per-image bias (ASPP pool branch)
  for (int i = tid; i < p.n_pad; i += nthreads) bias_s[i] = (b && i < p.Cout) ? __ldg(b + i) : 0.f;
}

// spin on an mbarrier with a short sleep between polls: for waiters with slack (producers, epilogue warps), so
// their polling does not take issue slots from the compute warps of the same SM sub-partition
#ifndef ADD_SLEEP_NS
#define ADD_SLEEP_NS 128
#endif
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(ADD_SLEEP_NS);
  }
}

// ---- epilogue ------------------------------------------------------------------------------------
// One 16-channel chunk of one pixel, bf16 output: (+= old y) -> round -> ReLU -> 16-byte stores.  `have_old`: the
// caller already holds the old y words of the two 8-channel groups (prefetched), else they are loaded here.
__device__ __forceinline__ void epi_chunk_bf16(const TcParams& p, float (&f)[16], bf16* dst, int c0, bool accum, bool relu_out,
                                               bool have_old, const uint4& o0, const uint4& o1) {
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    if (c0 + 8 * g + 8 <= p.Cout) {
      if (accum) {
        const uint4 old = have_old ? (g == 0 ? o0 : o1) : *reinterpret_cast<const uint4*>(dst + 8 * g);
        const uint32_t ow[4] = {old.x, old.y, old.z, old.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f[8 * g + 2 * j] += __uint_as_float(ow[j] << 16);
          f[8 * g + 2 * j + 1] += __uint_as_float(ow[j] & 0xffff0000u);
        }
      }
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 h = __floats2bfloat162_rn(f[8 * g + 2 * j], f[8 * g + 2 * j + 1]);
        pk[j] = *reinterpret_cast<uint32_t*>(&h);
        // ReLU on the rounded pair: rounding is monotonic and keeps the sign, so this equals round(relu(x))
        if (relu_out) asm("max.bf16x2 %0, %0, %1;" : "+r"(pk[j]) : "r"(0u));
      }
      *reinterpret_cast<uint4*>(dst + 8 * g) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    } else if (c0 + 8 * g < p.Cout) {
      for (int j = 8 * g; j < 8 * g + 8; ++j)
        if (c0 + j < p.Cout) {
          float o = f[j];
          if (accum) o += __bfloat162float(dst[j]);
          if (relu_out) o = fmaxf(o, 0.f);
          dst[j] = __float2bfloat16_rn(o);
        }
    }
  }
}

// Accumulate-into-slice (`ADD_ACCUMULATE`, the cell's node sum) needs the old y values.  Loaded inside the chunk
// loop they cost one exposed global-memory round trip PER 16-channel chunk (launch table r02d: a SepConv half with
// += took 29 us against 15 us without).  For the narrow outputs of the cell ops (Cout <= 80, bf16) all of a pixel's
// old words are instead fetched up front in one batch — before the wait on the accumulator barrier where the caller
// can (epilogue_prefetch_old), else at the top of epilogue_store — so the round trip is paid once and overlaps the MMA.
constexpr int EPI_PRE_MAX = 10;   // 8-channel (16-byte) groups held in registers: Cout <= 80
__device__ __forceinline__ bool epilogue_prefetchable(const TcParams& p) {
  return (p.flags & ADD_ACCUMULATE) && !p.y_is_f32 && p.Cout <= 8 * EPI_PRE_MAX && !(p.flags & 0x100u);
}
__device__ __forceinline__ void epilogue_prefetch_old(const TcParams& p, int warp, int lane, int n, int y0, int x0,
                                                      uint4 (&old)[EPI_PRE_MAX]) {
  const int BW = 1 << p.bw_log2;
  const int r = (warp & 3) * 32 + lane;
  const int oy = y0 + (r >> p.bw_log2), ox = x0 + (r & (BW - 1));
  const bool valid = (oy < p.Ho) && (ox < p.Wo);
  const size_t pix = ((size_t)n * p.Ho + oy) * p.Wo + ox;
  const bf16* src = static_cast<const bf16*>(p.y) + pix * p.ys;
#pragma unroll
  for (int g = 0; g < EPI_PRE_MAX; ++g) {
    if (valid && 8 * g + 8 <= p.Cout) old[g] = *reinterpret_cast<const uint4*>(src + 8 * g);
    else old[g] = make_uint4(0u, 0u, 0u, 0u);
  }
}

// TMEM accumulator -> +bias -> (+= y) -> ReLU -> bf16/fp32 stores.  Called by the four epilogue warps
// after the accumulator-complete barrier.  bias_s: shared-memory bias (stage_bias).  `pre`: old y words from
// epilogue_prefetch_old for this very tile, or nullptr.
__device__ __forceinline__ void epilogue_store(const TcParams& p, uint32_t tmem_base, const float* bias_s, int warp, int lane,
                                               int n, int y0, int x0, const uint4 (*pre)[EPI_PRE_MAX] = nullptr) {
  const int BW = 1 << p.bw_log2;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int q = warp & 3;                 // TMEM lane quadrant this warp may access
  const int r = q * 32 + lane;            // tile row = pixel
  const int oy = y0 + (r >> p.bw_log2), ox = x0 + (r & (BW - 1));
  const bool valid = (oy < p.Ho) && (ox < p.Wo);
  const size_t pix = ((size_t)n * p.Ho + oy) * p.Wo + ox;
  const bool relu_out = p.flags & ADD_RELU_OUT, accum = p.flags & ADD_ACCUMULATE;
  if (epilogue_prefetchable(p)) {
    // narrow accumulate path: old words in registers, chunk loop unrolled so they are indexed statically
    uint4 old_local[EPI_PRE_MAX];
    if (!pre) epilogue_prefetch_old(p, warp, lane, n, y0, x0, old_local);
    const uint4* old = pre ? &(*pre)[0] : &old_local[0];
    bf16* base = static_cast<bf16*>(p.y) + pix * p.ys;
#pragma unroll
    for (int ci = 0; ci < EPI_PRE_MAX / 2; ++ci) {
      const int c0 = 16 * ci;
      if (c0 < p.n_pad) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (valid) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + 4 * j);
            f[4 * j] = __uint_as_float(v[4 * j]) + b4.x; f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
            f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z; f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
          }
          epi_chunk_bf16(p, f, base + c0, c0, true, relu_out, true, old[2 * ci], old[2 * ci + 1]);
        }
      }
    }
    return;
  }
  for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (!valid || (p.flags & 0x100u)) continue;      // 0x100: experiment — skip the epilogue math and stores
    float f[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + 4 * j);
      f[4 * j] = __uint_as_float(v[4 * j]) + b4.x; f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
      f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z; f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
    }
    if (p.y_is_f32) {
      float* dst = static_cast<float*>(p.y) + pix * p.ys + c0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (c0 + 4 * g + 4 <= p.Cout) {
          float4 o = make_float4(f[4 * g], f[4 * g + 1], f[4 * g + 2], f[4 * g + 3]);
          if (accum) { float4 old = *reinterpret_cast<const float4*>(dst + 4 * g); o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
          if (relu_out) o = relu4(o);
          *reinterpret_cast<float4*>(dst + 4 * g) = o;
        } else if (c0 + 4 * g < p.Cout) {
          for (int j = 4 * g; j < 4 * g + 4; ++j)
            if (c0 + j < p.Cout) {
              float o = f[j];
              if (accum) o += dst[j];
              if (relu_out) o = fmaxf(o, 0.f);
              dst[j] = o;
            }
        }
      }
    } else {
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      epi_chunk_bf16(p, f, static_cast<bf16*>(p.y) + pix * p.ys + c0, c0, accum, relu_out, false, z, z);
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}


inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
inline int kchunks_of(int cin) { return (cin + TC_BK - 1) / TC_BK; }
inline int npad_of(int cout) { return round_up(cout, 16); }

}  // namespace
