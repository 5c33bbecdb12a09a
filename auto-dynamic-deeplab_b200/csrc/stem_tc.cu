// stem_tc.cu — the network's first layer on the bf16 path: NCHW fp32 image -> 3x3 stride-2 conv (3 -> 64)
// -> folded BN -> ReLU -> NHWC bf16, in ONE kernel (reference ADD.py:154-158 `stem0`, fed by the loader's NCHW
// fp32 tensor, eval.py:175).  Replaces the add_nchw_to_nhwc + add_conv2d_tc_fwd pair (1.7 ms per 8 images):
// the layout change, the fp32->bf16 conversion and the im2col all happen on the way into shared memory.
//
// One CTA = 128 consecutive output pixels of one output row.
//  1. the 3 channels x 3 input rows x 257 columns it needs are read coalesced from the NCHW planes into smem
//     (fp32, zero outside the image = the conv padding);
//  2. each of 128 threads builds ITS pixel's im2col row (27 taps, k = (ci*3+ky)*3+kx, padded to 32) as bf16,
//     directly in the UMMA K-major SWIZZLE_128B layout;
//  3. one thread issues two tcgen05.mma (M=128, N=64, K=16) into a 64-column TMEM accumulator; the 64x32
//     weights (BN scale folded) arrive by TMA;
//  4. epilogue: tcgen05.ld -> +bias -> ReLU -> bf16 -> swizzled smem tile -> ONE TMA store of the 128 px x 64 ch
//     tile (full 128-byte lines; columns past the image edge are clipped by the TMA unit).
// HBM-bound: 12 B/input pixel read + 128 B/output pixel written (= 44 B per input pixel).
#include "tc_common.cuh"

namespace {

constexpr int ST_COUT = 64;
constexpr int ST_IN_PITCH = 260;                 // 257 columns used, padded
constexpr int ST_THREADS = 160;                  // 4 worker warps + 1 control warp

struct StemParams {
  const float* x; const float* bias;
  int N, H, W, Ho, Wo, tiles_x, n_tiles;
  uint32_t flags;
};

constexpr int ST_LD = 19;                        // ceil(9 * 257 / 128) global loads per worker thread per tile

// Persistent: a CTA loops over tiles (TMEM, barriers and the weights are set up once).  The NCHW loads of tile
// i+1 are issued into registers before tile i's im2col / MMA / epilogue, so their DRAM latency is off the chain.
__global__ void __launch_bounds__(ST_THREADS, 4)
stem_conv3x3s2_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_y, const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3];            // weights landed, A tile built (128 arrivals), MMA done
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float bias_s[ST_COUT];

  const uint32_t a_base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // 16 KB: A tile, later the output tile
  const uint32_t b_base = a_base + TC_A_BYTES;                       // 8 KB: 64 x 128 B weights
  float* in_s = reinterpret_cast<float*>(smem_raw + (b_base + ST_COUT * 128u - smem_u32(smem_raw)));   // [9][ST_IN_PITCH]
  const uint32_t bar_w = smem_u32(&bars[0]), bar_a = smem_u32(&bars[1]), bar_mma = smem_u32(&bars[2]);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_ctrl = warp == 4;
  if (tid < ST_COUT) bias_s[tid] = __ldg(p.bias + tid);

  if (is_ctrl) {
    if (lane == 0) {
      mbar_init(bar_w, 1);
      mbar_init(bar_a, 128);
      mbar_init(bar_mma, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_smem)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  pdl_launch_dependents();
  pdl_wait();
  const int tiles_per_img = p.tiles_x * p.Ho;

  if (is_ctrl) {
    if (elect_one()) {
      mbar_expect_tx(bar_w, ST_COUT * 128u);
      tma_load_3d(b_base, &map_w, bar_w, 0, 0, 0);
      mbar_wait(bar_w, 0);
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(ST_COUT >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const uint64_t adesc = make_kmajor_sw128_desc(a_base), bdesc = make_kmajor_sw128_desc(b_base);
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ph ^= 1u) {
        mbar_wait(bar_a, ph);                               // all 128 im2col rows written (and fenced)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        umma_bf16(tmem_base, adesc, bdesc, idesc, 0u);
        umma_bf16(tmem_base, adesc + 2, bdesc + 2, idesc, 1u);
        umma_commit(bar_mma);
      }
    }
  } else {
    const size_t plane = (size_t)p.H * p.W;
    const bool relu_out = p.flags & ADD_RELU_OUT;
    float pre[ST_LD];
    // element e = r * 257 + j of the 9 x 257 input patch of a tile: thread handles e = tid, tid + 128, ...
    auto load_tile = [&](int tile) {
      const int n = tile / tiles_per_img, rr = tile - n * tiles_per_img;
      const int oy = rr / p.tiles_x, tx = rr - oy * p.tiles_x;
      const int ix0 = 2 * tx * TC_BM - 1;
      const float* xn = p.x + (size_t)n * 3 * plane;
#pragma unroll
      for (int i = 0; i < ST_LD; ++i) {
        const int e = tid + i * 128;
        float v = 0.f;
        if (e < 9 * 257) {
          const int r = e / 257, j = e - r * 257;
          const int ci = r / 3, ky = r - 3 * ci;
          const int iy = 2 * oy - 1 + ky, ix = ix0 + j;
          if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) v = __ldg(xn + (size_t)ci * plane + (size_t)iy * p.W + ix);
        }
        pre[i] = v;
      }
    };
    if (blockIdx.x < p.n_tiles) load_tile(blockIdx.x);
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ph ^= 1u) {
      const int n = tile / tiles_per_img, rr = tile - n * tiles_per_img;
      const int oy = rr / p.tiles_x, tx = rr - oy * p.tiles_x;
      const int ox0 = tx * TC_BM;
      // ---- 1. prefetched input patch -> smem; start the next tile's loads ----
#pragma unroll
      for (int i = 0; i < ST_LD; ++i) {
        const int e = tid + i * 128;
        if (e < 9 * 257) { const int r = e / 257; in_s[r * ST_IN_PITCH + (e - r * 257)] = pre[i]; }
      }
      if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous tile's TMA store has read the A/out tile
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (tile + (int)gridDim.x < p.n_tiles) load_tile(tile + gridDim.x);
      // ---- 2. im2col row of pixel m = tid: k = (ci*3+ky)*3+kx, 27 taps + 5 zeros, bf16, SW128 ----
      {
        const int m = tid;
        float v[32];
#pragma unroll
        for (int r = 0; r < 9; ++r) {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) v[r * 3 + kx] = in_s[r * ST_IN_PITCH + 2 * m + kx];
        }
#pragma unroll
        for (int i = 27; i < 32; ++i) v[i] = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 b = __floats2bfloat162_rn(v[u * 8 + 2 * j], v[u * 8 + 2 * j + 1]);
            w[j] = *reinterpret_cast<uint32_t*>(&b);
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_base + m * 128 + ((u ^ (m & 7)) << 4)),
                       "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(bar_a);
      // ---- 4. epilogue: the A tile is dead once the MMAs completed -> reuse it as the output tile ----
      mbar_wait(bar_mma, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int m = tid;                                   // TMEM lane = pixel; warp w owns lanes 32w..32w+31
#pragma unroll
      for (int c0 = 0; c0 < ST_COUT; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + 4 * j);
          __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(v[4 * j]) + b4.x, __uint_as_float(v[4 * j + 1]) + b4.y);
          __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(v[4 * j + 2]) + b4.z, __uint_as_float(v[4 * j + 3]) + b4.w);
          w[2 * j] = *reinterpret_cast<uint32_t*>(&h0);
          w[2 * j + 1] = *reinterpret_cast<uint32_t*>(&h1);
          if (relu_out) {
            asm("max.bf16x2 %0, %0, %1;" : "+r"(w[2 * j]) : "r"(0u));
            asm("max.bf16x2 %0, %0, %1;" : "+r"(w[2 * j + 1]) : "r"(0u));
          }
        }
        const int u = c0 >> 3;                               // 16-byte unit of the 128-byte pixel row
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_base + m * 128 + ((u ^ (m & 7)) << 4)),
                     "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_base + m * 128 + (((u + 1) ^ (m & 7)) << 4)),
                     "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (tid == 0) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                     ::"l"(&map_y), "r"(a_base), "r"(0), "r"(ox0), "r"(n * p.Ho + oy) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (is_ctrl) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
  }
}

}  // namespace

/* Pack the stem weights: w_oihw fp32 [64][3][3][3] (BN scale folded) -> bf16 [64][64], row co = the 27 taps in
 * (ci, ky, kx) order followed by zeros (K padded to the 128-byte UMMA row). */
extern "C" int64_t add_stem_tc_packed_bytes(void) { return (int64_t)ST_COUT * 64 * 2; }

extern "C" int add_stem_tc_pack(const float* w_oihw, void* packed_host) {
  ADD_CHECK_ARG(w_oihw && packed_host);
  uint16_t* out = static_cast<uint16_t*>(packed_host);
  std::memset(out, 0, (size_t)ST_COUT * 64 * 2);
  for (int co = 0; co < ST_COUT; ++co)
    for (int k = 0; k < 27; ++k) {
      uint32_t u; std::memcpy(&u, &w_oihw[co * 27 + k], 4);
      uint32_t rnd = 0x7FFFu + ((u >> 16) & 1u);
      out[co * 64 + k] = (uint16_t)((u + rnd) >> 16);
    }
  return ADD_OK;
}

extern "C" int add_stem_conv3x3s2_nchw_fwd(const float* x_nchw, int n, int h, int w, const add_tensor_t* y,
                                           const void* w_packed, const float* bias, uint32_t flags, void* stream) {
  ADD_CHECK_ARG(x_nchw && tensor_ok(y) && w_packed && bias && n > 0 && h > 0 && w > 0);
  ADD_CHECK_ARG(y->n == n && y->h == (h - 1) / 2 + 1 && y->w == (w - 1) / 2 + 1);
  ADD_CHECK_SUP(y->dtype == ADD_BF16 && y->c == ST_COUT && y->pix_stride % 8 == 0 && ((uintptr_t)y->ptr % 16) == 0 &&
                ((uintptr_t)w_packed % 16) == 0);
  ADD_CHECK_SUP(!(flags & (ADD_RELU_IN | ADD_ACCUMULATE)));
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) { g_add_last_cuda_error = (int)cudaErrorSymbolNotFound; return ADD_ERR_CUDA; }
  StemParams p;
  p.x = x_nchw; p.bias = bias; p.N = n; p.H = h; p.W = w; p.Ho = y->h; p.Wo = y->w;
  p.tiles_x = ceil_div(y->w, TC_BM); p.flags = flags;
  const long long n_tiles = (long long)p.tiles_x * p.Ho * n;
  ADD_CHECK_SUP(n_tiles < (1ll << 31));
  p.n_tiles = (int)n_tiles;
  CUtensorMap map_w, map_y;
  {
    cuuint64_t dims[3] = {64, (cuuint64_t)ST_COUT, 1};
    cuuint64_t strides[2] = {128, (cuuint64_t)ST_COUT * 128};
    cuuint32_t box[3] = {64, (cuuint32_t)ST_COUT, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (encode(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w_packed), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ADD_ERR_UNSUPPORTED;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)ST_COUT, (cuuint64_t)y->w, (cuuint64_t)y->h * y->n};
    cuuint64_t strides[2] = {(cuuint64_t)y->pix_stride * 2, (cuuint64_t)y->w * y->pix_stride * 2};
    cuuint32_t box[3] = {(cuuint32_t)ST_COUT, (cuuint32_t)TC_BM, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (encode(&map_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, y->ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ADD_ERR_UNSUPPORTED;
  }
  const size_t smem = TC_A_BYTES + ST_COUT * 128 + 9 * ST_IN_PITCH * sizeof(float) + 1024;
  static PerDeviceOnce once;
  once_per_device(once, [] {
    cudaFuncSetAttribute(stem_conv3x3s2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(stem_conv3x3s2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  });
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  long long grid = (long long)sms * 4;                 // 4 resident CTAs per SM (registers: 160 threads x ~100)
  if (grid > n_tiles) grid = n_tiles;
  launch_kernel(stem_conv3x3s2_kernel, dim3((unsigned)grid), dim3(ST_THREADS), smem, static_cast<cudaStream_t>(stream), map_w, map_y, p);
  ADD_RETURN_LAUNCH();
}
