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
  int N, H, W, Ho, Wo, tiles_x;
  uint32_t flags;
};

__global__ void __launch_bounds__(ST_THREADS)
stem_conv3x3s2_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_y, const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_base_smem;

  const uint32_t a_base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // 16 KB: A tile, later the output tile
  const uint32_t b_base = a_base + TC_A_BYTES;                       // 8 KB: 64 x 128 B weights
  float* in_s = reinterpret_cast<float*>(smem_raw + (b_base + ST_COUT * 128u - smem_u32(smem_raw)));   // [9][ST_IN_PITCH]
  const uint32_t bar_w = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_ctrl = warp == 4;
  int t = blockIdx.x;
  const int tx = t % p.tiles_x; t /= p.tiles_x;
  const int oy = t % p.Ho; const int n = t / p.Ho;
  const int ox0 = tx * TC_BM;

  if (is_ctrl) {
    if (lane == 0) {
      mbar_init(bar_w, 1);
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

  if (is_ctrl) {
    if (lane == 0) {
      mbar_expect_tx(bar_w, ST_COUT * 128u);
      tma_load_3d(b_base, &map_w, bar_w, 0, 0, 0);
    }
  } else {
    // ---- 1. input rows: 9 (channel, ky) rows x 257 columns, coalesced along W ----
    const int ix0 = 2 * ox0 - 1;
    const size_t plane = (size_t)p.H * p.W;
    const float* xn = p.x + (size_t)n * 3 * plane;
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int ci = r / 3, ky = r - 3 * ci;
      const int iy = 2 * oy - 1 + ky;
      const bool row_ok = iy >= 0 && iy < p.H;
      const float* src = xn + (size_t)ci * plane + (size_t)(row_ok ? iy : 0) * p.W;
      for (int j = tid; j < 257; j += 128) {
        const int ix = ix0 + j;
        in_s[r * ST_IN_PITCH + j] = (row_ok && ix >= 0 && ix < p.W) ? __ldg(src + ix) : 0.f;
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
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
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (is_ctrl) {
    if (lane == 0) {
      mbar_wait(bar_w, 0);
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(ST_COUT >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const uint64_t adesc = make_kmajor_sw128_desc(a_base), bdesc = make_kmajor_sw128_desc(b_base);
      umma_bf16(tmem_base, adesc, bdesc, idesc, 0u);
      umma_bf16(tmem_base, adesc + 2, bdesc + 2, idesc, 1u);
      umma_commit(bar_mma);
    }
  } else {
    // ---- 4. epilogue: the A tile is dead once the MMAs completed -> reuse it as the output tile ----
    mbar_wait(bar_mma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int m = tid;                                   // TMEM lane = pixel; warp w owns lanes 32w..32w+31
    const bool relu_out = p.flags & ADD_RELU_OUT;
#pragma unroll
    for (int c0 = 0; c0 < ST_COUT; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a = __uint_as_float(v[2 * j]) + __ldg(p.bias + c0 + 2 * j);
        float b = __uint_as_float(v[2 * j + 1]) + __ldg(p.bias + c0 + 2 * j + 1);
        if (relu_out) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        w[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      const int u = c0 >> 3;                               // 16-byte unit of the 128-byte pixel row
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_base + m * 128 + ((u ^ (m & 7)) << 4)),
                   "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_base + m * 128 + (((u + 1) ^ (m & 7)) << 4)),
                   "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (tid == 0) {
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                   ::"l"(&map_y), "r"(a_base), "r"(0), "r"(ox0), "r"(n * p.Ho + oy) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
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
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(stem_conv3x3s2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(stem_conv3x3s2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  });
  const long long grid = (long long)p.tiles_x * p.Ho * n;
  ADD_CHECK_SUP(grid < (1ll << 31));
  stem_conv3x3s2_kernel<<<(unsigned)grid, ST_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(map_w, map_y, p);
  ADD_RETURN_LAUNCH();
}
