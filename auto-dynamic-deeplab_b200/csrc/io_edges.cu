// io_edges.cu — the loader / dump edges either side of the network, all HBM-bound byte work:
//   * Cityscapes label ids -> train ids (dataloaders/datasets/cityscapes.py:85-91) as a 256-entry LUT, fused with the
//     pad-to-crop-size of the evaluation transform (custom_transforms.py:322-347: labels padded with 255);
//   * uint8 HWC image -> normalised fp32 NCHW, padded with zeros AFTER normalisation (same transform: ZeroPad2d follows
//     torchvision's ToTensor + Normalize: float32 arithmetic, bit-identical);
//   * class map -> colour image (dataloaders/utils.py:14-51 decode_segmap) as a 256-entry RGB LUT.
// One thread handles 16 output bytes (labels) / 4 output pixels (images); rows are independent, grids are sized in
// multiples of the SM count.
#include "common.cuh"

namespace {

// dst[n][Hp][Wp] = lut[src[n][h][w]] inside the image, `fill` in the padding.  lut == nullptr: identity.
__global__ void __launch_bounds__(256)
encode_pad_labels_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int h, int w, int Hp, int Wp,
                         const uint8_t* __restrict__ lut, int fill) {
  __shared__ uint8_t lut_s[256];
  lut_s[threadIdx.x] = lut ? lut[threadIdx.x] : (uint8_t)threadIdx.x;
  __syncthreads();
  const int n = blockIdx.y;
  const long long per = (long long)Hp * Wp;
  const uint8_t* in = src + (size_t)n * h * w;
  uint8_t* out = dst + (size_t)n * per;
  for (long long p = (blockIdx.x * 256ll + threadIdx.x) * 4; p < per; p += (long long)gridDim.x * 256 * 4) {
    uint8_t v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long q = p + j;
      const int y = (int)(q / Wp), x = (int)(q - (long long)y * Wp);
      v[j] = (q < per && y < h && x < w) ? lut_s[in[(size_t)y * w + x]] : (uint8_t)fill;
    }
    if (p + 3 < per && ((uintptr_t)(out + p) & 3) == 0) {
      *reinterpret_cast<uchar4*>(out + p) = make_uchar4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (p + j < per) out[p + j] = v[j];
    }
  }
}

__global__ void __launch_bounds__(256)
normalize_pad_u8_hwc_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int h, int w, int Hp, int Wp,
                            double m0, double m1, double m2, double s0, double s1, double s2) {
  const int n = blockIdx.y;
  const long long per = (long long)Hp * Wp;
  const uint8_t* in = src + (size_t)n * h * w * 3;
  float* out = dst + (size_t)n * per * 3;
  for (long long p = blockIdx.x * 256ll + threadIdx.x; p < per; p += (long long)gridDim.x * 256) {
    const int y = (int)(p / Wp), x = (int)(p - (long long)y * Wp);
    float o0 = 0.f, o1 = 0.f, o2 = 0.f;                       // ZeroPad2d after Normalize: the padding is exactly 0
    if (y < h && x < w) {
      const uint8_t* px = in + ((size_t)y * w + x) * 3;
      // full_image_eval_preprocess uses torchvision's transforms (custom_transforms.py:331-336): ToTensor = /255 in
      // float32, Normalize = tensor.sub_(mean).div_(std) with mean / std converted to float32 tensors — all float32,
      // each operation rounded to nearest (unlike the numpy Normalize class of the training transforms, which goes
      // through float64: that one is add_normalize_u8_hwc_to_nchw)
      const float r = __fdiv_rn((float)px[0], 255.0f), g = __fdiv_rn((float)px[1], 255.0f), b = __fdiv_rn((float)px[2], 255.0f);
      o0 = __fdiv_rn(__fsub_rn(r, (float)m0), (float)s0);
      o1 = __fdiv_rn(__fsub_rn(g, (float)m1), (float)s1);
      o2 = __fdiv_rn(__fsub_rn(b, (float)m2), (float)s2);
    }
    out[p] = o0; out[per + p] = o1; out[2 * per + p] = o2;
  }
}

// rgb[n][H][W][3] (uint8) = lut[label] ; labels are int64 (argmax output) or uint8
template <typename L>
__global__ void __launch_bounds__(256)
decode_segmap_kernel(const L* __restrict__ label, uint8_t* __restrict__ rgb, long long n_pix, const uint8_t* __restrict__ lut) {
  __shared__ uint8_t lut_s[768];
  for (int i = threadIdx.x; i < 768; i += 256) lut_s[i] = lut[i];
  __syncthreads();
  for (long long p = blockIdx.x * 256ll + threadIdx.x; p < n_pix; p += (long long)gridDim.x * 256) {
    const long long v = (long long)label[p];
    const int k = (int)(v < 0 ? 255 : (v > 255 ? 255 : v));
    rgb[3 * p] = lut_s[3 * k]; rgb[3 * p + 1] = lut_s[3 * k + 1]; rgb[3 * p + 2] = lut_s[3 * k + 2];
  }
}

inline unsigned io_blocks(long long work_items, int n) {
  long long b = (work_items + 255) / 256;
  const long long cap = (148ll * 16 + n - 1) / n;
  if (b > cap) b = cap;
  return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" int add_encode_pad_labels_u8(const uint8_t* src, uint8_t* dst, int n, int h, int w, int Hp, int Wp,
                                        const uint8_t* lut256_dev, int fill, void* stream) {
  ADD_CHECK_ARG(src && dst && n > 0 && h > 0 && w > 0 && Hp >= h && Wp >= w && fill >= 0 && fill <= 255);
  ADD_CHECK_SUP(n < 65536);
  const long long per = (long long)Hp * Wp;
  encode_pad_labels_kernel<<<dim3(io_blocks((per + 3) / 4, n), (unsigned)n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, dst, h, w, Hp, Wp, lut256_dev, fill);
  ADD_RETURN_LAUNCH();
}

extern "C" int add_normalize_pad_u8_hwc_to_nchw(const uint8_t* src, float* dst, int n, int h, int w, int Hp, int Wp,
                                                double mean0, double mean1, double mean2, double std0, double std1,
                                                double std2, void* stream) {
  ADD_CHECK_ARG(src && dst && n > 0 && h > 0 && w > 0 && Hp >= h && Wp >= w && std0 != 0.0 && std1 != 0.0 && std2 != 0.0);
  ADD_CHECK_SUP(n < 65536);
  const long long per = (long long)Hp * Wp;
  normalize_pad_u8_hwc_kernel<<<dim3(io_blocks(per, n), (unsigned)n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, dst, h, w, Hp, Wp, mean0, mean1, mean2, std0, std1, std2);
  ADD_RETURN_LAUNCH();
}

extern "C" int add_decode_segmap(const void* labels, int labels_are_int64, uint8_t* rgb, int64_t n_pixels,
                                 const uint8_t* lut768_dev, void* stream) {
  ADD_CHECK_ARG(labels && rgb && lut768_dev && n_pixels >= 0);
  if (n_pixels == 0) return ADD_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned blocks = io_blocks(n_pixels, 1);
  if (labels_are_int64) decode_segmap_kernel<long long><<<blocks, 256, 0, s>>>((const long long*)labels, rgb, n_pixels, lut768_dev);
  else decode_segmap_kernel<uint8_t><<<blocks, 256, 0, s>>>((const uint8_t*)labels, rgb, n_pixels, lut768_dev);
  ADD_RETURN_LAUNCH();
}
