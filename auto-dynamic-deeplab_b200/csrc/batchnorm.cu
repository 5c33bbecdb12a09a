// batchnorm.cu — training-mode (batch-statistics) BatchNorm forward for NHWC views, split the way the
// reference's SynchronizedBatchNorm splits it (modeling/sync_batchnorm/batchnorm.py:48-78, 113-125):
//   1. add_bn_stats_fwd     per-channel [sum(x), sum(x^2)] of this rank's shard   (HBM-bound: reads x once)
//   2. (host)               ONE all-reduce of the packed [sum | ssum | count] vector over the ranks (NCCL / gloo)
//   3. add_bn_finalize      mean, inv_std and the running-statistics update from the reduced sums (C threads)
//   4. add_bn_apply_fwd     y = (x - mean) * (inv_std * weight) + bias  [-> ReLU]  (HBM-bound: reads x, writes y)
// SURVEY §8f row 1 (training forward); the eval-mode BN of the inference path never runs here — it is folded into
// the conv weights.  All reductions are deterministic (fixed-order two-stage sums, no atomics).
#include "common.cuh"

namespace {

constexpr int BN_THREADS = 256;

__host__ __device__ inline int bn_splits(long long P) {
  long long s = (P + 1023) / 1024;               // >= 1024 pixels per block
  const long long cap = 148 * 8;
  return (int)(s < 1 ? 1 : (s > cap ? cap : s));
}

template <typename T> struct Vec4IO;
template <> struct Vec4IO<float> {
  static __device__ __forceinline__ float4 load(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
};
template <> struct Vec4IO<bf16> {
  static __device__ __forceinline__ float4 load(const bf16* p) { return ld4(p); }
};

// stage 1: grid = splits; a block sums a contiguous range of the N*H*W pixels for ALL channels.  thread = (4-channel
// vector v, pixel lane l); lanes meet in shared memory in fixed order.  part[split][2][C].
template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
bn_stats_partial_kernel(const T* __restrict__ x, float* __restrict__ part, long long P, int C, int xs) {
  extern __shared__ float red[];                  // [lanes][2][C]
  const int cv = C / 4;
  const int lanes = BN_THREADS / cv;
  const int S = gridDim.x;
  const long long per = (P + S - 1) / S;
  const long long p0 = (long long)blockIdx.x * per, p1 = p0 + per < P ? p0 + per : P;
  const int v = threadIdx.x % cv, l = threadIdx.x / cv;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  if (l < lanes) {
    const T* xv = x + v * 4;
    long long p = p0 + l;
    for (; p + 3 * lanes < p1; p += 4 * lanes) {          // four independent 4-channel loads in flight
      float4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] = Vec4IO<T>::load(xv + (size_t)(p + u * lanes) * xs);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s.x += t[u].x; s.y += t[u].y; s.z += t[u].z; s.w += t[u].w;
        q.x = fmaf(t[u].x, t[u].x, q.x); q.y = fmaf(t[u].y, t[u].y, q.y);
        q.z = fmaf(t[u].z, t[u].z, q.z); q.w = fmaf(t[u].w, t[u].w, q.w);
      }
    }
    for (; p < p1; p += lanes) {
      const float4 t = Vec4IO<T>::load(xv + (size_t)p * xs);
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      q.x = fmaf(t.x, t.x, q.x); q.y = fmaf(t.y, t.y, q.y); q.z = fmaf(t.z, t.z, q.z); q.w = fmaf(t.w, t.w, q.w);
    }
    float* r = red + (size_t)l * 2 * C + v * 4;
    r[0] = s.x; r[1] = s.y; r[2] = s.z; r[3] = s.w;
    r[C] = q.x; r[C + 1] = q.y; r[C + 2] = q.z; r[C + 3] = q.w;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += BN_THREADS) {
    float tot = red[i];
    for (int r = 1; r < lanes; ++r) tot += red[(size_t)r * 2 * C + i];
    part[(size_t)blockIdx.x * 2 * C + i] = tot;
  }
}

// stage 2: sums[2C] = fixed-order sum of the per-block partials, accumulated in double (the reference sums fp32
// tensors with ATen's pairwise reduction; a plain fp32 running sum over ~1000 partials would be the less accurate one)
__global__ void bn_stats_finalize_kernel(const float* __restrict__ part, float* __restrict__ sums, int S, int C2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C2) return;
  double t = 0.0;
  for (int s = 0; s < S; ++s) t += (double)part[(size_t)s * C2 + i];
  sums[i] = (float)t;
}

// batchnorm.py:113-125 (sync = 1: mean = sum / n, sumvar = ssum - sum * mean, inv_std = clamp(sumvar / n, eps)^-1/2)
// or F.batch_norm's training formula (sync = 0: inv_std = (sumvar / n + eps)^-1/2, batchnorm.py:50-53 path); both
// update the running statistics with the UNBIASED variance and `momentum`.
__global__ void bn_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ count_ptr, float count_val, int C,
                                   float eps, float momentum, int sync, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ mean_out, float* __restrict__ inv_std_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float n = count_ptr ? *count_ptr : count_val;
  const float sum = sums[c], ssum = sums[C + c];
  const float mean = sum / n;
  const float sumvar = ssum - sum * mean;
  const float unbias_var = sumvar / (n - 1.f);
  const float bias_var = sumvar / n;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
  if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbias_var;
  mean_out[c] = mean;
  inv_std_out[c] = sync ? 1.f / sqrtf(fmaxf(bias_var, eps)) : 1.f / sqrtf(bias_var + eps);
}

// y = (x - mean) * (inv_std * weight) + bias — the reference's own association (batchnorm.py:71)
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long P, int C, int xs, int ys, const float* __restrict__ mean,
                const float* __restrict__ inv_std, const float* __restrict__ weight, const float* __restrict__ bias,
                uint32_t flags) {
  const int cv = C / 4;
  const long long total = P * cv;
  const bool relu_out = flags & ADD_RELU_OUT;
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    const long long p = idx / cv;
    const int c = (int)(idx - p * cv) * 4;
    const float4 t = ld4(x + (size_t)p * xs + c);
    const float4 m = __ldg(reinterpret_cast<const float4*>(mean + c));
    float4 g = __ldg(reinterpret_cast<const float4*>(inv_std + c));
    if (weight) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(weight + c));
      g.x *= w.x; g.y *= w.y; g.z *= w.z; g.w *= w.w;
    }
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) b = __ldg(reinterpret_cast<const float4*>(bias + c));
    float4 o = make_float4((t.x - m.x) * g.x + b.x, (t.y - m.y) * g.y + b.y, (t.z - m.z) * g.z + b.z, (t.w - m.w) * g.w + b.w);
    if (relu_out) o = relu4(o);
    st4(y + (size_t)p * ys + c, o);
  }
}

}  // namespace

extern "C" int64_t add_bn_stats_workspace_bytes(int n, int h, int w, int c) {
  if (n <= 0 || h <= 0 || w <= 0 || c <= 0) return ADD_ERR_BAD_ARG;
  return (int64_t)bn_splits((long long)n * h * w) * 2 * c * sizeof(float);
}

extern "C" int add_bn_stats_fwd(const add_tensor_t* x, float* sums, void* workspace, int64_t workspace_bytes, void* stream) {
  ADD_CHECK_ARG(tensor_ok(x) && sums && workspace);
  ADD_CHECK_SUP(tensor_vec4_ok(x) && x->c / 4 <= BN_THREADS && ((uintptr_t)sums % 16) == 0);
  if (workspace_bytes < add_bn_stats_workspace_bytes(x->n, x->h, x->w, x->c)) return ADD_ERR_WORKSPACE;
  const long long P = (long long)x->n * x->h * x->w;
  const int S = bn_splits(P), cv = x->c / 4;
  const size_t smem = (size_t)(BN_THREADS / cv) * 2 * x->c * sizeof(float);
  ADD_CHECK_SUP(smem <= 48 * 1024);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // a batch-contiguous NHWC view: images are h*w*pix_stride elements apart, so the whole shard is one pixel range
  if (x->dtype == ADD_F32)
    bn_stats_partial_kernel<float><<<S, BN_THREADS, smem, s>>>((const float*)x->ptr, (float*)workspace, P, x->c, x->pix_stride);
  else
    bn_stats_partial_kernel<bf16><<<S, BN_THREADS, smem, s>>>((const bf16*)x->ptr, (float*)workspace, P, x->c, x->pix_stride);
  bn_stats_finalize_kernel<<<ceil_div(2 * x->c, 128), 128, 0, s>>>((const float*)workspace, sums, S, 2 * x->c);
  ADD_RETURN_LAUNCH();
}

extern "C" int add_bn_finalize(const float* sums, const float* count_dev, float count, int c, float eps, float momentum,
                               int sync, float* running_mean, float* running_var, float* mean, float* inv_std, void* stream) {
  ADD_CHECK_ARG(sums && mean && inv_std && c > 0 && eps >= 0.f);
  ADD_CHECK_ARG(count_dev || count > 1.f);      /* "BatchNorm computes unbiased standard-deviation, which requires size > 1" */
  bn_finalize_kernel<<<ceil_div(c, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(sums, count_dev, count, c, eps, momentum,
                                                                                     sync, running_mean, running_var, mean, inv_std);
  ADD_RETURN_LAUNCH();
}

extern "C" int add_bn_apply_fwd(const add_tensor_t* x, const add_tensor_t* y, const float* mean, const float* inv_std,
                                const float* weight, const float* bias, uint32_t flags, void* stream) {
  ADD_CHECK_ARG(tensor_ok(x) && tensor_ok(y) && mean && inv_std);
  ADD_CHECK_ARG(x->n == y->n && x->h == y->h && x->w == y->w && x->c == y->c);
  ADD_CHECK_SUP(tensor_vec4_ok(x) && tensor_vec4_ok(y));
  ADD_CHECK_SUP(((uintptr_t)mean % 16) == 0 && ((uintptr_t)inv_std % 16) == 0 && ((uintptr_t)weight % 16) == 0 &&
                ((uintptr_t)bias % 16) == 0);
  const long long P = (long long)x->n * x->h * x->w, total = P * (x->c / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148ll * 32) blocks = 148ll * 32;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define BA(TI, TO) bn_apply_kernel<TI, TO><<<(unsigned)blocks, 256, 0, s>>>((const TI*)x->ptr, (TO*)y->ptr, P, x->c, x->pix_stride, \
    y->pix_stride, mean, inv_std, weight, bias, flags)
  if (x->dtype == ADD_F32 && y->dtype == ADD_F32) BA(float, float);
  else if (x->dtype == ADD_BF16 && y->dtype == ADD_BF16) BA(bf16, bf16);
  else if (x->dtype == ADD_F32 && y->dtype == ADD_BF16) BA(float, bf16);
  else BA(bf16, float);
#undef BA
  ADD_RETURN_LAUNCH();
}
