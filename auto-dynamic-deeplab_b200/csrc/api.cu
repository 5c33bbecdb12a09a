// api.cu — library-level entry points (status strings, version, device query).
#include "common.cuh"

extern "C" const char* add_status_string(int status) {
  switch (status) {
    case ADD_OK: return "ok";
    case ADD_ERR_BAD_ARG: return "bad argument (null pointer, negative size, shape mismatch)";
    case ADD_ERR_UNSUPPORTED: return "unsupported shape/dtype/alignment for the sm_100a kernels";
    case ADD_ERR_CUDA: return "CUDA launch error (no device, or invalid launch configuration)";
    case ADD_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown status";
  }
}

thread_local int g_add_last_cuda_error = 0;

extern "C" const char* add_last_cuda_error(void) {
  return g_add_last_cuda_error == 0 ? "none" : cudaGetErrorString((cudaError_t)g_add_last_cuda_error);
}

extern "C" int add_version(void) { return 100; }  // 0.1.0

extern "C" int add_device_sm_count(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return ADD_ERR_CUDA;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return ADD_ERR_CUDA;
  return sms;
}

int g_add_pdl = 1;
/* 1 (default) = tcgen05 kernels are launched with programmatic stream serialization (their prologue overlaps the
 * previous kernel's tail; they wait for it with griddepcontrol.wait before touching activations); 0 = plain launches. */
extern "C" int add_set_pdl(int on) { g_add_pdl = on ? 1 : 0; return ADD_OK; }

int g_add_grid_pct = 100;
/* Tuning: persistent kernels launch (pct/100) x their default CTA count (default 100), leaving room for kernels of
 * other streams of the captured graph to co-reside. */
int g_add_conv_grid_pct = 100;     /* same, for the small 1x1 / few-tap convs of conv2d_tc_persistent_kernel only */
extern "C" int add_set_persistent_grid_pct(int pct) {
  /* pct in 10..100 scales every persistent kernel; 1000 * c + p (c, p in 10..100) scales the small convs by c % and the rest by p % */
  int conv = 0;
  if (pct >= 1000) { conv = pct / 1000; pct %= 1000; }
  if (pct < 10 || pct > 100 || (conv != 0 && (conv < 10 || conv > 100))) return ADD_ERR_BAD_ARG;
  g_add_grid_pct = pct;
  g_add_conv_grid_pct = conv ? conv : 100;
  return ADD_OK;
}
