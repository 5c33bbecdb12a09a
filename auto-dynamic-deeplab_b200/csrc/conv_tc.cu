// conv_tc.cu — tcgen05 / TMEM implicit-GEMM convolution for bf16 activations (sm_100a).
// PLACEHOLDER until the tensor-core kernel lands: the entry points exist (the C ABI is stable) and
// report ADD_ERR_UNSUPPORTED so callers fail loudly instead of silently taking another path.
#include "common.cuh"

extern "C" int64_t add_conv2d_tc_packed_bytes(int cin, int cout, int kh, int kw) {
  (void)cin; (void)cout; (void)kh; (void)kw;
  return ADD_ERR_UNSUPPORTED;
}

extern "C" int add_conv2d_tc_pack(const float* w_hwio, int cin, int cout, int kh, int kw, void* packed_host) {
  (void)w_hwio; (void)cin; (void)cout; (void)kh; (void)kw; (void)packed_host;
  return ADD_ERR_UNSUPPORTED;
}

extern "C" int add_conv2d_tc_fwd(const add_tensor_t* x, const add_tensor_t* y, const void* w_packed,
                                 const float* bias, int kh, int kw, int stride, int pad, int dil,
                                 uint32_t flags, void* stream) {
  (void)x; (void)y; (void)w_packed; (void)bias; (void)kh; (void)kw; (void)stride; (void)pad; (void)dil;
  (void)flags; (void)stream;
  return ADD_ERR_UNSUPPORTED;
}
