// C ABI of the implicit-GEMM convolution building blocks (bf16 tensor-core backend) so that they can be
// tested in isolation against torch.nn.functional.conv2d; the learner calls the same internals from
// atarinet.cu.  See include/torchbeast_b200.h.
#include "conv_implicit.cuh"
#include "gemm_tc.cuh"
#include "net_kernels.cuh"

using namespace tb;

extern "C" {

int tb_conv_nhwc_bf16_fwd(const void* act_bf16, const float* weight, const float* bias, int64_t Nf, int H, int W, int C, int KH,
                          int KW, int S, int O, int relu, void* out_bf16, void* pack_scratch_bf16, void* stream) {
  TB_REQUIRE(act_bf16 && weight && bias && out_bf16 && pack_scratch_bf16, "tb_conv_nhwc_bf16_fwd: null pointer");
  TB_REQUIRE(conv_tc_implicit_applicable(H, W, C, KH, KW, S, O), "tb_conv_nhwc_bf16_fwd: unsupported shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t K = int64_t(KH) * KW * C;
  int rc = pack_weights_bf16(weight, pack_scratch_bf16, O, KH * KW, C, K, st);
  if (rc) return rc;
  TcEpilogue te;
  te.C16 = static_cast<__nv_bfloat16*>(out_bf16); te.ldc16 = O; te.bias = bias; te.relu = relu; te.tag = "conv_fwd";
  return conv_tc_fwd_implicit(act_bf16, pack_scratch_bf16, Nf, H, W, C, KH, KW, S, O, te, st);
}

int tb_conv_nhwc_bf16_dgrad(const void* dy_bf16, const float* weight, const void* act_bf16, int64_t Nf, int H, int W, int C, int KH,
                            int KW, int S, int O, void* dx_bf16, void* pack_scratch_bf16, void* stream) {
  TB_REQUIRE(dy_bf16 && weight && dx_bf16 && pack_scratch_bf16, "tb_conv_nhwc_bf16_dgrad: null pointer");
  TB_REQUIRE(conv_tc_dgrad_implicit_applicable(H, W, C, KH, KW, S, O), "tb_conv_nhwc_bf16_dgrad: unsupported shape");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = pack_dgrad_weights_bf16(weight, pack_scratch_bf16, O, C, KH, KW, S, st);
  if (rc) return rc;
  TcEpilogue te;
  te.C16 = static_cast<__nv_bfloat16*>(dx_bf16); te.ldc16 = C;
  te.mask16 = static_cast<const __nv_bfloat16*>(act_bf16); te.ldmask = C; te.tag = "conv_dgrad";
  return conv_tc_dgrad_implicit(dy_bf16, pack_scratch_bf16, Nf, H, W, C, KH, KW, S, O, te, st);
}

int tb_conv_nhwc_bf16_wgrad(const void* dy_bf16, const void* act_bf16, int64_t Nf, int H, int W, int C, int KH, int KW, int S, int O,
                            float* dweight, float* partial, int64_t partial_floats, void* stream) {
  TB_REQUIRE(conv_tc_implicit_applicable(H, W, C, KH, KW, S, O) && O <= 64, "tb_conv_nhwc_bf16_wgrad: unsupported shape");
  return conv_tc_wgrad_implicit(dy_bf16, act_bf16, Nf, H, W, C, KH, KW, S, O, dweight, KH * KW, C, 1.0f, partial, partial_floats,
                                "conv_wgrad", (cudaStream_t)stream);
}

int tb_conv1_u8_fwd(const uint8_t* frame, const float* weight, const float* bias, int64_t N, int H, int W, int S, int relu,
                    void* out_bf16, void* image_bf16, void* pack_scratch_bf16, void* stream) {
  TB_REQUIRE(frame && weight && bias && out_bf16 && image_bf16 && pack_scratch_bf16, "tb_conv1_u8_fwd: null pointer");
  TB_REQUIRE(conv_u8_implicit_applicable(4, H, W, 8, 8, S, 32), "tb_conv1_u8_fwd: unsupported shape");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = pack_weights_bf16(weight, pack_scratch_bf16, 32, 1, 256, 256, st);  // reference flattening (c, kh, kw) kept
  if (rc) return rc;
  rc = frames_u8_to_bf16(frame, image_bf16, N * 4 * H * W, st);
  if (rc) return rc;
  TcEpilogue te;
  te.C16 = static_cast<__nv_bfloat16*>(out_bf16); te.ldc16 = 32; te.bias = bias; te.scale = 1.0f / 255.0f; te.relu = relu;
  te.tag = "conv1_fwd";
  return conv_u8_fwd_implicit(image_bf16, pack_scratch_bf16, N, H, W, S, te, st);
}

int tb_conv1_u8_wgrad(const void* dy_bf16, const void* image_bf16, int64_t N, int H, int W, int S, float* dweight, float* partial,
                      int64_t partial_floats, void* stream) {
  TB_REQUIRE(conv_u8_implicit_applicable(4, H, W, 8, 8, S, 32), "tb_conv1_u8_wgrad: unsupported shape");
  return conv_u8_wgrad_implicit(dy_bf16, image_bf16, N, H, W, S, dweight, 1.0f / 255.0f, partial, partial_floats, "conv1_wgrad",
                                (cudaStream_t)stream);
}

}  // extern "C"
