/*
 * masic_b200.h — C ABI of libmasic_b200.so (CUDA, sm_100a only).
 *
 * The reference (ywz978020607/MASIC, a CompressAI fork) has no FFI for this
 * path: every hot op is a stock ATen call reached through nn.Module.forward.
 * Each entry point below therefore names the reference Python call site it
 * replaces (file:line relative to the reference root) instead of an existing
 * binding.  INTEGRATION.md shows the ctypes stub a maintainer of the reference
 * would add to route those call sites here.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; all pointers are DEVICE pointers
 *     unless the parameter name ends in _host;
 *   - every function returns 0 on success, a negative MASIC_E* code on a
 *     rejected argument, or a positive cudaError_t;  no exceptions cross the ABI;
 *   - the caller owns every buffer; `stream` is a cudaStream_t passed as void*;
 *   - no hidden global state except the lazily resolved driver entry point for
 *     cuTensorMapEncodeTiled and per-function cudaFuncSetAttribute calls.
 *
 * Activation layout inside the library ("NHWC-bf16"): [N][H][W][Cpitch] with
 * bfloat16 elements, Cpitch a multiple of 8 (16-byte rows for TMA).  The
 * entropy/warp kernels take the reference's own NCHW fp32 tensors.
 */
#ifndef MASIC_B200_H
#define MASIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MASIC_OK 0
#define MASIC_EINVAL (-1)   /* bad argument (shape, alignment, enum)        */
#define MASIC_ENOSUP (-2)   /* valid request this build does not implement  */
#define MASIC_EDRIVER (-3)  /* cuTensorMapEncodeTiled unavailable / failed  */

/* ---------------------------------------------------------------- version */
int masic_abi_version(void);                 /* bumps on any signature change */
const char* masic_build_info(void);          /* "sm_100a nvcc 12.9 ..." */

/* ------------------------------------------------------------------ convs */
/* activation fused after bias (reference: nn.ReLU / nn.LeakyReLU(0.01) that
 * follow conv()/deconv() in MASIC.py:173-183, 338-376, 410-444, 678-691)    */
enum { MASIC_ACT_NONE = 0, MASIC_ACT_RELU = 1, MASIC_ACT_LEAKY = 2 };
/* GDN fused into the conv epilogue (reference: compressai/layers/gdn.py:77-92) */
enum { MASIC_GDN_NONE = 0, MASIC_GDN_FWD = 1, MASIC_GDN_INV = 2 };
/* what the implicit GEMM computes */
enum {
  MASIC_CONV = 0,          /* nn.Conv2d(k, stride s, pad k/2)  — models/utils.py:128-135  */
  MASIC_DECONV_S2 = 1,     /* nn.ConvTranspose2d(5, s=2, p=2, op=1) — models/utils.py:138-146 */
  MASIC_DECONV_S2_SUBPIX = 2 /* same op, all 4 output phases stacked on N (for tiny Cout):
                                out buffer is [N][H][W][4*Cout padded], phase-major   */
};

typedef struct MasicConvDesc {
  int kind;            /* MASIC_CONV | MASIC_DECONV_S2 | MASIC_DECONV_S2_SUBPIX          */
  int ksize;           /* 1, 3 or 5 (square)                                            */
  int stride;          /* 1 or 2 (MASIC_CONV only; deconv kinds are stride 2)           */
  uint32_t tap_mask;   /* bit (ky*ksize+kx) set = live tap; 0 means "all taps".
                          MaskedConv2d mask 'A' (layers.py:68-73) = rows 0,1 + (2,0),(2,1) */
  int n, h_in, w_in;   /* batch and INPUT spatial size                                   */
  int c_in;            /* input channels consumed (multiple of 16)                       */
  int c_out;           /* real output channels                                           */
  int c_out_pad;       /* rows per k-block in the packed weights (multiple of n_tile)    */
  int n_tile;          /* accumulator width: multiple of 16, <= 256                      */
  /* input activation buffer, NHWC bf16 */
  const void* in; int in_cpitch; int in_coff;   /* channel pitch / first channel read    */
  /* packed weights from masic_pack_conv_weights(), bf16 [kblocks][c_out_pad][64]        */
  const void* w_packed;
  const float* bias;   /* [c_out_pad] fp32 or NULL                                       */
  /* output buffer NHWC, bf16 or fp32, written at channel offset out_coff               */
  void* out; int out_cpitch; int out_coff; int out_fp32;
  /* per n-tile activation (MASIC_ACT_*), act[i] applies to channels [i*n_tile,(i+1)*n_tile) */
  uint8_t act[32];
  /* fused GDN: requires c_out == n_tile == 128.  gamma_packed: bf16 [128][128] = gamma'
   * (re-parametrised, gdn.py:80-83), beta: fp32 [128] = beta'                           */
  int gdn; const void* gamma_packed; const float* beta;
  /* optional per-pixel multiplier applied last (mask-weighted fusion, MASIC.py:827):
   * out *= rowscale[((n*H+y)*W+x)*rs_stride + rs_off]                                   */
  const float* rowscale; int rs_stride; int rs_off;
} MasicConvDesc;

typedef struct MasicConvPlan MasicConvPlan;   /* opaque: tensor maps + tile program */

/* Build the TMA tensor maps and the per-tile load/MMA program for one layer.
 * Buffers named in `desc` are bound into the plan (graph-capturable launches). */
int masic_conv_plan_create(const MasicConvDesc* desc, MasicConvPlan** plan_out);
int masic_conv_plan_launch(const MasicConvPlan* plan, void* stream);
void masic_conv_plan_destroy(MasicConvPlan* plan);
/* useful work of one launch, for roofline accounting */
int masic_conv_plan_info(const MasicConvPlan* plan, double* flops, double* hbm_bytes,
                         int* n_work_items, int* smem_bytes);

/* Pack torch-layout fp32 weights into the bf16 k-block layout the plan reads.
 *   transposed = 0: w is (c_out, c_in, k, k)   [nn.Conv2d]
 *   transposed = 1: w is (c_in, c_out, k, k)   [nn.ConvTranspose2d]
 *   subpix     = 1: build the 3x3 / 4-phase matrix of MASIC_DECONV_S2_SUBPIX
 * dst must hold masic_packed_weight_bytes() bytes.                             */
int64_t masic_packed_weight_bytes(int kind, int ksize, int c_in, int c_out_pad);
int masic_pack_conv_weights(const float* w, int kind, int transposed, int ksize,
                            int c_in, int c_out, int c_out_pad, void* dst, void* stream);
/* GDN re-parametrisation (parametrizers.py:61-64) + bf16 pack of gamma:
 *   beta'  = max(beta,  sqrt(beta_min + 2^-36))^2 - 2^-36
 *   gamma' = max(gamma, 2^-18)^2 - 2^-36                                        */
int masic_gdn_prepare(const float* beta, const float* gamma, int c, float beta_min,
                      float* beta_out, float* gamma_out_f32, void* gamma_out_bf16,
                      void* stream);

/* Plain direct convolution on CUDA cores (test oracle on the device + the
 * small-channel layers).  Same NHWC-bf16 activations, fp32 torch-layout weights. */
int masic_conv_direct_nhwc(const void* in, int n, int h_in, int w_in, int in_cpitch, int in_coff,
                           int c_in, const float* w, int transposed, int ksize, int stride,
                           uint32_t tap_mask, const float* bias, int c_out,
                           float* out_f32, int out_cpitch, int out_coff, int round_w_bf16,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MASIC_B200_H */
