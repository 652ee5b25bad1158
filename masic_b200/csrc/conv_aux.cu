// conv_aux.cu — weight packing for conv_tc.cu, the GDN re-parametrisation, and a plain
// CUDA-core direct convolution used (a) as the on-device cross-check of the tensor-core
// path in tests and (b) by callers that want fp32 torch-layout weights without packing.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/masic_b200.h"

namespace {

constexpr int KBLK = 64;

__global__ void pack_weights_kernel(const float* __restrict__ w, int kind, int transposed, int k,
                                    int c_in, int c_out, int c_out_pad, int ncb, long total,
                                    __nv_bfloat16* __restrict__ dst) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % KBLK);
  long r = i / KBLK;
  const int co = (int)(r % c_out_pad);
  const int kb = (int)(r / c_out_pad);
  const int tap = kb / ncb, cb = kb % ncb;
  const int ci = cb * KBLK + c;
  float v = 0.0f;
  if (ci < c_in || kind == MASIC_CONV_XFOLD4) {
    if (kind == MASIC_DECONV_S2_SUBPIX) {
      // 3x3 stride-1 taps over the INPUT grid, N = (py, px, co): ky = py + 2*(1 - dy)
      const int dy = tap / 3 - 1, dx = tap % 3 - 1;
      const int phase = co / c_out, cc = co % c_out;
      if (phase < 4) {
        const int py = phase >> 1, px = phase & 1;
        const int ky = py + 2 * (1 - dy), kx = px + 2 * (1 - dx);
        if (ky >= 0 && ky < 5 && kx >= 0 && kx < 5)
          v = w[((static_cast<long>(ci) * c_out + cc) * 5 + ky) * 5 + kx];
      }
    } else if (kind == MASIC_CONV_XFOLD4) {
      // k-block = (ky, group): group 0 holds taps kx = 0..3 as column j*16 + ch (window at pixel 2*ox),
      // group 1 holds kx = 4 in columns 32..47 (j = 2 of the window at pixel 2*ox+2);
      // `c_in` here is the REAL channel count of w (<= 16)
      const int ky = kb >> 1, grp = kb & 1;
      const int j = c >> 4, ch = c & 15;
      const int kx = grp ? (j == 2 ? 4 : 99) : j;
      if (co < c_out && ch < c_in && kx < 5)
        v = w[((static_cast<long>(co) * c_in + ch) * 5 + ky) * 5 + kx];
    } else if (co < c_out) {
      int ky = tap / k, kx = tap % k;
      if (transposed) {
        if (kind == MASIC_CONV) { ky = k - 1 - ky; kx = k - 1 - kx; }  // stride-1 transposed = flipped conv
        v = w[((static_cast<long>(ci) * c_out + co) * k + ky) * k + kx];
      } else {
        v = w[((static_cast<long>(co) * c_in + ci) * k + ky) * k + kx];
      }
    }
  }
  dst[i] = __float2bfloat16_rn(v);
}

__global__ void gdn_prepare_kernel(const float* __restrict__ beta, const float* __restrict__ gamma,
                                   int c, float beta_bound, float gamma_bound, float pedestal,
                                   float* __restrict__ beta_out, float* __restrict__ gamma_f32,
                                   __nv_bfloat16* __restrict__ gamma_bf16) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c) {
    const float b = fmaxf(beta[i], beta_bound);
    beta_out[i] = b * b - pedestal;
  }
  if (i < c * c) {
    const float g = fmaxf(gamma[i], gamma_bound);
    const float gp = g * g - pedestal;
    if (gamma_f32) gamma_f32[i] = gp;
    if (gamma_bf16) gamma_bf16[i] = __float2bfloat16_rn(gp);
  }
}

__global__ void conv_direct_kernel(const __nv_bfloat16* __restrict__ in, int n, int h_in, int w_in,
                                   int in_cpitch, int in_coff, int c_in,
                                   const float* __restrict__ w, int transposed, int k, int stride,
                                   uint32_t tap_mask, const float* __restrict__ bias, int c_out,
                                   int h_out, int w_out, float* __restrict__ out, int out_cpitch,
                                   int out_coff, int round_w) {
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = (long)n * h_out * w_out * c_out;
  if (idx >= total) return;
  const int co = (int)(idx % c_out);
  long r = idx / c_out;
  const int ox = (int)(r % w_out); r /= w_out;
  const int oy = (int)(r % h_out);
  const int ni = (int)(r / h_out);
  const int pad = k / 2;
  float acc = bias ? bias[co] : 0.0f;
  for (int ky = 0; ky < k; ++ky)
    for (int kx = 0; kx < k; ++kx) {
      if (tap_mask && !((tap_mask >> (ky * k + kx)) & 1)) continue;
      int iy, ix;
      if (!transposed) {
        iy = oy * stride - pad + ky; ix = ox * stride - pad + kx;
      } else {
        const int ty = oy + pad - ky, tx = ox + pad - kx;
        if (ty < 0 || tx < 0) continue;
        if (ty % stride || tx % stride) continue;
        iy = ty / stride; ix = tx / stride;
      }
      if (iy < 0 || iy >= h_in || ix < 0 || ix >= w_in) continue;
      const __nv_bfloat16* ip = in + ((static_cast<long>(ni) * h_in + iy) * w_in + ix) * in_cpitch + in_coff;
      for (int ci = 0; ci < c_in; ++ci) {
        float wv = transposed ? w[((static_cast<long>(ci) * c_out + co) * k + ky) * k + kx]
                              : w[((static_cast<long>(co) * c_in + ci) * k + ky) * k + kx];
        if (round_w) wv = __bfloat162float(__float2bfloat16_rn(wv));
        acc = fmaf(__bfloat162float(ip[ci]), wv, acc);
      }
    }
  out[((static_cast<long>(ni) * h_out + oy) * w_out + ox) * out_cpitch + out_coff + co] = acc;
}

}  // namespace

extern "C" int64_t masic_packed_weight_bytes(int kind, int ksize, int c_in, int c_out_pad) {
  if (kind == MASIC_CONV_XFOLD4) return static_cast<int64_t>(10) * c_out_pad * KBLK * 2;   // (ky, group) blocks
  const int taps = (kind == MASIC_DECONV_S2_SUBPIX) ? 9 : ksize * ksize;
  const int ncb = (c_in + KBLK - 1) / KBLK;
  return static_cast<int64_t>(taps) * ncb * c_out_pad * KBLK * 2;
}

extern "C" int masic_pack_conv_weights(const float* w, int kind, int transposed, int ksize, int c_in,
                                       int c_out, int c_out_pad, void* dst, void* stream) {
  if (!w || !dst || c_in <= 0 || c_out <= 0) return MASIC_EINVAL;
  if (kind == MASIC_DECONV_S2_SUBPIX) {
    if (ksize != 5 || 4 * c_out > c_out_pad) return MASIC_EINVAL;
  } else if (c_out > c_out_pad) {
    return MASIC_EINVAL;
  }
  if (kind == MASIC_CONV_XFOLD4 && (ksize != 5 || c_in > 16 || transposed)) return MASIC_EINVAL;
  const int ncb = (kind == MASIC_CONV_XFOLD4) ? 1 : (c_in + KBLK - 1) / KBLK;
  const long total = masic_packed_weight_bytes(kind, ksize, c_in, c_out_pad) / 2;
  const int bs = 256;
  pack_weights_kernel<<<(unsigned)((total + bs - 1) / bs), bs, 0, static_cast<cudaStream_t>(stream)>>>(
      w, kind, transposed, ksize, c_in, c_out, c_out_pad, ncb, total, static_cast<__nv_bfloat16*>(dst));
  return (int)cudaGetLastError();
}

extern "C" int masic_gdn_prepare(const float* beta, const float* gamma, int c, float beta_min,
                                 float* beta_out, float* gamma_out_f32, void* gamma_out_bf16,
                                 void* stream) {
  if (!beta || !gamma || !beta_out || c <= 0) return MASIC_EINVAL;
  // compressai/ops/parametrizers.py:49-64 — pedestal = (2^-18)^2, bound = sqrt(minimum + pedestal)
  const float pedestal = 1.4551915228366852e-11f;          // 2^-36
  const float beta_bound = sqrtf(beta_min + pedestal);
  const float gamma_bound = 3.814697265625e-06f;           // 2^-18
  const int bs = 256, total = c * c;
  gdn_prepare_kernel<<<(total + bs - 1) / bs, bs, 0, static_cast<cudaStream_t>(stream)>>>(
      beta, gamma, c, beta_bound, gamma_bound, pedestal, beta_out, gamma_out_f32,
      static_cast<__nv_bfloat16*>(gamma_out_bf16));
  return (int)cudaGetLastError();
}

extern "C" int masic_conv_direct_nhwc(const void* in, int n, int h_in, int w_in, int in_cpitch,
                                      int in_coff, int c_in, const float* w, int transposed, int ksize,
                                      int stride, uint32_t tap_mask, const float* bias, int c_out,
                                      float* out_f32, int out_cpitch, int out_coff, int round_w_bf16,
                                      void* stream) {
  if (!in || !w || !out_f32 || (stride != 1 && stride != 2)) return MASIC_EINVAL;
  int h_out, w_out;
  if (!transposed) { h_out = (h_in + stride - 1) / stride; w_out = (w_in + stride - 1) / stride; }
  else { h_out = h_in * stride; w_out = w_in * stride; }
  const long total = (long)n * h_out * w_out * c_out;
  const int bs = 128;
  conv_direct_kernel<<<(unsigned)((total + bs - 1) / bs), bs, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), n, h_in, w_in, in_cpitch, in_coff, c_in, w, transposed,
      ksize, stride, tap_mask, bias, c_out, h_out, w_out, out_f32, out_cpitch, out_coff, round_w_bf16);
  return (int)cudaGetLastError();
}
