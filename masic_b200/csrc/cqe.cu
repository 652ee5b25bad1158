// cqe.cu — memory-bound glue kernels of the cross quality enhancement network `Independent_EN`
// (coremasic/mywork/MASIC.py:1436-1501).  The 3x3 convolutions of its Enhancement_Blocks run on the
// tensor-core conv kernel (conv_tc.cu) with LeakyReLU and the residual adds fused into the epilogue; what is left is
// the mask-weighted blending of an image / feature map with the homography-warped map of the OTHER view:
//   MASIC.py:1470-1471   cat(x_other_warp * w[0], x_self * w[1])              3 + 3 channels, images (NCHW fp32)
//   MASIC.py:1479-1482   cat(out_self * w[1], warp(out_other, H) * w[0])      32 + 32 channels, features (NHWC bf16)
//   MASIC.py:1493-1496   x_hat2 = conv2(out) + identity
// kornia.warp_perspective = bilinear, zero padding, align_corners=True (same coordinate chain as image.cu).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/masic_b200.h"

namespace {

struct Bilin { int x0, y0; float w00, w01, w10, w11; bool in_x0, in_x1, in_y0, in_y1; };

__device__ __forceinline__ Bilin bilin_coords(const double* __restrict__ t, int x, int y, int h, int w, int ho, int wo) {
  const double xn = ((double)x / (double)(wo - 1) - 0.5) * 2.0;
  const double yn = ((double)y / (double)(ho - 1) - 0.5) * 2.0;
  const double q0 = xn * t[0] + yn * t[1] + t[2];
  const double q1 = xn * t[3] + yn * t[4] + t[5];
  const double q2 = xn * t[6] + yn * t[7] + t[8];
  const double den = fabs(q2) >= 0.25 ? q2 : q2 + 1e-8;
  const double sc = fabs(q2) > 1e-8 ? 1.0 / den : 1.0;
  const double ixd = ((q0 * sc + 1.0) / 2.0) * (double)(w - 1);
  const double iyd = ((q1 * sc + 1.0) / 2.0) * (double)(h - 1);
  const double fxd = floor(ixd), fyd = floor(iyd);
  const float ix = (float)(ixd - fxd), iy = (float)(iyd - fyd);
  Bilin b;
  const bool finite = fabs(ixd) < 1e9 && fabs(iyd) < 1e9;
  b.x0 = finite ? (int)fxd : -10; b.y0 = finite ? (int)fyd : -10;
  b.w00 = (1.0f - ix) * (1.0f - iy); b.w01 = ix * (1.0f - iy); b.w10 = (1.0f - ix) * iy; b.w11 = ix * iy;
  b.in_x0 = b.x0 >= 0 && b.x0 < w; b.in_x1 = b.x0 + 1 >= 0 && b.x0 + 1 < w;
  b.in_y0 = b.y0 >= 0 && b.y0 < h; b.in_y1 = b.y0 + 1 >= 0 && b.y0 + 1 < h;
  return b;
}

// out[p][0:3] = a[:, p] * w[0][p], out[p][3:6] = b[:, p] * w[1][p], out[p][6:16] = 0      (bf16, pitch 16)
__global__ void __launch_bounds__(256)
blend_images_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ wgt, int n,
                    long hw, __nv_bfloat16* __restrict__ out) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)n * hw) return;
  const int bi = (int)(i / hw);
  const long p = i - bi * hw;
  const float w0 = wgt[((long)bi * 2) * hw + p], w1 = wgt[((long)bi * 2 + 1) * hw + p];
  float v[6];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    v[c] = a[((long)bi * 3 + c) * hw + p] * w0;
    v[3 + c] = b[((long)bi * 3 + c) * hw + p] * w1;
  }
  uint32_t q[8];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    q[j] = *reinterpret_cast<uint32_t*>(&h2);
  }
#pragma unroll
  for (int j = 3; j < 8; ++j) q[j] = 0u;
  uint4* o = reinterpret_cast<uint4*>(out + i * 16);
  o[0] = make_uint4(q[0], q[1], q[2], q[3]);
  o[1] = make_uint4(q[4], q[5], q[6], q[7]);
}

// out[p][0:C] = self[p][0:C] * w[1][p];  out[p][C:2C] = bilinear(other, T)[p][0:C] * w[0][p]
// one thread per (pixel, 8-channel group): 16-byte loads / stores; C multiple of 8
__global__ void __launch_bounds__(256)
feature_fuse_kernel(const __nv_bfloat16* __restrict__ self, int self_pitch, const __nv_bfloat16* __restrict__ other,
                    int other_pitch, int C, const float* __restrict__ wgt, const double* __restrict__ T, int n, int h,
                    int w, __nv_bfloat16* __restrict__ out, int out_pitch) {
  const int groups = C / 8;
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = (long)n * h * w * 2 * groups;
  if (i >= total) return;
  const int g = (int)(i % (2 * groups));
  const long p = i / (2 * groups);
  const long hw = (long)h * w;
  const int bi = (int)(p / hw);
  const long pl = p - bi * hw;
  auto unpack = [](uint4 u, float* f) {
    const uint32_t q[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q[j]));
      f[2 * j] = t.x; f[2 * j + 1] = t.y;
    }
  };
  float acc[8];
  float scale;
  if (g < groups) {
    scale = wgt[((long)bi * 2 + 1) * hw + pl];
    unpack(__ldg(reinterpret_cast<const uint4*>(self + p * self_pitch + 8 * g)), acc);
  } else {
    const int gg = g - groups;
    scale = wgt[((long)bi * 2) * hw + pl];
    const int y = (int)(pl / w), x = (int)(pl - (long)y * w);
    const Bilin b = bilin_coords(T + bi * 9, x, y, h, w, h, w);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    const __nv_bfloat16* base = other + (long)bi * hw * other_pitch + 8 * gg;
    auto tap = [&](bool ok, int yy, int xx, float wt) {
      if (!ok) return;
      float f[8];
      unpack(__ldg(reinterpret_cast<const uint4*>(base + ((long)yy * w + xx) * other_pitch)), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j] * wt;
    };
    tap(b.in_y0 && b.in_x0, b.y0, b.x0, b.w00);
    tap(b.in_y0 && b.in_x1, b.y0, b.x0 + 1, b.w01);
    tap(b.in_y1 && b.in_x0, b.y0 + 1, b.x0, b.w10);
    tap(b.in_y1 && b.in_x1, b.y0 + 1, b.x0 + 1, b.w11);
  }
  uint32_t q[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[2 * j] * scale, acc[2 * j + 1] * scale);
    q[j] = *reinterpret_cast<uint32_t*>(&h2);
  }
  *reinterpret_cast<uint4*>(out + p * out_pitch + 8 * g) = make_uint4(q[0], q[1], q[2], q[3]);
}

// out_nchw[b][c][p] = conv_out_nhwc[b][p][c] + identity_nchw[b][c][p],  c < 3
__global__ void __launch_bounds__(256)
residual_image_kernel(const float* __restrict__ conv_out, int pitch, const float* __restrict__ identity, int n, long hw,
                      float* __restrict__ out) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)n * hw) return;
  const int bi = (int)(i / hw);
  const long p = i - bi * hw;
  const float4 v = *reinterpret_cast<const float4*>(conv_out + i * pitch);
  const float vv[3] = {v.x, v.y, v.z};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const long o = ((long)bi * 3 + c) * hw + p;
    out[o] = vv[c] + identity[o];
  }
}

}  // namespace

#define S(stream) static_cast<cudaStream_t>(stream)

extern "C" int masic_cqe_blend_images(const float* a_nchw, const float* b_nchw, const float* weights_nchw2, int n, int h,
                                      int w, void* out_nhwc16_bf16, void* stream) {
  if (!a_nchw || !b_nchw || !weights_nchw2 || !out_nhwc16_bf16 || n <= 0 || h <= 0 || w <= 0) return MASIC_EINVAL;
  const long total = (long)n * h * w;
  blend_images_kernel<<<(unsigned)((total + 255) / 256), 256, 0, S(stream)>>>(
      a_nchw, b_nchw, weights_nchw2, n, (long)h * w, static_cast<__nv_bfloat16*>(out_nhwc16_bf16));
  return (int)cudaGetLastError();
}

extern "C" int masic_cqe_feature_fuse(const void* self_bf16, int self_pitch, const void* other_bf16, int other_pitch,
                                      int c, const float* weights_nchw2, const double* t_prepared, int n, int h, int w,
                                      void* out_bf16, int out_pitch, void* stream) {
  if (!self_bf16 || !other_bf16 || !weights_nchw2 || !t_prepared || !out_bf16 || c <= 0 || (c % 8) || (self_pitch % 8) ||
      (other_pitch % 8) || (out_pitch % 8) || out_pitch < 2 * c || h < 2 || w < 2)
    return MASIC_EINVAL;
  const long total = (long)n * h * w * 2 * (c / 8);
  feature_fuse_kernel<<<(unsigned)((total + 255) / 256), 256, 0, S(stream)>>>(
      static_cast<const __nv_bfloat16*>(self_bf16), self_pitch, static_cast<const __nv_bfloat16*>(other_bf16), other_pitch,
      c, weights_nchw2, t_prepared, n, h, w, static_cast<__nv_bfloat16*>(out_bf16), out_pitch);
  return (int)cudaGetLastError();
}

extern "C" int masic_cqe_residual_image(const float* conv_out_nhwc, int pitch, const float* identity_nchw, int n, int h,
                                        int w, float* out_nchw, void* stream) {
  if (!conv_out_nhwc || !identity_nchw || !out_nchw || pitch < 4 || (pitch % 4) || n <= 0) return MASIC_EINVAL;
  const long total = (long)n * h * w;
  residual_image_kernel<<<(unsigned)((total + 255) / 256), 256, 0, S(stream)>>>(conv_out_nhwc, pitch, identity_nchw, n,
                                                                                (long)h * w, out_nchw);
  return (int)cudaGetLastError();
}
