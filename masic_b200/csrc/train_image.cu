// train_image.cu — image-domain backward kernels of the training step (NCHW fp32, <= 8 channels), sm_100a.
//
// Replaces the autograd backward (coremasic/mywork/newtrain_codec_real.py:134) of:
//   nn.MSELoss in RateDistortionLoss                     newtrain_codec_real.py:78-79
//   kornia.warp_perspective (F.grid_sample, bilinear, zeros, align_corners=True) w.r.t. its source   MASIC.py:821,833
//   Encoder2.pre_conv / Decoder2.after_conv / mask2weights.maskconv  MASIC.py:559,600,475-488 (conv_small_kernel's layers)
//   GDN(3) / GDN(3, inverse)  (pre_gdn, after_gdn)       MASIC.py:560,599; compressai/layers/gdn.py:77-92
//   the softmax over the 3 mask weights                  MASIC.py:497-502
//   g_a_conv1 (3 -> 128) and g_s_conv4 (128 -> 3) weight gradients (3-channel side: too thin for the tensor cores)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/masic_b200.h"

namespace {

constexpr int MAXC = 8;

// g = scale * (xh - x) [+ addend]
__global__ void __launch_bounds__(256)
mse_grad_kernel(const float* __restrict__ xh, const float* __restrict__ x, const float* __restrict__ addend, float scale,
                long n, float* __restrict__ g) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = scale * (xh[i] - x[i]);
  if (addend) v += addend[i];
  g[i] = v;
}

// ---------------------------------------------------------------- warp backward (scatter into the source grid)
__global__ void __launch_bounds__(256)
warp_bwd_kernel(const float* __restrict__ g0, const float* __restrict__ g1, int n, int c, int h, int w, int ho, int wo,
                const double* __restrict__ T, float* __restrict__ dsrc) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int b = blockIdx.z;
  if (x >= wo) return;
  const double* t = T + b * 9;
  // same coordinate chain as warp_kernel (image.cu)
  const double xn = ((double)x / (double)(wo - 1) - 0.5) * 2.0;
  const double yn = ((double)y / (double)(ho - 1) - 0.5) * 2.0;
  const double q0 = xn * t[0] + yn * t[1] + t[2];
  const double q1 = xn * t[3] + yn * t[4] + t[5];
  const double q2 = xn * t[6] + yn * t[7] + t[8];
  const double den = fabs(q2) >= 0.25 ? q2 : q2 + 1e-8;
  const double sc = fabs(q2) > 1e-8 ? 1.0 / den : 1.0;
  const double gx = q0 * sc, gy = q1 * sc;
  const double ixd = ((gx + 1.0) / 2.0) * (double)(w - 1);
  const double iyd = ((gy + 1.0) / 2.0) * (double)(h - 1);
  const double fxd = floor(ixd), fyd = floor(iyd);
  const float ix = (float)(ixd - fxd), iy = (float)(iyd - fyd);
  const float wx1 = ix, wx0 = 1.0f - ix, wy1 = iy, wy0 = 1.0f - iy;
  const bool finite = fabs(ixd) < 1e9 && fabs(iyd) < 1e9;
  const int x0 = finite ? (int)fxd : -10, y0 = finite ? (int)fyd : -10;
  const bool in_x0 = x0 >= 0 && x0 < w, in_x1 = x0 + 1 >= 0 && x0 + 1 < w;
  const bool in_y0 = y0 >= 0 && y0 < h, in_y1 = y0 + 1 >= 0 && y0 + 1 < h;
  for (int ch = 0; ch < c; ++ch) {
    const long o = ((long)(b * c + ch) * ho + y) * wo + x;
    float gv = g0[o];
    if (g1) gv += g1[o];
    float* s = dsrc + ((long)(b * c + ch) * h) * w;
    if (in_y0 && in_x0) atomicAdd(s + (long)y0 * w + x0, gv * wx0 * wy0);
    if (in_y0 && in_x1) atomicAdd(s + (long)y0 * w + x0 + 1, gv * wx1 * wy0);
    if (in_y1 && in_x0) atomicAdd(s + (long)(y0 + 1) * w + x0, gv * wx0 * wy1);
    if (in_y1 && in_x1) atomicAdd(s + (long)(y0 + 1) * w + x0 + 1, gv * wx1 * wy1);
  }
}

// ---------------------------------------------------------------- small conv backward
// Forward (conv_small_kernel): conv:   out[co][oy][ox] = b + sum in[ci][oy*s+ky-pad][ox*s+kx-pad] W[co][ci][ky][kx]
//                              transposed (s = 1): out[co][y][x] = b + sum in[ci][y+pad-ky][x+pad-kx] W[ci][co][ky][kx]
// g is dL/d(out) BEFORE the activation mask; act_out (forward output, post-ReLU) masks it: g' = g * (act_out > 0).
struct SCArgs {
  const float* in0; const float* in1; int c0, c1;
  int n, h, w, ho, wo, c_out, k, stride, transposed;
  const float* g; const float* act_out;
  const float* wt;
};
__device__ __forceinline__ float sc_in(const SCArgs& a, int b, int ci, int y, int x) {
  if (y < 0 || y >= a.h || x < 0 || x >= a.w) return 0.0f;
  return ci < a.c0 ? a.in0[((long)(b * a.c0 + ci) * a.h + y) * a.w + x]
                   : a.in1[((long)(b * a.c1 + ci - a.c0) * a.h + y) * a.w + x];
}
__device__ __forceinline__ float sc_g(const SCArgs& a, int b, int co, int y, int x) {
  if (y < 0 || y >= a.ho || x < 0 || x >= a.wo) return 0.0f;
  const long o = ((long)(b * a.c_out + co) * a.ho + y) * a.wo + x;
  const float v = a.g[o];
  return (a.act_out && !(a.act_out[o] > 0.0f)) ? 0.0f : v;
}

// Persistent blocks walk 16x16 tiles of OUTPUT pixels: the masked gradient tile (c_out x 256) and the input halo tile
// (cin x (16 s + k - 1)^2 for a conv, cin x (16 + k - 1)^2 for the transposed layer) are staged in shared memory;
// thread t owns weight elements t, t + 256, ... (and the bias), accumulates over all its tiles in registers and issues
// one atomicAdd per owned element at the end.
constexpr int SW_T = 16;
constexpr int SW_MAXW = 2;                       // owned weight elements per thread: 450 + 3 <= 2 * 256
__global__ void __launch_bounds__(256)
small_conv_wgrad_kernel(const SCArgs a, float* __restrict__ dw, float* __restrict__ db) {
  extern __shared__ float sm[];
  const int cin = a.c0 + a.c1, kk = a.k * a.k, pad = a.k / 2;
  const int nw = a.c_out * cin * kk;
  const int ntot = nw + (db ? a.c_out : 0);
  const int span = a.transposed ? SW_T + a.k - 1 : (SW_T - 1) * a.stride + a.k;      // input tile edge
  float* s_g = sm;                                // [c_out][256]
  float* s_in = sm + a.c_out * SW_T * SW_T;       // [cin][span][span]
  const int tiles_x = (a.wo + SW_T - 1) / SW_T, tiles_y = (a.ho + SW_T - 1) / SW_T;
  const int n_tiles = tiles_x * tiles_y * a.n;
  // decode owned elements once
  int e_co[SW_MAXW], e_ci[SW_MAXW], e_ky[SW_MAXW], e_kx[SW_MAXW];
  float acc[SW_MAXW];
#pragma unroll
  for (int j = 0; j < SW_MAXW; ++j) {
    const int e = threadIdx.x + j * 256;
    acc[j] = 0.0f;
    e_co[j] = -1; e_ci[j] = 0; e_ky[j] = 0; e_kx[j] = 0;
    if (e < nw) {
      const int tap = e % kk;
      if (a.transposed) { e_co[j] = (e / kk) % a.c_out; e_ci[j] = e / (kk * a.c_out); }
      else { e_ci[j] = (e / kk) % cin; e_co[j] = e / (kk * cin); }
      e_ky[j] = tap / a.k; e_kx[j] = tap % a.k;
    } else if (e < ntot) {
      e_co[j] = e - nw; e_ci[j] = -1;
    }
  }
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int tx = t % tiles_x, r = t / tiles_x;
    const int ty = r % tiles_y, b = r / tiles_y;
    const int oy0 = ty * SW_T, ox0 = tx * SW_T;
    // input tile origin: conv reads in[o*s + k - pad]; transposed reads in[o + pad - k]
    const int iy0 = a.transposed ? oy0 + pad - (a.k - 1) : oy0 * a.stride - pad;
    const int ix0 = a.transposed ? ox0 + pad - (a.k - 1) : ox0 * a.stride - pad;
    __syncthreads();
    for (int i = threadIdx.x; i < a.c_out * SW_T * SW_T; i += 256) {
      const int co = i / (SW_T * SW_T), q = i % (SW_T * SW_T);
      s_g[i] = sc_g(a, b, co, oy0 + q / SW_T, ox0 + q % SW_T);
    }
    for (int i = threadIdx.x; i < cin * span * span; i += 256) {
      const int ci = i / (span * span), q = i % (span * span);
      s_in[i] = sc_in(a, b, ci, iy0 + q / span, ix0 + q % span);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SW_MAXW; ++j) {
      if (e_co[j] < 0) continue;
      const float* g = s_g + e_co[j] * SW_T * SW_T;
      float s = 0.0f;
      if (e_ci[j] < 0) {
        for (int q = 0; q < SW_T * SW_T; ++q) s += g[q];
      } else {
        const float* in = s_in + e_ci[j] * span * span;
        // conv: in index = (oy*s + ky, ox*s + kx) relative to the tile; transposed: (oy + k-1-ky, ox + k-1-kx)
        const int by = a.transposed ? a.k - 1 - e_ky[j] : e_ky[j], bx = a.transposed ? a.k - 1 - e_kx[j] : e_kx[j];
        const int st = a.transposed ? 1 : a.stride;
        for (int oy = 0; oy < SW_T; ++oy)
#pragma unroll 4
          for (int ox = 0; ox < SW_T; ++ox)
            s += g[oy * SW_T + ox] * in[(oy * st + by) * span + ox * st + bx];
      }
      acc[j] += s;
    }
  }
#pragma unroll
  for (int j = 0; j < SW_MAXW; ++j) {
    const int e = threadIdx.x + j * 256;
    if (e < nw) atomicAdd(dw + e, acc[j]);
    else if (e < ntot) atomicAdd(db + (e - nw), acc[j]);
  }
}

// Weight gradient of the two 6 <-> 3 channel 5x5 stride-1 layers (pre_conv, after_conv), register-tiled: a thread owns
// the 5 x 3 elements (kx, co) of one (ci, ky) row of the kernel and one row of an 8 x 64 pixel tile; walking the row it
// keeps a sliding window of 5 inputs, so a pixel costs 4 shared-memory loads for 15 FMAs (the generic kernel: 2 per FMA).
// 30 (ci, ky) groups x 8 rows = 240 threads; threads 240..247 sum the bias gradient.  Persistent blocks keep their
// accumulators over all their tiles and add them to dW with one atomic per element at the end.
//   conv:        dW[co][ci][ky][kx]  = sum g[co][y][x] in[ci][y + ky - 2][x + kx - 2]
//   transposed:  dW[ci][co][ky][kx]  = sum g[co][y][x] in[ci][y + 2 - ky][x + 2 - kx]   (the same sums at flipped taps)
constexpr int WG_TW = 64, WG_TH = 8, WG_PW = WG_TW + 4, WG_PH = WG_TH + 4;
__global__ void __launch_bounds__(256)
conv5x5_6x3_wgrad_kernel(const float* __restrict__ in0, const float* __restrict__ in1, int c0, const float* __restrict__ g,
                         int n, int h, int w, int transposed, float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float s_in[6][WG_PH][WG_PW];
  __shared__ float s_g[3][WG_TH][WG_TW];
  const int tid = threadIdx.x, grp = tid >> 3, row = tid & 7;
  const int ci = grp / 5, ky = grp % 5;                 // valid for grp < 30
  float acc[5][3];
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[k][c] = 0.0f;
  const int tiles_x = (w + WG_TW - 1) / WG_TW, tiles_y = (h + WG_TH - 1) / WG_TH;
  const int n_tiles = tiles_x * tiles_y * n;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int tx = t % tiles_x, r = t / tiles_x, ty = r % tiles_y, b = r / tiles_y;
    const int y0 = ty * WG_TH, x0 = tx * WG_TW;
    __syncthreads();
    for (int i = tid; i < 6 * WG_PH * WG_PW; i += 256) {
      const int c = i % WG_PW, py = (i / WG_PW) % WG_PH, ch = i / (WG_PW * WG_PH);
      const int gy = y0 + py - 2, gx = x0 + c - 2;
      float v = 0.0f;
      if (gy >= 0 && gy < h && gx >= 0 && gx < w)
        v = ch < c0 ? __ldg(in0 + ((long)(b * c0 + ch) * h + gy) * w + gx)
                    : __ldg(in1 + ((long)(b * (6 - c0) + ch - c0) * h + gy) * w + gx);
      s_in[ch][py][c] = v;
    }
    for (int i = tid; i < 3 * WG_TH * WG_TW; i += 256) {
      const int c = i % WG_TW, py = (i / WG_TW) % WG_TH, co = i / (WG_TW * WG_TH);
      const int gy = y0 + py, gx = x0 + c;
      s_g[co][py][c] = (gy < h && gx < w) ? __ldg(g + ((long)(b * 3 + co) * h + gy) * w + gx) : 0.0f;
    }
    __syncthreads();
    if (grp < 30) {
      const float* ir = &s_in[ci][row + ky][0];
      float w0 = ir[0], w1 = ir[1], w2 = ir[2], w3 = ir[3];
#pragma unroll 4
      for (int x = 0; x < WG_TW; ++x) {
        const float w4 = ir[x + 4];
        const float g0 = s_g[0][row][x], g1 = s_g[1][row][x], g2 = s_g[2][row][x];
        acc[0][0] = fmaf(g0, w0, acc[0][0]); acc[0][1] = fmaf(g1, w0, acc[0][1]); acc[0][2] = fmaf(g2, w0, acc[0][2]);
        acc[1][0] = fmaf(g0, w1, acc[1][0]); acc[1][1] = fmaf(g1, w1, acc[1][1]); acc[1][2] = fmaf(g2, w1, acc[1][2]);
        acc[2][0] = fmaf(g0, w2, acc[2][0]); acc[2][1] = fmaf(g1, w2, acc[2][1]); acc[2][2] = fmaf(g2, w2, acc[2][2]);
        acc[3][0] = fmaf(g0, w3, acc[3][0]); acc[3][1] = fmaf(g1, w3, acc[3][1]); acc[3][2] = fmaf(g2, w3, acc[3][2]);
        acc[4][0] = fmaf(g0, w4, acc[4][0]); acc[4][1] = fmaf(g1, w4, acc[4][1]); acc[4][2] = fmaf(g2, w4, acc[4][2]);
        w0 = w1; w1 = w2; w2 = w3; w3 = w4;
      }
    } else if (grp == 30 && db) {
      for (int x = 0; x < WG_TW; ++x) {
        acc[0][0] += s_g[0][row][x]; acc[0][1] += s_g[1][row][x]; acc[0][2] += s_g[2][row][x];
      }
    }
  }
  // sum over the 8 rows (threads of a group are 8 consecutive lanes), then one atomic per element
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = acc[k][c];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      acc[k][c] = v;
    }
  if (row == 0) {
    if (grp < 30) {
#pragma unroll
      for (int kx = 0; kx < 5; ++kx)
#pragma unroll
        for (int co = 0; co < 3; ++co) {
          const int idx = transposed ? ((ci * 3 + co) * 5 + (4 - ky)) * 5 + (4 - kx) : ((co * 6 + ci) * 5 + ky) * 5 + kx;
          atomicAdd(dw + idx, acc[kx][co]);
        }
    } else if (grp == 30 && db) {
      for (int co = 0; co < 3; ++co) atomicAdd(db + co, acc[0][co]);
    }
  }
}

// one thread per input pixel: d_in[ci] for every input channel (written to din0 / din1; either may be NULL)
__global__ void __launch_bounds__(128)
small_conv_dgrad_kernel(const SCArgs a, float* __restrict__ din0, float* __restrict__ din1) {
  __shared__ float s_w[MAXC * MAXC * 25];
  const int cin = a.c0 + a.c1, kk = a.k * a.k, pad = a.k / 2;
  for (int i = threadIdx.x; i < a.c_out * cin * kk; i += blockDim.x) s_w[i] = a.wt[i];
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, b = blockIdx.z;
  if (x >= a.w) return;
  float acc[MAXC];
#pragma unroll
  for (int i = 0; i < MAXC; ++i) acc[i] = 0.0f;
  for (int ky = 0; ky < a.k; ++ky)
    for (int kx = 0; kx < a.k; ++kx) {
      int oy, ox;
      if (a.transposed) { oy = y - pad + ky; ox = x - pad + kx; }
      else {
        const int ty = y + pad - ky, tx = x + pad - kx;
        if (ty < 0 || tx < 0 || (ty % a.stride) || (tx % a.stride)) continue;
        oy = ty / a.stride; ox = tx / a.stride;
      }
      for (int co = 0; co < a.c_out; ++co) {
        const float gv = sc_g(a, b, co, oy, ox);
        if (gv == 0.0f) continue;
        for (int ci = 0; ci < cin; ++ci) {
          const float wv = a.transposed ? s_w[((ci * a.c_out + co) * a.k + ky) * a.k + kx]
                                        : s_w[((co * cin + ci) * a.k + ky) * a.k + kx];
          acc[ci] += gv * wv;
        }
      }
    }
  for (int ci = 0; ci < cin; ++ci) {
    if (ci < a.c0) { if (din0) din0[((long)(b * a.c0 + ci) * a.h + y) * a.w + x] = acc[ci]; }
    else if (din1) din1[((long)(b * a.c1 + ci - a.c0) * a.h + y) * a.w + x] = acc[ci];
  }
}

// Data gradient of the two 6 <-> 3 channel 5x5 stride-1 layers (pre_conv, after_conv) as a shared-memory tiled
// correlation of the 3-channel output gradient with K[ci][co][ky][kx] (a conv: W[co][ci] flipped; the transposed layer:
// W[ci][co] as stored): the tiling of conv5x5_6to3_kernel (image.cu) with the channel roles swapped.  A block of 16 x 8
// threads produces 128 x 8 pixels x 6 channels from a (128+4) x (8+4) x 3 patch; a thread owns 2 x 4 pixels x 6 channels.
constexpr int DG_TW = 128, DG_TH = 8, DG_PW = DG_TW + 8, DG_PH = DG_TH + 4;
__global__ void __launch_bounds__(128)
conv5x5_3to6_dgrad_kernel(const float* __restrict__ g, int h, int w, const float* __restrict__ wt, int transposed,
                          int c0, float* __restrict__ din0, float* __restrict__ din1) {
  __shared__ __align__(16) float s_g[3][DG_PH][DG_PW];
  __shared__ __align__(16) float s_w[3 * 5 * 32];       // [co][ky][kx * 6 + ci], 30 used of 32
  const int tid = threadIdx.y * 16 + threadIdx.x;
  for (int i = tid; i < 3 * 5 * 32; i += 128) {
    const int j = i & 31, ky = (i >> 5) % 5, co = i / 160;
    float v = 0.0f;
    if (j < 30) {
      const int kx = j / 6, ci = j % 6;
      v = transposed ? wt[((ci * 3 + co) * 5 + ky) * 5 + kx] : wt[((co * 6 + ci) * 5 + (4 - ky)) * 5 + (4 - kx)];
    }
    s_w[i] = v;
  }
  const int b = blockIdx.z, y0 = blockIdx.y * DG_TH, x0 = blockIdx.x * DG_TW;
  for (int i = tid; i < 3 * DG_PH * DG_PW; i += 128) {
    const int c = i % DG_PW, py = (i / DG_PW) % DG_PH, co = i / (DG_PW * DG_PH);
    const int gy = y0 + py - 2, gx = x0 + c - 2;
    float v = 0.0f;
    if (gy >= 0 && gy < h && gx >= 0 && gx < w) v = __ldg(g + ((long)(b * 3 + co) * h + gy) * w + gx);
    s_g[co][py][c] = v;
  }
  __syncthreads();
  float acc[2][4][6];
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int ci = 0; ci < 6; ++ci) acc[q][p][ci] = 0.0f;
  const int lx = threadIdx.x * 4, ly = threadIdx.y;
#pragma unroll 1
  for (int co = 0; co < 3; ++co) {
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) {
      const float4* wq = reinterpret_cast<const float4*>(&s_w[(co * 5 + ky) * 32]);
      float ww[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float4 t = wq[i]; ww[4 * i] = t.x; ww[4 * i + 1] = t.y; ww[4 * i + 2] = t.z; ww[4 * i + 3] = t.w; }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float4 a0 = *reinterpret_cast<const float4*>(&s_g[co][ly + ky][lx + 64 * q]);
        const float4 a1 = *reinterpret_cast<const float4*>(&s_g[co][ly + ky][lx + 64 * q + 4]);
        const float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int kx = 0; kx < 5; ++kx)
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float x = v[p + kx];
#pragma unroll
            for (int ci = 0; ci < 6; ++ci) acc[q][p][ci] = fmaf(x, ww[kx * 6 + ci], acc[q][p][ci]);
          }
      }
    }
  }
  const int oy = y0 + ly;
  if (oy >= h) return;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int ox = x0 + lx + 64 * q;
#pragma unroll
    for (int ci = 0; ci < 6; ++ci) {
      float* dst = ci < c0 ? (din0 ? din0 + ((long)(b * c0 + ci) * h + oy) * w : nullptr)
                           : (din1 ? din1 + ((long)(b * (6 - c0) + ci - c0) * h + oy) * w : nullptr);
      if (!dst) continue;
      if (ox + 3 < w && (w & 3) == 0) {
        *reinterpret_cast<float4*>(dst + ox) = make_float4(acc[q][0][ci], acc[q][1][ci], acc[q][2][ci], acc[q][3][ci]);
      } else {
        for (int p = 0; p < 4; ++p) if (ox + p < w) dst[ox + p] = acc[q][p][ci];
      }
    }
  }
}

// ---------------------------------------------------------------- GDN over <= 8 channels, backward (NCHW fp32)
// dx written; dbeta' [c] and dgamma' [c][c] accumulated (atomics) — chain to the stored parameters with masic_reparam_bwd
__global__ void __launch_bounds__(256)
gdn_small_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, int n, int c, long hw,
                     const float* __restrict__ beta, const float* __restrict__ gamma, float beta_bound, float gamma_bound,
                     float pedestal, int inverse, float* __restrict__ dx, float* __restrict__ dbeta, float* __restrict__ dgamma) {
  __shared__ float s_b[MAXC], s_g[MAXC * MAXC];
  __shared__ float red[8][MAXC + MAXC * MAXC];
  if (threadIdx.x < c) { const float v = fmaxf(beta[threadIdx.x], beta_bound); s_b[threadIdx.x] = v * v - pedestal; }
  if (threadIdx.x < c * c) { const float v = fmaxf(gamma[threadIdx.x], gamma_bound); s_g[threadIdx.x] = v * v - pedestal; }
  __syncthreads();
  const long total = (long)n * hw;
  float ab[MAXC], ag[MAXC * MAXC];
  for (int i = 0; i < MAXC; ++i) ab[i] = 0.0f;
  for (int i = 0; i < MAXC * MAXC; ++i) ag[i] = 0.0f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int b = (int)(i / hw);
    const long p = i - b * hw;
    float xv[MAXC], x2[MAXC], t[MAXC], gv[MAXC], nr[MAXC];
    for (int j = 0; j < c; ++j) {
      xv[j] = x[((long)b * c + j) * hw + p]; x2[j] = xv[j] * xv[j];
      gv[j] = g[((long)b * c + j) * hw + p];
    }
    for (int k = 0; k < c; ++k) {
      float nn = s_b[k];
      for (int j = 0; j < c; ++j) nn += s_g[k * c + j] * x2[j];
      nr[k] = nn;
      const float r = rsqrtf(nn);
      t[k] = inverse ? 0.5f * gv[k] * xv[k] * r : -0.5f * gv[k] * xv[k] * r * r * r;
      ab[k] += t[k];
      for (int j = 0; j < c; ++j) ag[k * c + j] += t[k] * x2[j];
    }
    for (int j = 0; j < c; ++j) {
      float v = 0.0f;
      for (int k = 0; k < c; ++k) v += s_g[k * c + j] * t[k];
      const float u = inverse ? gv[j] * sqrtf(nr[j]) : gv[j] * rsqrtf(nr[j]);
      dx[((long)b * c + j) * hw + p] = u + 2.0f * xv[j] * v;
    }
  }
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  for (int i = 0; i < c + c * c; ++i) {
    float a = i < c ? ab[i] : ag[i - c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) red[wp][i] = a;
  }
  __syncthreads();
  if (threadIdx.x < c + c * c) {
    float s = 0.0f;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    if (threadIdx.x < c) atomicAdd(dbeta + threadIdx.x, s);
    else atomicAdd(dgamma + threadIdx.x - c, s);
  }
}

// softmax over c channels, backward: w NHWC [P][c], dw NHWC [P][c] -> dlogits NCHW [n][c][hw]
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const float* __restrict__ w, const float* __restrict__ dw, int n, int c, long hw,
                   float* __restrict__ dl) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)n * hw) return;
  const int b = (int)(i / hw);
  const long p = i - b * hw;
  float dot = 0.0f;
  for (int j = 0; j < c; ++j) dot += w[i * c + j] * dw[i * c + j];
  for (int j = 0; j < c; ++j) dl[((long)b * c + j) * hw + p] = w[i * c + j] * (dw[i * c + j] - dot);
}

// dW[cl][ci][ky][kx] += sum_{n,p} LO[n,p,cl] * HI[n][ci][2p + k - 2]   (k = 5, stride 2, c_hi <= 3)
// LO NHWC bf16 [n][h][w][lo_pitch] (cl < c_lo <= 128 * gridDim.y), HI NCHW fp32.  block (128): one thread per cl with
// 75 accumulators; 32 pixels per step: their 5x5xc_hi patches are staged in shared memory (one barrier per 32
// pixels) and read back as broadcast float4.
constexpr int WS_PIX = 32, WS_NP = 76;          // 75 patch values padded to a multiple of 4
__global__ void __launch_bounds__(128, 4)
wgrad_small_kernel(const __nv_bfloat16* __restrict__ lo, int lo_pitch, int c_lo, const float* __restrict__ hi, int c_hi,
                   int n, int h, int w, float* __restrict__ dw) {
  __shared__ __align__(16) float patch[WS_PIX][WS_NP];
  __shared__ float lo_s[WS_PIX][128];
  const int cl = blockIdx.y * 128 + threadIdx.x;
  const long npix = (long)n * h * w;
  const long per = ((npix + gridDim.x - 1) / gridDim.x + WS_PIX - 1) / WS_PIX * WS_PIX;
  const long p0 = per * blockIdx.x, p1 = min(npix, p0 + per);
  float acc[WS_NP];
#pragma unroll
  for (int i = 0; i < WS_NP; ++i) acc[i] = 0.0f;
  const int np = c_hi * 25;
  const int hh = 2 * h, ww = 2 * w;
  for (long pb = p0; pb < p1; pb += WS_PIX) {
    __syncthreads();
    {   // patch staging: thread -> (pixel tid/4, values tid%4, +4, ...): pixel coordinates computed once, 32-bit math
      const int pl = threadIdx.x >> 2, sub = threadIdx.x & 3;
      const int p = (int)(pb - p0) + pl;                   // pixel index relative to this block's slab
      const long pg = p0 + p;
      const bool live = pg < p1;
      const int pgi = (int)pg;                             // n*h*w < 2^31 (checked on the host)
      const int ox = pgi % w, r = pgi / w;
      const int oy = r % h, b = r / h;
      const float* hb = hi + (long)b * c_hi * hh * ww;
#pragma unroll
      for (int i = sub; i < WS_NP; i += 4) {
        float v = 0.0f;
        if (live && i < np) {
          const int ci = i / 25, tap = i - ci * 25;
          const int iy = 2 * oy + tap / 5 - 2, ix = 2 * ox + tap % 5 - 2;
          if (iy >= 0 && iy < hh && ix >= 0 && ix < ww) v = hb[((long)ci * hh + iy) * ww + ix];
        }
        patch[pl][i] = v;
      }
    }
    // LO tile: groups of 8 independent coalesced loads per thread (one dependent load per pixel would expose latency)
#pragma unroll
    for (int g8 = 0; g8 < WS_PIX; g8 += 8) {
      float t[8];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        t[q] = (cl < c_lo && pb + g8 + q < p1) ? __bfloat162float(lo[(pb + g8 + q) * lo_pitch + cl]) : 0.0f;
#pragma unroll
      for (int q = 0; q < 8; ++q) lo_s[g8 + q][threadIdx.x] = t[q];
    }
    __syncthreads();
    if (cl < c_lo) {
      const int cnt = (int)min((long)WS_PIX, p1 - pb);
      for (int pl = 0; pl < cnt; ++pl) {
        const float lv = lo_s[pl][threadIdx.x];
        const float4* pp = reinterpret_cast<const float4*>(patch[pl]);
#pragma unroll
        for (int q = 0; q < WS_NP / 4; ++q) {
          const float4 v = pp[q];
          acc[4 * q] += lv * v.x; acc[4 * q + 1] += lv * v.y; acc[4 * q + 2] += lv * v.z; acc[4 * q + 3] += lv * v.w;
        }
      }
    }
  }
  if (cl < c_lo)
#pragma unroll
    for (int i = 0; i < WS_NP; ++i)
      if (i < np) atomicAdd(dw + (long)cl * np + i, acc[i]);     // [cl][ci][tap], Conv2d and ConvTranspose2d alike
}

// out[c] += sum over n, hw of g[n][c][hw]
__global__ void __launch_bounds__(256)
colsum_nchw_kernel(const float* __restrict__ g, int n, int c, long hw, float* __restrict__ out) {
  __shared__ float red[8];
  const int ch = blockIdx.x;
  float s = 0.0f;
  for (int b = 0; b < n; ++b)
    for (long p = blockIdx.y * (long)blockDim.x + threadIdx.x; p < hw; p += (long)gridDim.y * blockDim.x)
      s += g[((long)b * c + ch) * hw + p];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(out + ch, t);
  }
}

}  // namespace

#define S(stream) static_cast<cudaStream_t>(stream)

extern "C" int masic_mse_grad(const float* x_hat, const float* x, const float* addend, float scale, int64_t numel,
                              float* g, void* stream) {
  if (!x_hat || !x || !g || numel <= 0) return MASIC_EINVAL;
  mse_grad_kernel<<<(unsigned)((numel + 255) / 256), 256, 0, S(stream)>>>(x_hat, x, addend, scale, numel, g);
  return (int)cudaGetLastError();
}

extern "C" int masic_warp_perspective_bwd(const float* g0, const float* g1, int n, int c, int h, int w, int h_out,
                                          int w_out, const double* t_prepared, float* dsrc_zeroed, void* stream) {
  if (!g0 || !t_prepared || !dsrc_zeroed || n <= 0 || c <= 0 || c > 8 || h_out < 2 || w_out < 2) return MASIC_EINVAL;
  dim3 grid((w_out + 255) / 256, h_out, n);
  warp_bwd_kernel<<<grid, 256, 0, S(stream)>>>(g0, g1, n, c, h, w, h_out, w_out, t_prepared, dsrc_zeroed);
  return (int)cudaGetLastError();
}

extern "C" int masic_conv_small_bwd(const float* in0, int c0, const float* in1, int c1, int n, int h, int w,
                                    const float* weight, int transposed_s1, int c_out, int ksize, int stride,
                                    const float* g_out, const float* act_out, float* din0, float* din1,
                                    float* dweight_zeroed, float* dbias_zeroed, void* stream) {
  if (!in0 || !weight || !g_out || c_out <= 0 || c_out > MAXC || c0 <= 0 || c0 + c1 > MAXC) return MASIC_EINVAL;
  if ((ksize != 3 && ksize != 5) || (stride != 1 && stride != 2) || (transposed_s1 && stride != 1)) return MASIC_EINVAL;
  SCArgs a;
  a.in0 = in0; a.in1 = in1; a.c0 = c0; a.c1 = c1; a.n = n; a.h = h; a.w = w;
  a.ho = (h + stride - 1) / stride; a.wo = (w + stride - 1) / stride;
  a.c_out = c_out; a.k = ksize; a.stride = stride; a.transposed = transposed_s1;
  a.g = g_out; a.act_out = act_out; a.wt = weight;
  if (dweight_zeroed && ksize == 5 && stride == 1 && c0 + c1 == 6 && c_out == 3 && !act_out && c1 > 0) {
    const int n_tiles = ((w + WG_TW - 1) / WG_TW) * ((h + WG_TH - 1) / WG_TH) * n;
    const int blocks = n_tiles < 592 ? n_tiles : 592;
    conv5x5_6x3_wgrad_kernel<<<blocks, 256, 0, S(stream)>>>(in0, in1, c0, g_out, n, h, w, transposed_s1, dweight_zeroed,
                                                            dbias_zeroed);
  } else if (dweight_zeroed) {
    const int nw = c_out * (c0 + c1) * ksize * ksize;
    if (nw + c_out > SW_MAXW * 256) return MASIC_ENOSUP;
    const int span = transposed_s1 ? SW_T + ksize - 1 : (SW_T - 1) * stride + ksize;
    const int smem = (c_out * SW_T * SW_T + (c0 + c1) * span * span) * (int)sizeof(float);
    const int n_tiles = ((a.wo + SW_T - 1) / SW_T) * ((a.ho + SW_T - 1) / SW_T) * n;
    int blocks = n_tiles < 592 ? n_tiles : 592;
    small_conv_wgrad_kernel<<<blocks, 256, smem, S(stream)>>>(a, dweight_zeroed, dbias_zeroed);
  }
  if (din0 || din1) {
    if (ksize == 5 && stride == 1 && c0 + c1 == 6 && c_out == 3 && !act_out) {
      dim3 grid((w + DG_TW - 1) / DG_TW, (h + DG_TH - 1) / DG_TH, n), block(16, 8);
      conv5x5_3to6_dgrad_kernel<<<grid, block, 0, S(stream)>>>(g_out, h, w, weight, transposed_s1, c0, din0, din1);
    } else {
      dim3 grid((w + 127) / 128, h, n);
      small_conv_dgrad_kernel<<<grid, 128, 0, S(stream)>>>(a, din0, din1);
    }
  }
  return (int)cudaGetLastError();
}

extern "C" int masic_gdn_small_bwd(const float* x, const float* g, int n, int c, int hw, const float* beta,
                                   const float* gamma, float beta_min, int inverse, float* dx,
                                   float* dbeta_prime_zeroed, float* dgamma_prime_zeroed, void* stream) {
  if (!x || !g || !beta || !gamma || !dx || !dbeta_prime_zeroed || !dgamma_prime_zeroed || c <= 0 || c > MAXC)
    return MASIC_EINVAL;
  const long total = (long)n * hw;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  gdn_small_bwd_kernel<<<blocks, 256, 0, S(stream)>>>(x, g, n, c, hw, beta, gamma, sqrtf(beta_min + 1.4551915228366852e-11f),
                                                      3.814697265625e-06f, 1.4551915228366852e-11f, inverse, dx,
                                                      dbeta_prime_zeroed, dgamma_prime_zeroed);
  return (int)cudaGetLastError();
}

extern "C" int masic_softmax_channels_bwd(const float* w_nhwc, const float* dw_nhwc, int n, int c, int hw,
                                          float* dlogits_nchw, void* stream) {
  if (!w_nhwc || !dw_nhwc || !dlogits_nchw || c <= 0 || c > MAXC) return MASIC_EINVAL;
  const long total = (long)n * hw;
  softmax_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, S(stream)>>>(w_nhwc, dw_nhwc, n, c, hw, dlogits_nchw);
  return (int)cudaGetLastError();
}

extern "C" int masic_wgrad_small(const void* lo_bf16, int lo_pitch, int c_lo, const float* hi_nchw, int c_hi, int n,
                                 int h_lo, int w_lo, float* dw_accum, void* stream) {
  if (!lo_bf16 || !hi_nchw || !dw_accum || c_hi <= 0 || c_hi > 3 || c_lo <= 0 || n <= 0) return MASIC_EINVAL;
  const long npix = (long)n * h_lo * w_lo;
  if (npix >= (1L << 31)) return MASIC_EINVAL;
  int gx = (int)((npix + 127) / 128);
  if (gx > 1184) gx = 1184;                      // 8 resident blocks x 148 SMs
  wgrad_small_kernel<<<dim3(gx, (c_lo + 127) / 128), 128, 0, S(stream)>>>(
      static_cast<const __nv_bfloat16*>(lo_bf16), lo_pitch, c_lo, hi_nchw, c_hi, n, h_lo, w_lo, dw_accum);
  return (int)cudaGetLastError();
}

extern "C" int masic_colsum_nchw(const float* g, int n, int c, int64_t hw, float* out_accum, void* stream) {
  if (!g || !out_accum || n <= 0 || c <= 0 || hw <= 0) return MASIC_EINVAL;
  int gy = (int)((hw + 4095) / 4096);
  if (gy > 128) gy = 128;
  colsum_nchw_kernel<<<dim3(c, gy), 256, 0, S(stream)>>>(g, n, c, hw, out_accum);
  return (int)cudaGetLastError();
}
