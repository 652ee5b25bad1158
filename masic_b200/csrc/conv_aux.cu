// conv_aux.cu — weight packing for conv_tc.cu, the GDN re-parametrisation, and a plain
// CUDA-core direct convolution used (a) as the on-device cross-check of the tensor-core
// path in tests and (b) by callers that want fp32 torch-layout weights without packing.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/masic_b200.h"
#include "cvt16.cuh"

namespace {

constexpr int KBLK = 64;

__device__ __forceinline__ float pack_weight_value(const float* __restrict__ w, int kind, int transposed, int k,
                                                   int c_in, int c_out, int c_out_pad, int ncb, long i) {
  const int c = (int)(i % KBLK);
  long r = i / KBLK;
  const int co = (int)(r % c_out_pad);
  const int kb = (int)(r / c_out_pad);
  const int tap = kb / ncb, cb = kb % ncb;
  const int ci = cb * KBLK + c;
  float v = 0.0f;
  if (ci < c_in || kind == MASIC_CONV_XFOLD4 || kind == MASIC_CONV_XFOLD8) {
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
    } else if (kind == MASIC_CONV_XFOLD8) {
      // k-block = ky; column = pixel * 8 + channel of the 8-pixel window that starts at input pixel 2*ox - 2, so
      // pixel j is tap kx = j (j < 5); `c_in` is the REAL channel count of w (<= 8)
      const int ky = kb, kx = c >> 3, ch = c & 7;
      // channels c_in .. 2*c_in-1 repeat the weights: a producer may store the pixel as [hi | lo] (MASIC_FMT_SPLIT),
      // and hi * w + lo * w = x * w; without the split those slots hold zeros
      const int chr = (ch >= c_in && 2 * c_in <= 8) ? ch - c_in : ch;
      if (co < c_out && chr < c_in && kx < 5)
        v = w[((static_cast<long>(co) * c_in + chr) * 5 + ky) * 5 + kx];
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
  return v;
}

__global__ void pack_weights_kernel(const float* __restrict__ w, int kind, int transposed, int k,
                                    int c_in, int c_out, int c_out_pad, int ncb, long total,
                                    uint16_t* __restrict__ dst, int f16) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  dst[i] = masic::pack16(pack_weight_value(w, kind, transposed, k, c_in, c_out, c_out_pad, ncb, i), f16);
}

// Every weight pack (and padded bias copy) of a training step in ONE launch: block b belongs to the job j with
// first_block[j] <= b < first_block[j+1] (binary search), the tail blocks of a job copy its bias.
struct PackJobDev {
  const float* w; __nv_bfloat16* dst; const float* bias_src; float* bias_dst;
  int kind, transposed, k, c_in, c_out, c_out_pad, ncb, bias_n, bias_rep, pad;
  long total;           // packed elements
  int w_blocks;         // blocks of 256 covering `total`; the job's remaining blocks cover the bias
  int pad2;
};

// Plain conv / transposed-conv packs (MASIC_CONV, MASIC_DECONV_S2) go through a shared-memory tile: a block owns one
// 64-channel k-block column cb and PT_CO output channels for ALL taps, reads the fp32 weights in the order they lie in
// memory (runs of 64 * k * k floats of one output channel, or PT_CO * k * k floats of one input channel for the
// ConvTranspose2d layout) and writes 128-byte rows of the packed layout.  The element-wise form below reads with a
// stride of k * k floats between neighbouring threads (one 32-byte sector per value, re-fetched for every tap) and
// pays several 64-bit divisions per element: 0.6 ms for the training step's 110 packs; this form is bound by the
// 140 MB it reads.
constexpr int PT_CO = 4;
__device__ __forceinline__ bool pack_job_tiled(const PackJobDev& j) {
  return (j.kind == MASIC_CONV || j.kind == MASIC_DECONV_S2) && (j.c_out_pad % PT_CO) == 0 && j.k * j.k <= 25;
}

__global__ void __launch_bounds__(256)
pack_batch_kernel(const PackJobDev* __restrict__ jobs, const int* __restrict__ first_block, int n_jobs) {
  __shared__ float tile[PT_CO * KBLK * 25 + 1];
  int lo = 0, hi = n_jobs - 1;
  const int b = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (first_block[mid] <= b) lo = mid; else hi = mid - 1;
  }
  const PackJobDev j = jobs[lo];
  const int lb = b - first_block[lo];
  if (lb >= j.w_blocks) {
    const int i = (lb - j.w_blocks) * 256 + threadIdx.x;      // bias_dst[i] = bias_src[i % bias_n], i < bias_n * bias_rep
    if (i < j.bias_n * j.bias_rep) j.bias_dst[i] = j.bias_src[i % j.bias_n];
    return;
  }
  if (!pack_job_tiled(j)) {
    const long i = (long)lb * 256 + threadIdx.x;
    if (i < j.total)
      j.dst[i] = __float2bfloat16_rn(pack_weight_value(j.w, j.kind, j.transposed, j.k, j.c_in, j.c_out, j.c_out_pad, j.ncb, i));
    return;
  }
  const int kk = j.k * j.k;
  const int n_cog = j.c_out_pad / PT_CO;
  const int cb = lb / n_cog, co0 = (lb % n_cog) * PT_CO, ci0 = cb * KBLK;
  const bool flip = j.transposed && j.kind == MASIC_CONV;      // stride-1 transposed = flipped conv
  // tile[co_l][ci_l * kk + t], t = packed tap index
  if (!j.transposed) {
    const int run = KBLK * kk;                                  // (ci_l, tap) of one output channel: contiguous in w
    for (int e = threadIdx.x; e < PT_CO * run; e += 256) {
      const int co_l = e / run, r = e - co_l * run;
      const int ci_l = r / kk;
      const int co = co0 + co_l, ci = ci0 + ci_l;
      float v = 0.0f;
      if (co < j.c_out && ci < j.c_in) v = __ldg(j.w + ((long)co * j.c_in + ci0) * kk + r);
      tile[co_l * run + r] = v;
    }
  } else {
    const int run = PT_CO * kk;                                 // (co_l, tap) of one input channel: contiguous in w
    for (int e = threadIdx.x; e < KBLK * run; e += 256) {
      const int ci_l = e / run, r = e - ci_l * run;
      const int co_l = r / kk, ts = r - co_l * kk;
      const int co = co0 + co_l, ci = ci0 + ci_l;
      float v = 0.0f;
      if (co < j.c_out && ci < j.c_in) v = __ldg(j.w + ((long)ci * j.c_out + co) * kk + ts);
      tile[co_l * (KBLK * kk) + ci_l * kk + (flip ? kk - 1 - ts : ts)] = v;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < kk * PT_CO * KBLK; e += 256) {
    const int c = e & (KBLK - 1), co_l = (e >> 6) & (PT_CO - 1), t = e >> 8;
    j.dst[(((long)t * j.ncb + cb) * j.c_out_pad + co0 + co_l) * KBLK + c] =
        __float2bfloat16_rn(tile[co_l * (KBLK * kk) + c * kk + t]);
  }
}

__global__ void gdn_prepare_kernel(const float* __restrict__ beta, const float* __restrict__ gamma,
                                   int c, float beta_bound, float gamma_bound, float pedestal,
                                   float* __restrict__ beta_out, float* __restrict__ gamma_f32,
                                   uint16_t* __restrict__ gamma_bf16, int f16) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c) {
    const float b = fmaxf(beta[i], beta_bound);
    beta_out[i] = b * b - pedestal;
  }
  if (i < c * c) {
    const float g = fmaxf(gamma[i], gamma_bound);
    const float gp = g * g - pedestal;
    if (gamma_f32) gamma_f32[i] = gp;
    if (gamma_bf16) gamma_bf16[i] = masic::pack16(gp, f16);
  }
}

__global__ void conv_direct_kernel(const uint16_t* __restrict__ in, int f16, int n, int h_in, int w_in,
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
      const uint16_t* ip = in + ((static_cast<long>(ni) * h_in + iy) * w_in + ix) * in_cpitch + in_coff;
      for (int ci = 0; ci < c_in; ++ci) {
        float wv = transposed ? w[((static_cast<long>(ci) * c_out + co) * k + ky) * k + kx]
                              : w[((static_cast<long>(co) * c_in + ci) * k + ky) * k + kx];
        if (round_w) wv = masic::unpack16(masic::pack16(wv, f16), f16);
        acc = fmaf(masic::unpack16(ip[ci], f16), wv, acc);
      }
    }
  out[((static_cast<long>(ni) * h_out + oy) * w_out + ox) * out_cpitch + out_coff + co] = acc;
}

}  // namespace

extern "C" int64_t masic_packed_weight_bytes(int kind, int ksize, int c_in, int c_out_pad) {
  if (kind == MASIC_CONV_XFOLD4) return static_cast<int64_t>(10) * c_out_pad * KBLK * 2;   // (ky, group) blocks
  if (kind == MASIC_CONV_XFOLD8) return static_cast<int64_t>(5) * c_out_pad * KBLK * 2;    // one block per ky
  const int taps = (kind == MASIC_DECONV_S2_SUBPIX) ? 9 : ksize * ksize;
  const int ncb = (c_in + KBLK - 1) / KBLK;
  return static_cast<int64_t>(taps) * ncb * c_out_pad * KBLK * 2;
}

extern "C" int masic_pack_conv_weights(const float* w, int kind, int transposed, int ksize, int c_in,
                                       int c_out, int c_out_pad, void* dst, int f16, void* stream) {
  if (!w || !dst || c_in <= 0 || c_out <= 0) return MASIC_EINVAL;
  if (kind == MASIC_DECONV_S2_SUBPIX) {
    if (ksize != 5 || 4 * c_out > c_out_pad) return MASIC_EINVAL;
  } else if (c_out > c_out_pad) {
    return MASIC_EINVAL;
  }
  if (kind == MASIC_CONV_XFOLD4 && (ksize != 5 || c_in > 16 || transposed)) return MASIC_EINVAL;
  if (kind == MASIC_CONV_XFOLD8 && (ksize != 5 || c_in > 8 || transposed)) return MASIC_EINVAL;
  const int ncb = (kind == MASIC_CONV_XFOLD4 || kind == MASIC_CONV_XFOLD8) ? 1 : (c_in + KBLK - 1) / KBLK;
  const long total = masic_packed_weight_bytes(kind, ksize, c_in, c_out_pad) / 2;
  const int bs = 256;
  pack_weights_kernel<<<(unsigned)((total + bs - 1) / bs), bs, 0, static_cast<cudaStream_t>(stream)>>>(
      w, kind, transposed, ksize, c_in, c_out, c_out_pad, ncb, total, static_cast<uint16_t*>(dst), f16);
  return (int)cudaGetLastError();
}

extern "C" int masic_gdn_prepare(const float* beta, const float* gamma, int c, float beta_min,
                                 float* beta_out, float* gamma_out_f32, void* gamma_out_bf16,
                                 int f16, void* stream) {
  if (!beta || !gamma || !beta_out || c <= 0) return MASIC_EINVAL;
  // compressai/ops/parametrizers.py:49-64 — pedestal = (2^-18)^2, bound = sqrt(minimum + pedestal)
  const float pedestal = 1.4551915228366852e-11f;          // 2^-36
  const float beta_bound = sqrtf(beta_min + pedestal);
  const float gamma_bound = 3.814697265625e-06f;           // 2^-18
  const int bs = 256, total = c * c;
  gdn_prepare_kernel<<<(total + bs - 1) / bs, bs, 0, static_cast<cudaStream_t>(stream)>>>(
      beta, gamma, c, beta_bound, gamma_bound, pedestal, beta_out, gamma_out_f32,
      static_cast<uint16_t*>(gamma_out_bf16), f16);
  return (int)cudaGetLastError();
}

extern "C" int masic_conv_direct_nhwc(const void* in, int n, int h_in, int w_in, int in_cpitch,
                                      int in_coff, int c_in, const float* w, int transposed, int ksize,
                                      int stride, uint32_t tap_mask, const float* bias, int c_out,
                                      float* out_f32, int out_cpitch, int out_coff, int round_w_bf16,
                                      int f16, void* stream) {
  if (!in || !w || !out_f32 || (stride != 1 && stride != 2)) return MASIC_EINVAL;
  int h_out, w_out;
  if (!transposed) { h_out = (h_in + stride - 1) / stride; w_out = (w_in + stride - 1) / stride; }
  else { h_out = h_in * stride; w_out = w_in * stride; }
  const long total = (long)n * h_out * w_out * c_out;
  const int bs = 128;
  conv_direct_kernel<<<(unsigned)((total + bs - 1) / bs), bs, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint16_t*>(in), f16, n, h_in, w_in, in_cpitch, in_coff, c_in, w, transposed,
      ksize, stride, tap_mask, bias, c_out, h_out, w_out, out_f32, out_cpitch, out_coff, round_w_bf16);
  return (int)cudaGetLastError();
}

struct MasicPackBatch {
  PackJobDev* d_jobs = nullptr;
  int* d_first = nullptr;
  int n_jobs = 0, n_blocks = 0;
};

extern "C" int masic_pack_batch_create(const MasicPackJob* jobs, int n_jobs, MasicPackBatch** out) {
  if (!jobs || !out || n_jobs <= 0) return MASIC_EINVAL;
  std::vector<PackJobDev> hj(n_jobs);
  std::vector<int> first(n_jobs + 1, 0);
  for (int i = 0; i < n_jobs; ++i) {
    const MasicPackJob& s = jobs[i];
    if (!s.w || !s.dst || s.c_in <= 0 || s.c_out <= 0) return MASIC_EINVAL;
    if (s.kind == MASIC_DECONV_S2_SUBPIX ? (s.ksize != 5 || 4 * s.c_out > s.c_out_pad) : (s.c_out > s.c_out_pad)) return MASIC_EINVAL;
    if (s.kind == MASIC_CONV_XFOLD4 && (s.ksize != 5 || s.c_in > 16 || s.transposed)) return MASIC_EINVAL;
    if (s.kind == MASIC_CONV_XFOLD8 && (s.ksize != 5 || s.c_in > 8 || s.transposed)) return MASIC_EINVAL;
    if ((s.bias_src == nullptr) != (s.bias_dst == nullptr)) return MASIC_EINVAL;
    PackJobDev& d = hj[i];
    d.w = s.w; d.dst = static_cast<__nv_bfloat16*>(s.dst); d.bias_src = s.bias_src; d.bias_dst = s.bias_dst;
    d.kind = s.kind; d.transposed = s.transposed; d.k = s.ksize; d.c_in = s.c_in; d.c_out = s.c_out; d.c_out_pad = s.c_out_pad;
    d.ncb = (s.kind == MASIC_CONV_XFOLD4 || s.kind == MASIC_CONV_XFOLD8) ? 1 : (s.c_in + KBLK - 1) / KBLK;
    d.bias_n = s.bias_src ? s.c_out : 0;
    d.bias_rep = (s.kind == MASIC_DECONV_S2_SUBPIX) ? 4 : 1;
    d.total = masic_packed_weight_bytes(s.kind, s.ksize, s.c_in, s.c_out_pad) / 2;
    d.w_blocks = (int)((d.total + 255) / 256);
    if ((d.kind == MASIC_CONV || d.kind == MASIC_DECONV_S2) && d.c_out_pad % PT_CO == 0 && d.k * d.k <= 25)
      d.w_blocks = d.ncb * (d.c_out_pad / PT_CO);             // tiled form: one block per (k-block column, PT_CO channels)
    first[i + 1] = first[i] + d.w_blocks + (d.bias_n * d.bias_rep + 255) / 256;
  }
  MasicPackBatch* pb = new MasicPackBatch();
  pb->n_jobs = n_jobs; pb->n_blocks = first[n_jobs];
  cudaError_t e = cudaMalloc(&pb->d_jobs, sizeof(PackJobDev) * n_jobs);
  if (e == cudaSuccess) e = cudaMalloc(&pb->d_first, sizeof(int) * (n_jobs + 1));
  if (e == cudaSuccess) e = cudaMemcpy(pb->d_jobs, hj.data(), sizeof(PackJobDev) * n_jobs, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(pb->d_first, first.data(), sizeof(int) * (n_jobs + 1), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(pb->d_jobs); cudaFree(pb->d_first); delete pb; return (int)e; }
  *out = pb;
  return MASIC_OK;
}

extern "C" int masic_pack_batch_launch(const MasicPackBatch* pb, void* stream) {
  if (!pb) return MASIC_EINVAL;
  pack_batch_kernel<<<pb->n_blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(pb->d_jobs, pb->d_first, pb->n_jobs);
  return (int)cudaGetLastError();
}

extern "C" void masic_pack_batch_destroy(MasicPackBatch* pb) {
  if (!pb) return;
  cudaFree(pb->d_jobs); cudaFree(pb->d_first);
  delete pb;
}
