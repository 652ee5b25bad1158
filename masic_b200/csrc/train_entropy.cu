// train_entropy.cu — training-mode entropy models: forward with additive uniform noise AND the local
// backward of the rate term in the same pass (sm_100a, memory-bound fused kernels).
//
// Replaces (reference, file:line), for model.train():
//   EntropyModel._quantize('noise')                  compressai/entropy_models/entropy_models.py:98-110
//   GaussianMixtureConditional_gf.forward/_likelihood entropy_models.py:808-858      (+ softmax over K, MASIC.py:389-393)
//   EntropyBottleneck.forward/_likelihood/_logits_cumulative  entropy_models.py:350-411
//   EntropyBottleneck.loss                           entropy_models.py:345-348
//   LowerBound's custom gradient                     compressai/ops/bound_ops.py:36-58
//   the bpp term of RateDistortionLoss and its autograd backward  coremasic/mywork/newtrain_codec_real.py:66-79,134
//
// The rate loss is  sum(log(lik)) * c,  c = -1 / (ln2 * N*H*W): dLoss/dlik = c / lik depends on this element only,
// so each kernel produces the likelihood AND the gradients w.r.t. its inputs/parameters at once (no second pass,
// no saved intermediates).  Since c < 0 the likelihood floor's LowerBound gradient always passes.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/masic_b200.h"

namespace {

constexpr float kLikBound = 1e-9f;
constexpr float kInvSqrt2Neg = -0.70710678118654752440f;
constexpr float kInvSqrt2Pi = 0.3989422804014327f;

__device__ __forceinline__ float phi_cdf(float x) { return 0.5f * erfcf(kInvSqrt2Neg * x); }
__device__ __forceinline__ float phi_pdf(float x) { return kInvSqrt2Pi * __expf(-0.5f * x * x); }

// ------------------------------------------------------------------ K-mixture likelihood, training
// All tensors NHWC: y/noise/dy/lik [P][M], sigma/mu/wl and their gradients [P][K*M] (k-major channels).
// One thread per (pixel, m); consecutive threads walk m, so all 3K parameter loads are coalesced.
template <int K>
__global__ void __launch_bounds__(256)
gmm_train_kernel(const float* __restrict__ y, const float* __restrict__ noise, const float* __restrict__ sigma,
                 const float* __restrict__ mu, const float* __restrict__ wl, long total, int M, float bound, float c,
                 float* __restrict__ lik, __nv_bfloat16* __restrict__ yhat_bf, int bf_pitch, float* __restrict__ yhat,
                 float* __restrict__ dy, __nv_bfloat16* __restrict__ dsig, __nv_bfloat16* __restrict__ dmu,
                 __nv_bfloat16* __restrict__ dwl) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long p = i / M;
  const int m = (int)(i - p * M);
  const float v = y[i] + noise[i];
  const long pb = p * (long)(K * M) + m;
  float w[K], Pk[K], dad[K], ds[K], sg[K];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < K; ++k) { w[k] = wl[pb + (long)k * M]; mx = fmaxf(mx, w[k]); }
  float sum = 0.0f;
#pragma unroll
  for (int k = 0; k < K; ++k) { w[k] = expf(w[k] - mx); sum += w[k]; }
  float l = 0.0f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    w[k] = w[k] / sum;
    const float sraw = sigma[pb + (long)k * M];
    const float s = fmaxf(sraw, bound);
    const float d = v - mu[pb + (long)k * M];
    const float ad = fabsf(d);
    const float a = (0.5f - ad) / s, b = (-0.5f - ad) / s;
    Pk[k] = phi_cdf(a) - phi_cdf(b);
    const float pa = phi_pdf(a), pbb = phi_pdf(b);
    dad[k] = (pbb - pa) / s;                // dP/d|d|
    ds[k] = (b * pbb - a * pa) / s;         // dP/ds
    sg[k] = d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f);
    l += Pk[k] * w[k];
    // the sigma branch ends in a ReLU (MASIC.py:344,416) whose derivative is folded in here: sigma is the
    // post-ReLU value (the LowerBound(0.11) pass-through rule is applied below, once the gradient's sign is known)
    ds[k] = (sraw > 0.0f) ? ds[k] : 0.0f;
  }
  const float lb = fmaxf(l, kLikBound);
  const float gl = c / lb;                  // dLoss/dlik (negative): passes the 1e-9 floor's LowerBound
  float dot = 0.0f, dv = 0.0f;
#pragma unroll
  for (int k = 0; k < K; ++k) dot += w[k] * (gl * Pk[k]);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float dP = gl * w[k];
    const float g_ad = dP * dad[k];
    dv += g_ad * sg[k];
    const long o = pb + (long)k * M;
    if (dmu) dmu[o] = __float2bfloat16_rn(-g_ad * sg[k]);
    if (dsig) {
      const float sraw = sigma[o];
      float g_s = dP * ds[k];
      if (!(sraw >= bound) && !(g_s < 0.0f)) g_s = 0.0f;
      dsig[o] = __float2bfloat16_rn(g_s);
    }
    if (dwl) dwl[o] = __float2bfloat16_rn(w[k] * (gl * Pk[k] - dot));
  }
  if (lik) lik[i] = lb;
  if (yhat) yhat[i] = v;
  if (yhat_bf) yhat_bf[p * bf_pitch + m] = __float2bfloat16_rn(v);
  if (dy) dy[i] = dv;
}

// ------------------------------------------------------------------ EntropyBottleneck, training
struct EBRaw {          // per-channel parameters: transformed values and the derivative of the transform
  float M0[3], M1[9], M2[9], M3[9], M4[3];
  float B0[3], B1[3], B2[3], B3[3], B4[1];
  float F0[3], F1[3], F2[3], F3[3];
};
constexpr int EB_NPAR = 58;     // 33 matrix + 13 bias + 12 factor entries per channel

__device__ __forceinline__ float softplus_f(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

struct EBTrace { float pre[4][3], th[4][3], h[4][3]; };

// logits = L(v) keeping what the backward needs
__device__ __forceinline__ float eb_fwd_trace(const EBRaw& c, float v, EBTrace& t) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float a = c.M0[i] * v + c.B0[i];
    t.pre[0][i] = a; t.th[0][i] = tanhf(a); t.h[0][i] = a + c.F0[i] * t.th[0][i];
  }
  const float* Ms[3] = {c.M1, c.M2, c.M3};
  const float* Bs[3] = {c.B1, c.B2, c.B3};
  const float* Fs[3] = {c.F1, c.F2, c.F3};
#pragma unroll
  for (int l = 0; l < 3; ++l) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float a = Ms[l][i * 3] * t.h[l][0] + Ms[l][i * 3 + 1] * t.h[l][1] + Ms[l][i * 3 + 2] * t.h[l][2] + Bs[l][i];
      t.pre[l + 1][i] = a; t.th[l + 1][i] = tanhf(a); t.h[l + 1][i] = a + Fs[l][i] * t.th[l + 1][i];
    }
  }
  return c.M4[0] * t.h[3][0] + c.M4[1] * t.h[3][1] + c.M4[2] * t.h[3][2] + c.B4[0];
}

// accumulate g * dL/dparam into acc[58] (layout: M0 3, M1 9, M2 9, M3 9, M4 3, B0 3, B1 3, B2 3, B3 3, B4 1,
// F0 3, F1 3, F2 3, F3 3) and return g * dL/dv
__device__ __forceinline__ float eb_bwd_trace(const EBRaw& c, float v, const EBTrace& t, float g, float* acc) {
  float dh[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { acc[30 + i] += g * t.h[3][i]; dh[i] = c.M4[i] * g; }
  acc[45] += g;
  const float* Ms[3] = {c.M1, c.M2, c.M3};
  const float* Fs[4] = {c.F0, c.F1, c.F2, c.F3};
#pragma unroll
  for (int l = 3; l >= 1; --l) {
    float dpre[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      acc[46 + 3 * l + i] += dh[i] * t.th[l][i];
      dpre[i] = dh[i] * (1.0f + Fs[l][i] * (1.0f - t.th[l][i] * t.th[l][i]));
      acc[33 + 3 * l + i] += dpre[i];
    }
    float nh[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        acc[3 + 9 * (l - 1) + i * 3 + j] += dpre[i] * t.h[l - 1][j];
        nh[j] += Ms[l - 1][i * 3 + j] * dpre[i];
      }
#pragma unroll
    for (int j = 0; j < 3; ++j) dh[j] = nh[j];
  }
  float dv = 0.0f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    acc[46 + i] += dh[i] * t.th[0][i];
    const float dpre = dh[i] * (1.0f + c.F0[i] * (1.0f - t.th[0][i] * t.th[0][i]));
    acc[33 + i] += dpre;
    acc[i] += dpre * v;
    dv += c.M0[i] * dpre;
  }
  return dv;
}

__device__ __forceinline__ void eb_load_chan(EBRaw& sc, float* dtr, int c, const float* m0, const float* m1,
                                             const float* m2, const float* m3, const float* m4, const float* b0,
                                             const float* b1, const float* b2, const float* b3, const float* b4,
                                             const float* f0, const float* f1, const float* f2, const float* f3) {
  // thread i < 58 loads parameter i: transformed value into sc, d(transform)/d(raw) into dtr[i]
  const int i = threadIdx.x;
  if (i >= EB_NPAR) return;
  float* dst = reinterpret_cast<float*>(&sc);
  float raw, val, der;
  if (i < 33) {
    raw = i < 3 ? m0[c * 3 + i] : i < 12 ? m1[c * 9 + i - 3] : i < 21 ? m2[c * 9 + i - 12] : i < 30 ? m3[c * 9 + i - 21]
                                                                                                  : m4[c * 3 + i - 30];
    val = softplus_f(raw); der = sigmoid_f(raw);
  } else if (i < 46) {
    const int j = i - 33;
    raw = j < 3 ? b0[c * 3 + j] : j < 6 ? b1[c * 3 + j - 3] : j < 9 ? b2[c * 3 + j - 6] : j < 12 ? b3[c * 3 + j - 9] : b4[c];
    val = raw; der = 1.0f;
  } else {
    const int j = i - 46;
    raw = j < 3 ? f0[c * 3 + j] : j < 6 ? f1[c * 3 + j - 3] : j < 9 ? f2[c * 3 + j - 6] : f3[c * 3 + j - 9];
    val = tanhf(raw); der = 1.0f - val * val;
  }
  dst[i] = val;
  dtr[i] = der;
}

// grid (C): one block per channel walks all N*P elements of that channel (z is small: 128 x N*H/64*W/64),
// so the 58 parameter gradients of a channel are reduced inside one block — deterministic, no atomics.
// z / noise / dz: NHWC fp32 [N*P][C]; z_hat / lik: NCHW fp32 (the reference's layout); zq: NHWC bf16.
__global__ void __launch_bounds__(256)
eb_train_kernel(const float* __restrict__ z, const float* __restrict__ noise, int NP, int P, int C,
                const float* m0, const float* m1, const float* m2, const float* m3, const float* m4, const float* b0,
                const float* b1, const float* b2, const float* b3, const float* b4, const float* f0, const float* f1,
                const float* f2, const float* f3, float cscale, float* __restrict__ z_hat, float* __restrict__ lik,
                __nv_bfloat16* __restrict__ zq, int bf_pitch, float* __restrict__ dz, float* __restrict__ dpar /*[C][58]*/) {
  __shared__ EBRaw sc;
  __shared__ float dtr[EB_NPAR];
  __shared__ float red[8][EB_NPAR];
  const int c = blockIdx.x;
  eb_load_chan(sc, dtr, c, m0, m1, m2, m3, m4, b0, b1, b2, b3, b4, f0, f1, f2, f3);
  __syncthreads();
  float acc[EB_NPAR];
#pragma unroll
  for (int i = 0; i < EB_NPAR; ++i) acc[i] = 0.0f;
  for (int e = threadIdx.x; e < NP; e += blockDim.x) {
    const float v = z[(long)e * C + c] + noise[(long)e * C + c];
    EBTrace tl, tu;
    const float lower = eb_fwd_trace(sc, v - 0.5f, tl);
    const float upper = eb_fwd_trace(sc, v + 0.5f, tu);
    const float s0 = lower + upper;
    const float sign = s0 > 0.0f ? -1.0f : (s0 < 0.0f ? 1.0f : 0.0f);
    const float su = sigmoid_f(sign * upper), sl = sigmoid_f(sign * lower);
    const float diff = su - sl;
    const float l = fabsf(diff);
    const float lb = fmaxf(l, kLikBound);
    const float gl = cscale / lb;
    const float sd = diff > 0.0f ? 1.0f : (diff < 0.0f ? -1.0f : 0.0f);
    const float gu = gl * sd * su * (1.0f - su) * sign;
    const float glo = -gl * sd * sl * (1.0f - sl) * sign;
    float dv = eb_bwd_trace(sc, v + 0.5f, tu, gu, acc);
    dv += eb_bwd_trace(sc, v - 0.5f, tl, glo, acc);
    const int n = e / P, p = e - n * P;
    if (z_hat) z_hat[((long)n * C + c) * P + p] = v;
    if (lik) lik[((long)n * C + c) * P + p] = lb;
    if (zq) zq[(long)e * bf_pitch + c] = __float2bfloat16_rn(v);
    if (dz) dz[(long)e * C + c] = dv;
  }
  // block reduction of the 58 accumulators
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < EB_NPAR; ++i) {
    float a = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) red[wp][i] = a;
  }
  __syncthreads();
  if (threadIdx.x < EB_NPAR) {
    float s = 0.0f;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    dpar[c * EB_NPAR + threadIdx.x] = s * dtr[threadIdx.x];
  }
}

// EntropyBottleneck.loss (entropy_models.py:345-348): sum |L(quantiles) - target| with the network detached;
// grid (ceil(C*3/128)); dq = d loss / d quantiles; loss accumulated into *loss (pre-zeroed)
__global__ void eb_aux_kernel(const float* __restrict__ quantiles, int C, const float* m0, const float* m1,
                              const float* m2, const float* m3, const float* m4, const float* b0, const float* b1,
                              const float* b2, const float* b3, const float* b4, const float* f0, const float* f1,
                              const float* f2, const float* f3, float t0, float t1, float t2, float* __restrict__ loss,
                              float* __restrict__ dq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float contrib = 0.0f;
  if (i < C * 3) {
    const int c = i / 3, j = i - 3 * c;
    EBRaw r;
    for (int q = 0; q < 3; ++q) {
      r.M0[q] = softplus_f(m0[c * 3 + q]); r.M4[q] = softplus_f(m4[c * 3 + q]);
      r.B0[q] = b0[c * 3 + q]; r.B1[q] = b1[c * 3 + q]; r.B2[q] = b2[c * 3 + q]; r.B3[q] = b3[c * 3 + q];
      r.F0[q] = tanhf(f0[c * 3 + q]); r.F1[q] = tanhf(f1[c * 3 + q]); r.F2[q] = tanhf(f2[c * 3 + q]);
      r.F3[q] = tanhf(f3[c * 3 + q]);
    }
    for (int q = 0; q < 9; ++q) {
      r.M1[q] = softplus_f(m1[c * 9 + q]); r.M2[q] = softplus_f(m2[c * 9 + q]); r.M3[q] = softplus_f(m3[c * 9 + q]);
    }
    r.B4[0] = b4[c];
    EBTrace t;
    const float v = quantiles[i];
    const float lg = eb_fwd_trace(r, v, t);
    const float tgt = j == 0 ? t0 : (j == 1 ? t1 : t2);
    const float d = lg - tgt;
    contrib = fabsf(d);
    float acc[EB_NPAR];
    for (int q = 0; q < EB_NPAR; ++q) acc[q] = 0.0f;
    const float sg = d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f);
    dq[i] = eb_bwd_trace(r, v, t, sg, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
  if ((threadIdx.x & 31) == 0 && contrib != 0.0f) atomicAdd(loss, contrib);
}

}  // namespace

extern "C" int masic_gmm_likelihood_train(const float* y, const float* noise, const float* sigma, const float* mu,
                                          const float* wlogits, int64_t n_pixels, int m, int k, float scale_bound,
                                          float lik_grad_scale, float* lik, void* y_hat_bf16, int bf_pitch,
                                          float* y_hat, float* dy, void* dsigma_bf16, void* dmu_bf16, void* dwl_bf16,
                                          void* stream) {
  if (!y || !noise || !sigma || !mu || !wlogits || n_pixels <= 0 || m <= 0) return MASIC_EINVAL;
  if (k != 5) return MASIC_ENOSUP;
  if (y_hat_bf16 && bf_pitch < m) return MASIC_EINVAL;
  const long total = (long)n_pixels * m;
  gmm_train_kernel<5><<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      y, noise, sigma, mu, wlogits, total, m, scale_bound, lik_grad_scale, lik,
      static_cast<__nv_bfloat16*>(y_hat_bf16), bf_pitch, y_hat, dy, static_cast<__nv_bfloat16*>(dsigma_bf16),
      static_cast<__nv_bfloat16*>(dmu_bf16), static_cast<__nv_bfloat16*>(dwl_bf16));
  return (int)cudaGetLastError();
}

extern "C" int masic_eb_train(const float* z_nhwc, const float* noise_nhwc, int n, int c, int hw,
                              const float* const* matrices, const float* const* biases, const float* const* factors,
                              float lik_grad_scale, float* z_hat_nchw, float* lik_nchw, void* zq_bf16, int bf_pitch,
                              float* dz_nhwc, float* dparams, void* stream) {
  if (!z_nhwc || !noise_nhwc || !matrices || !biases || !factors || !dparams || n <= 0 || c <= 0 || hw <= 0)
    return MASIC_EINVAL;
  eb_train_kernel<<<c, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      z_nhwc, noise_nhwc, n * hw, hw, c, matrices[0], matrices[1], matrices[2], matrices[3], matrices[4], biases[0],
      biases[1], biases[2], biases[3], biases[4], factors[0], factors[1], factors[2], factors[3], lik_grad_scale,
      z_hat_nchw, lik_nchw, static_cast<__nv_bfloat16*>(zq_bf16), bf_pitch, dz_nhwc, dparams);
  return (int)cudaGetLastError();
}

extern "C" int masic_eb_aux_loss(const float* quantiles, int c, const float* const* matrices,
                                 const float* const* biases, const float* const* factors, const float* target3_host,
                                 float* loss, float* dquantiles, void* stream) {
  if (!quantiles || !matrices || !biases || !factors || !target3_host || !loss || !dquantiles || c <= 0)
    return MASIC_EINVAL;
  eb_aux_kernel<<<(c * 3 + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      quantiles, c, matrices[0], matrices[1], matrices[2], matrices[3], matrices[4], biases[0], biases[1], biases[2],
      biases[3], biases[4], factors[0], factors[1], factors[2], factors[3], target3_host[0], target3_host[1],
      target3_host[2], loss, dquantiles);
  return (int)cudaGetLastError();
}
