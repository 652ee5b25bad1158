// gdn.cu — stand-alone GDN / IGDN on NCHW fp32 (compressai/layers/gdn.py:77-92), used when
// the reference's MASIC.py drives the layers one module at a time.  Inside HSICEngine the
// same normalisation is fused into the producing conv's epilogue (conv_tc.cu).
//   norm_i = beta'_i + sum_j gamma'_ij x_j^2 ;  out_i = x_i * rsqrt(norm_i)   (inverse: * sqrt)
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/masic_b200.h"

namespace {
constexpr int PIX = 32;        // pixels per block
constexpr int MAXC = 256;

__global__ void __launch_bounds__(256)
gdn_nchw_kernel(const float* __restrict__ x, int c, int hw, const float* __restrict__ beta,
                const float* __restrict__ gamma, float beta_bound, int inverse, float* __restrict__ out) {
  extern __shared__ float sm[];                 // x2[c][PIX] then gamma'[c][c] is streamed from global
  float* x2 = sm;
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * PIX;
  const int px = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  const float ped = 1.4551915228366852e-11f;      // 2^-36
  const float* xb = x + (long)n * c * hw;
  for (int ch = ty; ch < c; ch += 8) {
    const int p = p0 + px;
    const float v = p < hw ? xb[(long)ch * hw + p] : 0.0f;
    x2[ch * PIX + px] = v * v;
  }
  __syncthreads();
  const int p = p0 + px;
  for (int i = ty; i < c; i += 8) {
    const float b = fmaxf(beta[i], beta_bound);
    float nrm = b * b - ped;
    const float* g = gamma + (long)i * c;
    for (int j = 0; j < c; ++j) {
      const float gv = fmaxf(__ldg(g + j), 3.814697265625e-06f);
      nrm = fmaf(gv * gv - ped, x2[j * PIX + px], nrm);
    }
    if (p < hw) {
      const float xv = xb[(long)i * hw + p];
      out[((long)n * c + i) * hw + p] = xv * (inverse ? sqrtf(nrm) : rsqrtf(nrm));
    }
  }
}
}  // namespace

extern "C" int masic_gdn_nchw(const float* x, int n, int c, int hw, const float* beta, const float* gamma,
                              float beta_min, int inverse, float* out, void* stream) {
  if (!x || !beta || !gamma || !out || n <= 0 || c <= 0 || c > MAXC || hw <= 0) return MASIC_EINVAL;
  dim3 grid((hw + PIX - 1) / PIX, n), block(PIX, 8);
  gdn_nchw_kernel<<<grid, block, c * PIX * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      x, c, hw, beta, gamma, sqrtf(beta_min + 1.4551915228366852e-11f), inverse, out);
  return (int)cudaGetLastError();
}
