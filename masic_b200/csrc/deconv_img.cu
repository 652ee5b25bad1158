// deconv_img.cu — the image-resolution transposed convolution g_s_conv4 (ConvTranspose2d(128 -> 3, k=5, s=2, p=2, op=1),
// compressai/models/utils.py:138-146, used at MASIC.py:542 / :596) with its pixel interleave and, for the right view,
// after_gdn (MASIC.py:599, :615) fused — output straight into the reference's NCHW fp32 image.
//
// Why not conv_tc.cu's sub-pixel form: with 3 output channels every tcgen05.mma of that form has N = 16 and still pays
// ~68 cycles for its 128x16 A sub-tile, 72 MMAs per 128 input pixels (measured 0.089 ms, tensor pipe 13 % active).
// Here the layer is the GEMM of its col2im form
//     Z[pixel][(ky, kx, co)] = sum_ci x[pixel][ci] * W[ci][co][ky][kx]          M = pixels, K = 128, N = 25 taps x 4 = 100 -> 112
//     out[co][2 iy - 2 + ky][2 ix - 2 + kx] += Z[iy][ix][(ky, kx, co)]
// i.e. 8 MMAs (N = 112) per 128 input pixels, followed by a gather-sum of <= 9 Z entries per output pixel done from
// shared memory.  A CTA walks DOWN a strip of 16 input columns (14 of them new: the horizontal halo is recomputed),
// eight input rows per step; Z rows live in a 12-row shared-memory ring, so the vertical halo costs one extra step
// only where a CTA's range starts in the middle of a strip.  Out-of-image pixels are TMA zero fill, hence Z = 0.
//
// Roles: warp 0 TMA producer (two 16 KB k-blocks per step, 3 stages), warp 1 MMA issuer (8 MMAs per step into one of
// two TMEM accumulators), warps 2-17 epilogue: TMEM -> Z ring, then 16 output rows x 28 columns x 3 channels per step.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/masic_b200.h"
#include "cvt16.cuh"
#include "ptx.cuh"

namespace {
using namespace masic;

constexpr int DI_COLS = 16, DI_NEW = 14, DI_ROWS = 8;     // strip width (loaded / new), rows per step
constexpr int DI_N = 112;                                  // accumulator columns: (ky*5+kx)*4 + co, 100 used
constexpr int DI_ZP = 116;                                 // floats per pixel in the Z ring (conflict-free float4 rows)
constexpr int DI_RING = 12;                                // Z ring rows: 10 live (R-2 .. R+7) while the next 8 are written
                                                           // ... after the barrier that ends the step's gather, so 12 suffice
constexpr int DI_THREADS = 576, DI_EPI = 512;
constexpr int DI_A_STAGE = 2 * 16384;                      // two k-blocks of [8 rows][16 cols][64 ch]
constexpr int DI_A_STAGES = 3;
constexpr int DI_B_BYTES = 2 * DI_N * 128;                 // two k-blocks of [112 rows][64 ch]
constexpr int DI_Z_BYTES = DI_RING * DI_COLS * DI_ZP * 4;
constexpr int DI_MISC = 256;
constexpr int DI_SMEM = DI_A_STAGES * DI_A_STAGE + DI_B_BYTES + DI_Z_BYTES + DI_MISC + 1024;

struct DIParams {
  CUtensorMap tmA, tmB;
  int n_img, h2, w2;            // input (half-resolution) size
  int n_strips, steps_per_strip, total_steps;
  const float* bias;            // [3] or null
  int gdn;                      // 0 or MASIC_GDN_INV (after_gdn)
  float beta[3], gamma[9];      // effective (re-parametrised) values
  float* out;                   // (n, 3, 2 h2, 2 w2) fp32, or null
  uint16_t* out16;              // optional NHWC 16-bit copy: pixel (y, x) at out16[((img*HO + y)*o16_row + x + o16_xoff)*o16_pitch + o16_coff + c]
  int o16_pitch, o16_row, o16_xoff, o16_coff, o16_f16;
  uint32_t idesc;
};

__global__ void __launch_bounds__(DI_THREADS, 1)
deconv_img_kernel(const __grid_constant__ DIParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base, sB = base + DI_A_STAGES * DI_A_STAGE;
  float* zring = reinterpret_cast<float*>(gen + DI_A_STAGES * DI_A_STAGE + DI_B_BYTES);
  const uint32_t sMisc = base + DI_A_STAGES * DI_A_STAGE + DI_B_BYTES + DI_Z_BYTES;
  // barriers: A full[3] @0, A empty[3] @24, acc full[2] @48, acc empty[2] @64, B full @80, tmem ptr @88
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(gen + DI_A_STAGES * DI_A_STAGE + DI_B_BYTES + DI_Z_BYTES + 88);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int i = 0; i < DI_A_STAGES; ++i) {
      mbar_init(sMisc + 8 * i, 1);
      mbar_init(sMisc + 24 + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(sMisc + 48 + 8 * i, 1);
      mbar_init(sMisc + 64 + 8 * i, DI_EPI);
    }
    mbar_init(sMisc + 80, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(sMisc + 88, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // this CTA's contiguous range of (image, strip, step) in strip-major order
  const long T = p.total_steps;
  const int u0 = static_cast<int>(T * blockIdx.x / gridDim.x), u1 = static_cast<int>(T * (blockIdx.x + 1) / gridDim.x);
  const int S = p.steps_per_strip;
  // iteration space of every role: for u in [u0, u1): if u is the first of the range and not the top of its strip, the
  // halo step u - 1 comes first (computed, nothing emitted)
  auto first_iter = [&]() -> int { return (u0 < u1 && (u0 % S) != 0) ? u0 - 1 : u0; };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (u0 < u1 && elect_one()) {
      mbar_expect_tx(sMisc + 80, DI_B_BYTES);
      tma_load_2d(sB, &p.tmB, sMisc + 80, 0, 0);
      tma_load_2d(sB + DI_N * 128, &p.tmB, sMisc + 80, 0, DI_N);
    }
    __syncwarp();
    uint32_t st = 0, ph = 0;
    for (int u = first_iter(); u < u1; ++u) {
      const int s = u % S, strip = (u / S) % p.n_strips, img = u / (S * p.n_strips);
      mbar_wait(sMisc + 24 + 8 * st, ph ^ 1);
      if (elect_one()) {
        const uint32_t full = sMisc + 8 * st;
        mbar_expect_tx(full, DI_A_STAGE);
        tma_load_4d(sA + st * DI_A_STAGE, &p.tmA, full, 0, strip * DI_NEW - 1, s * DI_ROWS, img);
        tma_load_4d(sA + st * DI_A_STAGE + 16384, &p.tmA, full, 64, strip * DI_NEW - 1, s * DI_ROWS, img);
      }
      __syncwarp();
      if (++st == DI_A_STAGES) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (u0 < u1) mbar_wait(sMisc + 80, 0);
    uint32_t st = 0, ph = 0;
    int it = 0;
    const uint64_t descA0 = umma_desc_sw128(sA), descB0 = umma_desc_sw128(sB);
    for (int u = first_iter(); u < u1; ++u, ++it) {
      const int buf = it & 1;
      mbar_wait(sMisc + 64 + 8 * buf, ((it >> 1) & 1) ^ 1);
      mbar_wait(sMisc + 8 * st, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tmem_base + buf * 128;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t a = descA0 + ((st * DI_A_STAGE + kb * 16384) >> 4), b = descB0 + ((kb * DI_N * 128) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d, a + 2 * k, b + 2 * k, p.idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(sMisc + 24 + 8 * st);
        umma_commit(sMisc + 48 + 8 * buf);
      }
      __syncwarp();
      if (++st == DI_A_STAGES) { st = 0; ph ^= 1; }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;                 // TMEM lane quarter of this warp
    const int cg = (warp - 2) >> 2;         // column group: columns 32 cg .. 32 cg + 31 (the last group holds 16)
    const int m = q * 32 + lane;            // pixel of the step's 8 x 16 block: row m >> 4, column m & 15
    const int et = (warp - 2) * 32 + lane;  // 0..511
    const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
    const int H2 = p.h2, W2 = p.w2, HO = 2 * p.h2, WO = 2 * p.w2;
    float bias[3] = {0.f, 0.f, 0.f};
    if (p.bias) { bias[0] = p.bias[0]; bias[1] = p.bias[1]; bias[2] = p.bias[2]; }
    // gather geometry of this thread, constant over the whole kernel: output (row ro of the step's 18, column oxl of the
    // strip's 28).  oy = 2R - 2 + ro, ox = 2 X0 + oxl  =>  taps ky = py + 2 jy read input row R - 1 + ((ro + 2 - ky) >> 1),
    // taps kx = px + 2 jx read strip column ((oxl + 2 - kx) >> 1) + 1; only R changes from step to step.
    const int ro = et / (2 * DI_NEW), oxl = et - ro * (2 * DI_NEW);
    const int py = ro & 1, px = oxl & 1;
    const int nty = py ? 2 : 3, ntx = px ? 2 : 3;
    int drow[3], coff[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int ky = py + 2 * j;
      drow[j] = ((ro + 2 - ky) >> 1) - 1;                               // input row relative to R
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int kx = px + 2 * i;
        coff[j][i] = (((oxl + 2 - kx) >> 1) + 1) * DI_ZP + (ky * 5 + kx) * 4;    // float offset inside a ring row
      }
    }
    int it = 0;
    for (int u = first_iter(); u < u1; ++u, ++it) {
      const int s = u % S, strip = (u / S) % p.n_strips, img = u / (S * p.n_strips);
      const int buf = it & 1, R = s * DI_ROWS, X0 = strip * DI_NEW;
      mbar_wait(sMisc + 48 + 8 * buf, (it >> 1) & 1);
      tc_fence_after();
      // ---- Z rows R .. R+7 -> ring (pixel m: ring row (R + (m >> 4)) & 15, column m & 15)
      {
        float* zp = zring + ((((R + (m >> 4)) % DI_RING) * DI_COLS + (m & 15)) * DI_ZP) + cg * 32;
        const uint32_t ta = tmem_base + lane_sel + buf * 128 + cg * 32;
        uint32_t r[32];
        tmem_ld16(ta, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));          // both loads in flight before the single wait
        if (cg < 3) tmem_ld16(ta + 16, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
        tmem_ld_wait();
        uint4* dst = reinterpret_cast<uint4*>(zp);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < 4 || cg < 3) dst[c] = make_uint4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
      }
      tc_fence_before();
      mbar_arrive(sMisc + 64 + 8 * buf);            // the accumulator may be refilled
      named_bar_sync(1, DI_EPI);
      // ---- output rows made complete by this step (none for the halo step of a range): rows 2R-2 .. 2R+13, and the
      // image's last two rows with the strip's last step
      if (u >= u0) {
        const int oy = 2 * R - 2 + ro, ox = 2 * X0 + oxl;
        if (ro < (s == S - 1 ? 18 : 16) && oy >= 0 && ox < WO) {
          float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const int iy = R + drow[j];
            if (j < nty && iy >= 0 && iy < H2) {
              const float* zr = zring + ((iy % DI_RING) * DI_COLS) * DI_ZP;
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                if (i < ntx) {
                  const float4 z = *reinterpret_cast<const float4*>(zr + coff[j][i]);
                  a0 += z.x; a1 += z.y; a2 += z.z;
                }
              }
            }
          }
          float v0 = a0 + bias[0], v1 = a1 + bias[1], v2 = a2 + bias[2];
          if (p.gdn) {                                   // after_gdn: IGDN over the 3 channels (gdn.py:77-92, inverse)
            const float s0 = v0 * v0, s1 = v1 * v1, s2 = v2 * v2;
            const float n0 = fmaf(p.gamma[2], s2, fmaf(p.gamma[1], s1, fmaf(p.gamma[0], s0, p.beta[0])));
            const float n1 = fmaf(p.gamma[5], s2, fmaf(p.gamma[4], s1, fmaf(p.gamma[3], s0, p.beta[1])));
            const float n2 = fmaf(p.gamma[8], s2, fmaf(p.gamma[7], s1, fmaf(p.gamma[6], s0, p.beta[2])));
            v0 *= sqrtf(n0); v1 *= sqrtf(n1); v2 *= sqrtf(n2);
          }
          if (p.out) {
            float* o = p.out + (static_cast<size_t>(img) * 3 * HO + oy) * WO + ox;
            o[0] = v0;
            o[static_cast<size_t>(HO) * WO] = v1;
            o[2 * static_cast<size_t>(HO) * WO] = v2;
          }
          if (p.out16) {
            uint16_t* o = p.out16 + ((static_cast<size_t>(img) * HO + oy) * p.o16_row + ox + p.o16_xoff) * p.o16_pitch + p.o16_coff;
            if (((p.o16_coff | p.o16_pitch) & 3) == 0) {                         // one aligned 8-byte store: [v0 v1 v2 0]
              *reinterpret_cast<uint2*>(o) = make_uint2(pack16x2(v0, v1, p.o16_f16), pack16x2(v2, 0.0f, p.o16_f16));
            } else {
              *reinterpret_cast<uint32_t*>(o) = pack16x2(v0, v1, p.o16_f16);    // o16_coff is even: 4-byte aligned
              o[2] = pack16(v2, p.o16_f16);
            }
          }
        }
      }
      named_bar_sync(1, DI_EPI);                    // the next step's Z rows overwrite ring rows this step still read
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

// weights [128][3][5][5] (ConvTranspose2d layout) -> B operand, 16-bit [2 k-blocks][112 rows][64 ch], row = (ky*5+kx)*4 + co
__global__ void di_pack_kernel(const float* __restrict__ w, uint16_t* __restrict__ dst, int f16) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * DI_N * 64) return;
  const int c = i & 63, row = (i >> 6) % DI_N, kb = i / (64 * DI_N);
  const int tap = row >> 2, co = row & 3, ci = kb * 64 + c;
  float v = 0.0f;
  if (tap < 25 && co < 3) v = w[(ci * 3 + co) * 25 + tap];
  dst[i] = pack16(v, f16);
}

}  // namespace

struct MasicDeconvImgPlan {
  DIParams kp;
  int grid;
};

extern "C" int64_t masic_deconv_img_weight_bytes(void) { return 2 * DI_N * 64 * 2; }

extern "C" int masic_deconv_img_pack_weights(const float* weight, void* dst_16, int f16, void* stream) {
  if (!weight || !dst_16) return MASIC_EINVAL;
  di_pack_kernel<<<(2 * DI_N * 64 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      weight, static_cast<uint16_t*>(dst_16), f16);
  return (int)cudaGetLastError();
}

extern "C" int masic_deconv_img_plan_create(const void* in_nhwc16, int n, int h_in, int w_in, int c_pitch,
                                            const void* w_packed, const float* bias3, int gdn, const float* beta3_host,
                                            const float* gamma9_host, float* out_nchw, int f16,
                                            MasicDeconvImgPlan** plan_out) {
  if (!in_nhwc16 || !w_packed || !plan_out || n <= 0 || h_in <= 0 || w_in <= 0 || (h_in % DI_ROWS) ||
      c_pitch != 128 || (gdn && (!beta3_host || !gamma9_host)) || (gdn && gdn != MASIC_GDN_INV))
    return MASIC_EINVAL;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return MASIC_EDRIVER;
  MasicDeconvImgPlan* pl = new MasicDeconvImgPlan();
  DIParams& kp = pl->kp;
  memset(&kp, 0, sizeof(kp));
  {
    cuuint64_t dims[4] = {128, (cuuint64_t)w_in, (cuuint64_t)h_in, (cuuint64_t)n};
    cuuint64_t strides[3] = {256, (cuuint64_t)w_in * 256, (cuuint64_t)h_in * w_in * 256};
    cuuint32_t box[4] = {64, DI_COLS, DI_ROWS, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&kp.tmA, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                     const_cast<void*>(in_nhwc16), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete pl; return MASIC_EDRIVER; }
  }
  {
    cuuint64_t dims[2] = {64, 2 * DI_N};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, DI_N};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&kp.tmB, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     const_cast<void*>(w_packed), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete pl; return MASIC_EDRIVER; }
  }
  kp.n_img = n; kp.h2 = h_in; kp.w2 = w_in;
  kp.n_strips = (w_in + DI_NEW - 1) / DI_NEW;
  kp.steps_per_strip = h_in / DI_ROWS;
  kp.total_steps = n * kp.n_strips * kp.steps_per_strip;
  kp.bias = bias3;
  kp.gdn = gdn;
  if (gdn) {   // NonNegativeParametrizer (parametrizers.py:61-64) applied on the host: 12 numbers
    const float ped = 1.4551915228366852e-11f, bb = sqrtf(1e-6f + ped), gb = 3.814697265625e-06f;
    for (int i = 0; i < 3; ++i) { const float b = beta3_host[i] > bb ? beta3_host[i] : bb; kp.beta[i] = b * b - ped; }
    for (int i = 0; i < 9; ++i) { const float g = gamma9_host[i] > gb ? gamma9_host[i] : gb; kp.gamma[i] = g * g - ped; }
  }
  kp.out = out_nchw;
  kp.idesc = umma_idesc_bf16(DI_N);
  if (f16) kp.idesc &= ~((1u << 7) | (1u << 10));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  pl->grid = kp.total_steps < sms ? kp.total_steps : sms;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(deconv_img_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DI_SMEM);
    if (e != cudaSuccess) { delete pl; return (int)e; }
    attr = true;
  }
  *plan_out = pl;
  return MASIC_OK;
}

extern "C" int masic_deconv_img_plan_set_out16(MasicDeconvImgPlan* pl, void* out_nhwc16, int c_pitch, int row_pixels,
                                               int xoff, int coff, int f16) {
  if (!pl || (out_nhwc16 && (c_pitch < coff + 3 || (c_pitch % 2) || (coff % 2) || coff < 0 || xoff < 0 ||
                             row_pixels < 2 * pl->kp.w2 + xoff || (reinterpret_cast<uintptr_t>(out_nhwc16) & 7))))
    return MASIC_EINVAL;
  pl->kp.out16 = static_cast<uint16_t*>(out_nhwc16);
  pl->kp.o16_pitch = c_pitch; pl->kp.o16_row = row_pixels; pl->kp.o16_xoff = xoff; pl->kp.o16_coff = coff;
  pl->kp.o16_f16 = f16 & 1;
  return MASIC_OK;
}

extern "C" int masic_deconv_img_plan_launch(const MasicDeconvImgPlan* pl, void* stream) {
  if (!pl || (!pl->kp.out && !pl->kp.out16)) return MASIC_EINVAL;
  deconv_img_kernel<<<pl->grid, DI_THREADS, DI_SMEM, static_cast<cudaStream_t>(stream)>>>(pl->kp);
  return (int)cudaGetLastError();
}

extern "C" void masic_deconv_img_plan_destroy(MasicDeconvImgPlan* pl) { delete pl; }
