// conv_tc.cu — implicit-GEMM convolution / transposed convolution on the 5th-gen tensor
// cores (tcgen05.mma, accumulators in TMEM), operands staged by TMA, GDN/IGDN fused
// into the epilogue.  sm_100a only.
//
// Replaces (reference, file:line):
//   conv()/deconv() factories          compressai/models/utils.py:128-146
//   GDN.forward                        compressai/layers/gdn.py:77-92
//   MaskedConv2d.forward               compressai/layers/layers.py:75-78
//   the nn.ReLU / nn.LeakyReLU that follow them in MASIC.py:173-183,338-444,678-691
//
// Design:
//   * activations NHWC bf16; one CTA tile = 16 rows x 8 cols of output positions (M = 128);
//   * the A operand of every filter tap is a *row-shifted window* of a "strip" —
//     (16 + taps-1) rows x 8 cols x 64 channels — that TMA drops into shared memory once,
//     so vertical taps re-use the same bytes (rows are 1024 B = one SWIZZLE_128B atom, so
//     every window start stays atom-aligned);
//   * stride-2 convs read the input through a 5-D "phase split" view
//     (2C, W/2, 2, H/2, N) of the same NHWC buffer, stride-2 transposed convs write their
//     output through the same view — no im2col, no scatter kernels;
//   * a per-layer *program* (A loads, B loads, MMA ops) built on the host drives three
//     single-thread roles (A producer, B producer, MMA issuer); 4 epilogue warps drain the
//     double-buffered TMEM accumulator: +bias, activation, optional GDN (a second
//     128x128x128 tcgen05.mma on the squared tile against gamma), optional per-pixel scale,
//     then swizzled smem staging and a TMA store.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/masic_b200.h"
#include "ptx.cuh"

namespace masic {

constexpr int TILE_W = 8;
constexpr int TILE_H = 16;
constexpr int KBLK = 64;            // channels per k-block: 128 B of bf16 = one swizzle row
constexpr int NUM_THREADS = 512;    // stream 0: warps 0 (A producer), 1 (B producer), 2 (MMA); warp 3: TMEM alloc;
                                    // warps 4-7: epilogue group 0; stream 1: warps 8 (A), 9 (B), 10 (MMA);
                                    // warp 11 idle; warps 12-15: epilogue group 1 (warp%4 = TMEM lane quarter)
constexpr int EPI_THREADS = 256;    // both epilogue groups
constexpr int NUM_STREAMS = 2;      // two independent load->MMA pipelines per CTA, one accumulator buffer each
constexpr int MAX_VARIANTS = 4;
constexpr int MAX_STAGES = 8;
constexpr int MAX_AOPS = 64, MAX_BOPS = 96, MAX_MOPS = 96;
constexpr int STAGE_BLK_BYTES = 16384;  // one 128-row x 128-B staging block
constexpr uint32_t TMEM_COLS = 512;

struct AOp { int16_t c0, dx, p2, dy; };          // TMA coordinates relative to the tile origin
struct BOp { int32_t row0; };                    // first row of the k-block in the packed weights
struct MOp { uint16_t a_row; uint8_t nk; uint8_t flags; };   // nk: low nibble = #K16 steps, high nibble = first step
enum { M_NEW_A = 1, M_NEW_B = 2, M_FIRST = 4, M_REL_A = 8, M_REL_B = 16 };

struct Variant {
  int n_aops, n_bops, n_mops;
  int aops_off, bops_off, mops_off;
  int out_p2;   // phase-row coordinate in the 5-D output view
  int out_c0;   // channel offset inside the output view (px * Cpitch)
};

struct KParams {
  CUtensorMap tmA, tmB, tmO, tmG;
  // per-layer programs live in the kernel-parameter constant bank: the role loops read them with
  // warp-uniform indices, so ptxas keeps ops / descriptors in uniform registers (no R2UR, no L1 trip)
  AOp aops[MAX_AOPS];
  BOp bops[MAX_BOPS];
  MOp mops[MAX_MOPS];
  Variant var[MAX_VARIANTS];
  int n_var;
  int tiles_x, tiles_y, n_img, n_ntiles;
  int n_tile;
  int a_stage_bytes, b_stage_bytes, a_stages, b_stages;
  int smem_b_off, smem_g_off, smem_stage_off, smem_misc_off;
  const float* bias;
  const float* beta;
  int gdn, out_fp32;
  int blk_ch;      // channels per staging block / TMA store
  int blk_pitch;   // bytes per row of a staging block (128 = swizzled)
  int stage_per_group;   // staging buffers per epilogue group (1 or 2)
  uint8_t act[32];
  int out_coff;
  const float* rowscale;
  int rs_stride, rs_off, rs_H, rs_W;
  uint32_t idesc;
  long long* trace; // optional [64 tiles][16] clock64() stamps of CTA 0 (MASIC_CONV_TRACE=1)
  int debug;       // bit0: skip A loads, bit1: skip B loads (timing experiments only; results are garbage)
};

// misc smem region layout (byte offsets from smem_misc_off)
constexpr int MISC_A_FULL = 0;                       // [2 streams][8] x u64
constexpr int MISC_A_EMPTY = 128;
constexpr int MISC_B_FULL = 256;
constexpr int MISC_B_EMPTY = 384;
constexpr int MISC_ACC_FULL = 512;                   // 2 x u64
constexpr int MISC_ACC_EMPTY = 528;                  // 2 x u64
constexpr int MISC_GDN_BAR = 544;
constexpr int MISC_G_FULL = 552;
constexpr int MISC_TMEM_PTR = 560;
constexpr int MISC_BYTES = 1024;

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == MASIC_ACT_RELU) return fmaxf(x, 0.0f);
  if (act == MASIC_ACT_LEAKY) return x > 0.0f ? x : 0.01f * x;
  return x;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                             uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c),
               "r"(d)
               : "memory");
}

#define TRACE(it, slot) do { if (p.trace && blockIdx.x == 0 && (it) < 64) p.trace[(it) * 16 + (slot)] = clock64(); } while (0)

struct Work { int n, ty, tx, var, nt; };
__device__ __forceinline__ Work decode_work(const KParams& p, int w) {
  Work r;
  r.nt = w % p.n_ntiles; w /= p.n_ntiles;
  r.var = w % p.n_var;   w /= p.n_var;
  r.tx = w % p.tiles_x;  w /= p.tiles_x;
  r.ty = w % p.tiles_y;  r.n = w / p.tiles_y;
  return r;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ KParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms must be 1024-B aligned in the shared window
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + p.smem_b_off;
  const uint32_t sG = smem_base + p.smem_g_off;
  const uint32_t sStage = smem_base + p.smem_stage_off;
  const uint32_t sMisc = smem_base + p.smem_misc_off;
  volatile uint32_t* tmem_ptr_s =
      reinterpret_cast<volatile uint32_t*>(smem_gen + p.smem_misc_off + MISC_TMEM_PTR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total = p.n_img * p.tiles_y * p.tiles_x * p.n_var * p.n_ntiles;
  const int acc_stride = p.n_tile <= 128 ? 128 : 256;   // TMEM columns between the two buffers

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmO);
    for (int i = 0; i < NUM_STREAMS * MAX_STAGES; ++i) {
      mbar_init(sMisc + MISC_A_FULL + 8 * i, 1);
      mbar_init(sMisc + MISC_A_EMPTY + 8 * i, 1);
      mbar_init(sMisc + MISC_B_FULL + 8 * i, 1);
      mbar_init(sMisc + MISC_B_EMPTY + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(sMisc + MISC_ACC_FULL + 8 * i, 1);
      mbar_init(sMisc + MISC_ACC_EMPTY + 8 * i, EPI_THREADS);
    }
    mbar_init(sMisc + MISC_GDN_BAR, 1);
    mbar_init(sMisc + MISC_G_FULL, 1);
    fence_mbar_init();
  }
  if (warp == 3) {
    tmem_alloc(sMisc + MISC_TMEM_PTR, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  // Role loops are WARP-UNIFORM (all 32 lanes walk them, barrier waits included) and only the
  // issue instructions sit under elect_one(): tcgen05.mma / TMA are uniform-datapath instructions,
  // and inside a divergent `if (lane == 0)` region ptxas wraps each of them in an ELECT/BRA.U.ANY
  // waterfall loop (measured: ~320 cycles per 128x128x16 MMA instead of 64).
  // stream s owns tiles it = s, s+2, ... of this CTA's sequence, accumulator buffer s, and its own rings
  const int sid = (warp >= 8 && warp < 12) ? 1 : 0;
  const int role = warp - 8 * sid;            // 0: A producer, 1: B producer, 2: MMA (warps 3..7 handled below)
  const uint32_t sAs = sA + sid * p.a_stages * p.a_stage_bytes;
  const uint32_t sBs = sB + sid * p.b_stages * p.b_stage_bytes;
  const uint32_t bar_off = sid * MAX_STAGES * 8;
  const int w_first = blockIdx.x + sid * gridDim.x, w_step = NUM_STREAMS * gridDim.x;
  if ((warp < 3 || (warp >= 8 && warp < 11)) && role == 0) {
    // ===================== A producer: activation strips =====================
    uint32_t st = 0, ph = 0;
    const uint32_t n_st = p.a_stages, st_bytes = p.a_stage_bytes;
    for (int w = w_first; w < total; w += w_step) {
      const Work wk = decode_work(p, w);
      const int x0 = wk.tx * TILE_W, y0 = wk.ty * TILE_H;
      const int i0 = p.var[wk.var].aops_off, i1 = i0 + p.var[wk.var].n_aops;
      for (int i = i0; i < i1; ++i) {
        const AOp op = p.aops[i];
        const uint32_t full = sMisc + MISC_A_FULL + bar_off + 8 * st;
        mbar_wait(sMisc + MISC_A_EMPTY + bar_off + 8 * st, ph ^ 1);
        if (elect_one()) {
          if (p.debug & 1) {
            mbar_arrive(full);
          } else {
            mbar_expect_tx(full, st_bytes);
            tma_load_5d(sAs + st * st_bytes, &p.tmA, full, op.c0, x0 + op.dx, op.p2, y0 + op.dy, wk.n);
          }
        }
        __syncwarp();
        if (++st == n_st) { st = 0; ph ^= 1; }
      }
    }
  } else if ((warp < 3 || (warp >= 8 && warp < 11)) && role == 1) {
    // ===================== B producer: weight k-blocks (+ gamma once) =====================
    if (p.gdn && sid == 0 && elect_one()) {
      tma_prefetch_desc(&p.tmG);
      mbar_expect_tx(sMisc + MISC_G_FULL, 2 * STAGE_BLK_BYTES);
      tma_load_2d(sG, &p.tmG, sMisc + MISC_G_FULL, 0, 0);
      tma_load_2d(sG + STAGE_BLK_BYTES, &p.tmG, sMisc + MISC_G_FULL, KBLK, 0);
    }
    __syncwarp();
    uint32_t st = 0, ph = 0;
    const uint32_t n_st = p.b_stages, st_bytes = p.b_stage_bytes;
    for (int w = w_first; w < total; w += w_step) {
      const Work wk = decode_work(p, w);
      const int nrow = wk.nt * p.n_tile;
      const int i0 = p.var[wk.var].bops_off, i1 = i0 + p.var[wk.var].n_bops;
      for (int i = i0; i < i1; ++i) {
        const BOp op = p.bops[i];
        const uint32_t full = sMisc + MISC_B_FULL + bar_off + 8 * st;
        mbar_wait(sMisc + MISC_B_EMPTY + bar_off + 8 * st, ph ^ 1);
        if (elect_one()) {
          if (p.debug & 2) {
            mbar_arrive(full);
          } else {
            mbar_expect_tx(full, st_bytes);
            tma_load_2d(sBs + st * st_bytes, &p.tmB, full, 0, op.row0 + nrow);
          }
        }
        __syncwarp();
        if (++st == n_st) { st = 0; ph ^= 1; }
      }
    }
  } else if ((warp < 3 || (warp >= 8 && warp < 11)) && role == 2) {
    // ===================== MMA issuer =====================
    const uint32_t n_sa = p.a_stages, n_sb = p.b_stages;
    uint32_t sa = n_sa - 1, pa = 1, sb = n_sb - 1, pb = 1;    // first advance lands on stage 0, phase 0
    int it = sid;
    // descriptor = constant high part | (smem address >> 4); +2 per K=16 step, +64 per strip row
    const uint64_t descA0 = umma_desc_sw128(sAs), descB0 = umma_desc_sw128(sBs);
    const uint32_t a_step = p.a_stage_bytes >> 4, b_step = p.b_stage_bytes >> 4;
    const uint32_t idesc = p.idesc;
    for (int w = w_first; w < total; w += w_step, it += NUM_STREAMS) {
      const Work wk = decode_work(p, w);
      const int buf = sid;                      // == it & 1
      if (lane == 0) TRACE(it, 0);
      mbar_wait(sMisc + MISC_ACC_EMPTY + 8 * buf, ((it >> 1) & 1) ^ 1);
      if (lane == 0) TRACE(it, 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * acc_stride;
      const int i0 = p.var[wk.var].mops_off, i1 = i0 + p.var[wk.var].n_mops;
      for (int i = i0; i < i1; ++i) {
        const MOp op = p.mops[i];
#define TRACE_OP(slot) do { if (p.trace && blockIdx.x == 0 && it == 4 && (i - i0) < 32 && lane == 0) p.trace[(32 + i - i0) * 16 + (slot)] = clock64(); } while (0)
        TRACE_OP(0);
        if (op.flags & M_NEW_A) {
          if (++sa == n_sa) { sa = 0; pa ^= 1; }
          mbar_wait(sMisc + MISC_A_FULL + bar_off + 8 * sa, pa);
        }
        if (op.flags & M_NEW_B) {
          if (++sb == n_sb) { sb = 0; pb ^= 1; }
          mbar_wait(sMisc + MISC_B_FULL + bar_off + 8 * sb, pb);
        }
        TRACE_OP(1);
        tc_fence_after();
        TRACE_OP(2);
        if (elect_one()) {
          const uint32_t nk = op.nk & 15u, k0 = op.nk >> 4;                      // +32 B (= 2) per K=16 step
          const uint64_t adesc = descA0 + (sa * a_step + op.a_row * 64u + 2u * k0);
          const uint64_t bdesc = descB0 + (sb * b_step + 2u * k0);
          umma_bf16(d_tmem, adesc, bdesc, idesc, (op.flags & M_FIRST) ? 0u : 1u);
          if (nk > 1) umma_bf16(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
          if (nk > 2) umma_bf16(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
          if (nk > 3) umma_bf16(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
          if (op.flags & M_REL_A) umma_commit(sMisc + MISC_A_EMPTY + bar_off + 8 * sa);
          if (op.flags & M_REL_B) umma_commit(sMisc + MISC_B_EMPTY + bar_off + 8 * sb);
          if (i + 1 == i1) umma_commit(sMisc + MISC_ACC_FULL + 8 * buf);
        }
        TRACE_OP(3);
        __syncwarp();
        TRACE_OP(4);
      }
      if (lane == 0) TRACE(it, 2);
    }
  } else if ((warp >= 4 && warp < 8) || warp >= 12) {
    // ===================== epilogue: TMEM -> regs -> smem -> TMA store =====================
    // Two groups of 4 warps; group g owns staging buffer g and the channel blocks j = g, g+2, ...
    // (with GDN: the 64 channels [64g, 64g+64)).  Bias/beta are read through L1 (uniform addresses).
    const int grp = warp >= 12 ? 1 : 0;
    const int ew = warp & 3;            // the TMEM lane quarter this warp may read
    const int t = ew * 32 + lane;       // accumulator row = tile position
    const uint32_t lane_sel = static_cast<uint32_t>(ew * 32) << 16;
    const uint32_t gbar = 1 + grp;      // named barrier of this group (128 threads); barrier 3 = both groups
    const bool leader = (t == 0);       // issues this group's TMA stores (bulk groups are per thread)
    const int nblk = p.n_tile / p.blk_ch;
    const int chunks_per_blk = p.blk_ch / 16;
    // staging: with GDN one 16 KB buffer per group (together they are the 32 KB A2 operand of the norm MMA);
    // without GDN (no gamma resident) two buffers per group, so a block's TMA store drains behind the next block
    const uint32_t sbuf = sStage + grp * p.stage_per_group * STAGE_BLK_BYTES;
    uint32_t flip = 0;
    const float* __restrict__ bias_g = p.bias;
    int it = 0;
    if (p.gdn && grp == 0 && ew == 0) mbar_wait(sMisc + MISC_G_FULL, 0);
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
      const Work wk = decode_work(p, w);
      const Variant& v = p.var[wk.var];
      const int buf = it & 1;
      const int act = p.act[wk.nt];
      const uint32_t acc_addr = tmem_base + lane_sel + buf * acc_stride;
      const float* __restrict__ bias_t = bias_g ? bias_g + wk.nt * p.n_tile : nullptr;
      if (t == 0 && grp == 0) TRACE(it, 4);
      float rs = 1.0f;
      if (p.rowscale) {
        const int y = wk.ty * TILE_H + (t >> 3), x = wk.tx * TILE_W + (t & 7);
        if (y < p.rs_H && x < p.rs_W)
          rs = p.rowscale[(static_cast<size_t>(wk.n * p.rs_H + y) * p.rs_W + x) * p.rs_stride + p.rs_off];
      }
      if (t == 0 && grp == 0) TRACE(it, 5);
      mbar_wait(sMisc + MISC_ACC_FULL + 8 * buf, (it >> 1) & 1);
      tc_fence_after();
      if (t == 0 && grp == 0) TRACE(it, 6);

      if (p.gdn) {
        // ---- pass 1: A2[:, 64g .. 64g+64) = bf16((acc + bias)^2), K-major SWIZZLE_128B block g
        if (leader) tma_store_wait_read<0>();          // staging buffer g (== A2 block g) free again
        named_bar_sync(gbar, 128);
        const uint32_t arow = sbuf + t * 128;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int c = grp * 64 + cc * 16;
          uint32_t r[16];
          tmem_ld16(acc_addr + c, r);
          tmem_ld_wait();
          uint32_t q[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float x0 = __uint_as_float(r[2 * i]) + __ldg(bias_t + c + 2 * i);
            const float x1 = __uint_as_float(r[2 * i + 1]) + __ldg(bias_t + c + 2 * i + 1);
            q[i] = pack_bf16x2(x0 * x0, x1 * x1);
          }
          st_shared_v4(arow + (((2 * cc) ^ (t & 7)) << 4), q[0], q[1], q[2], q[3]);
          st_shared_v4(arow + (((2 * cc + 1) ^ (t & 7)) << 4), q[4], q[5], q[6], q[7]);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        named_bar_sync(3, EPI_THREADS);
        if (t == 0 && grp == 0) TRACE(it, 7);
        if (grp == 0 && ew == 0) {                      // warp-uniform; one elected lane issues the 8 MMAs
          tc_fence_after();
          if (elect_one()) {
            const uint32_t d2 = tmem_base + 2 * acc_stride;
            const uint64_t dh = umma_desc_sw128(0);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
              const uint64_t a2 = dh | (((sStage + kb * STAGE_BLK_BYTES) & 0x3FFFFu) >> 4);
              const uint64_t g2 = dh | (((sG + kb * STAGE_BLK_BYTES) & 0x3FFFFu) >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16(d2, a2 + 2 * k, g2 + 2 * k, p.idesc, (kb | k) ? 1u : 0u);
            }
            umma_commit(sMisc + MISC_GDN_BAR);
          }
          __syncwarp();
        }
        mbar_wait(sMisc + MISC_GDN_BAR, it & 1);
        tc_fence_after();
        if (t == 0 && grp == 0) TRACE(it, 8);
        // ---- pass 2: out[:, 64g .. 64g+64) = x * rsqrt(beta + norm)  (IGDN: x * sqrt(.))
        const uint32_t norm_addr = tmem_base + lane_sel + 2 * acc_stride;
        const float* __restrict__ beta_g = p.beta;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int c = grp * 64 + cc * 16;
          uint32_t r[16], g[16];
          tmem_ld16(acc_addr + c, r);
          tmem_ld16(norm_addr + c, g);
          tmem_ld_wait();
          float o[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float x = __uint_as_float(r[i]) + __ldg(bias_t + c + i);
            const float nrm = __uint_as_float(g[i]) + __ldg(beta_g + c + i);
            // IGDN: sqrt(n) = n * rsqrt(n) on the MUFU fast path (2-ulp rsqrt is far below bf16 rounding)
            const float rs_n = rsqrtf(nrm);
            o[i] = ((p.gdn == MASIC_GDN_FWD) ? x * rs_n : x * (nrm * rs_n)) * rs;
          }
          st_shared_v4(arow + (((2 * cc) ^ (t & 7)) << 4), pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                       pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
          st_shared_v4(arow + (((2 * cc + 1) ^ (t & 7)) << 4), pack_bf16x2(o[8], o[9]), pack_bf16x2(o[10], o[11]),
                       pack_bf16x2(o[12], o[13]), pack_bf16x2(o[14], o[15]));
        }
        tc_fence_before();
        mbar_arrive(sMisc + MISC_ACC_EMPTY + 8 * buf);
        fence_proxy_async_smem();
        named_bar_sync(gbar, 128);
        if (leader && !(p.debug & 4)) {
          tma_store_5d(&p.tmO, sbuf, p.out_coff + v.out_c0 + wk.nt * p.n_tile + grp * 64, wk.tx * TILE_W, v.out_p2,
                       wk.ty * TILE_H, wk.n);
          tma_store_commit();
        }
      } else {
        // ---- plain epilogue: bias, activation, per-pixel scale; this group's blocks j = grp, grp+2, ...
        for (int j = grp; j < nblk; j += 2) {
          const uint32_t sb2 = sbuf + flip * STAGE_BLK_BYTES;
          if (p.stage_per_group == 2) {
            flip ^= 1;
            if (leader) tma_store_wait_read<1>();      // the store issued two blocks ago has left this buffer
          } else if (leader) {
            tma_store_wait_read<0>();
          }
          named_bar_sync(gbar, 128);
          const uint32_t row = sb2 + t * p.blk_pitch;
          const int sw = (p.blk_pitch == 128) ? (t & 7) : 0;
          for (int cc = 0; cc < chunks_per_blk; ++cc) {
            const int c = j * p.blk_ch + cc * 16;
            uint32_t r[16];
            float o[16];
            tmem_ld16(acc_addr + c, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              o[i] = apply_act(__uint_as_float(r[i]) + (bias_t ? __ldg(bias_t + c + i) : 0.0f), act) * rs;
            if (p.out_fp32) {
              const int ch0 = cc * 4;
#pragma unroll
              for (int i = 0; i < 4; ++i)
                st_shared_v4(row + (((ch0 + i) ^ sw) << 4), __float_as_uint(o[4 * i]), __float_as_uint(o[4 * i + 1]),
                             __float_as_uint(o[4 * i + 2]), __float_as_uint(o[4 * i + 3]));
            } else {
              const int ch0 = cc * 2;
#pragma unroll
              for (int i = 0; i < 2; ++i)
                st_shared_v4(row + (((ch0 + i) ^ sw) << 4), pack_bf16x2(o[8 * i], o[8 * i + 1]),
                             pack_bf16x2(o[8 * i + 2], o[8 * i + 3]), pack_bf16x2(o[8 * i + 4], o[8 * i + 5]),
                             pack_bf16x2(o[8 * i + 6], o[8 * i + 7]));
            }
          }
          if (j + 2 >= nblk) {      // this group's last TMEM read of the accumulator
            tc_fence_before();
            mbar_arrive(sMisc + MISC_ACC_EMPTY + 8 * buf);
          }
          fence_proxy_async_smem();
          named_bar_sync(gbar, 128);
          if (leader && !(p.debug & 4)) {
            tma_store_5d(&p.tmO, sb2, p.out_coff + v.out_c0 + wk.nt * p.n_tile + j * p.blk_ch, wk.tx * TILE_W,
                         v.out_p2, wk.ty * TILE_H, wk.n);
            tma_store_commit();
          }
        }
        if (grp >= nblk) {          // nothing to do for this group (single-block tile): just release the accumulator
          tc_fence_before();
          mbar_arrive(sMisc + MISC_ACC_EMPTY + 8 * buf);
        }
      }
      if (t == 0 && grp == 0) TRACE(it, 9);
    }
    if (leader) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ===================================================================== host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 5-D view (sx*Cp, W/sx, sy, H/sy, N) of an NHWC buffer; sx = sy = 1 gives (Cp, W, 1, H, N).
static int encode_nhwc_view(CUtensorMap* tm, const void* base, int elem_bytes, int n, int h, int w,
                            int cp, int split, int box_c, int box_rows, bool swizzle) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return MASIC_EDRIVER;
  const int s = split ? 2 : 1;
  if (h % s || w % s) return MASIC_EINVAL;
  cuuint64_t dims[5] = {(cuuint64_t)s * cp, (cuuint64_t)(w / s), (cuuint64_t)s, (cuuint64_t)(h / s),
                        (cuuint64_t)n};
  const cuuint64_t e = elem_bytes;
  cuuint64_t strides[4] = {(cuuint64_t)s * cp * e, (cuuint64_t)w * cp * e,
                           (cuuint64_t)s * w * cp * e, (cuuint64_t)h * w * cp * e};
  cuuint32_t box[5] = {(cuuint32_t)box_c, (cuuint32_t)TILE_W, 1, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < 4; ++i)
    if (strides[i] % 16) return MASIC_EINVAL;
  if (reinterpret_cast<uintptr_t>(base) % 16) return MASIC_EINVAL;
  CUresult r = enc(tm, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                   5, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MASIC_OK : MASIC_EDRIVER;
}

static int encode_rows64(CUtensorMap* tm, const void* base, long rows, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return MASIC_EDRIVER;
  cuuint64_t dims[2] = {64, (cuuint64_t)rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MASIC_OK : MASIC_EDRIVER;
}
// MASIC_CONV_XFOLD4 input: a [N][H][W + MASIC_IMG_XPAD][16] bf16 image (pixel x at column x + MASIC_IMG_XOFF, pad
// columns zero) seen as OVERLAPPING 4-pixel windows: dims (64 elements, (W+XPAD)/2 windows 64 B apart, 2 row phases,
// H/2, N).  Window hx starts at column 2*hx, i.e. pixel 2*hx - 2.
static int encode_xfold4_view(CUtensorMap* tm, const void* base, int n, int h, int w, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return MASIC_EDRIVER;
  const cuuint64_t row_bytes = (cuuint64_t)(w + MASIC_IMG_XPAD) * 32;
  cuuint64_t dims[5] = {64, (cuuint64_t)(w + MASIC_IMG_XPAD) / 2 - 1, 2, (cuuint64_t)h / 2, (cuuint64_t)n};
  cuuint64_t strides[4] = {64, row_bytes, 2 * row_bytes, (cuuint64_t)h * row_bytes};
  cuuint32_t box[5] = {64, (cuuint32_t)TILE_W, 1, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (reinterpret_cast<uintptr_t>(base) % 16 || (w % 2) || (h % 2)) return MASIC_EINVAL;
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MASIC_OK : MASIC_EDRIVER;
}

// gamma [128][128] bf16 row-major, loaded as two (64 x 128-row) K-blocks
static int encode_gamma(CUtensorMap* tm, const void* base) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return MASIC_EDRIVER;
  cuuint64_t dims[2] = {128, 128};
  cuuint64_t strides[1] = {256};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MASIC_OK : MASIC_EDRIVER;
}

}  // namespace masic

using namespace masic;

struct MasicConvPlan {
  KParams kp;
  void* d_tables = nullptr;
  void* d_trace = nullptr;
  int smem_bytes = 0;
  int grid = 0;
  int total_work = 0;
  double flops = 0, hbm_bytes = 0;
};

namespace {

struct TapList {   // live taps of one strip: (a_row, kb_tap_index)
  std::vector<std::pair<int, int>> taps;
};

inline int floordiv2(int t) { return t >= 0 ? t / 2 : -((-t + 1) / 2); }

// Build the per-variant programs.  Returns the strip height (rows) via *rows_out.
int build_programs(const MasicConvDesc& d, std::vector<AOp>& aops, std::vector<BOp>& bops,
                   std::vector<MOp>& mops, Variant* var, int* n_var, int* rows_out) {
  const int k = d.ksize, ncb = (d.c_in + KBLK - 1) / KBLK;
  const uint32_t mask = d.tap_mask ? d.tap_mask : 0xFFFFFFFFu;
  const int nk_last = (d.c_in - (ncb - 1) * KBLK) / 16;
  int rows = TILE_H;
  *n_var = 0;

  auto emit_variant = [&](const std::vector<std::pair<AOp, TapList>>& strips_cb0, int out_p2,
                          int out_c0) -> int {
    Variant& v = var[*n_var];
    v.aops_off = (int)aops.size();
    v.bops_off = (int)bops.size();
    v.mops_off = (int)mops.size();
    v.out_p2 = out_p2;
    v.out_c0 = out_c0;
    bool first = true;
    for (int cb = 0; cb < ncb; ++cb) {
      const int nk = (cb == ncb - 1) ? nk_last : 4;
      for (const auto& st : strips_cb0) {
        if (st.second.taps.empty()) continue;
        AOp a = st.first;
        a.c0 = (int16_t)(a.c0 + cb * KBLK);
        aops.push_back(a);
        for (size_t i = 0; i < st.second.taps.size(); ++i) {
          BOp b;
          b.row0 = (st.second.taps[i].second * ncb + cb) * d.c_out_pad;
          bops.push_back(b);
          MOp m;
          m.a_row = (uint16_t)st.second.taps[i].first;
          m.nk = (uint8_t)nk;
          m.flags = M_NEW_B | M_REL_B;
          if (i == 0) m.flags |= M_NEW_A;
          if (i + 1 == st.second.taps.size()) m.flags |= M_REL_A;
          if (first) m.flags |= M_FIRST;
          first = false;
          mops.push_back(m);
        }
      }
    }
    v.n_aops = (int)aops.size() - v.aops_off;
    v.n_bops = (int)bops.size() - v.bops_off;
    v.n_mops = (int)mops.size() - v.mops_off;
    if (v.n_mops == 0) return MASIC_EINVAL;
    ++*n_var;
    return MASIC_OK;
  };

  if (d.kind == MASIC_CONV || d.kind == MASIC_DECONV_S2_SUBPIX) {
    const int kk = (d.kind == MASIC_DECONV_S2_SUBPIX) ? 3 : k;
    const int stride = (d.kind == MASIC_DECONV_S2_SUBPIX) ? 1 : d.stride;
    const int pad = kk / 2;
    std::vector<std::pair<AOp, TapList>> strips;
    if (stride == 1) {
      rows = TILE_H + kk - 1;
      for (int kx = 0; kx < kk; ++kx) {
        AOp a; a.c0 = (int16_t)d.in_coff; a.dx = (int16_t)(kx - pad); a.p2 = 0; a.dy = (int16_t)(-pad);
        TapList tl;
        for (int ky = 0; ky < kk; ++ky)
          if (mask & (1u << (ky * kk + kx))) tl.taps.push_back({ky, ky * kk + kx});
        strips.push_back({a, tl});
      }
    } else if (stride == 2) {
      // input row = 2*(oy + half) + py with t = ky - pad, half = floor(t/2), py = t - 2*half
      const int half_min = floordiv2(-pad), half_max = floordiv2(kk - 1 - pad);
      rows = TILE_H + (half_max - half_min);
      for (int kx = 0; kx < kk; ++kx) {
        const int tx = kx - pad, hx = floordiv2(tx), px = tx - 2 * hx;
        for (int py = 0; py < 2; ++py) {
          AOp a; a.c0 = (int16_t)(d.in_coff + px * d.in_cpitch); a.dx = (int16_t)hx; a.p2 = (int16_t)py;
          a.dy = (int16_t)half_min;
          TapList tl;
          for (int ky = 0; ky < kk; ++ky) {
            const int ty = ky - pad, hy = floordiv2(ty);
            if (ty - 2 * hy != py) continue;
            if (mask & (1u << (ky * kk + kx))) tl.taps.push_back({hy - half_min, ky * kk + kx});
          }
          strips.push_back({a, tl});
        }
      }
    } else {
      return MASIC_EINVAL;
    }
    int rc = emit_variant(strips, 0, 0);
    if (rc) return rc;
  } else if (d.kind == MASIC_CONV_XFOLD4) {
    // Overlapping-window view of a padded 16-channel image (encode_xfold4_view): window hx holds pixels
    // 2*hx-2 .. 2*hx+1 as 64 consecutive bf16.  Taps kx = 0..3 are one K=64 block read at window ox;
    // tap kx = 4 is K sub-block j=2 of window ox+1.
    if (k != 5 || d.c_in != 64 || d.in_cpitch != 16 || d.in_coff != 0 || d.stride != 2 || d.tap_mask) return MASIC_EINVAL;
    rows = TILE_H + 2;
    std::vector<std::pair<AOp, TapList>> strips;
    for (int grp = 0; grp < 2; ++grp)
      for (int py = 0; py < 2; ++py) {
        AOp a; a.c0 = 0; a.dx = (int16_t)(grp ? 1 : 0); a.p2 = (int16_t)py; a.dy = -1;
        TapList tl;
        for (int ky = py; ky < 5; ky += 2) tl.taps.push_back({(ky >> 1), ky * 2 + grp});
        strips.push_back({a, tl});
      }
    // emit by hand: group 1 ops use a single K=16 MMA
    Variant& v = var[*n_var];
    v.aops_off = (int)aops.size(); v.bops_off = (int)bops.size(); v.mops_off = (int)mops.size();
    v.out_p2 = 0; v.out_c0 = 0;
    bool first = true;
    for (size_t si = 0; si < strips.size(); ++si) {
      aops.push_back(strips[si].first);
      const auto& taps = strips[si].second.taps;
      for (size_t i = 0; i < taps.size(); ++i) {
        BOp b; b.row0 = taps[i].second * d.c_out_pad; bops.push_back(b);
        MOp m; m.a_row = (uint16_t)taps[i].first; m.nk = (uint8_t)(si >= 2 ? (1 | (2 << 4)) : 4);
        m.flags = M_NEW_B | M_REL_B;
        if (i == 0) m.flags |= M_NEW_A;
        if (i + 1 == taps.size()) m.flags |= M_REL_A;
        if (first) m.flags |= M_FIRST;
        first = false;
        mops.push_back(m);
      }
    }
    v.n_aops = (int)aops.size() - v.aops_off; v.n_bops = (int)bops.size() - v.bops_off;
    v.n_mops = (int)mops.size() - v.mops_off;
    ++*n_var;
  } else if (d.kind == MASIC_DECONV_S2) {
    if (k != 5) return MASIC_ENOSUP;
    // out[2q+py] gets taps ky = py, py+2, .. from input row q + dy, dy = 1 - (ky-py)/2
    rows = TILE_H + 2;
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        std::vector<std::pair<AOp, TapList>> strips;
        for (int kx = px; kx < 5; kx += 2) {
          const int dx = 1 - (kx - px) / 2;
          AOp a; a.c0 = (int16_t)d.in_coff; a.dx = (int16_t)dx; a.p2 = 0; a.dy = -1;
          TapList tl;
          for (int ky = py; ky < 5; ky += 2) {
            const int dy = 1 - (ky - py) / 2;
            if (mask & (1u << (ky * 5 + kx))) tl.taps.push_back({dy + 1, ky * 5 + kx});
          }
          strips.push_back({a, tl});
        }
        int rc = emit_variant(strips, py, px * d.out_cpitch);
        if (rc) return rc;
      }
  } else {
    return MASIC_EINVAL;
  }
  *rows_out = rows;
  return MASIC_OK;
}

}  // namespace

extern "C" int masic_conv_plan_create(const MasicConvDesc* dp, MasicConvPlan** plan_out) {
  if (!dp || !plan_out) return MASIC_EINVAL;
  const MasicConvDesc& d = *dp;
  *plan_out = nullptr;
  if (d.c_in <= 0 || d.c_in % 16) return MASIC_EINVAL;
  if (d.n_tile % 16 || d.n_tile < 16 || d.n_tile > 256) return MASIC_EINVAL;
  if (d.c_out_pad % d.n_tile || d.c_out > d.c_out_pad) return MASIC_EINVAL;
  if (d.c_out_pad / d.n_tile > 32) return MASIC_EINVAL;
  if (d.in_cpitch % 8 || d.in_coff % 8 || d.out_coff % 8) return MASIC_EINVAL;
  if (d.kind == MASIC_CONV_XFOLD4 && d.c_in != 64) return MASIC_EINVAL;
  if (d.out_cpitch % (d.out_fp32 ? 4 : 8)) return MASIC_EINVAL;
  if (d.ksize != 1 && d.ksize != 3 && d.ksize != 5) return MASIC_EINVAL;
  if (d.gdn && (d.n_tile != 128 || d.c_out != 128 || !d.gamma_packed || !d.beta || !d.bias)) return MASIC_EINVAL;
  if ((d.kind == MASIC_CONV || d.kind == MASIC_CONV_XFOLD4) && d.stride == 2 && (d.h_in % 2 || d.w_in % 2))
    return MASIC_EINVAL;

  MasicConvPlan* pl = new MasicConvPlan();
  KParams& kp = pl->kp;
  memset(&kp, 0, sizeof(kp));

  std::vector<AOp> aops; std::vector<BOp> bops; std::vector<MOp> mops;
  int rows = 0;
  int rc = build_programs(d, aops, bops, mops, kp.var, &kp.n_var, &rows);
  if (rc) { delete pl; return rc; }

  // geometry of the tile grid (output positions for conv, input positions for deconv)
  int gh, gw, out_h, out_w, out_split = 0;
  if (d.kind == MASIC_CONV || d.kind == MASIC_CONV_XFOLD4) {
    gh = d.stride == 2 ? d.h_in / 2 : d.h_in;
    gw = d.stride == 2 ? d.w_in / 2 : d.w_in;
    out_h = gh; out_w = gw;
  } else if (d.kind == MASIC_DECONV_S2) {
    gh = d.h_in; gw = d.w_in; out_h = 2 * gh; out_w = 2 * gw; out_split = 1;
  } else {
    gh = d.h_in; gw = d.w_in; out_h = gh; out_w = gw;
  }
  kp.tiles_x = (gw + TILE_W - 1) / TILE_W;
  kp.tiles_y = (gh + TILE_H - 1) / TILE_H;
  kp.n_img = d.n;
  kp.n_tile = d.n_tile;
  kp.n_ntiles = d.c_out_pad / d.n_tile;
  kp.idesc = umma_idesc_bf16(d.n_tile);

  // staging block: 128-B swizzled rows when the n-tile is wide enough, else one narrow block
  const int esz = d.out_fp32 ? 4 : 2;
  const int full_blk_ch = 128 / esz;
  if (d.n_tile % full_blk_ch == 0) { kp.blk_ch = full_blk_ch; kp.blk_pitch = 128; }
  else if (d.n_tile * esz < 128) { kp.blk_ch = d.n_tile; kp.blk_pitch = d.n_tile * esz; }
  else { delete pl; return MASIC_EINVAL; }

  // shared memory carve-up
  kp.a_stage_bytes = rows * 1024;
  kp.b_stage_bytes = d.n_tile * 128;
  // staging: with GDN one block per epilogue group (the pair is the A2 operand); otherwise two per group when
  // the minimum rings (2 A + 2 B stages per stream) still fit, else one
  int n_stage_blk = d.gdn ? 2 : 4;
  int fixed = (d.gdn ? 2 * STAGE_BLK_BYTES : 0) + n_stage_blk * STAGE_BLK_BYTES + MISC_BYTES + 1024 /*align*/;
  int budget = (227 * 1024 - fixed) / NUM_STREAMS;            // ring bytes per stream
  int sa = 2, sb = 2;
  if (n_stage_blk == 4 && sa * kp.a_stage_bytes + sb * kp.b_stage_bytes > budget) {
    n_stage_blk = 2;
    fixed -= 2 * STAGE_BLK_BYTES;
    budget = (227 * 1024 - fixed) / NUM_STREAMS;
  }
  kp.stage_per_group = n_stage_blk / 2;
  if (sa * kp.a_stage_bytes + sb * kp.b_stage_bytes > budget) { delete pl; return MASIC_EINVAL; }
  // deepen the rings with what is left (B first: a B stage is consumed by a single op)
  for (bool grew = true; grew;) {
    grew = false;
    if (sb < 4 && sa * kp.a_stage_bytes + (sb + 1) * kp.b_stage_bytes <= budget) { ++sb; grew = true; }
    if (sa < 3 && (sa + 1) * kp.a_stage_bytes + sb * kp.b_stage_bytes <= budget) { ++sa; grew = true; }
  }
  kp.a_stages = sa; kp.b_stages = sb;
  kp.smem_b_off = NUM_STREAMS * sa * kp.a_stage_bytes;
  kp.smem_g_off = kp.smem_b_off + NUM_STREAMS * sb * kp.b_stage_bytes;
  kp.smem_stage_off = kp.smem_g_off + (d.gdn ? 2 * STAGE_BLK_BYTES : 0);
  kp.smem_misc_off = kp.smem_stage_off + n_stage_blk * STAGE_BLK_BYTES;
  pl->smem_bytes = kp.smem_misc_off + MISC_BYTES + 1024;
  if (pl->smem_bytes < 120 * 1024) pl->smem_bytes = 120 * 1024;   // keep 1 CTA/SM: 512 TMEM cols each

  // tensor maps
  const int split_in = ((d.kind == MASIC_CONV || d.kind == MASIC_CONV_XFOLD4) && d.stride == 2) ? 1 : 0;
  if (d.kind == MASIC_CONV_XFOLD4) rc = encode_xfold4_view(&kp.tmA, d.in, d.n, d.h_in, d.w_in, rows);
  else rc = encode_nhwc_view(&kp.tmA, d.in, 2, d.n, d.h_in, d.w_in, d.in_cpitch, split_in, KBLK, rows, true);
  const int ktaps = (d.kind == MASIC_DECONV_S2_SUBPIX) ? 9 : (d.kind == MASIC_CONV_XFOLD4 ? 10 : d.ksize * d.ksize);
  const int ncb = (d.kind == MASIC_CONV_XFOLD4) ? 1 : (d.c_in + KBLK - 1) / KBLK;
  if (!rc) rc = encode_rows64(&kp.tmB, d.w_packed, (long)ktaps * ncb * d.c_out_pad, d.n_tile);
  if (!rc) rc = encode_nhwc_view(&kp.tmO, d.out, esz, d.n, out_h, out_w, d.out_cpitch, out_split,
                                 kp.blk_ch, TILE_H, kp.blk_pitch == 128);
  if (!rc && d.gdn) rc = encode_gamma(&kp.tmG, d.gamma_packed);
  if (rc) { delete pl; return rc; }

  // the programs travel in the kernel parameters (constant bank)
  if (aops.size() > MAX_AOPS || bops.size() > MAX_BOPS || mops.size() > MAX_MOPS) { delete pl; return MASIC_ENOSUP; }
  memcpy(kp.aops, aops.data(), aops.size() * sizeof(AOp));
  memcpy(kp.bops, bops.data(), bops.size() * sizeof(BOp));
  memcpy(kp.mops, mops.data(), mops.size() * sizeof(MOp));
  cudaError_t ce = cudaSuccess;

  kp.bias = d.bias; kp.beta = d.beta; kp.gdn = d.gdn; kp.out_fp32 = d.out_fp32;
  memcpy(kp.act, d.act, sizeof(kp.act));
  kp.out_coff = d.out_coff;
  kp.rowscale = d.rowscale; kp.rs_stride = d.rs_stride; kp.rs_off = d.rs_off;
  kp.rs_H = gh; kp.rs_W = gw;
  { const char* e = getenv("MASIC_CONV_DEBUG"); kp.debug = e ? atoi(e) : 0; }
  if (getenv("MASIC_CONV_TRACE")) {
    if (cudaMalloc(&pl->d_trace, 64 * 16 * sizeof(long long)) == cudaSuccess) {
      cudaMemset(pl->d_trace, 0, 64 * 16 * sizeof(long long));
      kp.trace = static_cast<long long*>(pl->d_trace);
    }
  }

  pl->total_work = kp.n_img * kp.tiles_y * kp.tiles_x * kp.n_var * kp.n_ntiles;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  pl->grid = pl->total_work < sms ? pl->total_work : sms;

  // useful work: live taps only, real channels only
  {
    const uint32_t mask = d.tap_mask ? d.tap_mask : 0xFFFFFFFFu;
    int live = 0;
    for (int i = 0; i < d.ksize * d.ksize; ++i) live += (mask >> i) & 1;
    const double out_pos = (double)d.n * out_h * out_w;
    double macs;
    const double co_real = (d.kind == MASIC_DECONV_S2_SUBPIX) ? d.c_out / 4 : d.c_out;
    if (d.kind == MASIC_CONV_XFOLD4) macs = out_pos * 25.0 * 3.0 * d.c_out;       // the real 3-channel 5x5
    else if (d.kind == MASIC_CONV) macs = out_pos * live * d.c_in * d.c_out;
    else macs = (double)d.n * d.h_in * d.w_in * 25.0 * d.c_in * co_real;   // each input feeds 25 taps
    pl->flops = 2.0 * macs + (d.gdn ? 2.0 * out_pos * 128.0 * 128.0 : 0.0);
    const double out_ch = (d.kind == MASIC_DECONV_S2_SUBPIX) ? d.c_out_pad : d.c_out;
    pl->hbm_bytes = (double)d.n * d.h_in * d.w_in * (d.kind == MASIC_CONV_XFOLD4 ? 16 : d.c_in) * 2.0 + out_pos * out_ch * esz +
                    (double)ktaps * ncb * d.c_out_pad * 128.0;
  }

  static bool attr_set = false;
  if (!attr_set) {
    ce = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (ce != cudaSuccess) { delete pl; return (int)ce; }
    attr_set = true;
  }
  *plan_out = pl;
  return MASIC_OK;
}

extern "C" int masic_conv_plan_launch(const MasicConvPlan* pl, void* stream) {
  if (!pl) return MASIC_EINVAL;
  conv_tc_kernel<<<pl->grid, NUM_THREADS, pl->smem_bytes, static_cast<cudaStream_t>(stream)>>>(pl->kp);
  return (int)cudaGetLastError();
}

extern "C" void masic_conv_plan_destroy(MasicConvPlan* pl) {
  if (!pl) return;
  if (pl->d_tables) cudaFree(pl->d_tables);
  if (pl->d_trace) cudaFree(pl->d_trace);
  delete pl;
}

// Debug: copy out the clock64() stamps of CTA 0 (64 tiles x 16 slots); needs MASIC_CONV_TRACE=1 at plan creation.
extern "C" int masic_conv_plan_trace(const MasicConvPlan* pl, long long* out_host) {
  if (!pl || !out_host) return MASIC_EINVAL;
  if (!pl->d_trace) return MASIC_ENOSUP;
  return (int)cudaMemcpy(out_host, pl->d_trace, 64 * 16 * sizeof(long long), cudaMemcpyDeviceToHost);
}

extern "C" int masic_conv_plan_info(const MasicConvPlan* pl, double* flops, double* hbm_bytes,
                                    int* n_work_items, int* smem_bytes) {
  if (!pl) return MASIC_EINVAL;
  if (flops) *flops = pl->flops;
  if (hbm_bytes) *hbm_bytes = pl->hbm_bytes;
  if (n_work_items) *n_work_items = pl->total_work;
  if (smem_bytes) *smem_bytes = pl->smem_bytes;
  return MASIC_OK;
}
