// gemm.cu — persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = A[M,K] · B[N,K]^T   (bf16 operands, fp32 accumulate in TMEM)
//
// This one kernel serves every dense contraction on the Flux.2 denoising path:
//   * the DiT linears (reference: Flux2Attention.swift:115-123,185-189; Flux2ParallelAttention.swift:80-87,122;
//     Flux2FeedForward.swift:59-67,102-108; Flux2Transformer.swift:137-138,324),
//   * the VAE decoder 3x3 / 1x1 convolutions as implicit GEMM over NHWC (reference: VAEDecoder.swift:91-121,
//     ResnetBlock.swift:168-186,229-253) — the A operand is fetched by a 4-D TMA box whose out-of-bounds fill
//     implements the zero padding,
// with the elementwise work that follows each of them fused into the TMEM->register epilogue
// (gate·y + residual, SwiGLU, QK-RMSNorm + RoPE, bias, resnet shortcut add).
//
// Structure per CTA (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one thread) + TMEM allocator,
// warps 2..5 = epilogue (one TMEM lane quarter each). smem ring of kStages {A 128x64, B BNx64} tiles in the
// 128B-swizzled K-major layout; two accumulator stages in TMEM so the epilogue of tile i overlaps the MMAs of
// tile i+1. With kCtaGroup == 2 a CTA pair (cluster of 2) computes a 256xBN tile with cta_group::2 MMAs: each CTA
// loads its 128 rows of A and half of B, the leader CTA issues, commits are multicast to both CTAs.
#include "gemm.cuh"
#include "ptx.cuh"
#include "quant_dev.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <string>
#include <type_traits>

namespace f2b {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int CONV_TW = 16;  // spatial patch of one M tile in conv mode 1 (one TMA box per tap): 8 rows x 16 cols = 128 pixels
static constexpr int CONV_TH = 8;
// conv mode 2 (3x3, stride 1): the M tile is 16 rows x 8 cols and its 18 x 10-pixel halo (rows y0-1 .. y0+16, cols x0-1 .. x0+8)
// is loaded ONCE per 64-channel block and serves all nine taps: the A operand of tap (ky, kx) is the window of the halo that
// starts at halo row ky, column kx — a UMMA descriptor with an 8-row-group stride of one halo row (10 pixels x 128 B = 1280 B)
// and a start address off the 1024 B swizzle period. The descriptor's base-offset field stays 0: the tensor core applies the
// 128 B swizzle to absolute shared-memory address bits (as TMA does when it writes the tile), so neither a window that starts
// part-way into the pattern nor a group stride that is not a multiple of it needs a correction (measured: base offset kx or
// 8 - kx gives wrong taps for kx != 0, 0 is exact — profiles/r02_halo_base_offset.md). L2 -> shared-memory traffic for A drops
// from 9 x 16 KB to 22.5 KB per 64-channel block.
#ifndef F2B_HALO_W
#define F2B_HALO_W 10
#endif
static constexpr int HALO_W = F2B_HALO_W, HALO_H = 18;   // 10 = exactly the columns the nine windows touch (16: 8-row groups 2048 B apart)
static constexpr int HALO_TW = 8, HALO_TH = 16;
static constexpr int HALO_BYTES = HALO_W * HALO_H * 128;   // 22.5 KB of TMA payload per stage
static constexpr int HALO_STRIDE = (HALO_BYTES + 1023) / 1024 * 1024;   // stages start on the 1024 B swizzle repeat
static constexpr int HALO_STAGES = 3;
static constexpr int CONV_BIAS_MAX = 512;   // convolution bias vectors up to this many channels are staged in shared memory

struct KParams {
  int M, N, K;
  int num_m_units, num_n_blks, num_kb;
  // conv
  int taps, H, W, Cin, batch, tiles_x, tiles_y, kb_per_tap;  // H, W: OUTPUT extent
  int stride, tap_off;  // input coordinate of tap (ky, kx) for output (y, x): (y * stride + ky + tap_off, x * stride + kx + tap_off)
  int wres;             // conv mode 2: the layer's weights for this CTA's N rows fit the operand ring and stay resident in shared memory (loaded once)
  int up2;              // conv mode 2 only: nearest-2x upsample folded into the 3x3 convolution (four 2x2 phase kernels, see gemm.cuh)
  Epilogue epi;
  // block-scaled kinds: scale factors of A [ceil(M/128)][sfa_ld][512 B], of B [N/128][sfb_ld][512 B]; one 512 B block =
  // 128 rows x 4 scale bytes in the tcgen05 layout (quant.cuh); *_ld = blocks per 128-row block (the full K extent)
  const uint8_t* sfa;
  const uint8_t* sfb;
  int sfa_ld, sfb_ld;
  int split_units;  // two-problem launch: M units below this one read B through the second weight map (passed in the tmSFB slot)
  // W-only quantized weights (WQ kernels): B arrives as MLX's packed bytes and is dequantized to the 16-bit operand by four
  // extra warps on its way into the swizzled B stage. wq_mode = flux2b_quant (1 qint8, 2 int4, 3 mxfp8, 4 mxfp4, 5 nvfp4);
  // scales / biases row-major [N, wq_sb_ld groups] in their checkpoint type (affine: f16 / bf16, wq_sb_bf16; mx: one byte)
  int wq_mode, wq_sb_bf16, wq_sb_ld, wq_vec;   // wq_vec: 16 B loads of 8 (4 for nvfp4) k-blocks of scales are aligned
  const uint8_t* wq_s;  const uint8_t* wq_b;   // main problem
  const uint8_t* wq_s_lo; const uint8_t* wq_b_lo;  // rows below split_units (two-problem launch)
  int n_off;  // absolute output column of this launch's first B row (staged W-only chunks), 0 otherwise
  int n_fast; // tile order: 0 = consecutive tiles walk M (share a weight tile), 1 = consecutive tiles walk N (share an activation tile) — see launch_cfg
  int dbg;  // FLUX2B_GEMM_TIMELINE=1: cluster 0 prints where its producer / issuer / epilogue warps waited (debug aid)
};

// MXK: 0 = 16-bit operands, 1 = mxfp8 (E4M3, E8M0 scale / 32), 2 = mxfp4 (E2M1, E8M0 / 32), 3 = nvfp4 (E2M1, E4M3 / 16)
// WQ kernels: a ring of packed-weight slots (TMA destination, 64 B-/32 B-swizzled rows of 64 / 32 bytes = 64 K elements) beside
// the operand stages; 128 more threads (warps 6..9) turn slot -> 16-bit B stage
static constexpr int WQ_PSTAGES = 4;
static constexpr int WQ_DQ_WARPS = 8;   // two per scheduler: one warp per scheduler cannot hide its own dependent-issue latency (measured 2x)
template <int BN, int CG, int MXK = 0, int WQ = 0, int HALO = 0>
struct Cfg {
  static constexpr int B_ROWS = BN / CG;
  static constexpr int A_BYTES = BM * BK * 2;      // 128 rows x 128 B: 64 bf16, 128 fp8 or 256 fp4 elements along K
  // halo convolutions: one B stage carries the weight tiles of a GROUP of taps (one kernel row: 3 taps; folded upsample: 2), so
  // that one barrier round trip of the issuer covers 3 x 4 MMAs. With one tap per stage the issuer's ~160 instructions per
  // stage (waits, descriptor arithmetic, commit) cost ~560 clk against 96 - 192 clk of MMA work for the 96-wide VAE tiles
  // (ncu source page, profiles/r02_conv_halo.md). Single-CTA tiles wider than 128 keep one tap per stage (shared memory).
  static constexpr int TG = HALO ? (B_ROWS <= 128 ? 3 : 1) : 1;
  static constexpr int B_TAP_BYTES = B_ROWS * BK * 2;
  static constexpr int B_BYTES = TG * B_TAP_BYTES;
  static constexpr int KB_ELEMS = MXK == 0 ? BK : MXK == 1 ? 128 : 256;   // K elements per pipeline stage
  // 512 B scale-factor blocks (128 rows x 4 scale bytes) per 128 operand rows per stage
  static constexpr int SFPK = MXK == 0 ? 0 : MXK == 1 ? 1 : MXK == 2 ? 2 : 4;
  static constexpr int SFA_BYTES = 512 * SFPK;
  static constexpr int SFB_BYTES = (BN / 128) * 512 * SFPK;
  static constexpr int SF_BYTES = SFA_BYTES + SFB_BYTES;
  static constexpr int A_STAGE = HALO ? 0 : A_BYTES;                  // halo convolutions keep A in the halo ring, not in the stages
  static constexpr int STAGE_BYTES = A_STAGE + B_BYTES + SF_BYTES;
  static constexpr int PSLOT_BYTES = WQ ? B_ROWS * 64 : 0;            // sized for 8-bit codes; 4-bit modes use half of a slot
  static constexpr int PRING_BYTES = WQ_PSTAGES * PSLOT_BYTES;
  static constexpr int HRING_BYTES = HALO ? HALO_STAGES * HALO_STRIDE : 0;
  static constexpr int BUDGET = HALO ? 220 * 1024 : 196 * 1024;
  static constexpr int STAGES_RAW = (BUDGET - PRING_BYTES - HRING_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  // accumulator stages in TMEM: two, except the 256-wide block-scaled tile, whose scale factors need columns too
  static constexpr int ACC_STAGES = (MXK && BN == 256) ? 1 : 2;
  static constexpr int SF_COL0 = ACC_STAGES * BN;  // SFA: 4 columns per block, SFB: 4 per block per 128 rows, behind them
  static constexpr int SF_COLS = MXK ? 4 * SFPK * (1 + BN / 128) : 0;
  static constexpr int COLS_NEEDED = ACC_STAGES * BN + SF_COLS;
  static constexpr int TMEM_COLS = (COLS_NEEDED <= 32) ? 32 : (COLS_NEEDED <= 64) ? 64 : (COLS_NEEDED <= 128) ? 128 : (COLS_NEEDED <= 256) ? 256 : 512;
  static constexpr int BAR_BYTES = 512;
  static constexpr int BIAS_BYTES = CONV_BIAS_MAX * 4;   // convolutions: the bias vector, staged once per CTA
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + PRING_BYTES + HRING_BYTES + BAR_BYTES + BIAS_BYTES + 1024;  // +1024 for manual alignment
  static_assert(STAGES >= 2, "operand ring needs two stages");
  static constexpr int THREADS = WQ ? 192 + 32 * WQ_DQ_WARPS : 192;
};

// ------------------------------------------------------------------------------------------------ epilogue
__device__ __forceinline__ uint32_t pk2(float a, float b, int f16) {
  if (f16) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  return pack_bf16x2(a, b);
}
__device__ __forceinline__ float2 upk2(uint32_t u, int f16) {
  if (f16) return __half22float2(*reinterpret_cast<__half2*>(&u));
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}
// 32 consecutive 16-bit results of row `grow` starting at output column `col` -> block-scaled bytes + group scale(s)
template <int KIND>
__device__ __forceinline__ void store_quantised32(const Epilogue& e, const uint32_t (&pk)[16], bool row_ok, int64_t grow, int col) {
  const MxOut& m = e.mxo;
  constexpr int GROUP = KIND == 3 ? 16 : 32;
  uint8_t* sp = m.sf + sf_offset(grow, m.g0 + col / GROUP, m.sf_ld);
  if (!row_ok) {  // padding row of the last 128-row block: scale 1.0
    if constexpr (KIND == 3) *reinterpret_cast<uint16_t*>(sp) = 0x3838; else *sp = 127;
    return;
  }
  float f[32];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float2 t = upk2(pk[j], e.f16);
    f[2 * j] = t.x; f[2 * j + 1] = t.y;
  }
  if constexpr (KIND == 3) {
    uint32_t w[4], sb2 = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float amax = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) amax = fmaxf(amax, fabsf(f[16 * h + j]));
      const MxScale sc = mx_scale<3>(amax);
#pragma unroll
      for (int q = 0; q < 2; ++q)
        w[2 * h + q] = mx_pack4<3>(f[16 * h + 8 * q], f[16 * h + 8 * q + 1], f[16 * h + 8 * q + 2], f[16 * h + 8 * q + 3], sc.mul) |
                       (mx_pack4<3>(f[16 * h + 8 * q + 4], f[16 * h + 8 * q + 5], f[16 * h + 8 * q + 6], f[16 * h + 8 * q + 7], sc.mul) << 16);
      sb2 |= sc.sb << (8 * h);
    }
    *reinterpret_cast<uint4*>(m.q + grow * m.ldq + (col >> 1)) = make_uint4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<uint16_t*>(sp) = (uint16_t)sb2;
  } else {
    float amax = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) amax = fmaxf(amax, fabsf(f[j]));
    const MxScale sc = mx_scale<KIND>(amax);
    if constexpr (KIND == 2) {
      uint32_t w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        w[q] = mx_pack4<2>(f[8 * q], f[8 * q + 1], f[8 * q + 2], f[8 * q + 3], sc.mul) |
               (mx_pack4<2>(f[8 * q + 4], f[8 * q + 5], f[8 * q + 6], f[8 * q + 7], sc.mul) << 16);
      *reinterpret_cast<uint4*>(m.q + grow * m.ldq + (col >> 1)) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
      uint32_t w[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) w[q] = mx_pack4<1>(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3], sc.mul);
      uint4* dst = reinterpret_cast<uint4*>(m.q + grow * m.ldq + col);
      dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
      dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
    *sp = (uint8_t)sc.sb;
  }
}

template <int BN>
__device__ __forceinline__ void epilogue_tile(const KParams& p, uint32_t tmem_acc, int quarter, int lane, bool row_ok,
                                              int64_t grow, int n0, uint32_t sbias = 0) {
  const Epilogue& e = p.epi;
  const uint32_t tq = tmem_acc + ((uint32_t)(quarter * 32) << 16);
  uint32_t v[32];

  if (e.mode == EPI_SWIGLU) {
    // tile columns: [0, BN/2) gate, [BN/2, BN) value  ->  BN/2 outputs at column (n0/2 + j)
    constexpr int HALF = BN / 2;
    uint32_t u[32];
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(e.out);
#pragma unroll 1
    for (int c = 0; c < HALF / 32; ++c) {
      tmem_ld_32x32(tq + c * 32, v);
      tmem_ld_32x32(tq + HALF + c * 32, u);
      tmem_ld_wait();
      const int oc = n0 / 2 + c * 32;
      if ((row_ok || e.mxo.kind) && oc < p.N / 2) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float g0 = __uint_as_float(v[2 * j]), g1 = __uint_as_float(v[2 * j + 1]);
          float u0 = __uint_as_float(u[2 * j]), u1 = __uint_as_float(u[2 * j + 1]);
          pk[j] = pk2(silu_f(g0) * u0, silu_f(g1) * u1, e.f16);
        }
        if (e.mxo.kind == 0) {
          uint4* dst = reinterpret_cast<uint4*>(out + grow * e.ldo + oc);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        } else if (e.mxo.kind == 1) {
          store_quantised32<1>(e, pk, row_ok, grow, oc);
        } else if (e.mxo.kind == 2) {
          store_quantised32<2>(e, pk, row_ok, grow, oc);
        } else {
          store_quantised32<3>(e, pk, row_ok, grow, oc);
        }
      }
    }
    return;
  }

  if (e.mode == EPI_QKV_ROPE) {
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(e.out);
#pragma unroll 1
    for (int h = 0; h < BN / 128; ++h) {
      const int hc = n0 + h * 128;  // first column of this 128-wide head slab
      if (hc >= p.N) break;
      const int kc0 = e.k_col0 ? e.k_col0 : e.dmodel, vc0 = e.v_col0 ? e.v_col0 : 2 * e.dmodel;
      const bool is_v = hc >= vc0;
      const bool lo = grow < e.split_row;   // uniform over the tile (split_row is a multiple of the M tile)
      const float* nw = (hc < kc0) ? (lo ? e.norm_q_lo : e.norm_q) : (lo ? e.norm_k_lo : e.norm_k);
      float rstd = 1.0f;
      if (!is_v && nw) {
        float ss = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          tmem_ld_32x32(tq + h * 128 + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(v[j]);
            ss = fmaf(x, x, ss);
          }
        }
        rstd = rsqrtf(ss * (1.0f / 128.0f) + e.eps);
      }
      if (e.rope_half && !is_v) {
        // rotate-half RoPE of the text encoders (MLXFast.RoPE traditional = false; Qwen3Attention.swift:29-35): element j < 64
        // pairs with j + 64: (x1, x2) -> (x1 c - x2 s, x2 c + x1 s), angle table column j. Optional per-head RMSNorm first.
        uint32_t u[32];
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          tmem_ld_32x32(tq + h * 128 + c * 32, v);
          tmem_ld_32x32(tq + h * 128 + 64 + c * 32, u);
          tmem_ld_wait();
          if (!row_ok) continue;
          const float4* cs = reinterpret_cast<const float4*>(e.cos + grow * 128 + c * 32);
          const float4* sn = reinterpret_cast<const float4*>(e.sin + grow * 128 + c * 32);
          uint32_t p1[16], p2[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 cc = __ldg(cs + j), s4 = __ldg(sn + j);
            float4 w1 = make_float4(1.f, 1.f, 1.f, 1.f), w2 = w1;
            if (nw) { w1 = __ldg(reinterpret_cast<const float4*>(nw + c * 32) + j); w2 = __ldg(reinterpret_cast<const float4*>(nw + 64 + c * 32) + j); }
            const float a0 = __uint_as_float(v[4 * j + 0]) * rstd * w1.x, b0 = __uint_as_float(u[4 * j + 0]) * rstd * w2.x;
            const float a1 = __uint_as_float(v[4 * j + 1]) * rstd * w1.y, b1 = __uint_as_float(u[4 * j + 1]) * rstd * w2.y;
            const float a2 = __uint_as_float(v[4 * j + 2]) * rstd * w1.z, b2 = __uint_as_float(u[4 * j + 2]) * rstd * w2.z;
            const float a3 = __uint_as_float(v[4 * j + 3]) * rstd * w1.w, b3 = __uint_as_float(u[4 * j + 3]) * rstd * w2.w;
            p1[2 * j] = pk2(a0 * cc.x - b0 * s4.x, a1 * cc.y - b1 * s4.y, e.f16);
            p1[2 * j + 1] = pk2(a2 * cc.z - b2 * s4.z, a3 * cc.w - b3 * s4.w, e.f16);
            p2[2 * j] = pk2(b0 * cc.x + a0 * s4.x, b1 * cc.y + a1 * s4.y, e.f16);
            p2[2 * j + 1] = pk2(b2 * cc.z + a2 * s4.z, b3 * cc.w + a3 * s4.w, e.f16);
          }
          uint4* d1 = reinterpret_cast<uint4*>(out + grow * e.ldo + hc + c * 32);
          uint4* d2 = reinterpret_cast<uint4*>(out + grow * e.ldo + hc + 64 + c * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            d1[j] = make_uint4(p1[4 * j], p1[4 * j + 1], p1[4 * j + 2], p1[4 * j + 3]);
            d2[j] = make_uint4(p2[4 * j], p2[4 * j + 1], p2[4 * j + 2], p2[4 * j + 3]);
          }
        }
        continue;
      }
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        tmem_ld_32x32(tq + h * 128 + c * 32, v);
        tmem_ld_wait();
        if (!row_ok) continue;
        uint32_t pk[16];
        if (is_v) {
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pk2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]), e.f16);
        } else {
          const float4* cs = reinterpret_cast<const float4*>(e.cos + grow * 128 + c * 32);
          const float4* sn = reinterpret_cast<const float4*>(e.sin + grow * 128 + c * 32);
          const float4* w4 = reinterpret_cast<const float4*>(nw + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 cc = __ldg(cs + j), s4 = __ldg(sn + j), ww = nw ? __ldg(w4 + j) : make_float4(1.f, 1.f, 1.f, 1.f);
            float x0 = __uint_as_float(v[4 * j + 0]) * rstd * ww.x;
            float x1 = __uint_as_float(v[4 * j + 1]) * rstd * ww.y;
            float x2 = __uint_as_float(v[4 * j + 2]) * rstd * ww.z;
            float x3 = __uint_as_float(v[4 * j + 3]) * rstd * ww.w;
            // interleaved pairs (Flux2Attention.swift:225-226,452-460): (x0,x1) -> (x0 c - x1 s, x1 c + x0 s)
            pk[2 * j] = pk2(x0 * cc.x - x1 * s4.x, x1 * cc.y + x0 * s4.y, e.f16);
            pk[2 * j + 1] = pk2(x2 * cc.z - x3 * s4.z, x3 * cc.w + x2 * s4.w, e.f16);
          }
        }
        uint4* dst;
        if (e.sp_hp > 0) {
          const int which = hc / e.dmodel;                    // 0 q, 1 k, 2 v
          const int hh = (hc - which * e.dmodel) >> 7;        // head index
          const int dest = hh / e.sp_hp, hl = hh - dest * e.sp_hp;
          dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.sp_base[dest]) + grow * e.ldo +
                                         (which * e.sp_hp + hl) * 128 + c * 32);
        } else {
          dst = reinterpret_cast<uint4*>(out + grow * e.ldo + hc + c * 32);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      }
    }
    return;
  }

  // Software-pipelined over 32-column chunks: the TMEM load of chunk c + 1 (and, for convolutions, its bias from shared memory) is
  // in flight while chunk c is converted and stored. One chunk at a time exposed a TMEM round trip (~500 clk beside a running
  // MMA, which has priority on TMEM) plus a shared-memory round trip per chunk: 1.9 k clk per chunk on the 96-wide VAE tiles,
  // where the epilogue, not the tensor pipe, set the tile period (ncu source page, profiles/r02_conv_halo.md).
  constexpr int NC = BN / 32;
  uint32_t vb[32];
  uint4 bq[2][8];
  auto prefetch_bias = [&](int c, uint4 (&b)[8]) {
    if (sbias) {
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = lds128(sbias + (uint32_t)(n0 + c * 32 + 4 * j) * 4);
    }
  };
  auto process = [&](int c, const uint32_t (&vv)[32], const uint4 (&b)[8]) {
    const int col = n0 + c * 32;
    if (!row_ok || col >= p.N) return;
    float a[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) a[j] = __uint_as_float(vv[j]);
    const bool full = (col + 32 <= p.N);
    if (sbias) {   // convolutions: bias staged in shared memory, zero beyond N
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 b4 = b[j];
        a[4 * j] += __uint_as_float(b4.x); a[4 * j + 1] += __uint_as_float(b4.y);
        a[4 * j + 2] += __uint_as_float(b4.z); a[4 * j + 3] += __uint_as_float(b4.w);
      }
    } else if (e.bias) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (full || col + j < p.N) a[j] += __ldg(e.bias + col + j);
    }
    if (e.mode == EPI_GATE_RES) {
      float* out = reinterpret_cast<float*>(e.out) + grow * e.ldo + col;
      const float* res = e.res + grow * e.ldr + col;
      const float* gate = (grow < e.split_row) ? e.gate_lo : e.gate;
      if (full) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 r = *reinterpret_cast<const float4*>(res + 4 * j);
          float4 g = __ldg(reinterpret_cast<const float4*>(gate + col + 4 * j));
          float4 o;
          o.x = fmaf(g.x, a[4 * j + 0], r.x);
          o.y = fmaf(g.y, a[4 * j + 1], r.y);
          o.z = fmaf(g.z, a[4 * j + 2], r.z);
          o.w = fmaf(g.w, a[4 * j + 3], r.w);
          *reinterpret_cast<float4*>(out + 4 * j) = o;
        }
      } else {
        for (int j = 0; j < 32 && col + j < p.N; ++j) out[j] = fmaf(__ldg(gate + col + j), a[j], res[j]);
      }
    } else if (e.mode == EPI_F32) {
      float* out = reinterpret_cast<float*>(e.out) + grow * e.ldo + col;
      if (full) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(out + 4 * j) = make_float4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
      } else {
        for (int j = 0; j < 32 && col + j < p.N; ++j) out[j] = a[j];
      }
    } else {  // EPI_BF16
      __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(e.out) + grow * e.ldo + col;
      if (e.res16) {
        const uint16_t* r16 = reinterpret_cast<const uint16_t*>(e.res16) + grow * e.ldr + col;
        if (full) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 q = *reinterpret_cast<const uint4*>(r16 + 8 * j);
            const uint32_t h2[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float2 f = upk2(h2[k], e.f16);
              a[8 * j + 2 * k] += f.x;
              a[8 * j + 2 * k + 1] += f.y;
            }
          }
        } else {
          for (int j = 0; j < 32 && col + j < p.N; ++j) a[j] += upk2((uint32_t)r16[j], e.f16).x;
        }
      }
      if (full) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 q = make_uint4(pk2(a[8 * j], a[8 * j + 1], e.f16), pk2(a[8 * j + 2], a[8 * j + 3], e.f16),
                               pk2(a[8 * j + 4], a[8 * j + 5], e.f16), pk2(a[8 * j + 6], a[8 * j + 7], e.f16));
          *reinterpret_cast<uint4*>(out + 8 * j) = q;
        }
      } else {
        for (int j = 0; j < 32 && col + j < p.N; ++j) reinterpret_cast<uint16_t*>(out)[j] = (uint16_t)(pk2(a[j], 0.f, e.f16) & 0xffff);
      }
    }
  };
  tmem_ld_32x32(tq, v);
  prefetch_bias(0, bq[0]);
#pragma unroll 1
  for (int c = 0; c < NC; c += 2) {
    tmem_ld_wait();
    if (c + 1 < NC) { tmem_ld_32x32(tq + (c + 1) * 32, vb); prefetch_bias(c + 1, bq[1]); }
    process(c, v, bq[0]);
    if (c + 1 < NC) {
      tmem_ld_wait();
      if (c + 2 < NC) { tmem_ld_32x32(tq + (c + 2) * 32, v); prefetch_bias(c + 2, bq[0]); }
      process(c + 1, vb, bq[1]);
    }
  }
}


// ------------------------------------------------------------------------------------------------ W-only dequant (WQ kernels)
// One B row of one k-block: 64 packed codes from a slot row (TMA wrote the slot 64 B- / 32 B-swizzled, so the 16 B loads of a
// quarter warp hit distinct banks) -> 64 16-bit operands in the 128 B-swizzled K-major B stage. The arithmetic is
// dequantize_kernel's (quant.cu), operation for operation — q * scale + bias as separate fp32 multiply and add, block-scaled
// element * scale as one fp32 multiply, one rounding to the operand type — so the tile holds exactly the bits the dense working
// copy would hold and the GEMM result is bit-identical to the dense-copy path.
__device__ __forceinline__ float wq_u8_to_float(uint32_t word, int j) {
  // (float)byte_j without I2F (quarter rate): 0x4B000000 | q is the float 2^23 + q; the subtraction is exact
  return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540u | (uint32_t)j)) - 8388608.0f;
}
__device__ __forceinline__ float wq_sb_to_float(uint16_t h, int bf16) {
  return bf16 ? __uint_as_float((uint32_t)h << 16) : __half2float(__ushort_as_half(h));
}
__device__ __forceinline__ void wq_store_chunk(uint32_t brow, int r, int oc, const float (&v)[8], int f16) {
  sts128(brow + ((oc ^ (r & 7)) << 4), make_uint4(pk2(v[0], v[1], f16), pk2(v[2], v[3], f16), pk2(v[4], v[5], f16), pk2(v[6], v[7], f16)));
}
// MODE = flux2b_quant. `sc` / `bi`: this row's scales / biases for THIS k-block, already in registers:
//   affine: sc = scale, bi = bias (fp32 values of the stored 16-bit numbers); mx: sc[0..1] (group 32) / nv: sc[0..3] (group 16)
// PARTS = threads per row (1: the whole row; 2: `part` selects the first / second 32 elements).
// Affine modes: q * scale is exact in fp32 (q <= 8 bits, scale <= 11 bits), so one FFMA gives the same bits as
// dequantize_kernel's separate multiply and add; codes become floats through PRMT into the mantissa of 2^23 (no I2F).
template <int MODE, int PARTS>
__device__ __forceinline__ void wq_dequant_row(uint32_t slot, uint32_t bstage, int r, int part, const float (&sc)[4], float bi, int f16) {
  const uint32_t brow = bstage + r * 128;
  if constexpr (MODE == 1 || MODE == 3) {   // 8-bit codes: 64 bytes, slot rows of 64 B, 64 B swizzle: chunk c at c ^ ((r >> 1) & 3)
    const uint32_t src = slot + r * 64;
    const int sw = (r >> 1) & 3;
#pragma unroll
    for (int ci = 0; ci < 4 / PARTS; ++ci) {
      const int c = PARTS == 1 ? ci : 2 * part + ci;
      const uint4 q = lds128(src + ((c ^ sw) << 4));
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
      float s = sc[0];
      if constexpr (MODE == 3) s = PARTS == 1 ? sc[ci >> 1] : (part ? sc[1] : sc[0]);   // E8M0 scale per 32 elements
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v[8];
        if constexpr (MODE == 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __fmaf_rn(wq_u8_to_float(w[2 * h + (j >> 2)], j & 3), s, bi);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t word = w[2 * h + (j >> 1)];
            const __half2_raw hr = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)((word >> (16 * (j & 1))) & 0xffffu), __NV_E4M3);
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hr));
            v[2 * j] = __fmul_rn(f.x, s); v[2 * j + 1] = __fmul_rn(f.y, s);
          }
        }
        wq_store_chunk(brow, r, 2 * c + h, v, f16);
      }
    }
  } else {                            // 4-bit codes: 32 bytes, slot rows of 32 B, 32 B swizzle: chunk c at c ^ ((r >> 2) & 1)
    const uint32_t src = slot + r * 32;
    const int sw = (r >> 2) & 1;
#pragma unroll
    for (int ci = 0; ci < 2 / PARTS; ++ci) {
      const int c = PARTS == 1 ? ci : part;
      const uint4 q = lds128(src + ((c ^ sw) << 4));
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {   // one word = 8 codes = one 16 B operand chunk
        float v[8];
        if constexpr (MODE == 2) {    // int4 affine: element 2b = low nibble of byte b, 2b + 1 = its high nibble
          const uint32_t lo = w[i] & 0x0F0F0F0Fu, hi = (w[i] >> 4) & 0x0F0F0F0Fu;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            v[2 * b] = __fmaf_rn(wq_u8_to_float(lo, b), sc[0], bi);
            v[2 * b + 1] = __fmaf_rn(wq_u8_to_float(hi, b), sc[0], bi);
          }
        } else {                      // mxfp4 (scale per 32 = per 4 words) / nvfp4 (scale per 16 = per 2 words)
          float s;
          if constexpr (MODE == 4) s = PARTS == 1 ? sc[ci] : (part ? sc[1] : sc[0]);
          else s = PARTS == 1 ? sc[2 * ci + (i >> 1)] : (part ? sc[2 + (i >> 1)] : sc[i >> 1]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const __half2_raw hr = __nv_cvt_fp4x2_to_halfraw2((__nv_fp4x2_storage_t)((w[i] >> (8 * j)) & 0xffu), __NV_E2M1);
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hr));
            v[2 * j] = __fmul_rn(f.x, s); v[2 * j + 1] = __fmul_rn(f.y, s);
          }
        }
        wq_store_chunk(brow, r, 4 * c + i, v, f16);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ W-only staging kernel
// Packed codes [N, K * bits / 8] (row stride ldb bytes) -> 16-bit [N, K] operand. One thread per 8 elements (one 16 B store; a
// warp reads 128 / 256 contiguous bytes of codes and writes 512 contiguous bytes), eight elements never straddle a scale group.
// Arithmetic = wq_dequant_row's = dequantize_kernel's.
template <int MODE>
__global__ void __launch_bounds__(256) wq_stage_kernel(const uint8_t* __restrict__ packed, int64_t ldb, const uint8_t* __restrict__ scales,
                                                       const uint8_t* __restrict__ biases, int sb_ld, int sb_bf16, int N, int K,
                                                       uint16_t* __restrict__ out, int f16) {
  const int cpr = K >> 3;                                     // 8-element chunks per row
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * cpr) return;
  const int row = (int)(idx / cpr), c = (int)(idx - (int64_t)row * cpr);
  const int k0 = c << 3;
  float v[8];
  if constexpr (MODE == 1 || MODE == 3) {
    const uint2 q = __ldg(reinterpret_cast<const uint2*>(packed + (int64_t)row * ldb) + c);
    if constexpr (MODE == 1) {
      const int64_t gi = (int64_t)row * sb_ld + (k0 >> 6);
      const float s = wq_sb_to_float(__ldg(reinterpret_cast<const uint16_t*>(scales) + gi), sb_bf16);
      const float b = wq_sb_to_float(__ldg(reinterpret_cast<const uint16_t*>(biases) + gi), sb_bf16);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __fmaf_rn(wq_u8_to_float(j < 4 ? q.x : q.y, j & 3), s, b);
    } else {
      const float s = from_e8m0(__ldg(scales + (int64_t)row * sb_ld + (k0 >> 5)));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t word = j < 2 ? q.x : q.y;
        const __half2_raw hr = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)((word >> (16 * (j & 1))) & 0xffffu), __NV_E4M3);
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hr));
        v[2 * j] = __fmul_rn(f.x, s); v[2 * j + 1] = __fmul_rn(f.y, s);
      }
    }
  } else {
    const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(packed + (int64_t)row * ldb) + c);
    if constexpr (MODE == 2) {
      const int64_t gi = (int64_t)row * sb_ld + (k0 >> 6);
      const float s = wq_sb_to_float(__ldg(reinterpret_cast<const uint16_t*>(scales) + gi), sb_bf16);
      const float b = wq_sb_to_float(__ldg(reinterpret_cast<const uint16_t*>(biases) + gi), sb_bf16);
      const uint32_t lo = w & 0x0F0F0F0Fu, hi = (w >> 4) & 0x0F0F0F0Fu;
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        v[2 * bb] = __fmaf_rn(wq_u8_to_float(lo, bb), s, b);
        v[2 * bb + 1] = __fmaf_rn(wq_u8_to_float(hi, bb), s, b);
      }
    } else {
      const uint8_t sb = __ldg(scales + (int64_t)row * sb_ld + (MODE == 4 ? (k0 >> 5) : (k0 >> 4)));
      const float s = MODE == 4 ? from_e8m0(sb) : from_e4m3(sb);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __half2_raw hr = __nv_cvt_fp4x2_to_halfraw2((__nv_fp4x2_storage_t)((w >> (8 * j)) & 0xffu), __NV_E2M1);
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hr));
        v[2 * j] = __fmul_rn(f.x, s); v[2 * j + 1] = __fmul_rn(f.y, s);
      }
    }
  }
  *reinterpret_cast<uint4*>(out + (int64_t)row * K + k0) =
      make_uint4(pk2(v[0], v[1], f16), pk2(v[2], v[3], f16), pk2(v[4], v[5], f16), pk2(v[6], v[7], f16));
}

// ------------------------------------------------------------------------------------------------ kernel
// CONV: 0 = plain GEMM, 1 = implicit-GEMM convolution with one TMA box per tap, 2 = 3x3 stride-1 convolution from a halo tile
template <int BN, int CG, int CONV, int MXK = 0, int WQ = 0>
__global__ void __launch_bounds__(WQ ? 192 + 32 * WQ_DQ_WARPS : 192, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmSFA, const __grid_constant__ CUtensorMap tmSFB, const KParams p) {
  constexpr bool HALO = CONV == 2;
  using C = Cfg<BN, CG, MXK, WQ, HALO ? 1 : 0>;
  static_assert(!MXK || (!CONV && (BN == 128 || (BN == 256 && CG == 1))), "block-scaled tiles: BN 128 (1 or 2 CTAs) / 256 (1 CTA)");
  static_assert(!WQ || (!CONV && !MXK && BN == 256), "W-only quantized weights: plain GEMM, 256-wide tiles");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smA = smem;
  uint8_t* smB = smem + C::STAGES * C::A_STAGE;
  uint8_t* smSF = smB + C::STAGES * C::B_BYTES;  // [STAGES][SFA SFPK x 512 | SFB (BN/128) x SFPK x 512]   (block-scaled only)
  uint8_t* smP = smem + C::STAGES * C::STAGE_BYTES;   // [WQ_PSTAGES][PSLOT_BYTES] packed-weight slots (WQ only)
  uint8_t* smH = smP + C::PRING_BYTES;                // [HALO_STAGES][HALO_BYTES] halo tiles (conv mode 2; 1024 B aligned: stages are multiples of 1 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smH + C::HRING_BYTES);
  uint64_t* full = bars;                    // [STAGES]  TMA (+ dequant warps) -> MMA
  uint64_t* empty = bars + C::STAGES;       // [STAGES]  MMA -> TMA (+ dequant warps)
  uint64_t* tfull = bars + 2 * C::STAGES;   // [2]       MMA -> epilogue
  uint64_t* tempty = tfull + 2;             // [2]       epilogue -> MMA   (only ACC_STAGES of each are used)
  uint64_t* pfull = tempty + 2;             // [WQ_PSTAGES]  TMA -> dequant warps   (WQ only)
  uint64_t* pempty = pfull + WQ_PSTAGES;    // [WQ_PSTAGES]  dequant warps -> TMA
  uint64_t* hfull = pempty + WQ_PSTAGES;    // [HALO_STAGES] TMA -> MMA   (conv mode 2 only)
  uint64_t* hempty = hfull + HALO_STAGES;   // [HALO_STAGES] MMA -> TMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hempty + HALO_STAGES);
  static_assert(!HALO || (C::B_TAP_BYTES % 1024 == 0), "tap tiles and the halo ring behind them must stay 1024 B aligned");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint64_t t_entry = p.dbg ? globaltimer_ns() : 0;   // debug timeline: ns since this thread entered the kernel
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = (cta_rank == 0);
  const int unit_id = (CG == 2) ? (blockIdx.x >> 1) : blockIdx.x;
  const int num_units = (CG == 2) ? (gridDim.x >> 1) : gridDim.x;
  const int total_tiles = p.num_m_units * p.num_n_blks * ((HALO && p.up2) ? 4 : 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (MXK == 0 && !CONV && p.split_units > 0) tma_prefetch_desc(&tmSFB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      // one arrive.expect_tx per CTA of the pair (+ one arrive per dequant warp; resident conv weights: one per channel block)
      mbar_init(&full[s], WQ ? CG * (1 + WQ_DQ_WARPS) : (HALO && p.wres && s == 0) ? CG * p.kb_per_tap : CG);
      mbar_init(&empty[s], 1);  // one tcgen05.commit
    }
    if (WQ) {
      for (int s = 0; s < WQ_PSTAGES; ++s) {
        mbar_init(&pfull[s], 1);    // the producer's arrive.expect_tx (CTA-local)
        mbar_init(&pempty[s], WQ_DQ_WARPS);   // one arrive per dequant warp
      }
    }
    if (HALO) {
      for (int s = 0; s < HALO_STAGES; ++s) {
        mbar_init(&hfull[s], CG);   // one arrive.expect_tx per CTA of the pair (each loads the halo of its own 128 pixels)
        mbar_init(&hempty[s], 1);   // one tcgen05.commit
      }
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);        // one tcgen05.commit
      mbar_init(&tempty[s], 4 * CG);  // one arrive per epilogue warp (of both CTAs)
    }
    fence_mbar_init();
  }
  // convolutions: bias -> shared memory (a weight: safe to read before the dependency wait). The epilogue's per-chunk global
  // loads of it were an exposed L2 round trip per 32 columns (the 200+ KB of dynamic shared memory leaves almost no L1).
  const bool bias_staged = CONV && p.epi.bias && p.N <= CONV_BIAS_MAX;
  const uint32_t sbias = bias_staged ? smem_u32(reinterpret_cast<uint8_t*>(bars) + C::BAR_BYTES) : 0u;
  if (bias_staged) {
    float* sb = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + C::BAR_BYTES);
    for (int i = threadIdx.x; i < CONV_BIAS_MAX; i += blockDim.x) sb[i] = i < p.N ? __ldg(p.epi.bias + i) : 0.f;
  }
  if (warp == 1) tmem_alloc<CG>(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above ran beside the tail of the previous kernel (programmatic dependent launch); global memory from here on
  pdl_trigger();
  pdl_wait();
  const uint64_t t_prologue = p.dbg ? globaltimer_ns() - t_entry : 0;   // printed at the end: a printf here would stall the producer thread

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int pstage = 0, pk_tile = unit_id, pk_kb = 0;   // WQ: cursor of the packed-weight stream
      uint32_t pphase = 0;
      int hstage = 0;                                  // conv mode 2: halo ring
      uint32_t hphase = 0;
      (void)pstage; (void)pphase; (void)pk_tile; (void)pk_kb; (void)hstage; (void)hphase;
      const bool dbg = p.dbg && unit_id == 0;
      long long w_empty = 0, t_begin = dbg ? clock64() : 0;
      for (int t = unit_id; t < total_tiles; t += num_units) {
        // (folded upsample: the four output phases of an M tile are four consecutive tiles of the schedule)
        const int tq = (HALO && p.up2) ? t >> 2 : t;
        const int phase_idx = (HALO && p.up2) ? (t & 3) : 0;
        const int m_unit = p.n_fast ? tq / p.num_n_blks : tq % p.num_m_units;
        const int n_blk = p.n_fast ? tq % p.num_n_blks : tq / p.num_m_units;
        const int m_blk = m_unit * CG + (int)cta_rank;
        const int nrow0 = n_blk * BN + (int)cta_rank * C::B_ROWS;
        const CUtensorMap* tmBsel = (MXK == 0 && !CONV && m_unit < p.split_units) ? &tmSFB : &tmB;
        (void)tmBsel;
        int img = 0, y0 = 0, x0 = 0;
        if (CONV) {
          const int per_img = p.tiles_x * p.tiles_y;
          img = m_blk / per_img;
          const int r = m_blk % per_img;
          y0 = (r / p.tiles_x) * (HALO ? HALO_TH : CONV_TH);
          x0 = (r % p.tiles_x) * (HALO ? HALO_TW : CONV_TW);
        }
        if constexpr (HALO) {
          // channel-block major: one halo load per 64 input channels, then the nine taps' weight tiles through the ring
          for (int cb = 0; cb < p.kb_per_tap; ++cb) {
            mbar_wait(&hempty[hstage], hphase ^ 1, 8);
            uint8_t* h_dst = smH + hstage * HALO_STRIDE;
            if (CG == 1) {
              mbar_expect_tx(&hfull[hstage], HALO_BYTES);
              tma_load_4d(h_dst, &tmA, &hfull[hstage], cb * BK, x0 - 1, y0 - 1, img);
            } else {
              const uint32_t hbar = mapa_u32(smem_u32(&hfull[hstage]), 0);
              mbar_expect_tx_cluster(hbar, HALO_BYTES);
              tma_load_4d_cg2(h_dst, &tmA, hbar, cb * BK, x0 - 1, y0 - 1, img);
            }
            if (++hstage == HALO_STAGES) { hstage = 0; hphase ^= 1; }
            if (p.wres) {
              // resident weights: all (channel block, tap) tiles of this CTA's N rows go into the ring area once, behind the first
              // halo; every later tile only loads halos (the 96-wide 1024^2 layers re-fetched 108 KB of weights per 256 pixels)
              if (t == unit_id) {
                const uint32_t bytes = (uint32_t)(9 * C::B_TAP_BYTES);
                if (CG == 1) {
                  mbar_expect_tx(&full[0], bytes);
                  for (int tap = 0; tap < 9; ++tap)
                    tma_load_3d(smB + (cb * 9 + tap) * C::B_TAP_BYTES, &tmB, &full[0], cb * BK, tap, nrow0);
                } else {
                  const uint32_t lbar = mapa_u32(smem_u32(&full[0]), 0);
                  mbar_expect_tx_cluster(lbar, bytes);
                  for (int tap = 0; tap < 9; ++tap)
                    tma_load_4d_cg2(smB + (cb * 9 + tap) * C::B_TAP_BYTES, &tmB, lbar, cb * BK, tap, nrow0, 0);
                }
              }
              continue;
            }
            // weight tiles: one stage per tap group (one kernel row; folded upsample: one row of the 2 x 2 phase kernel)
            const int tg = p.up2 ? (C::TG == 3 ? 2 : 1) : C::TG;
            const int ngroups = (p.up2 ? 4 : 9) / tg, tap0 = p.up2 ? phase_idx * 4 : 0;
            for (int grp = 0; grp < ngroups; ++grp) {
              mbar_wait(&empty[stage], phase ^ 1, 1);
              uint8_t* b_dst = smB + stage * C::B_BYTES;
              if (CG == 1) {
                mbar_expect_tx(&full[stage], (uint32_t)(tg * C::B_TAP_BYTES));
                for (int j = 0; j < tg; ++j)
                  tma_load_3d(b_dst + j * C::B_TAP_BYTES, &tmB, &full[stage], cb * BK, tap0 + grp * tg + j, nrow0);
              } else {
                const uint32_t lbar = mapa_u32(smem_u32(&full[stage]), 0);
                mbar_expect_tx_cluster(lbar, (uint32_t)(tg * C::B_TAP_BYTES));
                for (int j = 0; j < tg; ++j)
                  tma_load_4d_cg2(b_dst + j * C::B_TAP_BYTES, &tmB, lbar, cb * BK, tap0 + grp * tg + j, nrow0, 0);
              }
              if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
        if constexpr (WQ != 0) {
          // W-only quantized weights: this CTA's B rows arrive as packed codes in their own slot ring, as one continuous stream
          // over (tile, k-block) that runs WQ_PSTAGES - 1 blocks ahead of the A loads — across tile boundaries too (the dequant
          // warps need the codes the moment a B stage frees). A goes the usual way, and only its bytes are expected on `full`:
          // the dequant warps arrive there once the 16-bit B tile is written.
          const int pcol = (p.wq_mode == 1 || p.wq_mode == 3) ? 64 : 32;   // packed bytes per row per k-block
          auto packed_next = [&]() {
            if (pk_tile >= total_tiles) return;
            const int pm_unit = p.n_fast ? pk_tile / p.num_n_blks : pk_tile % p.num_m_units, pn_blk = p.n_fast ? pk_tile % p.num_n_blks : pk_tile / p.num_m_units;
            const CUtensorMap* tmP = (pm_unit < p.split_units) ? &tmSFB : &tmB;
            mbar_wait(&pempty[pstage], pphase ^ 1, 5);
            mbar_expect_tx(&pfull[pstage], (uint32_t)(C::B_ROWS * pcol));
            tma_load_2d(smP + pstage * C::PSLOT_BYTES, tmP, &pfull[pstage], pk_kb * pcol, pn_blk * BN + (int)cta_rank * C::B_ROWS);
            if (++pstage == WQ_PSTAGES) { pstage = 0; pphase ^= 1; }
            if (++pk_kb == p.num_kb) { pk_kb = 0; pk_tile += num_units; }
          };
          if (t == unit_id)
            for (int i = 0; i < WQ_PSTAGES - 1; ++i) packed_next();
          for (int kb = 0; kb < p.num_kb; ++kb) {
            packed_next();
            mbar_wait(&empty[stage], phase ^ 1, 1);
            void* a_dst = smA + stage * C::A_BYTES;
            if (CG == 1) {
              mbar_expect_tx(&full[stage], C::A_BYTES);
              tma_load_2d(a_dst, &tmA, &full[stage], kb * BK, m_blk * BM);
            } else {
              const uint32_t lbar = mapa_u32(smem_u32(&full[stage]), 0);
              mbar_expect_tx_cluster(lbar, C::A_BYTES);
              tma_load_2d_cg2(a_dst, &tmA, lbar, kb * BK, m_blk * BM);
            }
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
          }
        }
        if constexpr (WQ == 0 && !HALO)
        for (int kb = 0; kb < p.num_kb; ++kb) {
          long long t0 = dbg ? clock64() : 0;
          mbar_wait(&empty[stage], phase ^ 1, 1);
          if (dbg) w_empty += clock64() - t0;
          void* a_dst = smA + stage * C::A_BYTES;
          void* b_dst = smB + stage * C::B_BYTES;
          if (CG == 1) {
            mbar_expect_tx(&full[stage], C::STAGE_BYTES);
            if (CONV) {
              const int tap = kb / p.kb_per_tap;
              const int c0 = (kb % p.kb_per_tap) * BK;
              const int ky = (p.taps == 9) ? tap / 3 + p.tap_off : 0;
              const int kx = (p.taps == 9) ? tap % 3 + p.tap_off : 0;
              tma_load_4d(a_dst, &tmA, &full[stage], c0, x0 * p.stride + kx, y0 * p.stride + ky, img);
              tma_load_3d(b_dst, &tmB, &full[stage], c0, tap, nrow0);
            } else {
              // (tensor-map coordinates are in elements: 64 bf16, or 128 bytes of fp8 / packed fp4, per 128 B row)
              tma_load_2d(a_dst, &tmA, &full[stage], kb * (MXK ? 2 * BK : BK), m_blk * BM);
              tma_load_2d(b_dst, tmBsel, &full[stage], kb * (MXK ? 2 * BK : BK), nrow0);
              if (MXK) {
                uint8_t* sf = smSF + stage * C::SF_BYTES;
                bulk_load(sf, p.sfa + ((size_t)m_blk * p.sfa_ld + kb * C::SFPK) * 512, C::SFA_BYTES, &full[stage]);
#pragma unroll
                for (int i = 0; i < BN / 128; ++i)
                  bulk_load(sf + C::SFA_BYTES + i * C::SFPK * 512,
                            p.sfb + ((size_t)(n_blk * (BN / 128) + i) * p.sfb_ld + kb * C::SFPK) * 512, C::SFPK * 512, &full[stage]);
              }
            }
          } else {
            const uint32_t lbar = mapa_u32(smem_u32(&full[stage]), 0);
            mbar_expect_tx_cluster(lbar, C::STAGE_BYTES);
            if (CONV) {
              const int tap = kb / p.kb_per_tap;
              const int c0 = (kb % p.kb_per_tap) * BK;
              const int ky = (p.taps == 9) ? tap / 3 + p.tap_off : 0;
              const int kx = (p.taps == 9) ? tap % 3 + p.tap_off : 0;
              tma_load_4d_cg2(a_dst, &tmA, lbar, c0, x0 * p.stride + kx, y0 * p.stride + ky, img);
              // 3-D weight box through the cta_group::2 path is expressed as 4-D with a unit outer dim
              tma_load_4d_cg2(b_dst, &tmB, lbar, c0, tap, nrow0, 0);
            } else {
              tma_load_2d_cg2(a_dst, &tmA, lbar, kb * (MXK ? 2 * BK : BK), m_blk * BM);
              tma_load_2d_cg2(b_dst, tmBsel, lbar, kb * (MXK ? 2 * BK : BK), nrow0);
              if (MXK) {
                // scale factors through tensor maps ([blocks][128 x u32]): a plain bulk copy cannot signal the leader's barrier.
                // Each CTA stages the scales of its own 128 A rows and of ALL BN weight rows of the tile.
                uint8_t* sf = smSF + stage * C::SF_BYTES;
                tma_load_2d_cg2(sf, &tmSFA, lbar, 0, m_blk * p.sfa_ld + kb * C::SFPK);
#pragma unroll
                for (int i = 0; i < BN / 128; ++i)
                  tma_load_2d_cg2(sf + C::SFA_BYTES + i * C::SFPK * 512, &tmSFB, lbar, 0,
                                  (n_blk * (BN / 128) + i) * p.sfb_ld + kb * C::SFPK);
              }
            }
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (dbg) printf("[gemm timeline] cta %d producer: total %lld clk, waited on empty slots %lld clk; last load issued at %llu ns\n", (int)cta_rank,
                      clock64() - t_begin, w_empty, (unsigned long long)(globaltimer_ns() - t_entry));
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only)
    // The whole warp walks the loop (warp-uniform control flow, barrier waits by all lanes); one elected lane issues.
    // Inside an `if (lane == 0)` region ptxas wraps every tcgen05.mma in an elect / broadcast loop with operand reloads,
    // which costs more issue time than a narrow (BN <= 128) MMA takes to execute.
    if (leader) {
      const uint32_t idesc = make_idesc_f16(BM * CG, BN, p.epi.f16 == 0, false, false);
      const uint64_t desc_hi = make_smem_desc(0, 16, 1024, SWZ_128B);
      const uint64_t desc_sf = make_smem_desc(0, 0, 128, SWZ_NONE);  // 32 x 16 B block: 8-row atoms 128 B apart
      const uint64_t desc_halo = make_smem_desc(0, 16, HALO_W * 128, SWZ_128B);   // halo windows: 8-row groups one halo row apart
      (void)desc_halo;
      const uint32_t a0 = smem_u32(smA) >> 4, b0 = smem_u32(smB) >> 4, sf0 = smem_u32(smSF) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int hstage = 0;
      uint32_t hphase = 0;
      (void)hstage; (void)hphase;
      const bool dbg = p.dbg && unit_id == 0;
      long long w_full = 0, w_tempty = 0, t_begin = dbg ? clock64() : 0;
      int tl_n = 0;
      unsigned tl_start[8], tl_drain[8], tl_oper[8];
      for (int t = unit_id; t < total_tiles; t += num_units) {
        long long t0 = dbg ? clock64() : 0;
        const long long w_full_before = w_full;
        mbar_wait<CG == 2>(&tempty[acc], acc_phase ^ 1, 2);
        if (dbg) w_tempty += clock64() - t0;
        if (dbg && tl_n < 8) { tl_start[tl_n] = (unsigned)(globaltimer_ns() - t_entry); tl_drain[tl_n] = (unsigned)(clock64() - t0); }
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        if constexpr (HALO) {
          // conv mode 2: per 64-channel block one halo tile, nine taps = nine windows of it (see HALO_* above). The last channel
          // block of a layer whose Cin is not a multiple of 64 (the 96-channel layers: 64 + 32) issues only the MMAs that carry
          // real channels instead of multiplying the TMA's zero fill.
          const uint32_t h0 = smem_u32(smH) >> 4;
          const int tail = p.Cin - (p.kb_per_tap - 1) * BK;          // channels in the last block, 1..64
          // folded upsample: phase (py, px) of the output reads the 2x2 source window that starts at halo (py, px)
          const int py = p.up2 ? ((t & 3) >> 1) : 0, px = p.up2 ? (t & 1) : 0;
          const int tg = p.up2 ? (C::TG == 3 ? 2 : 1) : C::TG;
          const int ngroups = (p.up2 ? 4 : 9) / tg;
          for (int cb = 0; cb < p.kb_per_tap; ++cb) {
            t0 = dbg ? clock64() : 0;
            mbar_wait<CG == 2>(&hfull[hstage], hphase, 9);
            if (dbg) w_full += clock64() - t0;
            const int nmma = (cb == p.kb_per_tap - 1) ? (tail + 15) / 16 : BK / 16;
            const uint32_t hbase = h0 + hstage * (HALO_STRIDE >> 4);
            for (int grp = 0; grp < ngroups; ++grp) {
              t0 = dbg ? clock64() : 0;
              if (!p.wres) mbar_wait<CG == 2>(&full[stage], phase, 3);
              else if (t == unit_id && cb == 0 && grp == 0) mbar_wait<CG == 2>(&full[0], 0, 3);   // resident weights have landed
              if (dbg) w_full += clock64() - t0;
              tc_fence_after();
              if (elect_one()) {
                const uint64_t bdesc0 = desc_hi + (p.wres ? b0 + (cb * 9 + grp * tg) * (C::B_TAP_BYTES >> 4) : b0 + stage * (C::B_BYTES >> 4));
#pragma unroll
                for (int j = 0; j < C::TG; ++j) {
                  if (j < tg) {
                    const int tap = grp * tg + j;
                    int ky, kx;
                    if (p.up2) { ky = py + (tap >> 1); kx = px + (tap & 1); } else { ky = tap / 3; kx = tap - 3 * ky; }
                    // window start: halo row ky, column kx; 8-row groups one halo row (2048 B) apart
                    const uint64_t adesc = desc_halo + (hbase + (((ky * HALO_W + kx) * 128) >> 4));
                    const uint64_t bdesc = bdesc0 + j * (C::B_TAP_BYTES >> 4);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                      if (k < nmma) umma_f16_ss<CG>(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (cb | tap | k) ? 1u : 0u);
                  }
                }
                if (!p.wres) { if (CG == 1) umma_commit(&empty[stage]); else umma_commit_cg2_mc(&empty[stage], 0x3); }
                if (grp == ngroups - 1) {
                  if (CG == 1) umma_commit(&hempty[hstage]); else umma_commit_cg2_mc(&hempty[hstage], 0x3);
                  if (cb == p.kb_per_tap - 1) {
                    if (CG == 1) umma_commit(&tfull[acc]); else umma_commit_cg2_mc(&tfull[acc], 0x3);
                  }
                }
              }
              __syncwarp();
              if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
            if (++hstage == HALO_STAGES) { hstage = 0; hphase ^= 1; }
          }
        }
        if constexpr (!HALO)
        for (int kb = 0; kb < p.num_kb; ++kb) {
          t0 = dbg ? clock64() : 0;
          mbar_wait<CG == 2>(&full[stage], phase, 3);
          if (dbg) w_full += clock64() - t0;
          tc_fence_after();
          if (elect_one()) {
            const uint64_t adesc = desc_hi + (a0 + stage * (C::A_BYTES >> 4));
            const uint64_t bdesc = desc_hi + (b0 + stage * (C::B_BYTES >> 4));
            if constexpr (MXK != 0) {
              // stage this k-block's scale factors: every 512 B block -> 4 TMEM columns (SFA: block j at +4j; SFB: block j
              // of weight rows [128 i, 128 i + 128) at +4 (j * BN/128 + i), so the columns one MMA reads are adjacent).
              // tcgen05.cp and tcgen05.mma of one thread execute in issue order: the single SF region needs no barrier.
              constexpr int NB = BN / 128;
              const uint32_t t_sfa = tmem_base + C::SF_COL0, t_sfb = t_sfa + 4 * C::SFPK;
              const uint32_t sfs = sf0 + stage * (C::SF_BYTES >> 4);
#pragma unroll
              for (int j = 0; j < C::SFPK; ++j) tmem_cp_32x128b_warpx4<CG>(t_sfa + 4 * j, desc_sf + (sfs + 32 * j));
#pragma unroll
              for (int i = 0; i < NB; ++i)
#pragma unroll
                for (int j = 0; j < C::SFPK; ++j)
                  tmem_cp_32x128b_warpx4<CG>(t_sfb + 4 * (j * NB + i), desc_sf + (sfs + 32 * C::SFPK + 32 * (i * C::SFPK + j)));
#pragma unroll
              for (int k = 0; k < 4; ++k) {  // 32 B of K per MMA: 32 fp8 or 64 fp4 elements
                const uint32_t accum = (kb | k) ? 1u : 0u;
                if constexpr (MXK == 1)       // one scale per row per MMA: byte k of the staged columns
                  umma_mxf8_ss<CG>(tmem_d, adesc + 2 * k, bdesc + 2 * k, make_idesc_mxf8(BM * CG, BN, k), t_sfa, t_sfb, accum);
                else if constexpr (MXK == 2)  // two scales (bytes 0,1 or 2,3) of block k / 2
                  umma_mxf4_ss<false, CG>(tmem_d, adesc + 2 * k, bdesc + 2 * k, make_idesc_mxf4(BM * CG, BN, true, (k & 1) * 2),
                                      t_sfa + 4 * (k >> 1), t_sfb + 4 * NB * (k >> 1), accum);
                else                          // four scales = the whole column of block k
                  umma_mxf4_ss<true, CG>(tmem_d, adesc + 2 * k, bdesc + 2 * k, make_idesc_mxf4(BM * CG, BN, false, 0),
                                     t_sfa + 4 * k, t_sfb + 4 * NB * k, accum);
              }
            } else {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                // advance 16 elements (32 B) along K inside the 128 B swizzle atom: +2 in the 16 B-unit address field
                umma_f16_ss<CG>(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
              }
            }
            if (CG == 1) umma_commit(&empty[stage]); else umma_commit_cg2_mc(&empty[stage], 0x3);
            if (kb == p.num_kb - 1) {
              if (CG == 1) umma_commit(&tfull[acc]); else umma_commit_cg2_mc(&tfull[acc], 0x3);
            }
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        if (dbg && tl_n < 8) { tl_oper[tl_n] = (unsigned)(w_full - w_full_before); ++tl_n; }
        if (++acc == C::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
      if (dbg && lane == 0)
        for (int i = 0; i < tl_n; ++i)
          printf("[gemm timeline] issuer tile %d: started at %u ns after waiting %u clk for its accumulator stage, then %u clk for operands\n",
                 i, tl_start[i], tl_drain[i], tl_oper[i]);
      if (dbg && lane == 0)
        printf("[gemm timeline] issuer: total %lld clk, waited on operands %lld clk, on accumulator drain %lld clk (%d k-blocks / tile); last MMA issued at %llu ns\n",
               clock64() - t_begin, w_full, w_tempty, p.num_kb, (unsigned long long)(globaltimer_ns() - t_entry));
    }
  } else if (warp >= 6) {
   if constexpr (WQ != 0) {
    // ===================================================== dequant warps (6..13, WQ only): packed slot -> 16-bit B stage
    // CTA pairs (128 B rows per CTA): two threads per row, 32 elements each; single-CTA tiles (256 rows): thread = row.
    // Scales / biases of 8 k-blocks (4 for nvfp4) sit in registers, fetched with one 16 B load per tensor when the row's
    // group array is 16 B aligned (K % 512 == 0), else one by one.
    constexpr int PARTS = (32 * WQ_DQ_WARPS) / C::B_ROWS;   // threads per row
    static_assert(PARTS == 1 || PARTS == 2, "dequant warps: one or two threads per B row");
    const int dt = (threadIdx.x - 192) % C::B_ROWS;
    const int part = (threadIdx.x - 192) / C::B_ROWS;       // warp-uniform
    auto run = [&](auto mode_tag) {
      constexpr int MODE = decltype(mode_tag)::value;
      constexpr int RPT = 1;                         // rows per thread
      constexpr int GPK = MODE <= 2 ? 1 : MODE == 5 ? 4 : 2;   // scale groups per k-block
      constexpr int KBV = MODE <= 2 ? 8 : 16 / GPK;            // k-blocks covered by one 16 B load
      int stage = 0, pstage = 0;
      uint32_t phase = 0, pphase = 0;
      for (int t = unit_id; t < total_tiles; t += num_units) {
        const int m_unit = p.n_fast ? t / p.num_n_blks : t % p.num_m_units;
        const int n_blk = p.n_fast ? t % p.num_n_blks : t / p.num_m_units;
        const bool lo = m_unit < p.split_units;
        const uint8_t* sbase = lo ? p.wq_s_lo : p.wq_s;
        const uint8_t* bbase = lo ? p.wq_b_lo : p.wq_b;
        const int nrow0 = n_blk * BN + (int)cta_rank * C::B_ROWS;
        uint4 sv[RPT], bv[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) { sv[i] = make_uint4(0, 0, 0, 0); bv[i] = make_uint4(0, 0, 0, 0); }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          if (p.wq_vec && (kb % KBV) == 0) {
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
              const int64_t grow = nrow0 + dt + 128 * i;
              if (grow < p.N) {
                const size_t off = ((size_t)grow * p.wq_sb_ld + (size_t)kb * GPK) * (MODE <= 2 ? 2 : 1);
                sv[i] = __ldg(reinterpret_cast<const uint4*>(sbase + off));
                if (MODE <= 2) bv[i] = __ldg(reinterpret_cast<const uint4*>(bbase + off));
              }
            }
          }
          mbar_wait(&pfull[pstage], pphase, 6);
          mbar_wait(&empty[stage], phase ^ 1, 7);
          const uint32_t slot = smem_u32(smP) + pstage * C::PSLOT_BYTES;
          const uint32_t bst = smem_u32(smB) + stage * C::B_BYTES;
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const int r = dt + 128 * i;
            const int64_t grow = nrow0 + r;
            float sc[4] = {0.f, 0.f, 0.f, 0.f};
            float bi = 0.f;
            if (grow < p.N) {
              if (MODE <= 2) {
                uint16_t sh, bh;
                if (p.wq_vec) {
                  const int j = kb % 8;
                  const uint32_t sw = j < 2 ? sv[i].x : j < 4 ? sv[i].y : j < 6 ? sv[i].z : sv[i].w;
                  const uint32_t bw = j < 2 ? bv[i].x : j < 4 ? bv[i].y : j < 6 ? bv[i].z : bv[i].w;
                  sh = (uint16_t)(sw >> (16 * (j & 1))); bh = (uint16_t)(bw >> (16 * (j & 1)));
                } else {
                  const size_t off = (size_t)grow * p.wq_sb_ld + kb;
                  sh = __ldg(reinterpret_cast<const uint16_t*>(sbase) + off); bh = __ldg(reinterpret_cast<const uint16_t*>(bbase) + off);
                }
                sc[0] = wq_sb_to_float(sh, p.wq_sb_bf16); bi = wq_sb_to_float(bh, p.wq_sb_bf16);
              } else {
                uint32_t word;   // this k-block's GPK scale bytes in the low bytes
                if (p.wq_vec) {
                  const int j = (kb % KBV) * GPK;   // byte index within the 16 B
                  const uint32_t w4 = j < 4 ? sv[i].x : j < 8 ? sv[i].y : j < 12 ? sv[i].z : sv[i].w;
                  word = w4 >> (8 * (j & 3));
                } else {
                  word = 0;
                  for (int g = 0; g < GPK; ++g) word |= (uint32_t)__ldg(sbase + (size_t)grow * p.wq_sb_ld + (size_t)kb * GPK + g) << (8 * g);
                }
#pragma unroll
                for (int g = 0; g < GPK; ++g) {
                  const uint8_t sb = (uint8_t)(word >> (8 * g));
                  sc[g] = MODE == 5 ? from_e4m3(sb) : from_e8m0(sb);
                }
              }
            }
            // rows past N: the TMA zero-filled the codes and the scales stay 0 -> zeros (never read back by a valid output column)
            wq_dequant_row<MODE, PARTS>(slot, bst, r, part, sc, bi, p.epi.f16);
          }
          fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
          __syncwarp();
          if (lane == 0) {
            if (CG == 1) mbar_arrive(&full[stage]); else mbar_arrive_cluster(mapa_u32(smem_u32(&full[stage]), 0));
            mbar_arrive(&pempty[pstage]);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
          if (++pstage == WQ_PSTAGES) { pstage = 0; pphase ^= 1; }
        }
      }
    };
    switch (p.wq_mode) {
      case 1: run(std::integral_constant<int, 1>{}); break;
      case 2: run(std::integral_constant<int, 2>{}); break;
      case 3: run(std::integral_constant<int, 3>{}); break;
      case 4: run(std::integral_constant<int, 4>{}); break;
      default: run(std::integral_constant<int, 5>{}); break;
    }
   }
  } else {
    // ===================================================== epilogue warps (2..5): TMEM lane quarter = warp % 4
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = unit_id; t < total_tiles; t += num_units) {
      const int tq = (HALO && p.up2) ? t >> 2 : t;
      const int m_unit = p.n_fast ? tq / p.num_n_blks : tq % p.num_m_units;
      const int n_blk = p.n_fast ? tq % p.num_n_blks : tq / p.num_m_units;
      const int m_blk = m_unit * CG + (int)cta_rank;
      const int r = quarter * 32 + lane;
      bool row_ok;
      int64_t grow;
      if (CONV) {
        const int per_img = p.tiles_x * p.tiles_y;
        const int img = m_blk / per_img;
        const int rr = m_blk % per_img;
        const int y = HALO ? (rr / p.tiles_x) * HALO_TH + r / HALO_TW : (rr / p.tiles_x) * CONV_TH + r / CONV_TW;
        const int x = HALO ? (rr % p.tiles_x) * HALO_TW + r % HALO_TW : (rr % p.tiles_x) * CONV_TW + r % CONV_TW;
        row_ok = (img < p.batch) && (y < p.H) && (x < p.W);
        grow = ((int64_t)img * p.H + y) * p.W + x;
        if (HALO && p.up2)   // source pixel (y, x) -> output pixel (2y + py, 2x + px) of the 2H x 2W image
          grow = ((int64_t)img * 2 * p.H + 2 * y + ((t & 3) >> 1)) * (2 * p.W) + 2 * x + (t & 1);
      } else {
        grow = (int64_t)m_blk * BM + r;
        row_ok = grow < p.M;
      }
      mbar_wait<CG == 2>(&tfull[acc], acc_phase, 4);
      tc_fence_after();
      const uint64_t t_e0 = p.dbg ? globaltimer_ns() : 0;
      epilogue_tile<BN>(p, tmem_base + acc * BN, quarter, lane, row_ok, grow, p.n_off + n_blk * BN, sbias);
      tc_fence_before();
      __syncwarp();
      const uint64_t t_e1 = p.dbg ? globaltimer_ns() : 0;
      if (lane == 0) {
        if (CG == 1) mbar_arrive(&tempty[acc]);
        else mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[acc]), 0));
      }
      // (debug print AFTER the accumulator stage has been handed back: a printf takes tens of microseconds)
      if (p.dbg && unit_id == 0 && warp == 2 && lane == 0 && t < 2 * num_units)
        printf("[gemm timeline] cta %d epilogue of tile %d: accumulator ready at %llu ns, stored at %llu ns\n", (int)cta_rank, t,
               (unsigned long long)(t_e0 - t_entry), (unsigned long long)(t_e1 - t_entry));
      if (++acc == C::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) tmem_dealloc<CG>(tmem_base, C::TMEM_COLS);
  if (p.dbg && unit_id == 0 && threadIdx.x == 0)
    printf("[gemm timeline] cta %d: prologue (barriers, TMEM alloc, cluster sync, dependency wait) done at %llu ns, kernel end at %llu ns\n",
           (int)cta_rank, (unsigned long long)t_prologue, (unsigned long long)(globaltimer_ns() - t_entry));
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;
static int g_num_sms = 0;
bool pdl_enabled() {
  static const bool on = !(getenv("FLUX2B_PDL") && atoi(getenv("FLUX2B_PDL")) == 0);
  return on;
}
static thread_local std::string g_err;
const char* gemm_last_error() { return g_err.c_str(); }

bool gemm_init() {
  static std::once_flag once;
  std::call_once(once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  });
  return g_encode != nullptr && g_num_sms > 0;
}

static bool make_tmap(CUtensorMap* m, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B,
                      const uint32_t* elem_strides = nullptr);
bool make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box) {
  return make_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box);
}
static bool make_tmap(CUtensorMap* m, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz, const uint32_t* elem_strides) {
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  // with an element stride s the box extent is given in traversed elements: s * (elements loaded)
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; es[i] = elem_strides ? elem_strides[i] : 1; bx[i] = box[i] * es[i]; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = g_encode(m, dtype, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_err = "cuTensorMapEncodeTiled failed, CUresult=" + std::to_string((int)r);
    return false;
  }
  return true;
}

template <int BN, int CG, int CONV, int MXK = 0>
static cudaError_t launch_cfg(const GemmProblem& g, cudaStream_t stream) {
  constexpr bool HALO = CONV == 2;
  using C = Cfg<BN, CG, MXK, 0, HALO ? 1 : 0>;
  KParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.epi = g.epi;
  static const bool timeline = getenv("FLUX2B_GEMM_TIMELINE") != nullptr;
  p.dbg = timeline ? 1 : 0;
  CUtensorMap tmA, tmB, tmSFA, tmSFB;
  int num_m_blks;
  if (CONV) {
    p.taps = g.conv_taps; p.H = g.H; p.W = g.W; p.Cin = g.Cin; p.batch = g.batch;
    p.up2 = (HALO && g.conv_up2) ? 1 : 0;
    static const bool wres_on = !(getenv("FLUX2B_CONV_WRES") && atoi(getenv("FLUX2B_CONV_WRES")) == 0);
    p.wres = (HALO && wres_on && !p.up2 && g.N <= BN && 9 * ((g.Cin + BK - 1) / BK) * C::B_TAP_BYTES <= C::STAGES * C::B_BYTES) ? 1 : 0;
    p.tiles_x = HALO ? (g.W + HALO_TW - 1) / HALO_TW : (g.W + CONV_TW - 1) / CONV_TW;
    p.tiles_y = HALO ? (g.H + HALO_TH - 1) / HALO_TH : (g.H + CONV_TH - 1) / CONV_TH;
    p.kb_per_tap = (g.Cin + BK - 1) / BK;
    p.num_kb = g.conv_taps * p.kb_per_tap;
    num_m_blks = g.batch * p.tiles_x * p.tiles_y;
    // activations NHWC: dims (C, W_in, H_in, N). Stride-2 convolutions (VAE encoder downsample: pad bottom / right only,
    // ResnetBlock.swift:203-213) traverse the input with element stride 2; out-of-bounds fill is the zero padding.
    const int st = g.conv_stride > 1 ? g.conv_stride : 1;
    const int Hin = g.Hin ? g.Hin : g.H * st, Win = g.Win ? g.Win : g.W * st;
    p.stride = st;
    p.tap_off = (g.conv_taps == 9 && st == 1) ? -1 : 0;
    uint64_t ad[4] = {(uint64_t)g.Cin, (uint64_t)Win, (uint64_t)Hin, (uint64_t)g.batch};
    uint64_t as[3] = {(uint64_t)g.lda * 2, (uint64_t)g.lda * 2 * Win, (uint64_t)g.lda * 2 * Win * Hin};
    uint32_t ab[4] = {BK, (uint32_t)(HALO ? HALO_W : CONV_TW), (uint32_t)(HALO ? HALO_H : CONV_TH), 1};
    uint32_t ae[4] = {1, (uint32_t)st, (uint32_t)st, 1};
    if (!make_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, g.A, 4, ad, as, ab, CU_TENSOR_MAP_SWIZZLE_128B, ae)) return cudaErrorInvalidValue;
    // weights OHWI: dims (Cin, taps, Cout[, 1])
    const int wtaps = (HALO && g.conv_up2) ? 16 : g.conv_taps;   // folded upsample: 4 phases x 4 pre-summed taps
    uint64_t bd[4] = {(uint64_t)g.Cin, (uint64_t)wtaps, (uint64_t)g.N, 1};
    uint64_t bs[3] = {(uint64_t)g.Cin * 2, (uint64_t)g.Cin * 2 * wtaps, (uint64_t)g.Cin * 2 * wtaps * g.N};
    uint32_t bb[4] = {BK, 1, (uint32_t)C::B_ROWS, 1};
    if (!make_tmap_bf16(&tmB, g.B, CG == 2 ? 4 : 3, bd, bs, bb)) return cudaErrorInvalidValue;
  } else if (MXK) {
    p.num_kb = g.K / C::KB_ELEMS;
    p.sfa = g.sfa; p.sfb = g.sfb;
    p.sfa_ld = g.sfa_ld ? g.sfa_ld : p.num_kb * C::SFPK;
    p.sfb_ld = g.sfb_ld ? g.sfb_ld : p.num_kb * C::SFPK;
    num_m_blks = (g.M + BM - 1) / BM;
    const uint64_t kbytes = (MXK == 1) ? (uint64_t)g.K : (uint64_t)g.K / 2;  // fp4: two elements per byte, dense
    uint64_t ad[2] = {kbytes, (uint64_t)g.M};
    uint64_t as[1] = {(uint64_t)g.lda};
    uint32_t ab[2] = {128, BM};
    if (!make_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_UINT8, g.A, 2, ad, as, ab)) return cudaErrorInvalidValue;
    uint64_t bd[2] = {kbytes, (uint64_t)g.N};
    uint64_t bs[1] = {(uint64_t)g.ldb};
    uint32_t bb[2] = {128, (uint32_t)C::B_ROWS};
    if (!make_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_UINT8, g.B, 2, bd, bs, bb)) return cudaErrorInvalidValue;
    if (CG == 2) {
      // scale factors as [512 B blocks][128 x u32]; extents are exact, so a block row past the operand (odd M-block count)
      // is zero-filled: scale 0 (E4M3) / 2^-127 (E8M0) times the zero-filled operand rows
      uint64_t sd[2] = {128, (uint64_t)(num_m_blks - 1) * p.sfa_ld + (uint64_t)p.num_kb * C::SFPK};
      uint64_t ss[1] = {512};
      uint32_t sb[2] = {128, (uint32_t)C::SFPK};
      if (!make_tmap(&tmSFA, CU_TENSOR_MAP_DATA_TYPE_UINT32, g.sfa, 2, sd, ss, sb, CU_TENSOR_MAP_SWIZZLE_NONE)) return cudaErrorInvalidValue;
      sd[1] = (uint64_t)(g.N / 128 - 1) * p.sfb_ld + (uint64_t)p.num_kb * C::SFPK;
      if (!make_tmap(&tmSFB, CU_TENSOR_MAP_DATA_TYPE_UINT32, g.sfb, 2, sd, ss, sb, CU_TENSOR_MAP_SWIZZLE_NONE)) return cudaErrorInvalidValue;
    }
  } else {
    p.num_kb = (g.K + BK - 1) / BK;
    num_m_blks = (g.M + BM - 1) / BM;
    uint64_t ad[2] = {(uint64_t)g.K, (uint64_t)g.M};
    uint64_t as[1] = {(uint64_t)g.lda * 2};
    uint32_t ab[2] = {BK, BM};
    if (!make_tmap_bf16(&tmA, g.A, 2, ad, as, ab)) return cudaErrorInvalidValue;
    uint64_t bd[2] = {(uint64_t)g.K, (uint64_t)(g.N - g.n_off)};
    uint64_t bs[1] = {(uint64_t)g.ldb * 2};
    uint32_t bb[2] = {BK, (uint32_t)C::B_ROWS};
    if (!make_tmap_bf16(&tmB, g.B, 2, bd, bs, bb)) return cudaErrorInvalidValue;
    p.n_off = g.n_off;
    if (g.B_lo) {
      if (!make_tmap_bf16(&tmSFB, g.B_lo, 2, bd, bs, bb)) return cudaErrorInvalidValue;
      p.split_units = g.M_lo / (BM * CG);
    }
  }
  p.num_m_units = (num_m_blks + CG - 1) / CG;
  p.num_n_blks = (g.N - g.n_off + BN - 1) / BN;
  const int total = p.num_m_units * p.num_n_blks * (p.up2 ? 4 : 1);
  const int max_units = g_num_sms / CG;
  const int units = std::min(total, max_units);
  if (!CONV) {
    // Tile order. The `units` tiles in flight at a time touch either ALL M units x a few N blocks (M fastest) or all N blocks x a
    // few M units (N fastest); the operand that is swept completely per wave has to stay in the 126 MB L2 from one wave to the next.
    // M fastest keeps A resident (fine while A = M x K fits beside a few weight tiles); for the K = 12288 out projections A is
    // 113 MB and was re-fetched from HBM by every wave (ncu: 484 MB read for 245 MB of operands) — there N fastest keeps the
    // 75 MB of weights resident and streams A once.
    static const int order = getenv("FLUX2B_GEMM_ORDER") ? atoi(getenv("FLUX2B_GEMM_ORDER")) : -1;   // -1 auto, 0 M fastest, 1 N fastest
    const double a_bytes = (double)g.M * g.K * (MXK ? 1 : 2), b_bytes = (double)(g.N - g.n_off) * g.K * (MXK ? 1 : 2) * (g.B_lo ? 2 : 1);
    const double waves_n = std::max(1.0, (double)units / p.num_m_units), waves_m = std::max(1.0, (double)units / p.num_n_blks);
    const double ws_mfast = a_bytes + b_bytes * std::min(1.0, waves_n / p.num_n_blks);   // all of A + the N blocks of one wave
    const double ws_nfast = b_bytes + a_bytes * std::min(1.0, waves_m / p.num_m_units);  // all of B + the M units of one wave
    const double l2 = 100e6;
    p.n_fast = order >= 0 ? order : (ws_mfast > l2 && ws_nfast < ws_mfast) ? 1 : 0;
  }

  auto kern = gemm_kernel<BN, CG, CONV, MXK>;
  static bool attr_set = false;  // per template instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(units * CG);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = CG;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  if (!(MXK && CG == 2)) { tmSFA = tmA; if (!(MXK == 0 && !CONV && g.B_lo)) tmSFB = tmA; }  // unused by those instantiations
  return cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmSFA, tmSFB, p);
}

// W-only quantized weights: 256-wide tiles, CTA pairs when there are at least two 128-row blocks
template <int CG>
static cudaError_t launch_wq(const GemmProblem& g, cudaStream_t stream) {
  constexpr int BN = 256;
  using C = Cfg<BN, CG, 0, 1>;
  KParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.epi = g.epi;
  p.num_kb = g.K / BK;
  const int bits = (g.wq == 1 || g.wq == 3) ? 8 : 4;
  const int group = g.wq <= 2 ? 64 : g.wq == 5 ? 16 : 32;
  p.wq_mode = g.wq; p.wq_sb_bf16 = g.wq_sb_bf16; p.wq_sb_ld = g.wq_sb_ld ? g.wq_sb_ld : g.K / group;
  p.wq_s = reinterpret_cast<const uint8_t*>(g.wq_scales); p.wq_b = reinterpret_cast<const uint8_t*>(g.wq_biases);
  p.wq_s_lo = reinterpret_cast<const uint8_t*>(g.wq_scales_lo); p.wq_b_lo = reinterpret_cast<const uint8_t*>(g.wq_biases_lo);
  {
    // 16 B loads of a row's scales cover 8 k-blocks (4 for nvfp4): rows and the K-slice start must be 16 B aligned, and the
    // k-block count a multiple of the span so that the last load of a row stays inside it
    const int esz = g.wq <= 2 ? 2 : 1, span = g.wq <= 2 ? 8 : g.wq == 5 ? 4 : 8;
    auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    p.wq_vec = ((size_t)p.wq_sb_ld * esz) % 16 == 0 && p.num_kb % span == 0 && al(g.wq_scales) && (g.wq > 2 || al(g.wq_biases)) &&
               (!g.B_lo || (al(g.wq_scales_lo) && (g.wq > 2 || al(g.wq_biases_lo))));
  }
  CUtensorMap tmA, tmB, tmBlo;
  const int num_m_blks = (g.M + BM - 1) / BM;
  uint64_t ad[2] = {(uint64_t)g.K, (uint64_t)g.M};
  uint64_t as[1] = {(uint64_t)g.lda * 2};
  uint32_t ab[2] = {BK, BM};
  if (!make_tmap_bf16(&tmA, g.A, 2, ad, as, ab)) return cudaErrorInvalidValue;
  // packed codes as bytes: one k-block of a row = 64 B (8-bit) or 32 B (4-bit); the slot is written 64 B- / 32 B-swizzled
  uint64_t bd[2] = {(uint64_t)g.K * bits / 8, (uint64_t)g.N};
  uint64_t bs[1] = {(uint64_t)g.ldb};
  uint32_t bb[2] = {(uint32_t)(bits == 8 ? 64 : 32), (uint32_t)C::B_ROWS};
  const CUtensorMapSwizzle swz = bits == 8 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  if (!make_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_UINT8, g.B, 2, bd, bs, bb, swz)) return cudaErrorInvalidValue;
  tmBlo = tmB;
  if (g.B_lo) {
    if (!make_tmap(&tmBlo, CU_TENSOR_MAP_DATA_TYPE_UINT8, g.B_lo, 2, bd, bs, bb, swz)) return cudaErrorInvalidValue;
    p.split_units = g.M_lo / (BM * CG);
  }
  p.num_m_units = (num_m_blks + CG - 1) / CG;
  p.num_n_blks = (g.N + BN - 1) / BN;
  const int total = p.num_m_units * p.num_n_blks;
  const int units = std::min(total, g_num_sms / CG);
  auto kern = gemm_kernel<BN, CG, 0, 0, 1>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(units * CG);
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = CG; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmA, tmBlo, p);
}

template <int CONV>
static cudaError_t dispatch(const GemmProblem& g, cudaStream_t s, int bn, int cg) {
#define F2B_CASE(BN_)                                                     \
  if (bn == BN_) {                                                        \
    if (cg == 2) return launch_cfg<BN_, 2, CONV>(g, s);                   \
    return launch_cfg<BN_, 1, CONV>(g, s);                                \
  }
  F2B_CASE(256)
  F2B_CASE(128)
  F2B_CASE(64)
  F2B_CASE(32)
  if constexpr (CONV != 0) {   // VAE channel counts (96, 192, 384) are multiples of 96, not of 128
    F2B_CASE(192)
    F2B_CASE(96)
  }
#undef F2B_CASE
  g_err = "unsupported BN";
  return cudaErrorInvalidValue;
}

// staged W-only GEMM: rows per N chunk. Default: the whole layer in one chunk. Chunks small enough to keep the stage L2-resident
// (option wq_stage_kb) were measured slower on Klein 9B int4 @1024^2 — 13.6 steps/s at 64 / 128 MB, 13.9 at 256 MB, 14.0 - 14.2 unchunked:
// the extra launches and the wave quantisation of the narrower GEMMs cost more than the saved HBM round trip of the stage.
static int wq_stage_chunk_rows(const GemmProblem& g) {
  if (g.wq_stage_kb <= 0) return std::max(256, (g.N + 255) / 256 * 256);
  const int64_t chunk_bytes = (int64_t)g.wq_stage_kb << 10;
  const int nchunks = (int)std::max<int64_t>(1, ((int64_t)g.N * g.K * 2 + chunk_bytes - 1) / chunk_bytes);
  return std::max(256, ((g.N + nchunks - 1) / nchunks + 255) / 256 * 256);
}
int gemm_launch_count(const GemmProblem& g) {
  if (!g.wq || !g.wq_stage) return 1;
  const int cn = wq_stage_chunk_rows(g);
  return ((g.N + cn - 1) / cn) * (1 + (g.B_lo ? 2 : 1));
}

cudaError_t gemm_launch(const GemmProblem& g, cudaStream_t stream) {
  if (!gemm_init()) {
    g_err = "gemm_init failed (driver entry point or device query)";
    return cudaErrorInitializationError;
  }
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return cudaSuccess;
  const bool conv = g.conv_taps != 0;
  if (g.mx) {
    const int kb_elems = g.mx == 1 ? 128 : 256;
    if (g.mx < 1 || g.mx > 3 || conv || g.K % kb_elems || g.N % 128 || g.lda % 16 || g.ldb % 16 || !g.sfa || !g.sfb ||
        (reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.B) & 15)) {
      g_err = "block-scaled GEMM needs K % 128 (fp8) / 256 (fp4) == 0, N % 128 == 0, 16 B aligned rows and both scale-factor tensors";
      return cudaErrorInvalidValue;
    }
    // N tile: 128 by default (two accumulator stages fit next to the scale-factor columns in TMEM, so the epilogue of one
    // tile overlaps the MMAs of the next); 256 on request. A SwiGLU producer must be launched with the tile its weight
    // rows were interleaved for (force_bn).
    const bool wide = g.force_bn == 256 && g.N % 256 == 0;
    if (g.epi.mode == EPI_SWIGLU && !g.force_bn) { g_err = "block-scaled SwiGLU epilogue needs the weight's tile size (force_bn)"; return cudaErrorInvalidValue; }
    // CTA pairs (cta_group::2, 256 x 128 tiles) by default: per MMA each tensor core then reads 128 A rows + 64 B rows from
    // shared memory instead of 128 + 128 — the single-CTA 128 x 128 tile sits exactly on the 128 B/clk smem read limit
    const bool pair = !wide && g.force_cta_group != 1 && (g.M + BM - 1) / BM >= 2;
    switch (g.mx * 4 + (wide ? 1 : 0) + (pair ? 2 : 0)) {
      case 4: return launch_cfg<128, 1, false, 1>(g, stream);
      case 5: return launch_cfg<256, 1, false, 1>(g, stream);
      case 6: return launch_cfg<128, 2, false, 1>(g, stream);
      case 8: return launch_cfg<128, 1, false, 2>(g, stream);
      case 9: return launch_cfg<256, 1, false, 2>(g, stream);
      case 10: return launch_cfg<128, 2, false, 2>(g, stream);
      case 12: return launch_cfg<128, 1, false, 3>(g, stream);
      case 13: return launch_cfg<256, 1, false, 3>(g, stream);
      default: return launch_cfg<128, 2, false, 3>(g, stream);
    }
  }
  if (g.wq) {
    const int group = g.wq <= 2 ? 64 : g.wq == 5 ? 16 : 32;
    if (g.wq < 1 || g.wq > 5 || conv || g.K % BK || g.lda % 8 || g.ldb % 16 || !g.wq_scales || (g.wq <= 2 && !g.wq_biases) ||
        (reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.B) & 15) ||
        (g.B_lo && (g.M_lo <= 0 || g.M_lo >= g.M || g.M_lo % (2 * BM) || (reinterpret_cast<uintptr_t>(g.B_lo) & 15) || !g.wq_scales_lo ||
                    (g.wq <= 2 && !g.wq_biases_lo) || g.epi.split_row != g.M_lo)) ||
        (g.epi.mode == EPI_SWIGLU && g.N % 256) || (g.wq_sb_ld && g.wq_sb_ld < g.K / group)) {
      g_err = "W-only quantized GEMM: plain GEMM, K % 64 == 0, 16 B aligned packed rows, scales (+ biases for the affine modes)";
      return cudaErrorInvalidValue;
    }
    if (g.wq_stage) {
      // staged: dequantize the layer once into the caller's 16-bit scratch, then the plain 16-bit kernel
      if ((g.B_lo && !g.wq_stage_lo) || (reinterpret_cast<uintptr_t>(g.wq_stage) & 15) || (reinterpret_cast<uintptr_t>(g.wq_stage_lo) & 15)) {
        g_err = "W-only staged GEMM: 16 B aligned stage buffers (one per weight set)";
        return cudaErrorInvalidValue;
      }
      // One chunk = the whole layer by default; with option wq_stage_kb the layer runs in N chunks that are dequantized into the SAME
      // stage (which then stays L2-resident) and multiplied by the plain kernel with a column window (n_off).
      const int esz_ = g.wq <= 2 ? 2 : 1;
      const int sb_ld = g.wq_sb_ld ? g.wq_sb_ld : g.K / group;
      const int cn_max = wq_stage_chunk_rows(g);
      for (int n0 = 0; n0 < g.N; n0 += cn_max) {
        const int cn = std::min(cn_max, g.N - n0);
        const int64_t nthreads = (int64_t)cn * (g.K / 8);
        const unsigned blocks = (unsigned)((nthreads + 255) / 256);
        for (int which = 0; which < (g.B_lo ? 2 : 1); ++which) {
          const uint8_t* pk = reinterpret_cast<const uint8_t*>(which ? g.B_lo : g.B) + (int64_t)n0 * g.ldb;
          const uint8_t* sc = reinterpret_cast<const uint8_t*>(which ? g.wq_scales_lo : g.wq_scales) + (int64_t)n0 * sb_ld * esz_;
          const uint8_t* bi0 = reinterpret_cast<const uint8_t*>(which ? g.wq_biases_lo : g.wq_biases);
          const uint8_t* bi = bi0 ? bi0 + (int64_t)n0 * sb_ld * esz_ : nullptr;
          uint16_t* out = reinterpret_cast<uint16_t*>(which ? g.wq_stage_lo : g.wq_stage);
          switch (g.wq) {
            case 1: wq_stage_kernel<1><<<blocks, 256, 0, stream>>>(pk, g.ldb, sc, bi, sb_ld, g.wq_sb_bf16, cn, g.K, out, g.epi.f16); break;
            case 2: wq_stage_kernel<2><<<blocks, 256, 0, stream>>>(pk, g.ldb, sc, bi, sb_ld, g.wq_sb_bf16, cn, g.K, out, g.epi.f16); break;
            case 3: wq_stage_kernel<3><<<blocks, 256, 0, stream>>>(pk, g.ldb, sc, bi, sb_ld, g.wq_sb_bf16, cn, g.K, out, g.epi.f16); break;
            case 4: wq_stage_kernel<4><<<blocks, 256, 0, stream>>>(pk, g.ldb, sc, bi, sb_ld, g.wq_sb_bf16, cn, g.K, out, g.epi.f16); break;
            default: wq_stage_kernel<5><<<blocks, 256, 0, stream>>>(pk, g.ldb, sc, bi, sb_ld, g.wq_sb_bf16, cn, g.K, out, g.epi.f16); break;
          }
          cudaError_t e = cudaGetLastError();
          if (e != cudaSuccess) return e;
        }
        GemmProblem h = g;
        h.wq = 0; h.wq_stage = nullptr; h.wq_stage_lo = nullptr;
        h.wq_scales = h.wq_biases = h.wq_scales_lo = h.wq_biases_lo = nullptr;
        h.B = g.wq_stage; h.ldb = g.K;
        if (g.B_lo) h.B_lo = g.wq_stage_lo;
        h.n_off = n0; h.N = n0 + cn;
        cudaError_t e = gemm_launch(h, stream);
        if (e != cudaSuccess) return e;
      }
      return cudaSuccess;
    }
    const int m_blks = (g.M + BM - 1) / BM;
    const bool pair = g.force_cta_group != 1 && m_blks >= 2;
    return pair ? launch_wq<2>(g, stream) : launch_wq<1>(g, stream);
  }
  if (!conv && (g.lda % 8 || g.ldb % 8)) { g_err = "lda/ldb must be multiples of 8 elements (TMA 16 B stride)"; return cudaErrorInvalidValue; }
  if (g.B_lo && (conv || g.M_lo <= 0 || g.M_lo >= g.M || g.M_lo % (2 * BM) || (reinterpret_cast<uintptr_t>(g.B_lo) & 15) || g.epi.split_row != g.M_lo)) {
    g_err = "two-problem GEMM: plain GEMM only, 0 < M_lo < M, M_lo a multiple of 256, epilogue split_row == M_lo";
    return cudaErrorInvalidValue;
  }
  if (conv && (g.Cin % 8 || g.lda % 8)) { g_err = "conv Cin / pixel stride must be multiples of 8"; return cudaErrorInvalidValue; }
  if ((reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.B) & 15)) { g_err = "A/B must be 16 B aligned"; return cudaErrorInvalidValue; }
  int bn = g.force_bn;
  const int Nw = g.N - g.n_off;   // columns this launch computes
  if (g.n_off && (conv || g.n_off < 0 || g.n_off >= g.N)) { g_err = "n_off: plain GEMM only, 0 <= n_off < N"; return cudaErrorInvalidValue; }
  if (!bn) bn = Nw > 128 ? 256 : Nw > 64 ? 128 : Nw > 32 ? 64 : 32;
  if (conv && !g.force_bn) {
    // N tile with the least padding (ties -> the wider tile): Cout = 384 -> 2 x 192 (256-wide tiles would compute 512 columns),
    // 192 -> 192, 96 -> 96. The small decoder's 96 / 192 / 384-channel layers lose a quarter of the tensor pipe otherwise.
    int best_pad = 1 << 30;
    for (int cand : {256, 192, 128, 96, 64, 32}) {
      const int pad = (g.N + cand - 1) / cand * cand;
      if (pad < best_pad) { best_pad = pad; bn = cand; }
    }
  }
  if (g.epi.mode == EPI_SWIGLU) bn = 256;
  if (g.epi.mode == EPI_QKV_ROPE && bn < 128) bn = 128;
  // CTA pairs (cta_group::2, 256 x BN tiles) by default: each CTA stages only half of B, which buys two more pipeline
  // stages and ~1/3 less L2 -> smem traffic per FLOP (measured +8..10 % over single-CTA tiles on the DiT shapes)
  int cg = g.force_cta_group;
  if (!cg) cg = 2;
  // 3x3 stride-1 convolutions take the halo-tile kernel (FLUX2B_CONV_HALO=0: one TMA box per tap, the cross-check)
  static const bool halo_on = !(getenv("FLUX2B_CONV_HALO") && atoi(getenv("FLUX2B_CONV_HALO")) == 0);
  const bool halo = conv && (g.conv_up2 || (halo_on && g.conv_taps == 9 && g.conv_stride <= 1 && !g.conv_no_halo));
  if (g.conv_up2 && (g.conv_taps != 9 || g.conv_stride > 1)) { g_err = "folded upsample: 3x3 stride-1 convolution only"; return cudaErrorInvalidValue; }
  const int m_blks = !conv ? (g.M + BM - 1) / BM
                     : halo ? g.batch * ((g.W + HALO_TW - 1) / HALO_TW) * ((g.H + HALO_TH - 1) / HALO_TH)
                            : g.batch * ((g.W + CONV_TW - 1) / CONV_TW) * ((g.H + CONV_TH - 1) / CONV_TH);
  if (cg == 2 && (m_blks < 2 || bn < 32)) cg = 1;
  if (!conv && !g.force_bn && !g.force_cta_group && bn == 256 && g.epi.mode != EPI_SWIGLU && Nw % 128 == 0) {
    // Few-row problems (the text encoder's M = 512 prefill: 256-wide tiles of an N = 2560 projection occupy 40 of 148 SMs):
    // pick the narrower tile when a simple wave model says it is clearly faster. Per k-step of 16 a CTA needs
    // max(MMA clocks = bn / 2, shared-memory operand reads = (128 + bn / cg) / 4) clocks; a wave costs ~3000 clocks of
    // pipeline fill + epilogue on top. The DiT shapes (M = 4608, many waves) keep the measured-best 256 x 256 pair tile.
    auto est = [&](int bn_, int cg_) {
      const long units = (long)((g.M + BM * cg_ - 1) / (BM * cg_)) * ((Nw + bn_ - 1) / bn_);
      const long slots = g_num_sms / cg_;
      const double per = (g.K / 16.0) * std::max(bn_ / 2.0, (128.0 + bn_ / cg_) / 4.0) + 3000.0;
      return (double)((units + slots - 1) / slots) * per;
    };
    const double cur = est(bn, cg);
    int best_bn = bn, best_cg = cg;
    double best = cur;
    for (int cb : {128, 64}) {
      if (g.epi.mode == EPI_QKV_ROPE && cb < 128) continue;
      for (int cc : {2, 1}) {
        if (cc == 2 && m_blks < 2) continue;
        const double e = est(cb, cc);
        if (e < best) { best = e; best_bn = cb; best_cg = cc; }
      }
    }
    if (best < 0.8 * cur) { bn = best_bn; cg = best_cg; }
  }
  return halo ? dispatch<2>(g, stream, bn, cg) : conv ? dispatch<1>(g, stream, bn, cg) : dispatch<0>(g, stream, bn, cg);
}

}  // namespace f2b
