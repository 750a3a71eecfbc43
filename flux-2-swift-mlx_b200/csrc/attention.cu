// attention.cu — fused joint (text+image) flash attention for sm_100a, head_dim 128, no mask.
//
//   O = softmax(Q K^T / sqrt(128)) V          (reference: MLXFast.scaledDotProductAttention call sites
//                                              Flux2Attention.swift:168-174, Flux2ParallelAttention.swift:104-110;
//                                              KV-cached variant Flux2Attention.swift:393-398)
//
// One CTA owns 256 query rows of one head: two 128-row Q tiles, each served by its own softmax warpgroup, share
// every K/V tile (ping-pong, so the tensor pipe works on one tile while the other is in softmax).
//   warps 0-3 / 4-7 : softmax warpgroup 0 / 1 (thread == query row == TMEM lane)
//   warp 8          : TMA producer (Q once, K/V ring)
//   warp 9          : tcgen05.mma issuer + TMEM allocator
// TMEM (512 columns): S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512); fp32 accumulators never touch registers
// except for the online-softmax correction of O (done only when a row maximum actually moved).
// K is consumed K-major (d contiguous), V MN-major (d contiguous, keys along the MMA K dimension) straight from the
// row-major [tokens, heads*128] activation matrices — the [txt|img] concatenation, the per-head split and the
// "extra cached reference K/V" of the klein-9b-kv path are all expressed as TMA coordinates / key segments, no copies.
// Two variants of the P operand: kPTmem=false stages P (bf16) through swizzled shared memory with 64-key tiles,
// kPTmem=true writes P back into the S columns of TMEM and issues the PV MMA with A from TMEM (128-key tiles).
#include "attention.cuh"
#include "ptx.cuh"
#include "gemm.cuh"

#include <cstdlib>
#include <string>

namespace f2b {

extern bool gemm_init();
bool make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box);

__device__ __forceinline__ uint32_t apk2(float a, float b, int f16) {
  if (f16) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  return pack_bf16x2(a, b);
}
static constexpr int HD = 128;  // head dim
static constexpr int QT = 128;  // query rows per softmax warpgroup

struct AttnKParams {
  CUtensorMap tmQ;
  CUtensorMap tmK[3];
  CUtensorMap tmV[3];
  int sq, num_heads;
  int spin;
  float scale_log2;
  int q_row0;
  long long q_bs;
  uint16_t* o;
  long long ldo;
  int o_row0;
  long long o_bs;
  int f16;
  int nseg;
  int seg_row0[3];
  int seg_len[3];
  long long seg_bs[3];
  int o_rows_per_peer, o_col0;
  uint16_t* o_peer[8];
  // variant 3 only — causal / padded / grouped-query mode of the text-encoder prefill (te.cu):
  int causal;        // key j is visible to query i iff j <= i (query index == key index, one segment)
  int kv_group;      // query head h reads K / V head h / kv_group
  int key_lo, key_hi;  // keys outside [key_lo, key_hi) get the additive padding mask pad_bias (key_hi == 0: no padding mask)
  const int* mask_dev; // if set: {key_lo, key_hi} are read from device memory (a captured CUDA graph is replayed for every prompt length)
  float pad_raw;     // raw (pre-scale) score given to padded keys: -(2^k) with 2^k >= |pad_bias| / scale
  int pingpong;      // variant 3: alternate the two warpgroups' exponential phases (default 1; FLUX2B_ATTN_PINGPONG=0 turns it off)
  int dbg;  // FLUX2B_ATTN_TIMELINE=1: CTA (0,0,0) prints its softmax / MMA time line (debug aid, off by default)
};

static bool fill_params(const struct AttnProblem& a, int BN, AttnKParams& p);

template <int BN, bool kPTmem>
struct ACfg {
  static constexpr int STAGES = kPTmem ? 2 : 3;
  static constexpr int Q_BYTES = QT * HD * 2;        // 32 KB per Q tile (2 panels of 128 rows x 128 B)
  static constexpr int KV_BYTES = BN * HD * 2;       // per K or V tile (2 panels of BN rows x 128 B)
  static constexpr int KV_PANEL = BN * 128;          // bytes between the d[0,64) and d[64,128) panels
  static constexpr int P_BYTES = kPTmem ? 0 : QT * BN * 2;
  static constexpr int OFF_K = 2 * Q_BYTES;
  static constexpr int OFF_V = OFF_K + STAGES * KV_BYTES;
  static constexpr int OFF_P = OFF_V + STAGES * KV_BYTES;
  static constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr int THREADS = 352;   // 8 softmax warps, TMA producer, two MMA issuers
};

template <int BN, bool kPTmem>
__global__ void __launch_bounds__(320, 1) attn_kernel(const __grid_constant__ AttnKParams p) {
  using C = ACfg<BN, kPTmem>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* q_full = bars;                      // [2]
  uint64_t* k_full = bars + 2;                  // [STAGES]
  uint64_t* v_full = k_full + C::STAGES;        // [STAGES]
  uint64_t* kv_empty = v_full + C::STAGES;      // [STAGES]
  uint64_t* s_ready = kv_empty + C::STAGES;     // [2]
  uint64_t* p_ready = s_ready + 2;              // [2]
  uint64_t* o_done = p_ready + 2;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int q_blk0 = blockIdx.x * 2 * QT;  // first query row (within the batch item) of this CTA

  // total number of key tiles over all segments
  int n_tiles = 0;
  for (int s = 0; s < p.nseg; ++s) n_tiles += (p.seg_len[s] + BN - 1) / BN;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    for (int s = 0; s < p.nseg; ++s) { tma_prefetch_desc(&p.tmK[s]); tma_prefetch_desc(&p.tmV[s]); }
  }
  if (warp == 9 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&s_ready[i], 1);
      mbar_init(&p_ready[i], 128);
      mbar_init(&o_done[i], 1);
    }
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ============================================================ TMA producer
    if (lane == 0) {
      const int qrow = p.q_row0 + (int)(b * p.q_bs) + q_blk0;
      for (int w = 0; w < 2; ++w) {
        mbar_expect_tx(&q_full[w], C::Q_BYTES);
        uint8_t* dst = smem + w * C::Q_BYTES;
        tma_load_2d(dst, &p.tmQ, &q_full[w], head * HD, qrow + w * QT);
        tma_load_2d(dst + QT * 128, &p.tmQ, &q_full[w], head * HD + 64, qrow + w * QT);
      }
      int j = 0;
      for (int s = 0; s < p.nseg; ++s) {
        const int tiles = (p.seg_len[s] + BN - 1) / BN;
        const int row_base = p.seg_row0[s] + (int)(b * p.seg_bs[s]);
        for (int t = 0; t < tiles; ++t, ++j) {
          const int st = j % C::STAGES;
          const uint32_t ph = (j / C::STAGES) & 1;
          mbar_wait(&kv_empty[st], ph ^ 1, 10);
          uint8_t* kd = smem + C::OFF_K + st * C::KV_BYTES;
          uint8_t* vd = smem + C::OFF_V + st * C::KV_BYTES;
          const int row = row_base + t * BN;
          mbar_expect_tx(&k_full[st], C::KV_BYTES);
          tma_load_2d(kd, &p.tmK[s], &k_full[st], head * HD, row);
          tma_load_2d(kd + C::KV_PANEL, &p.tmK[s], &k_full[st], head * HD + 64, row);
          mbar_expect_tx(&v_full[st], C::KV_BYTES);
          tma_load_2d(vd, &p.tmV[s], &v_full[st], head * HD, row);
          tma_load_2d(vd + C::KV_PANEL, &p.tmV[s], &v_full[st], head * HD + 64, row);
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ============================================================ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_f16(QT, BN, p.f16 == 0, false, false);  // S = Q K^T   (both K-major)
      const uint32_t idesc_o = make_idesc_f16(QT, HD, p.f16 == 0, false, true);   // O += P V    (V MN-major)
      auto issue_s = [&](int w, int st) {
        const uint32_t qa = smem_u32(smem + w * C::Q_BYTES);
        const uint32_t ka = smem_u32(smem + C::OFF_K + st * C::KV_BYTES);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint64_t ad = make_smem_desc(qa + (k / 4) * (QT * 128) + (k % 4) * 32, 16, 1024, SWZ_128B);
          const uint64_t bd = make_smem_desc(ka + (k / 4) * C::KV_PANEL + (k % 4) * 32, 16, 1024, SWZ_128B);
          umma_f16_ss<1>(tmem_base + w * 128, ad, bd, idesc_s, k ? 1u : 0u);
        }
        umma_commit(&s_ready[w]);
      };
      auto issue_pv = [&](int w, int st, bool accumulate) {
        const uint32_t va = smem_u32(smem + C::OFF_V + st * C::KV_BYTES);
#pragma unroll
        for (int k = 0; k < BN / 16; ++k) {
          // B = V tile, MN-major: 64-wide d panels KV_PANEL bytes apart (LBO), 8-key groups 1024 B apart (SBO);
          // 16 keys per MMA = 2048 B along the key (MMA-K) direction.
          const uint64_t bd = make_smem_desc(va + k * 2048, C::KV_PANEL, 1024, SWZ_128B);
          const uint32_t acc = (accumulate || k) ? 1u : 0u;
          if constexpr (kPTmem) {
            umma_f16_ts(tmem_base + 256 + w * 128, tmem_base + w * 128 + k * 8, bd, idesc_o, acc);
          } else {
            const uint32_t pa = smem_u32(smem + C::OFF_P + w * C::P_BYTES);
            const uint64_t ad = make_smem_desc(pa + (k / 4) * (QT * 128) + (k % 4) * 32, 16, 1024, SWZ_128B);
            umma_f16_ss<1>(tmem_base + 256 + w * 128, ad, bd, idesc_o, acc);
          }
        }
        umma_commit(&o_done[w]);
      };
      mbar_wait(&q_full[0], 0, 20);
      mbar_wait(&k_full[0], 0, 21);
      tc_fence_after();
      issue_s(0, 0);
      mbar_wait(&q_full[1], 0, 22);
      tc_fence_after();
      issue_s(1, 0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % C::STAGES;
        const uint32_t ph = (j / C::STAGES) & 1;
        const int stn = (j + 1) % C::STAGES;
        const uint32_t phn = ((j + 1) / C::STAGES) & 1;
        for (int w = 0; w < 2; ++w) {
          mbar_wait(&p_ready[w], j & 1, 23);
          if (w == 0) mbar_wait(&v_full[st], ph, 24);
          tc_fence_after();
          issue_pv(w, st, j > 0);
          if (w == 1) umma_commit(&kv_empty[st]);
          if (j + 1 < n_tiles) {
            if (w == 0) mbar_wait(&k_full[stn], phn, 25);
            tc_fence_after();
            issue_s(w, stn);
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ============================================================ softmax warpgroups
    const int w = warp >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;  // row inside the Q tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t t_s = tmem_base + lane_off + w * 128;
    const uint32_t t_o = tmem_base + lane_off + 256 + w * 128;
    const float sl2 = p.scale_log2;
    float m_run = -INFINITY, l_run = 0.f;
    uint32_t v[32];
    int j = 0;
    for (int s = 0; s < p.nseg; ++s) {
      const int tiles = (p.seg_len[s] + BN - 1) / BN;
      for (int t = 0; t < tiles; ++t, ++j) {
        const int nvalid = min(BN, p.seg_len[s] - t * BN);
        mbar_wait(&s_ready[w], j & 1, 30);
        tc_fence_after();
        // ---- pass 1: row maximum
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          tmem_ld_32x32(t_s + c * 32, v);
          tmem_ld_wait();
          if (nvalid >= (c + 1) * 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < nvalid) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
        }
        const float m_new = fmaxf(m_run, mx * sl2);
        const float alpha = fast_exp2(m_run - m_new);  // first tile: exp2(-inf) = 0
        if (j > 0) {
          // O of the previous tile must have landed before it is corrected / before P storage is reused
          mbar_wait(&o_done[w], (j - 1) & 1, 31);
          tc_fence_after();
          if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
            for (int c = 0; c < HD / 32; ++c) {
              tmem_ld_32x32(t_o + c * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
              tmem_st_32x32(t_o + c * 32, v);
            }
            tmem_st_wait();
          }
        }
        // ---- pass 2: P = exp2(S*scale - m), row sum, hand P to the tensor core
        float rs = 0.f;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          tmem_ld_32x32(t_s + c * 32, v);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float p0 = fast_exp2(fmaf(__uint_as_float(v[2 * i]), sl2, -m_new));
            float p1 = fast_exp2(fmaf(__uint_as_float(v[2 * i + 1]), sl2, -m_new));
            if (c * 32 + 2 * i >= nvalid) p0 = 0.f;
            if (c * 32 + 2 * i + 1 >= nvalid) p1 = 0.f;
            rs += p0 + p1;
            pk[i] = apk2(p0, p1, p.f16);
          }
          if constexpr (kPTmem) {
            tmem_st_32x16(t_s + c * 16, pk);
          } else {
            // K-major 128B-swizzled A tile: row r at r*128 B inside a 64-key panel, 16 B chunk index ^= (r & 7)
            uint8_t* pbase = smem + C::OFF_P + w * C::P_BYTES + ((c * 32) / 64) * (QT * 128) + row * 128;
            const int ch0 = ((c * 32) % 64) / 8;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const int ch = (ch0 + q4) ^ (row & 7);
              *reinterpret_cast<uint4*>(pbase + ch * 16) =
                  make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
            }
          }
        }
        l_run = l_run * alpha + rs;
        m_run = m_new;
        if constexpr (kPTmem) {
          tmem_st_wait();
        } else {
          fence_proxy_async_smem();
        }
        tc_fence_before();
        mbar_arrive(&p_ready[w]);
      }
    }
    // ---- finalize: O / l
    mbar_wait(&o_done[w], (n_tiles - 1) & 1, 32);
    tc_fence_after();
    const int qrow = q_blk0 + w * QT + row;
    const bool ok = qrow < p.sq;
    const float inv_l = 1.0f / l_run;
    uint16_t* orow = p.o + ((long long)p.o_row0 + b * p.o_bs + qrow) * p.ldo + head * HD;
#pragma unroll 1
    for (int c = 0; c < HD / 32; ++c) {
      tmem_ld_32x32(t_o + c * 32, v);
      tmem_ld_wait();
      if (ok) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 u;
          u.x = apk2(__uint_as_float(v[8 * q4 + 0]) * inv_l, __uint_as_float(v[8 * q4 + 1]) * inv_l, p.f16);
          u.y = apk2(__uint_as_float(v[8 * q4 + 2]) * inv_l, __uint_as_float(v[8 * q4 + 3]) * inv_l, p.f16);
          u.z = apk2(__uint_as_float(v[8 * q4 + 4]) * inv_l, __uint_as_float(v[8 * q4 + 5]) * inv_l, p.f16);
          u.w = apk2(__uint_as_float(v[8 * q4 + 6]) * inv_l, __uint_as_float(v[8 * q4 + 7]) * inv_l, p.f16);
          *reinterpret_cast<uint4*>(orow + c * 32 + q4 * 8) = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<1>(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ variant 3
// Same ping-pong structure as variant 2 (128-key tiles, P written back into the S columns of TMEM, PV with A from TMEM)
// with the softmax warpgroup reorganised around the two limits measured on the first kernels (ncu: tensor pipe 24 %,
// issue slots 41 %, MUFU 25 % busy):
//   * S is read from TMEM ONCE per tile (four 32-column loads in flight, one wait), row max / exp2 / pack run on
//     registers with independent accumulators;
//   * lazy rescaling: the running maximum only moves when it grows by more than 2^8 (P stays < 256, exact in fp32 /
//     bf16 range), so the O read-modify-write in TMEM happens a handful of times per row instead of every tile;
//   * no per-tile wait on the PV barrier: s_ready(j) already orders after PV(j-1) (tcgen05.commit tracks every MMA the
//     issuing thread issued before it), only the last tile signals o_done;
//   * separate K (3-deep) and V (2-deep) rings: K(j+1) is needed one softmax earlier than V(j+1).
struct A3 {
  static constexpr int BN = 128;
  static constexpr int KS = 3, VS = 2;
  static constexpr int Q_BYTES = QT * HD * 2;
  static constexpr int KV_BYTES = BN * HD * 2;
  static constexpr int KV_PANEL = BN * 128;
  static constexpr int OFF_K = 2 * Q_BYTES;
  static constexpr int OFF_V = OFF_K + KS * KV_BYTES;
  static constexpr int OFF_BAR = OFF_V + VS * KV_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr int THREADS = 352;   // 8 softmax warps, TMA producer, two MMA issuers
};

// 2^x on the FMA / ALU pipes instead of the MUFU: round-to-nearest split x = n + f by the 1.5 * 2^23 trick, a degree-3
// minimax polynomial for 2^f on [-0.5, 0.5] (max relative error 7.5e-5 — far below the 16-bit rounding of P), n added into
// the exponent field. x <= ~8 here (scores minus the running maximum); the clamp keeps -inf / very negative scores at 2^-126.
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float r = fmaf(f, 0.0551716648f, 0.2426111251f);
  r = fmaf(r, f, 0.6932609677f);
  r = fmaf(r, f, 0.9999280572f);
  return __uint_as_float(__float_as_uint(r) + (__float_as_uint(t) << 23));
}

__device__ __forceinline__ bool getenv_dbg2(const AttnKParams& p) { return p.dbg >= 2; }

// two elements at a time with the packed fp32x2 instructions
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.0f);
  x.y = fmaxf(x.y, -126.0f);
  const float2 magic = make_float2(12582912.0f, 12582912.0f);
  const float2 t = __fadd2_rn(x, magic);
  const float2 r = __fadd2_rn(t, make_float2(-12582912.0f, -12582912.0f));
  const float2 f = __ffma2_rn(r, make_float2(-1.0f, -1.0f), x);
  float2 q = __ffma2_rn(f, make_float2(0.0551716648f, 0.0551716648f), make_float2(0.2426111251f, 0.2426111251f));
  q = __ffma2_rn(q, f, make_float2(0.6932609677f, 0.6932609677f));
  q = __ffma2_rn(q, f, make_float2(0.9999280572f, 0.9999280572f));
  return make_float2(__uint_as_float(__float_as_uint(q.x) + (__float_as_uint(t.x) << 23)),
                     __uint_as_float(__float_as_uint(q.y) + (__float_as_uint(t.y) << 23)));
}

// kPoly: 0 = every exponential on the MUFU (ex2.approx); n > 0 = one element in n takes exp2_poly. At head dim 128 the MUFU
// (16 ex2 / clk / SM: 1024 clk per 128 x 128 tile) is as busy as the tensor pipe (QK^T + PV: 1024 clk), so moving a share of the
// exponentials onto the otherwise idle FMA lanes is what lets the MMAs run closer to back to back.
template <bool kF16, int kPoly, bool kMask = false>
__global__ void __launch_bounds__(A3::THREADS, 1) attn_kernel_v3(const __grid_constant__ AttnKParams p) {
  using C = A3;
  constexpr int BN = C::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* q_full = bars;               // [2]
  uint64_t* k_full = q_full + 2;         // [KS]
  uint64_t* k_empty = k_full + C::KS;    // [KS]
  uint64_t* v_full = k_empty + C::KS;    // [VS]
  uint64_t* v_empty = v_full + C::VS;    // [VS]
  uint64_t* s_ready = v_empty + C::VS;   // [2]
  uint64_t* p_ready = s_ready + 2;       // [2]
  uint64_t* o_done = p_ready + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int q_blk0 = blockIdx.x * 2 * QT;

  // causal: this CTA's 256 query rows see keys [0, q_blk0 + 256) at most (tiles above the diagonal are never loaded)
  // right padding (key_lo == 0): keys >= key_hi carry exactly zero weight for every row (key 0 is always visible), so they are not loaded either
  // (kMask: the causal / padded / grouped-query mode is a separate instantiation; the DiT kernel carries none of it)
  int key_lo = 0, key_hi = 0;
  if constexpr (kMask) {
    key_lo = p.key_lo; key_hi = p.key_hi;
    if (p.mask_dev) { key_lo = __ldg(p.mask_dev); key_hi = __ldg(p.mask_dev + 1); }
  }
  auto seg_len_of = [&](int s) {
    int n = p.seg_len[s];
    if constexpr (kMask) {
      if (p.causal && s == 0) n = min(n, q_blk0 + 2 * QT);
      if (key_hi > 0 && key_lo == 0 && s == 0) n = min(n, key_hi);
    }
    return n;
  };
  const int kv_col = kMask ? (head / p.kv_group) * HD : head * HD;
  int n_tiles = 0;
  for (int s = 0; s < p.nseg; ++s) n_tiles += (seg_len_of(s) + BN - 1) / BN;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    for (int s = 0; s < p.nseg; ++s) { tma_prefetch_desc(&p.tmK[s]); tma_prefetch_desc(&p.tmV[s]); }
  }
  if (warp == 9 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&s_ready[i], 1);
      mbar_init(&p_ready[i], 128);
      mbar_init(&o_done[i], 1);
    }
    // a K / V stage is free once BOTH issuers' MMAs on it have completed (tcgen05.commit tracks the committing thread's own MMAs)
    for (int i = 0; i < C::KS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 2); }
    for (int i = 0; i < C::VS; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 2); }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<1>(tmem_slot, 512);
  __shared__ long long t0_shared;   // common origin of the debug timeline (all warps of a CTA read the same SM clock)
  if (threadIdx.x == 0) t0_shared = clock64();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool dbg0 = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  pdl_trigger();   // programmatic dependent launch: the prologue above overlapped the previous kernel's tail
  pdl_wait();

  if (warp == 8) {
    // ============================================================ TMA producer
    if (lane == 0) {
      int pk_t[8], pv_w[8], pv_t[8];
      const int qrow = p.q_row0 + (int)(b * p.q_bs) + q_blk0;
      for (int w = 0; w < 2; ++w) {
        mbar_expect_tx(&q_full[w], C::Q_BYTES);
        uint8_t* dst = smem + w * C::Q_BYTES;
        tma_load_2d(dst, &p.tmQ, &q_full[w], head * HD, qrow + w * QT);
        tma_load_2d(dst + QT * 128, &p.tmQ, &q_full[w], head * HD + 64, qrow + w * QT);
      }
      // K runs one tile ahead of V: K(j + 1) is needed for S(j + 1), which is issued right behind PV(j), and its ring
      // (3 stages) is free long before; waiting for V(j)'s slot (2 stages) first would hold it back by a whole period.
      auto tile_at = [&](int j, int* seg, int* row) {
        for (int s = 0; s < p.nseg; ++s) {
          const int tiles = (seg_len_of(s) + BN - 1) / BN;
          if (j < tiles) { *seg = s; *row = p.seg_row0[s] + (int)(b * p.seg_bs[s]) + j * BN; return; }
          j -= tiles;
        }
      };
      auto load_k = [&](int j) {
        int s = 0, row = 0;
        tile_at(j, &s, &row);
        const int ks = j % C::KS;
        mbar_wait(&k_empty[ks], ((j / C::KS) & 1) ^ 1, 10);
        uint8_t* kd = smem + C::OFF_K + ks * C::KV_BYTES;
        mbar_expect_tx(&k_full[ks], C::KV_BYTES);
        tma_load_2d(kd, &p.tmK[s], &k_full[ks], kv_col, row);
        tma_load_2d(kd + C::KV_PANEL, &p.tmK[s], &k_full[ks], kv_col + 64, row);
        if (dbg0 && j < 8) pk_t[j] = (int)(clock64() - t0_shared);
      };
      load_k(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) load_k(j + 1);
        int s = 0, row = 0;
        tile_at(j, &s, &row);
        const int vs = j % C::VS;
        if (dbg0 && j < 8) pv_w[j] = (int)(clock64() - t0_shared);
        mbar_wait(&v_empty[vs], ((j / C::VS) & 1) ^ 1, 11);
        uint8_t* vd = smem + C::OFF_V + vs * C::KV_BYTES;
        mbar_expect_tx(&v_full[vs], C::KV_BYTES);
        tma_load_2d(vd, &p.tmV[s], &v_full[vs], kv_col, row);
        tma_load_2d(vd + C::KV_PANEL, &p.tmV[s], &v_full[vs], kv_col + 64, row);
        if (dbg0 && j < 8) pv_t[j] = (int)(clock64() - t0_shared);
      }
      if (dbg0)
        for (int j = 0; j < 8 && j < n_tiles; ++j)
          printf("[attn timeline] producer: tile %d  K issued at %6d  V slot awaited from %6d, V issued at %6d\n", j, pk_t[j], pv_w[j], pv_t[j]);
    }
    __syncwarp();
  } else if (warp == 9 || warp == 10) {
    // ============================================================ MMA issuers: one warp per query tile / softmax warpgroup
    // One issuing warp for both tiles was the kernel's limiter: per key tile it spent ~2 x 1030 clk blocked in the issue of
    // its 32 MMAs (the tensor pipe's queue is only a few MMAs deep) plus ~1000 clk in four mbarrier waits during which the
    // pipe ran dry (timeline: period 3100 clk against 2048 clk of MMA work). With one issuer per tile the waits of one warp
    // overlap the issue of the other; the tensor pipe executes the two streams in arrival order, and they only share the K / V
    // stages, whose release barriers therefore count two commits.
    // The whole warp walks the loop (warp-uniform control flow, barrier waits by all lanes) and one elected lane issues:
    // inside an `if (lane == 0)` region ptxas wraps every tcgen05.mma in an elect / broadcast loop and reloads its
    // operands from the stack (~25 instructions per MMA), which is slower than the 64 clk a 128x128x16 MMA takes.
    // The 512-column allocation starts at TMEM address 0 (checked below), so accumulator addresses are immediates, and
    // descriptors are (constant high word, 32-bit low word = address field + constants).
    if (tmem_base != 0) __trap();
    const int w = warp - 9;
    const uint32_t idesc_s = make_idesc_f16(QT, BN, !kF16, false, false);
    const uint32_t idesc_o = make_idesc_f16(QT, HD, !kF16, false, true);
    const uint32_t smem_base = smem_u32(smem);
    // K-major operands (Q, K): LBO 16 B, SBO 1024 B; V (MN-major): LBO = panel stride, SBO 1024 B
    const uint64_t desc_kmajor = make_smem_desc(0, 16, 1024, SWZ_128B);
    const uint64_t desc_v = make_smem_desc(0, C::KV_PANEL, 1024, SWZ_128B);
    auto issue_s = [&](int ks) {
      const uint32_t qa = (smem_base + w * C::Q_BYTES) >> 4;
      const uint32_t ka = (smem_base + C::OFF_K + ks * C::KV_BYTES) >> 4;
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) {
        const uint32_t off_q = ((k / 4) * (QT * 128) + (k % 4) * 32) >> 4;
        const uint32_t off_k = ((k / 4) * C::KV_PANEL + (k % 4) * 32) >> 4;
        umma_f16_ss<1>(w * 128, desc_kmajor + (qa + off_q), desc_kmajor + (ka + off_k), idesc_s, k ? 1u : 0u);
      }
      umma_commit(&s_ready[w]);
    };
    auto issue_pv = [&](int vs, bool accumulate) {
      const uint32_t va = (smem_base + C::OFF_V + vs * C::KV_BYTES) >> 4;
#pragma unroll
      for (int k = 0; k < BN / 16; ++k)
        umma_f16_ts(256 + w * 128, w * 128 + k * 8, desc_v + (va + k * (2048 >> 4)), idesc_o, (accumulate || k) ? 1u : 0u);
    };
    mbar_wait(&q_full[w], 0, 20);
    mbar_wait(&k_full[0], 0, 21);
    tc_fence_after();
    if (elect_one()) { issue_s(0); umma_commit(&k_empty[0]); }
    __syncwarp();
    const bool mdbg = dbg0;
    const long long mt0 = t0_shared;
    int m_seen[8], m_kv[8], m_iss[8], m_v[8];
    for (int j = 0; j < n_tiles; ++j) {
      const int vs = j % C::VS;
      const int ksn = (j + 1) % C::KS;
      const bool more = j + 1 < n_tiles;
      // operand barriers first: they completed long ago, but every mbarrier wait costs ~200 clk of latency, which
      // must not sit between "P is ready" and the MMAs that consume it
      mbar_wait(&v_full[vs], (j / C::VS) & 1, 24);
      if (mdbg && j < 8) m_v[j] = (int)(clock64() - mt0);
      if (more) mbar_wait(&k_full[ksn], ((j + 1) / C::KS) & 1, 25);
      if (p.spin) mbar_spin(&p_ready[w], j & 1, 23); else mbar_wait(&p_ready[w], j & 1, 23);
      if (mdbg && j < 8) m_seen[j] = (int)(clock64() - mt0);
      if (mdbg && j < 8) m_kv[j] = (int)(clock64() - mt0);
      tc_fence_after();
      if (elect_one()) {
        issue_pv(vs, j > 0);
        umma_commit(&v_empty[vs]);
        if (more) {
          issue_s(ksn);
          umma_commit(&k_empty[ksn]);
        } else {
          umma_commit(&o_done[w]);
        }
      }
      __syncwarp();
      if (mdbg && j < 8) m_iss[j] = (int)(clock64() - mt0);
      if (mdbg && more && getenv_dbg2(p)) {  // debug level 2: block until the S MMAs just issued have completed (perturbs the schedule)
        mbar_wait(&s_ready[w], (j + 1) & 1, 26);
        if (j < 8) m_kv[j] = (int)(clock64() - mt0);   // reuse the slot: completion time
      }
    }
    if (mdbg && lane == 0)
      for (int j = 0; j < 8 && j < n_tiles; ++j)
        printf("[attn timeline] mma warp: wg %d tile %d  V full at %6d  P seen at %6d  K/V ready at %6d  PV + next S issued at %6d\n",
               w, j, m_v[j], m_seen[j], m_kv[j], m_iss[j]);
  } else if (warp < 8) {
    // ============================================================ softmax warpgroups
    const int w = warp >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t t_s = tmem_base + lane_off + w * 128;
    const uint32_t t_o = tmem_base + lane_off + 256 + w * 128;
    const float sl2 = p.scale_log2;
    float m_run = -INFINITY, l_run = 0.f;  // m_run in the scaled (log2) domain
    const bool dbg = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (threadIdx.x & 127) == 0;
    const long long t_start = t0_shared;
    int tl_wake[16], tl_done[16], tl_ld[16], tl_max[16], tl_exp[16];
    int j = 0;
    for (int s = 0; s < p.nseg; ++s) {
      const int tiles = (seg_len_of(s) + BN - 1) / BN;
      for (int t = 0; t < tiles; ++t, ++j) {
        const int nvalid = min(BN, seg_len_of(s) - t * BN);
        // Ping-pong: the tiles of the two warpgroups alternate strictly (warpgroup 0 first). Left alone — with an issuer warp each —
        // they fall into lock-step: both in the MUFU-bound exponential phase at once (each 1.6x slower), then both MMA batches at
        // once, and the tensor pipe idles through every softmax; alternating, one warpgroup's softmax runs beside the other's MMAs.
        // Named barriers (bar.sync / bar.arrive, 128 + 128 threads): a waiting warpgroup is descheduled by the hardware; polling an
        // mbarrier with 128 threads for a whole exp phase slowed the working warpgroup down 2x. Taken at the top of the tile, where
        // nothing but the running state is live (between the max and exp phases it cost 400 B of spills in the hot loop).
        if (p.pingpong && (w == 1 || j > 0)) asm volatile("bar.sync %0, 256;" ::"r"(1 + w) : "memory");
        if (p.spin) mbar_spin(&s_ready[w], j & 1, 30); else mbar_wait(&s_ready[w], j & 1, 30);
        long long t_wake = dbg ? clock64() : 0;
        tc_fence_after();
        uint32_t v[4][32];
        tmem_ld_32x32(t_s, v[0]);
        tmem_ld_32x32(t_s + 32, v[1]);
        tmem_ld_32x32(t_s + 64, v[2]);
        tmem_ld_32x32(t_s + 96, v[3]);
        tmem_ld_wait();
        const long long t_ld = dbg ? clock64() : 0;
        if (nvalid < BN) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i >= nvalid) v[c][i] = 0xff800000u;  // -inf: ignored by the max, exp2 -> 0
        }
        if (kMask && key_hi > 0 && (t * BN < key_lo || t * BN + BN > key_hi)) {
          // additive padding mask of the reference (createCausalMask: -1e9 on padded keys, added to the scaled score in fp32).
          // ulp(1e9) = 64, so the score is absorbed: every padded key ends up with the SAME value, a row that sees real keys
          // gives them weight exp(-1e9) = 0, and a row that sees nothing but padding (left padding) attends uniformly. That is
          // reproduced with one constant for all padded keys: -2^k / scale-ish (pad_raw, a power of two so that pad_raw * sl2
          // is exact and the maximum subtraction below yields exactly 0, i.e. P = 1, in bf16 and f16 alike).
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int kj = t * BN + c * 32 + i;
              if (kj < key_lo || kj >= key_hi) v[c][i] = __float_as_uint(p.pad_raw);
            }
        }
        if (kMask && p.causal && t * BN + BN - 1 > q_blk0 + w * QT + quarter * 32) {
          const int qi = q_blk0 + w * QT + row;
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (t * BN + c * 32 + i > qi) v[c][i] = 0xff800000u;
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; i += 2) mx[c] = fmaxf(mx[c], fmaxf(__uint_as_float(v[c][i]), __uint_as_float(v[c][i + 1])));
        const float m_new = fmaxf(m_run, fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * sl2);
        // lazy rescale: keep the stale maximum while the new one is within 2^8 of it
        const bool move = (m_new - m_run) > 8.0f;   // first tile: m_run = -inf -> true
        const float m_use = move ? m_new : m_run;
        const float alpha = move ? fast_exp2(m_run - m_new) : 1.0f;
        const float neg_m = -m_use;
        const long long t_max = dbg ? clock64() : 0;
        // packed fp32x2 arithmetic (FFMA2 / FADD2): the loop is bound by the issue rate of its single warp per scheduler
        // as much as by the MUFU, so two elements per instruction wherever the ISA has it
        const float2 sl2v = make_float2(sl2, sl2), nmv = make_float2(neg_m, neg_m);
        float2 rs2[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        uint32_t pk[64];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(v[c][2 * i]), __uint_as_float(v[c][2 * i + 1])), sl2v, nmv);
            constexpr int kMod = kPoly > 0 ? kPoly : 1;
            float2 e;
            if (kPoly > 0 && i % kMod == kMod - 1) e = exp2_poly2(x);
            else e = make_float2(fast_exp2(x.x), fast_exp2(x.y));
            rs2[c] = __fadd2_rn(rs2[c], e);
            pk[c * 16 + i] = apk2(e.x, e.y, kF16 ? 1 : 0);
          }
        }
        const float rs[4] = {rs2[0].x + rs2[0].y, rs2[1].x + rs2[1].y, rs2[2].x + rs2[2].y, rs2[3].x + rs2[3].y};
        // ping-pong: the other warpgroup may start its tile now (this one's exponentials are done)
        if (p.pingpong && (w == 0 || j + 1 < n_tiles)) asm volatile("bar.arrive %0, 256;" ::"r"(2 - w) : "memory");
        const long long t_exp = dbg ? clock64() : 0;
        // P (16-bit) back into the first 64 columns of the S region
        tmem_st_32x32(t_s, *reinterpret_cast<const uint32_t(*)[32]>(&pk[0]));
        tmem_st_32x32(t_s + 32, *reinterpret_cast<const uint32_t(*)[32]>(&pk[32]));
        if (j > 0 && __any_sync(0xffffffffu, move)) {
          // rescale O (PV(j-1) has completed: s_ready(j) was committed after it)
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            uint32_t o0[32], o1[32];
            tmem_ld_32x32(t_o + h * 64, o0);
            tmem_ld_32x32(t_o + h * 64 + 32, o1);
            tmem_ld_wait();
            const float2 av = make_float2(alpha, alpha);
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float2 a0 = __fmul2_rn(make_float2(__uint_as_float(o0[i]), __uint_as_float(o0[i + 1])), av);
              const float2 a1 = __fmul2_rn(make_float2(__uint_as_float(o1[i]), __uint_as_float(o1[i + 1])), av);
              o0[i] = __float_as_uint(a0.x); o0[i + 1] = __float_as_uint(a0.y);
              o1[i] = __float_as_uint(a1.x); o1[i + 1] = __float_as_uint(a1.y);
            }
            tmem_st_32x32(t_o + h * 64, o0);
            tmem_st_32x32(t_o + h * 64 + 32, o1);
          }
        }
        tmem_st_wait();
        l_run = l_run * alpha + ((rs[0] + rs[1]) + (rs[2] + rs[3]));
        m_run = m_use;
        tc_fence_before();
        mbar_arrive(&p_ready[w]);
        if (dbg && j < 16) {
          tl_wake[j] = (int)(t_wake - t_start); tl_done[j] = (int)(clock64() - t_start);
          tl_ld[j] = (int)(t_ld - t_wake); tl_max[j] = (int)(t_max - t_ld); tl_exp[j] = (int)(t_exp - t_max);
        }
      }
    }
    if (dbg)
      for (int i = 0; i < 16 && i < j; ++i)
        printf("[attn timeline] wg %d tile %2d  S ready at %6d  P handed over at %6d  (softmax %5d = tmem ld %4d + max %4d + exp %4d + st/rescale %4d)\n",
               w, i, tl_wake[i], tl_done[i], tl_done[i] - tl_wake[i], tl_ld[i], tl_max[i], tl_exp[i],
               tl_done[i] - tl_wake[i] - tl_ld[i] - tl_max[i] - tl_exp[i]);
    // ---- finalize: O / l
    mbar_wait(&o_done[w], 0, 32);
    tc_fence_after();
    const int qrow = q_blk0 + w * QT + row;
    const bool ok = qrow < p.sq;
    const float inv_l = 1.0f / l_run;
    uint16_t* orow = p.o + ((long long)p.o_row0 + b * p.o_bs + qrow) * p.ldo + head * HD;
    if (p.o_rows_per_peer > 0 && ok) {
      // sequence parallel: this query row lives on rank `dest`; store its heads straight into that rank's buffer
      const int dest = qrow / p.o_rows_per_peer;
      orow = p.o_peer[dest] + (long long)(qrow - dest * p.o_rows_per_peer) * p.ldo + p.o_col0 + head * HD;
    }
    uint32_t o[4][32];
    tmem_ld_32x32(t_o, o[0]);
    tmem_ld_32x32(t_o + 32, o[1]);
    tmem_ld_32x32(t_o + 64, o[2]);
    tmem_ld_32x32(t_o + 96, o[3]);
    tmem_ld_wait();
    if (ok) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 u;
          u.x = apk2(__uint_as_float(o[c][8 * q4 + 0]) * inv_l, __uint_as_float(o[c][8 * q4 + 1]) * inv_l, kF16 ? 1 : 0);
          u.y = apk2(__uint_as_float(o[c][8 * q4 + 2]) * inv_l, __uint_as_float(o[c][8 * q4 + 3]) * inv_l, kF16 ? 1 : 0);
          u.z = apk2(__uint_as_float(o[c][8 * q4 + 4]) * inv_l, __uint_as_float(o[c][8 * q4 + 5]) * inv_l, kF16 ? 1 : 0);
          u.w = apk2(__uint_as_float(o[c][8 * q4 + 6]) * inv_l, __uint_as_float(o[c][8 * q4 + 7]) * inv_l, kF16 ? 1 : 0);
          *reinterpret_cast<uint4*>(orow + c * 32 + q4 * 8) = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<1>(tmem_base, 512);
}


// ------------------------------------------------------------------------------------------------ variant 4
// What bounded variant 3 (ncu: tensor pipe 57 % busy): with ONE S accumulator per query tile, S(j+1) cannot be issued before
// PV(j) has consumed P(j) (P overwrites S), so every key tile is a dependency chain softmax(j) [~1600 clk] -> PV(j) + S(j+1)
// [1024 clk] per warpgroup, and two chains in strict alternation keep the tensor pipe busy 2048 of ~3200 clk.
// Variant 4 halves the key tile to 64 keys and double-buffers S: TMEM = S 2 tiles x 2 buffers x 64 + O 2 x 128 = 512 columns.
//   issuer of tile w:   S(0) -> buf 0, S(1) -> buf 1;  then per key tile j:  [P(j) ready]  PV(j) from buf j%2,  S(j+2) -> buf j%2
//   softmax of tile w:  [S(j) ready in buf j%2]  load 64 scores, max / exp2 / pack, P(j) -> first 32 columns of buf j%2
// S(j+1) is therefore complete before softmax(j) ends: the softmax warpgroups never wait for the tensor pipe and the pipe always
// has the other tile's MMAs queued — no ping-pong barriers, the two warpgroups run concurrently. Cost: the N = 64 S-MMA is
// shared-memory bound (Q re-read per k-step: 6 KB per MMA = 48 clk instead of 32), which caps the pipe at 2048 / 2560 = 80 %.
// O may only be rescaled after PV(j-1) has completed; with S(j) no longer ordered behind PV(j-1), that is its own barrier
// (pv_done), waited on only in the rare tiles where a row maximum moved.
struct A4 {
  static constexpr int BN = 64;
  static constexpr int KS = 4, VS = 3;
  static constexpr int Q_BYTES = QT * HD * 2;
  static constexpr int KV_BYTES = BN * HD * 2;
  static constexpr int KV_PANEL = BN * 128;
  static constexpr int OFF_K = 2 * Q_BYTES;
  static constexpr int OFF_V = OFF_K + KS * KV_BYTES;
  static constexpr int OFF_BAR = OFF_V + VS * KV_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
  static constexpr int THREADS = 352;   // 8 softmax warps, TMA producer, two MMA issuers
};

template <bool kF16, int kPoly>
__global__ void __launch_bounds__(A4::THREADS, 1) attn_kernel_v4(const __grid_constant__ AttnKParams p) {
  using C = A4;
  constexpr int BN = C::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* q_full = bars;               // [2]
  uint64_t* k_full = q_full + 2;         // [KS]
  uint64_t* k_empty = k_full + C::KS;    // [KS]
  uint64_t* v_full = k_empty + C::KS;    // [VS]
  uint64_t* v_empty = v_full + C::VS;    // [VS]
  uint64_t* s_ready = v_empty + C::VS;   // [2 tiles][2 buffers]
  uint64_t* p_ready = s_ready + 4;       // [2][2]
  uint64_t* pv_done = p_ready + 4;       // [2]
  uint64_t* o_done = pv_done + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int q_blk0 = blockIdx.x * 2 * QT;
  const int kv_col = head * HD;
  int n_tiles = 0;
  for (int s = 0; s < p.nseg; ++s) n_tiles += (p.seg_len[s] + BN - 1) / BN;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    for (int s = 0; s < p.nseg; ++s) { tma_prefetch_desc(&p.tmK[s]); tma_prefetch_desc(&p.tmV[s]); }
  }
  if (warp == 9 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&pv_done[i], 1);
      mbar_init(&o_done[i], 1);
    }
    for (int i = 0; i < 4; ++i) { mbar_init(&s_ready[i], 1); mbar_init(&p_ready[i], 128); }
    // a K / V stage is free once BOTH issuers' MMAs on it have completed
    for (int i = 0; i < C::KS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 2); }
    for (int i = 0; i < C::VS; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 2); }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 8) {
    // ============================================================ TMA producer
    if (lane == 0) {
      const int qrow = p.q_row0 + (int)(b * p.q_bs) + q_blk0;
      for (int w = 0; w < 2; ++w) {
        mbar_expect_tx(&q_full[w], C::Q_BYTES);
        uint8_t* dst = smem + w * C::Q_BYTES;
        tma_load_2d(dst, &p.tmQ, &q_full[w], head * HD, qrow + w * QT);
        tma_load_2d(dst + QT * 128, &p.tmQ, &q_full[w], head * HD + 64, qrow + w * QT);
      }
      auto tile_at = [&](int j, int* seg, int* row) {
        for (int s = 0; s < p.nseg; ++s) {
          const int tiles = (p.seg_len[s] + BN - 1) / BN;
          if (j < tiles) { *seg = s; *row = p.seg_row0[s] + (int)(b * p.seg_bs[s]) + j * BN; return; }
          j -= tiles;
        }
      };
      auto load_k = [&](int j) {
        int s = 0, row = 0;
        tile_at(j, &s, &row);
        const int ks = j % C::KS;
        mbar_wait(&k_empty[ks], ((j / C::KS) & 1) ^ 1, 10);
        uint8_t* kd = smem + C::OFF_K + ks * C::KV_BYTES;
        mbar_expect_tx(&k_full[ks], C::KV_BYTES);
        tma_load_2d(kd, &p.tmK[s], &k_full[ks], kv_col, row);
        tma_load_2d(kd + C::KV_PANEL, &p.tmK[s], &k_full[ks], kv_col + 64, row);
      };
      // K runs two tiles ahead of V: S(j+2) is issued right behind PV(j)
      load_k(0);
      if (n_tiles > 1) load_k(1);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 2 < n_tiles) load_k(j + 2);
        int s = 0, row = 0;
        tile_at(j, &s, &row);
        const int vs = j % C::VS;
        mbar_wait(&v_empty[vs], ((j / C::VS) & 1) ^ 1, 11);
        uint8_t* vd = smem + C::OFF_V + vs * C::KV_BYTES;
        mbar_expect_tx(&v_full[vs], C::KV_BYTES);
        tma_load_2d(vd, &p.tmV[s], &v_full[vs], kv_col, row);
        tma_load_2d(vd + C::KV_PANEL, &p.tmV[s], &v_full[vs], kv_col + 64, row);
      }
    }
    __syncwarp();
  } else if (warp == 9 || warp == 10) {
    // ============================================================ MMA issuers: one warp per query tile (see variant 3)
    if (tmem_base != 0) __trap();
    const int w = warp - 9;
    const uint32_t idesc_s = make_idesc_f16(QT, BN, !kF16, false, false);
    const uint32_t idesc_o = make_idesc_f16(QT, HD, !kF16, false, true);
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t desc_kmajor = make_smem_desc(0, 16, 1024, SWZ_128B);
    const uint64_t desc_v = make_smem_desc(0, C::KV_PANEL, 1024, SWZ_128B);
    auto issue_s = [&](int ks, int buf) {
      const uint32_t qa = (smem_base + w * C::Q_BYTES) >> 4;
      const uint32_t ka = (smem_base + C::OFF_K + ks * C::KV_BYTES) >> 4;
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) {
        const uint32_t off_q = ((k / 4) * (QT * 128) + (k % 4) * 32) >> 4;
        const uint32_t off_k = ((k / 4) * C::KV_PANEL + (k % 4) * 32) >> 4;
        umma_f16_ss<1>(w * 128 + buf * 64, desc_kmajor + (qa + off_q), desc_kmajor + (ka + off_k), idesc_s, k ? 1u : 0u);
      }
      umma_commit(&s_ready[w * 2 + buf]);
      umma_commit(&k_empty[ks]);
    };
    auto issue_pv = [&](int vs, int buf, bool accumulate) {
      const uint32_t va = (smem_base + C::OFF_V + vs * C::KV_BYTES) >> 4;
#pragma unroll
      for (int k = 0; k < BN / 16; ++k)
        umma_f16_ts(256 + w * 128, w * 128 + buf * 64 + k * 8, desc_v + (va + k * (2048 >> 4)), idesc_o, (accumulate || k) ? 1u : 0u);
      umma_commit(&v_empty[vs]);
      umma_commit(&pv_done[w]);
    };
    mbar_wait(&q_full[w], 0, 20);
    mbar_wait(&k_full[0], 0, 21);
    tc_fence_after();
    if (elect_one()) issue_s(0, 0);
    __syncwarp();
    if (n_tiles > 1) {
      mbar_wait(&k_full[1], 0, 22);
      tc_fence_after();
      if (elect_one()) issue_s(1, 1);
      __syncwarp();
    }
    for (int j = 0; j < n_tiles; ++j) {
      const int vs = j % C::VS, buf = j & 1;
      const bool more = j + 2 < n_tiles;
      mbar_wait(&v_full[vs], (j / C::VS) & 1, 24);
      if (more) mbar_wait(&k_full[(j + 2) % C::KS], ((j + 2) / C::KS) & 1, 25);
      mbar_wait(&p_ready[w * 2 + buf], (j >> 1) & 1, 23);
      tc_fence_after();
      if (elect_one()) {
        issue_pv(vs, buf, j > 0);
        if (more) issue_s((j + 2) % C::KS, buf);
        if (j == n_tiles - 1) umma_commit(&o_done[w]);
      }
      __syncwarp();
    }
  } else if (warp < 8) {
    // ============================================================ softmax warpgroups (thread == query row == TMEM lane)
    const int w = warp >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t t_o = tmem_base + lane_off + 256 + w * 128;
    const float sl2 = p.scale_log2;
    float m_run = -INFINITY, l_run = 0.f;  // m_run in the scaled (log2) domain
    int j = 0;
    for (int s = 0; s < p.nseg; ++s) {
      const int tiles = (p.seg_len[s] + BN - 1) / BN;
      for (int t = 0; t < tiles; ++t, ++j) {
        const int nvalid = min(BN, p.seg_len[s] - t * BN);
        const int buf = j & 1;
        const uint32_t t_s = tmem_base + lane_off + w * 128 + buf * 64;
        mbar_wait(&s_ready[w * 2 + buf], (j >> 1) & 1, 30);
        tc_fence_after();
        uint32_t v[2][32];
        tmem_ld_32x32(t_s, v[0]);
        tmem_ld_32x32(t_s + 32, v[1]);
        tmem_ld_wait();
        if (nvalid < BN) {
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i >= nvalid) v[c][i] = 0xff800000u;  // -inf: ignored by the max, exp2 -> 0
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            mx[2 * c] = fmaxf(mx[2 * c], fmaxf(__uint_as_float(v[c][i]), __uint_as_float(v[c][i + 1])));
            mx[2 * c + 1] = fmaxf(mx[2 * c + 1], fmaxf(__uint_as_float(v[c][i + 2]), __uint_as_float(v[c][i + 3])));
          }
        const float m_new = fmaxf(m_run, fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * sl2);
        const bool move = (m_new - m_run) > 8.0f;   // lazy rescale (first tile: m_run = -inf -> true)
        const float m_use = move ? m_new : m_run;
        const float alpha = move ? fast_exp2(m_run - m_new) : 1.0f;
        const float2 sl2v = make_float2(sl2, sl2), nmv = make_float2(-m_use, -m_use);
        float2 rs2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(v[c][2 * i]), __uint_as_float(v[c][2 * i + 1])), sl2v, nmv);
            constexpr int kMod = kPoly > 0 ? kPoly : 1;
            float2 e;
            if (kPoly > 0 && i % kMod == kMod - 1) e = exp2_poly2(x);
            else e = make_float2(fast_exp2(x.x), fast_exp2(x.y));
            rs2[c] = __fadd2_rn(rs2[c], e);
            pk[c * 16 + i] = apk2(e.x, e.y, kF16 ? 1 : 0);
          }
        }
        // P (16-bit) into the first 32 columns of this S buffer
        tmem_st_32x32(t_s, pk);
        if (j > 0 && __any_sync(0xffffffffu, move)) {
          // rescale O: PV(j-1) must have completed (S(j) was issued behind PV(j-2) only)
          mbar_wait(&pv_done[w], (j - 1) & 1, 31);
          tc_fence_after();
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            uint32_t o0[32], o1[32];
            tmem_ld_32x32(t_o + h * 64, o0);
            tmem_ld_32x32(t_o + h * 64 + 32, o1);
            tmem_ld_wait();
            const float2 av = make_float2(alpha, alpha);
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float2 a0 = __fmul2_rn(make_float2(__uint_as_float(o0[i]), __uint_as_float(o0[i + 1])), av);
              const float2 a1 = __fmul2_rn(make_float2(__uint_as_float(o1[i]), __uint_as_float(o1[i + 1])), av);
              o0[i] = __float_as_uint(a0.x); o0[i + 1] = __float_as_uint(a0.y);
              o1[i] = __float_as_uint(a1.x); o1[i + 1] = __float_as_uint(a1.y);
            }
            tmem_st_32x32(t_o + h * 64, o0);
            tmem_st_32x32(t_o + h * 64 + 32, o1);
          }
        }
        tmem_st_wait();
        l_run = l_run * alpha + ((rs2[0].x + rs2[0].y) + (rs2[1].x + rs2[1].y));
        m_run = m_use;
        tc_fence_before();
        mbar_arrive(&p_ready[w * 2 + buf]);
      }
    }
    // ---- finalize: O / l
    mbar_wait(&o_done[w], 0, 32);
    tc_fence_after();
    const int qrow = q_blk0 + w * QT + row;
    const bool ok = qrow < p.sq;
    const float inv_l = 1.0f / l_run;
    uint16_t* orow = p.o + ((long long)p.o_row0 + b * p.o_bs + qrow) * p.ldo + head * HD;
    if (p.o_rows_per_peer > 0 && ok) {
      const int dest = qrow / p.o_rows_per_peer;
      orow = p.o_peer[dest] + (long long)(qrow - dest * p.o_rows_per_peer) * p.ldo + p.o_col0 + head * HD;
    }
    uint32_t o[4][32];
    tmem_ld_32x32(t_o, o[0]);
    tmem_ld_32x32(t_o + 32, o[1]);
    tmem_ld_32x32(t_o + 64, o[2]);
    tmem_ld_32x32(t_o + 96, o[3]);
    tmem_ld_wait();
    if (ok) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 u;
          u.x = apk2(__uint_as_float(o[c][8 * q4 + 0]) * inv_l, __uint_as_float(o[c][8 * q4 + 1]) * inv_l, kF16 ? 1 : 0);
          u.y = apk2(__uint_as_float(o[c][8 * q4 + 2]) * inv_l, __uint_as_float(o[c][8 * q4 + 3]) * inv_l, kF16 ? 1 : 0);
          u.z = apk2(__uint_as_float(o[c][8 * q4 + 4]) * inv_l, __uint_as_float(o[c][8 * q4 + 5]) * inv_l, kF16 ? 1 : 0);
          u.w = apk2(__uint_as_float(o[c][8 * q4 + 6]) * inv_l, __uint_as_float(o[c][8 * q4 + 7]) * inv_l, kF16 ? 1 : 0);
          *reinterpret_cast<uint4*>(orow + c * 32 + q4 * 8) = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<1>(tmem_base, 512);
}

static cudaError_t launch_attn_v4(const AttnProblem& a, cudaStream_t stream) {
  AttnKParams p{};
  if (!fill_params(a, A4::BN, p)) return cudaErrorInvalidValue;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaSuccess;
#define F2B_ATTR4(F16_, POLY_) if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_kernel_v4<F16_, POLY_>, cudaFuncAttributeMaxDynamicSharedMemorySize, A4::SMEM_BYTES)
    F2B_ATTR4(false, 0); F2B_ATTR4(false, 2); F2B_ATTR4(false, 3); F2B_ATTR4(false, 4);
    F2B_ATTR4(true, 0); F2B_ATTR4(true, 2); F2B_ATTR4(true, 3); F2B_ATTR4(true, 4);
#undef F2B_ATTR4
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((a.sq + 2 * QT - 1) / (2 * QT), a.num_heads, a.batch);
  const int poly = a.poly < 0 ? 0 : (a.poly == 0 ? 3 : a.poly);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = dim3(A4::THREADS); cfg.dynamicSmemBytes = A4::SMEM_BYTES; cfg.stream = stream;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs; cfg.numAttrs = pdl_enabled() ? 1 : 0;
#define F2B_GO4(F16_, POLY_) cudaLaunchKernelEx(&cfg, attn_kernel_v4<F16_, POLY_>, p)
  if (a.f16) { if (poly == 2) F2B_GO4(true, 2); else if (poly == 3) F2B_GO4(true, 3); else if (poly == 4) F2B_GO4(true, 4); else F2B_GO4(true, 0); }
  else { if (poly == 2) F2B_GO4(false, 2); else if (poly == 3) F2B_GO4(false, 3); else if (poly == 4) F2B_GO4(false, 4); else F2B_GO4(false, 0); }
#undef F2B_GO4
  return cudaGetLastError();
}

static bool fill_params(const AttnProblem& a, int BN, AttnKParams& p) {
  {
    uint64_t d[2] = {(uint64_t)a.ldq, (uint64_t)a.q_rows_total};
    uint64_t s[1] = {(uint64_t)a.ldq * 2};
    uint32_t bx[2] = {64, QT};
    if (!make_tmap_bf16(&p.tmQ, a.q, 2, d, s, bx)) return false;
  }
  for (int i = 0; i < a.num_segments; ++i) {
    const KVSegment& g = a.seg[i];
    uint64_t dk[2] = {(uint64_t)g.ldk, (uint64_t)g.rows_total};
    uint64_t sk[1] = {(uint64_t)g.ldk * 2};
    uint64_t dv[2] = {(uint64_t)g.ldv, (uint64_t)g.rows_total};
    uint64_t sv[1] = {(uint64_t)g.ldv * 2};
    uint32_t bx[2] = {64, (uint32_t)BN};
    if (!make_tmap_bf16(&p.tmK[i], g.k, 2, dk, sk, bx)) return false;
    if (!make_tmap_bf16(&p.tmV[i], g.v, 2, dv, sv, bx)) return false;
    p.seg_row0[i] = g.row0;
    p.seg_len[i] = g.len;
    p.seg_bs[i] = g.batch_stride;
  }
  p.nseg = a.num_segments;
  p.causal = a.causal;
  p.kv_group = a.kv_group > 0 ? a.kv_group : 1;
  p.key_lo = a.key_lo; p.key_hi = a.key_hi; p.mask_dev = a.mask_dev;
  p.pad_raw = -exp2f(ceilf(log2f(fabsf(a.pad_bias) / a.scale)));
  p.sq = a.sq;
  p.num_heads = a.num_heads;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.q_row0 = a.q_row0;
  p.q_bs = a.q_batch_stride;
  p.o = reinterpret_cast<uint16_t*>(a.o);
  p.f16 = a.f16;
  p.ldo = a.ldo;
  p.o_row0 = a.o_row0;
  p.o_bs = a.o_batch_stride;
  p.o_rows_per_peer = a.o_rows_per_peer;
  p.o_col0 = a.o_col0;
  for (int i = 0; i < 8; ++i) p.o_peer[i] = reinterpret_cast<uint16_t*>(a.o_peer[i]);
  static const int timeline = getenv("FLUX2B_ATTN_TIMELINE") ? atoi(getenv("FLUX2B_ATTN_TIMELINE")) : 0;
  p.dbg = timeline;
  static const int spin = getenv("FLUX2B_ATTN_SPIN") ? atoi(getenv("FLUX2B_ATTN_SPIN")) : 0;
  p.spin = spin;
  static const int pingpong = getenv("FLUX2B_ATTN_PINGPONG") ? atoi(getenv("FLUX2B_ATTN_PINGPONG")) : 1;
  p.pingpong = pingpong;
  return true;
}

#ifndef F2B_ATTN_POLY_DEFAULT
#define F2B_ATTN_POLY_DEFAULT 3   // share of the exponentials on the FMA pipe when AttnProblem::poly == 0 (1 element in n; 0 = none)
#endif
static cudaError_t launch_attn_v3(const AttnProblem& a, cudaStream_t stream) {
  AttnKParams p{};
  if (!fill_params(a, A3::BN, p)) return cudaErrorInvalidValue;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaSuccess;
#define F2B_ATTR(F16_, POLY_) if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_kernel_v3<F16_, POLY_>, cudaFuncAttributeMaxDynamicSharedMemorySize, A3::SMEM_BYTES)
    F2B_ATTR(false, 0); F2B_ATTR(false, 2); F2B_ATTR(false, 3); F2B_ATTR(false, 4);
    F2B_ATTR(true, 0); F2B_ATTR(true, 2); F2B_ATTR(true, 3); F2B_ATTR(true, 4);
#undef F2B_ATTR
#define F2B_ATTR_M(F16_, POLY_) if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_kernel_v3<F16_, POLY_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, A3::SMEM_BYTES)
    F2B_ATTR_M(false, 0); F2B_ATTR_M(false, 4); F2B_ATTR_M(true, 0); F2B_ATTR_M(true, 4);
#undef F2B_ATTR_M
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((a.sq + 2 * QT - 1) / (2 * QT), a.num_heads, a.batch);
  const int poly = a.poly < 0 ? 0 : (a.poly == 0 ? F2B_ATTN_POLY_DEFAULT : a.poly);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = dim3(A3::THREADS); cfg.dynamicSmemBytes = A3::SMEM_BYTES; cfg.stream = stream;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  if (a.causal || a.key_hi > 0 || a.kv_group > 1 || a.mask_dev) {
    // text-encoder mode: two exponential variants are enough (all on the MUFU, or the default one-in-four polynomial)
#define F2B_GO_M(F16_, POLY_) cudaLaunchKernelEx(&cfg, attn_kernel_v3<F16_, POLY_, true>, p)
    if (a.f16) { if (poly == 0) F2B_GO_M(true, 0); else F2B_GO_M(true, 4); }
    else { if (poly == 0) F2B_GO_M(false, 0); else F2B_GO_M(false, 4); }
#undef F2B_GO_M
    return cudaGetLastError();
  }
#define F2B_GO(F16_, POLY_) cudaLaunchKernelEx(&cfg, attn_kernel_v3<F16_, POLY_>, p)
  if (a.f16) { if (poly == 2) F2B_GO(true, 2); else if (poly == 3) F2B_GO(true, 3); else if (poly == 4) F2B_GO(true, 4); else F2B_GO(true, 0); }
  else { if (poly == 2) F2B_GO(false, 2); else if (poly == 3) F2B_GO(false, 3); else if (poly == 4) F2B_GO(false, 4); else F2B_GO(false, 0); }
#undef F2B_GO
  return cudaGetLastError();
}

template <int BN, bool kPTmem>
static cudaError_t launch_attn(const AttnProblem& a, cudaStream_t stream) {
  using C = ACfg<BN, kPTmem>;
  AttnKParams p{};
  if (!fill_params(a, BN, p)) return cudaErrorInvalidValue;

  auto kern = attn_kernel<BN, kPTmem>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((a.sq + 2 * QT - 1) / (2 * QT), a.num_heads, a.batch);
  kern<<<grid, 320, C::SMEM_BYTES, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t attention_launch(const AttnProblem& a, cudaStream_t stream) {
  if (!gemm_init()) return cudaErrorInitializationError;
  if (a.sq <= 0 || a.num_heads <= 0 || a.batch <= 0) return cudaSuccess;
  int total = 0;
  for (int i = 0; i < a.num_segments; ++i) total += a.seg[i].len;
  if (total <= 0 || a.num_segments < 1 || a.num_segments > 3) return cudaErrorInvalidValue;
  for (int i = 0; i < a.num_segments; ++i)
    if (a.seg[i].len <= 0) return cudaErrorInvalidValue;
  // default: variant 4 (64-key tiles, double-buffered S) for the DiT's unmasked joint attention, variant 3 for the text encoder's
  // causal / padded / grouped-query mode. FLUX2B_ATTN_VARIANT overrides the default (3 = the 128-key ping-pong kernel).
  static const int env_variant = getenv("FLUX2B_ATTN_VARIANT") ? atoi(getenv("FLUX2B_ATTN_VARIANT")) : 0;
  const bool masked = a.causal || a.key_hi > 0 || a.kv_group > 1 || a.mask_dev;
  const int variant = a.variant ? a.variant : masked ? 3 : (env_variant ? env_variant : 4);
  if (a.o_rows_per_peer > 0 && variant != 3 && variant != 4) return cudaErrorInvalidValue;
  if (variant == 4 && !masked) return launch_attn_v4(a, stream);
  if ((a.causal || a.key_hi > 0 || a.kv_group > 1 || a.mask_dev) && (variant != 3 || a.num_segments != 1 || a.seg[0].len != a.sq)) return cudaErrorInvalidValue;
  if (variant == 3) return launch_attn_v3(a, stream);
  if (variant == 2) return launch_attn<128, true>(a, stream);
  return launch_attn<64, false>(a, stream);
}

}  // namespace f2b
