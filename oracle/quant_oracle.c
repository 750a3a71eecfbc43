/* quant_oracle.c — CPU oracle (TEST INFRASTRUCTURE ONLY) for the five weight-quantization modes of the reference.
 *
 * The arithmetic lives in the third-party dependency ml-explore/mlx-swift pinned exact 0.31.6 (MLX core 0.31.1),
 * which is absent from /root/reference and from this container, so this file restates MLX's published algorithm;
 * PARITY STATUS: unpinned (the reference's tests pin only (bits, group, mode), biases != nil iff affine, and shapes:
 * Tests/Flux2CoreTests/Flux2CoreTests.swift:64-85,1100-1142 — those are checked in tests/test_oracle_pins.py).
 * Call sites in the reference: quantize(model:groupSize:bits:mode:) Pipeline/Flux2Pipeline.swift:567-578;
 * quantized()/dequantized() Loading/WeightLoader.swift:795-815; level table Configuration/QuantizationConfig.swift:51-60.
 *
 * Rules (each isolated in one function so it can be corrected the moment real MLX output is available):
 *   layout      groups along the input dim of W[out,in]; element j of a uint32 word at bits [j*bits,(j+1)*bits)
 *   affine      MLX affine_quantize: scale=max((max-min)/n_bins,1e-7); side=|min|>|max|; scale=side?scale:-scale;
 *               edge=side?min:max; q0=round(edge/scale); scale=q0?edge/q0:scale; bias=q0?edge:0;
 *               q=min(round((w-bias)/scale),n_bins) using the UNROUNDED fp32 scale/bias; scales/biases stored f16
 *   mx scale    E8M0 = clamp(round(log2(amax/fmax)),-127,127)+127, fmax = 448 (fp8) | 6 (fp4)
 *   nv scale    E4M3(amax/6), RNE, saturating at 448
 *   elements    E4M3 / E2M1 of w/scale, RNE, saturating
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC quant_oracle.c -o _build/libquant_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct { int bits, group, mode; } qspec; /* mode 0 affine, 1 mx, 2 nv */
static qspec spec(int quant) {
  qspec q = {16, 0, -1};
  switch (quant) {
    case 1: q.bits = 8; q.group = 64; q.mode = 0; break; /* qint8 */
    case 2: q.bits = 4; q.group = 64; q.mode = 0; break; /* int4  */
    case 3: q.bits = 8; q.group = 32; q.mode = 1; break; /* mxfp8 */
    case 4: q.bits = 4; q.group = 32; q.mode = 1; break; /* mxfp4 */
    case 5: q.bits = 4; q.group = 16; q.mode = 2; break; /* nvfp4 */
  }
  return q;
}
int oracle_quant_params(int quant, int* bits, int* group, int* has_biases) {
  qspec q = spec(quant);
  if (q.mode < 0) return -1;
  *bits = q.bits; *group = q.group; *has_biases = q.mode == 0;
  return 0;
}

/* ---- f16 <-> f32 (software, RNE) */
static uint16_t f32_to_f16(float f) {
  uint32_t x; memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t mant = x & 0x7fffffu;
  int exp = (int)((x >> 23) & 0xff);
  if (exp == 0xff) return (uint16_t)(sign | 0x7c00u | (mant ? 0x200u : 0));
  int e = exp - 127 + 15;
  if (e >= 0x1f) return (uint16_t)(sign | 0x7c00u);
  if (e <= 0) {
    if (e < -10) return (uint16_t)sign;
    mant |= 0x800000u;
    int shift = 14 - e;
    uint32_t half = mant >> shift;
    uint32_t rem = mant & ((1u << shift) - 1u);
    uint32_t mid = 1u << (shift - 1);
    if (rem > mid || (rem == mid && (half & 1))) half++;
    return (uint16_t)(sign | half);
  }
  uint32_t half = (uint32_t)(e << 10) | (mant >> 13);
  uint32_t rem = mant & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) half++;
  return (uint16_t)(sign | half);
}
static float f16_to_f32(uint16_t h) {
  uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1f, mant = h & 0x3ffu, x;
  if (exp == 0) {
    if (mant == 0) x = sign;
    else {
      int e = -1;
      do { e++; mant <<= 1; } while (!(mant & 0x400u));
      x = sign | ((uint32_t)(127 - 15 - e) << 23) | ((mant & 0x3ffu) << 13);
    }
  } else if (exp == 0x1f) x = sign | 0x7f800000u | (mant << 13);
  else x = sign | ((exp - 15 + 127) << 23) | (mant << 13);
  float f; memcpy(&f, &x, 4); return f;
}
static float bf16_to_f32(uint16_t h) { uint32_t x = (uint32_t)h << 16; float f; memcpy(&f, &x, 4); return f; }
static float load_w(const void* w, int dtype, int64_t i) { /* 0 f32, 1 f16, 2 bf16 */
  if (dtype == 0) return ((const float*)w)[i];
  if (dtype == 1) return f16_to_f32(((const uint16_t*)w)[i]);
  return bf16_to_f32(((const uint16_t*)w)[i]);
}

/* ---- E8M0: round(log2(x)) restated exactly: floor(log2 x) + (mantissa >= sqrt(2)); sqrt(2) is irrational so no tie */
uint8_t oracle_to_e8m0(float x) {
  if (!(x > 0.0f)) return 0;
  if (isinf(x)) return 0xFF;
  uint32_t u; memcpy(&u, &x, 4);
  int e = (int)((u >> 23) & 0xff);
  uint32_t m = u & 0x7fffffu;
  int n = (e == 0) ? -127 : (e - 127) + (m >= 0x3504F4u ? 1 : 0);
  if (n < -127) n = -127;
  if (n > 127) n = 127;
  return (uint8_t)(n + 127);
}
float oracle_from_e8m0(uint8_t b) {
  if (b == 0) return ldexpf(1.0f, -127);
  if (b == 255) return INFINITY;
  return ldexpf(1.0f, (int)b - 127);
}
/* ---- E4M3 (fn): RNE, saturating at 448 */
uint8_t oracle_to_e4m3(float x) {
  uint32_t u; memcpy(&u, &x, 4);
  uint8_t sign = (u >> 31) ? 0x80 : 0;
  float a = fabsf(x);
  if (a != a) return sign | 0x7F;
  if (a >= 448.0f) return sign | 0x7E;
  if (a < 0.015625f) return sign | (uint8_t)rintf(a * 512.0f);
  uint32_t au; memcpy(&au, &a, 4);
  int e = (int)(au >> 23) - 127;
  uint32_t m = au & 0x7fffffu, keep = m >> 20, rem = m & 0xfffffu;
  if (rem > 0x80000u || (rem == 0x80000u && (keep & 1))) keep++;
  if (keep == 8) { keep = 0; e++; }
  uint32_t code = ((uint32_t)(e + 7) << 3) | keep;
  if (code > 0x7E) code = 0x7E;
  return sign | (uint8_t)code;
}
float oracle_from_e4m3(uint8_t b) {
  float s = (b & 0x80) ? -1.0f : 1.0f;
  int e = (b >> 3) & 0xF, m = b & 7;
  if (e == 0) return s * (float)m * 0.001953125f;
  if (e == 15 && m == 7) return NAN;
  return s * ldexpf(1.0f + (float)m * 0.125f, e - 7);
}
/* ---- E2M1: grid {0,.5,1,1.5,2,3,4,6}, RNE (ties to the even mantissa), saturating */
uint8_t oracle_to_e2m1(float x) {
  uint32_t u; memcpy(&u, &x, 4);
  uint8_t sign = (u >> 31) ? 0x8 : 0x0, b;
  float a = fabsf(x);
  if (a != a) b = 0x7;
  else if (a > 5.0f) b = 0x7;
  else if (a >= 3.5f) b = 0x6;
  else if (a > 2.5f) b = 0x5;
  else if (a >= 1.75f) b = 0x4;
  else if (a > 1.25f) b = 0x3;
  else if (a >= 0.75f) b = 0x2;
  else if (a > 0.25f) b = 0x1;
  else b = 0x0;
  return b | sign;
}
float oracle_from_e2m1(uint8_t b) {
  static const float tab[8] = {0.f, 0.5f, 1.f, 1.5f, 2.f, 3.f, 4.f, 6.f};
  float v = tab[b & 7];
  return (b & 8) ? -v : v;
}

/* w [rows, cols] (dtype 0 f32 | 1 f16 | 2 bf16) -> packed uint32, scales (f16 bits or uint8), biases (f16 bits) */
int oracle_quantize(int quant, const void* w, int dtype, int64_t rows, int64_t cols, uint32_t* packed, void* scales,
                    uint16_t* biases) {
  qspec q = spec(quant);
  if (q.mode < 0 || cols % q.group) return -1;
  const int per_word = 32 / q.bits;
  const int64_t gpr = cols / q.group;
  for (int64_t r = 0; r < rows; ++r)
    for (int64_t g = 0; g < gpr; ++g) {
      const int64_t base = r * cols + g * q.group, gid = r * gpr + g;
      uint32_t* out = packed + base / per_word;
      if (q.mode == 0) {
        float wmax = -INFINITY, wmin = INFINITY;
        for (int i = 0; i < q.group; ++i) { float v = load_w(w, dtype, base + i); wmax = fmaxf(wmax, v); wmin = fminf(wmin, v); }
        const float n_bins = (float)((1 << q.bits) - 1);
        float scale = fmaxf((wmax - wmin) / n_bins, 1e-7f);
        const int side = fabsf(wmin) > fabsf(wmax);
        scale = side ? scale : -scale;
        const float edge = side ? wmin : wmax;
        const float q0 = roundf(edge / scale);
        const int at_zero = (q0 == 0.0f);
        scale = at_zero ? scale : edge / q0;
        const float bias = at_zero ? 0.0f : edge;
        ((uint16_t*)scales)[gid] = f32_to_f16(scale);
        biases[gid] = f32_to_f16(bias);
        for (int wd = 0; wd < q.group / per_word; ++wd) {
          uint32_t word = 0;
          for (int j = 0; j < per_word; ++j) {
            float v = load_w(w, dtype, base + wd * per_word + j);
            float rr = roundf((v - bias) / scale);
            rr = fminf(rr, n_bins);
            rr = fmaxf(rr, 0.0f);
            word |= ((uint32_t)rr) << (j * q.bits);
          }
          out[wd] = word;
        }
      } else {
        float amax = 0.0f;
        for (int i = 0; i < q.group; ++i) amax = fmaxf(amax, fabsf(load_w(w, dtype, base + i)));
        float scale = amax / (q.bits == 4 ? 6.0f : 448.0f);
        uint8_t sb;
        if (q.mode == 1) { sb = oracle_to_e8m0(scale); scale = oracle_from_e8m0(sb); }
        else { sb = oracle_to_e4m3(scale); scale = oracle_from_e4m3(sb); }
        ((uint8_t*)scales)[gid] = sb;
        for (int wd = 0; wd < q.group / per_word; ++wd) {
          uint32_t word = 0;
          for (int j = 0; j < per_word; ++j) {
            float v = load_w(w, dtype, base + wd * per_word + j);
            float x = (scale == 0.0f) ? 0.0f : v / scale;
            uint32_t code = (q.bits == 4) ? oracle_to_e2m1(x) : oracle_to_e4m3(x);
            word |= code << (j * q.bits);
          }
          out[wd] = word;
        }
      }
    }
  return 0;
}

/* -> fp32 [rows, cols]: q*scale + bias (two roundings, no FMA) | element * scale */
int oracle_dequantize(int quant, const uint32_t* packed, const void* scales, const uint16_t* biases, int64_t rows,
                      int64_t cols, float* out) {
  qspec q = spec(quant);
  if (q.mode < 0 || cols % q.group) return -1;
  const int per_word = 32 / q.bits;
  const uint32_t mask = (1u << q.bits) - 1u;
  const int64_t gpr = cols / q.group;
  for (int64_t r = 0; r < rows; ++r)
    for (int64_t c = 0; c < cols; ++c) {
      const int64_t e = r * cols + c, gid = r * gpr + c / q.group;
      const uint32_t code = (packed[e / per_word] >> ((e % per_word) * q.bits)) & mask;
      if (q.mode == 0) {
        const float s = f16_to_f32(((const uint16_t*)scales)[gid]), b = f16_to_f32(biases[gid]);
        volatile float prod = (float)code * s;
        out[e] = prod + b;
      } else {
        const uint8_t sb = ((const uint8_t*)scales)[gid];
        const float s = (q.mode == 1) ? oracle_from_e8m0(sb) : oracle_from_e4m3(sb);
        const float ev = (q.bits == 4) ? oracle_from_e2m1((uint8_t)code) : oracle_from_e4m3((uint8_t)code);
        out[e] = ev * s;
      }
    }
  return 0;
}
uint16_t oracle_f32_to_f16(float f) { return f32_to_f16(f); }
float oracle_f16_to_f32(uint16_t h) { return f16_to_f32(h); }
