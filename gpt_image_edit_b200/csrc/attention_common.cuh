// Shared definitions of the attention kernels (attention.cu: production kernels and host dispatch;
// attention_experiments.cu: alternative structures kept for the record, selected by B2F_ATTN_VARIANT).
#pragma once
#include "host_common.h"
#include "ptx.cuh"

namespace b2f {
namespace attn {

constexpr int DH = 128;
constexpr int BQ = 128;   // rows per query tile
constexpr int BKV = 128;  // rows per K/V block
constexpr int KV_SLOTS = 4;
constexpr int TILE_BYTES = 128 * DH * 2;  // 32 KB
constexpr int ATTN_THREADS = 320;
constexpr int ATTN_SMEM = (2 + KV_SLOTS) * TILE_BYTES + 256 + 1024;

struct AttnParams {
  int B, H, Hkv, Sq, Skv;
  float scale_log2;
  int causal;
  __nv_bfloat16* out;
  long long ldo;
  // additive score bias (BIAS kernels only): score = bias_scale * q.k + bias[h, q, kv]
  const __nv_bfloat16* bias;
  long long bias_h_stride, bias_row_stride;
  float bias_scale;
  // optional log-sum-exp output for the backward pass, base-2 domain: lse2[b, h, q] = max2 + log2(sum), where the scores
  // are scale_log2 * q.k; element (b, h, q) at lse + (b*H + h)*lse_stride + q
  float* lse;
  long long lse_stride;
};

static __device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA/ALU pipes (no MUFU): round-to-nearest split x = n + f, f in [-0.5, 0.5], degree-3
// minimax polynomial for 2^f (max rel. error 1.0e-4, far below the bf16 rounding of P), exponent
// add through the integer pipe.  Used for a fraction of the exponentials so the SFU (16 ex2/clk/SM)
// stops being co-critical with the tensor pipe.
static __device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;            // 1.5 * 2^23: low mantissa bits of t hold n
  const float f = x - (t - 12582912.0f);
  float r = fmaf(0.05592203512787819f, f, 0.24264007806777954f);
  r = fmaf(r, f, 0.6931210160255432f);
  r = fmaf(r, f, 0.9999244809150696f);
  return __int_as_float(__float_as_int(r) + (__float_as_int(t) << 23));
}

// Packed pair version (FFMA2 / FADD2): 2^x0, 2^x1 without MUFU in 11 issue slots.
static __device__ __forceinline__ void ex2_poly2(float x0, float x1, float& r0, float& r1) {
  x0 = fmaxf(x0, -125.0f);
  x1 = fmaxf(x1, -125.0f);
  float t0, t1, n0, n1, f0, f1;
  fadd2(t0, t1, x0, x1, 12582912.0f, 12582912.0f);
  fadd2(n0, n1, t0, t1, -12582912.0f, -12582912.0f);
  fadd2(f0, f1, x0, x1, -n0, -n1);
  ffma2(r0, r1, f0, f1, 0.05592203512787819f, 0.05592203512787819f, 0.24264007806777954f, 0.24264007806777954f);
  ffma2(r0, r1, r0, r1, f0, f1, 0.6931210160255432f, 0.6931210160255432f);
  ffma2(r0, r1, r0, r1, f0, f1, 0.9999244809150696f, 0.9999244809150696f);
  r0 = __int_as_float(__float_as_int(r0) + (__float_as_int(t0) << 23));
  r1 = __int_as_float(__float_as_int(r1) + (__float_as_int(t1) << 23));
}

typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnParams);

// launch shape of a kernel variant
struct Variant {
  KernelFn fn;
  int threads, smem;
  bool single_tile;   // one 128-row Q tile per CTA (grid.x = ceil(Sq / 128)) instead of two
};

// attention_experiments.cu: fills `out` for variants 10-12, 30-32, 40-42, 60-62; false for anything else
bool experimental_variant(int variant, Variant* out);

}  // namespace attn
}  // namespace b2f
