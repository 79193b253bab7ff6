// Fused non-causal / causal softmax attention for sm_100a, head_dim 128, bf16 in/out.
//
//   O[b, q, h, :] = softmax(Q[b, q, h, :] · K[b, :, hk, :]^T * scale) · V[b, :, hk, :]
//
// One CTA owns TWO 128-row query tiles of one (batch, head) and streams K/V in 128-row blocks:
//   warp 0 (1 lane)  TMA producer: Q tiles once, K/V blocks through a 4-slot 32 KB ring
//   warp 1 (1 lane)  tcgen05.mma issuer:  S_t = Q_t·K_j^T (SS),  O_t += P_t·V_j (A = P from TMEM)
//   warps 2..5       softmax warpgroup of tile 0   } one thread per query row: tcgen05.ld S,
//   warps 6..9       softmax warpgroup of tile 1   } online max/sum, exp2, P -> TMEM (bf16) over S
// TMEM (512 cols): S0|P0 [0,128)  S1|P1 [128,256)  O0 [256,384)  O1 [384,512).
// While one warpgroup runs softmax on its tile, the tensor core works on the other tile.
// O is rescaled lazily (only when the running max grows by more than 2^8), by the softmax
// warpgroup itself, between "S ready" (which also proves the previous P·V finished) and "P ready".
//
// Kernels in this file (B2F_ATTN_VARIANT selects at run time; all parity-green, tests/test_attention_gpu.py):
//   51 (default)  attn_fwd_kernel_2cta<4>: the same two-tile structure as a CTA PAIR (cta_group::2, M = 256 MMAs,
//                 each CTA stores half of every K / V tile) for non-causal, bias-free calls with >= 512 query rows
//                 (the FLUX joint attention); anything else falls through to
//   1             attn_fwd_kernel<4>: the single-CTA two-tile kernel described above (causal, GQA, bias, short)
//   0,2,5,6       other fractions of exponentials on the FMA pipe;  50/52 the same for the pair kernel
//   10-12, 30-32, 40-42, 60-62   experiments kept for the record (S sub-blocks / single tile with S double
//                 buffering / column-split warpgroups): none is faster, see DESIGN.md section 7
//
// Replaces F.scaled_dot_product_attention as reached by diffusers FluxAttnProcessor2_0
// (reference call site univa/utils/flux_pipeline.py:1067) and flash_attn as reached
// through transformers' attn_implementation="flash_attention_2" (univa/serve/cli.py:40).
#include <atomic>
#include <cmath>
#include <cstdlib>

#ifndef B2F_ATTN_DEFAULT_VARIANT
#define B2F_ATTN_DEFAULT_VARIANT 51
#endif

#include "host_common.h"
#include "ptx.cuh"

namespace b2f {

extern std::atomic<uint64_t> g_launch_count;

namespace {

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
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA/ALU pipes (no MUFU): round-to-nearest split x = n + f, f in [-0.5, 0.5], degree-3
// minimax polynomial for 2^f (max rel. error 1.0e-4, far below the bf16 rounding of P), exponent
// add through the integer pipe.  Used for a fraction of the exponentials so the SFU (16 ex2/clk/SM)
// stops being co-critical with the tensor pipe.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;            // 1.5 * 2^23: low mantissa bits of t hold n
  const float f = x - (t - 12582912.0f);
  float r = fmaf(0.05592203512787819f, f, 0.24264007806777954f);
  r = fmaf(r, f, 0.6931210160255432f);
  r = fmaf(r, f, 0.9999244809150696f);
  return __int_as_float(__float_as_int(r) + (__float_as_int(t) << 23));
}

// Packed pair version (FFMA2 / FADD2): 2^x0, 2^x1 without MUFU in 11 issue slots.
__device__ __forceinline__ void ex2_poly2(float x0, float x1, float& r0, float& r1) {
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

// POLY: one pair of exponentials in every POLY pairs goes to the polynomial (0 = never).
// TURNS: the two softmax warpgroups take turns on the exp section (forces anti-phase).
// ABL: timing ablations (wrong results), see attention_fwd.  BIAS: additive bf16 score bias (T5 relative
// position bias); the bias kernel folds the score scale into the bias step and runs with scale_log2 = log2(e).
template <int POLY, bool TURNS, int ABL = 0, bool BIAS = false>
__global__ void __launch_bounds__(ATTN_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* q_smem = smem;                       // 2 tiles
  uint8_t* kv_smem = smem + 2 * TILE_BYTES;     // KV_SLOTS tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (2 + KV_SLOTS) * TILE_BYTES);
  uint64_t* q_full = bars;            // 1
  uint64_t* kv_full = bars + 1;       // KV_SLOTS
  uint64_t* kv_empty = kv_full + KV_SLOTS;
  uint64_t* s_full = kv_empty + KV_SLOTS;  // 2
  uint64_t* p_full = s_full + 2;           // [tile][half] = 4: P columns [0,64) and [64,128) handed over separately
  uint64_t* o_done = p_full + 4;           // 2
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qpair = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.H / p.Hkv);
  const int q0 = qpair * 2 * BQ;

  // K/V blocks this CTA needs (causal: only up to its last query row; Sq == Skv assumed then)
  int kv_len = p.Skv;
  if (p.causal) kv_len = min(p.Skv, q0 + 2 * BQ);
  const int n_kv = (kv_len + BKV - 1) / BKV;
  // the second Q tile of the last pair may lie entirely beyond Sq (S = 8736 = 34*256 + 32): skip all of
  // its MMAs and its softmax warpgroup instead of multiplying zero rows
  const bool t1_active = TURNS || (q0 + BQ < p.Sq);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < KV_SLOTS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&o_done[i], 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&p_full[i], 4);  // one elected arrive per softmax warp
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      mbar_expect_tx(q_full, 2 * TILE_BYTES);
      for (int t = 0; t < 2; ++t)
        for (int half = 0; half < 2; ++half)
          tma_load_3d(q_smem + t * TILE_BYTES + half * (TILE_BYTES / 2), &tmQ, q_full,
                      h * DH + half * 64, q0 + t * BQ, b);
      int slot = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        for (int kv = 0; kv < 2; ++kv) {  // K_j then V_j
          mbar_wait(&kv_empty[slot], phase ^ 1);
          mbar_expect_tx(&kv_full[slot], TILE_BYTES);
          uint8_t* dst = kv_smem + slot * TILE_BYTES;
          const CUtensorMap* tm = kv == 0 ? &tmK : &tmV;
          tma_load_3d(dst, tm, &kv_full[slot], hk * DH, j * BKV, b);
          tma_load_3d(dst + TILE_BYTES / 2, tm, &kv_full[slot], hk * DH + 64, j * BKV, b);
          if (++slot == KV_SLOTS) {
            slot = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    // The WHOLE warp runs this loop (waits, descriptor arithmetic) so that the address math stays on
    // the uniform datapath; only the tcgen05.mma / tcgen05.commit instructions are predicated to one
    // lane.  (With the loop nested under `if (lane == 0)` every descriptor went through R2UR moves and
    // the issue thread, not the tensor pipe, paced the kernel: ncu showed it busy ~75 % of the time.)
    constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, BKV, 0);  // B = K tile, K-major
    constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, DH, 1);   // B = V tile, MN-major
    const uint32_t q_addr = smem_u32(q_smem);
    const uint32_t kv_addr = smem_u32(kv_smem);
    // descriptor of byte offset 0 of each buffer; every MMA operand is "base + constant" (one uniform
    // 64-bit add on the 14-bit address field, which cannot carry out for addresses < 256 KB)
    const uint64_t dq_base = make_sdesc_sw128(q_addr, 16, 1024);
    const uint64_t dk_base = make_sdesc_sw128(kv_addr, 16, 1024);
    const uint64_t dv_base = make_sdesc_sw128(kv_addr, TILE_BYTES / 2, 1024);
    int slot = 0;
    uint32_t phase = 0;
    auto issue_qk = [&](int t, int k_slot) {
      const uint32_t d = tmem_base + uint32_t(t * 128);
      const uint64_t qd = dq_base + uint64_t((t * TILE_BYTES) >> 4);
      const uint64_t kd = dk_base + uint64_t((k_slot * TILE_BYTES) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) {
          const uint64_t off = uint64_t(((k >> 2) * (TILE_BYTES / 2) + (k & 3) * 32) >> 4);
          umma_ss(d, qd + off, kd + off, idesc_qk, k != 0 ? 1u : 0u);
        }
      }
      __syncwarp();
    };
    // O_t += P_t[:, 64*hf : 64*hf+64] · V[64*hf : 64*hf+64, :]  (4 k-steps of 16 kv rows)
    auto issue_pv = [&](int t, int v_slot, int hf, bool first) {
      const uint32_t d = tmem_base + 256 + uint32_t(t * 128);
      const uint32_t pa = tmem_base + uint32_t(t * 128 + hf * 32);
      const uint64_t vd = dv_base + uint64_t((v_slot * TILE_BYTES + hf * 8192) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ts(d, pa + k * 8, vd + uint64_t((k * 2048) >> 4), idesc_pv, (first && k == 0) ? 0u : 1u);
      }
      __syncwarp();
    };
    auto commit = [&](uint64_t* bar) {
      if (elect_one()) umma_commit(bar);
      __syncwarp();
    };
    auto advance = [&]() {
      if (++slot == KV_SLOTS) {
        slot = 0;
        phase ^= 1;
      }
    };
    mbar_wait(q_full, 0);
    // prologue: S_t = Q_t K_0^T for both tiles
    mbar_wait(&kv_full[slot], phase);
    tc_fence_after();
    issue_qk(0, slot);
    commit(&s_full[0]);
    if (t1_active) {
      issue_qk(1, slot);
      commit(&s_full[1]);
    }
    commit(&kv_empty[slot]);
    advance();
    for (int j = 0; j < n_kv; ++j) {
      const int v_slot = slot;
      const uint32_t v_phase = phase;
      advance();
      const int k_slot = slot;  // K_{j+1} (if any)
      const uint32_t k_phase = phase;
      const bool more = (j + 1 < n_kv);
      if (more) advance();
      mbar_wait(&kv_full[v_slot], v_phase);
      // tile 0: the first half of P·V starts while the warpgroup still exponentiates the second half
      mbar_wait(&p_full[0], j & 1);
      tc_fence_after();
      issue_pv(0, v_slot, 0, j == 0);
      mbar_wait(&p_full[1], j & 1);
      tc_fence_after();
      issue_pv(0, v_slot, 1, false);
      if (more) {
        mbar_wait(&kv_full[k_slot], k_phase);
        tc_fence_after();
        issue_qk(0, k_slot);
        commit(&s_full[0]);
      }
      // tile 1
      if (t1_active) {
        mbar_wait(&p_full[2], j & 1);
        tc_fence_after();
        issue_pv(1, v_slot, 0, j == 0);
        mbar_wait(&p_full[3], j & 1);
        tc_fence_after();
        issue_pv(1, v_slot, 1, false);
      }
      commit(&kv_empty[v_slot]);
      if (more) {
        if (t1_active) {
          issue_qk(1, k_slot);
          commit(&s_full[1]);
        }
        commit(&kv_empty[k_slot]);
      }
    }
    commit(&o_done[0]);
    commit(&o_done[1]);
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    const int t = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    const int q_row = q0 + t * BQ + row_in_tile;  // query index inside the sequence
    const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_addr + uint32_t(t * 128);
    const uint32_t o_tmem = tmem_base + lane_addr + 256 + uint32_t(t * 128);
    float m = -INFINITY, l = 0.f;
    if (t == 0 || t1_active) {
    if (TURNS) {
    // named barriers 1/2 = "tile 0 / tile 1 may run its exp section"; tile 0 goes first
    if (t == 1) named_bar_arrive(1, 256);
    }
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      uint32_t sr[128];
      B2F_TMEM_LD_X32(s_tmem + 0, (sr + 0));
      B2F_TMEM_LD_X32(s_tmem + 32, (sr + 32));
      B2F_TMEM_LD_X32(s_tmem + 64, (sr + 64));
      B2F_TMEM_LD_X32(s_tmem + 96, (sr + 96));
      tmem_wait_ld();
      const int kv0 = j * BKV;
      if (BIAS) {
        if (q_row < p.Sq) {
          const __nv_bfloat16* brow = p.bias + (long long)h * p.bias_h_stride +
                                      (long long)q_row * p.bias_row_stride + kv0;
          const int n_ok = min(BKV, p.Skv - kv0);
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if (c < n_ok)
              sr[c] = __float_as_uint(fmaf(__uint_as_float(sr[c]), p.bias_scale, __bfloat162float(brow[c])));
        }
      }
      const bool need_mask = (kv0 + BKV > p.Skv) || (p.causal && kv0 + BKV > q0 + t * BQ);
      if (need_mask) {
        const int limit = p.causal ? min(p.Skv, q_row + 1) : p.Skv;
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if (kv0 + c >= limit) sr[c] = 0xff800000u;  // -inf
      }
      // row max with 3-input FMNMX3 in 4 independent chains (a single dependent chain costs ~4 clk per
      // link and only two softmax warps share an SM sub-partition, so nothing would hide it)
      float mx4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        mx4[i] = fmaxf(__uint_as_float(sr[2 * i]), __uint_as_float(sr[2 * i + 1]));
#pragma unroll
      for (int c = 8; c < 128; c += 8)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          mx4[i] = fmax3(mx4[i], __uint_as_float(sr[c + 2 * i]), __uint_as_float(sr[c + 2 * i + 1]));
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float m_new = fmaxf(m, mx * p.scale_log2);
      // lazy rescale: keep the stale max unless it grew by more than 2^8 (P stays < 256)
      const bool grow = (m_new - m) > 8.0f;
      const float m_use = grow ? m_new : m;
      const float alpha = grow ? ex2(m - m_use) : 1.0f;
      const float neg_m = (m_use == -INFINITY) ? 0.f : -m_use;  // fully masked row (causal tail)
      if (j > 0 && __any_sync(0xffffffffu, grow)) {
        // Rare path, BEFORE any P of this block is published: S_t(j) being ready proves P_t·V_{j-1}
        // completed and P_t·V_j cannot start before the arrives below, so O_t is quiescent here.
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t o[32];
          B2F_TMEM_LD_X32(o_tmem + c0, o);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
          B2F_TMEM_ST_X32(o_tmem + c0, o);
        }
      }
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};   // two packed (FADD2) accumulator pairs
      if (TURNS) named_bar_sync(1 + t, 256);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          float x0, x1;
          ffma2(x0, x1, __uint_as_float(sr[half * 64 + 2 * c]), __uint_as_float(sr[half * 64 + 2 * c + 1]),
                p.scale_log2, p.scale_log2, neg_m, neg_m);
          float p0, p1;
          if (POLY && (c % (POLY ? POLY : 1)) == (POLY ? POLY : 1) - 1) {
            ex2_poly2(x0, x1, p0, p1);
          } else if (ABL & 1) {
            p0 = x0 * 1e-3f;   // ablation: no MUFU
            p1 = x1 * 1e-3f;
          } else {
            p0 = ex2(x0);
            p1 = ex2(x1);
          }
          const int a = (c & 1) * 2;
          fadd2(sum4[a], sum4[a + 1], sum4[a], sum4[a + 1], p0, p1);
          pk[c] = pack_bf16x2(p0, p1);
        }
        B2F_TMEM_ST_X32(s_tmem + half * 32, pk);
        // publish this half of P (and, with the first half, the rescaled O): one arrive per warp
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t * 2 + half]);
      }
      if (TURNS) named_bar_arrive(1 + (t ^ 1), 256);
      l = l * alpha + ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
      m = m_use;
    }
    // ---------------------------------------------------------------- epilogue: O / l -> bf16
    mbar_wait(&o_done[t], 0);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const bool row_ok = q_row < p.Sq;
    __nv_bfloat16* out_row =
        p.out + ((long long)b * p.Sq + q_row) * p.ldo + (long long)h * DH;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t o[32];
      __syncwarp();
      B2F_TMEM_LD_X32(o_tmem + c0, o);
      tmem_wait_ld();
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
          v.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
          v.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
          v.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(out_row + c0 + g * 8) = v;
        }
      }
    }
    }  // t == 0 || t1_active
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ================================================================================================
// v2: the same CTA layout (two 128-row Q tiles, 4-slot K/V ring of 128-row tiles), but S is produced
// and consumed in 64-column SUB-BLOCKS with TWO S buffers per tile, so the tensor core computes
// S(u+1) (and S(u+2)) while the softmax warpgroup is still working on S(u):
//   TMEM per tile t:  S buffers at t*128 + {0, 64} (P bf16 aliased over the first 32 columns of each),
//                     O at 256 + t*128.
//   MMA lane:   QK(t,0) QK(t,1) | for u: wait P(t,u) -> PV(t,u) -> QK(t,u+2) into the buffer PV(t,u)
//               just released.  The softmax warpgroup never waits for the tensor pipe in steady state.
//   o_done[t] completes once per PV(t,u); the warpgroup only looks at it before the (rare, lazy) O
//   rescale and before the epilogue — completions can never run ahead by more than one phase because
//   PV(t,u) needs P(t,u) from the same warpgroup.
constexpr int SUB = 64;

template <int POLY>
__global__ void __launch_bounds__(ATTN_THREADS, 1)
attn_fwd_kernel_v2(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* q_smem = smem;
  uint8_t* kv_smem = smem + 2 * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (2 + KV_SLOTS) * TILE_BYTES);
  uint64_t* q_full = bars;                   // 1
  uint64_t* kv_full = bars + 1;              // KV_SLOTS
  uint64_t* kv_empty = kv_full + KV_SLOTS;   // KV_SLOTS
  uint64_t* s_full = kv_empty + KV_SLOTS;    // [tile][buf] = 4
  uint64_t* p_full = s_full + 4;             // [tile][buf] = 4
  uint64_t* o_done = p_full + 4;             // [tile] = 2
  uint64_t* o_final = o_done + 2;            // 1: everything issued has completed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_final + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qpair = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.H / p.Hkv);
  const int q0 = qpair * 2 * BQ;
  int kv_len = p.Skv;
  if (p.causal) kv_len = min(p.Skv, q0 + 2 * BQ);
  const int n_kv = (kv_len + BKV - 1) / BKV;
  const int n_sub = (kv_len + SUB - 1) / SUB;   // 64-column sub-blocks actually needed

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < KV_SLOTS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
    }
    mbar_init(&o_done[0], 1);
    mbar_init(&o_done[1], 1);
    mbar_init(o_final, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * TILE_BYTES);
      for (int t = 0; t < 2; ++t)
        for (int half = 0; half < 2; ++half)
          tma_load_3d(q_smem + t * TILE_BYTES + half * (TILE_BYTES / 2), &tmQ, q_full,
                      h * DH + half * 64, q0 + t * BQ, b);
      int slot = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        for (int kv = 0; kv < 2; ++kv) {
          mbar_wait(&kv_empty[slot], phase ^ 1);
          mbar_expect_tx(&kv_full[slot], TILE_BYTES);
          uint8_t* dst = kv_smem + slot * TILE_BYTES;
          const CUtensorMap* tm = kv == 0 ? &tmK : &tmV;
          tma_load_3d(dst, tm, &kv_full[slot], hk * DH, j * BKV, b);
          tma_load_3d(dst + TILE_BYTES / 2, tm, &kv_full[slot], hk * DH + 64, j * BKV, b);
          if (++slot == KV_SLOTS) {
            slot = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, SUB, 0);  // S sub-block: 128 x 64
    constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, DH, 1);   // O: 128 x 128, V MN-major
    const uint64_t dq_base = make_sdesc_sw128(smem_u32(q_smem), 16, 1024);
    const uint64_t dk_base = make_sdesc_sw128(smem_u32(kv_smem), 16, 1024);
    const uint64_t dv_base = make_sdesc_sw128(smem_u32(kv_smem), TILE_BYTES / 2, 1024);
    // K rows [64 hh, 64 hh + 64) of a 128-row tile start 8 KB into each dh-half sub-tile
    auto issue_qk = [&](int t, int buf, int k_slot, int hh) {
      const uint32_t d = tmem_base + uint32_t(t * 128 + buf * SUB);
      const uint64_t qd = dq_base + uint64_t((t * TILE_BYTES) >> 4);
      const uint64_t kd = dk_base + uint64_t((k_slot * TILE_BYTES + hh * (SUB * 128)) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) {
          const uint64_t off = uint64_t(((k >> 2) * (TILE_BYTES / 2) + (k & 3) * 32) >> 4);
          umma_ss(d, qd + off, kd + off, idesc_qk, k != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[t * 2 + buf]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int t, int buf, int v_slot, int hh, bool first) {
      const uint32_t d = tmem_base + 256 + uint32_t(t * 128);
      const uint32_t pa = tmem_base + uint32_t(t * 128 + buf * SUB);
      const uint64_t vd = dv_base + uint64_t((v_slot * TILE_BYTES + hh * (SUB * 128)) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < SUB / 16; ++k)
          umma_ts(d, pa + k * 8, vd + uint64_t((k * 2048) >> 4), idesc_pv, (first && k == 0) ? 0u : 1u);
        umma_commit(&o_done[t]);
      }
      __syncwarp();
    };
    auto commit = [&](uint64_t* bar) {
      if (elect_one()) umma_commit(bar);
      __syncwarp();
    };
    // ring bookkeeping: tile index i (0 = K_0, 1 = V_0, 2 = K_1, ...) lives in slot i % 4, phase (i / 4) & 1
    auto slot_of = [](int i) { return i % KV_SLOTS; };
    auto phase_of = [](int i) { return uint32_t((i / KV_SLOTS) & 1); };
    mbar_wait(q_full, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    for (int u0 = 0; u0 < 2 && u0 < n_sub; ++u0)
      for (int t = 0; t < 2; ++t) issue_qk(t, u0, 0, u0);
    if (n_sub <= 2) commit(&kv_empty[0]);   // K_0 fully consumed (otherwise released below)
    for (int u = 0; u < n_sub; ++u) {
      const int j = u >> 1, hh = u & 1, buf = u & 1;
      const int vi = 2 * j + 1;              // ring index of V_j
      if (hh == 0) {
        mbar_wait(&kv_full[slot_of(vi)], phase_of(vi));
        tc_fence_after();
      }
      const int u2 = u + 2;                  // the QK that reuses this S buffer
      const int ki2 = 2 * (u2 >> 1);         // ring index of K_{u2/2}
      for (int t = 0; t < 2; ++t) {
        mbar_wait(&p_full[t * 2 + buf], uint32_t(u >> 1) & 1);
        tc_fence_after();
        issue_pv(t, buf, slot_of(vi), hh, u == 0);
        if (u2 < n_sub) {
          if (t == 0 && hh == 0) {
            mbar_wait(&kv_full[slot_of(ki2)], phase_of(ki2));
            tc_fence_after();
          }
          issue_qk(t, buf, slot_of(ki2), hh);
        }
      }
      // releases: V_j after its second half (or the last sub-block); K tiles after their last QK
      if ((hh == 1) || (u == n_sub - 1)) commit(&kv_empty[slot_of(vi)]);
      if (u2 < n_sub && ((u2 & 1) == 1 || u2 == n_sub - 1)) commit(&kv_empty[slot_of(ki2)]);
      if (u == 0 && n_sub > 2) commit(&kv_empty[0]);   // K_0: consumed by the prologue QKs (u = 0, 1)
    }
    commit(o_final);
  } else {
    const int t = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    const int q_row = q0 + t * BQ + row_in_tile;
    const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
    const uint32_t s_base = tmem_base + lane_addr + uint32_t(t * 128);
    const uint32_t o_tmem = tmem_base + lane_addr + 256 + uint32_t(t * 128);
    float m = -INFINITY, l = 0.f;
    for (int u = 0; u < n_sub; ++u) {
      const int buf = u & 1;
      const uint32_t s_tmem = s_base + uint32_t(buf * SUB);
      mbar_wait(&s_full[t * 2 + buf], uint32_t(u >> 1) & 1);
      tc_fence_after();
      uint32_t sr[64];
      B2F_TMEM_LD_X32(s_tmem + 0, (sr + 0));
      B2F_TMEM_LD_X32(s_tmem + 32, (sr + 32));
      tmem_wait_ld();
      const int kv0 = u * SUB;
      const bool need_mask = (kv0 + SUB > p.Skv) || (p.causal && kv0 + SUB > q0 + t * BQ);
      if (need_mask) {
        const int limit = p.causal ? min(p.Skv, q_row + 1) : p.Skv;
#pragma unroll
        for (int c = 0; c < SUB; ++c)
          if (kv0 + c >= limit) sr[c] = 0xff800000u;
      }
      float mx4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) mx4[i] = fmaxf(__uint_as_float(sr[2 * i]), __uint_as_float(sr[2 * i + 1]));
#pragma unroll
      for (int c = 8; c < SUB; c += 8)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          mx4[i] = fmax3(mx4[i], __uint_as_float(sr[c + 2 * i]), __uint_as_float(sr[c + 2 * i + 1]));
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float m_new = fmaxf(m, mx * p.scale_log2);
      const bool grow = (m_new - m) > 8.0f;
      const float m_use = grow ? m_new : m;
      const float alpha = grow ? ex2(m - m_use) : 1.0f;
      const float neg_m = (m_use == -INFINITY) ? 0.f : -m_use;
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        float x0, x1;
        ffma2(x0, x1, __uint_as_float(sr[2 * c]), __uint_as_float(sr[2 * c + 1]), p.scale_log2, p.scale_log2,
              neg_m, neg_m);
        float p0, p1;
        if (POLY && (c % (POLY ? POLY : 1)) == (POLY ? POLY : 1) - 1) {
          ex2_poly2(x0, x1, p0, p1);
        } else {
          p0 = ex2(x0);
          p1 = ex2(x1);
        }
        const int a = (c & 1) * 2;
        fadd2(sum4[a], sum4[a + 1], sum4[a], sum4[a + 1], p0, p1);
        pk[c] = pack_bf16x2(p0, p1);
      }
      B2F_TMEM_ST_X32(s_tmem, pk);
      l = l * alpha + ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
      m = m_use;
      if (u > 0 && __any_sync(0xffffffffu, grow)) {
        // O must be quiescent: PV(t, u-1) is the last one issued (PV(t,u) waits for the arrive below)
        mbar_wait(&o_done[t], uint32_t(u - 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t o[32];
          B2F_TMEM_LD_X32(o_tmem + c0, o);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
          B2F_TMEM_ST_X32(o_tmem + c0, o);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t * 2 + buf]);
    }
    mbar_wait(o_final, 0);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const bool row_ok = q_row < p.Sq;
    __nv_bfloat16* out_row = p.out + ((long long)b * p.Sq + q_row) * p.ldo + (long long)h * DH;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t o[32];
      __syncwarp();
      B2F_TMEM_LD_X32(o_tmem + c0, o);
      tmem_wait_ld();
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
          v.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
          v.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
          v.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(out_row + c0 + g * 8) = v;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ================================================================================================
// v3: ONE 128-row Q tile per CTA, S double-buffered in TMEM.
//   TMEM: S0|P0 [0,128)  S1|P1 [128,256)  O [256,384)   (512 allocated)
//   MMA warp:  QK(0)->S0, QK(1)->S1, then for every block j:  wait P(j) -> PV(j) -> QK(j+2) into the
//              buffer PV(j) just released.  S(j+1) is therefore computed WHILE the warpgroup runs the
//              softmax of block j: in steady state the softmax warpgroup never waits for the tensor pipe
//              and the tensor pipe only waits for P — the two-tile kernel above serialises
//              QK -> softmax -> PV inside each tile and leaves its warpgroups waiting for S a third of
//              the time (ncu: 42 % of softmax-warp samples).
//   P is still handed over in two 64-column halves; O is rescaled lazily.  `o_done` completes once per
//   PV(j); the warpgroup consults it only before a rescale (unambiguous: PV(j) cannot be issued before the
//   warpgroup publishes P(j)); the epilogue waits on a separate one-shot `o_final` barrier.
constexpr int V3_SLOTS = 6;
constexpr int V3_THREADS = 192;
constexpr int V3_SMEM = (1 + V3_SLOTS) * TILE_BYTES + 256 + 1024;

template <int POLY>
__global__ void __launch_bounds__(V3_THREADS, 1)
attn_fwd_kernel_v3(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* q_smem = smem;
  uint8_t* kv_smem = smem + TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (1 + V3_SLOTS) * TILE_BYTES);
  uint64_t* q_full = bars;                   // 1
  uint64_t* kv_full = bars + 1;              // V3_SLOTS
  uint64_t* kv_empty = kv_full + V3_SLOTS;   // V3_SLOTS
  uint64_t* s_full = kv_empty + V3_SLOTS;    // [buf] = 2
  uint64_t* p_full = s_full + 2;             // [buf][half] = 4
  uint64_t* o_done = p_full + 4;             // 1
  uint64_t* o_final = o_done + 1;            // 1: everything issued has completed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_final + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.H / p.Hkv);
  const int q0 = qt * BQ;
  int kv_len = p.Skv;
  if (p.causal) kv_len = min(p.Skv, q0 + BQ);
  const int n_kv = (kv_len + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < V3_SLOTS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    for (int i = 0; i < 4; ++i) mbar_init(&p_full[i], 4);
    mbar_init(o_done, 1);
    mbar_init(o_final, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, TILE_BYTES);
      for (int half = 0; half < 2; ++half)
        tma_load_3d(q_smem + half * (TILE_BYTES / 2), &tmQ, q_full, h * DH + half * 64, q0, b);
      // ring order: K_0, K_1, V_0, K_2, V_1, K_3, ... (the order in which the MMA warp consumes tiles)
      int slot = 0;
      uint32_t phase = 0;
      auto load = [&](const CUtensorMap* tm, int j) {
        mbar_wait(&kv_empty[slot], phase ^ 1);
        mbar_expect_tx(&kv_full[slot], TILE_BYTES);
        uint8_t* dst = kv_smem + slot * TILE_BYTES;
        tma_load_3d(dst, tm, &kv_full[slot], hk * DH, j * BKV, b);
        tma_load_3d(dst + TILE_BYTES / 2, tm, &kv_full[slot], hk * DH + 64, j * BKV, b);
        if (++slot == V3_SLOTS) {
          slot = 0;
          phase ^= 1;
        }
      };
      load(&tmK, 0);
      if (n_kv > 1) load(&tmK, 1);
      for (int j = 0; j < n_kv; ++j) {
        load(&tmV, j);
        if (j + 2 < n_kv) load(&tmK, j + 2);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, BKV, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, DH, 1);
    const uint64_t dq_base = make_sdesc_sw128(smem_u32(q_smem), 16, 1024);
    const uint64_t dk_base = make_sdesc_sw128(smem_u32(kv_smem), 16, 1024);
    const uint64_t dv_base = make_sdesc_sw128(smem_u32(kv_smem), TILE_BYTES / 2, 1024);
    int slot = 0;
    uint32_t phase = 0;
    auto advance = [&]() {
      if (++slot == V3_SLOTS) {
        slot = 0;
        phase ^= 1;
      }
    };
    auto issue_qk = [&](int buf, int k_slot) {
      const uint32_t d = tmem_base + uint32_t(buf * 128);
      const uint64_t kd = dk_base + uint64_t((k_slot * TILE_BYTES) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) {
          const uint64_t off = uint64_t(((k >> 2) * (TILE_BYTES / 2) + (k & 3) * 32) >> 4);
          umma_ss(d, dq_base + off, kd + off, idesc_qk, k != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[buf]);
        umma_commit(&kv_empty[k_slot]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int buf, int v_slot, int hf, bool first, bool last_half) {
      const uint32_t d = tmem_base + 256;
      const uint32_t pa = tmem_base + uint32_t(buf * 128 + hf * 32);
      const uint64_t vd = dv_base + uint64_t((v_slot * TILE_BYTES + hf * 8192) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ts(d, pa + k * 8, vd + uint64_t((k * 2048) >> 4), idesc_pv, (first && k == 0) ? 0u : 1u);
        if (last_half) {
          umma_commit(o_done);
          umma_commit(&kv_empty[v_slot]);
        }
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    // prologue: S(0), S(1)
    for (int j = 0; j < 2 && j < n_kv; ++j) {
      mbar_wait(&kv_full[slot], phase);
      tc_fence_after();
      issue_qk(j, slot);
      advance();
    }
    for (int j = 0; j < n_kv; ++j) {
      const int buf = j & 1;
      mbar_wait(&kv_full[slot], phase);      // V_j
      const int v_slot = slot;
      advance();
      mbar_wait(&p_full[buf * 2 + 0], uint32_t(j >> 1) & 1);
      tc_fence_after();
      issue_pv(buf, v_slot, 0, j == 0, false);
      mbar_wait(&p_full[buf * 2 + 1], uint32_t(j >> 1) & 1);
      tc_fence_after();
      issue_pv(buf, v_slot, 1, false, true);
      if (j + 2 < n_kv) {
        mbar_wait(&kv_full[slot], phase);    // K_{j+2}
        tc_fence_after();
        issue_qk(buf, slot);
        advance();
      }
    }
    if (elect_one()) umma_commit(o_final);
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    const int q_row = q0 + row_in_tile;
    const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
    const uint32_t o_tmem = tmem_base + lane_addr + 256;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const int buf = j & 1;
      const uint32_t s_tmem = tmem_base + lane_addr + uint32_t(buf * 128);
      mbar_wait(&s_full[buf], uint32_t(j >> 1) & 1);
      tc_fence_after();
      uint32_t sr[128];
      B2F_TMEM_LD_X32(s_tmem + 0, (sr + 0));
      B2F_TMEM_LD_X32(s_tmem + 32, (sr + 32));
      B2F_TMEM_LD_X32(s_tmem + 64, (sr + 64));
      B2F_TMEM_LD_X32(s_tmem + 96, (sr + 96));
      tmem_wait_ld();
      const int kv0 = j * BKV;
      const bool need_mask = (kv0 + BKV > p.Skv) || (p.causal && kv0 + BKV > q0);
      if (need_mask) {
        const int limit = p.causal ? min(p.Skv, q_row + 1) : p.Skv;
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if (kv0 + c >= limit) sr[c] = 0xff800000u;
      }
      float mx4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) mx4[i] = fmaxf(__uint_as_float(sr[2 * i]), __uint_as_float(sr[2 * i + 1]));
#pragma unroll
      for (int c = 8; c < 128; c += 8)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          mx4[i] = fmax3(mx4[i], __uint_as_float(sr[c + 2 * i]), __uint_as_float(sr[c + 2 * i + 1]));
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float m_new = fmaxf(m, mx * p.scale_log2);
      const bool grow = (m_new - m) > 8.0f;
      const float m_use = grow ? m_new : m;
      const float alpha = grow ? ex2(m - m_use) : 1.0f;
      const float neg_m = (m_use == -INFINITY) ? 0.f : -m_use;
      if (j > 0 && __any_sync(0xffffffffu, grow)) {
        // O must be quiescent: PV(j-1) is the last one issued (PV(j) needs the arrives below).  Completed
        // phases of o_done are j-1 or j here, so the parity of phase j-1 is unambiguous.
        mbar_wait(o_done, uint32_t(j - 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t o[32];
          B2F_TMEM_LD_X32(o_tmem + c0, o);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
          B2F_TMEM_ST_X32(o_tmem + c0, o);
        }
      }
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          float x0, x1;
          ffma2(x0, x1, __uint_as_float(sr[half * 64 + 2 * c]), __uint_as_float(sr[half * 64 + 2 * c + 1]),
                p.scale_log2, p.scale_log2, neg_m, neg_m);
          float p0, p1;
          if (POLY && (c % (POLY ? POLY : 1)) == (POLY ? POLY : 1) - 1) {
            ex2_poly2(x0, x1, p0, p1);
          } else {
            p0 = ex2(x0);
            p1 = ex2(x1);
          }
          const int a = (c & 1) * 2;
          fadd2(sum4[a], sum4[a + 1], sum4[a], sum4[a + 1], p0, p1);
          pk[c] = pack_bf16x2(p0, p1);
        }
        B2F_TMEM_ST_X32(s_tmem + half * 32, pk);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[buf * 2 + half]);
      }
      l = l * alpha + ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
      m = m_use;
    }
    mbar_wait(o_final, 0);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const bool row_ok = q_row < p.Sq;
    __nv_bfloat16* out_row = p.out + ((long long)b * p.Sq + q_row) * p.ldo + (long long)h * DH;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t o[32];
      __syncwarp();
      B2F_TMEM_LD_X32(o_tmem + c0, o);
      tmem_wait_ld();
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
          v.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
          v.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
          v.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(out_row + c0 + g * 8) = v;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ================================================================================================
// v4: v3 (one 128-row Q tile per CTA, S double-buffered in TMEM so QK(j+1) runs during softmax(j)) with TWO
// softmax warpgroups that split the COLUMNS of every S block: WG h owns columns [64h, 64h+64) and produces the
// P half the MMA warp consumes as PV half h.  Two warps per SM sub-partition hide each other's TMEM-load /
// max / barrier latencies (v3's single warpgroup could not), while the softmax of block j still overlaps the
// tensor work of block j+1.  The row maximum is exchanged between the two halves through smem.
constexpr int V4_SLOTS = 5;
constexpr int V4_THREADS = 320;
constexpr int V4_XCH_BYTES = 3 * 2 * 128 * 4;   // row-max exchange [parity][wg][row] + row-sum exchange [wg][row]
constexpr int V4_SMEM = (1 + V4_SLOTS) * TILE_BYTES + 256 + V4_XCH_BYTES + 1024;

template <int POLY>
__global__ void __launch_bounds__(V4_THREADS, 1)
attn_fwd_kernel_v4(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* q_smem = smem;
  uint8_t* kv_smem = smem + TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (1 + V4_SLOTS) * TILE_BYTES);
  uint64_t* q_full = bars;                   // 1
  uint64_t* kv_full = bars + 1;              // V4_SLOTS
  uint64_t* kv_empty = kv_full + V4_SLOTS;   // V4_SLOTS
  uint64_t* s_full = kv_empty + V4_SLOTS;    // [buf] = 2
  uint64_t* p_full = s_full + 2;             // [buf][half] = 4
  uint64_t* o_done = p_full + 4;             // 1
  uint64_t* o_final = o_done + 1;            // 1: everything issued has completed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_final + 1);
  float* xch = reinterpret_cast<float*>(smem + (1 + V4_SLOTS) * TILE_BYTES + 256);   // [3][2][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.H / p.Hkv);
  const int q0 = qt * BQ;
  int kv_len = p.Skv;
  if (p.causal) kv_len = min(p.Skv, q0 + BQ);
  const int n_kv = (kv_len + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < V4_SLOTS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    for (int i = 0; i < 4; ++i) mbar_init(&p_full[i], 4);
    mbar_init(o_done, 1);
    mbar_init(o_final, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, TILE_BYTES);
      for (int half = 0; half < 2; ++half)
        tma_load_3d(q_smem + half * (TILE_BYTES / 2), &tmQ, q_full, h * DH + half * 64, q0, b);
      // ring order: K_0, K_1, V_0, K_2, V_1, K_3, ... (the order in which the MMA warp consumes tiles)
      int slot = 0;
      uint32_t phase = 0;
      auto load = [&](const CUtensorMap* tm, int j) {
        mbar_wait(&kv_empty[slot], phase ^ 1);
        mbar_expect_tx(&kv_full[slot], TILE_BYTES);
        uint8_t* dst = kv_smem + slot * TILE_BYTES;
        tma_load_3d(dst, tm, &kv_full[slot], hk * DH, j * BKV, b);
        tma_load_3d(dst + TILE_BYTES / 2, tm, &kv_full[slot], hk * DH + 64, j * BKV, b);
        if (++slot == V4_SLOTS) {
          slot = 0;
          phase ^= 1;
        }
      };
      load(&tmK, 0);
      if (n_kv > 1) load(&tmK, 1);
      for (int j = 0; j < n_kv; ++j) {
        load(&tmV, j);
        if (j + 2 < n_kv) load(&tmK, j + 2);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, BKV, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, DH, 1);
    const uint64_t dq_base = make_sdesc_sw128(smem_u32(q_smem), 16, 1024);
    const uint64_t dk_base = make_sdesc_sw128(smem_u32(kv_smem), 16, 1024);
    const uint64_t dv_base = make_sdesc_sw128(smem_u32(kv_smem), TILE_BYTES / 2, 1024);
    int slot = 0;
    uint32_t phase = 0;
    auto advance = [&]() {
      if (++slot == V4_SLOTS) {
        slot = 0;
        phase ^= 1;
      }
    };
    auto issue_qk = [&](int buf, int k_slot) {
      const uint32_t d = tmem_base + uint32_t(buf * 128);
      const uint64_t kd = dk_base + uint64_t((k_slot * TILE_BYTES) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) {
          const uint64_t off = uint64_t(((k >> 2) * (TILE_BYTES / 2) + (k & 3) * 32) >> 4);
          umma_ss(d, dq_base + off, kd + off, idesc_qk, k != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[buf]);
        umma_commit(&kv_empty[k_slot]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int buf, int v_slot, int hf, bool first, bool last_half) {
      const uint32_t d = tmem_base + 256;
      const uint32_t pa = tmem_base + uint32_t(buf * 128 + hf * 64);   // P half hf sits on WG hf's own S columns
      const uint64_t vd = dv_base + uint64_t((v_slot * TILE_BYTES + hf * 8192) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ts(d, pa + k * 8, vd + uint64_t((k * 2048) >> 4), idesc_pv, (first && k == 0) ? 0u : 1u);
        if (last_half) {
          umma_commit(o_done);
          umma_commit(&kv_empty[v_slot]);
        }
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    // prologue: S(0), S(1)
    for (int j = 0; j < 2 && j < n_kv; ++j) {
      mbar_wait(&kv_full[slot], phase);
      tc_fence_after();
      issue_qk(j, slot);
      advance();
    }
    for (int j = 0; j < n_kv; ++j) {
      const int buf = j & 1;
      mbar_wait(&kv_full[slot], phase);      // V_j
      const int v_slot = slot;
      advance();
      mbar_wait(&p_full[buf * 2 + 0], uint32_t(j >> 1) & 1);
      tc_fence_after();
      issue_pv(buf, v_slot, 0, j == 0, false);
      mbar_wait(&p_full[buf * 2 + 1], uint32_t(j >> 1) & 1);
      tc_fence_after();
      issue_pv(buf, v_slot, 1, false, true);
      if (j + 2 < n_kv) {
        mbar_wait(&kv_full[slot], phase);    // K_{j+2}
        tc_fence_after();
        issue_qk(buf, slot);
        advance();
      }
    }
    if (elect_one()) umma_commit(o_final);
    __syncwarp();
  } else {
    // Two warpgroups share the tile's rows: WG hw owns S/P columns [64 hw, 64 hw + 64) of every KV block.  Warps
    // `warp` and `warp ^ 4`... (2..5 = WG0, 6..9 = WG1; equal `warp & 3` = same TMEM lane quarter = same 32 rows)
    // exchange their partial row maxima through smem around a 64-thread named barrier.
    const int hw = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    const int q_row = q0 + row_in_tile;
    const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
    const uint32_t o_tmem = tmem_base + lane_addr + 256;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const int buf = j & 1;
      const uint32_t s_tmem = tmem_base + lane_addr + uint32_t(buf * 128 + hw * 64);
      mbar_wait(&s_full[buf], uint32_t(j >> 1) & 1);
      tc_fence_after();
      uint32_t sr[64];
      B2F_TMEM_LD_X32(s_tmem + 0, (sr + 0));
      B2F_TMEM_LD_X32(s_tmem + 32, (sr + 32));
      tmem_wait_ld();
      const int kv0 = j * BKV + hw * 64;
      const bool need_mask = (kv0 + 64 > p.Skv) || (p.causal && kv0 + 64 > q0);
      if (need_mask) {
        const int limit = p.causal ? min(p.Skv, q_row + 1) : p.Skv;
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (kv0 + c >= limit) sr[c] = 0xff800000u;
      }
      float mx4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) mx4[i] = fmaxf(__uint_as_float(sr[2 * i]), __uint_as_float(sr[2 * i + 1]));
#pragma unroll
      for (int c = 8; c < 64; c += 8)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          mx4[i] = fmax3(mx4[i], __uint_as_float(sr[c + 2 * i]), __uint_as_float(sr[c + 2 * i + 1]));
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      // row max over both column halves (slots double-buffered by block parity: a slot is rewritten only after
      // the partner has passed the next barrier, i.e. after it read this one)
      float* xm = xch + (j & 1) * 256;
      xm[hw * 128 + row_in_tile] = mx;
      named_bar_sync(1 + quarter, 64);
      mx = fmaxf(mx, xm[(hw ^ 1) * 128 + row_in_tile]);
      const float m_new = fmaxf(m, mx * p.scale_log2);
      const bool grow = (m_new - m) > 8.0f;
      const float m_use = grow ? m_new : m;
      const float alpha = grow ? ex2(m - m_use) : 1.0f;
      const float neg_m = (m_use == -INFINITY) ? 0.f : -m_use;
      if (hw == 0 && j > 0 && __any_sync(0xffffffffu, grow)) {
        // WG0 alone rescales O (all 128 columns of its rows) BEFORE it publishes its P half: PV(j) half 0 is the
        // first MMA that touches O again and it waits for WG0's arrive.  O is quiescent: PV(j-1) is the last one
        // issued, and completed phases of o_done are j-1 or j here, so the parity of phase j-1 is unambiguous.
        mbar_wait(o_done, uint32_t(j - 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t o[32];
          B2F_TMEM_LD_X32(o_tmem + c0, o);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
          B2F_TMEM_ST_X32(o_tmem + c0, o);
        }
      }
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        float x0, x1;
        ffma2(x0, x1, __uint_as_float(sr[2 * c]), __uint_as_float(sr[2 * c + 1]), p.scale_log2, p.scale_log2, neg_m,
              neg_m);
        float p0, p1;
        if (POLY && (c % (POLY ? POLY : 1)) == (POLY ? POLY : 1) - 1) {
          ex2_poly2(x0, x1, p0, p1);
        } else {
          p0 = ex2(x0);
          p1 = ex2(x1);
        }
        const int a = (c & 1) * 2;
        fadd2(sum4[a], sum4[a + 1], sum4[a], sum4[a + 1], p0, p1);
        pk[c] = pack_bf16x2(p0, p1);
      }
      B2F_TMEM_ST_X32(s_tmem, pk);   // bf16 P over the first 32 of this warpgroup's own 64 S columns
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[buf * 2 + hw]);
      l = l * alpha + ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
      m = m_use;
    }
    // total row sum = both halves' partial sums (identical rescale history)
    float* xl = xch + 512;
    xl[hw * 128 + row_in_tile] = l;
    named_bar_sync(1 + quarter, 64);
    l += xl[(hw ^ 1) * 128 + row_in_tile];
    mbar_wait(o_final, 0);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const bool row_ok = q_row < p.Sq;
    __nv_bfloat16* out_row = p.out + ((long long)b * p.Sq + q_row) * p.ldo + (long long)h * DH;
#pragma unroll 1
    for (int c0 = hw * 64; c0 < hw * 64 + 64; c0 += 32) {   // each warpgroup writes its half of the head
      uint32_t o[32];
      __syncwarp();
      B2F_TMEM_LD_X32(o_tmem + c0, o);
      tmem_wait_ld();
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
          v.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
          v.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
          v.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(out_row + c0 + g * 8) = v;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ================================================================================================
// CTA-PAIR kernel (cluster of 2, tcgen05 cta_group::2).  The two-tile kernel above is bound by shared-memory
// bandwidth, not by the softmax: per KV block and Q tile an M=128 SS-MMA reads the whole Q tile (32 KB) and the
// whole K tile (32 KB) — 128 B/clk during QK, the full smem bandwidth of an SM — plus V for P·V and the TMA
// writes of K/V (~125 B/clk on average; every single-tile / sub-block variant kept or raised that figure, which
// is why none was faster).  Here two CTAs own 512 query rows of a head: every MMA is M=256 (128 rows per CTA)
// and each CTA stores only HALF of every K tile (64 of its 128 kv rows) and HALF of every V tile (64 of its 128
// dh columns); the hardware shares the B halves between the two SMs.  Per tile and KV block an SM now moves
// 16 KB (TMA) + 48 KB (QK: Q 32 + K/2 16) + 16 KB (P·V: V/2) = 80 KB instead of 128 KB.
//   rows of pair pr:  tile t of CTA c = [512 pr + 256 t + 128 c, +128)
//   TMEM per CTA (its 128 lanes of the M=256 accumulators): S0|P0, S1|P1, O0, O1 as above
//   barriers: q_full / kv_full / p_full live on CTA 0 (the only MMA issuer; CTA 1's TMA bytes and softmax warps
//   signal them remotely), s_full / o_done / kv_empty exist in both CTAs and receive multicast commits.
// Non-causal, no bias (the FLUX joint attention); everything else goes to the kernels above.
constexpr int P2_SLOTS = 8;                       // half tiles of 16 KB: K_j/2, V_j/2 alternating
constexpr int P2_HALF = TILE_BYTES / 2;
constexpr int P2_SMEM = 2 * TILE_BYTES + P2_SLOTS * P2_HALF + 256 + 1024;

template <int POLY>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ATTN_THREADS, 1)
attn_fwd_kernel_2cta(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* q_smem = smem;                       // 2 tiles of this CTA's rows
  uint8_t* kv_smem = smem + 2 * TILE_BYTES;     // P2_SLOTS half tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * TILE_BYTES + P2_SLOTS * P2_HALF);
  uint64_t* q_full = bars;                      // 1   (CTA 0)
  uint64_t* kv_full = bars + 1;                 // P2_SLOTS (CTA 0)
  uint64_t* kv_empty = kv_full + P2_SLOTS;      // P2_SLOTS (both)
  uint64_t* s_full = kv_empty + P2_SLOTS;       // 2 (both)
  uint64_t* p_full = s_full + 2;                // [tile][half] = 4 (CTA 0)
  uint64_t* o_done = p_full + 4;                // 2 (both)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int pr = blockIdx.x >> 1, h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.H / p.Hkv);
  const int q0 = pr * 4 * BQ;                   // first row of the pair
  const int n_kv = (p.Skv + BKV - 1) / BKV;
  const bool t1_active = q0 + 2 * BQ < p.Sq;    // second tile of BOTH CTAs beyond Sq: skip it

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 2);                       // one arrive per CTA's producer (+ both CTAs' TMA bytes)
    for (int i = 0; i < P2_SLOTS; ++i) {
      mbar_init(&kv_full[i], 2);
      mbar_init(&kv_empty[i], 1);               // one multicast commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&o_done[i], 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&p_full[i], 8);   // 4 softmax warps x 2 CTAs
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(tmem_ptr, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer (one per CTA)
      if (leader)
        mbar_expect_tx(q_full, 4 * TILE_BYTES);
      else
        mbar_arrive_cta0(q_full);
      for (int t = 0; t < 2; ++t)
        for (int half = 0; half < 2; ++half)
          tma_load_3d_2cta(q_smem + t * TILE_BYTES + half * (TILE_BYTES / 2), &tmQ, q_full, h * DH + half * 64,
                           q0 + t * 2 * BQ + int(crank) * BQ, b);
      int slot = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        for (int kv = 0; kv < 2; ++kv) {  // K_j/2 then V_j/2
          mbar_wait(&kv_empty[slot], phase ^ 1);
          if (leader)
            mbar_expect_tx(&kv_full[slot], 2 * P2_HALF);
          else
            mbar_arrive_cta0(&kv_full[slot]);
          uint8_t* dst = kv_smem + slot * P2_HALF;
          if (kv == 0) {
            // this CTA's 64 kv rows of K_j, both 64-column dh halves (8 KB each)
            tma_load_3d_2cta(dst, &tmK, &kv_full[slot], hk * DH, j * BKV + int(crank) * 64, b);
            tma_load_3d_2cta(dst + P2_HALF / 2, &tmK, &kv_full[slot], hk * DH + 64, j * BKV + int(crank) * 64, b);
          } else {
            // this CTA's 64 dh columns of V_j, all 128 kv rows
            tma_load_3d_2cta(dst, &tmV, &kv_full[slot], hk * DH + int(crank) * 64, j * BKV, b);
          }
          if (++slot == P2_SLOTS) {
            slot = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------------------------------------------------------- MMA issuer (warp-uniform loop)
      constexpr uint32_t idesc_qk = make_idesc_bf16(2 * BQ, BKV, 0);  // M = 256, B = K (K-major, 64 rows per CTA)
      constexpr uint32_t idesc_pv = make_idesc_bf16(2 * BQ, DH, 1);   // M = 256, B = V (MN-major, 64 dh per CTA)
      const uint64_t dq_base = make_sdesc_sw128(smem_u32(q_smem), 16, 1024);
      const uint64_t dk_base = make_sdesc_sw128(smem_u32(kv_smem), 16, 1024);
      const uint64_t dv_base = make_sdesc_sw128(smem_u32(kv_smem), P2_HALF, 1024);
      int slot = 0;
      uint32_t phase = 0;
      auto issue_qk = [&](int t, int k_slot) {
        const uint32_t d = tmem_base + uint32_t(t * 128);
        const uint64_t qd = dq_base + uint64_t((t * TILE_BYTES) >> 4);
        const uint64_t kd = dk_base + uint64_t((k_slot * P2_HALF) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) {
            const uint64_t qoff = uint64_t(((k >> 2) * (TILE_BYTES / 2) + (k & 3) * 32) >> 4);
            const uint64_t koff = uint64_t(((k >> 2) * (P2_HALF / 2) + (k & 3) * 32) >> 4);
            umma_ss_2cta(d, qd + qoff, kd + koff, idesc_qk, k != 0 ? 1u : 0u);
          }
        }
        __syncwarp();
      };
      auto issue_pv = [&](int t, int v_slot, int hf, bool first) {
        const uint32_t d = tmem_base + 256 + uint32_t(t * 128);
        const uint32_t pa = tmem_base + uint32_t(t * 128 + hf * 32);
        const uint64_t vd = dv_base + uint64_t((v_slot * P2_HALF + hf * 8192) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ts_2cta(d, pa + k * 8, vd + uint64_t((k * 2048) >> 4), idesc_pv, (first && k == 0) ? 0u : 1u);
        }
        __syncwarp();
      };
      auto commit = [&](uint64_t* bar) {
        if (elect_one()) umma_commit_2cta(bar);
        __syncwarp();
      };
      auto advance = [&]() {
        if (++slot == P2_SLOTS) {
          slot = 0;
          phase ^= 1;
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[slot], phase);
      tc_fence_after();
      issue_qk(0, slot);
      commit(&s_full[0]);
      if (t1_active) {
        issue_qk(1, slot);
        commit(&s_full[1]);
      }
      commit(&kv_empty[slot]);
      advance();
      for (int j = 0; j < n_kv; ++j) {
        const int v_slot = slot;
        const uint32_t v_phase = phase;
        advance();
        const int k_slot = slot;
        const uint32_t k_phase = phase;
        const bool more = (j + 1 < n_kv);
        if (more) advance();
        mbar_wait(&kv_full[v_slot], v_phase);
        mbar_wait(&p_full[0], j & 1);
        tc_fence_after();
        issue_pv(0, v_slot, 0, j == 0);
        mbar_wait(&p_full[1], j & 1);
        tc_fence_after();
        issue_pv(0, v_slot, 1, false);
        if (more) {
          mbar_wait(&kv_full[k_slot], k_phase);
          tc_fence_after();
          issue_qk(0, k_slot);
          commit(&s_full[0]);
        }
        if (t1_active) {
          mbar_wait(&p_full[2], j & 1);
          tc_fence_after();
          issue_pv(1, v_slot, 0, j == 0);
          mbar_wait(&p_full[3], j & 1);
          tc_fence_after();
          issue_pv(1, v_slot, 1, false);
        }
        commit(&kv_empty[v_slot]);
        if (more) {
          if (t1_active) {
            issue_qk(1, k_slot);
            commit(&s_full[1]);
          }
          commit(&kv_empty[k_slot]);
        }
      }
      commit(&o_done[0]);
      commit(&o_done[1]);
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups (both CTAs)
    const int t = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    const int q_row = q0 + t * 2 * BQ + int(crank) * BQ + row_in_tile;
    const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_addr + uint32_t(t * 128);
    const uint32_t o_tmem = tmem_base + lane_addr + 256 + uint32_t(t * 128);
    float m = -INFINITY, l = 0.f;
    if (t == 0 || t1_active) {
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        uint32_t sr[128];
        B2F_TMEM_LD_X32(s_tmem + 0, (sr + 0));
        B2F_TMEM_LD_X32(s_tmem + 32, (sr + 32));
        B2F_TMEM_LD_X32(s_tmem + 64, (sr + 64));
        B2F_TMEM_LD_X32(s_tmem + 96, (sr + 96));
        tmem_wait_ld();
        const int kv0 = j * BKV;
        if (kv0 + BKV > p.Skv) {
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if (kv0 + c >= p.Skv) sr[c] = 0xff800000u;  // -inf
        }
        float mx4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) mx4[i] = fmaxf(__uint_as_float(sr[2 * i]), __uint_as_float(sr[2 * i + 1]));
#pragma unroll
        for (int c = 8; c < 128; c += 8)
#pragma unroll
          for (int i = 0; i < 4; ++i)
            mx4[i] = fmax3(mx4[i], __uint_as_float(sr[c + 2 * i]), __uint_as_float(sr[c + 2 * i + 1]));
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        const float m_new = fmaxf(m, mx * p.scale_log2);
        const bool grow = (m_new - m) > 8.0f;     // lazy rescale (P stays < 2^8)
        const float m_use = grow ? m_new : m;
        const float alpha = grow ? ex2(m - m_use) : 1.0f;
        const float neg_m = (m_use == -INFINITY) ? 0.f : -m_use;
        if (j > 0 && __any_sync(0xffffffffu, grow)) {
          // S_t(j) ready proves P_t.V_{j-1} completed (same issue order as the single-CTA kernel); P_t.V_j cannot
          // start before the arrives below (it needs all 8 warps of both CTAs), so this CTA's O_t lanes are quiescent
#pragma unroll 1
          for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t o[32];
            B2F_TMEM_LD_X32(o_tmem + c0, o);
            tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
            B2F_TMEM_ST_X32(o_tmem + c0, o);
          }
        }
        float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t pk[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            float x0, x1;
            ffma2(x0, x1, __uint_as_float(sr[half * 64 + 2 * c]), __uint_as_float(sr[half * 64 + 2 * c + 1]),
                  p.scale_log2, p.scale_log2, neg_m, neg_m);
            float p0, p1;
            if (POLY && (c % (POLY ? POLY : 1)) == (POLY ? POLY : 1) - 1) {
              ex2_poly2(x0, x1, p0, p1);
            } else {
              p0 = ex2(x0);
              p1 = ex2(x1);
            }
            const int a = (c & 1) * 2;
            fadd2(sum4[a], sum4[a + 1], sum4[a], sum4[a + 1], p0, p1);
            pk[c] = pack_bf16x2(p0, p1);
          }
          B2F_TMEM_ST_X32(s_tmem + half * 32, pk);
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cta0(&p_full[t * 2 + half]);   // the issuer's barrier (remote from CTA 1)
        }
        l = l * alpha + ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
        m = m_use;
      }
      mbar_wait(&o_done[t], 0);
      tc_fence_after();
      const float inv_l = 1.0f / l;
      const bool row_ok = q_row < p.Sq;
      __nv_bfloat16* out_row = p.out + ((long long)b * p.Sq + q_row) * p.ldo + (long long)h * DH;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t o[32];
        __syncwarp();
        B2F_TMEM_LD_X32(o_tmem + c0, o);
        tmem_wait_ld();
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
            v.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
            v.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
            v.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(out_row + c0 + g * 8) = v;
          }
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();   // both CTAs are done with each other's barriers / smem halves / TMEM
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

// ================================================================================================
// v6: the two-tile kernel with FOUR softmax warpgroups — (tile, column half): every S block's columns are split
// between two warpgroups (64 each), so four softmax warps share an SM sub-partition instead of two and the
// MUFU / FMA / ALU work of one overlaps the TMEM-load / max / barrier latencies of the others.  The MMA warp is
// unchanged (P is already consumed in two 64-column halves: half h is now produced by warpgroup (t, h)).
constexpr int V6_THREADS = 64 + 16 * 32;
constexpr int V6_SMEM = (2 + KV_SLOTS) * TILE_BYTES + 256 + 6 * 1024 + 1024;

template <int POLY>
__global__ void __launch_bounds__(V6_THREADS, 1)
attn_fwd_kernel_v6(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* q_smem = smem;                       // 2 tiles
  uint8_t* kv_smem = smem + 2 * TILE_BYTES;     // KV_SLOTS tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (2 + KV_SLOTS) * TILE_BYTES);
  uint64_t* q_full = bars;            // 1
  uint64_t* kv_full = bars + 1;       // KV_SLOTS
  uint64_t* kv_empty = kv_full + KV_SLOTS;
  uint64_t* s_full = kv_empty + KV_SLOTS;  // 2
  uint64_t* p_full = s_full + 2;           // [tile][half] = 4: P columns [0,64) and [64,128) handed over separately
  uint64_t* o_done = p_full + 4;           // 2
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_done + 2);
  float* xch = reinterpret_cast<float*>(smem + (2 + KV_SLOTS) * TILE_BYTES + 256);   // [3][tile][wg][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qpair = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.H / p.Hkv);
  const int q0 = qpair * 2 * BQ;

  // K/V blocks this CTA needs (causal: only up to its last query row; Sq == Skv assumed then)
  int kv_len = p.Skv;
  if (p.causal) kv_len = min(p.Skv, q0 + 2 * BQ);
  const int n_kv = (kv_len + BKV - 1) / BKV;
  // the second Q tile of the last pair may lie entirely beyond Sq (S = 8736 = 34*256 + 32): skip all of
  // its MMAs and its softmax warpgroup instead of multiplying zero rows
  const bool t1_active = q0 + BQ < p.Sq;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < KV_SLOTS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&o_done[i], 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&p_full[i], 4);  // one elected arrive per softmax warp
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      mbar_expect_tx(q_full, 2 * TILE_BYTES);
      for (int t = 0; t < 2; ++t)
        for (int half = 0; half < 2; ++half)
          tma_load_3d(q_smem + t * TILE_BYTES + half * (TILE_BYTES / 2), &tmQ, q_full,
                      h * DH + half * 64, q0 + t * BQ, b);
      int slot = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        for (int kv = 0; kv < 2; ++kv) {  // K_j then V_j
          mbar_wait(&kv_empty[slot], phase ^ 1);
          mbar_expect_tx(&kv_full[slot], TILE_BYTES);
          uint8_t* dst = kv_smem + slot * TILE_BYTES;
          const CUtensorMap* tm = kv == 0 ? &tmK : &tmV;
          tma_load_3d(dst, tm, &kv_full[slot], hk * DH, j * BKV, b);
          tma_load_3d(dst + TILE_BYTES / 2, tm, &kv_full[slot], hk * DH + 64, j * BKV, b);
          if (++slot == KV_SLOTS) {
            slot = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    // The WHOLE warp runs this loop (waits, descriptor arithmetic) so that the address math stays on
    // the uniform datapath; only the tcgen05.mma / tcgen05.commit instructions are predicated to one
    // lane.  (With the loop nested under `if (lane == 0)` every descriptor went through R2UR moves and
    // the issue thread, not the tensor pipe, paced the kernel: ncu showed it busy ~75 % of the time.)
    constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, BKV, 0);  // B = K tile, K-major
    constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, DH, 1);   // B = V tile, MN-major
    const uint32_t q_addr = smem_u32(q_smem);
    const uint32_t kv_addr = smem_u32(kv_smem);
    // descriptor of byte offset 0 of each buffer; every MMA operand is "base + constant" (one uniform
    // 64-bit add on the 14-bit address field, which cannot carry out for addresses < 256 KB)
    const uint64_t dq_base = make_sdesc_sw128(q_addr, 16, 1024);
    const uint64_t dk_base = make_sdesc_sw128(kv_addr, 16, 1024);
    const uint64_t dv_base = make_sdesc_sw128(kv_addr, TILE_BYTES / 2, 1024);
    int slot = 0;
    uint32_t phase = 0;
    auto issue_qk = [&](int t, int k_slot) {
      const uint32_t d = tmem_base + uint32_t(t * 128);
      const uint64_t qd = dq_base + uint64_t((t * TILE_BYTES) >> 4);
      const uint64_t kd = dk_base + uint64_t((k_slot * TILE_BYTES) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) {
          const uint64_t off = uint64_t(((k >> 2) * (TILE_BYTES / 2) + (k & 3) * 32) >> 4);
          umma_ss(d, qd + off, kd + off, idesc_qk, k != 0 ? 1u : 0u);
        }
      }
      __syncwarp();
    };
    // O_t += P_t[:, 64*hf : 64*hf+64] · V[64*hf : 64*hf+64, :]  (4 k-steps of 16 kv rows)
    auto issue_pv = [&](int t, int v_slot, int hf, bool first) {
      const uint32_t d = tmem_base + 256 + uint32_t(t * 128);
      const uint32_t pa = tmem_base + uint32_t(t * 128 + hf * 64);   // P half hf sits on warpgroup (t, hf)'s own S columns
      const uint64_t vd = dv_base + uint64_t((v_slot * TILE_BYTES + hf * 8192) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ts(d, pa + k * 8, vd + uint64_t((k * 2048) >> 4), idesc_pv, (first && k == 0) ? 0u : 1u);
      }
      __syncwarp();
    };
    auto commit = [&](uint64_t* bar) {
      if (elect_one()) umma_commit(bar);
      __syncwarp();
    };
    auto advance = [&]() {
      if (++slot == KV_SLOTS) {
        slot = 0;
        phase ^= 1;
      }
    };
    mbar_wait(q_full, 0);
    // prologue: S_t = Q_t K_0^T for both tiles
    mbar_wait(&kv_full[slot], phase);
    tc_fence_after();
    issue_qk(0, slot);
    commit(&s_full[0]);
    if (t1_active) {
      issue_qk(1, slot);
      commit(&s_full[1]);
    }
    commit(&kv_empty[slot]);
    advance();
    for (int j = 0; j < n_kv; ++j) {
      const int v_slot = slot;
      const uint32_t v_phase = phase;
      advance();
      const int k_slot = slot;  // K_{j+1} (if any)
      const uint32_t k_phase = phase;
      const bool more = (j + 1 < n_kv);
      if (more) advance();
      mbar_wait(&kv_full[v_slot], v_phase);
      // tile 0: the first half of P·V starts while the warpgroup still exponentiates the second half
      mbar_wait(&p_full[0], j & 1);
      tc_fence_after();
      issue_pv(0, v_slot, 0, j == 0);
      mbar_wait(&p_full[1], j & 1);
      tc_fence_after();
      issue_pv(0, v_slot, 1, false);
      if (more) {
        mbar_wait(&kv_full[k_slot], k_phase);
        tc_fence_after();
        issue_qk(0, k_slot);
        commit(&s_full[0]);
      }
      // tile 1
      if (t1_active) {
        mbar_wait(&p_full[2], j & 1);
        tc_fence_after();
        issue_pv(1, v_slot, 0, j == 0);
        mbar_wait(&p_full[3], j & 1);
        tc_fence_after();
        issue_pv(1, v_slot, 1, false);
      }
      commit(&kv_empty[v_slot]);
      if (more) {
        if (t1_active) {
          issue_qk(1, k_slot);
          commit(&s_full[1]);
        }
        commit(&kv_empty[k_slot]);
      }
    }
    commit(&o_done[0]);
    commit(&o_done[1]);
  } else {
    // ------------------------------------------------------------------ softmax: 4 warpgroups = (tile, column half)
    const int w = warp - 2;
    const int t = w >> 3;
    const int hw = (w >> 2) & 1;
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    const int q_row = q0 + t * BQ + row_in_tile;
    const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_addr + uint32_t(t * 128 + hw * 64);
    const uint32_t o_tmem = tmem_base + lane_addr + 256 + uint32_t(t * 128);
    const uint32_t bar_id = 1 + t * 4 + quarter;          // the two warps that hold the same 32 rows
    float* xt = xch + t * 256 + row_in_tile;              // + parity * 512 + wg * 128
    float m = -INFINITY, l = 0.f;
    if (t == 0 || t1_active) {
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      uint32_t sr[64];
      B2F_TMEM_LD_X32(s_tmem + 0, (sr + 0));
      B2F_TMEM_LD_X32(s_tmem + 32, (sr + 32));
      tmem_wait_ld();
      const int kv0 = j * BKV + hw * 64;
      const bool need_mask = (kv0 + 64 > p.Skv) || (p.causal && kv0 + 64 > q0 + t * BQ);
      if (need_mask) {
        const int limit = p.causal ? min(p.Skv, q_row + 1) : p.Skv;
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (kv0 + c >= limit) sr[c] = 0xff800000u;  // -inf
      }
      float mx4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) mx4[i] = fmaxf(__uint_as_float(sr[2 * i]), __uint_as_float(sr[2 * i + 1]));
#pragma unroll
      for (int c = 8; c < 64; c += 8)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          mx4[i] = fmax3(mx4[i], __uint_as_float(sr[c + 2 * i]), __uint_as_float(sr[c + 2 * i + 1]));
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      // row max across the two column halves (exchange slots double-buffered by block parity)
      float* xm = xt + (j & 1) * 512;
      xm[hw * 128] = mx;
      named_bar_sync(bar_id, 64);
      mx = fmaxf(mx, xm[(hw ^ 1) * 128]);
      const float m_new = fmaxf(m, mx * p.scale_log2);
      const bool grow = (m_new - m) > 8.0f;
      const float m_use = grow ? m_new : m;
      const float alpha = grow ? ex2(m - m_use) : 1.0f;
      const float neg_m = (m_use == -INFINITY) ? 0.f : -m_use;
      if (hw == 0 && j > 0 && __any_sync(0xffffffffu, grow)) {
        // warpgroup (t, 0) alone rescales O_t, before it publishes its P half (the first MMA to touch O_t again
        // waits for that arrive); S_t(j) being ready proves P_t.V_{j-1} completed
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 8) {   // 8 columns at a time: this thread also holds 64 S values
          uint32_t o[8];
          B2F_TMEM_LD_X8(o_tmem + c0, o);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 8; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
          B2F_TMEM_ST_X8(o_tmem + c0, o);
        }
      }
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        float x0, x1;
        ffma2(x0, x1, __uint_as_float(sr[2 * c]), __uint_as_float(sr[2 * c + 1]), p.scale_log2, p.scale_log2, neg_m,
              neg_m);
        float p0, p1;
        if (POLY && (c % (POLY ? POLY : 1)) == (POLY ? POLY : 1) - 1) {
          ex2_poly2(x0, x1, p0, p1);
        } else {
          p0 = ex2(x0);
          p1 = ex2(x1);
        }
        const int a = (c & 1) * 2;
        fadd2(sum4[a], sum4[a + 1], sum4[a], sum4[a + 1], p0, p1);
        pk[c] = pack_bf16x2(p0, p1);
      }
      B2F_TMEM_ST_X32(s_tmem, pk);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t * 2 + hw]);
      l = l * alpha + ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
      m = m_use;
    }
    float* xl = xch + 1024 + t * 256 + row_in_tile;
    xl[hw * 128] = l;
    named_bar_sync(bar_id, 64);
    l += xl[(hw ^ 1) * 128];
    mbar_wait(&o_done[t], 0);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const bool row_ok = q_row < p.Sq;
    __nv_bfloat16* out_row = p.out + ((long long)b * p.Sq + q_row) * p.ldo + (long long)h * DH;
#pragma unroll 1
    for (int c0 = hw * 64; c0 < hw * 64 + 64; c0 += 32) {
      uint32_t o[32];
      __syncwarp();
      B2F_TMEM_LD_X32(o_tmem + c0, o);
      tmem_wait_ld();
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
          v.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
          v.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
          v.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(out_row + c0 + g * 8) = v;
        }
      }
    }
    }  // t == 0 || t1_active
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


}  // namespace

static int attention_impl(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                          int64_t ldv, void* out, int64_t ldo, int B, int H, int Hkv, int Sq, int Skv,
                          int head_dim, float scale, int causal, const void* bias, int64_t bias_h_stride,
                          int64_t bias_row_stride, cudaStream_t stream) {
  if (!device_info().ok) return B2F_ERR_NODEVICE;
  if (!q || !k || !v || !out || B <= 0 || H <= 0 || Hkv <= 0 || Sq <= 0 || Skv <= 0)
    return B2F_ERR_INVALID;
  if (head_dim != DH) return B2F_ERR_UNSUPPORTED;
  if (H % Hkv) return B2F_ERR_INVALID;
  if (causal && Sq != Skv) return B2F_ERR_UNSUPPORTED;
  if ((ldq & 7) || (ldk & 7) || (ldv & 7) || (ldo & 7)) return B2F_ERR_ALIGN;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) |
       reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out)) & 15)
    return B2F_ERR_ALIGN;
  // kernel variant: compile-time (POLY, TURNS); B2F_ATTN_VARIANT selects at run time for tuning
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnParams);
  static KernelFn kernel = nullptr;
  static bool single_tile = false;
  static bool v4 = false;
  static bool v6 = false;
  if (!kernel) {
    const char* v = getenv("B2F_ATTN_VARIANT");
    const int variant = v ? atoi(v) : B2F_ATTN_DEFAULT_VARIANT;
    switch (variant) {
      case 1: case 50: case 51: case 52:                    // (50-52: what the CTA-pair kernel does not cover)
        kernel = attn_fwd_kernel<4, false>; break;          // 25 % of the exponentials on the FMA pipe
      case 2: kernel = attn_fwd_kernel<2, false>; break;   // 50 %
      case 3: kernel = attn_fwd_kernel<0, true>; break;
      case 4: kernel = attn_fwd_kernel<4, true>; break;
      case 5: kernel = attn_fwd_kernel<3, false>; break;   // 33 %
      case 6: kernel = attn_fwd_kernel<8, false>; break;   // 12.5 %
      case 21: kernel = attn_fwd_kernel<0, false, 1>; break;  // ablations for timing analysis only

      case 10: kernel = attn_fwd_kernel_v2<0>; break;
      case 11: kernel = attn_fwd_kernel_v2<4>; break;
      case 12: kernel = attn_fwd_kernel_v2<2>; break;
      case 60: kernel = attn_fwd_kernel_v6<0>; v6 = true; break;
      case 61: kernel = attn_fwd_kernel_v6<4>; v6 = true; break;
      case 62: kernel = attn_fwd_kernel_v6<2>; v6 = true; break;
      case 40: kernel = attn_fwd_kernel_v4<0>; single_tile = true; v4 = true; break;
      case 41: kernel = attn_fwd_kernel_v4<4>; single_tile = true; v4 = true; break;
      case 42: kernel = attn_fwd_kernel_v4<2>; single_tile = true; v4 = true; break;
      case 30: kernel = attn_fwd_kernel_v3<0>; single_tile = true; break;
      case 31: kernel = attn_fwd_kernel_v3<4>; single_tile = true; break;
      case 32: kernel = attn_fwd_kernel_v3<3>; single_tile = true; break;
      default: kernel = attn_fwd_kernel<0, false>; break;
    }
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         v6 ? V6_SMEM : v4 ? V4_SMEM : single_tile ? V3_SMEM : ATTN_SMEM);
    if (e != cudaSuccess) return cuda_err(e, "attention smem attribute");
    e = cudaFuncSetAttribute(attn_fwd_kernel<0, false, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             ATTN_SMEM);
    if (e != cudaSuccess) return cuda_err(e, "attention smem attribute");
  }
  // CTA-pair kernel (B2F_ATTN_VARIANT 50 / 51 / 52): non-causal, no bias, at least one full 512-row pair
  static const int pair_poly = [] {
    const char* v = getenv("B2F_ATTN_VARIANT");
    const int variant = v ? atoi(v) : B2F_ATTN_DEFAULT_VARIANT;
    return variant == 50 ? 0 : variant == 51 ? 4 : variant == 52 ? 2 : -1;
  }();
  if (pair_poly >= 0 && !causal && !bias && Sq >= 4 * BQ) {
    typedef void (*PairFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnParams);
    static PairFn pk = nullptr;
    if (!pk) {
      pk = pair_poly == 0 ? attn_fwd_kernel_2cta<0> : pair_poly == 4 ? attn_fwd_kernel_2cta<4> : attn_fwd_kernel_2cta<2>;
      cudaError_t e = cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM);
      if (e != cudaSuccess) return cuda_err(e, "attention pair smem attribute");
    }
    CUtensorMap tQ, tK, tV;
    int r2 = make_tmap_3d_rows(&tQ, q, (uint64_t)H * DH, Sq, B, ldq, (uint64_t)Sq * ldq);
    if (r2) return r2;
    r2 = make_tmap_3d_rows(&tK, k, (uint64_t)Hkv * DH, Skv, B, ldk, (uint64_t)Skv * ldk, 64);   // 64-row K halves
    if (r2) return r2;
    r2 = make_tmap_3d_rows(&tV, v, (uint64_t)Hkv * DH, Skv, B, ldv, (uint64_t)Skv * ldv);
    if (r2) return r2;
    AttnParams pp{};
    pp.B = B;
    pp.H = H;
    pp.Hkv = Hkv;
    pp.Sq = Sq;
    pp.Skv = Skv;
    pp.scale_log2 = scale * 1.4426950408889634f;
    pp.causal = 0;
    pp.out = static_cast<__nv_bfloat16*>(out);
    pp.ldo = ldo;
    dim3 grid_p(2 * ((Sq + 4 * BQ - 1) / (4 * BQ)), H, B);
    prof_begin(KC_ATTN, stream);
    pk<<<grid_p, ATTN_THREADS, P2_SMEM, stream>>>(tQ, tK, tV, pp);
    prof_end(KC_ATTN, stream, 4.0 * B * H * (double)Sq * Skv * DH, 2.0 * DH * B * (2.0 * H * Sq + 2.0 * Hkv * Skv));
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    B2F_CHECK_LAUNCH("attn_fwd_kernel_2cta");
    return B2F_OK;
  }
  CUtensorMap tmQ, tmK, tmV;
  int rc = make_tmap_3d_rows(&tmQ, q, (uint64_t)H * DH, Sq, B, ldq, (uint64_t)Sq * ldq);
  if (rc) return rc;
  rc = make_tmap_3d_rows(&tmK, k, (uint64_t)Hkv * DH, Skv, B, ldk, (uint64_t)Skv * ldk);
  if (rc) return rc;
  rc = make_tmap_3d_rows(&tmV, v, (uint64_t)Hkv * DH, Skv, B, ldv, (uint64_t)Skv * ldv);
  if (rc) return rc;
  AttnParams p{};
  p.B = B;
  p.H = H;
  p.Hkv = Hkv;
  p.Sq = Sq;
  p.Skv = Skv;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  if (bias) {
    // score = scale * q.k + bias, evaluated before the base-2 conversion
    p.bias = static_cast<const __nv_bfloat16*>(bias);
    p.bias_h_stride = bias_h_stride;
    p.bias_row_stride = bias_row_stride;
    p.bias_scale = scale;
    p.scale_log2 = 1.4426950408889634f;
    dim3 grid_b((Sq + 2 * BQ - 1) / (2 * BQ), H, B);
    prof_begin(KC_ATTN, stream);
    attn_fwd_kernel<0, false, 0, true><<<grid_b, ATTN_THREADS, ATTN_SMEM, stream>>>(tmQ, tmK, tmV, p);
    prof_end(KC_ATTN, stream, (causal ? 2.0 : 4.0) * B * H * (double)Sq * Skv * DH,
             2.0 * DH * B * (2.0 * H * Sq + 2.0 * Hkv * Skv) + 2.0 * H * (double)Sq * Skv);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    B2F_CHECK_LAUNCH("attn_fwd_kernel<bias>");
    return B2F_OK;
  }
  dim3 grid(single_tile ? (Sq + BQ - 1) / BQ : (Sq + 2 * BQ - 1) / (2 * BQ), H, B);
  prof_begin(KC_ATTN, stream);
  kernel<<<grid, v6 ? V6_THREADS : v4 ? V4_THREADS : single_tile ? V3_THREADS : ATTN_THREADS,
           v6 ? V6_SMEM : v4 ? V4_SMEM : single_tile ? V3_SMEM : ATTN_SMEM, stream>>>(tmQ, tmK, tmV, p);
  prof_end(KC_ATTN, stream, (causal ? 2.0 : 4.0) * B * H * (double)Sq * Skv * DH,
           2.0 * DH * B * (2.0 * H * Sq + 2.0 * Hkv * Skv));
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  B2F_CHECK_LAUNCH("attn_fwd_kernel");
  return B2F_OK;
}

int attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                  int64_t ldv, void* out, int64_t ldo, int B, int H, int Hkv, int Sq, int Skv,
                  int head_dim, float scale, int causal, cudaStream_t stream) {
  return attention_impl(q, ldq, k, ldk, v, ldv, out, ldo, B, H, Hkv, Sq, Skv, head_dim, scale, causal,
                        nullptr, 0, 0, stream);
}

int attention_bias_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                       int64_t ldv, void* out, int64_t ldo, int B, int H, int Hkv, int Sq, int Skv,
                       int head_dim, float scale, int causal, const void* bias, int64_t bias_h_stride,
                       int64_t bias_row_stride, cudaStream_t stream) {
  if (!bias || bias_row_stride < Skv || bias_h_stride < 0) return B2F_ERR_INVALID;
  return attention_impl(q, ldq, k, ldk, v, ldv, out, ldo, B, H, Hkv, Sq, Skv, head_dim, scale, causal,
                        bias, bias_h_stride, bias_row_stride, stream);
}

}  // namespace b2f
