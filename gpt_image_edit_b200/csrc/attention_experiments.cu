// Attention kernel EXPERIMENTS (B2F_ATTN_VARIANT 10-12, 30-32, 40-42, 60-62): alternative structures for the
// two-tile kernel of attention.cu, all parity-green (tests/test_attention_gpu.py runs with any variant) and none
// faster on B200 — kept so that the measurements in DESIGN.md section 7 can be reproduced:
//   v2 (10-12)  two tiles, 64-column S sub-blocks with two S buffers per tile
//   v3 (30-32)  one Q tile per CTA, S double-buffered (QK(j+1) overlaps softmax(j)), one softmax warpgroup
//   v4 (40-42)  v3 with two warpgroups splitting the S columns
//   v6 (60-62)  two tiles x two column warpgroups (16 softmax warps)
#include "attention_common.cuh"

namespace b2f {
namespace attn {

namespace {

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

bool experimental_variant(int variant, Variant* out) {
  switch (variant) {
    case 10: *out = {attn_fwd_kernel_v2<0>, ATTN_THREADS, ATTN_SMEM, false}; return true;
    case 11: *out = {attn_fwd_kernel_v2<4>, ATTN_THREADS, ATTN_SMEM, false}; return true;
    case 12: *out = {attn_fwd_kernel_v2<2>, ATTN_THREADS, ATTN_SMEM, false}; return true;
    case 30: *out = {attn_fwd_kernel_v3<0>, V3_THREADS, V3_SMEM, true}; return true;
    case 31: *out = {attn_fwd_kernel_v3<4>, V3_THREADS, V3_SMEM, true}; return true;
    case 32: *out = {attn_fwd_kernel_v3<3>, V3_THREADS, V3_SMEM, true}; return true;
    case 40: *out = {attn_fwd_kernel_v4<0>, V4_THREADS, V4_SMEM, true}; return true;
    case 41: *out = {attn_fwd_kernel_v4<4>, V4_THREADS, V4_SMEM, true}; return true;
    case 42: *out = {attn_fwd_kernel_v4<2>, V4_THREADS, V4_SMEM, true}; return true;
    case 60: *out = {attn_fwd_kernel_v6<0>, V6_THREADS, V6_SMEM, false}; return true;
    case 61: *out = {attn_fwd_kernel_v6<4>, V6_THREADS, V6_SMEM, false}; return true;
    case 62: *out = {attn_fwd_kernel_v6<2>, V6_THREADS, V6_SMEM, false}; return true;
    default: return false;
  }
}

}  // namespace attn
}  // namespace b2f
