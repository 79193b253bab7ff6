// FLUX-Kontext MMDiT forward composed from the libb2f kernels (no torch, no allocation, no host
// sync inside b2f_flux_forward — CUDA-graph capturable).
//
// Restates diffusers 0.32.2 FluxTransformer2DModel.forward as the sequence
//   x_embedder / context_embedder -> 19 x double block -> 38 x single block -> norm_out -> proj_out
// over ONE joint activation buffer h[B, S_txt + S_img, d] (text rows first), so the torch.cat calls
// of the reference ([txt;img] Q/K/V, [c;x] before the single blocks, [attn|mlp] before proj_out)
// become pointer offsets:
//   * Q/K/V of both streams are written by the QKV GEMMs straight into qkv[B, S, 3d];
//   * attention reads them through strided TMA maps and writes into cat[B, S, 5d][:, :, 0:d];
//   * the MLP up-projection (GELU fused) writes into cat[:, :, d:5d]; the single-block proj_out
//     GEMM reads cat with K = 5d.
// Reference call sites of this path: univa/utils/flux_pipeline.py:1067-1077,
// univa/models/modeling_univa_denoise_tower.py:103-110.
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "flux_ctx.h"
#include "host_common.h"

namespace b2f {

int gemm_bf16(const void* A, int64_t lda, int64_t a_bs, const void* W, int64_t ldw,
              const void* bias, void* out, int64_t ldc, int64_t out_bs, int batch, int M, int N,
              int K, int epilogue, const void* resid, int64_t ldr, int64_t resid_bs, const void* gate,
              int64_t gate_ld, cudaStream_t stream);
int gemm_qkv_norm_rope(const void* A, int64_t lda, int64_t a_bs, const void* W, int64_t ldw,
                       const void* bias, void* out, int64_t ldc, int64_t out_bs, int batch, int M,
                       int d_model, int K, const void* nw_q, const void* nw_k, const float* cos,
                       const float* sin, int rope_row0, float eps, int n_extra, void* out_extra,
                       int64_t ld_extra, int64_t bs_extra, int epi_extra, cudaStream_t stream);
int attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                  int64_t ldv, void* out, int64_t ldo, int B, int H, int Hkv, int Sq, int Skv,
                  int head_dim, float scale, int causal, cudaStream_t stream);
int ln_modulate(const void* x, int64_t ldx, int64_t x_batch_stride, const void* scale,
                const void* shift, int64_t mod_ld, void* out, int64_t ldo, int64_t out_batch_stride,
                int batch, int rows, int D, float eps, int split_row, const void* scale_b,
                const void* shift_b, cudaStream_t stream);
int rmsnorm_rope(void* q, void* k, int64_t ld, int64_t batch_stride, const void* wq_a,
                 const void* wk_a, const void* wq_b, const void* wk_b, const float* cos,
                 const float* sin, int batch, int S, int H, int head_dim, int n_a, float eps,
                 cudaStream_t stream);
int temb_sinusoid(const float* t, void* out, int rows, cudaStream_t stream);
int temb_combine(const void* t, const void* g, const void* txt, void* temb, void* silu_temb,
                 int64_t n, cudaStream_t stream);

static int64_t mod_width_of(const b2f_flux_cfg& c) {
  const int64_t d = (int64_t)c.num_heads * c.head_dim;
  return (int64_t)c.num_double * 12 * d + (int64_t)c.num_single * 3 * d + 2 * d;
}

static int expect(FluxCtx* c, const std::string& key, int64_t numel, const bf16_t** dst) {
  auto it = c->bound.find(key);
  if (it == c->bound.end()) {
    fprintf(stderr, "[b2f] flux_finalize: weight '%s' was never bound\n", key.c_str());
    return B2F_ERR_INVALID;
  }
  if (it->second.second != numel) {
    fprintf(stderr, "[b2f] flux_finalize: weight '%s' has %lld elements, expected %lld\n",
            key.c_str(), (long long)it->second.second, (long long)numel);
    return B2F_ERR_INVALID;
  }
  *dst = static_cast<const bf16_t*>(it->second.first);
  return B2F_OK;
}
static int expect_lin(FluxCtx* c, const std::string& name, int64_t out_f, int64_t in_f, Lin* l) {
  int rc = expect(c, name + ".weight", out_f * in_f, &l->w);
  if (rc) return rc;
  return expect(c, name + ".bias", out_f, &l->b);
}

}  // namespace b2f

using namespace b2f;

extern "C" {

int b2f_flux_create(b2f_flux** out, const b2f_flux_cfg* cfg) {
  if (!out || !cfg) return B2F_ERR_INVALID;
  if (cfg->head_dim != 128) return B2F_ERR_UNSUPPORTED;
  const int64_t d = (int64_t)cfg->num_heads * cfg->head_dim;
  if (d % 256 || cfg->in_channels % 8 || cfg->out_channels % 8 || cfg->joint_dim % 8 ||
      cfg->pooled_dim % 8 || cfg->num_double < 0 || cfg->num_single < 0 || cfg->mlp_ratio != 4)
    return B2F_ERR_UNSUPPORTED;
  FluxCtx* c = new (std::nothrow) FluxCtx();
  if (!c) return B2F_ERR_INVALID;
  c->cfg = *cfg;
  c->d = (int)d;
  c->mod_width = mod_width_of(*cfg);
  *out = reinterpret_cast<b2f_flux*>(c);
  return B2F_OK;
}

void b2f_flux_destroy(b2f_flux* h) { delete reinterpret_cast<FluxCtx*>(h); }

int b2f_flux_bind_weight(b2f_flux* h, const char* key, const void* dptr, int64_t numel) {
  FluxCtx* c = reinterpret_cast<FluxCtx*>(h);
  if (!c || !key || !dptr || numel <= 0) return B2F_ERR_INVALID;
  if (reinterpret_cast<uintptr_t>(dptr) & 15) return B2F_ERR_ALIGN;
  c->bound[key] = {dptr, numel};
  c->finalized = false;
  return B2F_OK;
}

int64_t b2f_flux_mod_width(const b2f_flux* h) {
  const FluxCtx* c = reinterpret_cast<const FluxCtx*>(h);
  return c ? c->mod_width : 0;
}

int b2f_flux_finalize(b2f_flux* h) {
  FluxCtx* c = reinterpret_cast<FluxCtx*>(h);
  if (!c) return B2F_ERR_INVALID;
  const int64_t d = c->d, d4 = 4 * d;
  const b2f_flux_cfg& g = c->cfg;
  int rc;
#define EL(name, o, i, dst) \
  if ((rc = expect_lin(c, name, o, i, dst)) != 0) return rc
#define EW(name, n, dst) \
  if ((rc = expect(c, name, n, dst)) != 0) return rc
  EL("x_embedder", d, g.in_channels, &c->x_embedder);
  EL("context_embedder", d, g.joint_dim, &c->context_embedder);
  EL("time_text_embed.timestep_embedder.linear_1", d, 256, &c->t1);
  EL("time_text_embed.timestep_embedder.linear_2", d, d, &c->t2);
  if (g.guidance_embeds) {
    EL("time_text_embed.guidance_embedder.linear_1", d, 256, &c->g1);
    EL("time_text_embed.guidance_embedder.linear_2", d, d, &c->g2);
  }
  EL("time_text_embed.text_embedder.linear_1", d, g.pooled_dim, &c->p1);
  EL("time_text_embed.text_embedder.linear_2", d, d, &c->p2);
  EL("adaln", c->mod_width, d, &c->adaln);
  EL("proj_out", g.out_channels, d, &c->proj_out);
  c->dbl.assign(g.num_double, DoubleW());
  for (int i = 0; i < g.num_double; ++i) {
    const std::string p = "transformer_blocks." + std::to_string(i) + ".";
    DoubleW& w = c->dbl[i];
    EL(p + "attn.qkv", 3 * d, d, &w.qkv);
    EL(p + "attn.add_qkv", 3 * d, d, &w.add_qkv);
    EL(p + "attn.to_out.0", d, d, &w.to_out);
    EL(p + "attn.to_add_out", d, d, &w.to_add_out);
    EL(p + "ff.net.0.proj", d4, d, &w.ff1);
    EL(p + "ff.net.2", d, d4, &w.ff2);
    EL(p + "ff_context.net.0.proj", d4, d, &w.ffc1);
    EL(p + "ff_context.net.2", d, d4, &w.ffc2);
    EW(p + "attn.norm_q.weight", g.head_dim, &w.norm_q);
    EW(p + "attn.norm_k.weight", g.head_dim, &w.norm_k);
    EW(p + "attn.norm_added_q.weight", g.head_dim, &w.norm_added_q);
    EW(p + "attn.norm_added_k.weight", g.head_dim, &w.norm_added_k);
  }
  c->sgl.assign(g.num_single, SingleW());
  for (int i = 0; i < g.num_single; ++i) {
    const std::string p = "single_transformer_blocks." + std::to_string(i) + ".";
    SingleW& w = c->sgl[i];
    EL(p + "qkv_mlp", 7 * d, d, &w.qkv_mlp);
    EL(p + "proj_out", d, 5 * d, &w.proj_out);
    EW(p + "attn.norm_q.weight", g.head_dim, &w.norm_q);
    EW(p + "attn.norm_k.weight", g.head_dim, &w.norm_k);
  }
#undef EL
#undef EW
  c->finalized = true;
  return B2F_OK;
}

int b2f_flux_set_rope(b2f_flux* h, const float* cos, const float* sin, int S) {
  FluxCtx* c = reinterpret_cast<FluxCtx*>(h);
  if (!c || !cos || !sin || S <= 0) return B2F_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(cos) | reinterpret_cast<uintptr_t>(sin)) & 15) return B2F_ERR_ALIGN;
  c->rope_cos = cos;
  c->rope_sin = sin;
  c->rope_S = S;
  return B2F_OK;
}

size_t b2f_flux_workspace_bytes(const b2f_flux* h, int B, int S_img, int S_txt) {
  const FluxCtx* c = reinterpret_cast<const FluxCtx*>(h);
  if (!c || B <= 0 || S_img <= 0 || S_txt < 0) return 0;
  const size_t S = (size_t)S_img + S_txt;
  // h[d] + xn[d] + qkv[3d] + cat[5d] per token, bf16
  return (size_t)B * S * (size_t)c->d * 10 * 2 + 1024;
}

size_t b2f_flux_temb_workspace_bytes(const b2f_flux* h, int rows) {
  const FluxCtx* c = reinterpret_cast<const FluxCtx*>(h);
  if (!c || rows <= 0) return 0;
  // 2 sinusoid tables [rows,256] + 6 activations [rows,d]
  return ((size_t)rows * 256 * 2 + (size_t)rows * c->d * 6) * 2 + 1024;
}

int b2f_flux_temb(b2f_flux* h, const float* timestep, const float* guidance, const void* pooled,
                  int64_t pooled_ld, int rows, void* temb, void* silu_temb, void* ws,
                  size_t ws_bytes, b2f_stream_t stream_) {
  FluxCtx* c = reinterpret_cast<FluxCtx*>(h);
  if (!c || !c->finalized || !timestep || !pooled || !temb || !silu_temb || rows <= 0 || !ws)
    return B2F_ERR_INVALID;
  if (c->cfg.guidance_embeds && !guidance) return B2F_ERR_INVALID;
  if (ws_bytes < b2f_flux_temb_workspace_bytes(h, rows)) return B2F_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int d = c->d;
  bf16_t* w = reinterpret_cast<bf16_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
  bf16_t* sin_t = w;
  bf16_t* sin_g = sin_t + (size_t)rows * 256;
  bf16_t* a1 = sin_g + (size_t)rows * 256;  // hidden
  bf16_t* et = a1 + (size_t)rows * d;
  bf16_t* eg = et + (size_t)rows * d;
  bf16_t* ep = eg + (size_t)rows * d;
  int rc;
  if ((rc = temb_sinusoid(timestep, sin_t, rows, st))) return rc;
  if ((rc = gemm_bf16(sin_t, 256, 0, c->t1.w, 256, c->t1.b, a1, d, 0, 1, rows, d, 256, B2F_EPI_SILU,
                      nullptr, 0, 0, nullptr, 0, st)))
    return rc;
  if ((rc = gemm_bf16(a1, d, 0, c->t2.w, d, c->t2.b, et, d, 0, 1, rows, d, d, B2F_EPI_BIAS, nullptr,
                      0, 0, nullptr, 0, st)))
    return rc;
  const bf16_t* eg_ptr = nullptr;
  if (c->cfg.guidance_embeds) {
    if ((rc = temb_sinusoid(guidance, sin_g, rows, st))) return rc;
    if ((rc = gemm_bf16(sin_g, 256, 0, c->g1.w, 256, c->g1.b, a1, d, 0, 1, rows, d, 256,
                        B2F_EPI_SILU, nullptr, 0, 0, nullptr, 0, st)))
      return rc;
    if ((rc = gemm_bf16(a1, d, 0, c->g2.w, d, c->g2.b, eg, d, 0, 1, rows, d, d, B2F_EPI_BIAS,
                        nullptr, 0, 0, nullptr, 0, st)))
      return rc;
    eg_ptr = eg;
  }
  if ((rc = gemm_bf16(pooled, pooled_ld, 0, c->p1.w, c->cfg.pooled_dim, c->p1.b, a1, d, 0, 1, rows, d,
                      c->cfg.pooled_dim, B2F_EPI_SILU, nullptr, 0, 0, nullptr, 0, st)))
    return rc;
  if ((rc = gemm_bf16(a1, d, 0, c->p2.w, d, c->p2.b, ep, d, 0, 1, rows, d, d, B2F_EPI_BIAS, nullptr,
                      0, 0, nullptr, 0, st)))
    return rc;
  return temb_combine(et, eg_ptr, ep, temb, silu_temb, (int64_t)rows * d, st);
}

int b2f_flux_modulation(b2f_flux* h, const void* silu_temb, int rows, void* mod, b2f_stream_t stream_) {
  FluxCtx* c = reinterpret_cast<FluxCtx*>(h);
  if (!c || !c->finalized || !silu_temb || !mod || rows <= 0) return B2F_ERR_INVALID;
  return gemm_bf16(silu_temb, c->d, 0, c->adaln.w, c->d, c->adaln.b, mod, c->mod_width, 0, 1, rows,
                   (int)c->mod_width, c->d, B2F_EPI_BIAS, nullptr, 0, 0, nullptr, 0,
                   static_cast<cudaStream_t>(stream_));
}

int b2f_flux_forward(b2f_flux* h, const void* hidden, const void* enc, const void* mod,
                     int64_t mod_ld, void* out, int B, int S_img, int S_txt, int n_out_rows,
                     void* ws, size_t ws_bytes, int first_block, int last_block,
                     b2f_stream_t stream_) {
  FluxCtx* c = reinterpret_cast<FluxCtx*>(h);
  if (!c || !c->finalized) return B2F_ERR_INVALID;
  if (!hidden || !enc || !mod || !out || !ws || B <= 0 || S_img <= 0 || S_txt <= 0)
    return B2F_ERR_INVALID;
  if (n_out_rows <= 0 || n_out_rows > S_img) return B2F_ERR_INVALID;
  const int S = S_img + S_txt;
  if (!c->rope_cos || c->rope_S != S) {
    fprintf(stderr, "[b2f] flux_forward: RoPE tables not set for S=%d (have %d)\n", S, c->rope_S);
    return B2F_ERR_INVALID;
  }
  if (ws_bytes < b2f_flux_workspace_bytes(h, B, S_img, S_txt)) return B2F_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const b2f_flux_cfg& g = c->cfg;
  const int64_t d = c->d;
  const int H = g.num_heads;
  const float eps = 1e-6f;
  const float scale = 1.0f / sqrtf((float)g.head_dim);

  bf16_t* base = reinterpret_cast<bf16_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
  const int64_t BS = (int64_t)B * S;
  bf16_t* hb = base;               // [B,S,d]
  bf16_t* xn = hb + BS * d;        // [B,S,d]
  bf16_t* qkv = xn + BS * d;       // [B,S,3d]
  bf16_t* cat = qkv + BS * 3 * d;  // [B,S,5d]
  const int64_t h_bs = (int64_t)S * d, qkv_bs = (int64_t)S * 3 * d, cat_bs = (int64_t)S * 5 * d;
  bf16_t* h_img = hb + (int64_t)S_txt * d;
  bf16_t* h_txt = hb;
  bf16_t* xn_img = xn + (int64_t)S_txt * d;
  bf16_t* xn_txt = xn;
  bf16_t* qkv_img = qkv + (int64_t)S_txt * 3 * d;
  bf16_t* qkv_txt = qkv;
  bf16_t* cat_img = cat + (int64_t)S_txt * 5 * d;
  bf16_t* cat_txt = cat;
  const bf16_t* modp = static_cast<const bf16_t*>(mod);
  int rc;
#define RUN(expr) \
  if ((rc = (expr)) != 0) return rc
  const int total_blocks = g.num_double + g.num_single;
  if (first_block < 0) first_block = 0;
  if (last_block < 0 || last_block > total_blocks) last_block = total_blocks;

  if (first_block == 0) {
    // x = x_embedder(hidden) -> h[:, S_txt:],  c = context_embedder(enc) -> h[:, :S_txt]
    RUN(gemm_bf16(hidden, g.in_channels, (int64_t)S_img * g.in_channels, c->x_embedder.w,
                  g.in_channels, c->x_embedder.b, h_img, d, h_bs, B, S_img, (int)d, g.in_channels,
                  B2F_EPI_BIAS, nullptr, 0, 0, nullptr, 0, st));
    RUN(gemm_bf16(enc, g.joint_dim, (int64_t)S_txt * g.joint_dim, c->context_embedder.w, g.joint_dim,
                  c->context_embedder.b, h_txt, d, h_bs, B, S_txt, (int)d, g.joint_dim, B2F_EPI_BIAS,
                  nullptr, 0, 0, nullptr, 0, st));
  }

  for (int blk = first_block; blk < last_block; ++blk) {
    if (blk < g.num_double) {
      const DoubleW& w = c->dbl[blk];
      // mod columns: [img: shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp | txt: same]
      const bf16_t* mi = modp + (int64_t)blk * 12 * d;
      const bf16_t* mt = mi + 6 * d;
      // both streams in one launch over the joint buffer: text rows use the context modulation
      RUN(ln_modulate(hb, d, h_bs, mt + d, mt, mod_ld, xn, d, h_bs, B, S, (int)d, eps, S_txt, mi + d, mi, st));
      // QKV projections with per-head RMSNorm + RoPE fused into the GEMM epilogue; both streams write
      // straight into the joint [txt; img] qkv buffer
      RUN(gemm_qkv_norm_rope(xn_img, d, h_bs, w.qkv.w, d, w.qkv.b, qkv_img, 3 * d, qkv_bs, B, S_img, (int)d,
                             (int)d, w.norm_q, w.norm_k, c->rope_cos, c->rope_sin, S_txt, eps, 0, nullptr, 0, 0, 0, st));
      RUN(gemm_qkv_norm_rope(xn_txt, d, h_bs, w.add_qkv.w, d, w.add_qkv.b, qkv_txt, 3 * d, qkv_bs, B, S_txt,
                             (int)d, (int)d, w.norm_added_q, w.norm_added_k, c->rope_cos, c->rope_sin, 0, eps, 0,
                             nullptr, 0, 0, 0, st));
      RUN(attention_fwd(qkv, 3 * d, qkv + d, 3 * d, qkv + 2 * d, 3 * d, cat, 5 * d, B, H, H, S, S,
                        g.head_dim, scale, 0, st));
      RUN(gemm_bf16(cat_img, 5 * d, cat_bs, w.to_out.w, d, w.to_out.b, h_img, d, h_bs, B, S_img,
                    (int)d, (int)d, B2F_EPI_GATE_RESID, h_img, d, h_bs, mi + 2 * d, mod_ld, st));
      RUN(gemm_bf16(cat_txt, 5 * d, cat_bs, w.to_add_out.w, d, w.to_add_out.b, h_txt, d, h_bs, B,
                    S_txt, (int)d, (int)d, B2F_EPI_GATE_RESID, h_txt, d, h_bs, mt + 2 * d, mod_ld, st));
      RUN(ln_modulate(hb, d, h_bs, mt + 4 * d, mt + 3 * d, mod_ld, xn, d, h_bs, B, S, (int)d, eps, S_txt,
                      mi + 4 * d, mi + 3 * d, st));
      RUN(gemm_bf16(xn_img, d, h_bs, w.ff1.w, d, w.ff1.b, cat_img + d, 5 * d, cat_bs, B, S_img,
                    (int)(4 * d), (int)d, B2F_EPI_GELU_TANH, nullptr, 0, 0, nullptr, 0, st));
      RUN(gemm_bf16(xn_txt, d, h_bs, w.ffc1.w, d, w.ffc1.b, cat_txt + d, 5 * d, cat_bs, B, S_txt,
                    (int)(4 * d), (int)d, B2F_EPI_GELU_TANH, nullptr, 0, 0, nullptr, 0, st));
      RUN(gemm_bf16(cat_img + d, 5 * d, cat_bs, w.ff2.w, 4 * d, w.ff2.b, h_img, d, h_bs, B, S_img,
                    (int)d, (int)(4 * d), B2F_EPI_GATE_RESID, h_img, d, h_bs, mi + 5 * d, mod_ld, st));
      RUN(gemm_bf16(cat_txt + d, 5 * d, cat_bs, w.ffc2.w, 4 * d, w.ffc2.b, h_txt, d, h_bs, B, S_txt,
                    (int)d, (int)(4 * d), B2F_EPI_GATE_RESID, h_txt, d, h_bs, mt + 5 * d, mod_ld, st));
    } else {
      const int si = blk - g.num_double;
      const SingleW& w = c->sgl[si];
      // mod columns: [shift, scale, gate]
      const bf16_t* ms = modp + (int64_t)g.num_double * 12 * d + (int64_t)si * 3 * d;
      RUN(ln_modulate(hb, d, h_bs, ms + d, ms, mod_ld, xn, d, h_bs, B, S, (int)d, eps, 0, nullptr, nullptr, st));
      // ONE launch for [to_q;to_k;to_v;proj_mlp] (N = 7d): Q/K get RMSNorm+RoPE, V passes through into
      // qkv, the MLP columns are GELU'd straight into cat[:, :, d:5d]
      RUN(gemm_qkv_norm_rope(xn, d, h_bs, w.qkv_mlp.w, d, w.qkv_mlp.b, qkv, 3 * d, qkv_bs, B, S, (int)d, (int)d,
                             w.norm_q, w.norm_k, c->rope_cos, c->rope_sin, 0, eps, (int)(4 * d), cat + d, 5 * d,
                             cat_bs, B2F_EPI_GELU_TANH, st));
      RUN(attention_fwd(qkv, 3 * d, qkv + d, 3 * d, qkv + 2 * d, 3 * d, cat, 5 * d, B, H, H, S, S,
                        g.head_dim, scale, 0, st));
      RUN(gemm_bf16(cat, 5 * d, cat_bs, w.proj_out.w, 5 * d, w.proj_out.b, hb, d, h_bs, B, S, (int)d,
                    (int)(5 * d), B2F_EPI_GATE_RESID, hb, d, h_bs, ms + 2 * d, mod_ld, st));
    }
  }

  if (last_block == total_blocks) {
    // norm_out (AdaLayerNormContinuous: chunk order scale, shift) + proj_out on the image rows
    const bf16_t* mo = modp + (int64_t)g.num_double * 12 * d + (int64_t)g.num_single * 3 * d;
    RUN(ln_modulate(h_img, d, h_bs, mo, mo + d, mod_ld, xn_img, d, h_bs, B, n_out_rows, (int)d, eps, 0, nullptr,
                    nullptr, st));
    RUN(gemm_bf16(xn_img, d, h_bs, c->proj_out.w, d, c->proj_out.b, out, g.out_channels,
                  (int64_t)n_out_rows * g.out_channels, B, n_out_rows, g.out_channels, (int)d,
                  B2F_EPI_BIAS, nullptr, 0, 0, nullptr, 0, st));
  }
#undef RUN
  return B2F_OK;
}

}  // extern "C"
