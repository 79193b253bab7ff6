// FLUX VAE (diffusers AutoencoderKL) encode / decode composed from libb2f kernels.
// Activations are NHWC bf16 inside; the ABI takes and returns the NCHW tensors the reference passes
// (univa/utils/flux_pipeline.py:600-613 encode -> latent_dist, :1127-1129 decode).
//   3x3 convs            tcgen05 implicit GEMM (conv.cu), residual add fused into conv2's epilogue
//   1x1 shortcut convs   plain tcgen05 GEMM over [pixels, C]
//   GroupNorm+SiLU       two HBM-bound passes (vae_kernels.cu)
//   mid-block attention  single head, dh = C: QK^T and PV as tcgen05 GEMMs around a row softmax
#include <map>
#include <string>
#include <vector>

#include "host_common.h"

namespace b2f {

int gemm_bf16(const void* A, int64_t lda, int64_t a_bs, const void* W, int64_t ldw,
              const void* bias, void* out, int64_t ldc, int64_t out_bs, int batch, int M, int N,
              int K, int epilogue, const void* resid, int64_t ldr, int64_t resid_bs, const void* gate,
              int64_t gate_ld, cudaStream_t stream);
int conv3x3(const void* in, const void* w, const void* bias, void* out, const void* resid, int N,
            int Hin, int Win, int Cin, int Cout, int stride, int out_nchw, cudaStream_t stream);
int groupnorm_silu(const void* x, const void* gamma, const void* beta, void* y, double* stats_ws,
                   int N, long long P, int C, float eps, int silu, cudaStream_t stream);
int upsample2x(const void* in, void* out, int N, int H, int W, int C, cudaStream_t stream);
int nchw_to_nhwc_pad(const void* in, int in_is_f32, void* out, int N, int C, int H, int W, int Cpad,
                     cudaStream_t stream);
int softmax_rows(void* s, int64_t ld, int rows, int L, float scale, cudaStream_t stream);
int transpose_bf16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int R, int Cc,
                   cudaStream_t stream);

typedef uint16_t bf16_t;

struct VaeCtx {
  b2f_vae_cfg cfg;
  std::map<std::string, std::pair<const void*, int64_t>> bound;
  bool finalized = false;
  int max_c = 0;
};

struct VaeRun {
  VaeCtx* c;
  cudaStream_t st;
  bf16_t *X, *A, *B;   // three activation buffers (NHWC)
  bf16_t* attn_ws;     // mid-attention scratch
  double* stats;       // [N,32,2]
  int N;
  int rc = 0;

  const bf16_t* w(const std::string& k) {
    auto it = c->bound.find(k);
    if (it == c->bound.end()) {
      if (!rc) fprintf(stderr, "[b2f] vae: weight '%s' not bound\n", k.c_str());
      rc = B2F_ERR_INVALID;
      return nullptr;
    }
    return static_cast<const bf16_t*>(it->second.first);
  }
  void run(int r) {
    if (!rc && r) rc = r;
  }
  void gn(const std::string& name, const bf16_t* x, bf16_t* y, long long P, int C, int silu) {
    if (rc) return;
    run(groupnorm_silu(x, w(name + ".weight"), w(name + ".bias"), y, stats, N, P, C, 1e-6f, silu, st));
  }
  void conv(const std::string& name, const bf16_t* in, bf16_t* out, const bf16_t* resid, int H, int W,
            int Cin, int Cout, int stride = 1, int nchw = 0) {
    if (rc) return;
    run(conv3x3(in, w(name + ".weight"), w(name + ".bias"), out, resid, N, H, W, Cin, Cout, stride, nchw, st));
  }
  // ResnetBlock2D in place on X: X[N,H,W,Cin] -> X[N,H,W,Cout]
  void resnet(const std::string& name, int H, int W, int Cin, int Cout) {
    const long long P = (long long)H * W;
    gn(name + ".norm1", X, A, P, Cin, 1);
    conv(name + ".conv1", A, B, nullptr, H, W, Cin, Cout);
    gn(name + ".norm2", B, A, P, Cout, 1);
    if (Cin != Cout) {
      if (rc) return;
      run(gemm_bf16(X, Cin, 0, w(name + ".conv_shortcut.weight"), Cin, w(name + ".conv_shortcut.bias"), B,
                    Cout, 0, 1, (int)(N * P), Cout, Cin, B2F_EPI_BIAS, nullptr, 0, 0, nullptr, 0, st));
      conv(name + ".conv2", A, X, B, H, W, Cout, Cout);
    } else {
      conv(name + ".conv2", A, X, X, H, W, Cout, Cout);
    }
  }
  // Attention(heads=1, residual) over the H*W tokens of each image, in place on X[N,P,C]
  void attention(const std::string& name, long long P, int C) {
    gn(name + ".group_norm", X, A, P, C, 0);
    if (rc) return;
    const bf16_t* wqkv = w(name + ".qkv.weight");
    const bf16_t* bqkv = w(name + ".qkv.bias");
    const bf16_t* wo = w(name + ".to_out.0.weight");
    const bf16_t* bo = w(name + ".to_out.0.bias");
    if (rc) return;
    bf16_t* qkv = attn_ws;                          // [P, 3C]
    bf16_t* vT = qkv + P * 3 * C;                   // [C, P]
    bf16_t* S = vT + (long long)C * P;              // [P, P]
    const float scale = 1.0f / sqrtf((float)C);
    for (int n = 0; n < N && !rc; ++n) {
      bf16_t* xa = A + (long long)n * P * C;
      bf16_t* xx = X + (long long)n * P * C;
      run(gemm_bf16(xa, C, 0, wqkv, C, bqkv, qkv, 3 * C, 0, 1, (int)P, 3 * C, C, B2F_EPI_BIAS, nullptr, 0, 0,
                    nullptr, 0, st));
      run(gemm_bf16(qkv, 3 * C, 0, qkv + C, 3 * C, nullptr, S, P, 0, 1, (int)P, (int)P, C, B2F_EPI_BIAS,
                    nullptr, 0, 0, nullptr, 0, st));
      run(softmax_rows(S, P, (int)P, (int)P, scale, st));
      run(transpose_bf16(qkv + 2 * C, 3 * C, vT, P, (int)P, C, st));
      run(gemm_bf16(S, P, 0, vT, P, nullptr, xa, C, 0, 1, (int)P, C, (int)P, B2F_EPI_BIAS, nullptr, 0, 0,
                    nullptr, 0, st));
      run(gemm_bf16(xa, C, 0, wo, C, bo, xx, C, 0, 1, (int)P, C, C, B2F_EPI_RESID, xx, C, 0, nullptr, 0, st));
    }
  }
  void mid(const std::string& name, int H, int W, int C) {
    resnet(name + ".resnets.0", H, W, C, C);
    attention(name + ".attentions.0", (long long)H * W, C);
    resnet(name + ".resnets.1", H, W, C, C);
  }
  void swapXA() {
    bf16_t* t = X;
    X = A;
    A = t;
  }
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// bytes: 3 activation buffers of `act` elements + attention scratch + stats
static void ws_layout(const b2f_vae_cfg& g, int N, int H, int W, size_t* act_elems, size_t* attn_elems) {
  // largest NHWC activation: full resolution x block_out[0] channels, or (decoder) the upsampled
  // tensor entering the last upsample conv: full resolution x block_out[1]
  const size_t full = (size_t)N * H * W;
  size_t m = full * (size_t)(g.block_out[1] > g.block_out[0] ? g.block_out[1] : g.block_out[0]);
  if (m < full * 64) m = full * 64;
  *act_elems = m;
  const size_t P = (size_t)(H / 8) * (W / 8);
  const size_t C = g.block_out[3];
  *attn_elems = P * 3 * C + C * P + P * P;
}

}  // namespace b2f

using namespace b2f;

extern "C" {

int b2f_vae_create(b2f_vae** out, const b2f_vae_cfg* cfg) {
  if (!out || !cfg) return B2F_ERR_INVALID;
  for (int i = 0; i < 4; ++i)
    if (cfg->block_out[i] <= 0 || cfg->block_out[i] % 32) return B2F_ERR_UNSUPPORTED;
  if (cfg->block_out[0] % 64 && cfg->block_out[0] != 32) return B2F_ERR_UNSUPPORTED;
  if (cfg->latent_channels <= 0 || cfg->latent_channels > 64 || cfg->layers_per_block <= 0)
    return B2F_ERR_UNSUPPORTED;
  VaeCtx* c = new (std::nothrow) VaeCtx();
  if (!c) return B2F_ERR_INVALID;
  c->cfg = *cfg;
  *out = reinterpret_cast<b2f_vae*>(c);
  return B2F_OK;
}
void b2f_vae_destroy(b2f_vae* h) { delete reinterpret_cast<VaeCtx*>(h); }

int b2f_vae_bind_weight(b2f_vae* h, const char* key, const void* dptr, int64_t numel) {
  VaeCtx* c = reinterpret_cast<VaeCtx*>(h);
  if (!c || !key || !dptr || numel <= 0) return B2F_ERR_INVALID;
  if (reinterpret_cast<uintptr_t>(dptr) & 15) return B2F_ERR_ALIGN;
  c->bound[key] = {dptr, numel};
  return B2F_OK;
}

size_t b2f_vae_workspace_bytes(const b2f_vae* h, int N, int H, int W) {
  const VaeCtx* c = reinterpret_cast<const VaeCtx*>(h);
  if (!c || N <= 0 || H <= 0 || W <= 0) return 0;
  size_t act, attn;
  ws_layout(c->cfg, N, H, W, &act, &attn);
  return 3 * align_up(act * 2, 256) + align_up(attn * 2, 256) + align_up(sizeof(double) * 64 * N, 256) + 1024;
}

static int vae_setup(VaeCtx* c, VaeRun* r, int N, int H, int W, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (ws_bytes < b2f_vae_workspace_bytes(reinterpret_cast<b2f_vae*>(c), N, H, W)) return B2F_ERR_WORKSPACE;
  size_t act, attn;
  ws_layout(c->cfg, N, H, W, &act, &attn);
  uint8_t* p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
  const size_t ab = align_up(act * 2, 256);
  r->c = c;
  r->st = st;
  r->N = N;
  r->X = reinterpret_cast<bf16_t*>(p);
  r->A = reinterpret_cast<bf16_t*>(p + ab);
  r->B = reinterpret_cast<bf16_t*>(p + 2 * ab);
  r->attn_ws = reinterpret_cast<bf16_t*>(p + 3 * ab);
  r->stats = reinterpret_cast<double*>(p + 3 * ab + align_up(attn * 2, 256));
  return B2F_OK;
}

int b2f_vae_encode(b2f_vae* h, const void* image_nchw, int image_is_f32, int N, int H, int W,
                   void* moments_nchw, void* ws, size_t ws_bytes, b2f_stream_t stream_) {
  VaeCtx* c = reinterpret_cast<VaeCtx*>(h);
  if (!c || !image_nchw || !moments_nchw || !ws || N <= 0) return B2F_ERR_INVALID;
  if (H % 8 || W % 8 || H <= 0 || W <= 0) return B2F_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  VaeRun r;
  int rc = vae_setup(c, &r, N, H, W, ws, ws_bytes, st);
  if (rc) return rc;
  const b2f_vae_cfg& g = c->cfg;
  r.run(nchw_to_nhwc_pad(image_nchw, image_is_f32, r.A, N, g.in_channels, H, W, 64, st));
  r.conv("encoder.conv_in", r.A, r.X, nullptr, H, W, 64, g.block_out[0]);
  int ch = g.block_out[0], hh = H, ww = W;
  for (int i = 0; i < 4; ++i) {
    for (int j = 0; j < g.layers_per_block; ++j) {
      r.resnet("encoder.down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), hh, ww, ch,
               g.block_out[i]);
      ch = g.block_out[i];
    }
    if (i != 3) {
      r.conv("encoder.down_blocks." + std::to_string(i) + ".downsamplers.0.conv", r.X, r.A, nullptr, hh, ww,
             ch, ch, 2);
      r.swapXA();
      hh /= 2;
      ww /= 2;
    }
  }
  r.mid("encoder.mid_block", hh, ww, ch);
  r.gn("encoder.conv_norm_out", r.X, r.A, (long long)hh * ww, ch, 1);
  r.conv("encoder.conv_out", r.A, static_cast<bf16_t*>(moments_nchw), nullptr, hh, ww, ch,
         2 * g.latent_channels, 1, 1);
  return r.rc;
}

static int vae_decode_impl(b2f_vae* h, const void* z_nchw, int N, int h_lat, int w_lat, void* image_nchw, int out_mode,
                           void* ws, size_t ws_bytes, b2f_stream_t stream_) {
  VaeCtx* c = reinterpret_cast<VaeCtx*>(h);
  if (!c || !z_nchw || !image_nchw || !ws || N <= 0 || h_lat <= 0 || w_lat <= 0) return B2F_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int H = h_lat * 8, W = w_lat * 8;
  VaeRun r;
  int rc = vae_setup(c, &r, N, H, W, ws, ws_bytes, st);
  if (rc) return rc;
  const b2f_vae_cfg& g = c->cfg;
  r.run(nchw_to_nhwc_pad(z_nchw, 0, r.A, N, g.latent_channels, h_lat, w_lat, 64, st));
  int ch = g.block_out[3], hh = h_lat, ww = w_lat;
  r.conv("decoder.conv_in", r.A, r.X, nullptr, hh, ww, 64, ch);
  r.mid("decoder.mid_block", hh, ww, ch);
  for (int i = 0; i < 4; ++i) {
    const int co = g.block_out[3 - i];
    for (int j = 0; j < g.layers_per_block + 1; ++j) {
      r.resnet("decoder.up_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), hh, ww, ch, co);
      ch = co;
    }
    if (i != 3) {
      r.run(upsample2x(r.X, r.A, N, hh, ww, ch, st));
      hh *= 2;
      ww *= 2;
      r.conv("decoder.up_blocks." + std::to_string(i) + ".upsamplers.0.conv", r.A, r.X, nullptr, hh, ww, ch, ch);
    }
  }
  r.gn("decoder.conv_norm_out", r.X, r.A, (long long)hh * ww, ch, 1);
  r.conv("decoder.conv_out", r.A, static_cast<bf16_t*>(image_nchw), nullptr, hh, ww, ch, g.out_channels, 1, out_mode);
  return r.rc;
}

int b2f_vae_decode(b2f_vae* h, const void* z_nchw, int N, int h_lat, int w_lat, void* image_nchw,
                   void* ws, size_t ws_bytes, b2f_stream_t stream_) {
  return vae_decode_impl(h, z_nchw, N, h_lat, w_lat, image_nchw, 1, ws, ws_bytes, stream_);
}

int b2f_vae_decode_u8(b2f_vae* h, const void* z_nchw, int N, int h_lat, int w_lat, void* image_u8_nhwc,
                      void* ws, size_t ws_bytes, b2f_stream_t stream_) {
  return vae_decode_impl(h, z_nchw, N, h_lat, w_lat, image_u8_nhwc, 2, ws, ws_bytes, stream_);
}

}  // extern "C"
