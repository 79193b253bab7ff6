/* b2f.h — C ABI of the B200-native FLUX-Kontext denoising engine (libb2f.so).
 *
 * The reference (wyhlovecpp/GPT-Image-Edit) has no FFI: its hot path sits behind Python object
 * protocols whose arithmetic lives in diffusers 0.32.2 / torch.  Every entry
 * point below names the reference interface it replaces (path:line under /root/reference).
 *
 * Conventions
 *   - every call returns B2F_OK (0) or a negative error code; nothing throws or aborts;
 *   - all pointers are DEVICE pointers owned by the caller (PyTorch owns storage), row-major,
 *     innermost dimension contiguous, 16-byte aligned; `ld*` arguments are row pitches in
 *     elements;
 *   - all work is enqueued on the caller's stream (a cudaStream_t passed as void*); no call
 *     allocates device memory or synchronises the host unless stated;
 *   - dtype is bf16 (raw uint16 storage) unless a parameter says otherwise.
 */
#ifndef B2F_H_
#define B2F_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2F_OK 0
#define B2F_ERR_INVALID (-1)     /* bad shape / argument */
#define B2F_ERR_CUDA (-2)        /* CUDA runtime/driver error (message on stderr) */
#define B2F_ERR_UNSUPPORTED (-3) /* shape or mode not implemented */
#define B2F_ERR_ALIGN (-4)       /* pointer or pitch not 16-byte aligned */
#define B2F_ERR_NODEVICE (-5)    /* no sm_100 device visible */
#define B2F_ERR_WORKSPACE (-6)   /* workspace too small */

typedef void* b2f_stream_t; /* cudaStream_t */

const char* b2f_strerror(int code);
/* ABI version; bumped on any signature change. */
int b2f_version(void);
/* Device facts the host needs for grid sizing / reporting. Returns B2F_ERR_NODEVICE without GPU. */
int b2f_device_info(int* num_sms, int* cc_major, int* cc_minor, size_t* smem_optin);
/* Number of kernels this library has launched since load (bench.py's gpu_launches claim). */
uint64_t b2f_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Linear layer: out[M,N] = epilogue(A[M,K] · W[N,K]^T + bias[N]).  tcgen05.mma, TMA, TMEM.
 * Replaces torch.nn.functional.linear → cuBLASLt as reached by diffusers' nn.Linear modules
 * (FluxTransformer2DModel: x_embedder, context_embedder, to_q/k/v, to_out, ff.net.*, proj_mlp,
 * proj_out, norm*.linear — SURVEY.md Appendix A.1/A.6; call site univa/utils/flux_pipeline.py:1067).
 *
 * epilogue:
 *   B2F_EPI_BIAS        out = bf16(acc + bias)
 *   B2F_EPI_GELU_TANH   out = bf16(gelu_tanh(bf16(acc + bias)))     (ff.net.0 / proj_mlp)
 *   B2F_EPI_SILU        out = bf16(silu(bf16(acc + bias)))          (time_text_embed MLPs, MLP2)
 *   B2F_EPI_GATE_RESID  out = bf16(resid + bf16(gate[b,n] * bf16(acc + bias)))
 *                       with b = row / rows_per_batch                (x = x + gate * proj(...))
 * bias may be NULL.  resid may alias out.  K % 8 == 0, N % 8 == 0.
 */
#define B2F_EPI_BIAS 0
#define B2F_EPI_GELU_TANH 1
#define B2F_EPI_SILU 2
#define B2F_EPI_GATE_RESID 3

int b2f_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias,
                  void* out, int64_t ldc, int M, int N, int K, int epilogue, const void* resid,
                  int64_t ldr, const void* gate, int64_t gate_ld, int rows_per_batch,
                  b2f_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused softmax attention, head_dim 128:  O = softmax(Q K^T * scale [+ causal mask]) V.
 * tcgen05 QK^T and PV with S/P/O in TMEM, K/V streamed by TMA, online softmax (FA-style).
 * Replaces F.scaled_dot_product_attention in diffusers' FluxAttnProcessor2_0 (joint [txt;img]
 * attention of FluxTransformerBlock / FluxSingleTransformerBlock; reference call
 * site univa/utils/flux_pipeline.py:1067) and flash_attn reached through
 * attn_implementation="flash_attention_2" (univa/serve/cli.py:40) for the Qwen2.5-VL prefill.
 *
 * Layout: token-major.  q points at element [b=0, s=0, head 0, 0]; head h of token s of batch b
 * is at q + (b*Sq + s)*ldq + h*128 (same for k, v with Skv, ldk/ldv and Hkv heads; GQA maps
 * query head h to kv head h / (H/Hkv)).  out is [B, Sq, H*128] with row pitch ldo.  This lets
 * Q/K/V be column slices of one fused QKV projection buffer and lets out be a column slice of
 * the single-stream [attn | mlp] buffer — no transposes or concatenations.
 * causal != 0 requires Sq == Skv.
 */
int b2f_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                      int64_t ldv, void* out, int64_t ldo, int B, int H, int Hkv, int Sq, int Skv,
                      int head_dim, float scale, int causal, b2f_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B2F_H_ */
