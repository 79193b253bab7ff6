// extern "C" surface of libb2f.so — thin argument marshalling over the b2f:: launchers.
#include <atomic>

#include "host_common.h"

namespace b2f {
extern std::atomic<uint64_t> g_launch_count;
int gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* out,
              int64_t ldc, int M, int N, int K, int epilogue, const void* resid, int64_t ldr,
              const void* gate, int64_t gate_ld, int rows_per_batch, cudaStream_t stream);
int attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                  int64_t ldv, void* out, int64_t ldo, int B, int H, int Hkv, int Sq, int Skv,
                  int head_dim, float scale, int causal, cudaStream_t stream);
}  // namespace b2f

extern "C" {

const char* b2f_strerror(int code) {
  switch (code) {
    case B2F_OK: return "ok";
    case B2F_ERR_INVALID: return "invalid argument or shape";
    case B2F_ERR_CUDA: return "CUDA error (see stderr)";
    case B2F_ERR_UNSUPPORTED: return "unsupported shape or mode";
    case B2F_ERR_ALIGN: return "pointer or pitch not 16-byte aligned";
    case B2F_ERR_NODEVICE: return "no sm_100 device";
    case B2F_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown error";
  }
}

int b2f_version(void) { return 1; }

int b2f_device_info(int* num_sms, int* cc_major, int* cc_minor, size_t* smem_optin) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return B2F_ERR_NODEVICE;
  }
  int dev = 0, v = 0;
  cudaGetDevice(&dev);
  if (num_sms) cudaDeviceGetAttribute(num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (cc_major) cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev);
  if (cc_minor) cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (smem_optin) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    *smem_optin = (size_t)v;
  }
  return B2F_OK;
}

uint64_t b2f_launch_count(void) { return b2f::g_launch_count.load(); }

int b2f_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias,
                  void* out, int64_t ldc, int M, int N, int K, int epilogue, const void* resid,
                  int64_t ldr, const void* gate, int64_t gate_ld, int rows_per_batch,
                  b2f_stream_t stream) {
  return b2f::gemm_bf16(A, lda, W, ldw, bias, out, ldc, M, N, K, epilogue, resid, ldr, gate,
                        gate_ld, rows_per_batch, static_cast<cudaStream_t>(stream));
}

int b2f_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                      int64_t ldv, void* out, int64_t ldo, int B, int H, int Hkv, int Sq, int Skv,
                      int head_dim, float scale, int causal, b2f_stream_t stream) {
  return b2f::attention_fwd(q, ldq, k, ldk, v, ldv, out, ldo, B, H, Hkv, Sq, Skv, head_dim, scale,
                            causal, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
