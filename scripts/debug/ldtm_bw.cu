// Micro-benchmark: tcgen05.ld (TMEM -> registers) throughput of one SM as a function of the number of warps reading
// and of the load shape.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ldtm_bw scripts/debug/ldtm_bw.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define LD_X32(taddr, r)                                                                             \
  asm volatile(                                                                                      \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                      \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"  \
      "%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                         \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),          \
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),      \
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),   \
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),   \
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                           \
      : "r"(taddr))

__global__ void __launch_bounds__(512, 1) ldtm_kernel(int iters, int cols_per_iter, long long* out_cycles, uint32_t* sink) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_ptr + (uint32_t((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int c = 0; c < cols_per_iter; c += 32) {
      uint32_t r[32];
      LD_X32(base + ((c + (warp >> 2) * 64) & 511), r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= r[i];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_ptr), "r"(512u) : "memory");
}

// the same with all loads of a 128-column row issued before one wait (what the softmax warps do)
__global__ void __launch_bounds__(512, 1) ldtm_kernel4(int iters, long long* out_cycles, uint32_t* sink) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_ptr + (uint32_t((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 128;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r0[32], r1[32], r2[32], r3[32];
    LD_X32(base + 0, r0);
    LD_X32(base + 32, r1);
    LD_X32(base + 64, r2);
    LD_X32(base + 96, r3);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= r0[i] ^ r1[i] ^ r2[i] ^ r3[i];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_ptr), "r"(512u) : "memory");
}

int main() {
  long long* d_cyc;
  uint32_t* d_sink;
  cudaMalloc(&d_cyc, 8 * 256);
  cudaMalloc(&d_sink, 64);
  const int iters = 2000;
  for (int warps : {1, 2, 4, 8, 16}) {
    for (int pass = 0; pass < 2; ++pass) {
      ldtm_kernel<<<1, warps * 32>>>(iters, 128, d_cyc, d_sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    }
    long long c;
    cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
    const double bytes = double(iters) * 128 * 32 * 4 * warps;
    printf("{\"kernel\": \"ld_x32_wait_each\", \"warps\": %d, \"cycles\": %lld, \"bytes_per_clk_sm\": %.1f, \"clk_per_x32\": %.1f}\n", warps, c,
           bytes / c, double(c) / (iters * 4.0));
    for (int pass = 0; pass < 2; ++pass) {
      ldtm_kernel4<<<1, warps * 32>>>(iters, d_cyc, d_sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("{\"kernel\": \"ld_4x_x32_one_wait\", \"warps\": %d, \"cycles\": %lld, \"bytes_per_clk_sm\": %.1f, \"clk_per_row128\": %.1f}\n", warps, c,
           bytes / c, double(c) / iters);
  }
  return 0;
}
