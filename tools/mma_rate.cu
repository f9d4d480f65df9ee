// tcgen05.mma issue / execution rate for the small tiles of the persistent kernel (tools/README.md).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sd-video-gen_b200/csrc -o tools/scratch/mma_rate tools/mma_rate.cu
// One thread issues `n` MMAs (cta_group::1, kind::f16, both operands from shared memory, SWIZZLE_128B K-major
// descriptors as in persistent.cuh) into `nacc` rotating accumulators, commits, waits: cycles per MMA.
// Operand contents are irrelevant (zeros).  Variants: M (64 / 128), N, accumulators, operand addresses advancing through
// a ring (as the kernel does) or fixed.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"

using namespace sdvg;

__device__ __forceinline__ void umma_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
      : "memory");
}

// 32 MMAs (8 blocks of 4 K steps, as one tile of the persistent kernel) fully unrolled, repeated `tiles` times
template <int M, int N, int NACC, int CE>
__global__ void __launch_bounds__(128, 1) rate_kernel(int tiles, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar[0], 1); ptx::mbar_init(&bar[1], 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc_f16(M, N, false);
    const uint32_t a0 = ((ptx::smem_u32(smem) & 0x3FFFF) >> 4) | (1u << 16);                 // weights: 8 blocks x 16 KB
    const uint32_t b0 = ((ptx::smem_u32(smem + 128 * 1024) & 0x3FFFF) >> 4) | (1u << 16);    // activations
    constexpr uint32_t acc_cols = (N + 31) & ~31;
    uint32_t phase = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int t = 0; t < tiles; ++t) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int blk = i >> 2, ks = i & 3;
          umma_lo(tmem + (i % NACC) * acc_cols, a0 + blk * 1024 + ks * 2, b0 + (blk & 1) * 1024 + ks * 2, idesc, i >= NACC ? 1u : 0u);
          if (CE && (i % CE) == CE - 1) ptx::umma_commit(&bar[1]);
        }
      }
      const long long t1 = clock64();
      ptx::umma_commit(&bar[0]);
      ptx::mbar_wait(&bar[0], phase);
      phase ^= 1;
      const long long t2 = clock64();
      if (rep == 2 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tmem, 512);
}

template <int M, int N, int NACC, int CE>
void run(long long* d, int ctas) {
  const int tiles = 8;
  auto k = rate_kernel<M, N, NACC, CE>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<<<ctas, 128, 200 * 1024>>>(tiles, d);
  long long h[2];
  cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  printf("%3d %3d %d %2d %3d : %6.1f %6.1f\n", M, N, NACC, CE, ctas, double(h[0]) / (32 * tiles), double(h[1]) / (32 * tiles));
}

template <int M, int N>
void shape(long long* d) {
  run<M, N, 1, 0>(d, 1); run<M, N, 2, 0>(d, 1); run<M, N, 4, 0>(d, 1);
  run<M, N, 2, 4>(d, 1); run<M, N, 2, 8>(d, 1); run<M, N, 2, 32>(d, 1); run<M, N, 2, 1>(d, 1);
  run<M, N, 2, 4>(d, 132);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  printf("M N nacc commit_every ctas : issue_clk_per_mma total_clk_per_mma\n");
  shape<128, 16>(d); shape<128, 48>(d); shape<128, 64>(d); shape<128, 128>(d);
  shape<64, 16>(d); shape<64, 48>(d); shape<64, 64>(d); shape<64, 128>(d);
  run<128, 256, 1, 0>(d, 1); run<128, 256, 2, 0>(d, 1); run<64, 256, 2, 0>(d, 1);
  return 0;
}
