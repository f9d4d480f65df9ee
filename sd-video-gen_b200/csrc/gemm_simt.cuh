// Plain fp32 FMA GEMM, C = A[M,K] * W[N,K]^T + Epilogue.  Bring-up / cross-check path ("fp32_simt" mode):
// exact fp32 products and accumulation on the CUDA cores, same epilogue as the tensor-core kernel, so the
// whole forward can be validated against the oracle independently of tcgen05/TMA.  Not a performance path.
#pragma once
#include "common.cuh"

namespace sdvg {

constexpr int kSimtBM = 64, kSimtBN = 64, kSimtBK = 16;

__global__ void __launch_bounds__(256) gemm_simt_kernel(const float* __restrict__ A, int lda,
                                                        const float* __restrict__ W, int ldw, int M, int N, int K,
                                                        const __grid_constant__ Epilogue e) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sA[kSimtBK][kSimtBM + 4];
  __shared__ float sW[kSimtBK][kSimtBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each a 4 x 4 micro-tile
  const int m0 = blockIdx.y * kSimtBM, n0 = blockIdx.x * kSimtBN;
  float acc[4][4] = {};
  // loader mapping: 64 rows x 16 k = 1024 elements, 4 per thread: row = tid / 4, k = (tid % 4) * 4 .. +3
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  for (int k0 = 0; k0 < K; k0 += kSimtBK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + lk + j;
      const int ra = m0 + lr, rw = n0 + lr;
      sA[lk + j][lr] = (ra < M && k < K) ? __ldg(A + static_cast<size_t>(ra) * lda + k) : 0.0f;
      sW[lk + j][lr] = (rw < N && k < K) ? __ldg(W + static_cast<size_t>(rw) * ldw + k) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSimtBK; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = sW[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= M) continue;
    const int pe_row = epi_pe_row(e, row);
    const int out_row = epi_out_row(e, row);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= N) continue;
      epi_store(e, epi_value(e, acc[i][j], row, col, pe_row), row, out_row, col);
    }
  }
}

inline cudaError_t launch_gemm_simt(const float* A, int lda, const float* W, int ldw, int M, int N, int K,
                                    const Epilogue& e, cudaStream_t stream) {
  dim3 grid(ceil_div(N, kSimtBN), ceil_div(M, kSimtBM));
  return launch_kernel(gemm_simt_kernel, dim3(grid), dim3(256), 0, stream, A, lda, W, ldw, M, N, K, e);
}

}  // namespace sdvg
