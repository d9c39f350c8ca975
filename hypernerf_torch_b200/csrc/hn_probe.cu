// Single-tile UMMA probe: D[128 x N] = A[128 x K] * B[N x K]^T with A/B staged in shared memory in the
// un-swizzled interleave layout, K-major or MN-major.  It exists to pin the descriptor conventions of
// hn_ptx.cuh against a plain matmul on real hardware (tests/test_umma_probe.py); the product kernels use
// exactly these layouts.
#include "hn_ptx.cuh"
#include "hn_api_internal.h"

namespace hn {

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, float* __restrict__ D,
                  int N, int K, int a_mn, int b_mn) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * K * 2;
  const int tid = threadIdx.x;

  // stage A
  for (int i = tid; i < 128 * K; i += 128) {
    int r, k;
    if (!a_mn) { r = i / K; k = i % K; } else { k = i / 128; r = i % 128; }
    uint32_t off = !a_mn ? (uint32_t)(k / 8) * (128 * 16) + r * 16 + (k % 8) * 2
                         : (uint32_t)(r / 8) * (K * 16) + k * 16 + (r % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sA + off) = A[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    int n, k;
    if (!b_mn) { n = i / K; k = i % K; } else { k = i / N; n = i % N; }
    uint32_t off = !b_mn ? (uint32_t)(k / 8) * (N * 16) + n * 16 + (k % 8) * 2
                         : (uint32_t)(n / 8) * (K * 16) + k * 16 + (n % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sB + off) = B[i];
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (tid < 32) {
    tmem_alloc(&tmem_base_s, 256);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, a_mn, b_mn);
    const uint32_t a_lbo = !a_mn ? 128 * 16 : 128, a_sbo = !a_mn ? 128 : K * 16;
    const uint32_t b_lbo = !b_mn ? N * 16 : 128, b_sbo = !b_mn ? 128 : K * 16;
    const uint32_t a_step = !a_mn ? 2 * 128 * 16 : 256, b_step = !b_mn ? 2 * N * 16 : 256;
    for (int ks = 0; ks < K / 16; ++ks) {
      uint64_t ad = make_smem_desc(smem_u32(sA) + ks * a_step, a_lbo, a_sbo);
      uint64_t bd = make_smem_desc(smem_u32(sB) + ks * b_step, b_lbo, b_sbo);
      umma_bf16(tmem_base, ad, bd, idesc, ks > 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();

  const int warp = tid / 32, lane = tid % 32;
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_base, 256);
}

}  // namespace hn

extern "C" int hn_umma_probe(const void* A, const void* B, float* D, int N, int K, int a_mn, int b_mn,
                             void* stream) {
  if (N < 16 || N > 256 || N % 16 || K < 16 || K > 256 || K % 16) return hn::set_error(-1, "hn_umma_probe: bad N/K");
  size_t smem = (size_t)(128 + N) * K * 2;
  cudaError_t e = cudaFuncSetAttribute(hn::umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return hn::set_cuda_error(e, "hn_umma_probe: smem attr");
  hn::umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, D,
                                                              N, K, a_mn, b_mn);
  return hn::set_cuda_error(cudaGetLastError(), "hn_umma_probe: launch");
}
