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

// ------------------------------------------------------------------------------------------------------
// UMMA issue-rate microbenchmark (test hook): back-to-back K=16 UMMAs from resident shared-memory operands,
// no epilogue.  Reports cycles per UMMA for a layout / shape, so the fused kernels' tensor-pipe time can be
// separated from everything else.  Operand contents are whatever shared memory holds.
// ------------------------------------------------------------------------------------------------------
namespace hn {
__global__ void __launch_bounds__(128, 1)
umma_rate_kernel(int N, int ksteps, int reps, int swizzle, int nsub, int inner, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x;
  // zero-fill operands (128 x 256 A per sub, N x 256 B)
  const int total = nsub * 128 * 256 * 2 + N * 64 * 2;   // A: 128 x 256 per sub; B: N x 64 (4 k-steps, re-used)
  for (int i = tid * 16; i < total; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); fence_barrier_init(); }
  if (tid < 32) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (tid < 32) {  // converged warp; tcgen05 instructions guarded by elect_one_sync()
    const uint32_t sA = smem_u32(smem), sB = sA + nsub * 128 * 256 * 2;
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    long long t0 = clock64();
    uint32_t phase = 0;
    for (int r = 0; r < reps; ++r) {
      if (elect_one_sync()) {
      for (int in = 0; in < inner; ++in)
      for (int ks = 0; ks < ksteps; ++ks) {
        for (int sub = 0; sub < nsub; ++sub) {
          uint64_t ad, bd;
          if (swizzle != 1) {
            ad = make_smem_desc(sA + sub * 65536 + ks * 4096, 2048, 128);
            bd = make_smem_desc(sB + (ks % 4) * 2 * N * 16, N * 16, 128);
          } else {
            // 128B swizzle, K-major: 64-element (128 B) rows, 8-row atoms of 1024 B; K advance = 32 B inside the atom row
            ad = make_smem_desc(sA + sub * 65536 + (ks / 4) * 16384 + (ks % 4) * 32, 16, 1024) | ((uint64_t)2 << 61);
            bd = make_smem_desc(sB + (ks % 4) * 32, 16, 1024) | ((uint64_t)2 << 61);
          }
          umma_bf16(tmem_base + sub * 256, ad, bd, idesc, ks > 0);
        }
        // swizzle >= 2: also commit (un-waited, on a second barrier) after every (swizzle-1)-th k-step
        if (swizzle >= 2 && (ks % (swizzle - 1)) == (swizzle - 2)) umma_commit(&bar2);
      }
      umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, phase);
      phase ^= 1;
    }
    if (tid == 0) out[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_base, 512);
}
}  // namespace hn

extern "C" int hn_umma_rate(int N, int ksteps, int reps, int swizzle, int nsub, int inner, int grid, void* out_cycles,
                            void* stream) {
  if (N < 16 || N > 256 || N % 16 || ksteps < 1 || ksteps > 16 || nsub < 1 || nsub > 2) return hn::set_error(-1, "hn_umma_rate: bad args");
  size_t smem = (size_t)nsub * 128 * 256 * 2 + (size_t)N * 64 * 2 + 1024;
  cudaError_t e = cudaFuncSetAttribute(hn::umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return hn::set_cuda_error(e, "hn_umma_rate: smem attr");
  hn::umma_rate_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(N, ksteps, reps, swizzle, nsub, inner, (unsigned long long*)out_cycles);
  return hn::set_cuda_error(cudaGetLastError(), "hn_umma_rate: launch");
}

// ------------------------------------------------------------------------------------------------------
// TMEM read-throughput microbenchmark (test hook): nwarps warps each read `cols` accumulator columns of their
// 32 lanes with tcgen05.ld.32x32b.x32, `reps` times.  mode 0: load+wait per block; 1: two loads in flight.
// ------------------------------------------------------------------------------------------------------
namespace hn {
__global__ void __launch_bounds__(256, 1) tmem_rate_kernel(int cols, int reps, int mode, unsigned long long* out) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (mode == 0) {
      for (int c = 0; c < cols; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= v[j];
      }
    } else {
      for (int c = 0; c < cols; c += 64) {
        uint32_t v[32], w[32];
        tmem_ld32(taddr + c, v);
        tmem_ld32(taddr + c + 32, w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= v[j] ^ w[j];
      }
    }
  }
  long long t1 = clock64();
  if (acc == 0x12345678u) out[1] = acc;
  if (threadIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}
}  // namespace hn

extern "C" int hn_tmem_rate(int nwarps, int cols, int reps, int mode, void* out_cycles, void* stream) {
  if (nwarps < 1 || nwarps > 8 || cols < 64 || cols > 256 || cols % 64) return hn::set_error(-1, "hn_tmem_rate: bad args");
  hn::tmem_rate_kernel<<<1, nwarps * 32, 0, (cudaStream_t)stream>>>(cols, reps, mode, (unsigned long long*)out_cycles);
  return hn::set_cuda_error(cudaGetLastError(), "hn_tmem_rate: launch");
}
