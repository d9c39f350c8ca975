// Single-tile UMMA probe: D[128 x N] = A[128 x K] * B[N x K]^T with A/B staged in shared memory in the
// un-swizzled interleave layout, K-major or MN-major.  It exists to pin the descriptor conventions of
// hn_ptx.cuh against a plain matmul on real hardware (tests/test_umma_probe.py); the product kernels use
// exactly these layouts.
#include "hn_ptx.cuh"
#include "hn_api_internal.h"
#include "../../include/hypernerf_b200_probe.h"

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


// ------------------------------------------------------------------------------------------------------
// Second issue-rate microbenchmark (test hook): separates the per-UMMA cost by accumulator rotation, M, operand
// source (A from shared memory or TMEM) and CTA pairing (cta_group::1 vs ::2 on a 2-CTA cluster).
//   nacc   accumulators rotated (N * nacc <= 512 columns minus the TMEM A operand)
//   order  0: k-step outer, accumulator inner (adjacent UMMAs hit different accumulators)
//          1: accumulator outer, k-step inner (16 chained UMMAs per accumulator)
//   a_src  0: A from shared memory, distinct K slice per step; 1: same A slice every step; 2: A from TMEM (.ts)
// ------------------------------------------------------------------------------------------------------
namespace hn {
template <int CG>
__device__ __forceinline__ void umma2(uint32_t d, uint64_t ad, uint32_t a_tmem, bool a_in_tmem, uint64_t bd, uint32_t idesc,
                                      uint32_t acc) {
  if (CG == 1) {
    if (!a_in_tmem)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                   ::"r"(d), "r"(a_tmem), "l"(bd), "r"(idesc), "r"(acc) : "memory");
  } else {
    if (!a_in_tmem)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                   ::"r"(d), "r"(a_tmem), "l"(bd), "r"(idesc), "r"(acc) : "memory");
  }
}

template <int CG>
__global__ void __launch_bounds__(128, 1)
umma_rate2_kernel(int M, int N, int nacc, int order, int a_src, int reps, int inner, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int total = 128 * 256 * 2 + 256 * 64 * 2;   // A: 128 x 256; B: up to 256 x 64 (4 k-steps, re-used)
  for (int i = tid * 16; i < total; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (tid < 32) {
    if (CG == 1) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (tid < 32 && rank == 0) {
    const uint32_t sA = smem_u32(smem), sB = sA + 128 * 256 * 2;
    const uint32_t idesc = make_idesc_bf16(M, N, 0, 0);
    const uint32_t nb = (CG == 2) ? N / 2 : N;  // B rows held by this CTA
    const uint32_t a_tmem = tmem_base + 384;    // 128 columns of packed bf16 A (K = 256)
    long long t0 = clock64();
    uint32_t phase = 0;
    for (int r = 0; r < reps; ++r) {
      if (elect_one_sync()) {
        for (int in = 0; in < inner; ++in) {
          const int n_outer = order == 0 ? 16 : nacc, n_inner = order == 0 ? nacc : 16;
          for (int o = 0; o < n_outer; ++o)
            for (int i2 = 0; i2 < n_inner; ++i2) {
              const int ks = order == 0 ? o : i2, acc = order == 0 ? i2 : o;
              const uint64_t ad = make_smem_desc(sA + (a_src == 1 ? 0 : ks * 4096), 2048, 128);
              const uint64_t bd = make_smem_desc(sB + (ks % 4) * 2 * nb * 16, nb * 16, 128);
              umma2<CG>(tmem_base + acc * N, ad, a_tmem + ks * 8, a_src == 2, bd, idesc, ks > 0);
            }
        }
        if (CG == 1) umma_commit(&bar);
        else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      __syncwarp();
      mbar_wait(&bar, phase);
      phase ^= 1;
    }
    if (tid == 0) out[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  if (tid < 32) {
    if (CG == 1) tmem_dealloc(tmem_base, 512);
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}
}  // namespace hn

extern "C" int hn_umma_rate2(int cta_group, int M, int N, int nacc, int order, int a_src, int reps, int inner, int grid,
                             void* out_cycles, void* stream) {
  if (N < 16 || N > 256 || N % 16 || nacc < 1 || N * nacc > 384 || (cta_group != 1 && cta_group != 2))
    return hn::set_error(-1, "hn_umma_rate2: bad args");
  size_t smem = 128 * 256 * 2 + 256 * 64 * 2 + 1024;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cta_group; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e;
  if (cta_group == 1) {
    e = cudaFuncSetAttribute(hn::umma_rate2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, hn::umma_rate2_kernel<1>, M, N, nacc, order, a_src, reps, inner, (unsigned long long*)out_cycles);
  } else {
    e = cudaFuncSetAttribute(hn::umma_rate2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, hn::umma_rate2_kernel<2>, M, N, nacc, order, a_src, reps, inner, (unsigned long long*)out_cycles);
  }
  return hn::set_cuda_error(e, "hn_umma_rate2: launch");
}


// ------------------------------------------------------------------------------------------------------
// Third issue-rate microbenchmark: the same UMMA stream as a fully unrolled, straight-line sequence with
// descriptors advanced by compile-time constants (no per-instruction descriptor arithmetic), to separate the
// tensor pipe's own rate from the cost of the issuing thread's loop.
// ------------------------------------------------------------------------------------------------------
namespace hn {
template <int N, int NACC>
__global__ void __launch_bounds__(128, 1) umma_rate3_kernel(int reps, int inner, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x;
  const int total = NACC * 128 * 256 * 2 + 256 * 64 * 2;
  for (int i = tid * 16; i < total; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (tid < 32) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (tid < 32) {
    const uint32_t sA = smem_u32(smem), sB = sA + NACC * 128 * 256 * 2;
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint64_t ad0 = make_smem_desc(sA, 2048, 128), bd0 = make_smem_desc(sB, N * 16, 128);
    long long t0 = clock64();
    uint32_t phase = 0;
    for (int r = 0; r < reps; ++r) {
      if (elect_one_sync()) {
        for (int in = 0; in < inner; ++in) {
#pragma unroll
          for (int ks = 0; ks < 16; ++ks) {
#pragma unroll
            for (int a = 0; a < NACC; ++a)
              umma_bf16(tmem_base + a * 256, ad0 + (uint64_t)((a * 65536 + ks * 4096) >> 4), bd0 + (uint64_t)(((ks % 4) * 2 * N * 16) >> 4),
                        idesc, ks > 0 ? 1u : 0u);
          }
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
template <int N, int NACC>
static cudaError_t launch_rate3(int reps, int inner, int grid, unsigned long long* out, cudaStream_t st) {
  size_t smem = (size_t)NACC * 128 * 256 * 2 + 256 * 64 * 2 + 1024;
  cudaError_t e = cudaFuncSetAttribute(umma_rate3_kernel<N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  umma_rate3_kernel<N, NACC><<<grid, 128, smem, st>>>(reps, inner, out);
  return cudaGetLastError();
}
}  // namespace hn

extern "C" int hn_umma_rate3(int N, int nacc, int reps, int inner, int grid, void* out_cycles, void* stream) {
  cudaError_t e = cudaErrorInvalidValue;
  unsigned long long* o = (unsigned long long*)out_cycles;
  cudaStream_t st = (cudaStream_t)stream;
#define HN_R3(n, a) if (N == n && nacc == a) e = hn::launch_rate3<n, a>(reps, inner, grid, o, st);
  HN_R3(256, 1) HN_R3(256, 2) HN_R3(128, 1) HN_R3(128, 2) HN_R3(64, 1) HN_R3(64, 2) HN_R3(16, 1) HN_R3(16, 2)
#undef HN_R3
  return hn::set_cuda_error(e, "hn_umma_rate3");
}


// ------------------------------------------------------------------------------------------------------
// Epilogue microbenchmark (test hook): 8 warps drain a 256-column fp32 accumulator (thread = row) `reps` times with
// a selectable subset of the forward epilogue's work, to find which pipe bounds it.
//   bit 0: convert (F2FP.RELU pack)       bit 1: st.shared 16 B packets (next layer's operand)
//   bit 2: st.global 16 B packets (stash) bit 3: bias from shared memory (8 x LDS.128 + 32 FADD per 32 columns)
//   bit 4: two TMEM loads in flight        bit 5: 64 B per thread contiguous global stores instead of 16 B
// ------------------------------------------------------------------------------------------------------
namespace hn {
template <int mode>
__global__ void __launch_bounds__(256, 1) epi_rate_kernel(int reps, uint8_t* gout, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float sbias[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  sbias[threadIdx.x] = 0.001f * threadIdx.x;
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int sub = warp >> 2, quarter = warp & 3, row = quarter * 32 + lane;
  const uint32_t taddr = tmem_base_s + ((uint32_t)(quarter * 32) << 16) + sub * 256;
  uint8_t* act_row = smem + sub * 65536 + row * 16;
  uint4* save_row = reinterpret_cast<uint4*>(gout + ((size_t)blockIdx.x * 4 + sub * 2 + (row >> 6)) * 32 * 1024) + (row & 63);
  uint32_t sink = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    uint32_t ra[32], rb[32];
    tmem_ld32(taddr, ra);
    if (mode & (64 | 128)) {   // previous layer's bulk stores must have finished reading the rows we overwrite
      bulk_wait_read<0>();
      if (mode & 128) asm volatile("bar.sync 1, 256;" ::: "memory"); else __syncwarp();
    }
#pragma unroll 1
    for (int c0 = 0; c0 < 256; c0 += 64) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t* cur = h ? rb : ra;
        uint32_t* nxt = h ? ra : rb;
        tmem_ld_wait();
        const int c = c0 + 32 * h;
        if (c + 32 < 256 && (mode & 16)) tmem_ld32(taddr + c + 32, nxt);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(cur[8 * q + j]);
          if (mode & 8) {
            const float4 b0 = *reinterpret_cast<const float4*>(sbias + c + 8 * q);
            const float4 b1 = *reinterpret_cast<const float4*>(sbias + c + 8 * q + 4);
            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
          }
          uint4 o;
          if (mode & 1) {
            o.x = pack_bf16_relu(v[0], v[1]); o.y = pack_bf16_relu(v[2], v[3]);
            o.z = pack_bf16_relu(v[4], v[5]); o.w = pack_bf16_relu(v[6], v[7]);
          } else {
            o.x = __float_as_uint(v[0]) ^ __float_as_uint(v[1]); o.y = __float_as_uint(v[2]) ^ __float_as_uint(v[3]);
            o.z = __float_as_uint(v[4]) ^ __float_as_uint(v[5]); o.w = __float_as_uint(v[6]) ^ __float_as_uint(v[7]);
          }
          if (mode & 2) *reinterpret_cast<uint4*>(act_row + ((c >> 3) + q) * 2048) = o;
          if (mode & 4) {
            if (mode & 32) reinterpret_cast<uint4*>(gout + ((size_t)blockIdx.x * 256 + threadIdx.x) * 512)[(c >> 3) + q] = o;
            else save_row[((c >> 3) + q) * 64] = o;
          }
          sink ^= o.x ^ o.y ^ o.z ^ o.w;
        }
        if (c + 32 < 256 && !(mode & 16)) tmem_ld32(taddr + c + 32, nxt);
      }
    }
    if (mode & 64) {          // per warp: 32 copies of 512 B (lane = chunk)
      fence_proxy_async_smem();
      __syncwarp();
      bulk_s2g(gout + ((size_t)blockIdx.x * 4 + sub * 2 + (quarter >> 1)) * 32 * 1024 + lane * 1024 + (quarter & 1) * 512,
               smem + sub * 65536 + lane * 2048 + quarter * 512, 512);
      bulk_commit();
    }
    if (mode & 128) {         // per sub tile: one copy of 64 KB
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if ((threadIdx.x & 127) == 0) {
        bulk_s2g(gout + ((size_t)blockIdx.x * 2 + sub) * 65536, smem + sub * 65536, 65536);
        bulk_commit();
      }
    }
  }
  if (mode & (64 | 128)) bulk_wait_read<0>();
  long long t1 = clock64();
  if (sink == 0x12345678u) out[1] = sink;
  if (threadIdx.x == 0) out[blockIdx.x * 2] = (unsigned long long)(t1 - t0);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}
}  // namespace hn

extern "C" int hn_epi_rate(int mode, int reps, int grid, void* gout /* grid * 128 KB */, void* out_cycles, void* stream) {
  cudaError_t e = cudaErrorInvalidValue;
  const int smem = 131072 + 1024;
#define HN_EPI(m)                                                                                              \
  if (mode == m) {                                                                                             \
    e = cudaFuncSetAttribute(hn::epi_rate_kernel<m>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);       \
    if (e == cudaSuccess) {                                                                                    \
      hn::epi_rate_kernel<m><<<grid, 256, smem, (cudaStream_t)stream>>>(reps, (uint8_t*)gout, (unsigned long long*)out_cycles); \
      e = cudaGetLastError();                                                                                  \
    }                                                                                                          \
  }
  HN_EPI(0) HN_EPI(16) HN_EPI(17) HN_EPI(19) HN_EPI(21) HN_EPI(23) HN_EPI(31) HN_EPI(27) HN_EPI(20) HN_EPI(52) HN_EPI(18) HN_EPI(24) HN_EPI(25) HN_EPI(7) HN_EPI(55) HN_EPI(83) HN_EPI(147) HN_EPI(91) HN_EPI(155)
#undef HN_EPI
  return hn::set_cuda_error(e, "hn_epi_rate");
}


// ------------------------------------------------------------------------------------------------------
// CTA-pair UMMA probe (test hook): D[256 x N] = A[256 x K] * B[N x K]^T on a 2-CTA cluster with cta_group::2.
// CTA r stages A rows [128 r, 128 r + 128) and B rows [r N/2, (r + 1) N/2) in the un-swizzled K-major interleave
// layout at the same shared-memory offsets; the leader issues, the commit is multicast to both CTAs, each CTA drains
// its own 128 accumulator rows.  Pins the operand split that the pair mode of the fused kernels relies on.
// ------------------------------------------------------------------------------------------------------
namespace hn {
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma_probe2_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, float* __restrict__ D, int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t rank = cluster_ctarank();
  const int tid = threadIdx.x, NH = N / 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * K * 2;
  for (int i = tid; i < 128 * K; i += 128) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sA + (uint32_t)(k / 8) * (128 * 16) + r * 16 + (k % 8) * 2) = A[(size_t)(rank * 128 + r) * K + k];
  }
  for (int i = tid; i < NH * K; i += 128) {
    const int n = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sB + (uint32_t)(k / 8) * (NH * 16) + n * 16 + (k % 8) * 2) = B[(size_t)(rank * NH + n) * K + k];
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (tid < 32) { tmem_alloc2(&tmem_base_s, 256); tmem_relinquish2(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (rank == 0 && tid < 32) {
    const uint32_t idesc = make_idesc_bf16(256, N, 0, 0);
    if (elect_one_sync()) {
      for (int ks = 0; ks < K / 16; ++ks) {
        uint64_t ad = make_smem_desc(smem_u32(sA) + ks * 2 * 128 * 16, 128 * 16, 128);
        uint64_t bd = make_smem_desc(smem_u32(sB) + ks * 2 * NH * 16, NH * 16, 128);
        umma2_bf16(tmem_base, ad, bd, idesc, ks > 0);
      }
      umma2_commit_mc(&bar, 3);
    }
    __syncwarp();
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
    for (int j = 0; j < 16; ++j) D[(size_t)(rank * 128 + row) * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (tid < 32) tmem_dealloc2(tmem_base, 256);
}
}  // namespace hn

extern "C" int hn_umma_probe2(const void* A, const void* B, float* D, int N, int K, void* stream) {
  if (N < 16 || N > 256 || N % 16 || K < 16 || K > 256 || K % 16) return hn::set_error(-1, "hn_umma_probe2: bad N/K");
  size_t smem = (size_t)(128 + N / 2) * K * 2;
  cudaError_t e = cudaFuncSetAttribute(hn::umma_probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return hn::set_cuda_error(e, "hn_umma_probe2: smem attr");
  hn::umma_probe2_kernel<<<2, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, D, N, K);
  return hn::set_cuda_error(cudaGetLastError(), "hn_umma_probe2: launch");
}


// ------------------------------------------------------------------------------------------------------
// Unrolled issue-rate microbenchmark for the CTA-pair form (cta_group::2, M = 256, each CTA holding N/2 rows of B).
// ------------------------------------------------------------------------------------------------------
namespace hn {
template <int N, int NACC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) umma_rate4_kernel(int reps, int inner, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x;
  const uint32_t rank = cluster_ctarank();
  const int total = NACC * 128 * 256 * 2 + 128 * 64 * 2;
  for (int i = tid * 16; i < total; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (tid < 32) { tmem_alloc2(&tmem_base_s, 512); tmem_relinquish2(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (tid < 32 && rank == 0) {
    const uint32_t sA = smem_u32(smem), sB = sA + NACC * 128 * 256 * 2;
    constexpr uint32_t idesc = make_idesc_bf16(256, N, 0, 0);
    const uint64_t ad0 = make_smem_desc(sA, 2048, 128), bd0 = make_smem_desc(sB, (N / 2) * 16, 128);
    long long t0 = clock64();
    uint32_t phase = 0;
    for (int r = 0; r < reps; ++r) {
      if (elect_one_sync()) {
        for (int in = 0; in < inner; ++in) {
#pragma unroll
          for (int ks = 0; ks < 16; ++ks) {
#pragma unroll
            for (int a = 0; a < NACC; ++a)
              umma2_bf16(tmem_base + a * 256, ad0 + (uint64_t)((a * 65536 + ks * 4096) >> 4), bd0 + (uint64_t)(((ks % 4) * 2 * (N / 2) * 16) >> 4),
                         idesc, ks > 0 ? 1u : 0u);
          }
        }
        umma2_commit_mc(&bar, 1);
      }
      __syncwarp();
      mbar_wait(&bar, phase);
      phase ^= 1;
    }
    if (tid == 0) out[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (tid < 32) tmem_dealloc2(tmem_base, 512);
}
template <int N, int NACC>
static cudaError_t launch_rate4(int reps, int inner, int grid, unsigned long long* out, cudaStream_t st) {
  size_t smem = (size_t)NACC * 128 * 256 * 2 + 128 * 64 * 2 + 1024;
  cudaError_t e = cudaFuncSetAttribute(umma_rate4_kernel<N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  umma_rate4_kernel<N, NACC><<<grid, 128, smem, st>>>(reps, inner, out);
  return cudaGetLastError();
}
}  // namespace hn

extern "C" int hn_umma_rate4(int N, int nacc, int reps, int inner, int grid, void* out_cycles, void* stream) {
  cudaError_t e = cudaErrorInvalidValue;
  unsigned long long* o = (unsigned long long*)out_cycles;
  cudaStream_t st = (cudaStream_t)stream;
#define HN_R4(n, a) if (N == n && nacc == a) e = hn::launch_rate4<n, a>(reps, inner, grid, o, st);
  HN_R4(256, 1) HN_R4(256, 2) HN_R4(128, 1) HN_R4(128, 2) HN_R4(64, 1) HN_R4(64, 2) HN_R4(16, 1) HN_R4(16, 2)
#undef HN_R4
  return hn::set_cuda_error(e, "hn_umma_rate4");
}


// ------------------------------------------------------------------------------------------------------
// Drain-under-UMMA microbenchmark (test hook): warps 0-3 drain a 128-row x 256-column accumulator (sub-tile 0) `reps`
// times with a selectable subset of the epilogue's work while, if umma != 0, warp 4 keeps the tensor pipe busy with
// back-to-back M = 128, N = umma_n, K = 16 UMMAs on sub-tile 1's accumulator from resident shared-memory operands, and, if
// tma != 0, warp 5 keeps re-loading a 16 KB weight stage from global memory by bulk copies (the weight ring's traffic).
//   mode bit 0: bf16 pack   bit 1: st.shared of the packed row   bit 2: st.global (stash layout, streaming)
//        bit 3: bias (8 x LDS.128 + 32 FADD per 32 columns)
// out[b*4 + 0] = drain cycles, [1] = UMMAs retired while the drain ran, [2] = issuer cycles, [3] = bulk copies completed.
// ------------------------------------------------------------------------------------------------------
namespace hn {
__global__ void __launch_bounds__(256, 1) overlap_rate_kernel(int mode, int umma, int umma_n, int tma, int reps, uint8_t* gout,
                                                              const uint8_t* gsrc, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // layout: [0, 64 KB) drain target (sub-tile 0 operand rows) | [64 KB, 80 KB) A: 128 x 64 | [80 KB, 112 KB) B: 256 x 64 | [112 KB, 128 KB) weight stage
  __shared__ uint64_t bar_mma, bar_tma;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int done;
  __shared__ __align__(16) float sbias[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  sbias[threadIdx.x] = 0.001f * threadIdx.x;
  for (int i = threadIdx.x * 16; i < 112 * 1024; i += 256 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  if (threadIdx.x == 0) { mbar_init(&bar_mma, 1); mbar_init(&bar_tma, 1); done = 0; fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp < 4) {
    const int row = warp * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint8_t* act_row = smem + row * 16;
    uint4* save_row = reinterpret_cast<uint4*>(gout + ((size_t)blockIdx.x * 2 + (row >> 6)) * 32 * 1024) + (row & 63);
    uint32_t sink = 0;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      uint32_t ra[32], rb[32];
      tmem_ld32(taddr, ra);
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 64) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t* cur = h ? rb : ra;
          uint32_t* nxt = h ? ra : rb;
          tmem_ld_wait();
          const int c = c0 + 32 * h;
          if (c + 32 < 256) tmem_ld32(taddr + c + 32, nxt);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(cur[8 * q + j]);
            if (mode & 8) {
              const float4 b0 = *reinterpret_cast<const float4*>(sbias + c + 8 * q);
              const float4 b1 = *reinterpret_cast<const float4*>(sbias + c + 8 * q + 4);
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            uint4 o;
            if (mode & 1) {
              o.x = pack_bf16_relu(v[0], v[1]); o.y = pack_bf16_relu(v[2], v[3]);
              o.z = pack_bf16_relu(v[4], v[5]); o.w = pack_bf16_relu(v[6], v[7]);
            } else {
              o.x = __float_as_uint(v[0]) ^ __float_as_uint(v[1]); o.y = __float_as_uint(v[2]) ^ __float_as_uint(v[3]);
              o.z = __float_as_uint(v[4]) ^ __float_as_uint(v[5]); o.w = __float_as_uint(v[6]) ^ __float_as_uint(v[7]);
            }
            if (mode & 2) *reinterpret_cast<uint4*>(act_row + ((c >> 3) + q) * 2048) = o;
            if (mode & 4) __stcs(&save_row[((c >> 3) + q) * 64], o);
            sink ^= o.x ^ o.y ^ o.z ^ o.w;
          }
        }
      }
    }
    long long t1 = clock64();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 0) { out[blockIdx.x * 4] = (unsigned long long)(t1 - t0); done = 1; }
    if (sink == 0x12345678u) out[1] = sink;
  } else if (warp == 4) {
    if (umma) {
      const uint32_t sA = smem_u32(smem) + 65536, sB = sA + 16384;
      const uint32_t idesc = make_idesc_bf16(128, umma_n, 0, 0);
      unsigned long long n = 0;
      uint32_t phase = 0;
      long long t0 = clock64();
      while (!done) {
        if (elect_one_sync()) {
#pragma unroll 1
          for (int ks = 0; ks < 16; ++ks) {
            const uint64_t ad = make_smem_desc(sA + (ks % 4) * 4096, 2048, 128);
            const uint64_t bd = make_smem_desc(sB + (ks % 4) * 2 * umma_n * 16, umma_n * 16, 128);
            umma_bf16(tmem_base + 256, ad, bd, idesc, ks > 0);
          }
          umma_commit(&bar_mma);
        }
        __syncwarp();
        mbar_wait(&bar_mma, phase);
        phase ^= 1;
        n += 16;
      }
      if (lane == 0) { out[blockIdx.x * 4 + 1] = n; out[blockIdx.x * 4 + 2] = (unsigned long long)(clock64() - t0); }
    }
  } else if (warp == 5) {
    if (tma && lane == 0) {
      unsigned long long n = 0;
      uint32_t phase = 0;
      while (!done) {
        mbar_arrive_expect_tx(&bar_tma, 16384);
        bulk_g2s(smem + 112 * 1024, gsrc + (size_t)((n * 7 + blockIdx.x) % 64) * 16384, 16384, &bar_tma);
        mbar_wait(&bar_tma, phase);
        phase ^= 1;
        ++n;
      }
      out[blockIdx.x * 4 + 3] = n;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}
}  // namespace hn

extern "C" int hn_overlap_rate(int mode, int umma, int umma_n, int tma, int reps, int grid, void* gout, const void* gsrc,
                               void* out, void* stream) {
  if (umma_n < 16 || umma_n > 256 || umma_n % 16) return hn::set_error(-1, "hn_overlap_rate: bad N");
  const int smem = 128 * 1024 + 1024;
  cudaError_t e = cudaFuncSetAttribute(hn::overlap_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return hn::set_cuda_error(e, "hn_overlap_rate: smem attr");
  hn::overlap_rate_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(mode, umma, umma_n, tma, reps, (uint8_t*)gout,
                                                                     (const uint8_t*)gsrc, (unsigned long long*)out);
  return hn::set_cuda_error(cudaGetLastError(), "hn_overlap_rate");
}
