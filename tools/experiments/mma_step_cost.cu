// Round-2 experiment: cycles per tcgen05.mma step by kind, N and CTA-pair mode (operands in shared memory, accumulator in
// TMEM), measured by issuing a long back-to-back run of steps from one thread and timing it with clock64 around the
// commit -> mbarrier wait.  Answers: does a kind::f8f6f4 K=32 step cost the same as a kind::f16 K=16 step (same operand
// bytes, twice the MACs)?  What does alternating the two kinds cost?  How does the step cost scale with N?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/experiments/mma_step_cost.cu -o /tmp/mma_step_cost
//   /tmp/mma_step_cost
// Not part of the library.  Operand contents are irrelevant (zeros); descriptors walk a 4-stage ring of 128-byte-row
// SWIZZLE_128B tiles exactly like conv_tc.cu's main loop (4 steps of 32 bytes per 128-byte row chunk).
#include <cuda.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
template <int CG, int F8>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 1 && !F8) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  if (CG == 1 && F8) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  if (CG == 2 && !F8) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  if (CG == 2 && F8) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done;
}

// mode: 0 = all f16, 1 = all f8, 2 = alternate every 4 steps (chunk-wise), 3 = f8 for the first half then f16,
//       4 = f16 with two independent accumulators (TMEM columns 0 and 256) alternating step by step: are back-to-back
//           steps into ONE accumulator latency-bound (dependent chain) rather than throughput-bound?
//       5 = f16, alternating accumulators per 4-step chunk
template <int CG>
__global__ void __launch_bounds__(128, 1) step_kernel(int N, int mode, int steps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < (4 * 48 * 1024) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_smem;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc16 = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((CG == 2 ? 256 : 128) >> 4) << 24);
    const long long t0 = clock64();
    for (int s = 0; s < steps; ++s) {
      const int stage = (s >> 2) & 3, k = s & 3;
      const uint64_t da = make_sw128_desc(base + stage * 48 * 1024) + 2 * k;            // A: 16 KB, B: up to 32 KB per stage
      const uint64_t db = make_sw128_desc(base + stage * 48 * 1024 + 16 * 1024) + 2 * k;
      const int chunk = s >> 2;
      const bool f8 = mode == 1 || (mode == 2 && (chunk & 1)) || (mode == 3 && s < steps / 2);
      const uint32_t acc_sel = mode == 4 ? (uint32_t)(s & 1) : (mode == 5 ? (uint32_t)(chunk & 1) : 0u);
      const uint32_t d = tmem + acc_sel * 256u;
      const uint32_t accum = (mode >= 4 ? s >= 8 : s > 0) ? 1u : 0u;
      if (f8) mma<CG, 1>(d, da, db, idesc16, accum);
      else mma<CG, 0>(d, da, db, idesc16, accum);
    }
    if (CG == 2) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
    else asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    while (!mbar_try_wait(smem_u32(&bar), 0)) {}
    out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <int CG>
static void run(int N, int mode, int grid, long long* d_out) {
  const int steps = 4096, smem = 4 * 48 * 1024 + 1024;
  cudaFuncSetAttribute(step_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, step_kernel<CG>, N, mode, steps, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CG=%d N=%d mode=%d: CUDA error %s\n", CG, N, mode, cudaGetErrorString(e)); return; }
  long long h[2] = {0, 0};
  cudaMemcpy(h, d_out, sizeof(long long) * (CG == 2 ? 2 : 1), cudaMemcpyDeviceToHost);
  const char* names[6] = {"f16 K=16", "f8 K=32", "alternating kinds per 4-step chunk", "f8 half then f16 half",
                          "f16, 2 accumulators per step", "f16, 2 accumulators per chunk"};
  printf("cta_group::%d grid %3d M=%d N=%3d %-36s %7.1f cycles/step\n", CG, grid, CG == 2 ? 256 : 128, N, names[mode], (double)h[0] / steps);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 1024 * sizeof(long long));
  cudaMemset(d_out, 0, 1024 * sizeof(long long));
  for (int grid : {1, 148}) {
    for (int N : {64, 128, 256})
      for (int mode = 0; mode < 6; ++mode) run<1>(N, mode, grid, d_out);
    for (int N : {64, 128, 256})
      for (int mode = 0; mode < 6; ++mode) run<2>(N, mode, grid == 1 ? 2 : 148, d_out);
  }
  return 0;
}
