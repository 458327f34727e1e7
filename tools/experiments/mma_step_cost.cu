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

// MODE: 0 = all f16, 1 = all f8, 2 = kinds alternate every 4-step chunk, 4 = f16 with two independent accumulators (TMEM
// columns 0 / 256) alternating step by step, 5 = the same alternating per chunk.  The issue loop is fully unrolled over
// 16 steps (4 ring stages x 4 K slices) with every descriptor precomputed in registers, so the elected thread spends
// ~2 instructions per MMA: what is measured is the tensor core's step time, not the issue overhead.
// (A first version of this tool computed descriptors inside the loop and measured ~215 cycles for EVERY shape: a lone
// thread retires a dependent instruction every ~10 cycles, i.e. it was issue-bound -- which is itself the lesson for
// conv_tc.cu's narrow-N layers.)
template <int CG, int MODE>
__global__ void __launch_bounds__(128, 1) step_kernel(int N, int iters, long long* out) {
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
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((CG == 2 ? 256 : 128) >> 4) << 24);
    uint64_t da[16], db[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      da[j] = make_sw128_desc(base + (j >> 2) * 48 * 1024) + 2 * (j & 3);             // A: 16 KB, B: up to 32 KB per stage
      db[j] = make_sw128_desc(base + (j >> 2) * 48 * 1024 + 16 * 1024) + 2 * (j & 3);
    }
    mma<CG, 0>(tmem, da[0], db[0], idesc, 0u);
    mma<CG, 0>(tmem + 256u, da[0], db[0], idesc, 0u);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const bool f8 = MODE == 1 || (MODE == 2 && ((j >> 2) & 1));
        const uint32_t d = tmem + ((MODE == 4 ? (j & 1) : (MODE == 5 ? ((j >> 2) & 1) : 0)) ? 256u : 0u);
        if (f8) mma<CG, 1>(d, da[j], db[j], idesc, 1u);
        else mma<CG, 0>(d, da[j], db[j], idesc, 1u);
      }
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

template <int CG, int MODE>
static void run(int N, int grid, long long* d_out) {
  const int iters = 256, smem = 4 * 48 * 1024 + 1024;
  cudaFuncSetAttribute(step_kernel<CG, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, step_kernel<CG, MODE>, N, iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CG=%d N=%d mode=%d: CUDA error %s\n", CG, N, MODE, cudaGetErrorString(e)); return; }
  long long h[2] = {0, 0};
  cudaMemcpy(h, d_out, sizeof(long long) * (CG == 2 ? 2 : 1), cudaMemcpyDeviceToHost);
  const char* names[6] = {"f16 K=16", "f8 K=32", "alternating kinds per 4-step chunk", "-",
                          "f16, 2 accumulators per step", "f16, 2 accumulators per chunk"};
  printf("cta_group::%d grid %3d M=%d N=%3d %-36s %7.1f cycles/step\n", CG, grid, CG == 2 ? 256 : 128, N, names[MODE],
         (double)h[0] / (iters * 16.0));
}

template <int CG>
static void run_all(int N, int grid, long long* d_out) {
  run<CG, 0>(N, grid, d_out); run<CG, 1>(N, grid, d_out); run<CG, 2>(N, grid, d_out);
  if (N <= 256) { run<CG, 4>(N, grid, d_out); run<CG, 5>(N, grid, d_out); }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 1024 * sizeof(long long));
  cudaMemset(d_out, 0, 1024 * sizeof(long long));
  for (int grid : {1, 148}) {
    for (int N : {16, 32, 64, 128, 256}) run_all<1>(N, grid, d_out);
    for (int N : {32, 64, 128, 256}) run_all<2>(N, grid == 1 ? 2 : 148, d_out);
  }
  return 0;
}
