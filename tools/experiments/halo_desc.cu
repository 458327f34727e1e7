// Round-2 experiment: can ONE shared-memory halo tile serve all nine taps of a 3x3 convolution?
//
// The conv kernel stages, per filter tap, a fresh 128-pixel A tile from L2 (9 x 16 KB per 64 channels): every conv of
// the step runs at the ~12 TB/s L2 -> SM cap (profiles/r2_ncu_summary.md).  A tile of 8 x 16 output pixels needs a
// 10 x 18 halo = 180 pixel rows of 128 bytes; tap (ty, tx) is then the view
//       start = base + (ty * 10 + tx) * 128 B,   8-row groups (one image row each) 10 * 128 = 1280 B apart
// i.e. a K-major SWIZZLE_128B descriptor whose start is not 1024-aligned (already used by the 7x7 halo-row scheme) AND
// whose stride-byte-offset is 1280 instead of 1024.  The 128-byte swizzle must then be a pure function of the absolute
// shared-memory address bits (chunk ^= (addr >> 7) & 7) for this to read the right bytes.  This tool fills the halo
// tile with plain stores using exactly that address function (what TMA SWIZZLE_128B writes), runs the nine views
// through tcgen05.mma (kind::f16 and kind::f8f6f4) and compares with the host.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/experiments/halo_desc.cu -o /tmp/halo_desc && /tmp/halo_desc
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int M = 128, N = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done;
}

// rows of 128 bytes at base + r * 128; 16-byte chunk c of a row lands at chunk c ^ ((address >> 7) & 7)
__device__ __forceinline__ void fill_rows(uint8_t* dst, const uint8_t* src, int rows) {
  for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    const uint32_t row_addr = smem_u32(dst) + r * 128;
    const uint4 v = *reinterpret_cast<const uint4*>(src + (size_t)r * 128 + c * 16);
    *reinterpret_cast<uint4*>(dst + r * 128 + ((c ^ ((row_addr >> 7) & 7)) << 4)) = v;
  }
}

// D[tap][128][N] = view_tap(A) * B^T ; f8 != 0: operands are e4m3 (128 K elements per row), else fp16 (64)
__global__ void __launch_bounds__(128, 1) halo_kernel(const uint8_t* A, const uint8_t* B, float* D, int f8, int pitch) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* sm = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  uint8_t* sA = sm; uint8_t* sB = sm + 28 * 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_smem;
  fill_rows(sA, A, 18 * pitch);
  fill_rows(sB, B, N);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  for (int tap = 0; tap < 9; ++tap) {
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int ty = tap / 3, tx = tap % 3;
      const uint64_t da = make_desc(smem_u32(sA) + (uint32_t)(ty * pitch + tx) * 128u, (uint32_t)pitch * 128u);
      const uint64_t db = make_desc(smem_u32(sB), 1024u);
      for (int k = 0; k < 4; ++k) {
        const uint32_t acc = k ? 1u : 0u;
        if (f8) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(da + 2 * k), "l"(db + 2 * k), "r"(idesc), "r"(acc) : "memory");
        else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                          ::"r"(tmem), "l"(da + 2 * k), "l"(db + 2 * k), "r"(idesc), "r"(acc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    while (!mbar_try_wait(smem_u32(&bar), (uint32_t)(tap & 1))) {}
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                   "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                     "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) D[((size_t)tap * M + row) * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main() {
  srand(3);
  int bad_total = 0;
  for (int pitch : {8, 10, 12}) {
    const int rows = 18 * pitch;
    for (int f8 = 0; f8 < 2; ++f8) {
      const int K = f8 ? 128 : 64;
      std::vector<uint8_t> a((size_t)rows * 128), b((size_t)N * 128);
      std::vector<float> af((size_t)rows * K), bf((size_t)N * K);
      auto put = [&](std::vector<uint8_t>& raw, std::vector<float>& f, size_t i) {
        const float v = (float)((rand() % 17) - 8) * 0.25f;           // exactly representable in fp16 and e4m3
        f[i] = v;
        if (f8) { __nv_fp8_e4m3 q(v); raw[i] = q.__x; }
        else { __half h = __float2half_rn(v); reinterpret_cast<__half*>(raw.data())[i] = h; }
      };
      for (size_t i = 0; i < af.size(); ++i) put(a, af, i);
      for (size_t i = 0; i < bf.size(); ++i) put(b, bf, i);
      uint8_t *dA, *dB; float* dD;
      cudaMalloc(&dA, a.size()); cudaMalloc(&dB, b.size()); cudaMalloc(&dD, (size_t)9 * M * N * 4);
      cudaMemcpy(dA, a.data(), a.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, b.data(), b.size(), cudaMemcpyHostToDevice);
      cudaMemset(dD, 0, (size_t)9 * M * N * 4);
      const int smem = 64 * 1024;
      cudaFuncSetAttribute(halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      halo_kernel<<<1, 128, smem>>>(dA, dB, dD, f8, pitch);
      cudaError_t e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("pitch %d f8 %d: CUDA error %s\n", pitch, f8, cudaGetErrorString(e)); return 1; }
      std::vector<float> out((size_t)9 * M * N);
      cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
      for (int tap = 0; tap < 9; ++tap) {
        int bad = 0; double worst = 0;
        for (int r = 0; r < M; ++r) for (int n = 0; n < N; ++n) {
          const int q = (r / 8 + tap / 3) * pitch + (r % 8) + tap % 3;
          double s = 0;
          for (int k = 0; k < K; ++k) s += (double)af[(size_t)q * K + k] * (double)bf[(size_t)n * K + k];
          const double d = fabs(s - (double)out[((size_t)tap * M + r) * N + n]);
          if (d > 1e-3) ++bad;
          worst = fmax(worst, d);
        }
        printf("pitch %2d (SBO %4d B) %s tap (%d,%d): %s  bad %d / %d  max|diff| %.3g\n", pitch, pitch * 128, f8 ? "f8 " : "f16",
               tap / 3, tap % 3, bad ? "MISMATCH" : "ok", bad, M * N, worst);
        bad_total += bad;
      }
      cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
  }
  printf(bad_total ? "RESULT: some views are wrong\n" : "RESULT: every shifted / strided view is exact\n");
  return 0;
}
