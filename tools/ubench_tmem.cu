// Micro-benchmarks that size the epilogue model of DESIGN.md section 4 (run on a B200):
//   1. tcgen05.ld throughput per SM for 4 / 8 / 16 concurrently loading warps (32x32b.x32, wait per load
//      or per two loads)
//   2. round-trip latency  epilogue-arrive -> issuer wake -> tcgen05.mma (N=128,K=128 or N=64,K=64) ->
//      tcgen05.commit -> epilogue wake
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neural-jacobian-field_b200/csrc \
//        -o gpurun_out/ubench_tmem tools/ubench_tmem.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"

using namespace njf;

__global__ void __launch_bounds__(512, 1) ld_bw(int nwarps, int per_wait, int iters, long long* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t0 = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long c0 = clock64();
  if (warp < nwarps) {
    for (int i = 0; i < iters; ++i) {
      uint32_t r[32], s[32];
      tmem_ld32(t0 + ((i * 32) & 255), r);
      if (per_wait == 2) tmem_ld32(t0 + ((i * 32 + 256) & 511), s);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= r[j];
      if (per_wait == 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= s[j];
      }
    }
  }
  __syncthreads();
  const long long c1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = c1 - c0;
  if (acc == 0x12345678u) out[1] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void wait_mode(uint64_t* bar, uint32_t parity, bool spin) {
  if (spin) {
    while (!mbar_test_wait(bar, parity)) {
    }
  } else {
    mbar_wait(bar, parity);
  }
}

// one epilogue warpgroup (128 threads) + one issuer thread; A/B tiles are zero-filled smem
__global__ void __launch_bounds__(160, 1) roundtrip(int n, int kblocks, int iters, int epi_work, int mode, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;             // 32 KB
  uint8_t* sB = smem + 32768;     // 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 128);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  const long long c0 = clock64();
  if (warp == 4) {
    if ((threadIdx.x & 31) == 0) {
      const uint32_t idesc = make_idesc_f16(n);
      uint32_t par = 0;
      for (int it = 0; it < ((mode & 8) ? 0 : iters); ++it) {
        wait_mode(&bars[0], par, mode & 1);
        par ^= 1u;
        tc_fence_after();
        for (int kb = 0; kb < kblocks; ++kb)
          for (int k = 0; k < 4; ++k)
            umma_f16(tm, make_sw128_desc(smem_u32(sA) + kb * 16384 + k * 32),
                     make_sw128_desc(smem_u32(sB) + kb * n * 128 + k * 32), idesc, (kb | k) ? 1u : 0u);
        umma_commit(&bars[1]);
      }
    }
  } else {
    uint32_t par = 0, acc = 0;
    const uint32_t t0 = tm + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t idesc = make_idesc_f16(n);
    for (int it = 0; it < iters; ++it) {
      if (!(mode & 4)) fence_proxy_async_smem();
      tc_fence_before();
      if (mode & 8) {  // no issuer thread: the epilogue group meets at a named barrier, its first thread issues
        named_bar_sync(1, 128);
        if (threadIdx.x == 0) {
          tc_fence_after();
          for (int kb = 0; kb < kblocks; ++kb)
            for (int k = 0; k < 4; ++k)
              umma_f16(tm, make_sw128_desc(smem_u32(sA) + kb * 16384 + k * 32),
                       make_sw128_desc(smem_u32(sB) + kb * n * 128 + k * 32), idesc, (kb | k) ? 1u : 0u);
          umma_commit(&bars[1]);
        }
      } else {
        mbar_arrive(&bars[0]);
      }
      wait_mode(&bars[1], par, mode & 2);
      par ^= 1u;
      tc_fence_after();
      for (int w = 0; w < epi_work; ++w) {
        uint32_t r[32];
        tmem_ld32(t0 + 32 * w, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= r[j];
      }
    }
    if (acc == 0x12345678u) out[1] = acc;
  }
  __syncthreads();
  const long long c1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = c1 - c0;
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tm, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  long long h[2];
  const int iters = 4096;
  for (int per_wait = 1; per_wait <= 2; ++per_wait)
    for (int nw : {1, 4, 8, 16}) {
      ld_bw<<<148, 512>>>(nw, per_wait, iters, d);
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      const double bytes = static_cast<double>(nw) * iters * per_wait * 4096.0;
      printf("ld_bw warps=%2d loads/wait=%d : %8lld cyc  %.1f B/clk/SM  %.0f cyc/load/warp\n", nw, per_wait, h[0],
             bytes / h[0], static_cast<double>(h[0]) / (iters * per_wait));
    }
  cudaFuncSetAttribute(roundtrip, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  for (int mode : {0, 8, 10})
    for (int work : {0, 2})
      for (auto nk : {std::pair<int, int>{128, 2}, {16, 1}}) {
        roundtrip<<<148, 160, 70000>>>(nk.first, nk.second, 2000, work, mode, d);
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("roundtrip mode=%d (issuer-spin=%d epi-spin=%d no-proxy-fence=%d self-issue=%d) N=%3d K=%3d tmem_ld_chunks=%d : %.0f cyc/iter\n", mode,
               mode & 1, (mode >> 1) & 1, (mode >> 2) & 1, (mode >> 3) & 1, nk.first, nk.second * 64, work, h[0] / 2000.0);
      }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
