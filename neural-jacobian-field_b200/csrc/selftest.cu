// Self-test of the tcgen05 layer-chain machinery (mlp_core.cuh): a miniature of the
// real trunk -- lin_in (K=64) -> x ; x += staged + bias (TMEM write-back) ; fc_0 ->
// net ; fc_1 accumulates onto x ; lin_out (N=16).  Run over several 128-row tiles by a
// persistent grid so the barrier phase tracking and the odd-tail (idle slot 1) path
// are exercised.  tests/test_gpu_selftest.py checks it against a numpy emulation.
#include "mlp_core.cuh"
#include "njf_internal.h"
#include <vector>

namespace njf {

struct SelftestParams {
  Program prog;
  const uint8_t* blob;   // packed weight images (4 layers)
  const float* bias;     // unused by the kernel (biases ride in the weight images)
  const float* a_in;     // [ntiles*128][64]  fp32 inputs (rounded to fp16 in-kernel)
  const float* tz_in;    // [ntiles*128][128] fp32 "gathered" term (rounded to fp16)
  float* x_out;          // [ntiles*128][128]
  float* y_out;          // [ntiles*128][16]
  int ntiles;
};

__global__ void __launch_bounds__(kThreads, 1) selftest_kernel(const __grid_constant__ SelftestParams p) {
  extern __shared__ uint8_t smem_raw[];
  CtaCtx c = cta_setup(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int nitems = (p.ntiles + 1) / 2;
  int my_items = 0;
  for (int it = blockIdx.x; it < nitems; it += gridDim.x) ++my_items;

  if (warp == kLoaderWarp) {
    if ((threadIdx.x & 31) == 0) loader_role(c, p.prog, p.blob, my_items);
  } else if (warp == kIssuerWarp) {
    if ((threadIdx.x & 31) == 0) {
      const int ntiles = p.ntiles;
      const int bx = blockIdx.x, gx = gridDim.x;
      issuer_role(c, p.prog, my_items, [=](int run) {
        const int it = bx + run * gx;
        return (2 * it + 1 < ntiles) ? 2 : 1;
      });
    }
  } else {
    EpiCtx e = epi_ctx(c);
    for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
      const int tile = 2 * it + e.slot;
      if (tile >= p.ntiles) continue;  // whole warpgroup idles; issuer skips this slot
      const size_t grow = static_cast<size_t>(tile) * kRows + e.row;
      // A tile K-block 0 <- fp16(a_in row): each of the row's two threads writes 32 columns
      {
        const float* src = p.a_in + grow * 64 + 32 * e.half;
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_f16x2(src[2 * j], src[2 * j + 1]);
        a_store32(e, 32 * e.half, pk);
      }
      epi_publish(e);  // step 0: x = a * W0^T
      // stage the "gathered" term the way the gather warps do: lane = 4 channels of one row,
      // warp (h, q) covers rows 32q + 16h .. + 16
      {
        const int lane = threadIdx.x & 31, wrow0 = e.q * 32 + 16 * e.half;
        for (int r = 0; r < 16; ++r) {
          const float* src = p.tz_in + (static_cast<size_t>(tile) * kRows + wrow0 + r) * 128 + 4 * lane;
          uint2 v;
          v.x = pack_f16x2(src[0], src[1]);
          v.y = pack_f16x2(src[2], src[3]);
          *reinterpret_cast<uint2*>(e.tz + tz_offset(wrow0 + r, lane >> 1) + (lane & 1) * 8) = v;
        }
        pair_bar(e);
      }
      epi_wait_acc(e);
      for (int c0 = e.col0; c0 < e.col0 + 64; c0 += 32) epi_x_update<true>(e, c0);
      epi_publish(e);  // step 1: net = relu(x) * W1^T
      epi_wait_acc(e);
      for (int c0 = e.col0; c0 < e.col0 + 64; c0 += 32) epi_relu_to_a(e, 128 + c0, c0);
      epi_publish(e);  // step 2: x += relu(net) * W2^T
      epi_wait_acc(e);
      for (int c0 = e.col0; c0 < e.col0 + 64; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(e.tmem + c0, r);
        tmem_ld_wait();
        float* dst = p.x_out + grow * 128 + c0;
#pragma unroll
        for (int j = 0; j < 32; ++j) dst[j] = __uint_as_float(r[j]);
      }
      for (int c0 = e.col0; c0 < e.col0 + 64; c0 += 32) epi_relu_to_a(e, c0, c0);
      epi_publish(e);  // step 3: y = relu(x) * W3^T (N=16)
      epi_wait_acc(e);
      if (e.half == 0) {
        uint32_t r[16];
        tmem_ld16(e.tmem + 128, r);
        tmem_ld_wait();
        float* dst = p.y_out + grow * 16;
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[j] = __uint_as_float(r[j]);
      }
      pair_bar(e);  // staging rows are rewritten by the partner warp in the next item
    }
  }
  cta_teardown(c);
}

}  // namespace njf

using namespace njf;

// C-ABI: see include/njf_b200.h
extern "C" int njf_selftest_chain(const float* w0, const float* w1, const float* w2,
                                  const float* w3, const float* bias, const float* a_in,
                                  const float* tz_in, float* x_out, float* y_out, int ntiles,
                                  int grid, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // inputs are HOST pointers for the weights/bias, DEVICE pointers for a_in/tz_in/x_out/y_out
  SelftestParams p{};
  std::vector<uint8_t> blob;
  // step s carries bias row `bias + 128*(s)` through the tensor core (kStepBias), the way the
  // render kernels do: x = a W0^T + b0 ; v = x + tz ; net = relu(v) W1^T + b1 ; x += relu(net) W2^T + b2
  auto add = [&](const float* w, const float* b, int n_real, int k_real, int n_pad, int k_pad, int d_col, int acc) {
    MmaStep st{};
    st.w_off = static_cast<uint32_t>(blob.size());
    st.w_bytes = static_cast<uint32_t>(n_pad * k_pad * 2);
    st.n = static_cast<uint16_t>(n_pad);
    st.kblocks = static_cast<uint8_t>(k_pad / 64);
    st.acc = static_cast<uint8_t>(acc);
    st.d_col = static_cast<uint16_t>(d_col);
    blob.resize(blob.size() + st.w_bytes);
    pack_sw128_f16(w, n_real, k_real, k_real, n_pad, k_pad, blob.data() + st.w_off);
    if (b) {
      st.flags |= kStepBias;
      const size_t off = blob.size();
      blob.resize(off + static_cast<size_t>(n_pad) * 32);
      pack_sw32_bias_f16(b, n_real, n_pad, blob.data() + off);
      st.w_bytes += static_cast<uint32_t>(n_pad * 32);
    }
    p.prog.steps[p.prog.nsteps++] = st;
  };
  add(w0, bias, 128, 64, 128, 64, 0, 0);
  add(w1, bias + 128, 128, 128, 128, 128, 128, 0);
  add(w2, bias + 256, 128, 128, 128, 128, 0, 1);
  add(w3, nullptr, 16, 128, 16, 128, 128, 0);
  uint8_t* d_blob = nullptr;
  float* d_bias = nullptr;
  NJF_CUDA(cudaMalloc(&d_blob, blob.size()));
  NJF_CUDA(cudaMalloc(&d_bias, 3 * 128 * sizeof(float)));
  NJF_CUDA(cudaMemcpyAsync(d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice, stream));
  NJF_CUDA(cudaMemcpyAsync(d_bias, bias, 3 * 128 * sizeof(float), cudaMemcpyHostToDevice, stream));
  p.blob = d_blob;
  p.bias = d_bias;
  p.a_in = a_in;
  p.tz_in = tz_in;
  p.x_out = x_out;
  p.y_out = y_out;
  p.ntiles = ntiles;
  const size_t smem = SmemMap::kScratch + 1024;
  NJF_CUDA(cudaFuncSetAttribute(selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  selftest_kernel<<<grid, kThreads, smem, stream>>>(p);
  NJF_CUDA(cudaGetLastError());
  NJF_CUDA(cudaStreamSynchronize(stream));
  cudaFree(d_blob);
  cudaFree(d_bias);
  return 0;
}
