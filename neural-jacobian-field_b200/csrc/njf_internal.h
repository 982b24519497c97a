// Internal host-side helpers shared by the translation units of libnjf_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace njf {

// thread-local message returned by njf_last_error()
std::string& last_error();

#define NJF_FAIL(...)                                             \
  do {                                                            \
    char buf_[512];                                               \
    snprintf(buf_, sizeof(buf_), __VA_ARGS__);                    \
    ::njf::last_error() = buf_;                                   \
    return 1;                                                     \
  } while (0)

#define NJF_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t err_ = (expr);                                                          \
    if (err_ != cudaSuccess)                                                            \
      NJF_FAIL("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(err_)); \
  } while (0)

// Pack a row-major fp32 matrix W[n_real][ld] (first k_real columns used) into the fp16
// K-major SWIZZLE_128B shared-memory image the tcgen05 B operand expects:
// k_pad/64 K-blocks of [n_pad rows x 128 B]; 16 B chunk c of row n lands at chunk
// c ^ (n & 7).  Padding rows/columns are zero.  `out` must hold n_pad*k_pad*2 bytes.
void pack_sw128_f16(const float* w, int n_real, int k_real, int ld, int n_pad, int k_pad,
                    uint8_t* out);

// [n_pad x 16] fp16 SWIZZLE_32B block whose column 0 holds bias[0..n_real) (kStepBias); n_pad*32 bytes
void pack_sw32_bias_f16(const float* bias, int n_real, int n_pad, uint8_t* out);

uint16_t f32_to_f16_bits(float f);

// number of kernels this library has launched (njf_debug_launch_count; bench.py's gpu_launches)
void count_launch(int n = 1);

}  // namespace njf
