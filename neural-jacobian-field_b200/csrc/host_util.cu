// Host-side utilities of libnjf_b200.so: error string, fp16 conversion, weight-image packer.
#include <cuda_fp16.h>

#include <atomic>
#include <cstring>

#include "njf_internal.h"
#include "ptx.cuh"

namespace njf {

std::string& last_error() {
  static thread_local std::string s;
  return s;
}

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

uint16_t f32_to_f16_bits(float f) {
  const __half h = __float2half_rn(f);
  uint16_t b;
  std::memcpy(&b, &h, 2);
  return b;
}

void pack_sw128_f16(const float* w, int n_real, int k_real, int ld, int n_pad, int k_pad,
                    uint8_t* out) {
  std::memset(out, 0, static_cast<size_t>(n_pad) * k_pad * 2);
  const uint32_t kb_stride = static_cast<uint32_t>(n_pad) * 128u;
  for (int n = 0; n < n_real; ++n) {
    for (int k = 0; k < k_real; ++k) {
      const uint16_t b = f32_to_f16_bits(w[static_cast<size_t>(n) * ld + k]);
      std::memcpy(out + sw128_offset(n, k, kb_stride), &b, 2);
    }
  }
}

void pack_sw32_bias_f16(const float* bias, int n_real, int n_pad, uint8_t* out) {
  std::memset(out, 0, static_cast<size_t>(n_pad) * 32);
  for (int n = 0; n < n_real; ++n) {
    const uint16_t b = f32_to_f16_bits(bias[n]);
    std::memcpy(out + sw32_offset(n, 0), &b, 2);
  }
}

}  // namespace njf

extern "C" const char* njf_last_error(void) { return njf::last_error().c_str(); }
extern "C" int njf_version(void) { return 100; }

extern "C" long long njf_debug_launch_count(int reset) {
  const long long v = njf::g_launches.load(std::memory_order_relaxed);
  if (reset) njf::g_launches.store(0, std::memory_order_relaxed);
  return v;
}
