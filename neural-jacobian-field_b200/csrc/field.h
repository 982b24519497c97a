// NjfField: device-resident packed weights + step programs (built by field.cu, consumed by render.cu).
#pragma once
#include "../../include/njf_b200.h"
#include "mlp_core.cuh"

namespace njf {

// fp32 side tables.  They travel BY VALUE inside the kernel parameter structs (constant bank:
// every epilogue access is warp-uniform, so it is a broadcast constant-cache read instead of an
// L1-thrashed global load).
// All other biases ride inside the weight images (kStepBias) and are accumulated by the tensor core.
struct TrunkTab {
  float4 e0[128];       // (W_in[:,60], W_in[:,61], W_in[:,62], b_in): raw-xyz columns kept in fp32
};
struct XfLayerTab {
  float ln1_g[64], ln1_b[64], ln2_g[64], ln2_b[64];
};
struct HeadTab {
  float4 q_e0[64];      // (Wq[:,60..62], bq)
  XfLayerTab layer[3];
};
struct ColorTab {
  float w3[3 * 64], b3[4];   // last colour layer (3 outputs) stays on the fp32 pipes
};

}  // namespace njf

#include <vector>

struct NjfField {
  struct HoistJobHost { int map, c0, N; uint32_t w_off; int bias_off; };
  NjfFieldDesc desc;
  uint8_t* d_hoist_img = nullptr;          // tcgen05 weight images of the hoist GEMM (hoist_tc.cu)
  mutable float* d_scratch = nullptr;      // grow-only per-field scratch (proposal weights between the
  mutable size_t scratch_bytes = 0;        // proposal kernel and the PDF kernel); one stream at a time
  std::vector<HoistJobHost> hoist_jobs;
  uint8_t* d_blob = nullptr;   // all layer images
  float* d_hoist_w = nullptr;  // [ch_total][512] rows of the hoisted linear maps
  float* d_hoist_b = nullptr;  // [ch_total]
  int ch_prop = 384;           // channels of each proposal map
  int ch_main = 0;             // 448 (transformer: 3x128 + 64 query) or 768 (MLP: 3x128 + 3x128)
  int ch_total = 0;
  njf::Program prop_prog[NJF_MAX_LEVELS];
  const uint8_t* prop_blob[NJF_MAX_LEVELS];
  njf::TrunkTab prop_trunk[NJF_MAX_LEVELS];
  njf::Program field_prog;
  const uint8_t* field_blob = nullptr;
  njf::TrunkTab dens_trunk, jac_trunk;
  njf::HeadTab head;
  njf::ColorTab color;
};

// hoist_tc.cu
int njf_hoist_build(NjfField* f, const std::vector<float>& w, const std::vector<float>& b);
int njf_hoist_launch(const NjfField* f, const float* feat_nchw, int B, int Hf, int Wf, void* maps_out,
                     cudaStream_t stream);
