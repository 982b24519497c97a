// NjfField: device-resident packed weights + step programs (built by field.cu, consumed by render.cu).
#pragma once
#include "../../include/njf_b200.h"
#include "mlp_core.cuh"

namespace njf {

// fp32 side table.  It travels BY VALUE inside the kernel parameter struct (constant bank: every access is
// warp-uniform, so it is a broadcast constant-cache read).  All biases ride inside the weight images
// (kStepBias) and are accumulated by the tensor core.
struct ColorTab {
  float w3[3 * 64], b3[4];   // last colour layer (3 outputs) stays on the fp32 pipes
};


// ---- cross-attention head kernel (xf_head.cu)
constexpr uint32_t kXfMaxBlob = 125 * 1024;  // all head layer images, resident in shared memory
struct XfParams {
  Program prog;        // filled by njf_xf_launch: 12 layer steps + jacobian_head; MmaStep::acc = accumulate onto x
  const uint8_t* blob;
  uint32_t blob_bytes;
  int A;
  // tiling of the pass: identical to field_kernel's (render.cuh PassGeom)
  int NR, S, G, T, NG, group0;
  const uint4* qs;     // [tile][8 chunks][128 rows] x 8 fp16: the 64-wide query embedding of every row
  const float* wts;    // [tile][128 rows] transmittance weight of the sample (0 for padding rows)
  float* jbar;         // [NR][3A] or null
  float* jac_out;      // [NR*S][3A] or null
};

}  // namespace njf

#include <vector>

struct NjfField {
  struct HoistJobHost { int map, c0, N; uint32_t w_off; int bias_off; };
  NjfFieldDesc desc;
  uint8_t* d_hoist_img = nullptr;          // tcgen05 weight images of the hoist GEMM (hoist_tc.cu)
  int q_enc_step = -1;                     // index of the q_enc step in field_prog (transformer head)
  njf::Program head_prog;                  // transformer head: steps of xf_kernel
  uint8_t* d_xf_blob = nullptr;
  uint32_t xf_bytes = 0;
  std::vector<HoistJobHost> hoist_jobs;
  uint8_t* d_blob = nullptr;   // all layer images
  float* d_hoist_w = nullptr;  // [ch_total][512] rows of the hoisted linear maps
  float* d_hoist_b = nullptr;  // [ch_total]
  int ch_prop = 384;           // channels of each proposal map
  int ch_main = 0;             // 448 (transformer: 3x128 + 64 query) or 768 (MLP: 3x128 + 3x128)
  int ch_total = 0;
  njf::Program prop_prog[NJF_MAX_LEVELS];
  const uint8_t* prop_blob[NJF_MAX_LEVELS];
  njf::Program field_prog;
  const uint8_t* field_blob = nullptr;
  njf::ColorTab color;
};

// hoist_tc.cu
int njf_hoist_build(NjfField* f, const std::vector<float>& w, const std::vector<float>& b);
int njf_hoist_launch(const NjfField* f, const float* feat_nchw, int B, int Hf, int Wf, void* maps_out,
                     cudaStream_t stream, int view0 = 0, int B_total = -1, const void* feat_nhwc_f16 = nullptr);
// xf_head.cu (prog / blob / A are taken from the field)
int njf_xf_launch(const NjfField* f, const njf::XfParams& params, cudaStream_t stream);
