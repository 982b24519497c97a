// hoist_tc_kernel: the per-image 1x1 "convolutions" (all lin_z layers + the feature half of
// jacobian_query_mlp) on tcgen05 tensor cores.
//
//   out[b][px][n] = fp16( sum_c feat[b][c][px] * W[n][c] + bias[n] ),  c < 512
//
// One CTA = 128 pixels (thread = pixel = TMEM lane) x one "job" (a slab of <= 448 output channels
// of one hoisted map).  K = 512 is walked in 8 blocks of 64 channels: the threads read the NCHW
// fp32 features coalesced along pixels, convert to fp16 and write the K-major 128B-swizzled A
// tile; the job's weight slab for that K block (pre-packed by njf_field_create) arrives by one
// cp.async.bulk; A and B are double-buffered so loads overlap the MMAs; accumulators (<= 448 fp32
// columns) stay in TMEM until the bias + fp16 epilogue.
#include "field.h"
#include "njf_internal.h"
#include <atomic>

namespace njf {

constexpr int kHoistMaxN = 448;
constexpr uint32_t kHoistABytes = 128 * 128;           // 128 px x 64 ch fp16
constexpr uint32_t kHoistBBytes = kHoistMaxN * 128;    // N x 64 ch fp16
constexpr uint32_t kHoistSmem = 2 * kHoistABytes + 2 * kHoistBBytes + 256 + 1024;

struct HoistJob {
  __half* out;          // map base [B][HW][CH]
  int CH;               // channel stride of that map
  int c0;               // first output channel of this slab inside the map
  int N;                // slab width (multiple of 16, <= 448)
  uint32_t w_off;       // byte offset of the slab's 8 K-block images inside the hoist blob
  int bias_off;         // offset into the hoist bias vector
};
struct HoistParams {
  const float* feat;    // [B][512][HW] fp32 (NCHW), or
  const __half* feat_h; // [B][HW][512] fp16 (NHWC: what a channels-last half-precision encoder emits; already K-major)
  const uint8_t* wimg;  // packed weight images
  const float* bias;
  int HW;
  HoistJob job[8];
};

template <bool kNhwcHalf>
__global__ void __launch_bounds__(128, 1) hoist_tc_kernel(const __grid_constant__ HoistParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                        // 2 x 16 KB
  uint8_t* sB = smem + 2 * kHoistABytes;     // 2 x 56 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 2 * kHoistBBytes);  // full[2], free[2], done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  const HoistJob& job = p.job[blockIdx.y];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int px = blockIdx.x * 128 + tid;
  const bool live = px < p.HW;
  const float* F = kNhwcHalf ? nullptr : p.feat + static_cast<size_t>(blockIdx.z) * 512 * p.HW + (live ? px : 0);
  // NHWC fp16: the pixel's 512 channels are contiguous, a K block is 8 ready-made 16-byte chunks of the A row
  const uint4* Fh = kNhwcHalf ? reinterpret_cast<const uint4*>(p.feat_h + (static_cast<size_t>(blockIdx.z) * p.HW + (live ? px : 0)) * 512)
                              : nullptr;
  const uint32_t bbytes = static_cast<uint32_t>(job.N) * 128u;
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], bbytes);
    bulk_g2s(sB, p.wimg + job.w_off, bbytes, &bars[0]);
  }
  float f[kNhwcHalf ? 1 : 64];
  uint4 fv[kNhwcHalf ? 8 : 1];
  if constexpr (kNhwcHalf) {
#pragma unroll
    for (int j = 0; j < 8; ++j) fv[j] = live ? __ldg(Fh + j) : make_uint4(0u, 0u, 0u, 0u);
  } else {
#pragma unroll
    for (int c = 0; c < 64; ++c) f[c] = live ? __ldg(F + static_cast<size_t>(c) * p.HW) : 0.f;
  }
  uint32_t full_par[2] = {0, 0}, free_par[2] = {0, 0};
  for (int kb = 0; kb < 8; ++kb) {
    const int buf = kb & 1;
    if (kb >= 2) {  // the MMAs of K block kb-2 have finished reading this buffer pair
      mbar_wait(&bars[2 + buf], free_par[buf]);
      free_par[buf] ^= 1u;
    }
    if (tid == 0 && kb + 1 < 8) {  // prefetch the next weight slab into the other buffer
      const int nb = buf ^ 1;
      if (kb >= 1) {               // ... once the MMAs of K block kb-1 are done with it
        mbar_wait(&bars[2 + nb], free_par[nb]);
        // (parity is consumed again by everyone at iteration kb+1; do not flip it here)
      }
      mbar_arrive_expect_tx(&bars[nb], bbytes);
      bulk_g2s(sB + nb * kHoistBBytes, p.wimg + job.w_off + static_cast<size_t>(kb + 1) * bbytes, bbytes, &bars[nb]);
    }
    // A tile: row = pixel, 64 channels -> 8 swizzled 16 B chunks
    {
      uint8_t* row = sA + buf * kHoistABytes + tid * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 v;
        if constexpr (kNhwcHalf) {
          v = fv[j];
        } else {
          v.x = pack_f16x2(f[8 * j + 0], f[8 * j + 1]);
          v.y = pack_f16x2(f[8 * j + 2], f[8 * j + 3]);
          v.z = pack_f16x2(f[8 * j + 4], f[8 * j + 5]);
          v.w = pack_f16x2(f[8 * j + 6], f[8 * j + 7]);
        }
        *reinterpret_cast<uint4*>(row + ((j ^ (tid & 7)) << 4)) = v;
      }
    }
    if (kb + 1 < 8) {
      if constexpr (kNhwcHalf) {
#pragma unroll
        for (int j = 0; j < 8; ++j) fv[j] = live ? __ldg(Fh + (kb + 1) * 8 + j) : make_uint4(0u, 0u, 0u, 0u);
      } else {
#pragma unroll
        for (int c = 0; c < 64; ++c) f[c] = live ? __ldg(F + static_cast<size_t>((kb + 1) * 64 + c) * p.HW) : 0.f;
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      mbar_wait(&bars[buf], full_par[buf]);
      tc_fence_after();
      const uint32_t a0 = smem_u32(sA + buf * kHoistABytes), b0 = smem_u32(sB + buf * kHoistBBytes);
      for (int n0 = 0; n0 < job.N; n0 += 256) {
        const int n = min(256, job.N - n0);
        const uint32_t idesc = make_idesc_f16(n);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem + n0, make_sw128_desc(a0 + k * 32), make_sw128_desc(b0 + n0 * 128 + k * 32), idesc,
                   (kb > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit(&bars[2 + buf]);
      if (kb == 7) umma_commit(&bars[4]);
    }
    full_par[buf] ^= 1u;
  }
  mbar_wait(&bars[4], 0);
  tc_fence_after();
  // epilogue: this thread's pixel row, 32 channels at a time
  const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  __half* orow = job.out + (static_cast<size_t>(blockIdx.z) * p.HW + (live ? px : 0)) * job.CH + job.c0;
  const float* bias = p.bias + job.bias_off;
  for (int c0 = 0; c0 < job.N; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(trow + c0, r);
    tmem_ld_wait();
    if (live) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 v;
        uint32_t* vw = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
        for (int t = 0; t < 4; ++t)
          vw[t] = pack_f16x2(__uint_as_float(r[8 * j + 2 * t]) + __ldg(bias + c0 + 8 * j + 2 * t),
                             __uint_as_float(r[8 * j + 2 * t + 1]) + __ldg(bias + c0 + 8 * j + 2 * t + 1));
        *reinterpret_cast<uint4*>(orow + c0 + 8 * j) = v;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace njf

using namespace njf;

// host: pack the hoist weight rows [n][512] into per-(job, K-block) SW128 images
int njf_hoist_build(NjfField* f, const std::vector<float>& w, const std::vector<float>& b) {
  struct Slab { int map; int c0; int n; };
  std::vector<Slab> slabs;
  int nmaps = f->desc.n_proposal + 1;
  for (int m = 0; m < nmaps; ++m) {
    const int CH = (m < f->desc.n_proposal) ? f->ch_prop : f->ch_main;
    for (int c0 = 0; c0 < CH;) {
      const int n = (CH - c0 > kHoistMaxN) ? 384 : CH - c0;
      slabs.push_back({m, c0, n});
      c0 += n;
    }
  }
  if (slabs.size() > 8) NJF_FAIL("internal: too many hoist slabs");
  std::vector<uint8_t> blob;
  f->hoist_jobs.clear();
  std::vector<int> map_row0(nmaps);
  for (int m = 0, r = 0; m < nmaps; ++m) {
    map_row0[m] = r;
    r += (m < f->desc.n_proposal) ? f->ch_prop : f->ch_main;
  }
  for (const Slab& s : slabs) {
    NjfField::HoistJobHost j{};
    j.map = s.map;
    j.c0 = s.c0;
    j.N = s.n;
    j.w_off = static_cast<uint32_t>(blob.size());
    j.bias_off = map_row0[s.map] + s.c0;
    for (int kb = 0; kb < 8; ++kb) {
      const size_t off = blob.size();
      blob.resize(off + static_cast<size_t>(s.n) * 128);
      pack_sw128_f16(w.data() + static_cast<size_t>(map_row0[s.map] + s.c0) * 512 + kb * 64, s.n, 64, 512, s.n, 64,
                     blob.data() + off);
    }
    f->hoist_jobs.push_back(j);
  }
  NJF_CUDA(cudaMalloc(&f->d_hoist_img, blob.size()));
  NJF_CUDA(cudaMemcpy(f->d_hoist_img, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  (void)b;
  return 0;
}

int njf_hoist_launch(const NjfField* f, const float* feat_nchw, int B, int Hf, int Wf, void* maps_out,
                     cudaStream_t stream, int view0, int B_total, const void* feat_nhwc_f16) {
  int dev = 0;
  cudaGetDevice(&dev);
  static std::atomic<bool> attr[64];  // the opt-in applies to the current device only
  if (dev >= 0 && dev < 64 && !attr[dev].load(std::memory_order_acquire)) {
    NJF_CUDA(cudaFuncSetAttribute(hoist_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(kHoistSmem)));
    NJF_CUDA(cudaFuncSetAttribute(hoist_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(kHoistSmem)));
    attr[dev].store(true, std::memory_order_release);
  }
  if (B_total < 0) B_total = B;
  HoistParams p{};
  p.feat = feat_nchw;
  p.feat_h = static_cast<const __half*>(feat_nhwc_f16);
  p.wimg = f->d_hoist_img;
  p.bias = f->d_hoist_b;
  p.HW = Hf * Wf;
  __half* base = static_cast<__half*>(maps_out);
  std::vector<__half*> map_ptr;
  for (int m = 0; m <= f->desc.n_proposal; ++m) {  // maps are [level][B_total views][pixel][channel]; write views view0..
    const size_t ch = (m < f->desc.n_proposal) ? f->ch_prop : f->ch_main;
    map_ptr.push_back(base + static_cast<size_t>(view0) * p.HW * ch);
    base += static_cast<size_t>(B_total) * p.HW * ch;
  }
  int nj = 0;
  for (const auto& j : f->hoist_jobs) {
    HoistJob& d = p.job[nj++];
    d.out = map_ptr[j.map];
    d.CH = (j.map < f->desc.n_proposal) ? f->ch_prop : f->ch_main;
    d.c0 = j.c0;
    d.N = j.N;
    d.w_off = j.w_off;
    d.bias_off = j.bias_off;
  }
  dim3 grid((p.HW + 127) / 128, nj, B);
  if (feat_nhwc_f16)
    hoist_tc_kernel<true><<<grid, 128, kHoistSmem, stream>>>(p);
  else
    hoist_tc_kernel<false><<<grid, 128, kHoistSmem, stream>>>(p);
  njf::count_launch();
  NJF_CUDA(cudaGetLastError());
  return 0;
}
