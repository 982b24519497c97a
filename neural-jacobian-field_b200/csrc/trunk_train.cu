// Training path of the ResnetFC trunks: the kernels behind the perception phase (density head,
// colour head, proposal networks, gradient into the encoder's feature map) and behind the MLP Jacobian head of the
// action phase.  The fused tcgen05 kernels of render.cu keep no activations, so training runs the trunks layer by
// layer in fp32 with the activations in HBM (training batches are a few thousand rays; 12 x M x 128 floats), and
// torch.autograd only does the book-keeping between these kernels (njf_b200/train_trunk.py):
//
//   tt_setup_kernel    world point -> context-camera point -> NeRFEncoding(63) + the four bilinear taps
//                      (pixel_aligned_features.py:11-35, geometry.py:59-65, 137-154; the same row_setup the render
//                      kernels use, so the taps are the ones the forward kernels gather with)
//   tt_gather_kernel   z[m][c] = sum_t w_t map[pix_t][c]       on the fp32 pixel-major lin_z maps (lin_z hoisted onto
//   tt_scatter_kernel  dmap[pix_t][c] += w_t g[m][c]            the feature map: lin_z(bilinear(f)) = bilinear(lin_z(f)))
//   tt_gemm_kernel     C = mask(act(A) . B + bias) + residual  forward  (B = W^T, act = ReLU on the layer input) and
//                                                               input-gradient (B = W, mask = ReLU' of the saved input)
//   tt_wgrad_kernel    dW += gY^T . act(X), db += sum_m gY     reduction over the sample rows, atomics into dW
//   tt_sh16_kernel     SH-16 of the view direction (action_decoder_jacobian.py:24-30, 284)
//
// All fp32 SIMT (register-tiled 8x8 per thread, operands staged in shared memory): gradients are exact to fp32
// round-off against autograd through the reference formulation.  Rows = samples; every matrix is row-major and
// contiguous, inner dimensions are multiples of 4 and at most 128 (the host pads 63 -> 64, 31 -> 32, 3 -> 4, ...).
#include <cuda_runtime.h>

#include "../../include/njf_b200.h"
#include "njf_internal.h"
#include "render.cuh"

namespace njf {
namespace {

constexpr int kTtThreads = 256;
constexpr int kTtBM = 128;   // sample rows per GEMM tile
constexpr int kTtLd = 132;   // shared-memory row pitch in floats (16-byte aligned, conflict-free transposing stores)

int tt_sms() {
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}

// ----------------------------------------------------------------------------- per-sample set-up
struct TtSetup {
  PassGeom g;
  float* enc;     // [M][64]
  int* tap_pix;   // [M][4]
  float* tap_w;   // [M][4]
};

__global__ void __launch_bounds__(128) tt_setup_kernel(const __grid_constant__ TtSetup p) {
  const int m = blockIdx.x * 128 + threadIdx.x;
  if (m >= p.g.NR) return;
  RowState rs;
  row_setup(p.g, blockIdx.x, 0, threadIdx.x, rs);
  // NeRFEncoding: [sin(2 pi x 2^k) dim-major / freq-minor | sin(. + pi/2) | x]; the argument is the same fp32
  // number the reference hands to torch.sin
  float4* out = reinterpret_cast<float4*>(p.enc + static_cast<size_t>(m) * 64);
  const float base[3] = {__fmul_rn(6.2831855f, rs.cam[0]), __fmul_rn(6.2831855f, rs.cam[1]),
                         __fmul_rn(6.2831855f, rs.cam[2])};
  float v[64];
#pragma unroll
  for (int j = 0; j < 30; ++j) {
    const int i = j / 10, k = j - 10 * i;
    const float t = base[i] * static_cast<float>(1 << k);
    v[j] = sinf(t);
    v[30 + j] = sinf(__fadd_rn(t, 1.5707964f));
  }
  v[60] = rs.cam[0];
  v[61] = rs.cam[1];
  v[62] = rs.cam[2];
  v[63] = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) out[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  // bilinear taps, F.grid_sample(align_corners=True, padding_mode="border") on the clamped coordinates
  const float x0 = floorf(rs.ix), y0 = floorf(rs.iy);
  const float x1 = x0 + 1.f, y1 = y0 + 1.f;
  const int xi = static_cast<int>(x0), yi = static_cast<int>(y0);
  const int xj = min(xi + 1, p.g.Wf - 1), yj = min(yi + 1, p.g.Hf - 1);  // out-of-range taps carry weight 0
  reinterpret_cast<float4*>(p.tap_w)[m] = make_float4((x1 - rs.ix) * (y1 - rs.iy), (rs.ix - x0) * (y1 - rs.iy),
                                                      (x1 - rs.ix) * (rs.iy - y0), (rs.ix - x0) * (rs.iy - y0));
  reinterpret_cast<int4*>(p.tap_pix)[m] = make_int4(rs.pixbase + yi * p.g.Wf + xi, rs.pixbase + yi * p.g.Wf + xj,
                                                    rs.pixbase + yj * p.g.Wf + xi, rs.pixbase + yj * p.g.Wf + xj);
}

// ----------------------------------------------------------------------------- gather / scatter on the lin_z maps
// one warp per sample row; a lane owns 4 consecutive channels of every 128-channel group
__global__ void __launch_bounds__(256) tt_gather_kernel(const float* __restrict__ map, const int* __restrict__ tap_pix,
                                                        const float* __restrict__ tap_w, int M, int CH, int ch0, int CW,
                                                        float* __restrict__ out) {
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= M) return;
  const int4 px = __ldg(reinterpret_cast<const int4*>(tap_pix) + m);
  const float4 w = __ldg(reinterpret_cast<const float4*>(tap_w) + m);
  const int pix[4] = {px.x, px.y, px.z, px.w};
  const float wt[4] = {w.x, w.y, w.z, w.w};
  for (int c = lane * 4; c < CW; c += 128) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(map + static_cast<size_t>(pix[t]) * CH + ch0 + c));
      acc.x = fmaf(wt[t], v.x, acc.x);
      acc.y = fmaf(wt[t], v.y, acc.y);
      acc.z = fmaf(wt[t], v.z, acc.z);
      acc.w = fmaf(wt[t], v.w, acc.w);
    }
    *reinterpret_cast<float4*>(out + static_cast<size_t>(m) * CW + c) = acc;
  }
}

__global__ void __launch_bounds__(256) tt_scatter_kernel(const float* __restrict__ g, const int* __restrict__ tap_pix,
                                                         const float* __restrict__ tap_w, int M, int CH, int ch0, int CW,
                                                         float* __restrict__ dmap) {
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= M) return;
  const int4 px = __ldg(reinterpret_cast<const int4*>(tap_pix) + m);
  const float4 w = __ldg(reinterpret_cast<const float4*>(tap_w) + m);
  const int pix[4] = {px.x, px.y, px.z, px.w};
  const float wt[4] = {w.x, w.y, w.z, w.w};
  for (int c = lane * 4; c < CW; c += 128) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g + static_cast<size_t>(m) * CW + c));
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (wt[t] == 0.f) continue;
      atomicAdd(reinterpret_cast<float4*>(dmap + static_cast<size_t>(pix[t]) * CH + ch0 + c),
                make_float4(wt[t] * v.x, wt[t] * v.y, wt[t] * v.z, wt[t] * v.w));
    }
  }
}

// ----------------------------------------------------------------------------- C = mask(act(A) . B + bias) + residual
struct TtGemm {
  const float* a;         // [M][KR]
  const float* w;         // trans_w ? [NO][KR] (a Linear weight used forward) : [KR][NO] (the same weight, backward)
  const float* bias;      // [NO] or null
  const float* residual;  // [M][NO] or null
  const float* mask_src;  // [M][NO] or null: the result is zeroed where mask_src <= 0 (ReLU' of the saved input)
  float* c;               // [M][NO]
  int M, NO, KR, trans_w, relu_in;
};

template <int NG>  // NG groups of 64 output columns
__global__ void __launch_bounds__(kTtThreads, 1) tt_gemm_kernel(const __grid_constant__ TtGemm p) {
  extern __shared__ __align__(16) float tt_smem[];
  float* As = tt_smem;                 // [KR][kTtLd]: the A tile transposed (k-major)
  float* Bs = tt_smem + 128 * kTtLd;   // [KR][kTtLd]
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int NO = p.NO, KR = p.KR;
  constexpr int NOp = 64 * NG;
  if (p.trans_w) {
    for (int i = tid; i < NOp * (KR / 4); i += kTtThreads) {
      const int no = i % NOp, k4 = i / NOp;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (no < NO) v = __ldg(reinterpret_cast<const float4*>(p.w + static_cast<size_t>(no) * KR + k4 * 4));
      Bs[(k4 * 4 + 0) * kTtLd + no] = v.x;
      Bs[(k4 * 4 + 1) * kTtLd + no] = v.y;
      Bs[(k4 * 4 + 2) * kTtLd + no] = v.z;
      Bs[(k4 * 4 + 3) * kTtLd + no] = v.w;
    }
  } else {
    for (int i = tid; i < KR * (NOp / 4); i += kTtThreads) {
      const int kr = i / (NOp / 4), no = (i % (NOp / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (no < NO) v = __ldg(reinterpret_cast<const float4*>(p.w + static_cast<size_t>(kr) * NO + no));
      *reinterpret_cast<float4*>(Bs + kr * kTtLd + no) = v;
    }
  }
  const int ntiles = (p.M + kTtBM - 1) / kTtBM;
  const int k4_per_half = (KR / 4 + 1) / 2;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();  // the previous tile's readers are done with As (and Bs is complete on the first pass)
    {
      const int row = tid & 127, half = tid >> 7;
      const int m = tile * kTtBM + row;
      const int k4_end = min(KR / 4, (half + 1) * k4_per_half);
      for (int k4 = half * k4_per_half; k4 < k4_end; ++k4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < p.M) v = __ldg(reinterpret_cast<const float4*>(p.a + static_cast<size_t>(m) * KR + k4 * 4));
        if (p.relu_in) {
          v.x = fmaxf(v.x, 0.f);
          v.y = fmaxf(v.y, 0.f);
          v.z = fmaxf(v.z, 0.f);
          v.w = fmaxf(v.w, 0.f);
        }
        As[(k4 * 4 + 0) * kTtLd + row] = v.x;
        As[(k4 * 4 + 1) * kTtLd + row] = v.y;
        As[(k4 * 4 + 2) * kTtLd + row] = v.z;
        As[(k4 * 4 + 3) * kTtLd + row] = v.w;
      }
    }
    __syncthreads();
    float acc[8][4 * NG];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4 * NG; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int kr = 0; kr < KR; ++kr) {
      const float4 a0 = *reinterpret_cast<const float4*>(As + kr * kTtLd + ty * 4);
      const float4 a1 = *reinterpret_cast<const float4*>(As + kr * kTtLd + 64 + ty * 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[4 * NG];
      {
        const float4 b0 = *reinterpret_cast<const float4*>(Bs + kr * kTtLd + tx * 4);
        bv[0] = b0.x, bv[1] = b0.y, bv[2] = b0.z, bv[3] = b0.w;
        if constexpr (NG == 2) {
          const float4 b1 = *reinterpret_cast<const float4*>(Bs + kr * kTtLd + 64 + tx * 4);
          bv[4] = b1.x, bv[5] = b1.y, bv[6] = b1.z, bv[7] = b1.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4 * NG; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = tile * kTtBM + (i >> 2) * 64 + ty * 4 + (i & 3);
      if (m >= p.M) continue;
#pragma unroll
      for (int gq = 0; gq < NG; ++gq) {
        const int no = gq * 64 + tx * 4;
        if (no >= NO) continue;
        float4 v = make_float4(acc[i][4 * gq], acc[i][4 * gq + 1], acc[i][4 * gq + 2], acc[i][4 * gq + 3]);
        if (p.bias) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + no));
          v.x += b.x, v.y += b.y, v.z += b.z, v.w += b.w;
        }
        const size_t o = static_cast<size_t>(m) * NO + no;
        if (p.mask_src) {
          const float4 s = __ldg(reinterpret_cast<const float4*>(p.mask_src + o));
          v.x = s.x > 0.f ? v.x : 0.f;
          v.y = s.y > 0.f ? v.y : 0.f;
          v.z = s.z > 0.f ? v.z : 0.f;
          v.w = s.w > 0.f ? v.w : 0.f;
        }
        if (p.residual) {
          const float4 r = __ldg(reinterpret_cast<const float4*>(p.residual + o));
          v.x += r.x, v.y += r.y, v.z += r.z, v.w += r.w;
        }
        *reinterpret_cast<float4*>(p.c + o) = v;
      }
    }
  }
}

// ----------------------------------------------------------------------------- dW += gY^T . act(X), db += sum gY
struct TtWgrad {
  const float* gy;  // [M][N]
  const float* x;   // [M][K]
  float* gw;        // [N][K], accumulated with atomics
  float* gb;        // [N] or null
  int M, N, K, relu_in, rows_per_cta;
};

template <bool N2, bool K2>  // a second group of 64 output rows (n >= 64) / columns (k >= 64)
__global__ void __launch_bounds__(kTtThreads, 2) tt_wgrad_kernel(const __grid_constant__ TtWgrad p) {
  extern __shared__ __align__(16) float tt_smem[];
  float* Ys = tt_smem;              // [64][128]
  float* Xs = tt_smem + 64 * 128;   // [64][128]
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int N = p.N, K = p.K;
  for (int i = tid; i < 2 * 64 * 128 / 4; i += kTtThreads)
    reinterpret_cast<float4*>(tt_smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);  // columns past N / K stay zero
  float acc[N2 ? 8 : 4][K2 ? 8 : 4];
  float by[N2 ? 8 : 4];
#pragma unroll
  for (int i = 0; i < (N2 ? 8 : 4); ++i) {
    by[i] = 0.f;
#pragma unroll
    for (int j = 0; j < (K2 ? 8 : 4); ++j) acc[i][j] = 0.f;
  }
  const int m_begin = blockIdx.x * p.rows_per_cta, m_end = min(p.M, m_begin + p.rows_per_cta);
  for (int m0 = m_begin; m0 < m_end; m0 += 64) {
    __syncthreads();
    for (int i = tid; i < 64 * (N / 4); i += kTtThreads) {
      const int r = i / (N / 4), c4 = i % (N / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < m_end) v = __ldg(reinterpret_cast<const float4*>(p.gy + static_cast<size_t>(m0 + r) * N + c4 * 4));
      *reinterpret_cast<float4*>(Ys + r * 128 + c4 * 4) = v;
    }
    for (int i = tid; i < 64 * (K / 4); i += kTtThreads) {
      const int r = i / (K / 4), c4 = i % (K / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < m_end) v = __ldg(reinterpret_cast<const float4*>(p.x + static_cast<size_t>(m0 + r) * K + c4 * 4));
      if (p.relu_in) {
        v.x = fmaxf(v.x, 0.f);
        v.y = fmaxf(v.y, 0.f);
        v.z = fmaxf(v.z, 0.f);
        v.w = fmaxf(v.w, 0.f);
      }
      *reinterpret_cast<float4*>(Xs + r * 128 + c4 * 4) = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < 64; ++r) {
      float yv[N2 ? 8 : 4], xv[K2 ? 8 : 4];
      {
        const float4 y0 = *reinterpret_cast<const float4*>(Ys + r * 128 + ty * 4);
        yv[0] = y0.x, yv[1] = y0.y, yv[2] = y0.z, yv[3] = y0.w;
        if constexpr (N2) {
          const float4 y1 = *reinterpret_cast<const float4*>(Ys + r * 128 + 64 + ty * 4);
          yv[4] = y1.x, yv[5] = y1.y, yv[6] = y1.z, yv[7] = y1.w;
        }
        const float4 x0 = *reinterpret_cast<const float4*>(Xs + r * 128 + tx * 4);
        xv[0] = x0.x, xv[1] = x0.y, xv[2] = x0.z, xv[3] = x0.w;
        if constexpr (K2) {
          const float4 x1 = *reinterpret_cast<const float4*>(Xs + r * 128 + 64 + tx * 4);
          xv[4] = x1.x, xv[5] = x1.y, xv[6] = x1.z, xv[7] = x1.w;
        }
      }
#pragma unroll
      for (int i = 0; i < (N2 ? 8 : 4); ++i) {
        by[i] += yv[i];
#pragma unroll
        for (int j = 0; j < (K2 ? 8 : 4); ++j) acc[i][j] = fmaf(yv[i], xv[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < (N2 ? 8 : 4); ++i) {
    const int n = (i >> 2) * 64 + ty * 4 + (i & 3);
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < (K2 ? 8 : 4); ++j) {
      const int k = (j >> 2) * 64 + tx * 4 + (j & 3);
      if (k < K) atomicAdd(p.gw + static_cast<size_t>(n) * K + k, acc[i][j]);
    }
    if (tx == 0 && p.gb) atomicAdd(p.gb + n, by[i]);
  }
}


// ============================================================================= TF32 tensor-core variants
// The same two GEMM shapes on tcgen05 (kind::tf32, fp32 accumulators in tensor memory), selected by the caller when
// torch's float32 matmul precision is not "highest" -- the reference trains with
// torch.set_float32_matmul_precision("high") (train.py:64-65), i.e. its nn.Linear layers run TF32 too.  Operands are
// staged from global memory by the CTA's threads straight into the K-major SWIZZLE_128B shared-memory image
// (rounded to TF32 with cvt.rna), one elected thread issues the MMAs, and both kernels are software pipelines over
// 128-row (64-row for the weight gradient) sample tiles so that the loads of tile j+1 overlap the MMAs of tile j;
// with K, N <= 128 they are bound by the HBM traffic of the activations, not by the tensor pipe.
__device__ __forceinline__ uint32_t f32_to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t n) {
  return (1u << 4)               // D format: f32
         | (2u << 7)             // A format: tf32
         | (2u << 10)            // B format: tf32
         | ((n >> 3) << 17)      // N >> 3
         | ((128u >> 4) << 24);  // M >> 4
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// byte offset of fp32 element (row, col) in a K-major SWIZZLE_128B tile whose K-blocks (32 columns = 128 B rows)
// are kb_stride bytes apart
__device__ __forceinline__ uint32_t sw128_f32(uint32_t row, uint32_t col, uint32_t kb_stride) {
  return (col >> 5) * kb_stride + row * 128u + (((((col & 31u) >> 2) ^ (row & 7u)) << 4) | ((col & 3u) << 2));
}

constexpr int kTcABytes = 4 * 128 * 128;  // one A buffer: 4 K-blocks x 128 rows x 128 B
constexpr int kTcGemmSmem = 3 * kTcABytes + 128 + 1024;
constexpr int kTcGemmThreads = 480;       // warps 0-5 stage A tiles, warps 6-13 run the epilogue, warp 14 issues the MMAs
constexpr int kTcLoadGroup = 96;          // threads per loader group (15 warps = at most 4 per SM sub-partition: 128 registers each)

// Warp-specialised: the three roles only meet at mbarriers (a_full / a_empty per A buffer, acc_full / acc_empty per
// accumulator), so the global loads of tile j+1, the MMAs of tile j and the stores of tile j-1 are all in flight.
// Loaders and epilogue warps each work as TWO groups of four warps that take alternate tiles (group g owns A buffer g /
// accumulator g): a group is latency-bound on its own loads (issue 16 loads, wait, store), so two groups keep two
// tiles' worth of requests in flight and one group's shared-memory / global stores overlap the other's load latency.
__global__ void __launch_bounds__(kTcGemmThreads, 1) tt_gemm_tc_kernel(const __grid_constant__ TtGemm p) {
  extern __shared__ uint8_t tc_raw[];
  uint8_t* base = tc_raw + ((1024u - (smem_u32(tc_raw) & 1023u)) & 1023u);
  uint8_t* Bt = base + 2 * kTcABytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(base + 3 * kTcABytes);
  uint64_t* a_empty = a_full + 2;
  uint64_t* acc_full = a_full + 4;
  uint64_t* acc_empty = a_full + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NO = p.NO, KR = p.KR;
  const int NOp = (NO + 15) & ~15, KRp = (KR + 7) & ~7, nkb = (KRp + 31) >> 5, kc = KR >> 2;
  const uint32_t b_kb = static_cast<uint32_t>(NOp) * 128u;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], kTcLoadGroup);
      mbar_init(&a_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 128);
    }
    mbar_fence_init();
  }
  if (warp == 14) tmem_alloc(tmem_slot, 256);
  for (int i = tid; i < 3 * kTcABytes / 16; i += kTcGemmThreads) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (p.trans_w) {  // B[no][kr] = W[no][kr]
    for (int i = tid; i < NO * kc; i += kTcGemmThreads) {
      const int no = i / kc, c4 = i - no * kc;
      const float4 v = __ldg(reinterpret_cast<const float4*>(p.w + static_cast<size_t>(no) * KR + c4 * 4));
      *reinterpret_cast<uint4*>(Bt + sw128_f32(no, c4 * 4, b_kb)) =
          make_uint4(f32_to_tf32(v.x), f32_to_tf32(v.y), f32_to_tf32(v.z), f32_to_tf32(v.w));
    }
  } else {  // B[no][kr] = W[kr][no]
    const int nc = NO >> 2;
    for (int i = tid; i < KR * nc; i += kTcGemmThreads) {
      const int kr = i / nc, no = (i - kr * nc) * 4;
      const float4 v = __ldg(reinterpret_cast<const float4*>(p.w + static_cast<size_t>(kr) * NO + no));
      *reinterpret_cast<uint32_t*>(Bt + sw128_f32(no + 0, kr, b_kb)) = f32_to_tf32(v.x);
      *reinterpret_cast<uint32_t*>(Bt + sw128_f32(no + 1, kr, b_kb)) = f32_to_tf32(v.y);
      *reinterpret_cast<uint32_t*>(Bt + sw128_f32(no + 2, kr, b_kb)) = f32_to_tf32(v.z);
      *reinterpret_cast<uint32_t*>(Bt + sw128_f32(no + 3, kr, b_kb)) = f32_to_tf32(v.w);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int ntiles = (p.M + kTtBM - 1) / kTtBM;
  const int n_my = (ntiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp < 6) {
    // ---- loaders: two groups of 96 threads, sixteen 16-byte loads in flight per thread before the first store
    const int grp = warp / 3, ltid = tid - grp * kTcLoadGroup;
    const int dr = kTcLoadGroup / kc, dc = kTcLoadGroup - dr * kc;
    const int total = kTtBM * kc;
    for (int j = grp; j < n_my; j += 2) {
      const int s = grp, tile = blockIdx.x + j * gridDim.x;
      uint8_t* A = base + s * kTcABytes;
      if (j >= 2) mbar_wait(&a_empty[s], ((j >> 1) & 1) ^ 1);  // the MMAs of tile j-2 are done reading this buffer
      int r = ltid / kc, c4 = ltid - r * kc;
      for (int i0 = ltid; i0 < total; i0 += 16 * kTcLoadGroup) {
        float4 v[16];
        uint32_t off[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int m = tile * kTtBM + r;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i0 + u * kTcLoadGroup < total && m < p.M)
            v[u] = __ldg(reinterpret_cast<const float4*>(p.a + static_cast<size_t>(m) * KR + c4 * 4));
          off[u] = sw128_f32(r, c4 * 4, 128u * 128u);
          r += dr;
          c4 += dc;
          if (c4 >= kc) {
            c4 -= kc;
            ++r;
          }
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          if (i0 + u * kTcLoadGroup >= total) break;
          float4 t = v[u];
          if (p.relu_in) {
            t.x = fmaxf(t.x, 0.f);
            t.y = fmaxf(t.y, 0.f);
            t.z = fmaxf(t.z, 0.f);
            t.w = fmaxf(t.w, 0.f);
          }
          *reinterpret_cast<uint4*>(A + off[u]) = make_uint4(f32_to_tf32(t.x), f32_to_tf32(t.y), f32_to_tf32(t.z), f32_to_tf32(t.w));
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&a_full[s]);
    }
  } else if (warp == 14) {
    // ---- issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(static_cast<uint32_t>(NOp));
      const uint32_t b0 = smem_u32(Bt);
      for (int j = 0; j < n_my; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&a_full[s], ph);
        if (j >= 2) mbar_wait(&acc_empty[s], ph ^ 1);  // the epilogue of tile j-2 has drained this accumulator
        tc_fence_after();
        const uint32_t a0 = smem_u32(base + s * kTcABytes);
        uint32_t acc = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          const int ks = min(4, (KRp - kb * 32) >> 3);
          for (int k = 0; k < ks; ++k) {
            umma_tf32(tmem + s * 128, make_sw128_desc(a0 + kb * (128 * 128) + k * 32), make_sw128_desc(b0 + kb * b_kb + k * 32),
                      idesc, acc);
            acc = 1;
          }
        }
        umma_commit(&a_empty[s]);
        umma_commit(&acc_full[s]);
      }
    }
  } else {
    // ---- epilogue: two groups of four warps; warp q of a group owns TMEM lanes 32 q .. 32 q + 31, a thread one output row
    const int ew = warp - 6;   // 0..7
    const int q = warp & 3, grp = ew >> 2;
    const int nch = NOp >> 4;
    for (int j = grp; j < n_my; j += 2) {
      const int s = grp, tile = blockIdx.x + j * gridDim.x;
      mbar_wait(&acc_full[s], (j >> 1) & 1);
      tc_fence_after();
      const int m = tile * kTtBM + 32 * q + lane;
      for (int ch = 0; ch < nch; ch += 2) {   // 32 columns per step: 16 mask / residual reads in flight per thread
        const bool two = ch + 1 < nch;
        uint32_t r[32];
        {
          uint32_t r0[16], r1[16];
          tmem_ld16(tmem + (static_cast<uint32_t>(32 * q) << 16) + s * 128 + ch * 16, r0);
          if (two) tmem_ld16(tmem + (static_cast<uint32_t>(32 * q) << 16) + s * 128 + ch * 16 + 16, r1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            r[i] = r0[i];
            r[16 + i] = two ? r1[i] : 0u;
          }
        }
        if (m < p.M) {
          const size_t o0 = static_cast<size_t>(m) * NO + ch * 16;
          float4 sg[8], rr[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const bool in = ch * 16 + jj * 4 < NO;
            sg[jj] = (p.mask_src && in) ? __ldg(reinterpret_cast<const float4*>(p.mask_src + o0 + jj * 4)) : make_float4(1.f, 1.f, 1.f, 1.f);
            rr[jj] = (p.residual && in) ? __ldg(reinterpret_cast<const float4*>(p.residual + o0 + jj * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int no = ch * 16 + jj * 4;
            if (no >= NO) break;
            float4 v = make_float4(__uint_as_float(r[4 * jj]), __uint_as_float(r[4 * jj + 1]),
                                   __uint_as_float(r[4 * jj + 2]), __uint_as_float(r[4 * jj + 3]));
            if (p.bias) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + no));
              v.x += bb.x, v.y += bb.y, v.z += bb.z, v.w += bb.w;
            }
            v.x = (sg[jj].x > 0.f ? v.x : 0.f) + rr[jj].x;
            v.y = (sg[jj].y > 0.f ? v.y : 0.f) + rr[jj].y;
            v.z = (sg[jj].z > 0.f ? v.z : 0.f) + rr[jj].z;
            v.w = (sg[jj].w > 0.f ? v.w : 0.f) + rr[jj].w;
            *reinterpret_cast<float4*>(p.c + o0 + jj * 4) = v;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[s]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 14) tmem_dealloc(tmem, 256);
}

// weight gradient: D[n][k] (n on the 128 TMEM lanes) += sum_m gY[m][n] . act(X)[m][k] over 64-sample tiles; an extra
// all-ones column k = K of the B operand makes column K of D the bias gradient.  Both operands are reduced over the
// SAMPLE index, i.e. they are needed "MN-major" (the M / N index contiguous in memory) -- exactly how the row-major
// gY [M][N] and X [M][K] tiles lie in global memory, so the loaders copy coalesced 16-byte chunks (no transposition)
// into the canonical MN-major SWIZZLE_128B layout (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>):
//   atom = 4 samples x 128 B (32 consecutive n), 32-byte unit u of sample row r at unit u ^ (r & 3);
//   atoms of the next 4 samples follow at SBO = 512 B, the next 32 n at LBO = 8 KB (64 samples per tile);
// one kind::tf32 MMA consumes K = 8 samples = two atoms, so the descriptor start address advances by 1 KB.
constexpr int kTcWgTile = 64;                            // samples per tile
constexpr int kTcWgSlab = (kTcWgTile / 4) * 512;         // one 32-column slab of a tile: 8 KB (= LBO)
constexpr int kTcWgA = 4 * kTcWgSlab;                    // gY tile: 128 n = 4 slabs
constexpr int kTcWgB = 5 * kTcWgSlab;                    // act(X) tile + ones column: K + 16 <= 144 -> 5 slabs
constexpr int kTcWgStages = 3;
constexpr int kTcWgradSmem = kTcWgStages * (kTcWgA + kTcWgB) + 128 + 1024;
constexpr int kTcWgradThreads = 288;       // warps 0-7 stage the operand tiles, warp 8 issues the MMAs

// byte offset of fp32 element (sample r < 64, column c) inside an MN-major tile.  For 32-bit MN-major operands the
// only layout the tensor core accepts is SWIZZLE_128B_BASE32B (cutlass sm100_common.inl: "for mn-major tf32 operands,
// SW128_32B is the only available smem layout"): atoms of 4 samples x 128 B whose 32-byte units are XOR-swizzled with
// the sample index (Swizzle<2,5,2> on the byte address: bits [5,7) ^= bits [7,9)).
__device__ __forceinline__ uint32_t mn128_f32(uint32_t r, uint32_t c) {
  return (c >> 5) * kTcWgSlab + (r >> 2) * 512u + (r & 3u) * 128u + (((((c & 31u) >> 3) ^ (r & 3u)) << 5) | ((c & 7u) << 2));
}
__device__ __forceinline__ uint64_t make_sw128_mn_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);        // start address
  d |= static_cast<uint64_t>(kTcWgSlab >> 4) << 16;           // LBO: next 32 columns of the M / N dimension
  d |= static_cast<uint64_t>(512 >> 4) << 32;                 // SBO: next 4 samples of the reduction dimension
  d |= static_cast<uint64_t>(1) << 46;                        // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(1) << 61;                        // layout: SWIZZLE_128B_BASE32B
  return d;
}

__global__ void __launch_bounds__(kTcWgradThreads, 1) tt_wgrad_tc_kernel(const __grid_constant__ TtWgrad p) {
  extern __shared__ uint8_t tc_raw[];
  uint8_t* base = tc_raw + ((1024u - (smem_u32(tc_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + kTcWgStages * (kTcWgA + kTcWgB));
  uint64_t* empty = full + kTcWgStages;
  uint64_t* done = full + 2 * kTcWgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 2 * kTcWgStages + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, K = p.K;
  const int Kp = (K + 1 + 15) & ~15;   // MMA N: the K weight columns + the ones column, rounded up to 16
  if (tid == 0) {
    for (int i = 0; i < kTcWgStages; ++i) {
      mbar_init(&full[i], 256);
      mbar_init(&empty[i], 1);
    }
    mbar_init(done, 1);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 256);
  for (int i = tid; i < kTcWgStages * (kTcWgA + kTcWgB) / 16; i += kTcWgradThreads)
    reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);   // columns past N / K stay zero
  __syncthreads();
  if (tid < kTcWgStages * kTcWgTile) {  // the ones column of every B buffer (never overwritten: the X chunks end at K)
    const int buf = tid / kTcWgTile, r = tid % kTcWgTile;
    *reinterpret_cast<uint32_t*>(base + buf * (kTcWgA + kTcWgB) + kTcWgA + mn128_f32(r, K)) = 0x3f800000u;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int m_begin = blockIdx.x * p.rows_per_cta, m_end = min(p.M, m_begin + p.rows_per_cta);
  const int n_my = m_end > m_begin ? (m_end - m_begin + kTcWgTile - 1) / kTcWgTile : 0;

  if (warp < 8) {
    // loaders: warp w copies sample rows 8 w .. 8 w + 7 of the tile, lane = 16-byte chunk of the row (coalesced)
    const bool has_y = 4 * lane < N, has_x = 4 * lane < K;
    for (int j = 0; j < n_my; ++j) {
      const int s = j % kTcWgStages;
      const uint32_t use = static_cast<uint32_t>(j / kTcWgStages);
      uint8_t* A = base + s * (kTcWgA + kTcWgB);
      uint8_t* Bm = A + kTcWgA;
      float4 vy[8], vx[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = m_begin + j * kTcWgTile + warp * 8 + i;
        vy[i] = (has_y && m < m_end) ? __ldg(reinterpret_cast<const float4*>(p.gy + static_cast<size_t>(m) * N + 4 * lane))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
        vx[i] = (has_x && m < m_end) ? __ldg(reinterpret_cast<const float4*>(p.x + static_cast<size_t>(m) * K + 4 * lane))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (use > 0) mbar_wait(&empty[s], (use & 1) ^ 1);  // the MMAs of tile j - stages are done with this buffer
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t r = warp * 8 + i;
        if (has_y)
          *reinterpret_cast<uint4*>(A + mn128_f32(r, 4 * lane)) =
              make_uint4(f32_to_tf32(vy[i].x), f32_to_tf32(vy[i].y), f32_to_tf32(vy[i].z), f32_to_tf32(vy[i].w));
        if (has_x) {
          float4 v = vx[i];
          if (p.relu_in) {
            v.x = fmaxf(v.x, 0.f);
            v.y = fmaxf(v.y, 0.f);
            v.z = fmaxf(v.z, 0.f);
            v.w = fmaxf(v.w, 0.f);
          }
          *reinterpret_cast<uint4*>(Bm + mn128_f32(r, 4 * lane)) =
              make_uint4(f32_to_tf32(v.x), f32_to_tf32(v.y), f32_to_tf32(v.z), f32_to_tf32(v.w));
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&full[s]);
    }
    if (n_my > 0) {
      mbar_wait(done, 0);
      tc_fence_after();
      const int q = warp & 3, h = warp >> 2;
      const int n = 32 * q + lane;
      const int nch = Kp >> 4, ch_begin = h ? (nch + 1) / 2 : 0, ch_end = h ? nch : (nch + 1) / 2;
      for (int ch = ch_begin; ch < ch_end; ++ch) {
        uint32_t r[16];
        tmem_ld16(tmem + (static_cast<uint32_t>(32 * q) << 16) + ch * 16, r);
        tmem_ld_wait();
        if (n < N) {
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const int k = ch * 16 + jj;
            if (k < K)
              atomicAdd(p.gw + static_cast<size_t>(n) * K + k, __uint_as_float(r[jj]));
            else if (k == K && p.gb)
              atomicAdd(p.gb + n, __uint_as_float(r[jj]));
          }
        }
      }
    }
  } else if (lane == 0) {
    const uint32_t idesc = make_idesc_tf32(static_cast<uint32_t>(Kp)) | (1u << 15) | (1u << 16);   // A and B MN-major
    for (int j = 0; j < n_my; ++j) {
      const int s = j % kTcWgStages;
      mbar_wait(&full[s], (j / kTcWgStages) & 1);
      tc_fence_after();
      const uint32_t a0 = smem_u32(base + s * (kTcWgA + kTcWgB)), b0 = a0 + kTcWgA;
      for (int kg = 0; kg < kTcWgTile / 8; ++kg)
        umma_tf32(tmem, make_sw128_mn_desc(a0 + kg * 1024), make_sw128_mn_desc(b0 + kg * 1024), idesc, (j | kg) ? 1u : 0u);
      umma_commit(&empty[s]);
      if (j == n_my - 1) umma_commit(done);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 256);
}

// ----------------------------------------------------------------------------- SH-16 of the view direction
__global__ void __launch_bounds__(256) tt_sh16_kernel(const float* __restrict__ dirs, int M, int conv, int fp16_round,
                                                      float* __restrict__ out) {
  const int m = blockIdx.x * 256 + threadIdx.x;
  if (m >= M) return;
  float sh[16];
  sh16(__ldg(dirs + 3 * static_cast<size_t>(m)), __ldg(dirs + 3 * static_cast<size_t>(m) + 1),
       __ldg(dirs + 3 * static_cast<size_t>(m) + 2), conv, sh);
  if (fp16_round) {
#pragma unroll
    for (int i = 0; i < 16; ++i) sh[i] = __half2float(__float2half_rn(sh[i]));  // tiny-cuda-nn returns fp16
  }
  float4* o = reinterpret_cast<float4*>(out + static_cast<size_t>(m) * 16);
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] = make_float4(sh[4 * i], sh[4 * i + 1], sh[4 * i + 2], sh[4 * i + 3]);
}

template <class K>
int tt_set_smem(K kernel, int bytes) {
  NJF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return 0;
}

bool tt_dim_ok(int d) { return d >= 4 && d <= 128 && d % 4 == 0; }

}  // namespace
}  // namespace njf

using namespace njf;

extern "C" {

int njf_train_sample_setup(const float* ctxt_w2c, const float* ctxt_k, const float* points, int B, int N, int Hf, int Wf,
                           float* enc, int* tap_pix, float* tap_w, void* stream) {
  if (!ctxt_w2c || !ctxt_k || !points || !enc || !tap_pix || !tap_w) NJF_FAIL("njf_train_sample_setup: null argument");
  if (B <= 0 || N <= 0 || Hf < 1 || Wf < 1) NJF_FAIL("njf_train_sample_setup: bad sizes");
  if (static_cast<long long>(B) * N > 0x7fffffffLL) NJF_FAIL("njf_train_sample_setup: too many points");
  TtSetup p{};
  p.g.NR = B * N;
  p.g.ray0 = 0;
  p.g.R = N;
  p.g.S = 1;
  p.g.G = 128;
  p.g.T = 1;
  p.g.NG = (B * N + 127) / 128;
  p.g.points = points;
  p.g.ctxt_w2c = ctxt_w2c;
  p.g.ctxt_k = ctxt_k;
  p.g.Hf = Hf;
  p.g.Wf = Wf;
  p.g.n_const_views = 0;
  p.enc = enc;
  p.tap_pix = tap_pix;
  p.tap_w = tap_w;
  tt_setup_kernel<<<p.g.NG, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
  NJF_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int njf_train_gather(const float* map, const int* tap_pix, const float* tap_w, int M, int CH, int ch0, int CW, float* out,
                     void* stream) {
  if (!map || !tap_pix || !tap_w || !out) NJF_FAIL("njf_train_gather: null argument");
  if (M <= 0 || CH <= 0 || CH % 4 || ch0 < 0 || ch0 % 4 || CW <= 0 || CW % 128 || ch0 + CW > CH)
    NJF_FAIL("njf_train_gather: channel window [%d, %d) of %d (CW must be a positive multiple of 128)", ch0, ch0 + CW, CH);
  tt_gather_kernel<<<(M + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(map, tap_pix, tap_w, M, CH, ch0, CW, out);
  NJF_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int njf_train_scatter(const float* g, const int* tap_pix, const float* tap_w, int M, int CH, int ch0, int CW, float* dmap,
                      void* stream) {
  if (!g || !tap_pix || !tap_w || !dmap) NJF_FAIL("njf_train_scatter: null argument");
  if (M <= 0 || CH <= 0 || CH % 4 || ch0 < 0 || ch0 % 4 || CW <= 0 || CW % 128 || ch0 + CW > CH)
    NJF_FAIL("njf_train_scatter: channel window [%d, %d) of %d (CW must be a positive multiple of 128)", ch0, ch0 + CW, CH);
  tt_scatter_kernel<<<(M + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(g, tap_pix, tap_w, M, CH, ch0, CW, dmap);
  NJF_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int njf_train_linear(const float* a, const float* w, const float* bias, const float* residual, const float* mask_src,
                     float* c, int M, int n_out, int k_red, int trans_w, int relu_in, int tensor_cores, void* stream) {
  if (!a || !w || !c) NJF_FAIL("njf_train_linear: null argument");
  if (M <= 0 || !tt_dim_ok(n_out) || !tt_dim_ok(k_red))
    NJF_FAIL("njf_train_linear: M=%d n_out=%d k_red=%d (inner sizes must be multiples of 4 in [4,128])", M, n_out, k_red);
  TtGemm p{a, w, bias, residual, mask_src, c, M, n_out, k_red, trans_w, relu_in};
  const int grid = std::min((M + kTtBM - 1) / kTtBM, tt_sms());
  if (tensor_cores) {
    if (tt_set_smem(tt_gemm_tc_kernel, kTcGemmSmem)) return 1;
    tt_gemm_tc_kernel<<<grid, kTcGemmThreads, kTcGemmSmem, static_cast<cudaStream_t>(stream)>>>(p);
    NJF_CUDA(cudaGetLastError());
    count_launch();
    return 0;
  }
  const int smem = 2 * 128 * kTtLd * 4;
  if (n_out > 64) {
    if (tt_set_smem(tt_gemm_kernel<2>, smem)) return 1;
    tt_gemm_kernel<2><<<grid, kTtThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
  } else {
    if (tt_set_smem(tt_gemm_kernel<1>, smem)) return 1;
    tt_gemm_kernel<1><<<grid, kTtThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
  }
  NJF_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int njf_train_linear_wgrad(const float* gy, const float* x, int M, int N, int K, int relu_in, float* gw, float* gb,
                           int tensor_cores, void* stream) {
  if (!gy || !x || !gw) NJF_FAIL("njf_train_linear_wgrad: null argument");
  if (M <= 0 || !tt_dim_ok(N) || !tt_dim_ok(K))
    NJF_FAIL("njf_train_linear_wgrad: M=%d N=%d K=%d (inner sizes must be multiples of 4 in [4,128])", M, N, K);
  const int chunks = (M + 63) / 64;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (tensor_cores) {
    const int grid_tc = std::min(chunks, tt_sms());
    TtWgrad ptc{gy, x, gw, gb, M, N, K, relu_in, ((chunks + grid_tc - 1) / grid_tc) * 64};
    if (tt_set_smem(tt_wgrad_tc_kernel, kTcWgradSmem)) return 1;
    tt_wgrad_tc_kernel<<<grid_tc, kTcWgradThreads, kTcWgradSmem, st>>>(ptc);
    NJF_CUDA(cudaGetLastError());
    count_launch();
    return 0;
  }
  const int grid = std::min(chunks, 2 * tt_sms());
  TtWgrad p{gy, x, gw, gb, M, N, K, relu_in, ((chunks + grid - 1) / grid) * 64};
  const int smem = 2 * 64 * 128 * 4;
  const bool n2 = N > 64, k2 = K > 64;
#define NJF_TT_WGRAD(A_, B_)                                              \
  do {                                                                    \
    if (tt_set_smem(tt_wgrad_kernel<A_, B_>, smem)) return 1;             \
    tt_wgrad_kernel<A_, B_><<<grid, kTtThreads, smem, st>>>(p);           \
  } while (0)
  if (n2 && k2) NJF_TT_WGRAD(true, true);
  else if (n2) NJF_TT_WGRAD(true, false);
  else if (k2) NJF_TT_WGRAD(false, true);
  else NJF_TT_WGRAD(false, false);
#undef NJF_TT_WGRAD
  NJF_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int njf_train_sh16(const float* dirs, int M, int sh_convention, int fp16_round, float* out, void* stream) {
  if (!dirs || !out || M <= 0) NJF_FAIL("njf_train_sh16: bad argument");
  if (sh_convention != NJF_SH_TCNN && sh_convention != NJF_SH_NERFSTUDIO_TORCH) NJF_FAIL("njf_train_sh16: unknown SH convention");
  tt_sh16_kernel<<<(M + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(dirs, M, sh_convention, fp16_round, out);
  NJF_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // extern "C"
